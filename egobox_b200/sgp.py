"""Host-side mirror of the sparse GP surface of egobox-gp (crates/gp/src/sparse_algorithm.rs,
sparse_parameters.rs) and of the `SparseGpx` / `SparseGpMix` Python classes
(python/src/sparse_gp_mix.rs:64-459).  All arithmetic runs in libegobox_gpu.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import EGX_OK, SgpParamsStruct, NUM_STAGES, STAGE_NAMES
from .gp import (_raise_status, InvalidValueError, SquaredExponentialCorr, CORR_NAMES, GP_COBYLA_MAX_EVAL,
                 GP_COBYLA_MIN_EVAL, GP_OPTIM_N_START)
from .context import DEFAULT_NUGGET

FITC, VFE = 0, 1
_dp = C.POINTER(C.c_double)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class ParamTuning:
    """ParamTuning{Fixed, Optimized}, sparse_parameters.rs:14-32."""

    def __init__(self, fixed, init, bounds=None):
        self.fixed, self.init, self.bounds = fixed, init, bounds

    @classmethod
    def Fixed(cls, value):
        return cls(True, value)

    @classmethod
    def Optimized(cls, init=1e-2, bounds=(100.0 * np.finfo(np.float64).eps, 1e10)):
        return cls(False, init, bounds)


class Inducings:
    """Inducings{Randomized(n), Located(z)}, sparse_parameters.rs:38-48."""

    def __init__(self, n=None, z=None):
        self.n, self.z = n, z

    @classmethod
    def Randomized(cls, n):
        return cls(n=n)

    @classmethod
    def Located(cls, z):
        return cls(z=_f64(z))


class SgpContext:
    """egx_sgp_ctx: likelihood / finalize / predict for given hyper-parameters (device seam)."""

    def __init__(self, x, y, z, corr=SquaredExponentialCorr, method=FITC, w_star=None, nugget=DEFAULT_NUGGET, device=0):
        self._lib = _lib.load()
        x, y, z = _f64(x), _f64(y).reshape(-1), _f64(z)
        if x.ndim == 1:
            x = x[:, None]
        n, d = x.shape
        w = np.eye(d) if w_star is None else _f64(w_star)
        self.n, self.d, self.m, self.h = n, d, z.shape[0], w.shape[1]
        self._h = C.c_void_p()
        st = self._lib.egx_sgp_create(C.byref(self._h), device, corr, method, x.ctypes.data_as(_dp), n, d,
                                      y.ctypes.data_as(_dp), z.ctypes.data_as(_dp), z.shape[0],
                                      _f64(w).ctypes.data_as(_dp), w.shape[1], float(nugget))
        if st != EGX_OK:
            _raise_status(st)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.egx_sgp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reduced_likelihood(self, theta, sigma2, noise):
        th = _f64(theta).reshape(-1)
        out = C.c_double()
        st = self._lib.egx_sgp_reduced_likelihood(self._h, th.ctypes.data_as(_dp), float(sigma2), float(noise),
                                                  C.byref(out))
        if st in (4, 5):
            _raise_status(st)
        return st, out.value

    def finalize(self, theta, sigma2, noise, want_inv=False):
        th = _f64(theta).reshape(-1)
        lik = C.c_double()
        vec = np.empty(self.m)
        inv = np.empty((self.m, self.m)) if want_inv else None
        st = self._lib.egx_sgp_finalize(self._h, th.ctypes.data_as(_dp), float(sigma2), float(noise), C.byref(lik),
                                        vec.ctypes.data_as(_dp), inv.ctypes.data_as(_dp) if want_inv else None)
        if st in (4, 5):
            _raise_status(st)
        return st, dict(likelihood=lik.value, w_vec=vec, w_inv=inv)

    def predict(self, x):
        x = _f64(x).reshape(-1, self.d)
        y = np.empty(x.shape[0])
        st = self._lib.egx_sgp_predict(self._h, x.ctypes.data_as(_dp), x.shape[0], y.ctypes.data_as(_dp))
        if st != EGX_OK:
            _raise_status(st)
        return y

    def predict_var(self, x):
        x = _f64(x).reshape(-1, self.d)
        v = np.empty(x.shape[0])
        st = self._lib.egx_sgp_predict_var(self._h, x.ctypes.data_as(_dp), x.shape[0], v.ctypes.data_as(_dp))
        if st != EGX_OK:
            _raise_status(st)
        return v

    def sample(self, x, z, method=1):
        """egx_sgp_sample: (m, n_traj) trajectories predict(x) + C z, C C^T = sigma2 r(x, x); z: (m, n_traj) normal draws,
        method 0 = Cholesky, 1 = eigen-decomposition (sparse_algorithm.rs:338-364)."""
        x = _f64(x).reshape(-1, self.d)
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(x.shape[0], -1)
        out = np.empty_like(z)
        st = self._lib.egx_sgp_sample(self._h, x.ctypes.data_as(_dp), x.shape[0], z.ctypes.data_as(_dp), z.shape[1], int(method),
                                      out.ctypes.data_as(_dp))
        if st != EGX_OK:
            _raise_status(st)
        return out

    def set_profiling(self, on=True):
        self._lib.egx_sgp_set_profiling(self._h, int(bool(on)))

    def profile(self):
        ms = np.zeros(NUM_STAGES)
        ln = np.zeros(NUM_STAGES, dtype=np.int64)
        self._lib.egx_sgp_get_profile(self._h, ms.ctypes.data_as(_dp), ln.ctypes.data_as(C.POINTER(C.c_longlong)))
        return {STAGE_NAMES[i]: (float(ms[i]), int(ln[i])) for i in range(NUM_STAGES)}


class SgpParams:
    """SgpParams builder (sparse_parameters.rs:151-293)."""

    def __init__(self, corr=SquaredExponentialCorr, inducings=None):
        self._corr = corr
        self._inducings = inducings or Inducings.Randomized(10)
        self._theta_init, self._theta_bounds, self._theta_fixed = None, None, False
        self._noise = ParamTuning.Optimized()
        self._method = FITC
        self._n_start, self._max_eval = GP_OPTIM_N_START, GP_COBYLA_MAX_EVAL
        self._nugget, self._seed, self._device = DEFAULT_NUGGET, None, 0
        self._kpls_dim, self._w_star = None, None
        self._ftol_rel = 1e-4

    def corr(self, corr):
        self._corr = corr
        return self

    def theta_init(self, init):
        self._theta_init = list(np.atleast_1d(init))
        return self

    def theta_bounds(self, bounds):
        self._theta_bounds = bounds
        return self

    def theta_fixed(self, init):
        self._theta_init, self._theta_fixed = list(np.atleast_1d(init)), True
        return self

    def noise_variance(self, tuning):
        self._noise = tuning
        return self

    def sparse_method(self, method):
        self._method = method
        return self

    def inducings(self, inducings):
        self._inducings = inducings
        return self

    def n_start(self, n):
        self._n_start = n
        return self

    def max_eval(self, n):
        self._max_eval = max(GP_COBYLA_MIN_EVAL, n)
        return self

    def nugget(self, v):
        self._nugget = v
        return self

    def seed(self, seed):
        self._seed = seed
        return self

    def device(self, device):
        self._device = device
        return self

    def kpls_dim(self, k, w_star=None):
        self._kpls_dim, self._w_star = k, w_star
        return self

    def fit(self, x, y):
        """impl Fit for SgpValidParams, sparse_algorithm.rs:416-648."""
        lib = _lib.load()
        x = _f64(x)
        if x.ndim == 1:
            x = x[:, None]
        y = _f64(y).reshape(-1)
        n, d = x.shape
        if y.shape[0] != n:
            raise InvalidValueError("x and y should have the same number of rows")
        prm = SgpParamsStruct()
        lib.egx_sgp_params_default(C.byref(prm))
        keep = []
        prm.corr, prm.method = int(self._corr), int(self._method)
        prm.theta_fixed = int(self._theta_fixed)
        if self._theta_init is not None:
            a = _f64(self._theta_init)
            prm.theta_init, prm.n_theta_init = a.ctypes.data_as(_dp), a.size
            keep.append(a)
        if self._theta_bounds is not None:
            b = _f64(self._theta_bounds).reshape(-1, 2)
            prm.theta_bounds, prm.n_theta_bounds = b.ctypes.data_as(_dp), b.shape[0]
            keep.append(b)
        prm.noise_fixed = int(self._noise.fixed)
        prm.noise_init = float(self._noise.init)
        if self._noise.bounds is not None:
            prm.noise_lo, prm.noise_hi = float(self._noise.bounds[0]), float(self._noise.bounds[1])
        if self._inducings.z is not None:
            z = _f64(self._inducings.z)
            if z.shape[1] != d:
                raise InvalidValueError("inducing points should have %d columns" % d)
            prm.z, prm.n_inducings = z.ctypes.data_as(_dp), z.shape[0]
            keep.append(z)
        else:
            prm.n_inducings = int(self._inducings.n)
        prm.n_start, prm.max_eval, prm.nugget = int(self._n_start), int(self._max_eval), float(self._nugget)
        if self._kpls_dim is not None:
            if self._kpls_dim > d:
                raise InvalidValueError("Dimension reduction %d should be smaller than actual training input "
                                        "dimensions %d" % (self._kpls_dim, d))
            if self._w_star is None:      # rotations computed by the fit driver (sparse_algorithm.rs:442-455)
                prm.kpls_dim = int(self._kpls_dim)
            else:
                w = _f64(self._w_star)
                prm.w_star, prm.kpls_dim = w.ctypes.data_as(_dp), w.shape[1]
                keep.append(w)
        prm.device = int(self._device)
        prm.seed = int(self._seed if self._seed is not None else np.random.SeedSequence().entropy % (2 ** 63))
        prm.cobyla_ftol_rel = float(self._ftol_rel)
        h = C.c_void_p()
        st = lib.egx_sgp_fit(C.byref(prm), x.ctypes.data_as(_dp), n, d, y.ctypes.data_as(_dp), C.byref(h))
        if st != EGX_OK:
            _raise_status(st)
        return SparseGaussianProcess(h, self, (x.copy(), y.copy()))


class SparseGaussianProcess:
    """Trained sparse GP (sparse_algorithm.rs:145-169)."""

    def __init__(self, handle, params, training_data):
        self._lib = _lib.load()
        self._h = handle
        self.params_ = params
        self.training_data = training_data
        n, d, h, m = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._lib.egx_sgp_model_dims(self._h, C.byref(n), C.byref(d), C.byref(h), C.byref(m))
        self._n, self._d, self._hdim, self._m = n.value, d.value, h.value, m.value

    @staticmethod
    def params(corr=SquaredExponentialCorr, inducings=None):
        return SgpParams(corr, inducings)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.egx_sgp_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _x(self, x):
        x = _f64(x)
        if x.ndim == 1:
            x = x.reshape(-1, self._d)
        return x

    def predict(self, x):
        x = self._x(x)
        y = np.empty(x.shape[0])
        st = self._lib.egx_sgp_model_predict(self._h, x.ctypes.data_as(_dp), x.shape[0], y.ctypes.data_as(_dp))
        if st != EGX_OK:
            _raise_status(st)
        return y

    def predict_var(self, x):
        x = self._x(x)
        v = np.empty(x.shape[0])
        st = self._lib.egx_sgp_model_predict_var(self._h, x.ctypes.data_as(_dp), x.shape[0], v.ctypes.data_as(_dp))
        if st != EGX_OK:
            _raise_status(st)
        return v

    def _sample(self, x, n_traj, method, seed=None, z=None):
        x = self._x(x)
        if z is None:
            # the reference draws from an unseeded generator (gp/src/algorithm.rs:1191-1192)
            z = np.random.default_rng(seed).standard_normal((x.shape[0], int(n_traj)))
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(x.shape[0], -1)
        out = np.empty_like(z)
        st = self._lib.egx_sgp_model_sample(self._h, x.ctypes.data_as(_dp), x.shape[0], z.ctypes.data_as(_dp), z.shape[1],
                                            int(method), out.ctypes.data_as(_dp))
        if st != EGX_OK:
            _raise_status(st)
        return out

    def sample_chol(self, x, n_traj, seed=None, z=None):
        """sparse_algorithm.rs:338-341: (n, n_traj) trajectories, Cholesky of sigma2 r(x, x)."""
        return self._sample(x, n_traj, 0, seed, z)

    def sample_eig(self, x, n_traj, seed=None, z=None):
        """sparse_algorithm.rs:343-346: eigen-decomposition (eigenvalues < 1e-9 dropped)."""
        return self._sample(x, n_traj, 1, seed, z)

    def sample(self, x, n_traj, seed=None, z=None):
        """sparse_algorithm.rs:348-351: alias of sample_eig."""
        return self.sample_eig(x, n_traj, seed, z)

    # sparse_algorithm.rs:298-336: the reference differentiates predict / predict_var by central differences with the
    # fixed step sqrt(eps) of the `finitediff` crate, one point and one coordinate at a time; here all 2 * n * nx shifted
    # points go through ONE batched device prediction
    _FD_STEP = 1.4901161193847656e-08

    def _central_diff(self, fn, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x.reshape(-1, self._d)
        n, nx = x.shape
        h = self._FD_STEP
        shifts = np.zeros((2, nx, nx))
        shifts[0][np.diag_indices(nx)] = h
        shifts[1][np.diag_indices(nx)] = -h
        pts = (x[None, None, :, :] + shifts[:, :, None, :]).reshape(2 * nx * n, nx)
        v = fn(pts).reshape(2, nx, n)
        return ((v[0] - v[1]) / (2.0 * h)).T.copy()

    def predict_gradients(self, x):
        """sparse_algorithm.rs:298-316 -> (n, nx)."""
        return self._central_diff(self.predict, x)

    def predict_var_gradients(self, x):
        """sparse_algorithm.rs:318-336 -> (n, nx)."""
        return self._central_diff(self.predict_var, x)

    def q2_score(self, kfold, fit=None):
        """`PredictScore::q2_score` for the sparse GP, gp/src/metrics.rs:35-53, 77-91."""
        from .gp import _q2_score
        return _q2_score(self.training_data, kfold, fit if fit is not None else self.params_.fit)

    def looq2_score(self, fit=None):
        return self.q2_score(self.training_data[0].shape[0], fit)

    def theta(self):
        th = np.empty(self._hdim)
        self._lib.egx_sgp_model_theta(self._h, th.ctypes.data_as(_dp))
        return th

    def variance(self):
        return float(self._lib.egx_sgp_model_variance(self._h))

    def noise_variance(self):
        return float(self._lib.egx_sgp_model_noise_variance(self._h))

    def likelihood(self):
        return float(self._lib.egx_sgp_model_likelihood(self._h))

    def n_evals(self):
        return int(self._lib.egx_sgp_model_n_evals(self._h))

    def inducings(self):
        z = np.empty((self._m, self._d))
        self._lib.egx_sgp_model_inducings(self._h, z.ctypes.data_as(_dp))
        return z

    def woodbury(self, with_inv=True):
        vec = np.empty((self._m, 1))
        inv = np.empty((self._m, self._m)) if with_inv else None
        st = self._lib.egx_sgp_model_woodbury(self._h, vec.ctypes.data_as(_dp),
                                              inv.ctypes.data_as(_dp) if with_inv else None)
        if st != EGX_OK:
            _raise_status(st)
        return dict(vec=vec, inv=inv)

    def dims(self):
        return (self._d, 1)

    def __str__(self):
        return "SGP(corr=%s, theta=%s, variance=%s, noise variance=%s, likelihood=%s)" % (
            CORR_NAMES[self.params_._corr], self.theta(), self.variance(), self.noise_variance(), self.likelihood())


class SparseKriging:
    """SparseKriging = SgpParams<SquaredExponentialCorr> (sparse_algorithm.rs:171-180)."""

    @staticmethod
    def params(inducings):
        return SgpParams(SquaredExponentialCorr, inducings)


class SparseMethod:
    FITC, VFE = 0, 1
    Fitc, Vfe = 0, 1


def _expert_seed(mixture_seed):
    """moe/src/algorithm.rs:330: the mixture hands its expert `self.rng().gen()` read as Option<u64> (surrogates.rs:42),
    i.e. rand 0.8.5's Standard for Option: one bool (sign bit of next_u32) and, if true, one next_u64 of
    Xoshiro256Plus::seed_from_u64(mixture_seed); None = the expert seeds itself from entropy (sparse_algorithm.rs:457-460)."""
    if mixture_seed is None:
        return None
    m64 = (1 << 64) - 1
    z, st = int(mixture_seed) & m64, []
    for _ in range(4):                                   # SplitMix64 seeding of rand_xoshiro 0.6.0
        z = (z + 0x9E3779B97F4A7C15) & m64
        x = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m64
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & m64
        st.append(x ^ (x >> 31))

    def nxt():
        r, t = (st[0] + st[3]) & m64, (st[1] << 17) & m64
        st[2] ^= st[0]
        st[3] ^= st[1]
        st[1] ^= st[2]
        st[0] ^= st[3]
        st[2] ^= t
        st[3] = ((st[3] << 45) | (st[3] >> 19)) & m64
        return r
    return nxt() if (nxt() >> 63) == 1 else None


class SparseGpMix:
    """python/src/sparse_gp_mix.rs:64-219 (single cluster)."""

    def __init__(self, corr_spec=1, theta_init=None, theta_bounds=None, kpls_dim=None, n_start=10, nz=None, z=None,
                 method=SparseMethod.FITC, seed=None, w_star=None, device=0):
        self.corr_spec, self.theta_init, self.theta_bounds, self.kpls_dim = corr_spec, theta_init, theta_bounds, kpls_dim
        self.n_start, self.nz, self.z, self.method, self.seed = n_start, nz, z, method, seed
        self.w_star, self.device = w_star, device

    def fit(self, xt, yt):
        from .gpx import _CORR
        xt = np.asarray(xt, dtype=np.float64)
        if xt.ndim == 1:
            xt = xt[:, None]
        yt = np.asarray(yt, dtype=np.float64)
        if yt.ndim == 2:
            if yt.shape[1] != 1:
                raise ValueError("Bad training output data")
            yt = yt[:, 0]
        corr_spec, self.cv_errors_ = self.corr_spec, None
        if corr_spec not in _CORR:
            # several correlation models: moe/src/algorithm.rs:209-260 decides by the 5-fold cross-validation error of the
            # DENSE expert with constant mean (`compute_errors!` builds full GPs whatever the `gp_type`), then trains the
            # sparse expert with the winner (:306-327)
            from . import gp as _gp
            from .moe import GpMixtureParams, _CORRS
            allowed = [(nm, bit, c) for nm, bit, c in _CORRS if int(corr_spec) & bit]
            if not allowed:
                raise InvalidValueError("empty correlation specification")
            cv = GpMixtureParams().set(kpls_dim=self.kpls_dim, w_star=self.w_star, device=self.device)
            errs = [(nm, bit, cv._cv_error("Constant", _gp.ConstantMean, c, xt, yt)) for nm, bit, c in allowed]
            self.cv_errors_ = {"Constant_%s" % nm: e for nm, _, e in errs}
            corr_spec = min(errs, key=lambda e: e[2] if not np.isnan(e[2]) else np.inf)[1]
        if self.z is not None:
            ind = Inducings.Located(self.z)
        elif self.nz is not None:
            ind = Inducings.Randomized(self.nz)
        else:
            raise ValueError("You must specify inducing points")      # sparse_gp_mix.rs:176-178
        p = SgpParams(_CORR[corr_spec], ind).sparse_method(self.method).n_start(self.n_start)
        p = p.seed(_expert_seed(self.seed))
        p = p.device(self.device)
        if self.theta_init is not None:
            p = p.theta_init(self.theta_init)
        if self.theta_bounds is not None:
            p = p.theta_bounds(self.theta_bounds)
        if self.kpls_dim is not None:
            p = p.kpls_dim(self.kpls_dim, self.w_star)
        model = SparseGpx(p.fit(xt, yt))
        model.cv_errors_ = self.cv_errors_
        return model


class SparseGpx:
    """python/src/sparse_gp_mix.rs:224-459."""

    def __init__(self, model):
        self._gp = model

    @staticmethod
    def builder(corr_spec=1, theta_init=None, theta_bounds=None, kpls_dim=None, n_start=10, nz=None, z=None,
                method=SparseMethod.FITC, seed=None, **kw):
        return SparseGpMix(corr_spec, theta_init, theta_bounds, kpls_dim, n_start, nz, z, method, seed, **kw)

    def predict(self, x):
        return self._gp.predict(x)

    def predict_var(self, x):
        return self._gp.predict_var(x)

    def thetas(self):
        return self._gp.theta()[None, :]

    def variances(self):
        return np.array([self._gp.variance()])

    def likelihoods(self):
        return np.array([self._gp.likelihood()])

    def gp(self):
        return self._gp

    def predict_gradients(self, x):
        """sparse_gp_mix.rs (SparseGpx.predict_gradients) -> sparse_algorithm.rs:298-316."""
        return self._gp.predict_gradients(np.asarray(x, dtype=np.float64))

    def predict_var_gradients(self, x):
        """sparse_gp_mix.rs (SparseGpx.predict_var_gradients) -> sparse_algorithm.rs:318-336."""
        return self._gp.predict_var_gradients(np.asarray(x, dtype=np.float64))

    def sample(self, x, n_traj, seed=None, z=None):
        """SparseGpx.sample (sparse_gp_mix.rs) -> sparse_algorithm.rs:348-364."""
        return self._gp.sample(np.asarray(x, dtype=np.float64), n_traj, seed=seed, z=z)
