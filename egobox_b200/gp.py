"""Host-side mirror of the egobox-gp builder / fit / predict surface for the kriging path.

Names, argument meaning and error behaviour follow crates/gp (relf/egobox @ be16128):
``GaussianProcess.params(mean, corr)`` -> ``GpParams`` builder (parameters.rs:163-273) ->
``fit(x, y)`` (algorithm.rs:791-979) -> ``GaussianProcess`` with ``predict``,
``predict_var``, ``predict_valvar``, ``theta()``, ``variance()``, ``likelihood()``,
``dims()`` (algorithm.rs:253-307, 413-439).  ``Kriging.params()`` is the constant-mean /
squared-exponential alias (algorithm.rs:200-207).  Everything numerical happens in
libegobox_gpu.so (C++ fit driver + CUDA kernels); this file only marshals arrays."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import EGX_OK, GpuError, GpParamsStruct, STATUS_NAMES
from .context import GpContext, DEFAULT_NUGGET

# correlation / mean model tags (correlation_models.rs, mean_models.rs)
SquaredExponentialCorr, AbsoluteExponentialCorr, Matern32Corr, Matern52Corr = 0, 1, 2, 3
ConstantMean, LinearMean, QuadraticMean = 0, 1, 2
CORR_NAMES = ["SquaredExponential", "AbsoluteExponential", "Matern32", "Matern52"]
MEAN_NAMES = ["ConstantMean", "LinearMean", "QuadraticMean"]

GP_OPTIM_N_START = 10        # algorithm.rs:33
GP_COBYLA_MIN_EVAL = 25      # algorithm.rs:35
GP_COBYLA_MAX_EVAL = 1000    # algorithm.rs:37


class GpError(Exception):
    """GpError (gp/src/errors.rs:8-40)."""


class LinalgError(GpError):
    """GpError::LinalgError -- Cholesky of a non positive definite R."""


class LikelihoodComputationError(GpError):
    """GpError::LikelihoodComputationError -- ill-conditioned F / ft."""


class InvalidValueError(GpError):
    """GpError::InvalidValueError."""


def _raise_status(st):
    msg = _lib.last_error()
    if st == 1:
        raise LinalgError(msg)
    if st in (2, 3):
        raise LikelihoodComputationError(msg)
    if st == 4:
        raise InvalidValueError(msg)
    raise GpuError(st, msg)


class ThetaTuning:
    """ThetaTuning{Fixed, Full, Partial}, parameters.rs:14-78."""
    DEFAULT_INIT = 1e-1
    DEFAULT_BOUNDS = (1e-2, 1e1)

    def __init__(self, kind, init, bounds=None, active=None):
        self.kind, self.init, self.bounds, self.active = kind, list(np.atleast_1d(init)), bounds, active

    @classmethod
    def Fixed(cls, init):
        return cls(0, init)

    @classmethod
    def Full(cls, init=None, bounds=None):
        return cls(1, [cls.DEFAULT_INIT] if init is None else init,
                   [cls.DEFAULT_BOUNDS] if bounds is None else bounds)

    @classmethod
    def Partial(cls, init, bounds, active):
        return cls(2, init, bounds, list(active))


class GpParams:
    """GpParams builder (parameters.rs:163-273); defaults = GpValidParams::default (:105-120)."""

    def __init__(self, mean=ConstantMean, corr=SquaredExponentialCorr):
        self._mean, self._corr = mean, corr
        self._theta_tuning = ThetaTuning.Full()
        self._kpls_dim = None
        self._w_star = None
        self._n_start = GP_OPTIM_N_START
        self._max_eval = GP_COBYLA_MAX_EVAL
        self._nugget = DEFAULT_NUGGET
        self._device = 0
        self._seed = 42
        self._ftol_rel = 1e-4
        self._optimizer = "cobyla"

    def mean(self, mean):
        self._mean = mean
        return self

    def corr(self, corr):
        self._corr = corr
        return self

    def theta_tuning(self, tuning):
        self._theta_tuning = tuning
        return self

    def theta_init(self, init):
        t = self._theta_tuning
        self._theta_tuning = ThetaTuning(t.kind, init, t.bounds, t.active)
        return self

    def theta_bounds(self, bounds):
        t = self._theta_tuning
        self._theta_tuning = ThetaTuning(t.kind if t.kind != 0 else 1, t.init, bounds, t.active)
        return self

    def kpls_dim(self, kpls_dim, w_star=None):
        """KPLS reduction (algorithm.rs:798-813, 843-855).  The PLS rotations are computed by the
        fit driver (``egx_pls_rotations``, the NIPALS regression of linfa-pls) unless ``w_star``
        (d x kpls_dim) is given."""
        self._kpls_dim, self._w_star = kpls_dim, w_star
        return self

    def n_start(self, n_start):
        self._n_start = n_start
        return self

    def max_eval(self, max_eval):
        self._max_eval = max(GP_COBYLA_MIN_EVAL, max_eval)     # parameters.rs:262-265
        return self

    def nugget(self, nugget):
        self._nugget = nugget
        return self

    def device(self, device):
        self._device = device
        return self

    def cobyla_ftol_rel(self, ftol_rel):
        self._ftol_rel = ftol_rel
        return self

    def optimizer(self, name):
        """"cobyla" (default, the reference's optimiser) or "lbfgsb": projected L-BFGS per start on the closed-form
        theta gradient -- same starts and evaluation budget; not in the reference."""
        if name not in ("cobyla", "lbfgsb"):
            raise InvalidValueError("optimizer should be 'cobyla' or 'lbfgsb', got %r" % (name,))
        self._optimizer = name
        return self

    def check(self):
        """ParamGuard::check_ref, parameters.rs:283-313."""
        if self._kpls_dim is not None and self._kpls_dim < 1:
            raise InvalidValueError("`kpls_dim` canot be 0!")
        t = self._theta_tuning
        if t.kind == 1 and t.bounds is not None and len(t.init) != len(t.bounds) and \
                len(t.init) != 1 and len(t.bounds) != 1:
            raise InvalidValueError("theta_tuning: init and bounds should have the same size")
        return self

    def chain_shard(self, rank, world, exchange):
        """Run only the multistart chains c with c % world == rank here; `exchange(f_best, z_best) -> (f, z)` returns the
        best (objective, log10 theta) over all ranks (see parallel.fit_multistart)."""
        self._chain_shard = (rank, world, exchange)
        return self

    def fit(self, x, y):
        """impl Fit for GpValidParams, algorithm.rs:791-979."""
        self.check()
        lib = _lib.load()
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x[:, None]
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        n, d = x.shape
        if y.shape[0] != n:
            raise InvalidValueError("x and y should have the same number of rows")
        if self._kpls_dim is not None:
            if self._kpls_dim > d:
                raise InvalidValueError(
                    "Dimension reduction %d should be smaller than actual training input dimensions %d"
                    % (self._kpls_dim, d))
        prm = GpParamsStruct()
        lib.egx_gp_params_default(C.byref(prm))
        prm.corr, prm.mean = int(self._corr), int(self._mean)
        t = self._theta_tuning
        prm.theta_tuning = t.kind
        init = np.ascontiguousarray(t.init, dtype=np.float64)
        prm.theta_init = init.ctypes.data_as(C.POINTER(C.c_double))
        prm.n_theta_init = init.size
        keep = [init]
        if t.bounds is not None:
            b = np.ascontiguousarray(t.bounds, dtype=np.float64).reshape(-1, 2)
            prm.theta_bounds = b.ctypes.data_as(C.POINTER(C.c_double))
            prm.n_theta_bounds = b.shape[0]
            keep.append(b)
        if t.active is not None:
            a = np.ascontiguousarray(t.active, dtype=np.int32)
            prm.active = a.ctypes.data_as(C.POINTER(C.c_int))
            prm.n_active = a.size
            keep.append(a)
        prm.n_start, prm.max_eval, prm.nugget = int(self._n_start), int(self._max_eval), float(self._nugget)
        if self._w_star is not None:
            w = np.ascontiguousarray(self._w_star, dtype=np.float64)
            prm.w_star = w.ctypes.data_as(C.POINTER(C.c_double))
            prm.kpls_dim = w.shape[1]
            keep.append(w)
        elif self._kpls_dim is not None:
            prm.kpls_dim = int(self._kpls_dim)
        prm.device, prm.seed, prm.cobyla_ftol_rel = int(self._device), int(self._seed), float(self._ftol_rel)
        prm.optimizer = _lib.EGX_OPT_LBFGSB if self._optimizer == "lbfgsb" else _lib.EGX_OPT_COBYLA
        shard = getattr(self, "_chain_shard", None)
        if shard is not None:                 # multistart chains sharded over ranks (egobox_b200/parallel.py::fit_multistart)
            rank, world, exchange = shard

            def _cb(f_ptr, z_ptr, nz, _user):
                try:
                    z = np.ctypeslib.as_array(z_ptr, shape=(nz,))
                    f, zb = exchange(float(f_ptr[0]), z.copy())
                    f_ptr[0] = f
                    z[:] = zb
                    return 0
                except Exception:             # never unwind through the C frames
                    return 1
            cb = _lib.EXCHANGE_FN(_cb)
            keep.append(cb)
            prm.chain_rank, prm.chain_world, prm.exchange = int(rank), int(world), cb
        h = C.c_void_p()
        st = lib.egx_gp_fit(C.byref(prm), x.ctypes.data_as(C.POINTER(C.c_double)), n, d,
                            y.ctypes.data_as(C.POINTER(C.c_double)), C.byref(h))
        if st != EGX_OK:
            _raise_status(st)
        return GaussianProcess(h, self, (x.copy(), y.copy()))


class GaussianProcess:
    """Trained model (algorithm.rs:174-192); owns the device-resident state."""

    def __init__(self, handle, params, training_data):
        self._lib = _lib.load()
        self._h = handle
        self.params_ = params
        self.training_data = training_data
        n, d, h, p = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._lib.egx_gp_model_dims(self._h, C.byref(n), C.byref(d), C.byref(h), C.byref(p))
        self._n, self._d, self._hdim, self._p = n.value, d.value, h.value, p.value

    @staticmethod
    def params(mean=ConstantMean, corr=SquaredExponentialCorr):
        return GpParams(mean, corr)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.egx_gp_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _x(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x.reshape(-1, self._d)
        if x.shape[1] != self._d:
            raise InvalidValueError("x should have %d columns" % self._d)
        return x

    def predict(self, x):
        x = self._x(x)
        y = np.empty(x.shape[0])
        st = self._lib.egx_gp_model_predict(self._h, x.ctypes.data_as(C.POINTER(C.c_double)), x.shape[0],
                                            y.ctypes.data_as(C.POINTER(C.c_double)))
        if st != EGX_OK:
            _raise_status(st)
        return y

    def predict_var(self, x):
        x = self._x(x)
        v = np.empty(x.shape[0])
        st = self._lib.egx_gp_model_predict_var(self._h, x.ctypes.data_as(C.POINTER(C.c_double)), x.shape[0],
                                                v.ctypes.data_as(C.POINTER(C.c_double)))
        if st != EGX_OK:
            _raise_status(st)
        return v

    def predict_valvar(self, x):
        x = self._x(x)
        y = np.empty(x.shape[0])
        v = np.empty(x.shape[0])
        st = self._lib.egx_gp_model_predict_valvar(self._h, x.ctypes.data_as(C.POINTER(C.c_double)), x.shape[0],
                                                   y.ctypes.data_as(C.POINTER(C.c_double)),
                                                   v.ctypes.data_as(C.POINTER(C.c_double)))
        if st != EGX_OK:
            _raise_status(st)
        return y, v

    def predict_gradients(self, x):
        """algorithm.rs:518-529: (n, nx) matrix of output derivatives."""
        x = self._x(x)
        g = np.empty((x.shape[0], self._d))
        st = self._lib.egx_gp_model_predict_gradients(self._h, x.ctypes.data_as(C.POINTER(C.c_double)), x.shape[0],
                                                      g.ctypes.data_as(C.POINTER(C.c_double)))
        if st != EGX_OK:
            _raise_status(st)
        return g

    def predict_var_gradients(self, x):
        """algorithm.rs:697-704: (n, nx) matrix of variance derivatives."""
        x = self._x(x)
        g = np.empty((x.shape[0], self._d))
        st = self._lib.egx_gp_model_predict_var_gradients(self._h, x.ctypes.data_as(C.POINTER(C.c_double)), x.shape[0],
                                                          g.ctypes.data_as(C.POINTER(C.c_double)))
        if st != EGX_OK:
            _raise_status(st)
        return g

    def predict_valvar_gradients(self, x):
        """algorithm.rs:708-727."""
        return self.predict_gradients(x), self.predict_var_gradients(x)

    def covariance(self, x):
        """algorithm.rs:310-326 `_compute_covariance`: (n, n) conditional covariance at the rows of x."""
        x = self._x(x)
        cov = np.empty((x.shape[0], x.shape[0]))
        st = self._lib.egx_gp_model_covariance(self._h, x.ctypes.data_as(C.POINTER(C.c_double)), x.shape[0],
                                               cov.ctypes.data_as(C.POINTER(C.c_double)))
        if st != EGX_OK:
            _raise_status(st)
        return cov

    def _sample(self, x, n_traj, method, seed=None, z=None):
        x = self._x(x)
        if z is None:
            # the reference draws from an unseeded generator (algorithm.rs:1191-1192)
            z = np.random.default_rng(seed).standard_normal((x.shape[0], int(n_traj)))
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(x.shape[0], -1)
        out = np.empty_like(z)
        st = self._lib.egx_gp_model_sample(self._h, x.ctypes.data_as(C.POINTER(C.c_double)), x.shape[0],
                                           z.ctypes.data_as(C.POINTER(C.c_double)), z.shape[1], int(method),
                                           out.ctypes.data_as(C.POINTER(C.c_double)))
        if st != EGX_OK:
            _raise_status(st)
        return out

    def sample_chol(self, x, n_traj, seed=None, z=None):
        """algorithm.rs:383-385: (n, n_traj) trajectories, Cholesky of the conditional covariance."""
        return self._sample(x, n_traj, 0, seed, z)

    def sample_eig(self, x, n_traj, seed=None, z=None):
        """algorithm.rs:388-390: eigen-decomposition of the conditional covariance (eigenvalues < 1e-9 dropped)."""
        return self._sample(x, n_traj, 1, seed, z)

    def sample(self, x, n_traj, seed=None, z=None):
        """algorithm.rs:393-395: alias of sample_eig."""
        return self.sample_eig(x, n_traj, seed, z)

    def predict_kth_derivatives(self, x, kx):
        """gp/src/algorithm.rs:443-508: derivative of the mean with respect to the kx-th input, (n,) -- one column of the
        batched `predict_gradients` (the reference evaluates it on its own)."""
        kx = int(kx)
        if kx < 0 or kx >= self._d:
            raise InvalidValueError("kx should be in 0..%d, got %d" % (self._d - 1, kx))
        return np.ascontiguousarray(self.predict_gradients(x)[:, kx])

    def q2_score(self, kfold, fit=None):
        """`PredictScore::q2_score`, gp/src/metrics.rs:35-53: 1 - PRESS / TSS over `kfold` refits with the model's own
        parameters (linfa `fold`: validation rows [i fs, (i + 1) fs), fs = n / kfold; TSS around the mean of ALL targets).
        `fit(x, y) -> model with predict` defaults to this model's parameter set (a fan-out of GPU fits)."""
        return _q2_score(self.training_data, kfold, fit if fit is not None else self.params_.fit)

    def looq2_score(self, fit=None):
        """`PredictScore::looq2_score`, gp/src/metrics.rs:55-58: leave-one-out."""
        return self.q2_score(self.training_data[0].shape[0], fit)

    def theta(self):
        th = np.empty(self._hdim)
        self._lib.egx_gp_model_theta(self._h, th.ctypes.data_as(C.POINTER(C.c_double)))
        return th

    def variance(self):
        return float(self._lib.egx_gp_model_variance(self._h))

    def likelihood(self):
        return float(self._lib.egx_gp_model_likelihood(self._h))

    def n_evals(self):
        return int(self._lib.egx_gp_model_n_evals(self._h))

    def dims(self):
        return (self._d, 1)

    def kpls_dim(self):
        return self._hdim if self._hdim < self._d else None

    def inner_params(self, with_chol=True):
        """GpInnerParams (algorithm.rs:47-60) downloaded from the device."""
        n, p = self._n, self._p
        beta, gamma = np.empty((p, 1)), np.empty((n, 1))
        ft, g = np.empty((n, p)), np.empty((p, p))
        chol = np.empty((n, n)) if with_chol else None
        dp = C.POINTER(C.c_double)
        st = self._lib.egx_gp_model_inner_params(self._h, beta.ctypes.data_as(dp), gamma.ctypes.data_as(dp),
                                                 chol.ctypes.data_as(dp) if with_chol else None,
                                                 ft.ctypes.data_as(dp), g.ctypes.data_as(dp))
        if st != EGX_OK:
            _raise_status(st)
        return dict(sigma2=self.variance(), beta=beta, gamma=gamma, r_chol=chol, ft=ft, ft_qr_r=g)

    def normalization(self):
        d, h = self._d, self._hdim
        xm, xs, w = np.empty(d), np.empty(d), np.empty((d, h))
        ym, ys = C.c_double(), C.c_double()
        dp = C.POINTER(C.c_double)
        self._lib.egx_gp_model_normalization(self._h, xm.ctypes.data_as(dp), xs.ctypes.data_as(dp),
                                             C.byref(ym), C.byref(ys), w.ctypes.data_as(dp))
        return dict(x_mean=xm, x_std=xs, y_mean=ym.value, y_std=ys.value, w_star=w)

    def context(self):
        """Borrowed GpContext view on the model's device state (profiling, device-pointer predict)."""
        ctx = GpContext.__new__(GpContext)
        ctx._lib = self._lib
        ctx._h = C.c_void_p(self._lib.egx_gp_model_context(self._h))
        ctx.n, ctx.d, ctx.h, ctx.p = self._n, self._d, self._hdim, self._p
        ctx.close = lambda: None          # owned by the model
        return ctx

    def __str__(self):
        p = self.params_
        return "GP(mean=%s, corr=%s, theta=%s, variance=%s, likelihood=%s)" % (
            MEAN_NAMES[p._mean], CORR_NAMES[p._corr], self.theta(), self.variance(), self.likelihood())


class Kriging:
    """Kriging = GpParams<ConstantMean, SquaredExponentialCorr> (algorithm.rs:200-207)."""

    @staticmethod
    def params():
        return GpParams(ConstantMean, SquaredExponentialCorr)


def _q2_score(training_data, kfold, fit):
    from . import metrics
    x, _y = training_data
    if kfold < 1 or kfold > x.shape[0]:
        raise InvalidValueError("kfold should be in 1..%d, got %d" % (x.shape[0], kfold))
    return metrics.q2_k_score(training_data, kfold, fit)


def bound_cobyla_minimize(fun, x0, bounds, rhobeg=0.5, ftol_rel=1e-4, maxeval=200):
    """Host-only access to the chain optimiser (optimization.rs:122-169 semantics)."""
    lib = _lib.load()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = x0.size
    lo = np.ascontiguousarray([b[0] for b in bounds], dtype=np.float64)
    hi = np.ascontiguousarray([b[1] for b in bounds], dtype=np.float64)

    def _cb(xp, nn, _user):
        return float(fun(np.ctypeslib.as_array(xp, shape=(nn,)).copy()))

    cb = _lib.OBJECTIVE_FN(_cb)
    xopt = np.empty(n)
    fopt = C.c_double()
    nev = C.c_int()
    dp = C.POINTER(C.c_double)
    st = lib.egx_bound_cobyla_minimize(cb, None, n, x0.ctypes.data_as(dp), lo.ctypes.data_as(dp),
                                       hi.ctypes.data_as(dp), rhobeg, ftol_rel, maxeval, xopt.ctypes.data_as(dp),
                                       C.byref(fopt), C.byref(nev))
    if st != EGX_OK:
        _raise_status(st)
    return xopt, fopt.value, nev.value


def bound_lbfgs_minimize(fun_and_grad, x0, bounds, ftol_rel=1e-9, gtol=1e-7, maxeval=200):
    """Host-only access to the per-start optimiser of optimizer("lbfgsb"); fun_and_grad(x) -> (f, grad)."""
    lib = _lib.load()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = x0.size
    lo = np.ascontiguousarray([b[0] for b in bounds], dtype=np.float64)
    hi = np.ascontiguousarray([b[1] for b in bounds], dtype=np.float64)

    def _cb(xp, nn, gp, _user):
        f, g = fun_and_grad(np.ctypeslib.as_array(xp, shape=(nn,)).copy())
        np.ctypeslib.as_array(gp, shape=(nn,))[:] = np.asarray(g, dtype=np.float64)
        return float(f)

    cb = _lib.OBJECTIVE_GRAD_FN(_cb)
    xopt = np.empty(n)
    fopt = C.c_double()
    nev = C.c_int()
    dp = C.POINTER(C.c_double)
    st = lib.egx_bound_lbfgs_minimize(cb, None, n, x0.ctypes.data_as(dp), lo.ctypes.data_as(dp), hi.ctypes.data_as(dp),
                                      ftol_rel, gtol, maxeval, xopt.ctypes.data_as(dp), C.byref(fopt), C.byref(nev))
    if st != EGX_OK:
        _raise_status(st)
    return xopt, fopt.value, nev.value


def prepare_multistart(n_start, theta0, bounds, seed=42):
    lib = _lib.load()
    theta0 = np.ascontiguousarray(theta0, dtype=np.float64)
    dim = theta0.size
    b = np.ascontiguousarray(bounds, dtype=np.float64).reshape(-1, 2)
    if b.shape[0] == 1:
        b = np.repeat(b, dim, axis=0)
    out = np.zeros((n_start + 1, dim))
    dp = C.POINTER(C.c_double)
    st = lib.egx_prepare_multistart(n_start, theta0.ctypes.data_as(dp), np.ascontiguousarray(b).ctypes.data_as(dp),
                                    dim, seed, out.ctypes.data_as(dp))
    if st != EGX_OK:
        _raise_status(st)
    return out


def symmetric_eig(a):
    """Host eigen-decomposition used by the eigenvalue sampler (`cov_x.eigh()`, algorithm.rs:1171-1173):
    returns (w, v) with the eigenvectors as the columns of v (unsorted)."""
    lib = _lib.load()
    v = np.array(a, dtype=np.float64, order="C")
    n = v.shape[0]
    w = np.empty(n)
    dp = C.POINTER(C.c_double)
    st = lib.egx_symmetric_eig(n, v.ctypes.data_as(dp), w.ctypes.data_as(dp))
    if st != EGX_OK:
        _raise_status(st)
    return w, v
