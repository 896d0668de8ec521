"""`Gpx` / `GpMix`: the Python surface of python/src/gp_mix.rs:31-496 for the kriging path.

Same builder signature, defaults and return shapes as the PyO3 classes; the model behind
it is the B200-resident GaussianProcess of ``egobox_b200.gp``.  The mixture-of-experts
control plane of egobox-moe (GMM clustering, multi-spec cross-validated model selection,
smooth recombination) is the CALLER of this path and is out of scope (SURVEY.md section 8,
rows (f)-2): ``n_clusters`` other than 1 and multi-bit ``regr_spec`` / ``corr_spec`` raise
``NotImplementedError`` instead of silently doing something else."""
from __future__ import annotations

import json

import numpy as np

from . import gp as _gp


class RegressionSpec:            # python/src/types.rs
    CONSTANT, LINEAR, QUADRATIC, ALL = 1, 2, 4, 7


class CorrelationSpec:
    SQUARED_EXPONENTIAL, ABSOLUTE_EXPONENTIAL, MATERN32, MATERN52, ALL = 1, 2, 4, 8, 15


class Recombination:
    HARD, SMOOTH = 0, 1
    Hard, Smooth = 0, 1


_REGR = {1: _gp.ConstantMean, 2: _gp.LinearMean, 4: _gp.QuadraticMean}
_CORR = {1: _gp.SquaredExponentialCorr, 2: _gp.AbsoluteExponentialCorr, 4: _gp.Matern32Corr, 8: _gp.Matern52Corr}
EGO_GP_OPTIM_N_START = 10      # ego/src/solver/egor_config.rs:13
EGO_GP_OPTIM_MAX_EVAL = 50     # ego/src/solver/egor_config.rs:15
MOE_GP_MAX_EVAL = 1000         # moe/src/parameters.rs:153 -- what GpMix.fit really runs with (gp_mix.rs:222-231)


class GpMix:
    """Gaussian processes mixture builder (python/src/gp_mix.rs:31-236)."""

    def __init__(self, regr_spec=RegressionSpec.CONSTANT, corr_spec=CorrelationSpec.SQUARED_EXPONENTIAL,
                 kpls_dim=None, n_clusters=1, recombination=Recombination.HARD, theta_init=None,
                 theta_bounds=None, n_start=EGO_GP_OPTIM_N_START, max_eval=EGO_GP_OPTIM_MAX_EVAL, seed=None,
                 w_star=None, device=0):
        self.regr_spec, self.corr_spec, self.kpls_dim = regr_spec, corr_spec, kpls_dim
        self.n_clusters, self.recombination = n_clusters, recombination
        self.theta_init, self.theta_bounds = theta_init, theta_bounds
        self.n_start, self.max_eval, self.seed = n_start, max_eval, seed
        self.w_star, self.device = w_star, device

    def fit(self, xt, yt):
        """gp_mix.rs:140-236: accepts 1-D or 2-D xt, yt must be single-output."""
        xt = np.asarray(xt, dtype=np.float64)
        if xt.ndim == 1:
            xt = xt[:, None]
        elif xt.ndim != 2:
            raise ValueError("Bad training input data")           # gp_mix.rs:147-150
        yt = np.asarray(yt, dtype=np.float64)
        if yt.ndim == 2:
            if yt.shape[1] != 1:
                raise ValueError("Bad training output data")      # gp_mix.rs:157-160
            yt = yt[:, 0]
        elif yt.ndim != 1:
            raise ValueError("Bad training output data")
        if self.n_clusters != 1:
            raise NotImplementedError("n_clusters != 1: GMM clustering is egobox-moe's control plane (out of scope)")
        if self.regr_spec not in _REGR or self.corr_spec not in _CORR:
            raise NotImplementedError("multi-model selection by cross-validation (moe/src/algorithm.rs:209-347) "
                                      "is out of scope: pass a single regr_spec / corr_spec")
        tuning = _gp.ThetaTuning.Full()
        if self.theta_init is not None:
            tuning = _gp.ThetaTuning.Full(list(self.theta_init), [_gp.ThetaTuning.DEFAULT_BOUNDS])
        if self.theta_bounds is not None:
            tuning = _gp.ThetaTuning.Full(tuning.init, [tuple(b) for b in self.theta_bounds])
        n_start = self.n_start
        if n_start < 0:                                            # gp_mix.rs:200-206
            tuning = _gp.ThetaTuning.Fixed(tuning.init)
            n_start = 0
        params = (_gp.GaussianProcess.params(_REGR[self.regr_spec], _CORR[self.corr_spec])
                  .theta_tuning(tuning).n_start(n_start).max_eval(MOE_GP_MAX_EVAL).device(self.device))
        if self.kpls_dim is not None:
            params = params.kpls_dim(self.kpls_dim, self.w_star)
        return Gpx(params.fit(xt, yt), self)


class Gpx:
    """A trained Gaussian processes mixture with one expert (python/src/gp_mix.rs:242-496)."""

    def __init__(self, model, builder):
        self._gp = model
        self._builder = builder

    @staticmethod
    def builder(regr_spec=RegressionSpec.CONSTANT, corr_spec=CorrelationSpec.SQUARED_EXPONENTIAL, kpls_dim=None,
                n_clusters=1, recombination=Recombination.HARD, theta_init=None, theta_bounds=None,
                n_start=EGO_GP_OPTIM_N_START, max_eval=EGO_GP_OPTIM_MAX_EVAL, seed=None, **kw):
        return GpMix(regr_spec, corr_spec, kpls_dim, n_clusters, recombination, theta_init, theta_bounds,
                     n_start, max_eval, seed, **kw)

    def predict(self, x):
        return self._gp.predict(np.asarray(x, dtype=np.float64))

    def predict_var(self, x):
        return self._gp.predict_var(np.asarray(x, dtype=np.float64))

    def predict_valvar(self, x):
        return self._gp.predict_valvar(np.asarray(x, dtype=np.float64))

    def predict_gradients(self, x):
        """gp_mix.rs:373-383: (nsamples, nx) output derivatives."""
        return self._gp.predict_gradients(np.asarray(x, dtype=np.float64))

    def predict_var_gradients(self, x):
        """gp_mix.rs:394-404: (nsamples, nx) variance derivatives."""
        return self._gp.predict_var_gradients(np.asarray(x, dtype=np.float64))

    def sample(self, x, n_traj, seed=None):
        """gp_mix.rs:415-425: (nsamples, n_traj) trajectories of the (single-cluster) surrogate;
        moe/src/algorithm.rs:550-558 -> GaussianProcess::sample (eigenvalue variant)."""
        return self._gp.sample(np.asarray(x, dtype=np.float64), int(n_traj), seed=seed)

    def dims(self):
        return self._gp.dims()

    def training_data(self):
        x, y = self._gp.training_data
        return x.copy(), y.copy()

    def thetas(self):
        return self._gp.theta()[None, :]

    def variances(self):
        return np.array([self._gp.variance()])

    def likelihoods(self):
        return np.array([self._gp.likelihood()])

    def gp(self):
        return self._gp

    # ---- persistence (gp_mix.rs:310-337; moe/src/algorithm.rs:510-524, 1096-1106) -----------------
    def _expert_dict(self):
        """The `experts[0]` object in the reference's serde-JSON layout (ndarray = {"v":1,"dim":[..],"data":[..]}),
        as printed in doc/Gpx_Tutorial.ipynb:421 (GpInnerParams gp/src/algorithm.rs:41-60, GaussianProcess :165-192)."""
        gp = self._gp
        p = gp.params_
        ip = gp.inner_params()
        nz = gp.normalization()
        x, y = gp.training_data
        xn = (x - nz["x_mean"]) / nz["x_std"]
        yn = ((y - nz["y_mean"]) / nz["y_std"])[:, None]

        def arr(a):
            a = np.asarray(a, dtype=np.float64)
            return {"v": 1, "dim": list(a.shape), "data": a.reshape(-1).tolist()}
        t = p._theta_tuning
        if t.kind == 0:
            tuning = {"Fixed": arr(t.init)}
        else:
            b = [list(map(float, bb)) for bb in (t.bounds or [_gp.ThetaTuning.DEFAULT_BOUNDS])]
            tuning = {"Full": {"init": arr(t.init), "bounds": {"v": 1, "dim": [len(b)], "data": b}}}
            if t.kind == 2:
                tuning = {"Partial": {"init": arr(t.init), "bounds": {"v": 1, "dim": [len(b)], "data": b},
                                      "active": list(t.active)}}
        mean_name = _gp.MEAN_NAMES[p._mean]
        return {
            "type_fullgp": "Gp%s%sSurrogate" % (mean_name.replace("Mean", ""), _gp.CORR_NAMES[p._corr]),
            "theta": arr(gp.theta()), "likelihood": gp.likelihood(),
            "inner_params": {"sigma2": ip["sigma2"], "beta": arr(ip["beta"]), "gamma": arr(ip["gamma"]),
                             "r_chol": arr(ip["r_chol"]), "ft": arr(ip["ft"]), "ft_qr_r": arr(ip["ft_qr_r"])},
            "w_star": arr(nz["w_star"]),
            "xt_norm": {"data": arr(xn), "mean": arr(nz["x_mean"]), "std": arr(nz["x_std"])},
            "yt_norm": {"data": arr(yn), "mean": arr([nz["y_mean"]]), "std": arr([nz["y_std"]])},
            "training_data": [arr(x), arr(y)],
            "params": {"theta_tuning": tuning, "mean": mean_name, "corr": _gp.CORR_NAMES[p._corr],
                       "kpls_dim": p._kpls_dim, "n_start": p._n_start, "max_eval": p._max_eval, "nugget": p._nugget},
        }

    def save(self, filename):
        """JSON only: {"recombination": "Hard", "experts": [<expert in the reference layout>]}.  The mixture-level
        blocks of the reference file (gmx, gp_type, params with the RNG state) belong to egobox-moe's control plane
        and are not written, so stock egobox cannot load this file as a GpMixture."""
        if not str(filename).endswith(".json"):
            raise NotImplementedError("bincode persistence is out of scope; use a .json filename")
        with open(filename, "w") as f:
            json.dump({"recombination": "Hard", "experts": [self._expert_dict()]}, f)
        return True

    @staticmethod
    def load(filename, device=0):
        """Rebuild the device-resident model from a file written by save() -- or from the `experts[0]` block of a
        stock egobox JSON -- by one final evaluation at the stored theta (Fixed tuning)."""
        with open(filename) as f:
            obj = json.load(f)
        e = obj["experts"][0] if "experts" in obj else obj

        def arr(o):
            return np.array(o["data"], dtype=np.float64).reshape(o["dim"])
        prm = e["params"]
        mean = _gp.MEAN_NAMES.index(prm["mean"])
        corr = _gp.CORR_NAMES.index(prm["corr"])
        x, y = arr(e["training_data"][0]), arr(e["training_data"][1])
        params = (_gp.GaussianProcess.params(mean, corr).theta_tuning(_gp.ThetaTuning.Fixed(arr(e["theta"])))
                  .nugget(prm.get("nugget", _gp.DEFAULT_NUGGET)).device(device))
        w = arr(e["w_star"])
        if w.shape[1] < w.shape[0]:
            params = params.kpls_dim(w.shape[1], w)
        builder = GpMix(regr_spec=1 << mean, corr_spec=1 << corr, n_start=-1, theta_init=arr(e["theta"]).tolist())
        return Gpx(params.fit(x, y), builder)

    def __str__(self):
        p = self._gp.params_
        return "Mixture[Hard](%s_%sGP(mean=%s, corr=%s, theta=%s, variance=%s, likelihood=%s))" % (
            _gp.MEAN_NAMES[p._mean].replace("Mean", ""), _gp.CORR_NAMES[p._corr], _gp.MEAN_NAMES[p._mean],
            _gp.CORR_NAMES[p._corr], self._gp.theta().tolist(), self._gp.variance(), self._gp.likelihood())

    def __repr__(self):
        return json.dumps({"recombination": "Hard", "experts": [{
            "theta": self._gp.theta().tolist(), "likelihood": self._gp.likelihood(),
            "variance": self._gp.variance()}]})
