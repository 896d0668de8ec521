"""`Gpx` / `GpMix`: the Python surface of python/src/gp_mix.rs:31-496 for the kriging path.

Same builder signature, defaults and return shapes as the PyO3 classes; the model behind it is a
``egobox_b200.moe.GpMixture``: one B200-resident GaussianProcess expert per cluster, multi-bit
``regr_spec`` / ``corr_spec`` resolved by the 5-fold cross-validation of moe/src/algorithm.rs:209-347
(a batch of GPU fits), hard / smooth recombination on the device; ``n_clusters <= 0`` picks the number of clusters (and hard / smooth recombination) by the
cross-validated search of moe/src/clustering.rs:59-390 (``moe.find_best_number_of_clusters``)."""
from __future__ import annotations

import json

import numpy as np

from . import gp as _gp
from . import moe as _moe


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
                 w_star=None, device=0, gmx=None):
        self.regr_spec, self.corr_spec, self.kpls_dim = regr_spec, corr_spec, kpls_dim
        self.n_clusters, self.recombination = n_clusters, recombination
        self.theta_init, self.theta_bounds = theta_init, theta_bounds
        self.n_start, self.max_eval, self.seed = n_start, max_eval, seed
        self.w_star, self.device = w_star, device
        self.gmx = gmx            # optional preset moe.GaussianMixture (GpMixtureParams::gmx, moe/src/parameters.rs)

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
        tuning = _gp.ThetaTuning.Full()
        if self.theta_init is not None:
            tuning = _gp.ThetaTuning.Full(list(self.theta_init), [_gp.ThetaTuning.DEFAULT_BOUNDS])
        if self.theta_bounds is not None:
            tuning = _gp.ThetaTuning.Full(tuning.init, [tuple(b) for b in self.theta_bounds])
        n_start = self.n_start
        if n_start < 0:                                            # gp_mix.rs:200-206
            tuning = _gp.ThetaTuning.Fixed(tuning.init)
            n_start = 0
        # n_clusters = 0: automatic, < 0: automatic up to -n_clusters (gp_mix.rs:197-201); one tuning then serves all experts
        params = _moe.GpMixtureParams().set(
            n_clusters=int(self.n_clusters), recombination=int(self.recombination),
            regression_spec=int(self.regr_spec), correlation_spec=int(self.corr_spec),
            theta_tunings=[tuning] * max(int(self.n_clusters), 1), kpls_dim=self.kpls_dim, w_star=self.w_star,
            n_start=n_start, max_eval=MOE_GP_MAX_EVAL, gmx=self.gmx, seed=self.seed, device=self.device)
        return Gpx(params.fit(xt, yt), self)


def _arr(a):
    a = np.asarray(a, dtype=np.float64)
    return {"v": 1, "dim": list(a.shape), "data": a.reshape(-1).tolist()}


def _unarr(o):
    return np.array(o["data"], dtype=np.float64).reshape(o["dim"])


def _tuning_dict(t):
    if t.kind == 0:
        return {"Fixed": _arr(t.init)}
    b = [list(map(float, bb)) for bb in (t.bounds or [_gp.ThetaTuning.DEFAULT_BOUNDS])]
    full = {"init": _arr(t.init), "bounds": {"v": 1, "dim": [len(b)], "data": b}}
    if t.kind == 2:
        return {"Partial": dict(full, active=list(t.active))}
    return {"Full": full}


def _spec_names(bits, names):
    return " | ".join(n for n, b in names if bits & b)


class Gpx:
    """A trained Gaussian processes mixture (python/src/gp_mix.rs:242-496) over a `moe.GpMixture`."""

    def __init__(self, model, builder):
        if isinstance(model, _gp.GaussianProcess):          # a single expert: wrap it in a one-cluster mixture
            x, y = model.training_data
            w, mu, cov = _moe.fit_gmm(np.concatenate([x, y[:, None]], axis=1), 1)
            nx = x.shape[1]
            gmx = _moe.GaussianMixture(w, mu[:, :nx], cov[:, :nx, :nx], 1.0, getattr(builder, "device", 0))
            model = _moe.GpMixture([model], gmx, Recombination.HARD, (x, y), None)
        self._mix = model
        self._gp = model.experts[0]
        self._builder = builder

    @staticmethod
    def builder(regr_spec=RegressionSpec.CONSTANT, corr_spec=CorrelationSpec.SQUARED_EXPONENTIAL, kpls_dim=None,
                n_clusters=1, recombination=Recombination.HARD, theta_init=None, theta_bounds=None,
                n_start=EGO_GP_OPTIM_N_START, max_eval=EGO_GP_OPTIM_MAX_EVAL, seed=None, **kw):
        return GpMix(regr_spec, corr_spec, kpls_dim, n_clusters, recombination, theta_init, theta_bounds,
                     n_start, max_eval, seed, **kw)

    def predict(self, x):
        return self._mix.predict(np.asarray(x, dtype=np.float64))

    def predict_var(self, x):
        return self._mix.predict_var(np.asarray(x, dtype=np.float64))

    def predict_valvar(self, x):
        return self._mix.predict_valvar(np.asarray(x, dtype=np.float64))

    def predict_gradients(self, x):
        """gp_mix.rs:373-383: (nsamples, nx) output derivatives."""
        return self._mix.predict_gradients(np.asarray(x, dtype=np.float64))

    def predict_var_gradients(self, x):
        """gp_mix.rs:394-404: (nsamples, nx) variance derivatives."""
        return self._mix.predict_var_gradients(np.asarray(x, dtype=np.float64))

    def sample(self, x, n_traj, seed=None):
        """gp_mix.rs:415-425: (nsamples, n_traj) trajectories of the (single-cluster) surrogate;
        moe/src/algorithm.rs:550-558 -> GaussianProcess::sample (eigenvalue variant)."""
        return self._mix.sample(np.asarray(x, dtype=np.float64), int(n_traj), seed=seed)

    def dims(self):
        return self._mix.dims()

    def training_data(self):
        x, y = self._mix.training_data
        return x.copy(), y.copy()

    def thetas(self):
        """gp_mix.rs:457-468: (n_clusters, theta dimension)."""
        return np.stack([e.theta() for e in self._mix.experts])

    def variances(self):
        return np.array([e.variance() for e in self._mix.experts])

    def likelihoods(self):
        return np.array([e.likelihood() for e in self._mix.experts])

    def gp(self):
        return self._gp

    def mixture(self):
        return self._mix

    # ---- persistence (gp_mix.rs:310-337; moe/src/algorithm.rs:510-524, 1096-1106) -----------------
    @staticmethod
    def _expert_dict(gp):
        """One `experts[i]` object in the reference's serde-JSON layout (ndarray = {"v":1,"dim":[..],"data":[..]}),
        as printed in doc/Gpx_Tutorial.ipynb:421 (GpInnerParams gp/src/algorithm.rs:41-60, GaussianProcess :165-192)."""
        p = gp.params_
        ip = gp.inner_params()
        nz = gp.normalization()
        x, y = gp.training_data
        xn = (x - nz["x_mean"]) / nz["x_std"]
        yn = ((y - nz["y_mean"]) / nz["y_std"])[:, None]
        mean_name = _gp.MEAN_NAMES[p._mean]
        return {
            "type_fullgp": "Gp%s%sSurrogate" % (mean_name.replace("Mean", ""), _gp.CORR_NAMES[p._corr]),
            "theta": _arr(gp.theta()), "likelihood": gp.likelihood(),
            "inner_params": {"sigma2": ip["sigma2"], "beta": _arr(ip["beta"]), "gamma": _arr(ip["gamma"]),
                             "r_chol": _arr(ip["r_chol"]), "ft": _arr(ip["ft"]), "ft_qr_r": _arr(ip["ft_qr_r"])},
            "w_star": _arr(nz["w_star"]),
            "xt_norm": {"data": _arr(xn), "mean": _arr(nz["x_mean"]), "std": _arr(nz["x_std"])},
            "yt_norm": {"data": _arr(yn), "mean": _arr([nz["y_mean"]]), "std": _arr([nz["y_std"]])},
            "training_data": [_arr(x), _arr(y)],
            "params": {"theta_tuning": _tuning_dict(p._theta_tuning), "mean": mean_name, "corr": _gp.CORR_NAMES[p._corr],
                       "kpls_dim": p._kpls_dim, "n_start": p._n_start, "max_eval": p._max_eval, "nugget": p._nugget},
        }

    def _recombination_value(self):
        if self._mix.recombination == Recombination.HARD:
            return "Hard"
        return {"Smooth": self._mix.gmx.heaviside_factor()}

    def to_dict(self):
        """The GpMixture serde layout of doc/Gpx_Tutorial.ipynb:421: recombination, experts, gmx, gp_type,
        training_data, params (the `rng` state of the params block is the reference's Xoshiro stream, which this
        implementation does not carry: it is written as zeros)."""
        mix, b = self._mix, self._builder
        x, y = mix.training_data
        prm = mix.params_
        regr = prm.regression_spec if prm is not None else 1 << self._gp.params_._mean
        corr = prm.correlation_spec if prm is not None else 1 << self._gp.params_._corr
        tunings = prm.theta_tunings if prm is not None else [self._gp.params_._theta_tuning]
        return {
            "recombination": self._recombination_value(),
            "experts": [self._expert_dict(e) for e in mix.experts],
            "gmx": mix.gmx.to_dict(),
            "gp_type": "FullGp",
            "training_data": [_arr(x), _arr(y)],
            "params": {"gp_type": "FullGp", "n_clusters": {"Fixed": {"nb": mix.n_clusters()}},
                       "recombination": "Hard" if mix.recombination == Recombination.HARD else {"Smooth": None},
                       "regression_spec": _spec_names(regr, [("CONSTANT", 1), ("LINEAR", 2), ("QUADRATIC", 4)]),
                       "correlation_spec": _spec_names(corr, [("SQUAREDEXPONENTIAL", 1), ("ABSOLUTEEXPONENTIAL", 2),
                                                              ("MATERN32", 4), ("MATERN52", 8)]),
                       "theta_tunings": [_tuning_dict(t) for t in tunings],
                       "kpls_dim": getattr(b, "kpls_dim", None), "n_start": max(getattr(b, "n_start", 10), 0),
                       "max_eval": MOE_GP_MAX_EVAL, "gmm": None, "gmx": None, "rng": {"s": [0, 0, 0, 0]}},
        }

    def save(self, filename):
        """gp_mix.rs:310-320: `.json` -> serde JSON in the reference's GpMixture layout (see to_dict), any other name ->
        bincode 2 standard configuration of the same structure (egobox_b200/bincode.py)."""
        if str(filename).endswith(".json"):
            with open(filename, "w") as f:
                json.dump(self.to_dict(), f)
        else:
            from . import bincode
            with open(filename, "wb") as f:
                f.write(bincode.encode_mixture(self.to_dict()))
        return True

    @staticmethod
    def _load_expert(e, device):
        prm = e["params"]
        mean = _gp.MEAN_NAMES.index(prm["mean"])
        corr = _gp.CORR_NAMES.index(prm["corr"])
        x, y = _unarr(e["training_data"][0]), _unarr(e["training_data"][1])
        params = (_gp.GaussianProcess.params(mean, corr).theta_tuning(_gp.ThetaTuning.Fixed(_unarr(e["theta"])))
                  .nugget(prm.get("nugget", _gp.DEFAULT_NUGGET)).device(device))
        w = _unarr(e["w_star"])
        # KPLS is on when the stored parameters say so (kpls_dim == d is legal: w_star is then a d x d rotation, not the
        # identity); files without the key fall back to the shape test
        if prm.get("kpls_dim") is not None or w.shape[1] < w.shape[0]:
            params = params.kpls_dim(w.shape[1], w)
        return params.fit(x, y)

    @staticmethod
    def load(filename, device=0):
        """Rebuild the device-resident mixture from a file written by save() or by stock egobox (JSON): every expert
        by one final evaluation at its stored theta (Fixed tuning), the Gaussian mixture from the `gmx` block."""
        if str(filename).endswith(".json"):
            with open(filename) as f:
                obj = json.load(f)
        else:
            from . import bincode
            with open(filename, "rb") as f:
                obj = bincode.decode_mixture(f.read())
        experts_json = obj["experts"] if "experts" in obj else [obj]
        experts = [Gpx._load_expert(e, device) for e in experts_json]
        first = experts_json[0]
        builder = GpMix(regr_spec=1 << _gp.MEAN_NAMES.index(first["params"]["mean"]),
                        corr_spec=1 << _gp.CORR_NAMES.index(first["params"]["corr"]), n_start=-1,
                        n_clusters=len(experts), theta_init=_unarr(first["theta"]).tolist(), device=device)
        if "gmx" not in obj:
            return Gpx(experts[0], builder)
        gmx = _moe.GaussianMixture.from_dict(obj["gmx"], device)
        rec = obj.get("recombination", "Hard")
        recomb = Recombination.HARD if rec == "Hard" else Recombination.SMOOTH
        if isinstance(rec, dict) and rec.get("Smooth") is not None:
            gmx.set_heaviside_factor(float(rec["Smooth"]))
        builder.recombination = recomb
        td = obj.get("training_data")
        training = (_unarr(td[0]), _unarr(td[1])) if td else experts[0].training_data
        return Gpx(_moe.GpMixture(experts, gmx, recomb, training, None), builder)

    def __str__(self):
        return str(self._mix)

    def __repr__(self):
        return json.dumps({"recombination": self._recombination_value(), "experts": [{
            "theta": e.theta().tolist(), "likelihood": e.likelihood(), "variance": e.variance()}
            for e in self._mix.experts]})
