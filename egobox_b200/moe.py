"""Mixture of GP experts: host-side mirror of crates/moe (`GaussianMixture`, `GpMixtureParams`, `GpMixture`).

What runs where:
* the experts (fit, predict*, gradients) are the B200-resident GaussianProcess of ``egobox_b200.gp``;
* the predict-side Gaussian mixture (responsibilities, their derivatives, arg-max clusters) and the hard / smooth
  recombination run on the device behind ``egx_moe_*`` (csrc/moe.cu): points are uploaded once, every expert
  predicts its batch through the device-pointer entry points, one download at the end -- the reference calls
  every expert once PER POINT (moe/src/algorithm.rs:691-1010);
* the cross-validated expert selection (`find_best_expert`, moe/src/algorithm.rs:209-347) is a batch of GPU fits;
* the EM clustering of the training set is the reference's control plane (third-party linfa-clustering
  `GaussianMixtureModel`, moe/src/algorithm.rs:118-124): `fit_gmm` below is a plain numpy EM used only when the
  caller does not bring a mixture (`gmx=`); it runs once per fit on (n, nx+1) data.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from . import gp as _gp

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
HARD, SMOOTH = 0, 1


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class GaussianMixture:
    """moe/src/gaussian_mixture.rs:28-47: weights (k), means (k, nx), covariances (k, nx, nx) + heaviside factor.
    Precisions by Cholesky on the host (k small nx x nx matrices, :182-216); everything per point on the device."""

    def __init__(self, weights, means, covariances, heaviside_factor=1.0, device=0):
        self._lib = _lib.load()
        self._weights = _f64(weights).reshape(-1)
        self._means = _f64(means)
        if self._means.ndim == 1:
            self._means = self._means[None, :]
        k, nx = self._means.shape
        self._covariances = _f64(covariances).reshape(k, nx, nx)
        if self._weights.shape[0] != k:
            raise ValueError("weights / means / covariances disagree on the number of clusters")
        self._factor = float(heaviside_factor)
        self._device = device
        self._h = C.c_void_p()
        st = self._lib.egx_moe_create(C.byref(self._h), int(device), k, nx, self._weights.ctypes.data_as(_dp),
                                      self._means.ctypes.data_as(_dp), self._covariances.ctypes.data_as(_dp),
                                      self._factor)
        if st != _gp.EGX_OK:
            _gp._raise_status(st)
        self._experts = []

    # -- accessors (gaussian_mixture.rs:85-106) -------------------------------------------------
    def n_clusters(self):
        return self._means.shape[0]

    def weights(self):
        return self._weights

    def means(self):
        return self._means

    def covariances(self):
        return self._covariances

    def heaviside_factor(self):
        return self._factor

    def with_heaviside_factor(self, factor):
        """`heaviside_factor(f)` setter :101-106, as a new mixture (the reference clones, algorithm.rs:366-367)."""
        return GaussianMixture(self._weights, self._means, self._covariances, factor, self._device)

    def set_heaviside_factor(self, factor):
        st = self._lib.egx_moe_set_heaviside_factor(self._h, float(factor))
        if st != _gp.EGX_OK:
            _gp._raise_status(st)
        self._factor = float(factor)
        return self

    def parameters(self):
        k, nx = self._means.shape
        prec, pc, ld = np.empty((k, nx, nx)), np.empty((k, nx, nx)), np.empty(k)
        self._lib.egx_moe_parameters(self._h, prec.ctypes.data_as(_dp), pc.ctypes.data_as(_dp), ld.ctypes.data_as(_dp))
        return prec, pc, ld

    def _x(self, x):
        x = _f64(x)
        if x.ndim == 1:
            x = x.reshape(-1, self._means.shape[1])
        if x.shape[1] != self._means.shape[1]:
            raise _gp.InvalidValueError("x should have %d columns" % self._means.shape[1])
        return x

    def predict_probas(self, x):
        """:109-116 -> (m, k) responsibilities."""
        x = self._x(x)
        p = np.empty((x.shape[0], self.n_clusters()))
        st = self._lib.egx_moe_predict_probas(self._h, x.ctypes.data_as(_dp), x.shape[0], p.ctypes.data_as(_dp), None)
        if st != _gp.EGX_OK:
            _gp._raise_status(st)
        return p

    def predict(self, x):
        """:306-318 -> (m,) cluster with the largest responsibility."""
        x = self._x(x)
        c = np.empty(x.shape[0], dtype=np.int32)
        st = self._lib.egx_moe_predict_probas(self._h, x.ctypes.data_as(_dp), x.shape[0], None, c.ctypes.data_as(_ip))
        if st != _gp.EGX_OK:
            _gp._raise_status(st)
        return c.astype(np.int64)

    def predict_probas_derivatives(self, x):
        """:158-170 -> (m, k, nx)."""
        x = self._x(x)
        d = np.empty((x.shape[0], self.n_clusters(), x.shape[1]))
        st = self._lib.egx_moe_predict_probas_derivatives(self._h, x.ctypes.data_as(_dp), x.shape[0],
                                                          d.ctypes.data_as(_dp))
        if st != _gp.EGX_OK:
            _gp._raise_status(st)
        return d

    # -- experts (borrowed device contexts) --------------------------------------------------------
    def _bind_experts(self, experts):
        self._experts = list(experts)             # keep them alive
        for c, e in enumerate(self._experts):
            ctx = self._lib.egx_gp_model_context(e._h)
            st = self._lib.egx_moe_set_expert(self._h, c, ctx)
            if st != _gp.EGX_OK:
                _gp._raise_status(st)

    def _predict(self, recombination, x, want):
        x = self._x(x)
        m, nx = x.shape
        y = np.empty(m) if "y" in want else None
        v = np.empty(m) if "v" in want else None
        gy = np.empty((m, nx)) if "gy" in want else None
        gv = np.empty((m, nx)) if "gv" in want else None

        def ptr(a):
            return a.ctypes.data_as(_dp) if a is not None else None
        st = self._lib.egx_moe_predict(self._h, int(recombination), x.ctypes.data_as(_dp), m, ptr(y), ptr(v), ptr(gy),
                                       ptr(gv))
        if st != _gp.EGX_OK:
            _gp._raise_status(st)
        return y, v, gy, gv

    # -- serde layout (doc/Gpx_Tutorial.ipynb:421 `gmx` block) -----------------------------------------
    def to_dict(self):
        prec, pc, ld = self.parameters()

        def arr(a):
            a = np.asarray(a, dtype=np.float64)
            return {"v": 1, "dim": list(a.shape), "data": a.reshape(-1).tolist()}
        return {"weights": arr(self._weights), "means": arr(self._means), "covariances": arr(self._covariances),
                "precisions": arr(prec), "precisions_chol": arr(pc), "heaviside_factor": self._factor,
                "log_det": arr(ld)}

    @staticmethod
    def from_dict(o, device=0):
        def arr(a):
            return np.array(a["data"], dtype=np.float64).reshape(a["dim"])
        return GaussianMixture(arr(o["weights"]), arr(o["means"]), arr(o["covariances"]),
                               o.get("heaviside_factor", 1.0), device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.egx_moe_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------------
# control plane: EM clustering of the training set (linfa-clustering GaussianMixtureModel in the reference)
# ----------------------------------------------------------------------------------------------------
def _kmeans(data, k, rng, n_iter=100):
    n = data.shape[0]
    centers = [data[rng.integers(n)]]
    for _ in range(1, k):                                           # k-means++ seeding
        d2 = np.min([((data - c) ** 2).sum(axis=1) for c in centers], axis=0)
        tot = d2.sum()
        centers.append(data[rng.choice(n, p=d2 / tot)] if tot > 0 else data[rng.integers(n)])
    centers = np.array(centers)
    labels = np.zeros(n, dtype=int)
    for _ in range(n_iter):
        d2 = ((data[:, None, :] - centers[None, :, :]) ** 2).sum(axis=2)
        new = d2.argmin(axis=1)
        if np.array_equal(new, labels) and _ > 0:
            break
        labels = new
        for c in range(k):
            if np.any(labels == c):
                centers[c] = data[labels == c].mean(axis=0)
    return labels


def _m_step(data, resp, reg_covar):
    nk = resp.sum(axis=0) + 10.0 * np.finfo(np.float64).eps
    means = (resp.T @ data) / nk[:, None]
    k, dim = means.shape
    covs = np.empty((k, dim, dim))
    for c in range(k):
        diff = data - means[c]
        covs[c] = (resp[:, c, None] * diff).T @ diff / nk[c]
        covs[c].flat[::dim + 1] += reg_covar
    return nk / data.shape[0], means, covs


def _e_step(data, weights, means, covs):
    n, dim = data.shape
    k = means.shape[0]
    lp = np.empty((n, k))
    for c in range(k):
        chol = np.linalg.cholesky(covs[c])
        z = np.linalg.solve(chol, (data - means[c]).T)
        lp[:, c] = -0.5 * ((z * z).sum(axis=0) + dim * math.log(2 * math.pi)) - np.log(np.diag(chol)).sum()
    wlp = lp + np.log(weights)
    mx = wlp.max(axis=1, keepdims=True)
    lpn = mx[:, 0] + np.log(np.exp(wlp - mx).sum(axis=1))
    return lpn, wlp - lpn[:, None]


def fit_gmm(data, n_clusters, n_runs=20, tol=1e-3, max_iter=100, reg_covar=1e-6, seed=None):
    """EM for a full-covariance Gaussian mixture (linfa-clustering defaults: k-means initialisation, tolerance 1e-3,
    100 iterations, reg_covar 1e-6; n_runs = 20 as set at moe/src/algorithm.rs:120-123), best lower bound of the runs.
    Returns (weights, means, covariances) over ALL columns of `data` (the caller slices the x block, :127-129)."""
    data = _f64(data)
    rng = np.random.default_rng(seed)
    best = None
    runs = 1 if n_clusters == 1 else n_runs
    for _ in range(runs):
        labels = _kmeans(data, n_clusters, rng) if n_clusters > 1 else np.zeros(data.shape[0], dtype=int)
        resp = np.zeros((data.shape[0], n_clusters))
        resp[np.arange(data.shape[0]), labels] = 1.0
        try:
            w, mu, cov = _m_step(data, resp, reg_covar)
            prev = -math.inf
            for _it in range(max_iter):
                lpn, log_resp = _e_step(data, w, mu, cov)
                w, mu, cov = _m_step(data, np.exp(log_resp), reg_covar)
                lb = lpn.mean()
                if abs(lb - prev) < tol:
                    break
                prev = lb
        except np.linalg.LinAlgError:
            continue
        if best is None or lb > best[0]:
            best = (lb, w, mu, cov)
    if best is None:
        raise _gp.GpError("GMM clustering failed (singular covariances in every run)")
    return best[1], best[2], best[3]


def extract_part(data, quantile):
    """moe/src/algorithm.rs:1111-1122: rows 0, q, 2q, ... are held out; returns (test, train)."""
    n = data.shape[0]
    mask = (np.arange(n) % quantile) != 0
    return data[~mask], data[mask]


def _folds(n, k):
    """linfa `Dataset::fold(k)`: validation rows [i fs, (i + 1) fs) with fs = n / k, everything else trains."""
    fs = n // k
    for i in range(k):
        va = np.arange(i * fs, (i + 1) * fs)
        tr = np.concatenate([np.arange(0, i * fs), np.arange((i + 1) * fs, n)])
        yield tr, va


def _median(v):
    return float(np.median(v)) if len(v) else math.nan


def find_best_number_of_clusters(x, y, max_nb_clusters, fit_mixture, seed=None, gmm_fit=None):
    """moe/src/clustering.rs:59-390 (`NbClusters::Auto`): for 1, 2, ... clusters, 5-fold cross-validation of a mixture
    trained with that many clusters; hard-recombination error sum|pred - y| / sum|y| and smooth(1.0) error
    sum|pred - y| (:196-227, the two are scaled differently in the reference and compared as they are, :354-362);
    the count with the smallest MEDIAN error wins, and with it hard or smooth(None) recombination.  The search stops
    when both medians rose twice in a row (:290-300) and only looks at counts whose every cluster kept more than 3
    values (rows x columns, :169-172) in every fold.

    fit_mixture(n_clusters, xtrain, ytrain) -> object with predict_hard(x), predict_smooth1(x), close(); a raised
    GpError counts as the reference's failed `fit` (:229-233).  Returns (n_clusters, recombination, heaviside) with
    heaviside None = to be optimised."""
    x = _f64(x)
    y = _f64(y).reshape(-1)
    n, nx = x.shape
    if gmm_fit is None:
        gmm_fit = fit_gmm
    if max_nb_clusters == 0:
        max_nb_clusters = n // 10 + 1
    data = np.concatenate([x, y[:, None]], axis=1)
    med_h, med_s, ok_counts = [], [], []
    ok1 = True
    i, stop = 0, False
    while i < max_nb_clusters and not stop:
        k = i + 1
        h_errors, s_errors = [], []
        ok = True
        try:
            gmm = gmm_fit(data, k, seed=seed)
        except _gp.GpError:
            gmm = None
        if gmm is not None:
            for tr, va in _folds(n, 5):
                try:
                    mix = fit_mixture(k, x[tr], y[tr])
                except _gp.GpError:
                    ok = False
                    s_errors.append(1.0)
                    h_errors.append(1.0)
                    continue
                try:
                    labels = np.argmax(_e_step(data[tr], *gmm)[1], axis=1)
                    for c in range(k):
                        ok = ok and int((labels == c).sum()) * (nx + 1) > 3
                    actual = y[va]
                    for which, errs in (("h", h_errors), ("s", s_errors)):
                        try:
                            pred = mix.predict_hard(x[va]) if which == "h" else mix.predict_smooth1(x[va])
                        except _gp.GpError:
                            ok = False
                            errs.append(1.0)
                            continue
                        if np.any(np.isinf(pred)):
                            errs.append(1.0)
                        elif np.any(np.isnan(pred)):
                            ok = False
                            errs.append(1.0)
                        elif which == "h":
                            errs.append(float(np.abs(pred - actual).sum() / np.abs(actual).sum()))
                        else:
                            errs.append(float(np.abs(pred - actual).sum()))
                finally:
                    mix.close()
        if ok and s_errors and h_errors:
            ok_counts.append(i)
        med_s.append(_median(s_errors))
        med_h.append(_median(h_errors))
        if i > 3:
            ok2, ok1 = ok1, ok
            stop = (not ok) and (not ok1) and (not ok2)
            stop = (med_h[i - 1] >= med_h[i - 2] and med_s[i - 1] >= med_s[i - 2] and
                    med_h[i] >= med_h[i - 1] and med_s[i] >= med_s[i - 1])          # the median rule has the last word
        i += 1
    if not ok_counts:
        return 1, SMOOTH, None
    cluster_h = cluster_s = 1
    min_h, min_s = med_h[ok_counts[0]], med_s[ok_counts[0]]
    for k in ok_counts:
        if min_h > med_h[k]:
            min_h, cluster_h = med_h[k], k + 1
        if min_s > med_s[k]:
            min_s, cluster_s = med_s[k], k + 1
    if med_h[cluster_h - 1] < med_s[cluster_s - 1]:
        return cluster_h, HARD, None
    return cluster_s, SMOOTH, None


class _CvMixture:
    """The two predictions the cluster-count search needs from a trained mixture (hard, and smooth with factor 1)."""

    def __init__(self, mix):
        self.mix = mix

    def predict_hard(self, x):
        m = self.mix
        return m.experts[0].predict(x) if m._one() else m.gmx._predict(HARD, x, ("y",))[0]

    def predict_smooth1(self, x):
        m = self.mix
        if m._one():
            return m.experts[0].predict(x)
        if m.gmx.heaviside_factor() != 1.0:
            m.gmx.set_heaviside_factor(1.0)
        return m.gmx._predict(SMOOTH, x, ("y",))[0]

    def close(self):
        self.mix.close()


# ----------------------------------------------------------------------------------------------------
# GpMixtureParams / GpMixture
# ----------------------------------------------------------------------------------------------------
_MEANS = [("Constant", 1, _gp.ConstantMean), ("Linear", 2, _gp.LinearMean), ("Quadratic", 4, _gp.QuadraticMean)]
_CORRS = [("SquaredExponential", 1, _gp.SquaredExponentialCorr), ("AbsoluteExponential", 2, _gp.AbsoluteExponentialCorr),
          ("Matern32", 4, _gp.Matern32Corr), ("Matern52", 8, _gp.Matern52Corr)]


class GpMixtureParams:
    """moe/src/parameters.rs GpMixtureParams: n_clusters, recombination, regression / correlation spec bit sets,
    theta_tunings (one per cluster or one for all), kpls_dim, n_start, max_eval, optional preset `gmx`."""

    def __init__(self):
        self.n_clusters = 1
        self.recombination = HARD
        self.heaviside = None                      # Smooth(Some(f)) when set
        self.regression_spec = 1
        self.correlation_spec = 1
        self.theta_tunings = [_gp.ThetaTuning.Full()]
        self.kpls_dim = None
        self.w_star = None
        self.n_start = 10
        self.max_eval = 1000                       # moe/src/parameters.rs:153
        self.gmx = None
        self.seed = None
        self.device = 0

    def set(self, **kw):
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)
        return self

    # -- expert construction -------------------------------------------------------------------------
    def _gp_params(self, mean, corr, tuning, n_start, max_eval):
        p = (_gp.GaussianProcess.params(mean, corr).theta_tuning(tuning).n_start(n_start).max_eval(max_eval)
             .device(self.device))
        if self.kpls_dim is not None:
            p = p.kpls_dim(self.kpls_dim, self.w_star)
        return p

    def _cv_error(self, mean_name, mean, corr, x, y):
        """compute_error!, moe/src/expertise_macros.rs:14-51 (default GP params + kpls_dim; folds of linfa iter_fold)."""
        n, nx = x.shape
        k = min(n, 5)
        if k < 4 * nx and mean_name == "Quadratic":
            return math.inf
        if k < 3 * nx and mean_name == "Linear":
            return math.inf
        fs = n // k
        errs = []
        for i in range(k):
            va = np.arange(i * fs, (i + 1) * fs)
            tr = np.concatenate([np.arange(0, i * fs), np.arange((i + 1) * fs, n)])
            prm = self._gp_params(mean, corr, _gp.ThetaTuning.Full(), 10, 1000)
            gp = prm.fit(x[tr], y[tr])             # `.unwrap()` in the reference: a failed fit propagates
            errs.append(float(np.linalg.norm(y[va] - gp.predict(x[va]))))
            gp.close()
        return sum(errs) / len(errs)

    def find_best_expert(self, nc, x, y):
        """moe/src/algorithm.rs:209-347: single spec -> train it; several -> 5-fold CV error decides."""
        means = [(nm, m) for nm, bit, m in _MEANS if self.regression_spec & bit]
        corrs = [(nm, c) for nm, bit, c in _CORRS if self.correlation_spec & bit]
        if not means or not corrs:
            raise _gp.InvalidValueError("empty regression / correlation specification")
        errors = None
        if len(means) == 1 and len(corrs) == 1:
            best = (means[0], corrs[0])
        else:
            errors = [((mn, m), (cn, c), self._cv_error(mn, m, c, x, y)) for mn, m in means for cn, c in corrs]
            best = min(errors, key=lambda e: e[2] if not math.isnan(e[2]) else math.inf)[:2]
        tuning = self.theta_tunings[0] if (nc > 0 and len(self.theta_tunings) == 1) else self.theta_tunings[nc]
        gp = self._gp_params(best[0][1], best[1][1], tuning, self.n_start, self.max_eval).fit(x, y)
        gp.cv_errors_ = None if errors is None else {"%s_%s" % (e[0][0], e[1][0]): e[2] for e in errors}
        return gp

    # -- training ------------------------------------------------------------------------------------------
    def fit(self, xt, yt):
        """GpMixtureValidParams::train, moe/src/algorithm.rs:73-143."""
        xt = _f64(xt)
        yt = _f64(yt).reshape(-1)
        if self.n_clusters < 1:
            # NbClusters::Auto { max } (moe/src/algorithm.rs:85-101): n_clusters = 0 -> up to n / 10 + 1, -m -> up to m
            def cv_fit(k, xtr, ytr):
                # GpMixtureParams::default() + the three specs (clustering.rs:148-154); its preset `gmm` is never read
                # by `train` (algorithm.rs:118 looks at `gmx` only), so every fold clusters its own training rows
                p = GpMixtureParams().set(n_clusters=k, regression_spec=self.regression_spec,
                                          correlation_spec=self.correlation_spec, kpls_dim=self.kpls_dim,
                                          seed=self.seed, device=self.device)
                return _CvMixture(p.fit(xtr, ytr))

            max_nb = -self.n_clusters if self.n_clusters < 0 else xt.shape[0] // 10 + 1
            k, recomb, heaviside = find_best_number_of_clusters(xt, yt, max_nb, cv_fit, seed=self.seed)
            chosen = GpMixtureParams()
            chosen.__dict__.update(self.__dict__)
            chosen.n_clusters, chosen.recombination, chosen.heaviside = k, recomb, heaviside
            if len(chosen.theta_tunings) != k:
                chosen.theta_tunings = [self.theta_tunings[0]]          # one tuning for all experts (gp_mix.rs:210-214)
            return chosen.fit(xt, yt)
        # GpMixtureParams::check (moe/src/parameters.rs): one theta tuning for all experts or one per cluster
        if len(self.theta_tunings) not in (1, self.n_clusters):
            raise _gp.InvalidValueError("Number of theta tunings should be 1 or the number of clusters (%d), got %d"
                                        % (self.n_clusters, len(self.theta_tunings)))
        nx = xt.shape[1]
        data = np.concatenate([xt, yt[:, None]], axis=1)
        multi = self.n_clusters > 1
        smooth_auto = self.recombination == SMOOTH and self.heaviside is None and multi
        training = extract_part(data, 5)[1] if smooth_auto else data        # :108-116
        if self.gmx is not None:
            gmx = self.gmx
        else:
            w, mu, cov = fit_gmm(training, self.n_clusters, seed=self.seed)     # :118-124
            factor = self.heaviside if (self.recombination == SMOOTH and self.heaviside is not None) else 1.0
            gmx = GaussianMixture(w, mu[:, :nx], cov[:, :nx, :nx], factor, self.device)     # :127-136
        return self.train_on_clusters(xt, yt, gmx)

    def train_on_clusters(self, xt, yt, gmx):
        """moe/src/algorithm.rs:147-206."""
        xt = _f64(xt)
        yt = _f64(yt).reshape(-1)
        nx = xt.shape[1]
        k = gmx.n_clusters()
        labels = gmx.predict(xt)
        clusters = [np.nonzero(labels == c)[0] for c in range(k)]          # sort_by_cluster, clustering.rs:33-57
        if k > 1:                                                         # check_number_of_points :383-407
            need = (nx + 1) * (nx + 2) // 2 if self.regression_spec & 4 else (nx + 1 if self.regression_spec & 2 else 1)
            for rows in clusters:
                if rows.size * (nx + 1) < need:
                    raise _gp.GpError("Not enough points in training set. Need %d points, got %d"
                                      % (need, rows.size * (nx + 1)))
                if rows.size < 3:
                    raise _gp.GpError("Not enough points in cluster, requires at least 3, got %d" % rows.size)
        experts = [self.find_best_expert(c, xt[rows], yt[rows]) for c, rows in enumerate(clusters)]
        smooth_auto = self.recombination == SMOOTH and self.heaviside is None and k > 1
        if smooth_auto:                                                   # :182-195
            data = np.concatenate([xt, yt[:, None]], axis=1)
            test = extract_part(data, 5)[0]
            factor = optimize_heaviside_factor(experts, gmx, test[:, :nx], test[:, nx])
            for e in experts:
                e.close()
            # moe/src/algorithm.rs:186-192: retrain with `GpMixtureParams::from(self.clone())` and the optimised factor; a mixture
            # the caller preset (self.gmx) is kept -- train() reuses it (:118-119) -- only its heaviside factor changes
            again = GpMixtureParams()
            again.__dict__.update(self.__dict__)
            again.heaviside = factor
            if self.gmx is not None:
                preset = GaussianMixture(self.gmx.weights(), self.gmx.means(), self.gmx.covariances(), factor, self.gmx._device)
                again.gmx = preset
            return again.fit(xt, yt)
        if self.recombination == SMOOTH and self.heaviside is not None and gmx.heaviside_factor() != self.heaviside:
            gmx.set_heaviside_factor(self.heaviside)
        return GpMixture(experts, gmx, self.recombination, (xt.copy(), yt.copy()), self)


def optimize_heaviside_factor(experts, gmx, xtest, ytest):
    """moe/src/algorithm.rs:353-380: the factor of linspace(0.1, 2.1, 20) with the smallest smooth-prediction error
    on the held-out rows (1 when every error is below 1e-6)."""
    factors = np.linspace(0.1, 2.1, 20)
    probe = GaussianMixture(gmx.weights(), gmx.means(), gmx.covariances(), 1.0, gmx._device)
    probe._bind_experts(experts)
    errs = []
    xn = math.sqrt(float((xtest * xtest).sum()))
    for f in factors:
        probe.set_heaviside_factor(float(f))
        y = probe._predict(SMOOTH, xtest, ("y",))[0]
        errs.append(math.sqrt(float(((y - ytest) ** 2).sum())) / xn)
    probe.close()
    if max(errs) < 1e-6:
        return 1.0
    return float(factors[int(np.argmin(errs))])


class GpMixture:
    """moe/src/algorithm.rs:426-440: experts + gmx + recombination; prediction dispatch :455-541."""

    def __init__(self, experts, gmx, recombination, training_data, params):
        self.experts = list(experts)
        self.gmx = gmx
        self.recombination = recombination
        self.training_data = training_data
        self.params_ = params
        gmx._bind_experts(self.experts)

    @staticmethod
    def params():
        return GpMixtureParams()

    def n_clusters(self):
        return self.gmx.n_clusters()

    def dims(self):
        return self.experts[0].dims()

    def _one(self):
        return len(self.experts) == 1

    def predict(self, x):
        return self.experts[0].predict(x) if self._one() else self.gmx._predict(self.recombination, x, ("y",))[0]

    def predict_var(self, x):
        return self.experts[0].predict_var(x) if self._one() else self.gmx._predict(self.recombination, x, ("v",))[1]

    def predict_valvar(self, x):
        if self._one():
            return self.experts[0].predict_valvar(x)
        return self.gmx._predict(self.recombination, x, ("y", "v"))[:2]

    def predict_gradients(self, x):
        if self._one():
            return self.experts[0].predict_gradients(x)
        return self.gmx._predict(self.recombination, x, ("gy",))[2]

    def predict_var_gradients(self, x):
        if self._one():
            return self.experts[0].predict_var_gradients(x)
        return self.gmx._predict(self.recombination, x, ("gv",))[3]

    def predict_valvar_gradients(self, x):
        if self._one():
            return self.experts[0].predict_valvar_gradients(x)
        return self.gmx._predict(self.recombination, x, ("gy", "gv"))[2:]

    def sample(self, x, n_traj, seed=None):
        """moe/src/algorithm.rs:543-558: only for a single cluster."""
        if not self._one():
            raise _gp.GpError("Can not sample when several clusters %d" % self.n_clusters())
        return self.experts[0].sample(x, n_traj, seed=seed)

    def sample_expert(self, ith, x, n_traj, seed=None):
        """moe/src/algorithm.rs:1010-1020: trajectories of the ith expert alone."""
        if ith < 0 or ith >= len(self.experts):
            raise _gp.InvalidValueError("expert index should be in 0..%d, got %d" % (len(self.experts) - 1, ith))
        return self.experts[ith].sample(x, n_traj, seed=seed)

    def set_recombination(self, recombination):
        """moe/src/algorithm.rs:632-639: switches hard / smooth; like the reference it leaves the mixture's heaviside
        factor alone (the smooth prediction reads it from `gmx`)."""
        if recombination not in (HARD, SMOOTH):
            raise _gp.InvalidValueError("recombination should be HARD or SMOOTH")
        self.recombination = recombination
        return self

    # -- cross-validation scores (moe/src/algorithm.rs:566-600 -> moe/src/metrics.rs) ------------------------------
    def _refit(self):
        if self.params_ is None:
            raise _gp.GpError("this mixture was not built by GpMixtureParams.fit: no parameters to refit with")
        return self.params_.fit

    def q2_k(self, kfold):
        from . import metrics
        return metrics.q2_k_score(self.training_data, kfold, self._refit())

    def q2(self):
        return self.q2_k(self.training_data[0].shape[0])

    def pva_k(self, kfold):
        from . import metrics
        return metrics.pva_k_score(self.training_data, kfold, self._refit())

    def pva(self):
        return self.pva_k(self.training_data[0].shape[0])

    def iae_alpha_k(self, kfold, with_plot_data=False):
        from . import metrics
        score, alphas, deltas = metrics.iae_alpha_k_score(self.training_data, kfold, self._refit())
        return (score, alphas, deltas) if with_plot_data else score

    def iae_alpha(self, with_plot_data=False):
        return self.iae_alpha_k(self.training_data[0].shape[0], with_plot_data)

    def recombination_name(self):
        if self.recombination == HARD:
            return "Hard"
        return "Smooth(%s)" % self.gmx.heaviside_factor()

    def __str__(self):
        return "Mixture[%s](%s)" % (self.recombination_name(), ", ".join(_expert_str(e) for e in self.experts))

    def close(self):
        self.gmx.close()
        for e in self.experts:
            e.close()


def _expert_str(gp):
    p = gp.params_
    return "%s_%sGP(mean=%s, corr=%s, theta=%s, variance=%s, likelihood=%s)" % (
        _gp.MEAN_NAMES[p._mean].replace("Mean", ""), _gp.CORR_NAMES[p._corr], _gp.MEAN_NAMES[p._mean],
        _gp.CORR_NAMES[p._corr], gp.theta().tolist(), gp.variance(), gp.likelihood())
