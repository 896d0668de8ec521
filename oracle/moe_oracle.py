"""CPU oracle for the mixture-of-experts recombination  --  TEST INFRASTRUCTURE ONLY.

numpy restatement of crates/moe/src/gaussian_mixture.rs (the predict-side Gaussian mixture with
its heaviside factor) and of the recombination formulas of crates/moe/src/algorithm.rs
(`predict_smooth` :411-423, `predict_var_smooth` :670-685, `predict_gradients_smooth` :691-733,
`predict_var_gradients_smooth` :739-783, the `*_hard` variants :879-1010) and the cross-validated
expert selection (`find_best_expert` :209-347, `compute_error!` expertise_macros.rs:14-51).
Pinned by tests/test_moe_oracle.py on the reference's `test_pdfs` known answers
(gaussian_mixture.rs:371-397) and `test_gmx_one_cluster` (:344-358).  Citations are file:line under
/root/reference/crates/moe/src/.  Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np

EPS = np.finfo(np.float64).eps
MIN_10_EXP = -307.0            # f64::MIN_10_EXP, gaussian_mixture.rs:246


class GaussianMixture:
    """gaussian_mixture.rs:28-47, 62-83."""

    def __init__(self, weights, means, covariances, heaviside_factor=1.0):
        self.weights = np.asarray(weights, dtype=np.float64)
        self.means = np.atleast_2d(np.asarray(means, dtype=np.float64))
        self.covariances = np.asarray(covariances, dtype=np.float64)
        k, nx = self.means.shape
        assert self.covariances.shape == (k, nx, nx) and self.weights.shape == (k,)
        # compute_precisions_cholesky :182-205: (L^-1)^T with L = chol(cov)
        self.precisions_chol = np.stack([np.linalg.inv(np.linalg.cholesky(c)).T for c in self.covariances])
        # compute_precisions :208-216
        self.precisions = np.stack([pc @ pc.T for pc in self.precisions_chol])
        self.heaviside_factor = float(heaviside_factor)

    def n_clusters(self):
        return self.means.shape[0]

    def with_heaviside_factor(self, f):
        return GaussianMixture(self.weights, self.means, self.covariances, f)

    def _log_det(self):
        # compute_log_det :221-227 + compute_log_det_cholesky :286-299
        f = self.heaviside_factor ** -0.5
        return np.array([np.log(np.diag(pc * f)).sum() for pc in self.precisions_chol])

    def log_gaussian_prob(self, x):
        """compute_log_gaussian_prob :260-283."""
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        k, nx = self.means.shape
        f = self.heaviside_factor ** -0.5
        out = np.zeros((x.shape[0], k))
        for c in range(k):
            diff = (x - self.means[c]) @ (self.precisions_chol[c] * f)
            out[:, c] = (diff * diff).sum(axis=1)
        cst = nx * math.log(2.0 * math.pi)
        return -0.5 * (out + cst) + self._log_det()

    def log_prob_resp(self, x):
        """compute_log_prob_resp :231-256 (with its underflow / zero clamps)."""
        wlp = self.log_gaussian_prob(x) + np.log(self.weights)
        e = np.where(wlp <= MIN_10_EXP, 0.0, np.exp(wlp))
        s = e.sum(axis=1)
        with np.errstate(divide="ignore"):
            lpn = np.where(np.abs(s) < EPS, 0.0, np.log(np.where(s > 0, s, 1.0)))
        return lpn, wlp - lpn[:, None]

    def predict_probas(self, x):
        """:109-116."""
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        if self.n_clusters() == 1:
            return np.ones((x.shape[0], 1))
        return np.exp(self.log_prob_resp(x)[1])

    def predict(self, x):
        """PredictInplace :306-318: index of the largest responsibility."""
        return np.argmax(np.exp(self.log_prob_resp(x)[1]), axis=1)

    def pdfs(self, x):
        """:173-176."""
        return np.exp(self.log_gaussian_prob(np.asarray(x, dtype=np.float64)[None, :]))[0]

    def predict_single_probas_derivatives(self, x):
        """:122-152 -> (k, nx)."""
        x = np.asarray(x, dtype=np.float64)
        pdf = self.pdfs(x)
        v = self.weights @ pdf
        precs = self.precisions / self.heaviside_factor
        deriv = np.stack([(x - self.means[c]) @ precs[c] for c in range(self.n_clusters())])
        vprime = (deriv * (-self.weights * pdf)[:, None]).sum(axis=0)
        u = (self.weights * pdf)[:, None]
        uprime = -(deriv * u)
        return (uprime * v - u * vprime[None, :]) / (v * v)

    def predict_probas_derivatives(self, x):
        """:158-170 -> (m, k, nx)."""
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        return np.stack([self.predict_single_probas_derivatives(xi) for xi in x])


# --------------------------------------------------------------------------------------------
# recombination (algorithm.rs); `experts` expose predict / predict_var / predict_gradients /
# predict_var_gradients on (m, nx) arrays (oracle.gp_oracle.GaussianProcess does)
# --------------------------------------------------------------------------------------------
def predict_smooth(experts, gmx, x):
    p = gmx.predict_probas(x)
    return sum(e.predict(x) * p[:, i] for i, e in enumerate(experts))


def predict_var_smooth(experts, gmx, x):
    p = gmx.predict_probas(x)
    return sum(e.predict_var(x) * p[:, i] * p[:, i] for i, e in enumerate(experts))


def predict_hard(experts, gmx, x, what="predict"):
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    cl = gmx.predict(x)
    out = None
    for c, e in enumerate(experts):
        idx = np.nonzero(cl == c)[0]
        if idx.size == 0:
            continue
        r = getattr(e, what)(x[idx])
        if out is None:
            out = np.zeros((x.shape[0],) + r.shape[1:])
        out[idx] = r
    return out


def predict_gradients_smooth(experts, gmx, x):
    """:691-733: sum_k p_k grad y_k + sum_k grad p_k * y_k."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    p = gmx.predict_probas(x)
    dp = gmx.predict_probas_derivatives(x)
    out = np.zeros_like(x)
    for i, e in enumerate(experts):
        out += p[:, i:i + 1] * e.predict_gradients(x) + dp[:, i, :] * e.predict(x)[:, None]
    return out


def predict_var_gradients_smooth(experts, gmx, x):
    """:739-783: sum_k p_k^2 grad v_k + 2 sum_k p_k grad p_k v_k."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    p = gmx.predict_probas(x)
    dp = gmx.predict_probas_derivatives(x)
    out = np.zeros_like(x)
    for i, e in enumerate(experts):
        out += (p[:, i:i + 1] ** 2) * e.predict_var_gradients(x) \
            + 2.0 * p[:, i:i + 1] * dp[:, i, :] * e.predict_var(x)[:, None]
    return out


def optimize_heaviside_factor(experts, gmx, xtest, ytest):
    """:353-380: best factor among linspace(0.1, 2.1, 20) on the held-out rows; 1 if all errors < 1e-6."""
    factors = np.linspace(0.1, 2.1, 20)
    errs = np.array([np.sqrt(((predict_smooth(experts, gmx.with_heaviside_factor(f), xtest) - ytest) ** 2).sum())
                     / np.sqrt((xtest ** 2).sum()) for f in factors])
    if errs.max() < 1e-6:
        return 1.0
    return float(factors[int(np.argmin(errs))])


def extract_part(data, quantile):
    """clustering.rs `extract_part`: every `quantile`-th row is held out (test), the rest trains."""
    n = data.shape[0]
    idx_test = np.arange(0, n, quantile)
    mask = np.ones(n, dtype=bool)
    mask[idx_test] = False
    return data[idx_test], data[mask]


def cv_folds(n, k):
    """linfa `iter_fold(k)`: fold size n // k, fold i validates rows [i*fs, (i+1)*fs), the rest trains."""
    fs = n // k
    for i in range(k):
        valid = np.arange(i * fs, (i + 1) * fs)
        train = np.concatenate([np.arange(0, i * fs), np.arange((i + 1) * fs, n)])
        yield train, valid


def cv_error(fit_fn, x, y, regr_name):
    """compute_error!, expertise_macros.rs:14-51: mean over folds of the L2 norm of the validation residual;
    +inf when there are too few points for a Linear / Quadratic trend."""
    n, nx = x.shape
    k = min(n, 5)
    if k < 4 * nx and regr_name == "Quadratic":
        return math.inf
    if k < 3 * nx and regr_name == "Linear":
        return math.inf
    errs = []
    for tr, va in cv_folds(n, k):
        gp = fit_fn(x[tr], y[tr])
        errs.append(float(np.linalg.norm(y[va] - gp.predict(x[va]))))
    return sum(errs) / len(errs)
