"""CPU oracle for the sparse GP (FITC / VFE) branch of the path -- TEST INFRASTRUCTURE ONLY.

numpy/scipy (fp64) restatement of crates/gp/src/sparse_algorithm.rs @ be16128:
  compute_k :676-691, fitc :695-765, vfe :769-830, predict :237-241, predict_var :245-257,
  fit driver :416-648 (params = [theta..., sigma2, (noise)] on a log10 scale, zero mean, NO input
  normalisation), make_inducings :833-847.
Pinning.  `make_inducings` (the seeded choice of the inducing points) IS pinned: the Rust stream behind it
(rand_xoshiro 0.6.0 + rand 0.8.5 shuffle, oracle/rust_rng.py) reproduces every digit of the reference's
seed-42 LHS fixture (crates/doe/src/lhs.rs:332-347; tests/test_host_rng.py).
The LIKELIHOOD VALUES stay "parity unpinned" by the reference: its tests hold only loose smoke bounds (0.5 /
0.3 in sparse_algorithm.rs:941,944) and the one printed value, doc/SparseGpx_Tutorial.ipynb:226
(likelihood 281.125279453634 at theta 9.7394, sigma2 0.6625, noise 0.009574; numpy RandomState(0) data, nz = 30,
seed = 42), cannot be reproduced even by the reference itself: the mixture hands its experts
`let seed = self.rng().gen()` (crates/moe/src/algorithm.rs:330) whose target type is Option<u64>
(surrogates.rs:42) -- rand's Standard draws a bool first, and for Xoshiro256Plus::seed_from_u64(42) that bool is
false -> seed None -> `Xoshiro256Plus::from_entropy()` (sparse_algorithm.rs:457-460): the 30 inducing points of
the notebook run were drawn from OS entropy.  tests/test_host_rng.py::test_sparse_notebook_value_is_entropy_seeded
records this, together with the three deterministic readings of the seed (u64 draw, Option draw forced to Some,
42 itself), which give 274.19 / 289.61 / 229.12 at the printed hyper-parameters.  Beyond that this file is pinned
by its internal consistency tests (Woodbury identities against the dense FITC / VFE formulas) in
tests/test_sgp_oracle.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import scipy.linalg as sla

from . import gp_oracle as O

FITC, VFE = 0, 1
SGP_DEFAULT_NUGGET = O.DEFAULT_NUGGET                         # GpValidParams::default, parameters.rs:118
SGP_THETA_BOUNDS = (1e-2, 1e2)                               # sparse_parameters.rs:155-168
NOISE_DEFAULT = (1e-2, (100.0 * np.finfo(np.float64).eps, 1e10))   # sparse_parameters.rs:25-32


USE_FAST_KERNEL = False      # full-size tests: build K with the C/OpenMP restatement (no (M N) x d temporaries)


def compute_k(corr, a, b, w_star, theta, sigma2):
    """sparse_algorithm.rs:676-691."""
    if USE_FAST_KERNEL:
        from oracle import fast
        return fast.cross_corr(corr, a, b, theta, w_star) * sigma2
    dx = O.pairwise_differences(a, b)
    r = O.corr_value(corr, dx, theta, w_star)
    return r.reshape(a.shape[0], b.shape[0]) * sigma2


@dataclass
class WoodburyData:
    vec: np.ndarray      # (M, 1)
    inv: np.ndarray      # (M, M)


def _tri_inv(l):
    return sla.solve_triangular(l, np.eye(l.shape[0]), lower=True, check_finite=False)


def fitc(corr, theta, sigma2, noise, w_star, xtrain, ytrain, z, nugget):
    """sparse_algorithm.rs:695-765 (natural logs)."""
    nz = z.shape[0]
    knn = np.full(xtrain.shape[0], sigma2)
    kmm = compute_k(corr, z, z, w_star, theta, sigma2) + np.eye(nz) * nugget
    kmn = compute_k(corr, z, xtrain, w_star, theta, sigma2)
    u = sla.cholesky(kmm, lower=True, check_finite=False)
    ui = _tri_inv(u)
    v = ui.dot(kmn)
    nu = knn - (v * v).sum(axis=0) + noise
    beta = 1.0 / nu
    a = np.eye(nz) + (v * beta[None, :]).dot(v.T)
    l = sla.cholesky(a, lower=True, check_finite=False)
    li = _tri_inv(l)
    ay = ytrain * beta[:, None]
    b = li.dot(v).dot(ay)
    term1 = np.log(nu).sum()
    term2 = 2.0 * np.log(np.diag(l)).sum()
    term3 = float(ay.T.dot(ytrain)[0, 0])
    term4 = -float((b * b).sum())
    lik = -0.5 * (term1 + term2 + term3 + term4)
    li_ui = li.dot(ui)
    return lik, WoodburyData(vec=li_ui.T.dot(b), inv=ui.T.dot(ui) - li_ui.T.dot(li_ui))


def vfe(corr, theta, sigma2, noise, w_star, xtrain, ytrain, z, nugget):
    """sparse_algorithm.rs:769-830."""
    nz = z.shape[0]
    n = ytrain.shape[0]
    kmm = compute_k(corr, z, z, w_star, theta, sigma2) + np.eye(nz) * nugget
    kmn = compute_k(corr, z, xtrain, w_star, theta, sigma2)
    u = sla.cholesky(kmm, lower=True, check_finite=False)
    ui = _tri_inv(u)
    v = ui.dot(kmn)
    beta = 1.0 / max(noise, nugget)
    a = v.dot(v.T) * beta
    l = sla.cholesky(np.eye(nz) + a, lower=True, check_finite=False)
    li = _tri_inv(l)
    b = li.dot(v).dot(ytrain) * beta
    term1 = -n * math.log(beta)
    term2 = 2.0 * np.log(np.diag(l)).sum()
    term3 = beta * float((ytrain * ytrain).sum())
    term4 = -float(b.T.dot(b)[0, 0])
    term5 = n * beta * sigma2
    term6 = -float(np.trace(a))
    lik = -0.5 * (term1 + term2 + term3 + term4 + term5 + term6)
    li_ui = li.dot(ui)
    bi = np.eye(nz) + li.T.dot(li)
    return lik, WoodburyData(vec=li_ui.T.dot(b), inv=ui.T.dot(bi).dot(ui))


def reduced_likelihood(method, corr, theta, sigma2, noise, w_star, xtrain, ytrain, z, nugget=SGP_DEFAULT_NUGGET):
    """sparse_algorithm.rs:654-673.  Raises numpy LinAlgError where the reference `.unwrap()`s."""
    f = fitc if method == FITC else vfe
    return f(corr, np.asarray(theta, dtype=np.float64), float(sigma2), float(noise), w_star,
             np.asarray(xtrain, dtype=np.float64), np.asarray(ytrain, dtype=np.float64).reshape(-1, 1),
             np.asarray(z, dtype=np.float64), nugget)


@dataclass
class SparseGaussianProcess:
    """sparse_algorithm.rs:145-169."""
    corr: int
    method: int
    theta: np.ndarray
    sigma2: float
    noise: float
    likelihood: float
    w_data: WoodburyData
    w_star: np.ndarray
    inducings: np.ndarray

    def predict(self, x):
        """:237-241"""
        kx = compute_k(self.corr, np.asarray(x, dtype=np.float64), self.inducings, self.w_star, self.theta, self.sigma2)
        return kx.dot(self.w_data.vec)[:, 0]

    def predict_var(self, x):
        """:245-257"""
        x = np.asarray(x, dtype=np.float64)
        kx = compute_k(self.corr, self.inducings, x, self.w_star, self.theta, self.sigma2)
        var = self.sigma2 - (self.w_data.inv.T.dot(kx) * kx).sum(axis=0)
        return np.where(var < 1e-15, 1e-15 + self.noise, var + self.noise)


def make_inducings(n_inducing, xt, rng):
    """:833-847.  `rng`: an int seed (-> the reference's own index stream, oracle/rust_rng.py) or a numpy Generator
    (synthetic test inputs that need no particular stream)."""
    if isinstance(rng, (int, np.integer)):
        from . import rust_rng
        idx = rust_rng.inducing_indices(xt.shape[0], n_inducing, int(rng))
    else:
        idx = rng.permutation(xt.shape[0])[: min(n_inducing, xt.shape[0])]
    return xt[idx].copy()


def build(method, corr, theta, sigma2, noise, x, y, z, w_star=None, nugget=SGP_DEFAULT_NUGGET):
    """The trained object for given hyper-parameters (the tail of fit, :621-647)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.eye(x.shape[1]) if w_star is None else w_star
    lik, wd = reduced_likelihood(method, corr, theta, sigma2, noise, w, x, y, z, nugget)
    return SparseGaussianProcess(corr=corr, method=method, theta=np.asarray(theta, dtype=np.float64),
                                 sigma2=float(sigma2), noise=float(noise), likelihood=lik, w_data=wd, w_star=w,
                                 inducings=np.asarray(z, dtype=np.float64))
