"""Internal-consistency pins of the sparse-GP oracle (the reference has no tight fixture here):
the Woodbury quantities of fitc / vfe must reproduce the dense formulas they come from."""
import numpy as np
import pytest

from oracle import gp_oracle as O
from oracle import sgp_oracle as S


def _data(n=60, m=9, d=2, seed=0):
    rng = np.random.default_rng(seed)
    x = 2 * rng.random((n, d)) - 1
    y = np.sin(3 * np.pi * x[:, 0]) + 0.3 * np.cos(2 * x[:, -1]) + rng.normal(0, 0.1, n)
    z = S.make_inducings(m, x, rng)
    return x, y, z


@pytest.mark.parametrize("corr", [O.SQEXP, O.MATERN52])
def test_fitc_matches_dense_formulas(corr):
    x, y, z = _data()
    theta, sigma2, noise = np.array([1.5, 0.7]), 0.9, 0.02
    nug = 1e-10
    w = np.eye(2)
    lik, wd = S.reduced_likelihood(S.FITC, corr, theta, sigma2, noise, w, x, y, z, nug)
    kmm = S.compute_k(corr, z, z, w, theta, sigma2) + nug * np.eye(len(z))
    kmn = S.compute_k(corr, z, x, w, theta, sigma2)
    qnn = kmn.T @ np.linalg.solve(kmm, kmn)
    lam = np.diag(sigma2 - np.diag(qnn) + noise)
    cov = qnn + lam                                      # FITC covariance of y
    sign, logdet = np.linalg.slogdet(cov)
    dense = -0.5 * (logdet + y @ np.linalg.solve(cov, y))
    assert lik == pytest.approx(dense, rel=1e-9)
    # predictive mean / variance against the dense FITC posterior
    xs = np.linspace(-1, 1, 7)[:, None] * np.ones((1, 2))
    ks = S.compute_k(corr, xs, z, w, theta, sigma2)
    sig = np.linalg.inv(kmm + kmn @ np.linalg.solve(lam, kmn.T))
    mean = ks @ sig @ kmn @ np.linalg.solve(lam, y)
    var = sigma2 - np.einsum("ij,jk,ik->i", ks, np.linalg.inv(kmm) - sig, ks) + noise
    gp = S.build(S.FITC, corr, theta, sigma2, noise, x, y, z, nugget=nug)
    np.testing.assert_allclose(gp.predict(xs), mean, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(gp.predict_var(xs), var, rtol=1e-7, atol=1e-9)


def test_vfe_matches_dense_bound():
    x, y, z = _data(seed=3)
    theta, sigma2, noise = np.array([2.0, 1.0]), 1.1, 0.05
    nug = 1e-10
    w = np.eye(2)
    lik, wd = S.reduced_likelihood(S.VFE, O.SQEXP, theta, sigma2, noise, w, x, y, z, nug)
    kmm = S.compute_k(O.SQEXP, z, z, w, theta, sigma2) + nug * np.eye(len(z))
    kmn = S.compute_k(O.SQEXP, z, x, w, theta, sigma2)
    qnn = kmn.T @ np.linalg.solve(kmm, kmn)
    n = len(y)
    cov = qnn + noise * np.eye(n)
    sign, logdet = np.linalg.slogdet(cov)
    dense = -0.5 * (logdet + y @ np.linalg.solve(cov, y)) - 0.5 / noise * (n * sigma2 - np.trace(qnn))
    assert lik == pytest.approx(dense, rel=1e-9)


def test_predict_var_floor():
    x, y, z = _data(seed=5)
    gp = S.build(S.FITC, O.SQEXP, [1.0, 1.0], 1.0, 0.01, x, y, z)
    v = gp.predict_var(z)              # at the inducing points the latent variance is ~0
    assert np.all(v >= 0.01 - 1e-12)
