"""-m gpu: BASELINE.json's FULL sizes (C2 n=8192 d=10, C3 N=100 000 M=1024, C4 n=4096 d=20).

Two kinds of checks, per the parity rules:
  * direct parity against the oracle where it still finishes in seconds -- the C/OpenMP correlation build of
    `oracle/fast.py` + LAPACK (about 2 s per likelihood at n = 8192 on a many-core host);
  * size-independent properties of the domain: L L^T = R through random probe vectors (a checksum of the
    factorisation), interpolation of the training data with vanishing variance, bit-identical re-evaluation
    (idempotence) and batch == single, linearity of the predictor in y and y-independence of var / sigma2.
"""
import numpy as np
import pytest

from oracle import gp_oracle as O
from oracle import sgp_oracle as S

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1200)]


def _fast():
    from oracle import fast
    return fast


def _workload(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    z = 4.0 * x - 2.0
    y = np.sum(100.0 * (z[:, 1:] - z[:, :-1] ** 2) ** 2 + (1.0 - z[:, :-1]) ** 2, axis=1)
    return x, y


@pytest.fixture(scope="module")
def c2():
    """C2 context, finalized at theta = 1 (shared by the tests of this module: 0.55 GB per workspace)."""
    import egobox_b200 as eg
    n, d = 8192, 10
    x, y = _workload(n, d, 2024)
    xn, xm, xs = O.normalize(x)
    yn, ym, ys = O.normalize(y.reshape(-1, 1))
    ctx = eg.GpContext(xn, yn[:, 0], xm, xs, float(ym[0]), float(ys[0]), eg.MATERN52, eg.CONSTANT)
    theta = np.full(d, 1.0)
    yield dict(ctx=ctx, x=x, y=y, xn=xn, xm=xm, xs=xs, yn=yn, ym=float(ym[0]), ys=float(ys[0]), theta=theta, n=n, d=d)
    ctx.close()


def test_c2_likelihood_and_predict_against_oracle(c2):
    """Direct parity at n = 8192: rlf, sigma2, beta, predict, predict_var (bars 1e-6; held to 1e-9 / 1e-7 / 1e-6)."""
    fast = _fast()
    ctx, theta, d = c2["ctx"], c2["theta"], c2["d"]
    fx = O.mean_value(O.CONSTANT, c2["xn"])
    rlf_ref, inner = fast.reduced_likelihood(O.MATERN52, c2["xn"], fx, c2["yn"], c2["ys"], theta, np.eye(d))
    st, rlf = ctx.reduced_likelihood(theta)
    assert st == 0
    assert rlf == pytest.approx(rlf_ref, rel=1e-9)
    st, res = ctx.finalize(theta, want_ft=False)
    assert st == 0
    assert res["sigma2"] == pytest.approx(inner.sigma2, rel=1e-8)
    np.testing.assert_allclose(res["beta"], inner.beta[:, 0], rtol=1e-7, atol=1e-10)
    gp = O.GaussianProcess(corr=O.MATERN52, mean=O.CONSTANT, theta=theta, likelihood=rlf_ref, inner=inner,
                           w_star=np.eye(d), xt_norm=c2["xn"], x_mean=c2["xm"], x_std=c2["xs"], yt_norm=c2["yn"],
                           y_mean=c2["ym"], y_std=c2["ys"])
    xs = np.random.default_rng(5).random((384, d))
    yo, vo = fast.predict_valvar(gp, xs, chunk=384)
    yg, vg = ctx.predict_valvar(xs)
    np.testing.assert_allclose(yg, yo, rtol=1e-7, atol=1e-8 * np.abs(yo).max())
    np.testing.assert_allclose(vg, vo, rtol=1e-6, atol=1e-8 * inner.sigma2)


def test_c2_factorisation_checksum_and_interpolation(c2):
    """L (L^T v) = R v for random probes v (checksum of the whole factor, O(n^2) on the host), the GP interpolates
    its training data and its variance vanishes there."""
    ctx, theta, n = c2["ctx"], c2["theta"], c2["n"]
    R = ctx.correlation_matrix(theta)               # K1 alone (also clears the trained state)
    st, res = ctx.finalize(theta, want_ft=False)
    assert st == 0
    L = ctx.download_chol()
    assert np.all(np.diag(L) > 0) and np.allclose(np.triu(L, 1), 0.0)
    rng = np.random.default_rng(1)
    for _ in range(3):
        v = rng.standard_normal(n)
        lhs = L.dot(L.T.dot(v))
        rhs = R.dot(v)
        assert np.abs(lhs - rhs).max() <= 1e-11 * np.abs(rhs).max()
    # log-determinant term of the likelihood from the downloaded factor
    logdet = 2.0 / n * np.log10(np.diag(L)).sum()
    assert res["rlf"] == pytest.approx(-n * (np.log10(res["sigma2"] / c2["ys"] ** 2) + logdet), rel=1e-10)
    idx = rng.choice(n, 1500, replace=False)
    yv, vv = ctx.predict_valvar(c2["x"][idx])
    np.testing.assert_allclose(yv, c2["y"][idx], rtol=0, atol=1e-6 * np.abs(c2["y"]).max())
    assert vv.max() <= 1e-6 * res["sigma2"]


def test_c2_idempotence_and_batch(c2):
    """Re-evaluation is bit-identical, and so are equal candidates inside a batch.  The batched entry point
    (up to 8 workspaces in flight, no look-ahead: every K = 256 update through the tcgen05 kernel) and the single
    call (look-ahead: the next pair's two block columns are updated by the DMMA kernel) may round differently in
    the last bit."""
    ctx, theta = c2["ctx"], c2["theta"]
    st1, a = ctx.reduced_likelihood(theta)
    st2, b = ctx.reduced_likelihood(theta)
    assert st1 == 0 and st2 == 0 and a == b
    thetas = np.tile(theta, (6, 1)) * np.array([1.0, 0.7, 1.3, 1.0, 2.0, 0.7])[:, None]
    st, rl = ctx.reduced_likelihood_batch(thetas)
    assert np.all(st == 0)
    assert rl[0] == rl[3] and rl[1] == rl[5] and rl[0] == pytest.approx(a, rel=1e-13)
    for k in (1, 2, 4):
        assert ctx.reduced_likelihood(thetas[k])[1] == pytest.approx(rl[k], rel=1e-13)


def test_c2_linearity_in_y(c2):
    """At fixed theta the kriging predictor is linear in the observations and var / sigma2 does not depend on them."""
    import egobox_b200 as eg
    d, theta = c2["d"], c2["theta"]
    rng = np.random.default_rng(9)
    y1 = c2["yn"][:, 0]
    y2 = np.sin(3.0 * c2["xn"][:, 0]) + 0.5 * c2["xn"][:, 1]
    xs = rng.random((256, d))
    out = []
    for yy in (y1, y2, y1 + 2.0 * y2):
        ctx = eg.GpContext(c2["xn"], yy, c2["xm"], c2["xs"], 0.0, 1.0, eg.MATERN52, eg.CONSTANT)
        st, res = ctx.finalize(theta, want_ft=False)
        assert st == 0
        yv, vv = ctx.predict_valvar(xs)
        out.append((yv, vv / res["sigma2"]))
        ctx.close()
    scale = np.abs(out[2][0]).max()
    np.testing.assert_allclose(out[2][0], out[0][0] + 2.0 * out[1][0], rtol=0, atol=1e-9 * scale)
    np.testing.assert_allclose(out[1][1], out[0][1], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(out[2][1], out[0][1], rtol=1e-8, atol=1e-12)


def test_c4_expert_size_against_oracle():
    """C4: one expert, n = 4096, d = 20 (the CUDA-graph replay path): likelihood parity + batch == single."""
    import egobox_b200 as eg
    fast = _fast()
    n, d = 4096, 20
    x, y = _workload(n, d, 4)
    xn, xm, xs = O.normalize(x)
    yn, ym, ys = O.normalize(y.reshape(-1, 1))
    ctx = eg.GpContext(xn, yn[:, 0], xm, xs, float(ym[0]), float(ys[0]), eg.MATERN52, eg.CONSTANT)
    fx = O.mean_value(O.CONSTANT, xn)
    thetas = 10.0 ** np.random.default_rng(3).uniform(-0.7, 0.3, size=(7, d))
    st, rl = ctx.reduced_likelihood_batch(thetas)
    assert np.all(st == 0)
    for k in (0, 3, 6):
        ref, _ = fast.reduced_likelihood(O.MATERN52, xn, fx, yn, float(ys[0]), thetas[k], np.eye(d))
        assert rl[k] == pytest.approx(ref, rel=1e-9)
    st2, rl2 = ctx.reduced_likelihood_batch(thetas)                 # replayed graphs
    assert np.array_equal(rl, rl2)
    assert ctx.reduced_likelihood(thetas[2])[1] == rl[2]
    ctx.close()


def test_c3_sparse_fullsize_against_oracle():
    """C3: FITC, N = 100 000, d = 6, M = 1024 inducing points -- likelihood and predictions against the oracle
    (its K matrices built by the C/OpenMP kernel, the rest numpy + LAPACK)."""
    import egobox_b200 as eg
    N, d, M = 100_000, 6, 1024
    rng = np.random.default_rng(17)
    x = 2 * rng.random((N, d)) - 1
    y = np.sum(np.sin(3 * x), axis=1) + rng.normal(0, 0.1, N)
    z = S.make_inducings(M, x, rng)
    theta, sigma2, noise, nug = np.full(d, 1.5), 0.9, 0.02, 1e-8
    S.USE_FAST_KERNEL = True
    try:
        ref = S.build(S.FITC, O.MATERN52, theta, sigma2, noise, x, y, z, nugget=nug)
        xs = 2 * rng.random((1000, d)) - 1
        yo, vo = ref.predict(xs), ref.predict_var(xs)
    finally:
        S.USE_FAST_KERNEL = False
    ctx = eg.SgpContext(x, y, z, corr=O.MATERN52, method=S.FITC, nugget=nug)
    st, lik = ctx.reduced_likelihood(theta, sigma2, noise)
    assert st == 0
    assert lik == pytest.approx(ref.likelihood, rel=1e-8)
    st, res = ctx.finalize(theta, sigma2, noise)
    assert st == 0
    np.testing.assert_allclose(ctx.predict(xs), yo, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(ctx.predict_var(xs), vo, rtol=2e-5, atol=1e-8)
    ctx.close()
