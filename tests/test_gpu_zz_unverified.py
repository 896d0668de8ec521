"""-m gpu: device mirrors of reference tests -- automatic number of clusters (`test_moe_auto`, moe/src/algorithm.rs:1291-1311), the constant-function edge case
(gp/src/algorithm.rs:1217-1237), the cross-validation scores (gp/src/metrics.rs:117-150, moe/src/metrics.rs:239-261).  Their
host logic is covered on the CPU (tests/test_moe_host.py, tests/test_metrics_host.py, tests/test_host_optimizer.py) with
oracle stand-ins; here every fit is a device fit.  Sorted last on purpose."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# r01: written after that round's GPU minutes were spent and carried a non-strict xfail; all four XPASSED on the driver's B200
# (GPUTEST_r01.json), so the marker is gone: they are ordinary tests now and can fail.


def _unverified(f):
    return f


def _f_test_1d(x):
    x = x[:, 0]
    return np.where(x < 0.4, x * x, np.where(x < 0.8, 3.0 * x + 1.0, np.sin(10.0 * x)))


@_unverified
@pytest.mark.timeout(600)
def test_moe_auto():
    import egobox_b200 as egx
    xt = np.random.default_rng(42).random((60, 1))
    yt = _f_test_1d(xt)
    gpx = egx.Gpx.builder(n_clusters=0, seed=42).fit(xt, yt)
    assert 2 <= gpx.thetas().shape[0] <= 60 // 10 + 1               # the CPU stand-in of the same search finds 3
    assert gpx.predict(np.array([[0.37]])).item() == pytest.approx(0.37 * 0.37, abs=1e-3)      # algorithm.rs:1305-1309
    # bounded search: n_clusters = -2 looks at 1 and 2 clusters only (gp_mix.rs:200)
    gpx2 = egx.Gpx.builder(n_clusters=-2, seed=42).fit(xt, yt)
    assert 1 <= gpx2.thetas().shape[0] <= 2
    assert np.all(np.isfinite(gpx2.predict(xt)))


@_unverified
@pytest.mark.timeout(300)
def test_constant_function():
    """gp/src/algorithm.rs:1217-1237: y = 3.1 everywhere, KPLS with one component.  The PLS power method reports a constant
    residual -> all-zero rotations (:846-851) -> R = ones + nugget I, sigma2 = 0, likelihood +inf; the fit must still come
    back and predict the constant."""
    import egobox_b200 as egx
    from oracle import gp_oracle as O
    xt = O.lhs_classic(np.array([[0.0, 1.0]] * 3), 5, np.random.default_rng(42))
    yt = np.full(5, 3.1)
    gp = egx.GaussianProcess.params().theta_init([0.1]).kpls_dim(1).fit(xt, yt)
    xtest = O.lhs_classic(np.array([[0.0, 1.0]] * 3), 5, np.random.default_rng(43))
    np.testing.assert_allclose(gp.predict(xtest), np.full(5, 3.1), atol=1e-6)
    assert np.all(gp.predict_var(xtest) >= 0.0)
    gp.close()


@_unverified
@pytest.mark.timeout(600)
def test_q2_gp_griewank():
    """gp/src/metrics.rs:117-150: KPLS(3) kriging on 100 LHS points of the 5-D Griewank function; leave-one-out and
    10-fold Q2 within 1e-2 of 1 (100 + 10 refits on the device)."""
    import egobox_b200 as egx
    from oracle import gp_oracle as O
    dim, nt = 5, 100
    xt = O.lhs_classic(np.array([[-600.0, 600.0]] * dim), nt, np.random.default_rng(42))
    dd = np.sqrt(np.linspace(1.0, dim, dim))
    yt = (xt ** 2).sum(axis=1) / 4000.0 - np.prod(np.cos(xt / dd), axis=1) + 1.0
    gp = egx.GaussianProcess.params().kpls_dim(3).fit(xt, yt)
    assert gp.looq2_score() == pytest.approx(1.0, abs=1e-2)
    assert gp.q2_score(10) == pytest.approx(1.0, abs=1e-2)
    gp.close()


@_unverified
@pytest.mark.timeout(600)
def test_gpqa_x_squared():
    """moe/src/metrics.rs:239-261 (`test_gpqa_griewank`, which fits x -> sum x^2): 20 LHS points in [-10, 10]^2, default
    mixture (one cluster): Q2 (10-fold and leave-one-out) within 1e-3 of 1, predictive variance adequacy within 0.2 of 0."""
    from egobox_b200 import moe as E
    from oracle import gp_oracle as O
    xt = O.lhs_classic(np.array([[-10.0, 10.0]] * 2), 20, np.random.default_rng(42))
    yt = (xt ** 2).sum(axis=1)
    mix = E.GpMixtureParams().fit(xt, yt)
    assert mix.q2_k(10) == pytest.approx(1.0, abs=1e-3)
    assert mix.q2() == pytest.approx(1.0, abs=1e-3)
    assert mix.pva_k(10) == pytest.approx(0.0, abs=2e-1)
    assert mix.pva() == pytest.approx(0.0, abs=3e-1)         # reference: 2e-1 on ITS sample; the oracle gives 0.19 on this one
    assert 0.0 <= mix.iae_alpha_k(10) <= 0.5
    mix.close()
