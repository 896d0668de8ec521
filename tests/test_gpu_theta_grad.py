"""-m gpu: the closed-form theta gradient (egx_gp_reduced_likelihood_grad_analytic, kernels_thetagrad.cu) against the
oracle's closed form (oracle/gp_oracle.py::reduced_likelihood_grad, itself checked against central differences of the
oracle's rlf in tests/test_oracle_golden.py).  The reference has no theta gradient (algorithm.rs:880)."""
import numpy as np
import pytest

from oracle import gp_oracle as O
from tests.gpu_util import make_problem, make_context

pytestmark = pytest.mark.gpu

CORRS = [O.SQEXP, O.ABSEXP, O.MATERN32, O.MATERN52]


def _check(ctx, corr, mean, xn, yn, ys, theta, w, rtol):
    st, rlf, g = ctx.reduced_likelihood_grad_analytic(theta)
    assert st == 0
    fx = O.mean_value(mean, xn)
    rlf_o, g_o = O.reduced_likelihood_grad(corr, xn, fx, yn[:, 0], ys, theta, w)
    assert rlf == pytest.approx(rlf_o, rel=1e-9)
    scale = np.max(np.abs(g_o))
    assert np.max(np.abs(g - g_o)) <= rtol * scale, (g, g_o)
    return g


@pytest.mark.parametrize("corr", CORRS)
@pytest.mark.parametrize("n,d,mean", [(5, 1, O.CONSTANT), (150, 3, O.LINEAR), (333, 10, O.CONSTANT)])
def test_closed_form_gradient(corr, n, d, mean):
    x, y = make_problem(n, d, seed=n + 1)
    ctx, (xn, xm, xs, yn, ym, ys) = make_context(x, y, corr, mean)
    theta = np.linspace(0.6, 1.9, d)
    # bound: 1e-8 of the largest component (the two terms of the gradient cancel to ~1e-3 of their size here)
    _check(ctx, corr, mean, xn, yn, ys, theta, np.eye(d), 1e-8)
    ctx.close()


@pytest.mark.parametrize("corr", CORRS)
def test_closed_form_gradient_kpls_weights(corr):
    n, d, h = 200, 6, 2
    x, y = make_problem(n, d, seed=9)
    rng = np.random.default_rng(11)
    w = rng.normal(size=(d, h))
    w[2, 0] = 0.0
    ctx, (xn, xm, xs, yn, ym, ys) = make_context(x, y, corr, O.CONSTANT, w_star=w)
    _check(ctx, corr, O.CONSTANT, xn, yn, ys, np.array([0.7, 1.4]), w, 1e-8)
    ctx.close()


def test_closed_form_gradient_matches_central_differences_and_is_repeatable():
    n, d = 150, 3
    x, y = make_problem(n, d, seed=77)
    ctx, _ = make_context(x, y, O.MATERN52, O.CONSTANT)
    theta = np.array([1.5, 0.9, 2.2])
    st, rlf, g = ctx.reduced_likelihood_grad_analytic(theta)
    st2, rlf2, g2 = ctx.reduced_likelihood_grad_analytic(theta)
    assert st == 0 and st2 == 0
    assert rlf == rlf2 and np.array_equal(g, g2)                 # fixed summation order
    st, rlf_fd, g_fd = ctx.reduced_likelihood_grad(theta, rel_step=1e-5)
    assert st == 0 and rlf == pytest.approx(rlf_fd, rel=1e-10)
    np.testing.assert_allclose(g, g_fd, rtol=1e-4, atol=1e-6)
    # the context still evaluates and finalises normally afterwards
    st, v = ctx.reduced_likelihood(theta)
    assert st == 0 and v == pytest.approx(rlf, rel=1e-10)
    ctx.close()


def test_closed_form_gradient_on_the_tcgen05_path():
    """n = 1664 (13 block columns): the sweep on the identity and the W W^T triangles of 8+ tile rows run on tcgen05."""
    n, d = 1664, 4
    x, y = make_problem(n, d, seed=5)
    ctx, (xn, xm, xs, yn, ym, ys) = make_context(x, y, O.MATERN52, O.CONSTANT)
    theta = np.array([0.8, 1.1, 0.6, 1.4])
    ctx.set_profiling(True)
    ctx.reset_profile()
    _check(ctx, O.MATERN52, O.CONSTANT, xn, yn, ys, theta, np.eye(d), 1e-7)
    prof = ctx.profile()
    assert prof["theta_grad"][1] == 2
    assert prof["ozaki_syrk"][1] > 0
    ctx.close()


def test_closed_form_gradient_rejects_more_than_32_components():
    n, d = 40, 33
    x, y = make_problem(n, d, seed=2)
    ctx, _ = make_context(x, y, O.SQEXP, O.CONSTANT)
    from egobox_b200._lib import GpuError
    with pytest.raises(GpuError) as e:
        ctx.reduced_likelihood_grad_analytic(np.full(d, 0.5))
    assert "32 components" in str(e.value)
    ctx.close()


def test_fit_with_the_gradient_based_multistart():
    """optimizer("lbfgsb"): same starts and budget as the COBYLA chains.  Checker: the same host
    minimiser driven by the ORACLE's closed-form gradient from the same starts -- the device only supplies (rlf, gradient)."""
    import math
    import egobox_b200 as egx
    from egobox_b200 import gp as G
    n, d = 200, 4
    x, y = make_problem(n, d, seed=9)
    base = lambda: egx.GaussianProcess.params().corr(O.MATERN32).n_start(4)
    gp = base().optimizer("lbfgsb").fit(x, y)
    gp_cobyla = base().fit(x, y)

    xn, _, _ = O.normalize(x)
    yn, _, ys = O.normalize(y.reshape(-1, 1))
    fx = O.mean_value(O.CONSTANT, xn)

    def fg(z):
        th = 10.0 ** z
        try:
            rlf, g = O.reduced_likelihood_grad(O.MATERN32, xn, fx, yn[:, 0], float(ys[0]), th, np.eye(d))
        except (O.LinalgError, O.LikelihoodComputationError):
            return math.inf, np.zeros(d)
        return -rlf, -math.log(10.0) * th * g

    starts = G.prepare_multistart(4, np.full(d, 0.1), [(1e-2, 1e1)] * d, seed=42)
    best_f, best_z, evals = math.inf, None, 0
    for s0 in starts:
        z, f, nev = G.bound_lbfgs_minimize(fg, s0, [(-2.0, 1.0)] * d, ftol_rel=1e-9, gtol=1e-7, maxeval=40)
        evals += nev
        if f < best_f:
            best_f, best_z = f, z
    assert gp.likelihood() == pytest.approx(-best_f, rel=1e-6)
    np.testing.assert_allclose(gp.theta(), 10.0 ** best_z, rtol=5e-3)
    assert gp.n_evals() <= 5 * 40 + 1
    # likelihood at the returned theta is the oracle's, and the optimum is at least as good as the derivative-free one
    rlf_o, _ = O.reduced_likelihood(O.MATERN32, xn, fx, yn[:, 0], float(ys[0]), gp.theta(), np.eye(d))
    assert gp.likelihood() == pytest.approx(rlf_o, rel=1e-9)
    assert gp.likelihood() >= gp_cobyla.likelihood() - 1e-6 * abs(gp_cobyla.likelihood())
    xs = np.random.default_rng(1).random((50, d))
    assert np.all(np.isfinite(gp.predict(xs))) and np.all(gp.predict_var(xs) >= 0.0)
    gp.close()
    gp_cobyla.close()
