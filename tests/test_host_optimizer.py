"""CPU tests of the host-side fit machinery (C++ in libegobox_gpu.so, no GPU needed):
the bound-constrained COBYLA-family optimiser and the multistart seeds
(gp/src/optimization.rs:26-71, 122-169).  The kriging objective here is the ORACLE's
(checker), so these tests also pin the optimiser on the reference's published optimum."""
import json
import math
import os

import numpy as np
import pytest

from egobox_b200 import gp as G
from oracle import gp_oracle as O


def test_quadratic_in_box():
    f = lambda z: (z[0] - 0.3) ** 2 + 2.0 * (z[1] + 0.2) ** 2 + 1.0
    x, fv, nev = G.bound_cobyla_minimize(f, [-0.9, 0.9], [(-1, 1), (-1, 1)], ftol_rel=0.0, maxeval=200)
    assert fv == pytest.approx(1.0, abs=1e-8)
    np.testing.assert_allclose(x, [0.3, -0.2], atol=1e-4)
    assert nev <= 200


def test_active_bound():
    f = lambda z: (z[0] - 2.0) ** 2 + (z[1] - 0.5) ** 2
    x, fv, _ = G.bound_cobyla_minimize(f, [0.0, 0.0], [(-1, 1), (-1, 1)], ftol_rel=0.0, maxeval=150)
    np.testing.assert_allclose(x, [1.0, 0.5], atol=1e-4)
    assert fv == pytest.approx(1.0, abs=1e-6)


def test_budget_and_inf_values():
    calls = []

    def f(z):
        calls.append(z.copy())
        return math.inf if z[0] < -0.5 else (z[0] - 0.1) ** 2 + (z[1] + 0.4) ** 2 + (z[2] - 0.2) ** 2

    x, fv, nev = G.bound_cobyla_minimize(f, [-0.9, 0.0, 0.0], [(-1, 1)] * 3, ftol_rel=0.0, maxeval=60)
    assert nev == len(calls) <= 60
    assert all(np.all(np.abs(c) <= 1.0 + 1e-15) for c in calls)     # never leaves the box
    assert fv < 1e-3


def test_rosenbrock_5d_budget():
    def f(z):
        return float(np.sum(100.0 * (z[1:] - z[:-1] ** 2) ** 2 + (1 - z[:-1]) ** 2))
    # linear-model methods crawl along the Rosenbrock valley: hold ours to the same league as
    # scipy's COBYLA (PRIMA) with the same radius and budget
    from scipy.optimize import minimize
    x, fv, nev = G.bound_cobyla_minimize(f, np.zeros(5), [(-2, 2)] * 5, ftol_rel=0.0, maxeval=500)
    ref = minimize(f, np.zeros(5), method="COBYLA", bounds=[(-2, 2)] * 5,
                   options=dict(rhobeg=0.5, maxiter=500, tol=1e-10))
    assert nev == 500
    assert fv < 2.5 * ref.fun


def test_prepare_multistart_layout():
    s = G.prepare_multistart(10, [0.1, 0.1, 0.1], [(1e-2, 10.0)])
    assert s.shape == (11, 3)
    np.testing.assert_allclose(s[0], -1.0)
    assert np.all(s[1:] >= -2.0) and np.all(s[1:] <= 1.0)
    # LHS: exactly one point per stratum in every column
    for j in range(3):
        strata = np.floor((s[1:, j] + 2.0) / 3.0 * 10).astype(int)
        assert sorted(strata) == list(range(10))
    np.testing.assert_array_equal(s, G.prepare_multistart(10, [0.1, 0.1, 0.1], [(1e-2, 10.0)]))   # seeded


def test_kriging5_optimum_with_reference_settings(golden_dir):
    """doc/Gpx_Tutorial.ipynb:165-167: theta* = 1.83209405, rlf = 0.5781740714613353."""
    with open(os.path.join(golden_dir, "gpx_tutorial_kriging5.json")) as f:
        k = json.load(f)
    x = np.array(k["xt"])[:, None]
    y = np.array(k["yt"])
    xn, _, _ = O.normalize(x)
    yn, _, ys = O.normalize(y.reshape(-1, 1))
    fx = O.mean_value(O.CONSTANT, xn)
    obj = lambda z: O.objective(O.SQEXP, xn, fx, yn, float(ys[0]), 10.0 ** z, np.eye(1))
    starts = G.prepare_multistart(10, [0.1], [(1e-2, 10.0)])
    best = (math.inf, None)
    for s in starts:
        z, fv, nev = G.bound_cobyla_minimize(obj, s, [(-2.0, 1.0)], rhobeg=0.5, ftol_rel=1e-4, maxeval=25)
        assert nev <= 25
        if fv < best[0]:
            best = (fv, z)
    assert -best[0] == pytest.approx(k["likelihood"], rel=1e-6)
    assert 10.0 ** best[1][0] == pytest.approx(k["theta"], rel=5e-3)


# ------------------------------------------------------ host eigen-decomposition (sampler) --------
@pytest.mark.parametrize("n", [1, 2, 7, 40, 150])
def test_symmetric_eig_against_numpy(n):
    """egx_symmetric_eig is the `cov_x.eigh()` of gp/src/algorithm.rs:1171-1173 (host in the reference too)."""
    rng = np.random.default_rng(n)
    b = rng.normal(size=(n, n))
    a = b.dot(b.T) + np.diag(rng.random(n))
    if n >= 7:                                       # rank-deficient block, as conditional covariances are
        a[:, -3:] = a[:, :3]
        a[-3:, :] = a[:3, :]
        a = 0.5 * (a + a.T)
    w, v = G.symmetric_eig(a)
    scale = np.abs(a).max()
    np.testing.assert_allclose(v.T.dot(v), np.eye(n), atol=1e-12)
    np.testing.assert_allclose(v.dot(np.diag(w)).dot(v.T), a, atol=1e-12 * scale * n)
    np.testing.assert_allclose(np.sort(w), np.linalg.eigvalsh(a), atol=1e-12 * scale * n)


def test_oracle_covariance_consistent_with_predict_var():
    """The oracle's conditional covariance (algorithm.rs:310-326) has predict_var on its diagonal and the
    sampler reproduces it: C C^T = cov for both decompositions (eigenvalues below 1e-9 dropped)."""
    rng = np.random.default_rng(5)
    x = rng.random((40, 2))
    y = np.sin(4 * x[:, 0]) + x[:, 1] ** 2
    gp = O.fit(x, y, corr=O.MATERN52, mean=O.CONSTANT, theta_init=np.array([1.0, 1.0]), fixed=True)
    xs = rng.random((15, 2))
    cov = gp.compute_covariance(xs)
    np.testing.assert_allclose(np.diag(cov), gp.predict_var(xs), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(cov, cov.T, atol=1e-12)
    mean = gp.predict(xs)[:, None]
    for method in ("chol", "eig"):
        c = gp.sample(xs, np.eye(15), method) - mean
        np.testing.assert_allclose(c.dot(c.T), cov, atol=2e-9 * 15 + 1e-9 * np.abs(cov).max())


# --------------------------------------------------------------------------- projected L-BFGS (optimizer "lbfgsb", not in the reference)
def _rosen(x):
    f = np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1.0 - x[:-1]) ** 2)
    g = np.zeros_like(x)
    g[:-1] = -400.0 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2.0 * (1.0 - x[:-1])
    g[1:] += 200.0 * (x[1:] - x[:-1] ** 2)
    return f, g


def test_lbfgs_rosenbrock_against_scipy():
    from scipy.optimize import minimize
    x0 = np.array([-1.2, 1.0, -0.5, 0.8, 1.5])
    bounds = [(-2.0, 2.0)] * 5
    x, f, nev = G.bound_lbfgs_minimize(_rosen, x0, bounds, ftol_rel=1e-14, gtol=1e-9, maxeval=2000)
    ref = minimize(_rosen, x0, jac=True, method="L-BFGS-B", bounds=bounds, options={"ftol": 1e-15, "gtol": 1e-10})
    assert f <= 1e-10 and ref.fun <= 1e-10
    np.testing.assert_allclose(x, ref.x, atol=1e-4)
    np.testing.assert_allclose(x, np.ones(5), atol=1e-4)
    assert nev < 400


def test_lbfgs_active_bounds_against_scipy():
    from scipy.optimize import minimize
    c = np.array([2.0, -1.0, 0.3])

    def fg(x):
        return float(np.sum((x - c) ** 2) + 0.5 * x[0] * x[2]), 2.0 * (x - c) + 0.5 * np.array([x[2], 0.0, x[0]])

    bounds = [(0.0, 1.0)] * 3
    x, f, nev = G.bound_lbfgs_minimize(fg, np.full(3, 0.5), bounds, ftol_rel=1e-15, gtol=1e-10, maxeval=200)
    ref = minimize(fg, np.full(3, 0.5), jac=True, method="L-BFGS-B", bounds=bounds, options={"ftol": 1e-15, "gtol": 1e-12})
    np.testing.assert_allclose(x, ref.x, atol=1e-6)
    assert x[0] == 1.0 and x[1] == 0.0                       # both on their bounds, exactly
    assert f == pytest.approx(ref.fun, rel=1e-10)


def test_lbfgs_budget_inf_and_start_outside_the_box():
    calls = []

    def fg(x):
        calls.append(x.copy())
        if x[0] > 0.8:                                       # Err(_) -> +inf region (algorithm.rs:893-896)
            return math.inf, np.zeros(2)
        return float((x[0] - 1.0) ** 2 + (x[1] - 0.2) ** 2), np.array([2.0 * (x[0] - 1.0), 2.0 * (x[1] - 0.2)])

    x, f, nev = G.bound_lbfgs_minimize(fg, np.array([-3.0, 5.0]), [(0.0, 1.0), (0.0, 1.0)], maxeval=60)
    assert np.all(calls[0] == [0.0, 1.0])                    # start projected into the box
    assert nev == len(calls) <= 60
    # the wall is a discontinuity: a line-search method stops against it, short of the constrained optimum (0.8, 0.2)
    assert 0.7 <= x[0] <= 0.8 and math.isfinite(f) and f < fg(np.array([0.0, 1.0]))[0] and f == fg(x)[0]
    # a start where the objective is +inf (a failed likelihood) probes four points towards the centre of the box before it
    # gives up and reports +inf ...
    x, f, nev = G.bound_lbfgs_minimize(lambda z: (math.inf, np.zeros(1)), np.array([0.5]), [(0.0, 1.0)])
    assert nev == 5 and math.isinf(f)
    # ... and carries on from the first finite one: f = +inf for z > 0.8, (z - 0.3)^2 below; start at 1.0, centre 0.5
    wall = lambda z: (math.inf, np.zeros(1)) if z[0] > 0.8 else ((z[0] - 0.3) ** 2, np.array([2.0 * (z[0] - 0.3)]))   # noqa: E731
    x, f, nev = G.bound_lbfgs_minimize(wall, np.array([1.0]), [(0.0, 1.0)], maxeval=40)
    assert math.isfinite(f) and abs(x[0] - 0.3) < 1e-4


def test_multistart_reaches_the_likelihood_of_powells_cobyla():
    """The reference optimises with the third-party `cobyla 0.8.0` crate (Powell's COBYLA, optimization.rs:122-169), which is
    not under /root/reference; the chain optimiser here is of the same family, not a transcription, so its TRAJECTORY differs
    from Powell's and -- at the reference's tiny budget of clamp(10 d, 25, 1000) evaluations per start, where neither has
    converged -- so does the best likelihood over the multistart.  With the reference's own starts (Lhs Maximin, seed 42:
    tests/test_host_rng.py), rhobeg 0.5 and ftol_rel 1e-4 the two agree to ~1 % in the worst case and to 0.2 % on
    average over a set of problems, each winning some (measured r02: -3.4e-3 .. +1.1e-2, mean +1.3e-3 relative)."""
    from scipy.optimize import minimize
    from tests.gpu_util import make_problem
    cases = [(60, 2, O.SQEXP, 1), (80, 3, O.MATERN52, 2), (90, 5, O.ABSEXP, 5), (70, 3, O.SQEXP, 3), (100, 4, O.MATERN52, 4),
             (120, 6, O.MATERN32, 6), (80, 3, O.MATERN52, 7), (150, 8, O.MATERN52, 9)]
    rel = []
    for n, d, corr, seed in cases:
        x, y = make_problem(n, d, seed=seed)
        xn, _, _ = O.normalize(x)
        yn, _, ys = O.normalize(y.reshape(-1, 1))
        fx = O.mean_value(O.CONSTANT, xn)
        obj = lambda z: O.objective(corr, xn, fx, yn[:, 0], float(ys[0]), 10.0 ** np.asarray(z), np.eye(d))   # noqa: E731
        starts = G.prepare_multistart(10, np.full(d, 0.1), [(1e-2, 1e1)] * d, seed=42)
        bounds = [(-2.0, 1.0)] * d
        maxeval = min(max(10 * d, 25), 1000)
        mine = min(G.bound_cobyla_minimize(obj, s0, bounds, rhobeg=0.5, ftol_rel=1e-4, maxeval=maxeval)[1] for s0 in starts)
        powell = min(minimize(obj, s0, method="COBYLA", bounds=bounds,
                              options={"rhobeg": 0.5, "maxiter": maxeval, "tol": 1e-4}).fun for s0 in starts)
        rel.append((mine - powell) / abs(powell))
    rel = np.array(rel)
    assert rel.max() <= 2e-2, rel            # never more than 2 % behind Powell's optimiser ...
    assert rel.mean() <= 4e-3, rel           # ... 0.4 % on average ...
    assert rel.min() < 0.0, rel              # ... and ahead of it on some problems


def test_q2_score_fold_logic_against_a_direct_computation():
    """gp/src/metrics.rs:35-58 (`q2_score`, `looq2_score`): the fold arithmetic is host logic; the refits are oracle kriging
    models here (on the device they are `params.fit`)."""
    from egobox_b200.gp import _q2_score
    rng = np.random.default_rng(0)
    x = rng.random((23, 2))
    y = np.sin(3.0 * x[:, 0]) + x[:, 1] ** 2
    fit = lambda xs, ys: O.fit(xs, ys, corr=O.SQEXP, mean=O.CONSTANT, theta_init=[1.0], fixed=True)
    for kfold in (4, 23):
        fs = 23 // kfold
        press = tss = 0.0
        for i in range(kfold):
            va = list(range(i * fs, (i + 1) * fs))
            tr = [j for j in range(23) if j not in va]               # the remainder rows (n % kfold) always train
            pred = fit(x[tr], y[tr]).predict(x[va])
            press += ((y[va] - pred) ** 2).sum()
            tss += ((y[va] - y.mean()) ** 2).sum()
        assert _q2_score((x, y), kfold, fit) == pytest.approx(1.0 - press / tss, rel=1e-13)
    assert _q2_score((x, y), 23, fit) > 0.9                           # leave-one-out on a smooth function
    with pytest.raises(G.InvalidValueError):
        _q2_score((x, y), 24, fit)
