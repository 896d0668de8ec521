"""Pins the CPU oracle against every tight fixture the reference holds for the path.

Fixture values below are transcribed verbatim from the reference's own tests /
stored notebook outputs (file:line cited per test)."""
import json
import math
import os

import numpy as np
import pytest

from oracle import gp_oracle as O


# ---- correlation_models.rs:597-641, 718-726 --------------------------------
def test_squared_exponential_1d():
    xt = np.array([[4.5], [1.2], [2.0], [3.0], [4.0]])
    d, _ = O.diff_matrix(xt)
    res = O.corr_value(O.SQEXP, d, [math.sqrt(0.2)], np.array([[1.0]]))
    expected = [0.336552878364737, 0.5352614285189903, 0.7985162187593771, 0.9753099120283326,
                0.9380049995307295, 0.7232502423798424, 0.4565760496233148, 0.9048374180359595,
                0.6703200460356393, 0.9048374180359595]
    np.testing.assert_allclose(res, expected, atol=1e-12)


def test_squared_exponential_2d():
    xt = np.array([[0.0, 1.0], [2.0, 3.0], [4.0, 5.0]])
    d, _ = O.diff_matrix(xt)
    res = O.corr_value(O.SQEXP, d, [math.sqrt(2.0), 2.0], np.eye(2))
    np.testing.assert_allclose(res, [6.14421235e-06, 1.42516408e-21, 6.14421235e-06], atol=1e-6, rtol=1e-8)


def test_matern32_2d():
    xt = np.array([[0.0, 1.0], [2.0, 3.0], [4.0, 5.0]])
    d, _ = O.diff_matrix(xt)
    res = O.corr_value(O.MATERN32, d, [1.0, 2.0], np.eye(2))
    np.testing.assert_allclose(res, [1.08539595e-03, 1.10776401e-07, 1.08539595e-03], atol=1e-6, rtol=1e-8)


def test_matern52_2d():
    xt = np.array([[0.0, 1.0], [2.0, 3.0], [4.0, 5.0]])
    d, _ = O.diff_matrix(xt)
    res = O.corr_value(O.MATERN52, d, [1.0, 2.0], np.eye(2))
    np.testing.assert_allclose(res, [6.62391590e-04, 1.02117882e-08, 6.62391590e-04], atol=1e-6, rtol=1e-8)


# ---- utils.rs:150-242 -------------------------------------------------------
def test_pairwise_differences():
    x = np.array([[-0.9486833], [-0.82219219]])
    y = np.array([[-1.26491106], [-0.63245553], [0.0], [0.63245553], [1.26491106]])
    exp = [0.31622777, -0.31622777, -0.9486833, -1.58113883, -2.21359436,
           0.44271887, -0.18973666, -0.82219219, -1.45464772, -2.08710326]
    np.testing.assert_allclose(O.pairwise_differences(x, y)[:, 0], exp, atol=1e-6)


def test_normalized_matrix():
    xn, mean, std = O.normalize(np.array([[1.0, 2.0], [3.0, 4.0]]))
    assert list(mean) == [2.0, 3.0]
    assert list(std) == [math.sqrt(2.0), math.sqrt(2.0)]


def test_normalize_constant_column():
    xn, mean, std = O.normalize(np.array([[1.0, 2.0], [1.0, 4.0]]))
    assert std[0] == 1.0 and np.all(xn[:, 0] == 0.0)


def test_diff_matrix():
    xt = np.array([[0.5], [1.2], [2.0], [3.0], [4.0]])
    d, idx = O.diff_matrix(xt)
    np.testing.assert_allclose(d[:, 0], [0.7, 1.5, 2.5, 3.5, 0.8, 1.8, 2.8, 1.0, 2.0, 1.0], atol=1e-15)
    assert idx.tolist() == [[0, 1], [0, 2], [0, 3], [0, 4], [1, 2], [1, 3], [1, 4], [2, 3], [2, 4], [3, 4]]


# ---- mean_models.rs:169-213 -------------------------------------------------
def test_quadratic_basis():
    a = np.array([[1.0, 2.0, 3.0], [3.0, 4.0, 5.0]])
    exp = np.array([[1.0, 1.0, 2.0, 3.0, 1.0, 2.0, 3.0, 4.0, 6.0, 9.0],
                    [1.0, 3.0, 4.0, 5.0, 9.0, 12.0, 15.0, 16.0, 20.0, 25.0]])
    np.testing.assert_array_equal(O.mean_value(O.QUADRATIC, a), exp)
    assert O.mean_value(O.LINEAR, a).shape == (2, 4)
    assert O.mean_value(O.CONSTANT, a).shape == (2, 1)
    assert O.mean_nbasis(O.QUADRATIC, 3) == 10


# ---- doc/Gpx_Tutorial.ipynb:420-421 : the 16-digit full model ----------------
def _arr(o):
    return np.array(o["data"], dtype=np.float64).reshape(o["dim"])


@pytest.fixture(scope="module")
def nb_model(golden_dir):
    with open(os.path.join(golden_dir, "gpx_tutorial_linear_matern52.json")) as f:
        return json.load(f)


def test_notebook_full_model(nb_model):
    m = nb_model
    xt, yt = _arr(m["training_data"][0]), _arr(m["training_data"][1])
    theta = _arr(m["theta"])
    gp = O.fit(xt, yt, corr=O.MATERN52, mean=O.LINEAR, theta_init=theta, fixed=True,
               nugget=m["params"]["nugget"])
    ip = m["inner_params"]
    np.testing.assert_allclose(gp.xt_norm, _arr(m["xt_norm"]["data"]), rtol=0, atol=5e-16)
    np.testing.assert_allclose(gp.x_std, _arr(m["xt_norm"]["std"]), rtol=1e-15)
    np.testing.assert_allclose(gp.yt_norm, _arr(m["yt_norm"]["data"]), rtol=0, atol=5e-16)
    assert gp.likelihood == pytest.approx(m["likelihood"], rel=1e-13)
    assert gp.inner.sigma2 == pytest.approx(ip["sigma2"], rel=1e-12)
    np.testing.assert_allclose(gp.inner.r_chol, _arr(ip["r_chol"]), rtol=0, atol=1e-15)
    np.testing.assert_allclose(gp.inner.ft, _arr(ip["ft"]), rtol=0, atol=2e-15)
    np.testing.assert_allclose(gp.inner.ft_qr_r, _arr(ip["ft_qr_r"]), rtol=0, atol=2e-15)
    np.testing.assert_allclose(gp.inner.beta, _arr(ip["beta"]), rtol=0, atol=1e-14)
    np.testing.assert_allclose(gp.inner.gamma, _arr(ip["gamma"]), rtol=0, atol=1e-14)


# ---- doc/Gpx_Tutorial.ipynb:165-167 + python/egobox/tests/test_gpmix.py:37-53 ----
@pytest.fixture(scope="module")
def krg5(golden_dir):
    with open(os.path.join(golden_dir, "gpx_tutorial_kriging5.json")) as f:
        return json.load(f)


def test_kriging5_at_published_theta(krg5):
    xt = np.array(krg5["xt"])[:, None]
    gp = O.fit(xt, krg5["yt"], theta_init=[krg5["theta"]], fixed=True)
    # printed theta has 9 significant digits: likelihood is flat at the optimum
    assert gp.likelihood == pytest.approx(krg5["likelihood"], rel=1e-9)
    assert gp.inner.sigma2 == pytest.approx(krg5["variance"], rel=2e-8)
    # test_gpmix.py:37-46
    assert gp.predict(np.array([[1.0]]))[0] == pytest.approx(1.0, abs=1e-7)
    assert gp.predict_var(np.array([[1.0]]))[0] == pytest.approx(0.0, abs=1e-7)
    assert gp.predict(np.array([[1.1]]))[0] == pytest.approx(1.1163, abs=1e-3)
    assert gp.predict_var(np.array([[1.1]]))[0] == pytest.approx(0.0, abs=1e-3)


def test_kriging5_optimised(krg5):
    xt = np.array(krg5["xt"])[:, None]
    gp = O.fit(xt, krg5["yt"])              # default multistart COBYLA
    assert gp.theta[0] == pytest.approx(krg5["theta"], rel=2e-3)
    assert gp.likelihood == pytest.approx(krg5["likelihood"], rel=1e-6)


def test_fixed_theta_no_optim(krg5):
    # test_gpmix.py:137-142
    xt = np.array(krg5["xt"])[:, None]
    gp = O.fit(xt, krg5["yt"], theta_init=[0.314], fixed=True)
    assert gp.theta[0] == 0.314


def test_valvar_matches_val_and_var(krg5):
    # moe/src/algorithm.rs:1549-1552
    xt = np.array(krg5["xt"])[:, None]
    gp = O.fit(xt, krg5["yt"], theta_init=[1.0], fixed=True)
    x = np.linspace(-1, 5, 17)[:, None]
    y, v = gp.predict_valvar(x)
    np.testing.assert_allclose(y, gp.predict(x), rtol=0, atol=1e-12)
    np.testing.assert_allclose(v, gp.predict_var(x), rtol=0, atol=1e-12)


def test_non_pd_maps_to_inf():
    # duplicated rows + nugget 0 -> singular R -> LinalgError -> +inf objective
    x = np.array([[0.0], [0.0], [1.0], [2.0]])
    y = np.array([0.0, 0.0, 1.0, 2.0])
    xn, _, _ = O.normalize(x)
    yn, _, ys = O.normalize(y.reshape(-1, 1))
    fx = O.mean_value(O.CONSTANT, xn)
    v = O.objective(O.SQEXP, xn, fx, yn, float(ys[0]), [1.0], np.eye(1), nugget=0.0)
    assert v == math.inf
    assert O.objective(O.SQEXP, xn, fx, yn, float(ys[0]), [float("nan")], np.eye(1)) == math.inf


def test_corr_matrix_equals_scatter():
    rng = np.random.default_rng(0)
    x = rng.random((23, 3))
    theta = np.array([0.3, 1.2, 2.0])
    for kind in (O.SQEXP, O.ABSEXP, O.MATERN32, O.MATERN52):
        d, idx = O.diff_matrix(x)
        r = O.corr_value(kind, d, theta, np.eye(3))
        R = np.eye(23) * (1.0 + O.DEFAULT_NUGGET)
        R[idx[:, 0], idx[:, 1]] = r
        R[idx[:, 1], idx[:, 0]] = r
        np.testing.assert_allclose(O.corr_matrix(kind, x, theta, np.eye(3)), R, rtol=1e-12, atol=0)


def test_c_restatement_matches_numpy():
    """oracle/corr_oracle.c (OpenMP, the CPU-baseline kernel) against the numpy restatement."""
    from oracle import fast
    rng = np.random.default_rng(9)
    x = rng.normal(size=(57, 4))
    xs = rng.normal(size=(13, 4))
    for w in (np.eye(4), rng.normal(size=(4, 2))):
        theta = np.array([0.4, 1.1, 0.7, 2.0])[: w.shape[1]]
        for kind in (O.SQEXP, O.ABSEXP, O.MATERN32, O.MATERN52):
            np.testing.assert_allclose(fast.corr_matrix(kind, x, theta, w), O.corr_matrix(kind, x, theta, w),
                                       rtol=1e-12, atol=0)
            ref = O.corr_value(kind, O.pairwise_differences(xs, x), theta, w).reshape(13, 57)
            np.testing.assert_allclose(fast.cross_corr(kind, xs, x, theta, w), ref, rtol=1e-12, atol=0)


# ---- correlation_models.rs:643-716: analytic jacobians vs central finite differences -------
XT16 = np.array([[-9.375, -5.625], [-5.625, -4.375], [9.375, 1.875], [8.125, 5.625], [-4.375, -0.625],
                 [6.875, -3.125], [4.375, 9.375], [3.125, 4.375], [5.625, -8.125], [-8.125, 3.125],
                 [1.875, -6.875], [-0.625, 8.125], [-1.875, -1.875], [0.625, 0.625], [-6.875, -9.375],
                 [-3.125, 6.875]])


@pytest.mark.parametrize("kpls", [False, True])
@pytest.mark.parametrize("kind", [O.SQEXP, O.ABSEXP, O.MATERN32, O.MATERN52])
def test_corr_jacobian_vs_finite_differences(kind, kpls):
    x = np.array([3.0, 5.0])
    xn_t, mean, std = O.normalize(XT16)
    if kpls:
        theta, w = np.array([0.31059002]), np.array([[-0.02701716], [-0.99963497]])
    else:
        theta, w = np.array([0.34599115925909146, 0.32083374253611624]), np.eye(2)
    jac = O.corr_jacobian(kind, (x - mean) / std, xn_t, theta, w) / std
    e = 1e-5
    for k in range(2):
        xp, xm = x.copy(), x.copy()
        xp[k] += e
        xm[k] -= e
        rp = O.corr_value(kind, (xp - mean) / std - xn_t, theta, w)
        rm = O.corr_value(kind, (xm - mean) / std - xn_t, theta, w)
        np.testing.assert_allclose((rp - rm) / (2 * e), jac[:, k], atol=1e-6)


def test_kriging5_predict_gradients(krg5):
    # python/egobox/tests/test_gpmix.py:48-50: predict_gradients(1.1) = 1.1204 (+- 1e-3)
    xt = np.array(krg5["xt"])[:, None]
    gp = O.fit(xt, krg5["yt"], theta_init=[krg5["theta"]], fixed=True)
    g = gp.predict_gradients(np.array([[1.1]]))
    assert g.shape == (1, 1) and g[0, 0] == pytest.approx(1.1204, abs=1e-3)
    e = 1e-6
    fd = (gp.predict(np.array([[1.1 + e]])) - gp.predict(np.array([[1.1 - e]]))) / (2 * e)
    assert g[0, 0] == pytest.approx(fd[0], rel=1e-6)


def test_quadratic_mean_jacobian():
    # mean_models.rs:188-213
    x = np.array([1.0, 2.0, 3.0])
    jac = O.mean_jacobian(O.QUADRATIC, x)
    e = 1e-6
    for k in range(3):
        xp, xm = x.copy(), x.copy()
        xp[k] += e
        xm[k] -= e
        fd = (O.mean_value(O.QUADRATIC, xp[None])[0] - O.mean_value(O.QUADRATIC, xm[None])[0]) / (2 * e)
        np.testing.assert_allclose(jac[:, k], fd, atol=1e-8)


def test_bug_var_derivatives_fixture():
    """gp/src/algorithm.rs:1723-1786: fixed theta, fixed 12-point data, analytic variance gradient vs
    central finite differences of predict_var to 1e-5."""
    xt = np.array([[6.875, -4.375], [-3.125, 1.875], [1.875, -1.875], [-4.375, 3.125], [8.125, 9.375],
                   [4.375, 4.375], [0.625, 0.625], [9.375, 6.875], [5.625, 8.125], [-0.625, -3.125],
                   [3.125, 5.625], [-1.875, -0.625]])
    yt = np.array([2.43286801, 13.10840811, 5.32908578, 17.81862219, 74.08849877, 39.68137781, 14.96009727,
                   63.17475741, 61.26331775, -7.46009727, 44.39159189, 2.17091422])
    theta = [np.sqrt(2 * 0.0437386), np.sqrt(2 * 0.00697978)]
    gp = O.fit(xt, yt, corr=O.SQEXP, mean=O.CONSTANT, theta_init=theta, fixed=True)
    e, xa, xb = 5e-6, -1.3, 2.5
    xq = np.array([[xa, xb], [xa + e, xb], [xa - e, xb], [xa, xb + e], [xa, xb - e]])
    v = gp.predict_var(xq)
    g = gp.predict_var_gradients(np.array([[xa, xb]]))
    assert g[0, 0] == pytest.approx((v[1] - v[2]) / (2 * e), abs=1e-5)
    assert g[0, 1] == pytest.approx((v[3] - v[4]) / (2 * e), abs=1e-5)


def test_kriging5_predict_var_gradients(krg5):
    # python/egobox/tests/test_gpmix.py:51-53: predict_var_gradients(1.1) = 0.0145 (+- 1e-3)
    xt = np.array(krg5["xt"])[:, None]
    gp = O.fit(xt, krg5["yt"], theta_init=[krg5["theta"]], fixed=True)
    g = gp.predict_var_gradients(np.array([[1.1]]))
    assert g[0, 0] == pytest.approx(0.0145, abs=1e-3)


@pytest.mark.parametrize("corr,mean", [(O.MATERN52, O.LINEAR), (O.MATERN32, O.QUADRATIC), (O.ABSEXP, O.CONSTANT)])
def test_var_gradients_vs_fd_general(corr, mean):
    rng = np.random.default_rng(4)
    x = rng.random((40, 2)) * 3
    y = np.sin(x[:, 0]) + x[:, 1] ** 2
    gp = O.fit(x, y, corr=corr, mean=mean, theta_init=[0.8, 1.3], fixed=True)
    xq = np.array([[1.234, 0.777]])
    g = gp.predict_var_gradients(xq)
    e = 1e-6
    for k in range(2):
        xp, xm = xq.copy(), xq.copy()
        xp[0, k] += e
        xm[0, k] -= e
        fd = (gp.predict_var(xp)[0] - gp.predict_var(xm)[0]) / (2 * e)
        assert g[0, k] == pytest.approx(fd, rel=1e-4, abs=1e-7)


# --------------------------------------------------------------------------- closed-form theta gradient (not in the reference)
@pytest.mark.parametrize("kind", [O.SQEXP, O.ABSEXP, O.MATERN32, O.MATERN52])
@pytest.mark.parametrize("kpls", [False, True])
def test_oracle_theta_gradient_against_central_differences(kind, kpls):
    """The reference has no theta gradient (algorithm.rs:880 ignores `_gradient`); the oracle's closed form is held to
    central differences of the oracle's own rlf, which the notebook fixture pins."""
    rng = np.random.default_rng(3)
    n, d = 90, 4
    x = rng.uniform(-1.0, 1.0, (n, d))
    y = np.sin(x.sum(axis=1)) + 0.1 * x[:, 0] ** 2
    xn, _, _ = O.normalize(x)
    yn, _, ys = O.normalize(y.reshape(-1, 1))
    w = rng.normal(size=(d, 2)) if kpls else np.eye(d)
    theta = rng.uniform(0.5, 1.5, w.shape[1])
    fx = O.mean_value(O.CONSTANT, xn)
    rlf, g = O.reduced_likelihood_grad(kind, xn, fx, yn[:, 0], float(ys[0]), theta, w)
    f = lambda t: O.reduced_likelihood(kind, xn, fx, yn[:, 0], float(ys[0]), t, w)[0]
    assert rlf == f(theta)
    for k in range(theta.size):
        e = np.zeros(theta.size)
        e[k] = 1e-6 * theta[k]
        fd = (f(theta + e) - f(theta - e)) / (2.0 * e[k])
        assert g[k] == pytest.approx(fd, rel=2e-6, abs=1e-6)
