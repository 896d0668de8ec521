"""-m gpu: mixture of experts on the device (csrc/moe.cu behind egx_moe_*) against
oracle/moe_oracle.py, and the reference's own mixture tests (moe/src/algorithm.rs:1207-1290, 1679-1731)."""
import json
import os

import numpy as np
import pytest

from oracle import gp_oracle as O
from oracle import moe_oracle as M

pytestmark = pytest.mark.gpu


def _random_mixture(rng, k, nx):
    a = rng.random((k, nx, nx))
    covs = np.stack([m @ m.T + 0.3 * np.eye(nx) for m in a])
    w = rng.random(k) + 0.2
    return w / w.sum(), rng.random((k, nx)) * 2.0, covs


@pytest.mark.parametrize("k,nx,factor", [(3, 4, 0.7), (2, 1, 1.0), (5, 7, 1.9), (1, 3, 1.0)])
def test_gaussian_mixture_device_vs_oracle(k, nx, factor):
    import egobox_b200 as egx
    rng = np.random.default_rng(10 * k + nx)
    w, mu, cov = _random_mixture(rng, k, nx)
    g = egx.GaussianMixture(w, mu, cov, factor)
    o = M.GaussianMixture(w, mu, cov, factor)
    x = rng.random((257, nx)) * 2.5 - 0.25
    np.testing.assert_allclose(g.predict_probas(x), o.predict_probas(x), rtol=1e-11, atol=1e-300)
    np.testing.assert_array_equal(g.predict(x), o.predict(x))
    if k > 1:
        np.testing.assert_allclose(g.predict_probas_derivatives(x), o.predict_probas_derivatives(x),
                                   rtol=1e-9, atol=1e-13)
    prec, pc, ld = g.parameters()
    np.testing.assert_allclose(prec, o.precisions, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(pc, o.precisions_chol, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(ld, o._log_det(), rtol=1e-12, atol=1e-14)
    g.close()


def test_gmx_block_of_the_notebook_model(golden_dir):
    """doc/Gpx_Tutorial.ipynb:421: precisions / precisions_chol / log_det of the stored one-cluster mixture."""
    import egobox_b200 as egx
    with open(os.path.join(golden_dir, "gpx_tutorial_mixture.json")) as f:
        ref = json.load(f)["gmx"]
    arr = lambda o_: np.array(o_["data"], dtype=np.float64).reshape(o_["dim"])
    g = egx.GaussianMixture.from_dict(ref)
    d = g.to_dict()
    for key in ("weights", "means", "covariances", "precisions", "precisions_chol", "log_det"):
        assert d[key]["dim"] == ref[key]["dim"]
        np.testing.assert_allclose(arr(d[key]), arr(ref[key]), rtol=1e-14)
    assert d["heaviside_factor"] == ref["heaviside_factor"]


def _two_cluster_problem(seed=0, nx=2, n=90):
    rng = np.random.default_rng(seed)
    x = rng.random((n, nx))
    y = np.where(x[:, 0] < 0.5, np.sin(4 * x).sum(axis=1), 3.0 + (x ** 2).sum(axis=1))
    w = np.array([0.5, 0.5])
    mu = np.full((2, nx), 0.5)
    mu[0, 0], mu[1, 0] = 0.25, 0.75
    cov = np.stack([np.diag([0.02] + [0.08] * (nx - 1))] * 2)
    return x, y, w, mu, cov


@pytest.mark.parametrize("recomb,factor", [("hard", 1.0), ("smooth", 1.0), ("smooth", 0.4)])
def test_mixture_predictions_vs_oracle(recomb, factor):
    """train_on_clusters with a preset mixture and fixed theta: experts, responsibilities and the recombined
    values / variances / gradients must equal the oracle's composition of the same formulas."""
    import egobox_b200 as egx
    x, y, w, mu, cov = _two_cluster_problem()
    theta = [0.8, 1.3]
    code = egx.Recombination.HARD if recomb == "hard" else egx.Recombination.SMOOTH
    gmx = egx.GaussianMixture(w, mu, cov, factor)
    prm = egx.GpMixtureParams().set(n_clusters=2, recombination=code, heaviside=factor, correlation_spec=8,
                                    theta_tunings=[egx.ThetaTuning.Fixed(theta)] * 2)
    mix = prm.train_on_clusters(x, y, gmx)
    ogmx = M.GaussianMixture(w, mu, cov, factor)
    labels = ogmx.predict(x)
    oexp = [O.fit(x[labels == c], y[labels == c], corr=O.MATERN52, mean=O.CONSTANT, theta_init=theta, fixed=True)
            for c in range(2)]
    for c in range(2):
        assert mix.experts[c].likelihood() == pytest.approx(oexp[c].likelihood, rel=1e-9)
    rng = np.random.default_rng(5)
    xq = rng.random((300, 2))
    if recomb == "hard":
        yo = M.predict_hard(oexp, ogmx, xq, "predict")
        vo = M.predict_hard(oexp, ogmx, xq, "predict_var")
        gyo = M.predict_hard(oexp, ogmx, xq, "predict_gradients")
        gvo = M.predict_hard(oexp, ogmx, xq, "predict_var_gradients")
    else:
        yo = M.predict_smooth(oexp, ogmx, xq)
        vo = M.predict_var_smooth(oexp, ogmx, xq)
        gyo = M.predict_gradients_smooth(oexp, ogmx, xq)
        gvo = M.predict_var_gradients_smooth(oexp, ogmx, xq)
    s2 = max(e.inner.sigma2 for e in oexp)
    np.testing.assert_allclose(mix.predict(xq), yo, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(mix.predict_var(xq), vo, rtol=1e-6, atol=1e-9 * s2)
    yv = mix.predict_valvar(xq)
    np.testing.assert_allclose(yv[0], yo, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(yv[1], vo, rtol=1e-6, atol=1e-9 * s2)
    np.testing.assert_allclose(mix.predict_gradients(xq), gyo, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(mix.predict_var_gradients(xq), gvo, rtol=1e-5, atol=1e-8 * s2)
    g2 = mix.predict_valvar_gradients(xq)
    np.testing.assert_allclose(g2[0], gyo, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(g2[1], gvo, rtol=1e-5, atol=1e-8 * s2)
    mix.close()


def f_test_1d(x):
    """moe/src/algorithm.rs:1175-1188."""
    x = np.asarray(x).reshape(-1)
    return np.where(x < 0.4, x * x, np.where(x < 0.8, 3.0 * x + 1.0, np.sin(10.0 * x)))


def test_moe_hard_three_clusters():
    """moe/src/algorithm.rs:1207-1239 (test_moe_hard): 50 points, 3 clusters, hard recombination."""
    import egobox_b200 as egx
    rng = np.random.default_rng(0)
    xt = rng.random((50, 1))
    gpx = egx.Gpx.builder(n_clusters=3, recombination=egx.Recombination.HARD, seed=0).fit(xt, f_test_1d(xt))
    assert gpx.thetas().shape == (3, 1) and gpx.likelihoods().shape == (3,) and gpx.variances().shape == (3,)
    assert gpx.predict(np.array([[0.39]])).item() == pytest.approx(0.39 * 0.39, abs=1e-3)
    assert gpx.predict(np.array([[0.82]])).item() == pytest.approx(np.sin(8.2), abs=1e-3)
    assert gpx.predict_gradients(np.array([[0.2], [0.6]])).shape == (2, 1)
    assert str(gpx).startswith("Mixture[Hard](Constant_SquaredExponentialGP(")
    with pytest.raises(egx.GpError):
        gpx.sample(np.array([[0.5]]), 2)             # "Can not sample when several clusters"


def test_moe_smooth_three_clusters_and_heaviside_search(tmp_path):
    """moe/src/algorithm.rs:1242-1288 (test_moe_smooth, Smooth(None) -> optimised heaviside factor) and
    :1361-1380 (save / load round trip of a 3-cluster mixture)."""
    import egobox_b200 as egx
    rng = np.random.default_rng(42)
    xt = rng.random((60, 1))
    gpx = egx.Gpx.builder(n_clusters=3, recombination=egx.Recombination.SMOOTH, seed=42).fit(xt, f_test_1d(xt))
    assert gpx.predict(np.array([[0.37]])).item() == pytest.approx(0.37 * 0.37, abs=5e-3)
    assert 0.1 <= gpx.mixture().gmx.heaviside_factor() <= 2.1
    fn = str(tmp_path / "saved_moe.json")
    assert gpx.save(fn)
    obj = json.load(open(fn))
    assert set(obj) == {"recombination", "experts", "gmx", "gp_type", "training_data", "params"}
    assert len(obj["experts"]) == 3 and obj["gmx"]["weights"]["dim"] == [3]
    again = egx.Gpx.load(fn)
    xq = np.array([[0.6], [0.1], [0.95]])
    np.testing.assert_allclose(again.predict(xq), gpx.predict(xq), atol=1e-6)
    np.testing.assert_allclose(again.predict_var(xq), gpx.predict_var(xq), atol=1e-6)
    # the binary format of the same three-expert mixture (moe/src/algorithm.rs:1361-1380 saves with GpFileFormat::Binary too)
    fb = str(tmp_path / "saved_moe.bin")
    assert gpx.save(fb)
    again_b = egx.Gpx.load(fb)
    np.testing.assert_allclose(again_b.predict(xq), gpx.predict(xq), atol=1e-6)
    np.testing.assert_allclose(again_b.predict_var(xq), gpx.predict_var(xq), atol=1e-6)


def test_smooth_equals_hard_for_one_cluster():
    """moe/src/algorithm.rs:1679-1731."""
    import egobox_b200 as egx
    rng = np.random.default_rng(42)
    xt = rng.random((50, 2))
    yt = (xt * xt).sum(axis=1)
    hard = egx.Gpx.builder(n_clusters=1, recombination=egx.Recombination.HARD).fit(xt, yt)
    smooth = egx.Gpx.builder(n_clusters=1, recombination=egx.Recombination.SMOOTH).fit(xt, yt)
    x = np.random.default_rng(43).random((1, 2))
    for f in ("predict", "predict_var", "predict_gradients", "predict_var_gradients"):
        np.testing.assert_allclose(getattr(hard, f)(x), getattr(smooth, f)(x), atol=1e-5)


def test_cross_validated_expert_selection_reproduces_notebook_choice(golden_dir):
    """doc/Gpx_Tutorial.ipynb:420-421: regression_spec CONSTANT|LINEAR|QUADRATIC and correlation_spec
    SQUAREDEXPONENTIAL|MATERN52 -- the stored model is the expert the reference's 5-fold cross-validation SELECTED
    (Linear + Matern52, theta 5.03288..., likelihood 3.0156045880805125)."""
    import egobox_b200 as egx
    with open(os.path.join(golden_dir, "gpx_tutorial_mixture.json")) as f:
        ref = json.load(f)
    arr = lambda o_: np.array(o_["data"], dtype=np.float64).reshape(o_["dim"])
    xt, yt = arr(ref["training_data"][0]), arr(ref["training_data"][1])
    gpx = egx.Gpx.builder(regr_spec=egx.RegressionSpec.ALL,
                          corr_spec=egx.CorrelationSpec.SQUARED_EXPONENTIAL | egx.CorrelationSpec.MATERN52).fit(xt, yt)
    assert gpx.to_dict()["experts"][0]["type_fullgp"] == ref["selected_expert"]
    assert gpx.gp().cv_errors_ is not None and len(gpx.gp().cv_errors_) == 6
    assert gpx.thetas().item() == pytest.approx(5.0328871070499, rel=2e-2)
    assert gpx.likelihoods().item() == pytest.approx(3.0156045880805125, rel=1e-4)
    assert str(gpx).startswith("Mixture[Hard](Linear_Matern52GP(mean=LinearMean, corr=Matern52, theta=[5.0")
    d = gpx.to_dict()
    assert d["params"]["regression_spec"] == ref["params"]["regression_spec"]
    assert d["params"]["correlation_spec"] == ref["params"]["correlation_spec"]
    assert d["params"]["n_clusters"] == ref["params"]["n_clusters"]
    for key in ("weights", "means", "covariances", "precisions", "precisions_chol", "log_det"):
        np.testing.assert_allclose(arr(d["gmx"][key]), arr(ref["gmx"][key]), rtol=1e-13)
