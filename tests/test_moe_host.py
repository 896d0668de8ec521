"""CPU checks of the mixture's host logic: the EM clustering used when the caller brings no mixture
(control plane, linfa-clustering in the reference) and the held-out split, against the reference's stored
`gmx` block (doc/Gpx_Tutorial.ipynb:421) and scikit-learn's GaussianMixture."""
import json
import os

import numpy as np
import pytest

from oracle import moe_oracle as M


@pytest.fixture(scope="module")
def mix_json(golden_dir):
    with open(os.path.join(golden_dir, "gpx_tutorial_mixture.json")) as f:
        return json.load(f)


def _arr(o):
    return np.array(o["data"], dtype=np.float64).reshape(o["dim"])


def test_one_cluster_gmm_reproduces_notebook_gmx(mix_json):
    """The reference trains its GMM on [x | y] and keeps the x block (moe/src/algorithm.rs:118-136): for one
    cluster that is the sample mean and the MLE covariance + reg_covar 1e-6."""
    from egobox_b200.moe import fit_gmm
    x, y = _arr(mix_json["training_data"][0]), _arr(mix_json["training_data"][1])
    w, mu, cov = fit_gmm(np.concatenate([x, y[:, None]], axis=1), 1)
    g = mix_json["gmx"]
    np.testing.assert_allclose(w, _arr(g["weights"]), rtol=1e-15)
    np.testing.assert_allclose(mu[:, :1], _arr(g["means"]), rtol=1e-14)
    np.testing.assert_allclose(cov[:, :1, :1], _arr(g["covariances"]), rtol=1e-14)


def test_oracle_gmx_parameters_match_notebook(mix_json):
    g = mix_json["gmx"]
    o = M.GaussianMixture(_arr(g["weights"]), _arr(g["means"]), _arr(g["covariances"]), g["heaviside_factor"])
    np.testing.assert_allclose(o.precisions, _arr(g["precisions"]), rtol=1e-14)
    np.testing.assert_allclose(o.precisions_chol, _arr(g["precisions_chol"]), rtol=1e-14)
    np.testing.assert_allclose(o._log_det(), _arr(g["log_det"]), rtol=1e-14)


def test_gmm_separated_blobs_match_sklearn():
    sk = pytest.importorskip("sklearn.mixture")
    from egobox_b200.moe import fit_gmm
    rng = np.random.default_rng(0)
    centers = np.array([[0.0, 0.0, 0.0], [6.0, 6.0, -4.0], [-7.0, 5.0, 5.0]])
    data = np.concatenate([c + rng.standard_normal((80, 3)) * [1.0, 0.5, 1.5] for c in centers])
    w, mu, cov = fit_gmm(data, 3, seed=1)
    ref = sk.GaussianMixture(3, covariance_type="full", reg_covar=1e-6, tol=1e-6, random_state=0).fit(data)
    order = [int(np.argmin(((mu - m) ** 2).sum(axis=1))) for m in ref.means_]
    assert sorted(order) == [0, 1, 2]
    np.testing.assert_allclose(w[order], ref.weights_, atol=1e-3)
    np.testing.assert_allclose(mu[order], ref.means_, atol=1e-2)
    np.testing.assert_allclose(cov[order], ref.covariances_, atol=2e-2)


def test_extract_part_matches_reference_rule():
    from egobox_b200.moe import extract_part
    data = np.arange(46.0).reshape(23, 2)
    test, train = extract_part(data, 5)
    o_test, o_train = M.extract_part(data, 5)
    np.testing.assert_array_equal(test, o_test)
    np.testing.assert_array_equal(train, o_train)
