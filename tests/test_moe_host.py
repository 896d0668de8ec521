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


# --------------------------------------------------------------------------- automatic number of clusters
def _f_test_1d(x):
    """moe/src/clustering.rs:398-410 (`function_test_1d`): three regimes on [0, 1]."""
    x = x[:, 0]
    return np.where(x < 0.4, x * x, np.where(x < 0.8, 3.0 * x + 1.0, np.sin(10.0 * x)))


class _OracleMixture:
    """Stand-in for the device mixture: EM clustering + one fixed-theta oracle kriging per cluster."""

    def __init__(self, k, xtr, ytr):
        from egobox_b200.moe import fit_gmm
        from oracle import gp_oracle as O
        nx = xtr.shape[1]
        w, mu, cov = fit_gmm(np.concatenate([xtr, ytr[:, None]], axis=1), k, seed=7)
        self.gmx = M.GaussianMixture(w, mu[:, :nx], cov[:, :nx, :nx], 1.0)
        labels = self.gmx.predict(xtr)
        self.experts = []
        for c in range(k):
            rows = np.nonzero(labels == c)[0]
            if rows.size < 3:
                from egobox_b200.gp import GpError
                raise GpError("Not enough points in cluster")
            self.experts.append(O.fit(xtr[rows], ytr[rows], corr=O.SQEXP, mean=O.CONSTANT, theta_init=[3.0], fixed=True))

    def predict_hard(self, x):
        return M.predict_hard(self.experts, self.gmx, x)

    def predict_smooth1(self, x):
        return M.predict_smooth(self.experts, self.gmx, x)

    def close(self):
        pass


def test_find_best_number_of_clusters_on_the_reference_test_function():
    """moe/src/clustering.rs:421-443 (`test_find_best_cluster_nb_1d`): 50 LHS points of the three-regime function, up to 3
    clusters -> 3.  The search itself is host logic; the mixtures it cross-validates are oracle stand-ins here."""
    from egobox_b200.moe import find_best_number_of_clusters, HARD, SMOOTH
    from oracle import gp_oracle as O
    for seed in (42, 1):
        x = O.lhs_classic(np.array([[0.0, 1.0]]), 50, np.random.default_rng(seed))     # Lhs::sample(50), unordered rows
        y = _f_test_1d(x)
        k, recomb, heaviside = find_best_number_of_clusters(x, y, 3, _OracleMixture, seed=42)
        assert k == 3 and recomb in (HARD, SMOOTH) and heaviside is None


def test_cluster_search_rules_with_scripted_errors():
    """Selection and stopping rules of clustering.rs:270-362 on scripted fold errors."""
    from egobox_b200 import moe as E

    def run(table, max_nb, nan_for=()):
        calls = []

        class Scripted:
            def __init__(self, k, xtr, ytr):
                self.k = k
                calls.append(k)

            def predict_hard(self, xv):
                if self.k in nan_for:
                    return np.full(xv.shape[0], np.nan)
                return np.full(xv.shape[0], 1.0 + table[self.k][0])          # actual = 1 -> hard error = table value

            def predict_smooth1(self, xv):
                return np.full(xv.shape[0], 1.0 + table[self.k][1] / xv.shape[0])   # smooth error is a SUM

            def close(self):
                pass

        x = np.linspace(0.0, 1.0, 100)[:, None]
        y = np.ones(100)
        # k equal clusters along x (y is constant): every fold keeps > 3 values per cluster
        gmm = lambda data, k, seed=None: (np.full(k, 1.0 / k),
                                          np.stack([(np.arange(k) + 0.5) / k, np.ones(k)], axis=1),
                                          np.tile(np.eye(2) * 0.01, (k, 1, 1)))
        return E.find_best_number_of_clusters(x, y, max_nb, Scripted, gmm_fit=gmm), calls

    # hard error smallest at 2 clusters and below every smooth error -> (2, HARD)
    (k, recomb, _), calls = run({1: (0.5, 0.9), 2: (0.1, 0.8), 3: (0.3, 0.7)}, 3)
    assert (k, recomb) == (2, E.HARD) and calls == [1] * 5 + [2] * 5 + [3] * 5
    # smooth error smallest overall -> its count, SMOOTH with the factor left to the heaviside search
    (k, recomb, hv), _ = run({1: (0.5, 0.9), 2: (0.4, 0.8), 3: (0.3, 0.05)}, 3)
    assert (k, recomb, hv) == (3, E.SMOOTH, None)
    # both medians rise twice in a row after 3 clusters: the search stops at i = 4 (5 clusters), never tries 6
    t = {1: (0.5, 0.5), 2: (0.4, 0.4), 3: (0.1, 0.3), 4: (0.2, 0.35), 5: (0.3, 0.4), 6: (0.0, 0.0), 7: (0.0, 0.0)}
    (k, recomb, _), calls = run(t, 7)
    assert (k, recomb) == (3, E.HARD) and max(calls) == 5
    # NaN predictions disqualify a count even when its error would be the smallest
    (k, recomb, _), _ = run({1: (0.5, 0.9), 2: (0.0, 0.0), 3: (0.3, 0.7)}, 3, nan_for=(2,))
    assert k in (1, 3) and k != 2
    # nothing usable: 1 cluster, smooth
    (k, recomb, _), _ = run({1: (0.5, 0.9)}, 1, nan_for=(1,))
    assert (k, recomb) == (1, E.SMOOTH)
    # max_nb_clusters = 0 -> n / 10 + 1 = 11 candidates at most (the rise rule may stop earlier)
    (_, _, _), calls = run({k: (1.0 / k, 1.0 / k) for k in range(1, 12)}, 0)
    assert max(calls) == 11


def test_auto_cluster_glue_in_fit_without_a_device(monkeypatch):
    """GpMixtureParams.fit with n_clusters <= 0: the search result is turned into a fixed-count fit with the chosen
    recombination.  The device classes are replaced by recorders, only the host glue runs here."""
    from egobox_b200 import moe as E
    fits = []

    class FakeGmx:
        def __init__(self, w, mu, cov, factor, device):
            self.k = len(w)

        def n_clusters(self):
            return self.k

    class FakeMix:
        def __init__(self, k):
            self.k = k

    def fake_train(self, xt, yt, gmx):
        fits.append((self.n_clusters, self.recombination, self.heaviside, xt.shape[0], len(self.theta_tunings)))
        return FakeMix(gmx.n_clusters())

    monkeypatch.setattr(E, "GaussianMixture", FakeGmx)
    monkeypatch.setattr(E.GpMixtureParams, "train_on_clusters", fake_train)
    monkeypatch.setattr(E, "find_best_number_of_clusters",
                        lambda x, y, max_nb, fit_mixture, seed=None: (fits.append(("search", max_nb)), (3, E.SMOOTH, None))[1])
    x = np.linspace(0.0, 1.0, 60)[:, None]
    y = _f_test_1d(x)
    mix = E.GpMixtureParams().set(n_clusters=0, seed=1).fit(x, y)
    assert fits[0] == ("search", 7)                                   # n / 10 + 1
    # 3 clusters, Smooth(None) (the heaviside search follows inside train_on_clusters), one tuning for all experts; the
    # experts see every row -- only the clustering is trained on the 95 % split (algorithm.rs:108-116, 143)
    assert fits[1] == (3, E.SMOOTH, None, 60, 1) and mix.k == 3
    fits.clear()
    E.GpMixtureParams().set(n_clusters=-4, seed=1).fit(x, y)
    assert fits[0] == ("search", 4)
