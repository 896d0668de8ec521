"""Pins oracle/moe_oracle.py (Gaussian-mixture responsibilities + recombination) on the known answers the
reference holds: gaussian_mixture.rs:371-397 (test_pdfs), :344-358 (test_gmx_one_cluster)."""
import numpy as np
import pytest

from oracle import moe_oracle as M


@pytest.mark.parametrize("means,covs,expected,x", [
    ([[0.0, 0.0]], [[[1.0, 0.0], [0.0, 1.0]]], 0.05854983152431917, [1.0, 1.0]),
    ([[0.0, 0.0]], [[[1.0, 0.0], [0.0, 1.0]]], 0.013064233284684921, [1.0, 2.0]),
    ([[0.5, -0.2]], [[[2.0, 0.3], [0.3, 0.5]]], 0.00014842259203296995, [-1.0, 2.0]),
])
def test_pdfs(means, covs, expected, x):
    """gaussian_mixture.rs:360-397."""
    g = M.GaussianMixture([1.0], means, covs)
    assert g.pdfs(x)[0] == pytest.approx(expected, abs=1e-15, rel=1e-12)


def test_one_cluster():
    """gaussian_mixture.rs:344-358."""
    g = M.GaussianMixture([1.0], [[4.0, 4.0]], [[[3.0, 0.0], [0.0, 3.0]]], 1.0)
    obs = np.repeat(np.linspace(0.0, 4.0, 11)[:, None], 2, axis=1)
    assert np.all(g.predict(obs) == 0)
    assert np.all(g.predict_probas(obs) == 1.0)


def _two():
    return M.GaussianMixture([0.5, 0.5], [[0.0, 0.0], [4.0, 4.0]],
                             [[[3.0, 0.0], [0.0, 3.0]], [[3.0, 0.0], [0.0, 3.0]]], 0.99)


def test_two_clusters_symmetry_and_normalisation():
    """gaussian_mixture.rs:325-341 (test_gmx, which only prints): symmetric mixture -> the midpoint is 50/50,
    responsibilities sum to one and the hard label switches at the midpoint."""
    g = _two()
    obs = np.repeat(np.linspace(0.0, 4.0, 11)[:, None], 2, axis=1)
    p = g.predict_probas(obs)
    np.testing.assert_allclose(p.sum(axis=1), 1.0, rtol=1e-14)
    assert p[5, 0] == pytest.approx(0.5, rel=1e-12)
    np.testing.assert_allclose(p[:, 0], p[::-1, 1], rtol=1e-12)
    assert list(g.predict(obs)) == [0] * 6 + [1] * 5          # ties go to the first cluster


def test_heaviside_factor_sharpens():
    obs = np.array([[1.5, 1.5]])
    p_sharp = _two().with_heaviside_factor(0.2).predict_probas(obs)[0, 0]
    p_soft = _two().with_heaviside_factor(2.0).predict_probas(obs)[0, 0]
    assert p_sharp > _two().predict_probas(obs)[0, 0] > p_soft > 0.5


def test_probas_derivatives_match_finite_differences():
    rng = np.random.default_rng(0)
    a = rng.random((3, 3, 3))
    covs = np.stack([m @ m.T + 0.5 * np.eye(3) for m in a])
    g = M.GaussianMixture([0.2, 0.5, 0.3], rng.random((3, 3)) * 2, covs, 0.7)
    x = rng.random((4, 3)) * 2
    d = g.predict_probas_derivatives(x)
    h = 1e-6
    for j in range(3):
        e = np.zeros(3)
        e[j] = h
        fd = (g.predict_probas(x + e) - g.predict_probas(x - e)) / (2 * h)
        np.testing.assert_allclose(d[:, :, j], fd, rtol=1e-6, atol=1e-9)


def test_extract_part_and_folds():
    data = np.arange(23)[:, None].astype(float)
    test, train = M.extract_part(data, 5)
    assert list(test[:, 0]) == [0, 5, 10, 15, 20] and train.shape[0] == 18
    folds = list(M.cv_folds(23, 5))
    assert [len(v) for _, v in folds] == [4] * 5 and all(len(t) == 19 for t, _ in folds)
