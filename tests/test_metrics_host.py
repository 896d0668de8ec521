"""CPU: the cross-validation scores of egobox_b200/metrics.py (gp/src/metrics.rs, moe/src/metrics.rs) with oracle kriging
models as the refitted surrogates -- the fold arithmetic, the variance adequacy and the coverage integral are host logic; on
the device the refits are `params.fit`."""
import math

import numpy as np
import pytest
from scipy.stats import norm

from egobox_b200 import metrics
from oracle import gp_oracle as O


class _Model:
    def __init__(self, x, y, inflate=1.0):
        self.gp = O.fit(x, y, corr=O.SQEXP, mean=O.CONSTANT, theta_init=[2.0], fixed=True)
        self.inflate = inflate
        self.closed = False

    def predict(self, x):
        return self.gp.predict(x)

    def predict_valvar(self, x):
        y, v = self.gp.predict_valvar(x)
        return y, v * self.inflate

    def close(self):
        self.closed = True


@pytest.fixture(scope="module")
def data():
    rng = np.random.default_rng(3)
    x = rng.random((41, 2))
    y = np.sin(4.0 * x[:, 0]) * np.cos(2.0 * x[:, 1]) + 0.05 * rng.normal(size=41)
    return x, y


def test_folds_are_linfas():
    got = [(tr.tolist(), va.tolist()) for tr, va in metrics.folds(7, 3)]
    assert got == [([2, 3, 4, 5, 6], [0, 1]), ([0, 1, 4, 5, 6], [2, 3]), ([0, 1, 2, 3, 6], [4, 5])]   # row 6 never validates
    with pytest.raises(ValueError):
        list(metrics.folds(5, 6))


def test_pva_and_q2_against_direct_formulas(data):
    x, y = data
    made = []

    def fit(xs, ys):
        made.append(_Model(xs, ys))
        return made[-1]

    for k in (5, 41):
        press = tss = varss = 0.0
        n = 0
        for tr, va in metrics.folds(41, k):
            m = _Model(x[tr], y[tr])
            p, v = m.predict_valvar(x[va])
            press += ((y[va] - p) ** 2).sum()
            tss += ((y[va] - y.mean()) ** 2).sum()
            varss += (((y[va] - p) ** 2) / v).sum()
            n += len(va)
        assert metrics.q2_k_score((x, y), k, fit) == pytest.approx(1.0 - press / tss, rel=1e-13)
        assert metrics.pva_k_score((x, y), k, fit) == pytest.approx(abs(math.log(varss / n)), rel=1e-13)
    assert all(m.closed for m in made) and len(made) == 2 * (5 + 41)        # every refit is released


def test_iae_alpha_coverage(data):
    x, y = data
    score, alphas, deltas = metrics.iae_alpha_k_score((x, y), 5, lambda xs, ys: _Model(xs, ys))
    assert alphas[0] == 0.02 and alphas[-1] == pytest.approx(0.98) and alphas.size == 20
    # direct: coverage of the central (1 - alpha) intervals, averaged over folds, against 1 - alpha
    cov = np.zeros(20)
    sc = []
    for tr, va in metrics.folds(41, 5):
        p, v = _Model(x[tr], y[tr]).predict_valvar(x[va])
        z = np.abs(y[va] - p) / np.sqrt(v)
        d = np.array([(z <= norm.ppf(1.0 - a / 2.0)).mean() for a in alphas])
        cov += d
        sc.append(np.abs(d - (1.0 - alphas)).mean())
    np.testing.assert_allclose(deltas, cov / 5, rtol=1e-12)
    assert score == pytest.approx(np.mean(sc), rel=1e-12)
    # absurdly wide intervals cover everything: coverage 1 for every alpha, IAE = mean(alpha) = 0.5
    wide, _, d_wide = metrics.iae_alpha_k_score((x, y), 5, lambda xs, ys: _Model(xs, ys, inflate=1e12))
    assert np.all(d_wide == 1.0) and wide == pytest.approx(0.5, rel=1e-12)
