"""-m gpu: automatic number of clusters (`Gpx.builder(n_clusters=0)`, NbClusters::Auto) end to end on the device.
Mirrors `test_moe_auto`, moe/src/algorithm.rs:1291-1311.  The search itself (moe.find_best_number_of_clusters) is covered on
the CPU in tests/test_moe_host.py with oracle stand-ins; here every cross-validated mixture is a device mixture.
(Sorted last on purpose: ~140 small GPU fits.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _f_test_1d(x):
    x = x[:, 0]
    return np.where(x < 0.4, x * x, np.where(x < 0.8, 3.0 * x + 1.0, np.sin(10.0 * x)))


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
