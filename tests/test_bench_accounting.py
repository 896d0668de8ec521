"""bench.py's roofline bookkeeping: the useful flops attributed to the tcgen05 launches and to the remaining DMMA
launches must add up to the flops of the K = 128 / K = 256 updates of the factorisations and of the predict_var solve."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_tcgen05_and_dmma_flops_partition_the_total():
    b = _bench()
    n, evals, pts = 8192, 3, 8192
    total = b.gemm_algorithmic_flops(n, evals, pts)
    fact, l_fact, t_fact = b.ozaki_algorithmic_flops(n, evals + 1)
    pred, l_pred, t_pred = b.ozaki_predict_flops(n, pts)
    assert 0 < fact < total and 0 < pred < total and fact + pred < total
    assert l_fact == 28 * (evals + 1) and l_pred == 28          # pair steps with >= 8 trailing block columns at T = 64
    assert (total - fact - pred) / total < 0.05                 # partner columns + small blocks: a few per cent
    # the factorisation share of the tcgen05 kernel approaches n^3/3 per factorisation
    assert 0.8 < fact / ((evals + 1) * n ** 3 / 3.0) < 1.0
    # below 8 row tiles a predict chunk stays on DMMA
    assert b.ozaki_predict_flops(n, 512)[0] == 0.0
    # tiles: lower triangle + one appended tile row per launch
    assert t_fact == (evals + 1) * sum(t * (t + 1) // 2 + t for t in range(8, 63, 2))
