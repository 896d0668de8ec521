"""K1 alone: time of the correlation-matrix build (per-launch CUDA events of the stage profiler) for the four kernels at n = 8192,
at the headline dimension (d = 10) and at d = 6; algorithmic bytes = 8 n (n + 1) / 2 written.
    python tools/corr_probe.py  ->  one JSON line per (kernel, d)"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as eg                                   # noqa: E402
from tools._util import make_problem, make_context         # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for d in (10, 6):
    x, y = make_problem(n, d, seed=42)
    for name, corr in (("SquaredExponential", eg.SQUARED_EXPONENTIAL), ("AbsoluteExponential", eg.ABSOLUTE_EXPONENTIAL),
                       ("Matern32", eg.MATERN32), ("Matern52", eg.MATERN52)):
        ctx = make_context(x, y, corr, eg.CONSTANT)
        theta = np.full(d, 1.0)
        ctx.reduced_likelihood(theta)
        ctx.set_profiling(True)
        ctx.reset_profile()
        reps = 3
        for _ in range(reps):
            ctx.reduced_likelihood(theta)
        ms = ctx.profile()["corr_build"][0] / reps
        ctx.close()
        nbytes = 8.0 * n * (n + 1) / 2
        print(json.dumps({"kernel": name, "n": n, "d": d, "corr_build_ms": round(ms, 4), "algorithmic_GBs": round(nbytes / ms * 1e-6, 1),
                          "pairs_per_s": round(n * (n + 1) / 2 / ms * 1e3, 0)}), flush=True)
