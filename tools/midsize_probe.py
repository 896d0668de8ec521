"""Mid-size regime (the EGO loop's usual n): likelihood batches and full default-budget fits, warm.
 usage: midsize_probe.py [n ...]      (EGX_GRAPHS=0 disables the CUDA-graph replay for an A/B)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as eg                                   # noqa: E402
from tools._util import make_problem, make_context         # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [300, 500, 1000, 2000]
    d, B = 10, 132
    for n in sizes:
        x, y = make_problem(n, d, seed=7)
        ctx = make_context(x, y, eg.MATERN52, eg.CONSTANT)
        thetas = 10.0 ** np.random.default_rng(3).uniform(-1.5, 0.5, size=(B, d))
        for _ in range(2):
            ctx.reduced_likelihood_batch(thetas[:48])
        t0 = time.perf_counter()
        st, rlf = ctx.reduced_likelihood_batch(thetas)
        t1 = time.perf_counter()
        ctx.close()
        eg.GaussianProcess.params(eg.ConstantMean, eg.Matern52Corr).fit(x, y)
        fits = []
        for _ in range(3):
            t2 = time.perf_counter()
            gp = eg.GaussianProcess.params(eg.ConstantMean, eg.Matern52Corr).fit(x, y)
            t3 = time.perf_counter()
            fits.append((t3 - t2) * 1e3)
            nev, lik = gp.n_evals(), gp.likelihood()
            gp.close()
        t2, t3 = 0.0, min(fits) * 1e-3
        print(json.dumps({"n": n, "d": d, "graphs": os.environ.get("EGX_GRAPHS", "1"),
                          "batch_us_per_eval": (t1 - t0) / B * 1e6, "ok": int(np.sum(st == 0)),
                          "fit_ms": (t3 - t2) * 1e3, "fit_ms_all": [round(f, 2) for f in fits], "fit_evals": nev,
                          "fit_us_per_eval": (t3 - t2) / nev * 1e6, "likelihood": lik}),
              flush=True)


if __name__ == "__main__":
    main()
