"""Latency of one lock-step round of the fit driver (a batch of 11 likelihood evaluations) vs the fit itself."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as eg                                   # noqa: E402
from tools._util import make_problem, make_context         # noqa: E402

d = 10
for n in [int(a) for a in sys.argv[1:]] or [300, 500, 1000]:
    x, y = make_problem(n, d, seed=7)
    ctx = make_context(x, y, eg.MATERN52, eg.CONSTANT)
    thetas = 10.0 ** np.random.default_rng(3).uniform(-1.5, 0.5, size=(11, d))
    for _ in range(5):
        ctx.reduced_likelihood_batch(thetas)
    ts = []
    for _ in range(200):
        t0 = time.perf_counter()
        ctx.reduced_likelihood_batch(thetas)
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    t0 = time.perf_counter()
    ctx.reduced_likelihood(thetas[0])
    one = (time.perf_counter() - t0) * 1e6
    ctx.close()
    fits = []
    for _ in range(4):
        t2 = time.perf_counter()
        gp = eg.GaussianProcess.params(eg.ConstantMean, eg.Matern52Corr).fit(x, y)
        fits.append((time.perf_counter() - t2) * 1e3)
        nev = gp.n_evals()
        gp.close()
    print(json.dumps({"n": n, "graphs": os.environ.get("EGX_GRAPHS", "1"), "round11_us_median": float(np.median(ts)),
                      "round11_us_p90": float(np.percentile(ts, 90)), "single_eval_us": one,
                      "fit_ms": [round(f, 1) for f in fits], "evals": nev,
                      "rounds_est": nev / 11, "fit_us_per_round": min(fits) * 1e3 / (nev / 11)}), flush=True)
