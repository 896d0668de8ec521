"""Fit through the public API (multistart driver): wall time, evaluations, optimum -- to compare the independent-chain
driver (default) with the lock-step one (EGX_FIT_LOCKSTEP=1); both must find the same optimum."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as eg                                   # noqa: E402
from tools._util import make_problem                       # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [2000, 8192]:
    d = 10
    x, y = make_problem(n, d, seed=42)
    prm = eg.GaussianProcess.params(eg.ConstantMean, eg.Matern52Corr).cobyla_ftol_rel(0.0)
    if os.environ.get("PROBE_NSTART"):                     # n_start + 1 chains (what one rank of a sharded fit carries)
        prm = prm.n_start(int(os.environ["PROBE_NSTART"]))
    if n <= 2500:
        prm.fit(x, y).close()                              # warm-up (module load, graph captures)
    for rep in range(2 if n > 2500 else 1):                # the second fit of a large problem is the warm one (block cache, modules)
        t0 = time.perf_counter()
        gp = prm.fit(x, y)
        t1 = time.perf_counter()
        xs = np.random.default_rng(1).random((100000, d))
        t2 = time.perf_counter()
        var = gp.predict_var(xs)
        t3 = time.perf_counter()
        print(json.dumps({"n": n, "rep": rep, "lockstep": os.environ.get("EGX_FIT_LOCKSTEP", "0"), "fit_s": t1 - t0, "evals": gp.n_evals(),
                          "ms_per_eval": (t1 - t0) / gp.n_evals() * 1e3, "predict_var_100k_s": t3 - t2, "likelihood": gp.likelihood(),
                          "theta0": float(gp.theta()[0])}), flush=True)
        gp.close()
