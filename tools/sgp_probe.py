"""BASELINE config 3 probe: FITC N=100000 d=6 M=1024 -- likelihood evaluation and prediction timings."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as eg   # noqa: E402

N, d, M = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (100000, 6, 1024)))
rng = np.random.default_rng(42)
x = rng.random((N, d))
y = np.sum(np.sin(3 * np.pi * x), axis=1) + rng.normal(0, 0.1, N)
z = x[rng.permutation(N)[:M]].copy()
ctx = eg.SgpContext(x, y, z, corr=eg.MATERN52, method=eg.SparseMethod.FITC)
theta = np.full(d, 1.0)
for _ in range(2):
    st, lik = ctx.reduced_likelihood(theta, 1.0, 0.01)
if not os.environ.get("PROBE_NOPROF"):      # per-launch events serialise the launches: unset for stage times only
    ctx.set_profiling(True)
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    st, lik = ctx.reduced_likelihood(theta, 1.0, 0.01)
t1 = time.perf_counter()
prof = ctx.profile()
print(json.dumps({"N": N, "d": d, "M": M, "status": st, "lik": lik, "ms_per_eval": (t1 - t0) / reps * 1e3,
                  "stage_ms_per_eval": {k: round(v[0] / reps, 3) for k, v in prof.items() if v[1]},
                  "flops_2M2N": 2.0 * M * M * N, "tflops": 2.0 * M * M * N / ((t1 - t0) / reps) / 1e12}))
ctx.finalize(theta, 1.0, 0.01)
xs = rng.random((100000, d))
t0 = time.perf_counter()
v = ctx.predict_var(xs)
t1 = time.perf_counter()
print(json.dumps({"predict_var_100k_ms": (t1 - t0) * 1e3, "var_mean": float(v.mean())}))
ctx.close()
