"""Steady-state batched likelihood throughput vs the number of workspaces kept in flight (EGX_BATCH_STREAMS)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as eg                                   # noqa: E402
from tools._util import make_problem, make_context         # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
B = int(sys.argv[2]) if len(sys.argv) > 2 else 48
d = 10
x, y = make_problem(n, d, seed=42)
ctx = make_context(x, y, eg.MATERN52, eg.CONSTANT)
thetas = np.tile(np.full(d, 1.0), (B, 1)) * np.linspace(0.8, 1.2, B)[:, None]
ctx.reduced_likelihood_batch(thetas[:8])
ctx.reduced_likelihood_batch(thetas)          # includes the graph captures of the replicas (n <= 4096)
t0 = time.perf_counter()
st, rl = ctx.reduced_likelihood_batch(thetas)
t1 = time.perf_counter()
print(json.dumps({"n": n, "B": B, "W": os.environ.get("EGX_BATCH_STREAMS", "default"),
                  "ms_per_eval": (t1 - t0) / B * 1e3, "ok": int((st == 0).sum())}), flush=True)
ctx.close()
