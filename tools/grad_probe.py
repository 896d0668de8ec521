"""Closed-form theta gradient against the batched central-difference one on the C2 shape (Matern-5/2, n = 8192, d = 10):
wall time of each call after a warm-up, per-stage device time of the closed form, and their agreement.
    python tools/grad_probe.py [n] [d]  ->  one JSON line per run on stdout"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as eg
from egobox_b200._lib import load

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rng = np.random.default_rng(42)
x = rng.random((n, d))
z = 4.0 * x - 2.0
y = np.sum(100.0 * (z[:, 1:] - z[:, :-1] ** 2) ** 2 + (1.0 - z[:, :-1]) ** 2, axis=1)
xm, xs = x.mean(0), x.std(0, ddof=1)
ym, ys = y.mean(), y.std(ddof=1)
ctx = eg.GpContext((x - xm) / xs, (y - ym) / ys, xm, xs, float(ym), float(ys), 3, 0)
theta = np.full(d, 0.7)
out = {"n": n, "d": d}
for name, fn in (("closed_form", lambda: ctx.reduced_likelihood_grad_analytic(theta)),
                 ("central_differences", lambda: ctx.reduced_likelihood_grad(theta, rel_step=1e-5))):
    fn()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        st, rlf, g = fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    out[name + "_ms"] = min(ts)
    out[name + "_grad"] = [float(v) for v in g]
    out[name + "_status"] = int(st)
ga, gf = np.array(out["closed_form_grad"]), np.array(out["central_differences_grad"])
out["max_rel_diff"] = float(np.max(np.abs(ga - gf)) / np.max(np.abs(ga)))
ctx.set_profiling(True)
ctx.reset_profile()
ctx.reduced_likelihood_grad_analytic(theta)
out["closed_form_stage_ms"] = {k: round(v[0], 3) for k, v in ctx.profile().items() if v[1] > 0}
print(json.dumps(out), flush=True)
ctx.close()
