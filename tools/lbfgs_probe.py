"""COBYLA multistart against the gradient-based one (optimizer "lbfgsb") on the C2 shape with 3 starts: fit time,
likelihood reached, evaluations spent.    python tools/lbfgs_probe.py [n] [d] [n_start]  ->  one JSON line"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as egx

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(42)
x = rng.random((n, d))
z = 4.0 * x - 2.0
y = np.sum(100.0 * (z[:, 1:] - z[:, :-1] ** 2) ** 2 + (1.0 - z[:, :-1]) ** 2, axis=1)
egx.GaussianProcess.params().corr(3).n_start(0).max_eval(25).fit(x[:512], y[:512]).close()      # warm-up
out = {"n": n, "d": d, "n_start": ns}
for name in ("cobyla", "lbfgsb"):
    t0 = time.perf_counter()
    gp = egx.GaussianProcess.params().corr(3).n_start(ns).optimizer(name).fit(x, y)
    out[name] = {"fit_s": time.perf_counter() - t0, "likelihood": gp.likelihood(), "n_evals": gp.n_evals(),
                 "theta": [float(v) for v in gp.theta()]}
    gp.close()
print(json.dumps(out), flush=True)
