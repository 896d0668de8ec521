"""Secondary BASELINE.json configs on one GPU (not the bench headline): prints one JSON line each.
 C1 Kriging SquaredExponential n=200 d=1 : full fit (default budget) + predict_var
 C4 one MoE expert n=4096 d=20 Matern52 : fit with the default budget (1 of the 8 experts / GPUs)
 C5 theta sweep of 512 candidates, d=20, n in {21, 100, 2048}"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as eg                                   # noqa: E402
from tools._util import make_problem, make_context         # noqa: E402


def c1():
    x = np.sort(np.random.default_rng(42).random((200, 1)) * 25.0, axis=0)
    y = (x[:, 0] - 3.5) * np.sin((x[:, 0] - 3.5) / np.pi)
    eg.Kriging.params().fit(x[:50], y[:50]).predict_var(x[:10])     # CUDA context / module load outside the timing
    t0 = time.perf_counter()
    gp = eg.Kriging.params().fit(x, y)
    t1 = time.perf_counter()
    xs = np.linspace(0, 25, 100000)[:, None]
    v = gp.predict_var(xs)
    t2 = time.perf_counter()
    print(json.dumps({"config": "C1 Kriging SqExp n=200 d=1", "fit_ms": (t1 - t0) * 1e3, "evals": gp.n_evals(),
                      "theta": gp.theta().tolist(), "likelihood": gp.likelihood(),
                      "predict_var_100k_ms": (t2 - t1) * 1e3}), flush=True)


def c4():
    n, d = 4096, 20
    x, y = make_problem(n, d, seed=42)
    t0 = time.perf_counter()
    gp = eg.GaussianProcess.params(eg.ConstantMean, eg.Matern52Corr).fit(x, y)
    t1 = time.perf_counter()
    print(json.dumps({"config": "C4 one expert n=4096 d=20 Matern52 (default budget 11 x 200 + 1)",
                      "fit_s": t1 - t0, "evals": gp.n_evals(), "evals_per_s": gp.n_evals() / (t1 - t0),
                      "likelihood": gp.likelihood()}), flush=True)


def c5():
    d, B = 20, 512
    for n in (21, 100, 2048):
        x, y = make_problem(n, d, seed=5)
        ctx = make_context(x, y, eg.MATERN52, eg.CONSTANT)
        thetas = 10.0 ** np.random.default_rng(42).uniform(-2.0, 1.0, size=(B, d))
        ctx.reduced_likelihood_batch(thetas[:32])
        t0 = time.perf_counter()
        st, rlf = ctx.reduced_likelihood_batch(thetas)
        t1 = time.perf_counter()
        print(json.dumps({"config": "C5 theta sweep 512 candidates d=20 n=%d" % n, "ms": (t1 - t0) * 1e3,
                          "evals_per_s": B / (t1 - t0), "ok": int((st == 0).sum())}), flush=True)
        ctx.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c5", "c4"]
    for w in which:
        {"c1": c1, "c4": c4, "c5": c5}[w]()
