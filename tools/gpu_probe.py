"""First-light probe: per-stage device timings of one likelihood evaluation and a
predict_valvar chunk at a few sizes (not a bench; prints JSON lines to stdout)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as eg                                  # noqa: E402
from tools._util import make_problem, make_context        # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 4096, 8192]
    for n in sizes:
        d = 10
        x, y = make_problem(n, d, seed=42)
        ctx = make_context(x, y, eg.MATERN52, eg.CONSTANT)
        theta = np.full(d, 1.0)
        for _ in range(2):
            ctx.reduced_likelihood(theta)
        ctx.set_profiling(True)
        ctx.reset_profile()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            st, rlf = ctx.reduced_likelihood(theta)
        t1 = time.perf_counter()
        prof = ctx.profile()
        print(json.dumps({"n": n, "status": st, "rlf": rlf, "wall_ms_per_eval": (t1 - t0) / reps * 1e3,
                          "stage_ms_per_eval": {k: round(v[0] / reps, 4) for k, v in prof.items()},
                          "launches_per_eval": {k: v[1] // reps for k, v in prof.items()}}), flush=True)
        ctx.set_profiling(False)
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.reduced_likelihood(theta)
        t1 = time.perf_counter()
        print(json.dumps({"n": n, "wall_ms_per_eval_noprof": (t1 - t0) / reps * 1e3}), flush=True)
        thetas = np.tile(theta, (12, 1)) * np.linspace(0.8, 1.2, 12)[:, None]
        ctx.reduced_likelihood_batch(thetas[:4])
        t0 = time.perf_counter()
        stb, rlfb = ctx.reduced_likelihood_batch(thetas)
        t1 = time.perf_counter()
        print(json.dumps({"n": n, "batch12_ms_per_eval": (t1 - t0) / 12 * 1e3, "batch_status_ok": int((stb == 0).sum()),
                          "batch_streams": __import__("os").environ.get("EGX_BATCH_STREAMS", "3")}), flush=True)
        st, res = ctx.finalize(theta)
        m = 8192
        xs = np.random.default_rng(43).random((m, d))
        ctx.predict_valvar(xs[:256])
        ctx.set_profiling(True)
        ctx.reset_profile()
        t0 = time.perf_counter()
        yv = ctx.predict_valvar(xs)
        t1 = time.perf_counter()
        prof = ctx.profile()
        t2 = time.perf_counter()
        ctx.predict_valvar(xs)
        t3 = time.perf_counter()
        print(json.dumps({"n": n, "m": m, "predict_valvar_wall_ms": (t1 - t0) * 1e3,
                          "predict_valvar_wall_ms_second_call": (t3 - t2) * 1e3,
                          "stage_ms": {k: round(v[0], 4) for k, v in prof.items()},
                          "var_min": float(yv[1].min()), "var_max": float(yv[1].max())}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
