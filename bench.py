#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native kriging hot path.

Metric (BASELINE.json): GP fit+predict throughput in points/s at n=8192, d=10.
Workload (BASELINE.json configs[1]): Kriging Matern-5/2, n=8192, d=10, fp64, constant mean;
one STEP = one fit with the reference's default optimisation budget (n_start=10 ->
11 chains x clamp(10*d, 25, max_eval)=100 likelihood evaluations + 1 final = 1101
evaluations, gp/src/algorithm.rs:33-37, 936-937) followed by predict_var on m=100000 points.
    points/s = m / (t_fit + t_predict_var)

 value : the same work with the training set, the theta sequence and x* already resident
         in HBM (C ABI device-pointer entry points), CUDA-event timed.
 e2e   : the reference-facing call a user makes -- GaussianProcess.params(..).fit(x, y)
         then predict_var(x*) -- with host buffers (x* pinned), H2D/D2H inside the region.
 N > 1 : one process per GPU (torchrun), STRONG scaling of the same step on ONE training set: the
         likelihood evaluations of the fit are sharded round-robin over the ranks (the rayon
         multistart + `reduce` by min of gp/src/algorithm.rs:928-945 across GPUs), one NCCL
         all-gather finds the best theta, every rank finalises there and predicts its slice of the
         m points, one NCCL all-gather assembles the variances -- both collectives INSIDE the timed
         region.  Two more legs are timed the same way and reported beside the headline:
         c5 = the 512-candidate theta sweep (n = 2048, d = 20; BASELINE configs[4]) and
         c4 = 8 experts x n = 4096, d = 20 fitted one per rank (configs[3], moe/src/algorithm.rs:165-177).
 --impl reference : the CPU path (oracle restatement: C/OpenMP correlation + LAPACK through
         scipy, all host threads) on a bounded sample of the same step (a few likelihood
         evaluations + a prediction chunk, actually run every step), scaled to the full step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CORR_MATERN52, MEAN_CONSTANT = 3, 0


# ----------------------------------------------------------------------------- workload
def lhs(n, d, seed):
    rng = np.random.default_rng(seed)
    u = rng.random((n, d))
    pts = (np.arange(n)[:, None] + u) / n
    for j in range(d):
        pts[:, j] = pts[rng.permutation(n), j]
    return pts


def rosenbrock(x):
    z = 4.0 * x - 2.0
    return np.sum(100.0 * (z[:, 1:] - z[:, :-1] ** 2) ** 2 + (1.0 - z[:, :-1]) ** 2, axis=1)


def make_workload(n, d, m, evals, rank):
    x = lhs(n, d, 42 + 1000 * rank)
    y = rosenbrock(x)
    xs = np.random.default_rng(43 + 1000 * rank).random((m, d))
    # theta sequence: log10-uniform LHS in the default bounds [1e-2, 10] (parameters.rs:51)
    thetas = 10.0 ** (-2.0 + 3.0 * lhs(evals, d, 44))
    return x, y, xs, thetas


def normalize(a):
    mean = a.mean(axis=0)
    std = a.std(axis=0, ddof=1)
    std = np.where(std == 0.0, 1.0, std)
    return (a - mean) / std, mean, std


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smmax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.3:
                continue
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smmax,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_sample(n, d, m, evals, sample_pts=2048, sample_evals=4):
    """Bounded sample of one step on the host cores, actually run: `sample_evals` likelihood evaluations at full n (different
    thetas of the step's own sequence) plus predict_valvar on `sample_pts` points; the step time is that sample scaled to
    `evals` evaluations and m points."""
    from oracle import fast, gp_oracle as O
    cores = len(os.sched_getaffinity(0))
    # all host threads, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)
    omp_threads = fast.set_threads(cores)
    blas_threads = None
    try:
        import scipy.linalg  # noqa: F401  (loads OpenBLAS before the pool is resized)
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=cores)
        blas_threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        pass
    x, y, xs, thetas = make_workload(n, d, sample_pts, max(evals - 1, sample_evals), 0)
    xn, xm, xsd = normalize(x)
    yn, ym, ysd = normalize(y.reshape(-1, 1))
    fx = O.mean_value(O.CONSTANT, xn)
    theta = np.full(d, 1.0)
    nw = min(n, 1024)       # warm the OpenMP / BLAS thread pools outside the timed sample
    fast.reduced_likelihood(O.MATERN52, xn[:nw], fx[:nw], yn[:nw], float(ysd[0]), theta, np.eye(d))
    t_start = time.perf_counter()
    n_ok = 0
    for k in range(sample_evals):
        try:
            fast.reduced_likelihood(O.MATERN52, xn, fx, yn, float(ysd[0]), thetas[k], np.eye(d))
            n_ok += 1
        except Exception:       # a non positive definite candidate costs its Cholesky attempt, like Err(_) -> +inf in the reference
            pass
    t_evals = time.perf_counter() - t_start
    rlf, inner = fast.reduced_likelihood(O.MATERN52, xn, fx, yn, float(ysd[0]), theta, np.eye(d))
    gp = O.GaussianProcess(corr=O.MATERN52, mean=O.CONSTANT, theta=theta, likelihood=rlf, inner=inner,
                           w_star=np.eye(d), xt_norm=xn, x_mean=xm, x_std=xsd, yt_norm=yn,
                           y_mean=float(ym[0]), y_std=float(ysd[0]))
    t0 = time.perf_counter()
    fast.predict_valvar(gp, xs, chunk=min(sample_pts, 1024))
    t_chunk = time.perf_counter() - t0
    t_eval = t_evals / sample_evals
    t_fit = evals * t_eval
    t_pred = (m / sample_pts) * t_chunk
    return {"value": m / (t_fit + t_pred), "unit": "points/s", "cores": cores, "blas_threads": blas_threads,
            "omp_threads": omp_threads,
            "kind": "port",
            "sample": "%d likelihood evals at n=%d (%.2f s each, step = x%d) + predict_var on %d points (%.2f s, step = x%.1f), "
                      "all run; oracle port: C/OpenMP correlation (-O3 -march=native) + scipy LAPACK" % (
                          sample_evals, n, t_eval, evals, sample_pts, t_chunk, m / sample_pts),
            "sample_evals": sample_evals, "sample_points": sample_pts, "sample_evals_ok": n_ok,
            "sample_wall_s": t_evals + t_chunk,
            "t_eval_s": t_eval, "t_predict_chunk_s": t_chunk, "rlf": rlf,
            "full_step_s_scaled_from_sample": t_fit + t_pred}


def cpu_sample_clean_env(n, d, m, evals, flag="--cpu-sample-only"):
    """Run cpu_sample in a child process whose environment does not pin the thread pools to one thread
    (torchrun exports OMP_NUM_THREADS=1 and OpenBLAS sizes its pool from it at load time)."""
    env = {k: v for k, v in os.environ.items()
           if k not in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "GOTO_NUM_THREADS")}
    cmd = [sys.executable, os.path.abspath(__file__), flag, "--ntrain", str(n), "--dim", str(d),
           "--npred", str(m), "--evals", str(evals)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    for line in reversed(r.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("cpu sample failed: %s" % r.stderr[-2000:])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, last = [], None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        last = cpu_sample_clean_env(args.n, args.d, args.m, args.evals)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    val = last["value"]
    out = {"impl": "reference", "metric": "GP fit+predict throughput (points/s) at n=%d d=%d" % (args.n, args.d),
           "value": val, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           # what one timed step of THIS run is: the bounded sample (child process start-up included); the full step it
           # stands for would take full_step_ms_scaled_from_sample on these cores
           "ms_per_step": 1e3 * float(np.mean(times)),
           "full_step_ms_scaled_from_sample": 1e3 * args.m / val,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "same_config_as_gpu_arm": True,          # at every N the GPU arm works on ONE n=8192 problem, like this arm
           "config": workload_config(args), "cpu_baseline": last,
           "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(args):
    return {"workload": "Kriging Matern52 n=%d d=%d fp64 constant mean: fit (%d likelihood evals = 11 chains x 100 "
                        "+ final) + predict_var on m=%d points" % (args.n, args.d, args.evals, args.m),
            "n": args.n, "d": args.d, "m": args.m, "likelihood_evals_per_fit": args.evals,
            "l2": "inputs larger than L2 (R/L workspace %.0f MB, predict chunk >= %.0f MB vs 126 MB L2)" % (
                8e-6 * args.n * args.n, 8e-6 * min(args.m, 8192) * args.n),
            "parallelism": "likelihood evaluations and prediction points of ONE training set sharded over the GPUs; "
                           "collectives: one all-gather of (status, likelihood) per sweep, one all-gather of the variances"}


# ----------------------------------------------------------------------------- GPU arm
def gemm_algorithmic_flops(n, evals, m, chunk=8192, nb=128, q=2):
    """Useful flops executed by the K4 GEMM launches of one step (SURVEY 8d figures restated per launch)."""
    T = -(-n // nb)
    fl = 0.0
    for k in range(T - 1):
        r = max(n - (k + 1) * nb, 0)
        fl += 2.0 * nb * (r * (r + 1) / 2.0 + q * r)
    chol = fl * (evals + 1)
    pred = 0.0
    for i0 in range(0, m, chunk):
        mc = min(chunk, m - i0)
        for k in range(T - 1):
            pred += 2.0 * nb * mc * max(n - (k + 1) * nb, 0)
    return chol + pred


def ozaki_algorithmic_flops(n, nfact, nb=128, q=2, min_tri=8):
    """Useful fp64 flops of the tcgen05 (int8-sliced) trailing-update launches of `nfact` factorisations run WITHOUT
    look-ahead (the roofline pass): pair step k updates the lower triangle of the (T-k-2) trailing block columns plus
    the q appended right-hand-side rows with K = 256; steps with fewer than `min_tri` block columns stay on DMMA
    (csrc/sweep.cu::trailing_syrk).  Returns (flops, launches, tiles)."""
    T = -(-n // nb)
    fl, launches, tiles = 0.0, 0, 0
    for k in range(0, T, 2):
        tri = T - k - 2
        if tri < min_tri:
            continue
        r = min(tri * nb, max(n - (k + 2) * nb, 0))
        fl += 2.0 * (2 * nb) * (r * (r + 1) / 2.0 + q * r)
        launches += 1
        tiles += tri * (tri + 1) // 2 + tri
    return fl * nfact, launches * nfact, tiles * nfact


def ozaki_predict_flops(n, m, chunk=8192, nb=128, min_tri=8):
    """Useful flops of the tcgen05 launches of the multi-RHS solve of predict_var on `m` points: pair step k updates the
    (T-k-2) block columns right of the pair with K = 256 (chunks of >= 8 row tiles, >= min_tri block columns; the
    partner-column updates and the rest stay on DMMA).  Returns (flops, launches, tiles)."""
    T = -(-n // nb)
    fl, launches, tiles = 0.0, 0, 0
    for i0 in range(0, m, chunk):
        mc = min(chunk, m - i0)
        row_tiles = -(-mc // nb)
        if row_tiles < 8:
            continue
        for k in range(0, T, 2):
            tri = T - k - 2
            if tri < min_tri:
                continue
            fl += 2.0 * (2 * nb) * mc * min(tri * nb, max(n - (k + 2) * nb, 0))
            launches += 1
            tiles += row_tiles * tri
    return fl, launches, tiles


def run_ours(args):
    import torch
    import torch.distributed as dist
    import egobox_b200 as eg
    from egobox_b200 import parallel as P

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if eg.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n, d, m, E = args.n, args.d, args.m, args.evals
    # ONE training set, ONE theta sequence and ONE set of prediction points for the whole job, whatever the number of GPUs
    x, y, xs, thetas = make_workload(n, d, m, E - 1, 0)
    xn, xm, xsd = normalize(x)
    yn, ym, ysd = normalize(y.reshape(-1, 1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value leg: everything resident in HBM --------------------------------
    # every rank holds a replica of the training set (0.7 MB) and its own R / L workspaces
    ctx = eg.GpContext(xn, yn[:, 0], xm, xsd, float(ym[0]), float(ysd[0]), eg.MATERN52, eg.CONSTANT,
                       device=local_rank)
    bounds = [(m * r) // world for r in range(world + 1)]
    lo, hi = bounds[rank], bounds[rank + 1]
    cap = max(bounds[r + 1] - bounds[r] for r in range(world))
    xs_dev = torch.from_numpy(np.ascontiguousarray(xs[lo:hi])).cuda()
    y_loc = torch.zeros(cap, dtype=torch.float64, device="cuda")
    v_loc = torch.zeros(cap, dtype=torch.float64, device="cuda")
    v_all = torch.zeros(cap * world, dtype=torch.float64, device="cuda") if world > 1 else v_loc
    split = {"fit_ms": 0.0, "predict_ms": 0.0}
    state = {}

    def device_step():
        # fit: the E - 1 candidate evaluations sharded round-robin, ONE all-gather of (status, likelihood) per rank
        # (parallel.theta_sweep; NCCL for world > 1), the `reduce` by min, the final evaluation at the winner on EVERY rank
        def evaluate(th):
            ctx.timer_start()
            out = ctx.reduced_likelihood_batch(th)
            split["fit_ms"] += ctx.timer_stop()
            return out
        status, rlf, best = P.theta_sweep(evaluate, thetas)
        ctx.timer_start()
        st, _ = ctx.finalize(thetas[best], want_ft=False)
        split["fit_ms"] += ctx.timer_stop()
        assert st == 0
        # predict_var: the m points sharded in contiguous slices, ONE all-gather of the variances
        ctx.timer_start()
        ctx.predict_valvar_dev(xs_dev.data_ptr(), hi - lo, y_loc.data_ptr(), v_loc.data_ptr())
        split["predict_ms"] += ctx.timer_stop()
        if world > 1:
            dist.all_gather_into_tensor(v_all, v_loc)
        state.update(status=status, rlf=rlf, best=best)

    for _ in range(args.warmup):
        device_step()
    ctx.set_profiling(False)     # production path: evaluations replayed as CUDA graphs, no per-launch events
    ctx.reset_profile()          # launch counters are kept either way
    split["fit_ms"] = split["predict_ms"] = 0.0
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_wall0 = time.time()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        device_step()
    barrier()
    step_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    # The kernels run on the context's own stream, which torch events do not see: the step time is taken between two
    # (barrier + device synchronize) brackets, i.e. it ends when the last kernel and the last collective of the last step
    # have completed on every rank; the CUDA-event time of the kernel regions on the context's stream (egx_gp_timer_*)
    # is reported beside it (device_ms_per_step: no collectives, no host gaps between regions).
    dev_ms = (split["fit_ms"] + split["predict_ms"]) / args.steps
    prof = ctx.profile()
    n_fail = int(np.sum(state["status"] != 0))
    launches = sum(v[1] for v in prof.values())
    theta_fin = thetas[state["best"]]

    # ---------------- roofline pass for the dominant kernel -----------------
    # The timed region keeps several evaluations in flight on separate streams (graph replays), where a CUDA-event
    # bracket around one launch would also contain the time it queued behind other streams.  The per-kernel
    # figure is therefore taken right after it, on the same context and data, with the launches back to back on
    # ONE stream (look-ahead off, no batch concurrency, per-launch events on): 3 likelihood evaluations +
    # finalize + predict_var on one 8192-point chunk.
    ctx.set_lookahead(False)
    ctx.set_profiling(True)
    ctx.reset_profile()
    roof_evals, roof_pts = 3, min(hi - lo, 8192)
    for _ in range(roof_evals):
        ctx.reduced_likelihood(theta_fin)
    ctx.finalize(theta_fin, want_ft=False)
    ctx.predict_valvar_dev(xs_dev.data_ptr(), roof_pts, y_loc.data_ptr(), v_loc.data_ptr())
    roof_prof = ctx.profile()
    ctx.set_lookahead(True)
    ctx.set_profiling(False)
    ctx.close()

    # ---------------- e2e leg: public API, host buffers ------------------------------------
    xs_pinned = torch.from_numpy(xs).pin_memory().numpy()
    x_h, y_h = np.ascontiguousarray(x), np.ascontiguousarray(y)

    e2e_parts = []        # (fit, predict, close) wall ms of every e2e step of this rank, warm-up included

    def e2e_step():
        # same evaluation budget as the value leg: (n_start + 1) chains x clamp(10 d, 25, 1000) + 1 final; the chains are
        # sharded over the ranks (parallel.fit_multistart: chain c on rank c % world, one all-gather of (objective, theta)),
        # every rank finalises at the winner and predicts its slice of the points (parallel.predict_sharded: one all-gather)
        per_chain = min(max(10 * d, 25), 1000)
        params = (eg.GaussianProcess.params(eg.ConstantMean, eg.Matern52Corr)
                  .n_start(max((E - 1) // per_chain - 1, 0)).max_eval(1000)
                  .cobyla_ftol_rel(0.0).device(local_rank))
        t_a = time.perf_counter()
        gp = P.fit_multistart(params, x_h, y_h)
        t_b = time.perf_counter()
        var = P.predict_sharded(gp.predict_var, xs_pinned)
        t_c = time.perf_counter()
        nev = gp.n_evals()
        lik, th = gp.likelihood(), gp.theta()
        gp.close()
        t_d = time.perf_counter()
        e2e_parts.append(((t_b - t_a) * 1e3, (t_c - t_b) * 1e3, (t_d - t_c) * 1e3))
        return var, nev, lik, th

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(args.warmup, 1)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        var, nev, lik, th = e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps

    # ---------------- the other multi-GPU configurations of BASELINE.json ----------------------------
    extra = {}
    if not args.no_extra:
        if world == 1:
            extra["c3_sparse_gp"] = leg_c3(args, eg)
        extra["c5_theta_sweep"] = leg_c5(args, eg, P, torch, dist, local_rank, world, barrier)
        extra["c4_moe_experts"] = leg_c4(args, eg, P, torch, dist, local_rank, rank, world, barrier)

    # ---------------- reductions over ranks -------------------------------------------------
    step_ms_all, e2e_ms_all, dev_ms_all, nev_all = step_ms, e2e_ms, dev_ms, float(nev)
    collective_us = None
    if world > 1:
        # what the exchanges of the sharded path cost on this box: the (status, likelihood) all-gather of the sweep through
        # parallel.all_gather_rows (host rows -> device -> NCCL -> host), as it runs inside the timed region, wall time
        rows = np.zeros((len(P.shard_indices(E - 1, rank, world)), 2))
        cnts = [len(P.shard_indices(E - 1, r, world)) for r in range(world)]
        for _ in range(5):
            P.all_gather_rows(rows, 2, cnts)
        barrier()
        t0c = time.perf_counter()
        for _ in range(20):
            P.all_gather_rows(rows, 2, cnts)
        torch.cuda.synchronize()
        collective_us = (time.perf_counter() - t0c) / 20 * 1e6
        t = torch.tensor([step_ms, e2e_ms, dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms_all, e2e_ms_all, dev_ms_all = float(t[0]), float(t[1]), float(t[2])
        t = torch.tensor([float(nev - 1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        nev_all = float(t[0]) + 1.0

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback (sustained 1400)"
        sm_clock = clocks.get("sm_mhz") or 1965.0
        fp64_peak_at_clock = 148 * 64 * 2 * sm_clock * 1e6 / 1e12
        # int8 rate of the tensor pipe: one M = 128, N = 256, K = 32 MMA per 128.4 clk per SM, MEASURED on this pool
        # (profiles/r02/pattern_probe.txt, modes 2 / 5: 1027 clk per 8 MMAs) = 8166 MAC/clk/SM, at the clock sampled under load
        int8_mac_per_clk_sm = 128.0 * 256 * 32 * 8 / 1027.0
        int8_peak_at_clock = 148 * int8_mac_per_clk_sm * 2 * sm_clock * 1e6 / 1e12
        roof_total_ms = sum(v[0] for v in roof_prof.values())
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        except Exception:
            pass
        measured_in = ("roofline pass after the timed region: %d evaluations + predict_var(%d) with launches back to "
                       "back on one stream" % (roof_evals + 1, roof_pts))
        # the DMMA kernel (look-ahead / partner-column / small trailing updates, multi-RHS solve of predict_var)
        gemm_ms, gemm_launches = roof_prof["syrk_gemm"]
        oz_ms, oz_launches = roof_prof.get("ozaki_syrk", (0.0, 0))
        oz_flops, _, oz_tiles = ozaki_algorithmic_flops(n, roof_evals + 1) if oz_launches else (0.0, 0, 0)
        if oz_launches:
            pf, _, pt = ozaki_predict_flops(n, roof_pts)
            oz_flops, oz_tiles = oz_flops + pf, oz_tiles + pt
        dmma_flops = gemm_algorithmic_flops(n, roof_evals, roof_pts) - oz_flops
        dmma_achieved = dmma_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        dmma = {"kernel": "gemm_nt_sub_kernel (DMMA fp64: partner-column / look-ahead / small trailing updates)",
                "achieved": dmma_achieved, "unit": "TFLOP/s", "launches": gemm_launches,
                "avg_launch_ms": gemm_ms / max(gemm_launches, 1),
                "share_of_kernel_time_in_roofline_pass": gemm_ms / max(roof_total_ms, 1e-9),
                "fp64_pipe_peak_tflops_at_sampled_clock": fp64_peak_at_clock,
                "frac_of_fp64_pipe": (dmma_achieved / fp64_peak_at_clock) if dmma_achieved else None,
                "traffic": ncu.get("gemm_nt_sub_kernel", {}).get("dram_bytes_per_launch")}
        if oz_launches:
            # dominant kernel: the trailing SYRK update on tcgen05 (UTCIMMA int8, TMEM accumulators)
            achieved = oz_flops / (oz_ms * 1e-3) / 1e12
            int8_tops = 28.0 * oz_tiles * 2.0 * 128 * 128 * 256 / (oz_ms * 1e-3) / 1e12    # executed int8 ops
            roofline = {"bound": "tensor", "kernel": "ozaki_syrk5_kernel (tcgen05.mma kind::i8 on 7 balanced base-256 "
                                                     "digit slices of the fp64 operands; fp64 result): trailing updates "
                                                     "of the factorisations + multi-RHS solve updates of predict_var",
                        "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s",
                        "frac": achieved / bf16_peak, "peak_source": peak_src,
                        "traffic": ncu.get("ozaki_syrk5_kernel", ncu.get("ozaki_syrk_kernel", {})).get("dram_bytes_per_launch"),
                        "launches": oz_launches, "avg_launch_ms": oz_ms / max(oz_launches, 1),
                        "measured_in": measured_in,
                        "share_of_kernel_time_in_roofline_pass": oz_ms / max(roof_total_ms, 1e-9),
                        "executed_int8_tops": int8_tops,
                        "int8_peak_tops_at_sampled_clock": int8_peak_at_clock,
                        "int8_peak_source": "148 SM x 8166 int8 MAC/clk/SM (measured: profiles/r02/pattern_probe.txt) x 2 x the "
                                            "SM clock sampled under load (%.0f MHz)" % sm_clock,
                        "tensor_pipe_frac": int8_tops / int8_peak_at_clock,
                        "fp64_pipe_peak_tflops_at_sampled_clock": fp64_peak_at_clock,
                        "speedup_over_fp64_pipe_peak": achieved / fp64_peak_at_clock,
                        "note": "`achieved` counts the useful fp64 flops of the update (2 K per entry of the lower "
                                "triangle), `peak` is the mandated bf16 figure, so `frac` compares fp64 work with a bf16 "
                                "rate; each fp64 multiply-add costs 28 int8 multiply-adds on the tensor core "
                                "(executed_int8_tops): tensor_pipe_frac = executed int8 rate / the int8 rate of the pipe, "
                                "the figure ncu reports as sm__pipe_tensor_subpipe_imma_cycles_active; for scale: the fp64 "
                                "(DMMA / FMA) pipe of the chip peaks at fp64_pipe_peak_tflops_at_sampled_clock",
                        "dmma_kernel": dmma}
        else:
            roofline = dict(dmma, bound="tensor", peak=bf16_peak, peak_source=peak_src,
                            frac=(dmma_achieved / bf16_peak) if dmma_achieved else None, measured_in=measured_in,
                            note="EGX_OZAKI=0: fp64 contraction on the DMMA pipe; the bf16 figure is the mandated "
                                 "denominator, the fp64-pipe line (148 SM x 64 FMA/clk) is the physical bound")
        # correlation build (K1): algorithmic bytes = lower 128-block triangle written once + X read once
        k1_ms, k1_launches = roof_prof.get("corr_build", (0.0, 0))
        if k1_launches:
            T = -(-n // 128)
            k1_bytes = 8.0 * 128 * 128 * T * (T + 1) / 2 + 8.0 * n * d
            k1_gbs = k1_bytes / (k1_ms / k1_launches * 1e-3) / 1e9
            roofline["corr_build_kernel"] = {
                "avg_launch_ms": k1_ms / k1_launches, "algorithmic_bytes": k1_bytes, "achieved_gbs": k1_gbs,
                "hbm_frac": k1_gbs / peaks.get("hbm_gbs", 6650.0),
                "pairs_per_s": n * (n + 1) / 2.0 / (k1_ms / k1_launches * 1e-3),
                "note": "Matern-5/2 at d = 10 is bound by fp64 issue, not by HBM (DESIGN.md section 4)"}
        value = m / (step_ms_all * 1e-3)
        e2e_value = m / (e2e_ms_all * 1e-3)
        out = {"metric": "GP fit+predict throughput (points/s) at n=%d d=%d" % (n, d),
               "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": step_ms_all, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": workload_config(args),
               "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": e2e_ms_all, "steps": e2e_steps,
                       "likelihood_evals": int(nev_all),
                       "h2d_bytes_per_step": int(world * (x_h.nbytes + y_h.nbytes) + xs_pinned.nbytes + nev_all * d * 8),
                       "d2h_bytes_per_step": int(var.nbytes + nev_all * 64),
                       "fit_predict_close_ms_rank0": [[round(v, 1) for v in p3] for p3 in e2e_parts]},
               "sweep_allgather_us_rank0": collective_us,
               "gpu_launches": int(launches) * world,
               "launches_per_step_rank0": {k: int(v[1] // args.steps) for k, v in prof.items()},
               "stage_ms_roofline_pass": {k: round(v[0], 3) for k, v in roof_prof.items()},
               "device_ms_per_step": dev_ms_all,
               "fit_ms_per_step_rank0": split["fit_ms"] / args.steps, "predict_ms_per_step_rank0": split["predict_ms"] / args.steps,
               "likelihood_evals_per_s": E / (step_ms_all * 1e-3 - split["predict_ms"] / args.steps * 1e-3),
               "failed_theta_in_sweep": n_fail,
               "roofline": roofline}
        out.update(extra)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_sample_clean_env(n, d, m, E)
            if "c3_sparse_gp" in out:
                try:
                    out["c3_sparse_gp"]["cpu_baseline"] = cpu_sample_clean_env(n, d, m, E, flag="--cpu-c3-only")
                    out["c3_sparse_gp"]["speedup_vs_cpu_port"] = (out["c3_sparse_gp"]["cpu_baseline"]["ms_per_likelihood_eval_scaled"]
                                                                  / out["c3_sparse_gp"]["ms_per_likelihood_eval"])
                except Exception as exc:       # the secondary baseline must not take the headline line down
                    out["c3_sparse_gp"]["cpu_baseline"] = {"error": str(exc)[:200]}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def c3_inputs(N, d, M):
    rng = np.random.default_rng(42)
    x = rng.random((N, d))
    y = np.sum(np.sin(3 * np.pi * x), axis=1) + rng.normal(0, 0.1, N)
    z = x[rng.permutation(N)[:M]].copy()
    return x, y, z


def leg_c3(args, eg):
    """BASELINE configs[2]: sparse GP (FITC), N = 100000, d = 6, M = 1024 inducing points, one GPU: likelihood evaluations
    (what the fit driver repeats) and predict_var on 100000 points, through the device seam egx_sgp_*."""
    N, d, M = 100000, 6, 1024
    x, y, z = c3_inputs(N, d, M)
    ctx = eg.SgpContext(x, y, z, corr=eg.MATERN52, method=eg.SparseMethod.FITC)
    theta = np.full(d, 1.0)
    for _ in range(2):
        ctx.reduced_likelihood(theta, 1.0, 0.01)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        st, lik = ctx.reduced_likelihood(theta, 1.0, 0.01)
    ms = (time.perf_counter() - t0) * 1e3 / reps
    ctx.finalize(theta, 1.0, 0.01)
    xs = np.random.default_rng(43).random((100000, d))
    ctx.predict_var(xs[:1024])
    t0 = time.perf_counter()
    v = ctx.predict_var(xs)
    pms = (time.perf_counter() - t0) * 1e3
    ctx.close()
    return {"N": N, "d": d, "M": M, "method": "FITC", "status": int(st), "likelihood": float(lik), "ms_per_likelihood_eval": ms,
            "likelihood_evals_per_s": 1e3 / ms, "algorithmic_flops_2M2N": 2.0 * M * M * N,
            "tflops_fp64_equivalent": 2.0 * M * M * N / (ms * 1e-3) / 1e12, "predict_var_100k_ms": pms,
            "predict_var_points_per_s": 1e5 / (pms * 1e-3), "var_mean": float(v.mean())}


def cpu_c3_sample(Ns=20000):
    """The oracle's FITC likelihood (C/OpenMP kernel + LAPACK, all host threads) on the first Ns points of the C3 inputs, the
    same 1024 inducing points; the cost is linear in N (2 M^2 N flops), so the full evaluation is the sample x N / Ns."""
    from oracle import fast, gp_oracle as O, sgp_oracle as S
    cores = len(os.sched_getaffinity(0))
    fast.set_threads(cores)
    try:
        import scipy.linalg  # noqa: F401
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    N, d, M = 100000, 6, 1024
    x, y, z = c3_inputs(N, d, M)
    S.USE_FAST_KERNEL = True
    theta = np.full(d, 1.0)
    S.reduced_likelihood(S.FITC, O.MATERN52, theta, 1.0, 0.01, np.eye(d), x[:2000], y[:2000], z)       # warm the pools
    t0 = time.perf_counter()
    lik, _ = S.reduced_likelihood(S.FITC, O.MATERN52, theta, 1.0, 0.01, np.eye(d), x[:Ns], y[:Ns], z)
    t = time.perf_counter() - t0
    return {"kind": "port", "cores": cores, "sample": "FITC likelihood on %d of the %d points (%.2f s), scaled x%.1f" % (Ns, N, t, N / Ns),
            "sample_wall_s": t, "ms_per_likelihood_eval_scaled": 1e3 * t * N / Ns, "likelihood_of_sample": float(lik)}


def leg_c5(args, eg, P, torch, dist, local_rank, world, barrier):
    """BASELINE configs[4]: a 512-candidate reduced-likelihood sweep on ONE training set (d = 20), candidates sharded over the
    ranks, NCCL all-gather + argmin inside the timed region (parallel.theta_sweep).  Strong scaling."""
    out = {}
    for nn in (2048,):
        dd, B = 20, 512
        x = lhs(nn, dd, 7)
        y = rosenbrock(x)
        xn, xm, xsd = normalize(x)
        yn, ym, ysd = normalize(y.reshape(-1, 1))
        thetas = 10.0 ** (-2.0 + 3.0 * lhs(B, dd, 8))
        ctx = eg.GpContext(xn, yn[:, 0], xm, xsd, float(ym[0]), float(ysd[0]), eg.MATERN52, eg.CONSTANT, device=local_rank)
        for _ in range(3):
            P.theta_sweep(ctx.reduced_likelihood_batch, thetas)
        reps = 5
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            status, rlf, best = P.theta_sweep(ctx.reduced_likelihood_batch, thetas)
        barrier()
        ms = (time.perf_counter() - t0) * 1e3 / reps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        ctx.close()
        out["n%d" % nn] = {"n": nn, "d": dd, "candidates": B, "ms_per_sweep": ms, "likelihood_evals_per_s": B / (ms * 1e-3),
                           "best_index": int(best), "failed": int(np.sum(status != 0)), "scaling": "strong",
                           "collective": "all_gather of (status, likelihood) + argmin, inside the timed region"}
    return out


def leg_c4(args, eg, P, torch, dist, local_rank, rank, world, barrier):
    """BASELINE configs[3]: 8 experts x n = 4096, d = 20, one GP fit per expert (default budget: 11 chains x 200 + 1
    evaluations), expert e on rank e % world (parallel.fit_experts; moe/src/algorithm.rs:165-177), the table of
    (likelihood, variance, theta) all-gathered inside the timed region.  Strong scaling: N = 1 fits all 8 one after the other."""
    ne, nn, dd = 8, 4096, 20
    centres = 10.0 * lhs(ne, dd, 42)

    def data(e):
        xe = centres[e] + np.random.default_rng(100 + e).random((nn, dd)) - 0.5
        return xe, rosenbrock(xe / 10.0)

    def fit_one(e):
        xe, ye = data(e)
        gp = eg.GaussianProcess.params(eg.ConstantMean, eg.Matern52Corr).device(local_rank).fit(xe, ye)
        return gp, gp.likelihood(), gp.variance(), gp.theta()
    # warm-up: one short fit per rank (workspaces, graph captures)
    xe, ye = data(rank % ne)
    eg.GaussianProcess.params(eg.ConstantMean, eg.Matern52Corr).n_start(0).max_eval(25).device(local_rank).fit(xe, ye).close()
    barrier()
    t0 = time.perf_counter()
    models, table = P.fit_experts(fit_one, ne, dd)
    barrier()
    ms = (time.perf_counter() - t0) * 1e3
    evals = float(sum(g.n_evals() for g in models.values()))
    for g in models.values():
        g.close()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        t = torch.tensor([evals], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        evals = float(t[0])
    return {"experts": ne, "n": nn, "d": dd, "ms_per_mixture_fit": ms, "experts_per_s": ne / (ms * 1e-3),
            "likelihood_evals": int(evals), "likelihood_evals_per_s": evals / (ms * 1e-3), "scaling": "strong",
            "min_likelihood": float(table[:, 0].min()),
            "collective": "all_gather of (likelihood, variance, theta) per expert, inside the timed region"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ntrain", dest="n", type=int, default=8192)
    ap.add_argument("--dim", dest="d", type=int, default=10)
    ap.add_argument("--npred", dest="m", type=int, default=100000)
    ap.add_argument("--evals", type=int, default=1101)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the c5 / c4 legs")
    ap.add_argument("--cpu-sample-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-c3-only", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_sample_only:
        print(json.dumps(cpu_sample(args.n, args.d, args.m, args.evals)), flush=True)
        return
    if args.cpu_c3_only:
        print(json.dumps(cpu_c3_sample()), flush=True)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
