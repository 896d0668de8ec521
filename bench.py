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
 N > 1 : one process per GPU (torchrun), weak scaling: every rank fits and predicts its own
         n=8192 expert (the MoE expert loop moe/src/algorithm.rs:167-177 sharded over GPUs);
         the only collective is the final NCCL all-gather of (likelihood, theta) per expert.
 --impl reference : the CPU path (oracle restatement: C/OpenMP correlation + LAPACK through
         scipy, all host threads) on a bounded sample of the same step, extrapolated.
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
def cpu_sample(n, d, m, evals, sample_pts=1024):
    """Bounded sample of one step on the host cores: ONE likelihood evaluation at full n plus
    predict_valvar on `sample_pts` points, extrapolated to `evals` evaluations and m points."""
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
    x, y, xs, thetas = make_workload(n, d, sample_pts, 4, 0)
    xn, xm, xsd = normalize(x)
    yn, ym, ysd = normalize(y.reshape(-1, 1))
    fx = O.mean_value(O.CONSTANT, xn)
    theta = np.full(d, 1.0)
    nw = min(n, 1024)       # warm the OpenMP / BLAS thread pools outside the timed sample
    fast.reduced_likelihood(O.MATERN52, xn[:nw], fx[:nw], yn[:nw], float(ysd[0]), theta, np.eye(d))
    t0 = time.perf_counter()
    rlf, inner = fast.reduced_likelihood(O.MATERN52, xn, fx, yn, float(ysd[0]), theta, np.eye(d))
    t_eval = time.perf_counter() - t0
    gp = O.GaussianProcess(corr=O.MATERN52, mean=O.CONSTANT, theta=theta, likelihood=rlf, inner=inner,
                           w_star=np.eye(d), xt_norm=xn, x_mean=xm, x_std=xsd, yt_norm=yn,
                           y_mean=float(ym[0]), y_std=float(ysd[0]))
    t0 = time.perf_counter()
    fast.predict_valvar(gp, xs, chunk=sample_pts)
    t_chunk = time.perf_counter() - t0
    t_fit = evals * t_eval
    t_pred = (m / sample_pts) * t_chunk
    return {"value": m / (t_fit + t_pred), "unit": "points/s", "cores": cores, "blas_threads": blas_threads,
            "omp_threads": omp_threads,
            "kind": "port",
            "sample": "1 likelihood eval at n=%d (%.2f s, x%d) + predict_var on %d points (%.2f s, x%.1f); "
                      "oracle port: C/OpenMP correlation + scipy LAPACK" % (n, t_eval, evals, sample_pts, t_chunk,
                                                                            m / sample_pts),
            "t_eval_s": t_eval, "t_predict_chunk_s": t_chunk, "rlf": rlf}


def cpu_sample_clean_env(n, d, m, evals):
    """Run cpu_sample in a child process whose environment does not pin the thread pools to one thread
    (torchrun exports OMP_NUM_THREADS=1 and OpenBLAS sizes its pool from it at load time)."""
    env = {k: v for k, v in os.environ.items()
           if k not in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "GOTO_NUM_THREADS")}
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-sample-only", "--ntrain", str(n), "--dim", str(d),
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
           "ms_per_step": 1e3 * args.m / val, "sample_wall_ms_per_step": 1e3 * float(np.mean(times)),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args), "cpu_baseline": last,
           "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(args):
    return {"workload": "Kriging Matern52 n=%d d=%d fp64 constant mean: fit (%d likelihood evals = 11 chains x 100 "
                        "+ final) + predict_var on m=%d points" % (args.n, args.d, args.evals, args.m),
            "n": args.n, "d": args.d, "m": args.m, "likelihood_evals_per_fit": args.evals,
            "l2": "inputs larger than L2 (R/L workspace %.0f MB, predict chunk %.0f MB vs 126 MB L2)" % (
                8e-6 * args.n * args.n, 8e-6 * min(args.m, 8192) * args.n),
            "parallelism": "1 expert per GPU, no data-path collective"}


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

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if eg.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n, d, m, E = args.n, args.d, args.m, args.evals
    x, y, xs, thetas = make_workload(n, d, m, E - 1, rank)
    xn, xm, xsd = normalize(x)
    yn, ym, ysd = normalize(y.reshape(-1, 1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value leg: everything resident in HBM --------------------------------
    ctx = eg.GpContext(xn, yn[:, 0], xm, xsd, float(ym[0]), float(ysd[0]), eg.MATERN52, eg.CONSTANT,
                       device=local_rank)
    xs_dev = torch.from_numpy(xs).cuda()
    y_dev = torch.empty(m, dtype=torch.float64, device="cuda")
    v_dev = torch.empty(m, dtype=torch.float64, device="cuda")
    theta_fin = np.full(d, 1.0)

    split = {"fit_ms": 0.0, "predict_ms": 0.0}

    def device_step():
        ctx.timer_start()
        status, rlf = ctx.reduced_likelihood_batch(thetas)
        st, _ = ctx.finalize(theta_fin, want_ft=False)
        assert st == 0
        split["fit_ms"] += ctx.timer_stop()
        ctx.timer_start()
        ctx.predict_valvar_dev(xs_dev.data_ptr(), m, y_dev.data_ptr(), v_dev.data_ptr())
        split["predict_ms"] += ctx.timer_stop()
        return status, rlf

    for _ in range(args.warmup):
        device_step()
    ctx.set_profiling(False)     # production path: evaluations replayed as CUDA graphs, no per-launch events
    ctx.reset_profile()          # launch counters are kept either way
    split["fit_ms"] = split["predict_ms"] = 0.0
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        status, rlf = device_step()
    ev1.record()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    # The kernels run on the context's own stream, which torch events do not see: the step time is
    # the CUDA-event time measured on THAT stream (egx_gp_timer_*: fit + predict regions, GPU idle
    # gaps while the host prepares the next launch included).  Host wall time is kept as a check.
    dev_ms = split["fit_ms"] + split["predict_ms"]
    wall_ms = (t_wall1 - t_wall0) * 1e3
    step_ms = dev_ms / args.steps
    prof = ctx.profile()
    n_fail = int(np.sum(status != 0))
    launches = sum(v[1] for v in prof.values())

    # ---------------- roofline pass for the dominant kernel (K4, gemm_nt_sub_kernel) -----------------
    # The timed region keeps 4 evaluations in flight on separate streams (graph replays), where a CUDA-event
    # bracket around one launch would also contain the time it queued behind other streams.  The per-kernel
    # figure is therefore taken right after it, on the same context and data, with the launches back to back on
    # ONE stream (look-ahead off, no batch concurrency, per-launch events on): 3 likelihood evaluations +
    # finalize + predict_var on one 8192-point chunk.
    ctx.set_lookahead(False)
    ctx.set_profiling(True)
    ctx.reset_profile()
    roof_evals, roof_pts = 3, min(m, 8192)
    for _ in range(roof_evals):
        ctx.reduced_likelihood(theta_fin)
    ctx.finalize(theta_fin, want_ft=False)
    ctx.predict_valvar_dev(xs_dev.data_ptr(), roof_pts, y_dev.data_ptr(), v_dev.data_ptr())
    roof_prof = ctx.profile()
    ctx.set_lookahead(True)
    ctx.set_profiling(False)

    # ---------------- e2e leg: public API, host buffers ------------------------------------
    xs_pinned = torch.from_numpy(xs).pin_memory().numpy()
    x_h, y_h = np.ascontiguousarray(x), np.ascontiguousarray(y)

    def e2e_step():
        # same evaluation budget as the value leg: (n_start + 1) chains x clamp(10 d, 25, 1000) + 1 final
        per_chain = min(max(10 * d, 25), 1000)
        gp = (eg.GaussianProcess.params(eg.ConstantMean, eg.Matern52Corr)
              .n_start(max((E - 1) // per_chain - 1, 0)).max_eval(1000)
              .cobyla_ftol_rel(0.0).device(local_rank).fit(x_h, y_h))
        var = gp.predict_var(xs_pinned)
        nev = gp.n_evals()
        lik, th = gp.likelihood(), gp.theta()
        gp.close()
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

    # ---------------- reductions over ranks -------------------------------------------------
    step_ms_all, e2e_ms_all = step_ms, e2e_ms
    if world > 1:
        t = torch.tensor([step_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms_all, e2e_ms_all = float(t[0]), float(t[1])
        # the one collective of the design: gather (likelihood, theta) of every expert
        pack = torch.tensor([lik] + list(th), dtype=torch.float64, device="cuda")
        gathered = [torch.empty_like(pack) for _ in range(world)]
        dist.all_gather(gathered, pack)

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
            roofline = {"bound": "tensor", "kernel": "ozaki_syrk_kernel (tcgen05.mma kind::i8 on 7 balanced base-256 "
                                                     "digit slices of the fp64 operands; fp64 result): trailing updates "
                                                     "of the factorisations + multi-RHS solve updates of predict_var",
                        "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s",
                        "frac": achieved / bf16_peak, "peak_source": peak_src,
                        "traffic": ncu.get("ozaki_syrk_kernel", {}).get("dram_bytes_per_launch"),
                        "launches": oz_launches, "avg_launch_ms": oz_ms / max(oz_launches, 1),
                        "measured_in": measured_in,
                        "share_of_kernel_time_in_roofline_pass": oz_ms / max(roof_total_ms, 1e-9),
                        "executed_int8_tops": int8_tops,
                        "int8_peak_tops": 2.0 * bf16_peak,
                        "frac_of_int8_peak": int8_tops / (2.0 * bf16_peak),
                        "fp64_pipe_peak_tflops_at_sampled_clock": fp64_peak_at_clock,
                        "speedup_over_fp64_pipe_peak": achieved / fp64_peak_at_clock,
                        "note": "`achieved` counts the useful fp64 flops of the update (2 K per entry of the lower "
                                "triangle), `peak` is the mandated bf16 figure; each fp64 multiply-add costs 28 int8 "
                                "multiply-adds on the tensor core (executed_int8_tops), whose peak is twice the bf16 "
                                "rate (int8_peak_tops = 2 x the measured bf16 figure); for scale: the fp64 (DMMA / "
                                "FMA) pipe of the chip peaks at fp64_pipe_peak_tflops_at_sampled_clock",
                        "dmma_kernel": dmma}
        else:
            roofline = dict(dmma, bound="tensor", peak=bf16_peak, peak_source=peak_src,
                            frac=(dmma_achieved / bf16_peak) if dmma_achieved else None, measured_in=measured_in,
                            note="EGX_OZAKI=0: fp64 contraction on the DMMA pipe; the bf16 figure is the mandated "
                                 "denominator, the fp64-pipe line (148 SM x 64 FMA/clk) is the physical bound")
        value = world * m / (step_ms_all * 1e-3)
        e2e_value = world * m / (e2e_ms_all * 1e-3)
        out = {"metric": "GP fit+predict throughput (points/s) at n=%d d=%d" % (n, d),
               "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": step_ms_all, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": workload_config(args),
               "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": e2e_ms_all, "steps": e2e_steps,
                       "likelihood_evals": nev,
                       "h2d_bytes_per_step": int(x_h.nbytes + y_h.nbytes + xs_pinned.nbytes + nev * d * 8),
                       "d2h_bytes_per_step": int(var.nbytes + nev * 64)},
               "gpu_launches": int(launches),
               "launches_per_step": {k: int(v[1] // args.steps) for k, v in prof.items()},
               "stage_ms_roofline_pass": {k: round(v[0], 3) for k, v in roof_prof.items()},
               "fit_ms_per_step": split["fit_ms"] / args.steps, "predict_ms_per_step": split["predict_ms"] / args.steps,
               "likelihood_evals_per_s": world * E / (split["fit_ms"] / args.steps * 1e-3),
               "predict_var_points_per_s": world * m / (split["predict_ms"] / args.steps * 1e-3),
               "host_wall_ms_per_step": wall_ms / args.steps,
               "failed_theta_in_sweep": n_fail,
               "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_sample_clean_env(n, d, m, E)
        print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-only", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_sample_only:
        print(json.dumps(cpu_sample(args.n, args.d, args.m, args.evals)), flush=True)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
