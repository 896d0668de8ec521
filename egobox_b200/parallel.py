"""Multi-GPU fan-out of the kriging path: one process per GPU, torch.distributed for plumbing.

The path shards as independent units:
  * theta candidates of a likelihood sweep / multistart chains  -> `theta_sweep`
  * per-cluster experts of egobox-moe (moe/src/algorithm.rs:167-177) -> `fit_experts`
There is no data-path collective: every rank works on a replica (or its own expert) and the
only exchange is one `all_gather` of a few doubles per rank -- the replacement of the rayon
`reduce` by min at gp/src/algorithm.rs:942-945.  Works with the `nccl` backend on GPUs and
with `gloo` on CPU (the tests run world_size 2 over gloo with a stand-in evaluator)."""
from __future__ import annotations

import numpy as np


def shard_indices(n_items, rank, world):
    """Round-robin shard: item i belongs to rank i % world (balanced to within one item)."""
    return list(range(rank, n_items, world))


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def _device_for_backend():
    import torch
    dist = _dist()
    if dist is not None and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_rows(local_rows, width, counts):
    """Gather a ragged set of float64 rows (each `width` wide) from every rank.

    counts[r] = number of rows rank r contributes (known to everybody from the sharding rule).
    Returns a list (per rank) of (counts[r], width) arrays."""
    import torch
    dist = _dist()
    local = np.asarray(local_rows, dtype=np.float64).reshape(-1, width)
    if dist is None:
        return [local]
    world = dist.get_world_size()
    cap = max(counts) if counts else 0
    dev = _device_for_backend()
    buf = torch.zeros((cap, width), dtype=torch.float64, device=dev)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(local).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return [o[: counts[r]].cpu().numpy() for r, o in enumerate(out)]


def theta_sweep(evaluate_batch, thetas):
    """Evaluate B candidate thetas sharded over the ranks and return the global result.

    evaluate_batch(thetas_local) -> (status[int], rlf[float]) for the local shard (on a GPU rank this
    is GpContext.reduced_likelihood_batch of the rank's replica of the training set).
    Returns (status[B], rlf[B], best_index) on every rank; failed candidates count as -inf
    likelihood (objective +inf, gp/src/algorithm.rs:893-896)."""
    thetas = np.asarray(thetas, dtype=np.float64)
    B = thetas.shape[0]
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    mine = shard_indices(B, rank, world)
    if mine:
        st, rl = evaluate_batch(thetas[mine])
        rows = np.stack([np.asarray(st, dtype=np.float64), np.asarray(rl, dtype=np.float64)], axis=1)
    else:
        rows = np.zeros((0, 2))
    counts = [len(shard_indices(B, r, world)) for r in range(world)]
    parts = all_gather_rows(rows, 2, counts)
    status = np.zeros(B, dtype=np.int32)
    rlf = np.full(B, np.nan)
    for r, part in enumerate(parts):
        idx = shard_indices(B, r, world)
        status[idx] = part[:, 0].astype(np.int32)
        rlf[idx] = part[:, 1]
    score = np.where((status == 0) & ~np.isnan(rlf), rlf, -np.inf)
    best = int(np.argmax(score)) if B else -1
    return status, rlf, best


def fit_experts(fit_one, n_experts, theta_dim):
    """MoE fan-out: expert e is fitted by rank e % world; (likelihood, variance, theta[theta_dim])
    of every expert is then known on every rank.

    fit_one(e) -> (model, likelihood, variance, theta) ; the model (device state) stays on its rank.
    Returns (local_models: dict e -> model, table: (n_experts, 2 + theta_dim) array)."""
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    mine = shard_indices(n_experts, rank, world)
    models, rows = {}, []
    for e in mine:
        model, lik, var, theta = fit_one(e)
        models[e] = model
        rows.append(np.concatenate([[lik, var], np.asarray(theta, dtype=np.float64).reshape(-1)]))
    width = 2 + theta_dim
    counts = [len(shard_indices(n_experts, r, world)) for r in range(world)]
    parts = all_gather_rows(np.array(rows).reshape(-1, width), width, counts)
    table = np.zeros((n_experts, width))
    for r, part in enumerate(parts):
        table[shard_indices(n_experts, r, world)] = part
    return models, table


def argmin_exchange(f_best, z_best):
    """The `reduce` by min of gp/src/algorithm.rs:942-945 completed across ranks: every rank contributes its best
    (objective, log10 theta); all get those of the rank with the smallest objective (ties: lowest rank; +inf / NaN lose)."""
    z = np.asarray(z_best, dtype=np.float64).reshape(-1)
    dist = _dist()
    if dist is None:
        return float(f_best), z
    world = dist.get_world_size()
    f = float(f_best)
    row = np.concatenate([[f if np.isfinite(f) else np.inf], z])[None, :]
    parts = all_gather_rows(row, z.size + 1, [1] * world)
    table = np.concatenate(parts, axis=0)
    win = int(np.argmin(table[:, 0]))
    return float(table[win, 0]), table[win, 1:].copy()


def fit_multistart(params, x, y):
    """`Fit::fit` (gp/src/algorithm.rs:791-979) with the n_start + 1 optimiser chains sharded over the ranks: chain c runs on
    rank c % world against that rank's replica of the training set, one all-gather of (objective, theta) -- 8 (h + 1) bytes
    per rank -- replaces the rayon `reduce` (:942-945), and EVERY rank finalises at the winning theta, so each holds the
    complete trained model (needed for point-sharded prediction)."""
    dist = _dist()
    if dist is None:
        return params.fit(x, y)
    return params.chain_shard(dist.get_rank(), dist.get_world_size(), argmin_exchange).fit(x, y)


def predict_sharded(predict_fn, x, width=1):
    """Point-sharded prediction: rank r evaluates `predict_fn` (e.g. gp.predict_var) on the contiguous slice r of the rows
    of x; one all-gather returns the full result on every rank."""
    x = np.asarray(x, dtype=np.float64)
    dist = _dist()
    if dist is None:
        return np.asarray(predict_fn(x))
    rank, world = dist.get_rank(), dist.get_world_size()
    m = x.shape[0]
    bounds = [(m * r) // world for r in range(world + 1)]
    lo, hi = bounds[rank], bounds[rank + 1]
    local = np.asarray(predict_fn(x[lo:hi]), dtype=np.float64).reshape(hi - lo, width) if hi > lo else np.zeros((0, width))
    parts = all_gather_rows(local, width, [bounds[r + 1] - bounds[r] for r in range(world)])
    out = np.concatenate(parts, axis=0)
    return out[:, 0] if width == 1 else out


class HostComm:
    """The exchange of the sharded path without torch: `egx_comm_*` of the C ABI (csrc/host_comm.cpp), a TCP star on
    addr:port (rank 0 listens).  What a Rust / C caller of the library uses; the functions above do the same exchange over a
    torch process group (NCCL / gloo)."""

    def __init__(self, nranks, rank, addr="127.0.0.1", port=29600, timeout_ms=60000):
        import ctypes as C
        from . import _lib
        self._C, self._lib = C, _lib.load()
        self._h = C.c_void_p()
        st = self._lib.egx_comm_init(C.byref(self._h), int(nranks), int(rank), str(addr).encode(), int(port), int(timeout_ms))
        if st != 0:
            raise _lib.GpuError(st, _lib.last_error())
        self.rank, self.size = int(rank), int(nranks)

    def allgather(self, values):
        """(count,) doubles of this rank -> (size, count) array of every rank, on every rank."""
        C = self._C
        v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        out = np.empty((self.size, v.size))
        dp = C.POINTER(C.c_double)
        st = self._lib.egx_comm_allgather(self._h, v.ctypes.data_as(dp), v.size, out.ctypes.data_as(dp))
        if st != 0:
            from . import _lib
            raise _lib.GpuError(st, _lib.last_error())
        return out

    def argmin(self, value, payload):
        """-> (value, payload, winner rank) of the rank with the smallest value (gp/src/algorithm.rs:942-945 across ranks)."""
        C = self._C
        v = C.c_double(float(value))
        p = np.ascontiguousarray(payload, dtype=np.float64).reshape(-1).copy()
        w = C.c_int(-1)
        dp = C.POINTER(C.c_double)
        st = self._lib.egx_argmin_allreduce(self._h, C.byref(v), p.ctypes.data_as(dp), p.size, C.byref(w))
        if st != 0:
            from . import _lib
            raise _lib.GpuError(st, _lib.last_error())
        return v.value, p, w.value

    def close(self):
        if self._h:
            self._lib.egx_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
