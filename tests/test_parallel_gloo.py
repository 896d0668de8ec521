"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding of theta candidates /
experts and the single all_gather reduction.  The per-rank evaluator is the ORACLE here
(stand-in for a GPU replica), so the distributed result must equal the single-process one."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    from oracle import gp_oracle as O
    rng = np.random.default_rng(0)
    x = rng.random((40, 2))
    y = np.sin(3 * x[:, 0]) + x[:, 1] ** 2
    xn, _, _ = O.normalize(x)
    yn, _, ys = O.normalize(y.reshape(-1, 1))
    fx = O.mean_value(O.CONSTANT, xn)
    thetas = 10.0 ** np.random.default_rng(1).uniform(-1.5, 1.0, size=(7, 2))
    thetas[3, 0] = np.nan                         # a failing candidate

    def evaluate(th):
        st, rl = [], []
        for t in th:
            v = O.objective(O.SQEXP, xn, fx, yn, float(ys[0]), t, np.eye(2))
            st.append(0 if np.isfinite(v) else 4)
            rl.append(-v if np.isfinite(v) else np.nan)
        return st, rl
    return thetas, evaluate


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from egobox_b200 import parallel as P
    thetas, evaluate = _problem()
    calls = []

    def counted(th):
        calls.append(len(th))
        return evaluate(th)
    status, rlf, best = P.theta_sweep(counted, thetas)

    def fit_one(e):
        return ("model%d" % e, 10.0 * e + 1.0, 0.5 * e, [e, e + 0.5, e + 0.25])
    models, table = P.fit_experts(fit_one, 5, 3)
    q.put((rank, status.tolist(), rlf.tolist(), best, sum(calls), sorted(models), table.tolist()))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_theta_sweep_and_experts_world2():
    from egobox_b200 import parallel as P
    thetas, evaluate = _problem()
    st1, rlf1, best1 = P.theta_sweep(evaluate, thetas)      # single process reference
    assert st1[3] != 0 and best1 != 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(2))
    for p in procs:
        p.join(30)
    for rank, status, rlf, best, ncalls, models, table in res:
        assert status == st1.tolist()
        np.testing.assert_allclose(np.array(rlf), rlf1, rtol=0, atol=0, equal_nan=True)
        assert best == best1
        assert ncalls == len(P.shard_indices(7, rank, 2))        # each rank evaluated only its shard
        assert models == P.shard_indices(5, rank, 2)
        t = np.array(table)
        np.testing.assert_allclose(t[:, 0], 10.0 * np.arange(5) + 1.0)
        np.testing.assert_allclose(t[:, 2], np.arange(5))


def test_shard_indices_cover():
    from egobox_b200 import parallel as P
    for n, w in [(0, 2), (1, 4), (7, 2), (512, 8), (11, 8)]:
        allidx = sorted(i for r in range(w) for i in P.shard_indices(n, r, w))
        assert allidx == list(range(n))
        sizes = [len(P.shard_indices(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
