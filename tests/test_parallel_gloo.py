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
    # the reduce by min of the sharded multistart and the point-sharded prediction
    f_loc = [3.5, 1.25][rank] if world == 2 else 3.5
    fwin, zwin = P.argmin_exchange(f_loc, np.array([10.0 * rank + 1.0, 10.0 * rank + 2.0]))
    finf, zinf = P.argmin_exchange(float("inf") if rank == 0 else float("nan"), np.array([float(rank)]))
    xs = np.arange(22.0).reshape(11, 2)
    pred = P.predict_sharded(lambda a: a[:, 0] * 2.0 + a[:, 1], xs)
    q.put((rank, status.tolist(), rlf.tolist(), best, sum(calls), sorted(models), table.tolist(),
           (fwin, zwin.tolist(), finf, zinf.tolist(), pred.tolist())))
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
    for rank, status, rlf, best, ncalls, models, table, extra in res:
        fwin, zwin, finf, zinf, pred = extra
        assert fwin == 1.25 and zwin == [11.0, 12.0]             # rank 1 holds the smaller objective
        assert finf == float("inf") and zinf == [0.0]            # nobody succeeded: +inf, lowest rank
        xs = np.arange(22.0).reshape(11, 2)
        np.testing.assert_array_equal(np.array(pred), xs[:, 0] * 2.0 + xs[:, 1])
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


# ---------------------------------------------------------------- mixture of experts -----------
class _OracleExpert:
    """Stand-in for a GPU expert on the CPU ranks: the ORACLE's fixed-theta GP (checker code, test only)."""

    def __init__(self, x, y):
        from oracle import gp_oracle as O
        self.gp = O.fit(x, y, corr=O.SQEXP, mean=O.CONSTANT, theta_init=np.full(x.shape[1], 2.0), fixed=True)

    def predict_valvar(self, x):
        return self.gp.predict_valvar(x)

    def likelihood(self):
        return float(self.gp.likelihood)

    def variance(self):
        return float(self.gp.inner.sigma2)

    def theta(self):
        return np.asarray(self.gp.theta)

    def close(self):
        pass


def _mix_problem():
    rng = np.random.default_rng(3)
    x = rng.random((90, 2))
    y = np.where(x[:, 0] < 0.5, np.sin(6 * x[:, 0]) + x[:, 1], 3.0 + x[:, 0] * x[:, 1])
    centers = np.array([[0.2, 0.5], [0.55, 0.5], [0.85, 0.5]])

    def probas(xq):
        d2 = ((xq[:, None, :] - centers[None, :, :]) ** 2).sum(axis=2)
        w = np.exp(-d2 / 0.02)
        return w / w.sum(axis=1, keepdims=True)
    labels = np.argmax(probas(x), axis=1)
    xs = rng.random((57, 2))
    return x, y, labels, probas, xs


def _mix_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from egobox_b200 import mixture as M
    x, y, labels, probas, xs = _mix_problem()
    out = {}
    for rec in (M.Recombination.HARD, M.Recombination.SMOOTH):
        mix = M.ExpertMixture.fit(x, y, labels, probas, recombination=rec, expert_fit=_OracleExpert)
        yv, vv = mix.predict_valvar(xs)
        out[rec] = (yv.tolist(), vv.tolist(), sorted(mix.experts), mix.table.tolist())
    q.put((rank, out))
    dist.destroy_process_group()


def test_recombination_formulas():
    """moe/src/algorithm.rs:417-421 (sum_k p_k y_k), :675-683 (sum_k p_k^2 var_k), :880 (argmax cluster)."""
    from egobox_b200 import mixture as M
    rng = np.random.default_rng(0)
    preds, var = rng.normal(size=(3, 11)), rng.random((3, 11))
    p = rng.random((11, 3))
    p /= p.sum(axis=1, keepdims=True)
    y, v = M.recombine_smooth(preds, var, p)
    for i in range(11):
        assert y[i] == pytest.approx(sum(p[i, k] * preds[k, i] for k in range(3)), rel=1e-14)
        assert v[i] == pytest.approx(sum(p[i, k] ** 2 * var[k, i] for k in range(3)), rel=1e-14)
    np.testing.assert_array_equal(M.hard_clusters(p), np.argmax(p, axis=1))


@pytest.mark.timeout(240)
def test_expert_mixture_world2_matches_single_process():
    """Experts sharded over 2 gloo ranks (cluster c on rank c % 2) give the single-process mixture."""
    from egobox_b200 import mixture as M
    x, y, labels, probas, xs = _mix_problem()
    ref = {}
    for rec in (M.Recombination.HARD, M.Recombination.SMOOTH):
        mix = M.ExpertMixture.fit(x, y, labels, probas, recombination=rec, expert_fit=_OracleExpert)
        assert mix.n_clusters == 3
        ref[rec] = mix.predict_valvar(xs)
        # against the reference formulas applied to the three experts directly
        e = [mix.experts[c].predict_valvar(xs) for c in range(3)]
        p = probas(xs)
        if rec == M.Recombination.SMOOTH:
            yy, vv = M.recombine_smooth(np.array([a[0] for a in e]), np.array([a[1] for a in e]), p)
        else:
            cl = np.argmax(p, axis=1)
            yy = np.array([e[cl[i]][0][i] for i in range(len(xs))])
            vv = np.array([e[cl[i]][1][i] for i in range(len(xs))])
        np.testing.assert_allclose(ref[rec][0], yy, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(ref[rec][1], vv, rtol=1e-12, atol=1e-12)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mix_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = sorted((q.get(timeout=200) for _ in range(2)), key=lambda t: t[0])
    for p_ in procs:
        p_.join(30)
    for rank, out in res:
        for rec in (M.Recombination.HARD, M.Recombination.SMOOTH):
            yv, vv, mine, table = out[rec]
            assert mine == [c for c in range(3) if c % 2 == rank]
            np.testing.assert_allclose(np.array(yv), ref[rec][0], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(np.array(vv), ref[rec][1], rtol=1e-12, atol=1e-12)
            assert np.array(table).shape == (3, 4)
