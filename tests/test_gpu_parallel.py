"""The multi-GPU fan-out (egobox_b200/parallel.py) with REAL device contexts: two ranks, one process each.
On a one-GPU box both ranks share cuda:0 and exchange over gloo (NCCL refuses two ranks on one device); with two or
more GPUs the same test also runs over NCCL, rank r on cuda:r."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from tests.gpu_util import make_problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    n, d = 700, 3
    x, y = make_problem(n, d, seed=5)
    thetas = 10.0 ** np.random.default_rng(2).uniform(-1.0, 0.7, size=(13, d))
    xs = np.random.default_rng(3).random((301, d))
    return x, y, thetas, xs


def _fit_params(eg, device):
    return (eg.GaussianProcess.params(eg.ConstantMean, eg.Matern52Corr).n_start(4).max_eval(30).device(device))


def _worker(rank, world, port, backend, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    device = rank if backend == "nccl" else 0
    torch.cuda.set_device(device)
    dist.init_process_group(backend, rank=rank, world_size=world)
    import egobox_b200 as eg
    from egobox_b200 import parallel as P
    from tools._util import make_context
    x, y, thetas, xs = _inputs()
    ctx = make_context(x, y, eg.MATERN52, eg.CONSTANT) if device == 0 else None
    if ctx is None:
        from tools._util import normalize
        xn, xm, xsd = normalize(x)
        yn, ym, ysd = normalize(y.reshape(-1, 1))
        ctx = eg.GpContext(xn, yn[:, 0], xm, xsd, float(ym[0]), float(ysd[0]), eg.MATERN52, eg.CONSTANT, device=device)
    ncalls = []

    def evaluate(th):
        ncalls.append(len(th))
        return ctx.reduced_likelihood_batch(th)
    status, rlf, best = P.theta_sweep(evaluate, thetas)
    ctx.close()
    gp = P.fit_multistart(_fit_params(eg, device), x, y)
    var = P.predict_sharded(gp.predict_var, xs)
    q.put((rank, status.tolist(), rlf.tolist(), best, sum(ncalls), gp.theta().tolist(), gp.likelihood(), gp.n_evals(),
           var.tolist()))
    gp.close()
    dist.destroy_process_group()


def _run(backend):
    import egobox_b200 as eg
    from egobox_b200 import parallel as P
    from tools._util import make_context
    x, y, thetas, xs = _inputs()
    ctx = make_context(x, y, eg.MATERN52, eg.CONSTANT)
    st1, rlf1 = ctx.reduced_likelihood_batch(thetas)
    ctx.close()
    gp1 = _fit_params(eg, 0).fit(x, y)
    th1, lik1, nev1, var1 = gp1.theta(), gp1.likelihood(), gp1.n_evals(), gp1.predict_var(xs)
    gp1.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, backend, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=280) for _ in range(2))
    for p in procs:
        p.join(60)
    evals = 0
    for rank, status, rlf, best, ncalls, th, lik, nev, var in res:
        assert status == st1.tolist()
        np.testing.assert_allclose(np.array(rlf), rlf1, rtol=1e-12)
        assert best == int(np.argmax(np.where(st1 == 0, rlf1, -np.inf)))
        assert ncalls == len(P.shard_indices(len(thetas), rank, 2))       # only its shard was evaluated here
        # every chain sees the sequence it would see alone, so the sharded multistart finds the single-process optimum
        np.testing.assert_allclose(np.array(th), th1, rtol=1e-9)
        assert lik == pytest.approx(lik1, rel=1e-10)
        np.testing.assert_allclose(np.array(var), var1, rtol=1e-9, atol=1e-12 * np.abs(var1).max())
        evals += nev - 1                                                  # the final evaluation is done on every rank
    assert evals == nev1 - 1                                              # the chains were split, not repeated


@pytest.mark.timeout(300)
def test_sharded_sweep_fit_predict_two_ranks_one_gpu_gloo():
    _run("gloo")


@pytest.mark.timeout(300)
def test_sharded_sweep_fit_predict_two_gpus_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL refuses two ranks on one device)")
    _run("nccl")
