"""CPU, world size 3: the C-ABI exchange of the sharded path (egx_comm_init / egx_comm_allgather / egx_argmin_allreduce,
csrc/host_comm.cpp) between real processes on 127.0.0.1 -- SURVEY section 8 (b) / (e)."""
import math
import multiprocessing as mp
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    try:
        from egobox_b200.parallel import HostComm
        comm = HostComm(world, rank, "127.0.0.1", port, timeout_ms=20000)
        gathered = comm.allgather([rank + 0.5, 10.0 * rank])
        # rank r proposes value (r - 1)^2 (rank 1 wins) with payload [r, r, r]; a second round with a NaN and a tie
        v, p, w = comm.argmin((rank - 1.0) ** 2, np.full(3, float(rank)))
        v2, p2, w2 = comm.argmin(math.nan if rank == 0 else 7.0, [float(rank)])
        comm.close()
        q.put((rank, gathered.tolist(), v, p.tolist(), w, v2, p2.tolist(), w2))
    except Exception as e:          # pragma: no cover
        q.put((rank, "error", repr(e)))


def test_exchange_between_three_processes():
    world, port = 3, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in reversed(procs):       # rank 0 starts LAST: the others must retry their connect until it listens
        p.start()
    results = sorted(q.get(timeout=60) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, gathered, v, payload, w, v2, p2, w2 in results:
        assert gathered == [[0.5, 0.0], [1.5, 10.0], [2.5, 20.0]]
        assert (v, payload, w) == (0.0, [1.0, 1.0, 1.0], 1)
        assert (v2, p2, w2) == (7.0, [1.0], 1)                  # NaN never wins; of the tied ranks 1 and 2 the first


def test_single_rank_needs_no_address_and_bad_arguments_are_refused():
    from egobox_b200.parallel import HostComm
    from egobox_b200._lib import GpuError
    comm = HostComm(1, 0, addr=None, port=0)
    assert comm.allgather([1.0, 2.0]).tolist() == [[1.0, 2.0]]
    assert comm.argmin(3.0, [4.0])[::2] == (3.0, 0)
    comm.close()
    with pytest.raises(GpuError):
        HostComm(2, 2, "127.0.0.1", 29999)                      # rank out of range
    with pytest.raises(GpuError):
        HostComm(2, 1, "127.0.0.1", _free_port(), timeout_ms=300)      # nobody listens: times out, does not hang
