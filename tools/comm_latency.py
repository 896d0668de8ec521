"""Latency of the C-ABI exchange (egx_comm_allgather / egx_argmin_allreduce, csrc/host_comm.cpp: a TCP star on loopback) between
`world` real processes: the message of the sharded path is 8 (h + 2) bytes per rank, so this latency -- not a bandwidth -- is what
the choice of transport costs.   python tools/comm_latency.py [world] [h]  ->  one JSON line"""
import json
import multiprocessing as mp
import socket
import sys
import time

import numpy as np


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, h, reps, q):
    sys.path.insert(0, ".")
    from egobox_b200.parallel import HostComm
    comm = HostComm(world, rank, "127.0.0.1", port, timeout_ms=20000)
    payload = np.full(h, float(rank))
    for _ in range(50):
        comm.argmin(float(rank), payload)
    t0 = time.perf_counter()
    for _ in range(reps):
        comm.argmin(float(rank), payload)
    t1 = time.perf_counter()
    for _ in range(reps):
        comm.allgather(payload)
    t2 = time.perf_counter()
    comm.close()
    q.put((rank, (t1 - t0) / reps * 1e6, (t2 - t1) / reps * 1e6))


if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    h = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    reps, port = 2000, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, h, reps, q)) for r in range(world)]
    for p in reversed(procs):
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    print(json.dumps({"world": world, "payload_doubles": h, "reps": reps,
                      "argmin_allreduce_us_max_over_ranks": round(max(r[1] for r in res), 1),
                      "allgather_us_max_over_ranks": round(max(r[2] for r in res), 1),
                      "transport": "TCP star on 127.0.0.1 (csrc/host_comm.cpp), ctypes call overhead included"}))
