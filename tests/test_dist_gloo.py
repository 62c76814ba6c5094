"""World-size-2 gloo test (CPU) of the multi-GPU host logic: contiguous sharding and the SPC all-reduce."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from qcgraft import dist as qd
    r, lr, w = qd.init("gloo")
    assert (r, w) == (rank, world)
    lo, hi = qd.shard_range(2048 + 3, rank, world)
    v = torch.arange(9, dtype=torch.float64) * (rank + 1)
    v[0] = hi - lo
    h = qd.spc_allreduce(v, async_op=True)
    h.wait()
    mx = qd.max_over_ranks(10.0 + rank)
    qd.barrier()
    out[rank] = (lo, hi, v.tolist(), mx)


def test_shard_and_spc_allreduce_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (lo0, hi0, v0, m0), (lo1, hi1, v1, m1) = out[0], out[1]
    assert lo0 == 0 and hi0 == lo1 and hi1 == 2051 and (hi0 - lo0) - (hi1 - lo1) in (0, 1)
    expect = (np.arange(9) * 1.0 + np.arange(9) * 2.0)
    expect[0] = 2051
    assert v0 == v1 == expect.tolist()
    assert m0 == m1 == 11.0


def test_shard_range_covers_everything():
    from qcgraft import dist as qd
    for n in (0, 1, 7, 64, 2048):
        for world in (1, 2, 3, 8):
            spans = [qd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_spc_summary():
    from qcgraft import dist as qd
    s = qd.spc_summary(np.array([4, 8, 20, 6.4, 8, 2, 2, 4, 1000.0]))
    assert s["c_bar"] == 2.0 and abs(s["avg_confidence"] - 0.8) < 1e-12 and s["critical_rate"] == 0.25
    assert abs(s["ucl"] - (2 + 3 * np.sqrt(2))) < 1e-12 and s["lcl"] == 0.0
