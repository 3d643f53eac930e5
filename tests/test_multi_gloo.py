"""world_size-2 test of the N>1 host logic on CPU (gloo): ranks are independent pipelines, the only
exchange is barrier + max/sum reductions of scalars (no data-path collective exists)."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import numpy as np
    from gstreamer_b200 import multi
    from oracle import bindings as ob
    dist = multi.init(backend="gloo")
    assert dist is not None and dist.get_world_size() == world
    # every rank converts its OWN stream (distinct seeds) with the CPU checker standing in for the
    # per-GPU pipeline; frames must differ between ranks and results stay rank-local
    d = ob.vcs_desc(64, 48, 32, 24, 3)
    frame = ob.nv12_random_frame(64, 48, multi.stream_seed(rank, 0))
    out = ob.oracle_vcs_convert(d, frame)
    local_ms = 10.0 + 5.0 * rank
    dist.barrier()
    (mx,) = multi.reduce_max(dist, [local_ms])
    (total,) = multi.reduce_sum(dist, [float(out.sum())])
    value = multi.whole_job_throughput(32, world, mx * 1e-3)
    q.put((rank, mx, total, value, int(out.sum()), int(frame[:64].sum())))
    dist.destroy_process_group()


def test_two_rank_stream_parallel():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mx0, tot0, v0, s0, f0), (r1, mx1, tot1, v1, s1, f1) = res
    assert mx0 == mx1 == 15.0                       # slowest rank defines the job time
    assert tot0 == tot1 == s0 + s1                  # sum over ranks
    assert f0 != f1                                 # distinct streams per rank
    assert v0 == v1 == pytest.approx(32 * 2 / 0.015)


def test_single_rank_has_no_process_group(monkeypatch):
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    from gstreamer_b200 import multi
    assert multi.init() is None
    assert multi.reduce_max(None, [3.0, 1.0]) == [3.0, 1.0]
    assert multi.rank_info() == (0, 1, 0)
