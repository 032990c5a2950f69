"""CPU, world_size 2, gloo: the only collective of the inference path -- the counter all-reduce -- and
the contiguous-by-volume sharding give the same metrics as a single process."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import yolo_u_b200 as ysp
    g = torch.Generator().manual_seed(7)
    n_vol, per = 6, 5
    counts = torch.randint(0, 100, (n_vol * per, 3), generator=g)
    counts[:, 0] = torch.minimum(counts[:, 0], torch.minimum(counts[:, 1], counts[:, 2]))
    lo, hi = ysp.shard_slices(n_vol, per, world, rank)
    m = ysp.SegMetrics()
    for a, b in ysp.batches(lo, hi, 4):
        m.update(counts[a:b].int())
    m.reduce()
    if rank == 0:
        full = ysp.SegMetrics()
        full.update(counts.int())
        q.put((m.compute(), full.compute()))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_counter_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, want = q.get(timeout=100)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert got["slices"] == want["slices"] == 30
    assert (got["TP"], got["FP"], got["FN"]) == (want["TP"], want["FP"], want["FN"])
    assert got["dice"] == pytest.approx(want["dice"], abs=1e-9)
