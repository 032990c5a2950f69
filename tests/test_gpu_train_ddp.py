"""Data-parallel seg-head training (BASELINE cfg 4): two ranks, each with its own batch shard, ONE all-reduce
of the flat gradient buffer per step, BN statistics local.  Both ranks share cuda:0 here (gloo carries the CUDA tensors);
the NCCL path is the same `dist.all_reduce(self.grads)` call and is exercised by `bench.py --workload train --gpus N`."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.model import build_models, synth_inputs
        from yolo_u_b200.trainer import SegHeadTrainer
        _, seg = build_models(0)
        sd = seg.state_dict()
        B, S = 2, 64
        shards = [synth_inputs(B, S, seed=10 + r) for r in range(world)]
        mk = lambda: SegHeadTrainer(sd, batch_size=B, image_size=S, lr=1e-3, epochs=10, device="cuda:0")
        # distributed run: 2 steps on this rank's shard
        tr = mk()
        x, lg, tg = (t.cuda() for t in shards[rank])
        for _ in range(2):
            tr.step(x, tg, lg)
        # emulation on one process: both shards' gradients, averaged, same optimiser
        emu = [mk() for _ in range(world)]
        for _ in range(2):
            for r, e in enumerate(emu):
                ex, elg, etg = (t.cuda() for t in shards[r])
                e.forward_backward(ex, etg, elg)
            avg = sum(e.grads for e in emu) / world
            for e in emu:
                e.grads.copy_(avg)
                e.pg = "skip"
                e.step_count += 1
                from yolo_u_b200._lib import check, lib
                check(lib().ysp_adamw(e.params.data_ptr(), e.grads.data_ptr(), e.adam_m.data_ptr(), e.adam_v.data_ptr(),
                                      e.params.numel(), e.lr, 0.9, 0.999, e.eps, e.weight_decay, e.step_count, 1.0, 0.0,
                                      e._scratch.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        start = mk().params
        diff = ((tr.params - emu[rank].params).norm() / (emu[rank].params - start).norm()).item()
        # ranks hold identical parameters after the all-reduced step; BN running statistics stay local (differ)
        gathered = [torch.zeros_like(tr.params) for _ in range(world)]
        dist.all_gather(gathered, tr.params)
        same = max((g - tr.params).abs().max().item() for g in gathered)
        stats = [torch.zeros_like(tr.stats) for _ in range(world)]
        dist.all_gather(stats, tr.stats)
        ret[rank] = (diff, same, (stats[0] - stats[1]).abs().max().item())
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_emulation():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for r in range(2):
        diff, same, stat_diff = ret[r]
        assert same == 0.0                 # bit-identical parameters on both ranks
        # = single-process emulation with averaged gradients.  Relative L2 of the parameter update: atomics reorder
        # the gradient sums and Adam turns rounding-level gradients into +-lr moves (see test_gpu_train.py)
        assert diff <= 2e-2
        assert stat_diff > 0.0             # BN running statistics are per-rank, as without SyncBN
