"""NCCL, world_size 2, one rank per GPU: the evaluation of BASELINE cfg 3 (evaluate_model.py:134-187) sharded by volume gives
the same reduced metrics as one process over all slices -- SegMetrics.reduce() over NCCL (the gloo twin is
tests/test_dist_cpu.py) -- and the data-parallel seg-head training step of cfg 4 with its gradient all-reduce over NCCL.
Needs 2 GPUs (`gpurun --gpus 2`); skipped on a 1-GPU box."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_VOL, PER, BATCH = 5, 12, 16          # odd volume count -> ranks get 3 and 2 volumes; 36 = 16 + 16 + 4: ragged last batch

WORKER = r"""
import json, os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
import yolo_u_b200 as ysp
from yolo_u_b200.synth import synth_state_dicts
n_vol, per, batch = (int(v) for v in sys.argv[2:5])
world, rank, local = (int(os.environ[k]) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
det_sd, seg_sd = synth_state_dicts(0)
P = ysp.Predictor(det_sd, seg_sd, device=dev, mode="tc32")
lo, hi = ysp.shard_slices(n_vol, per, world, rank)
m = ysp.SegMetrics()
g = torch.Generator()
for a, b in ysp.batches(lo, hi, batch):
    imgs, tgs = [], []
    for s in range(a, b):                                   # data depends on the GLOBAL slice index only
        g.manual_seed(500 + s)
        imgs.append(torch.randint(0, 256, (240, 240, 4), dtype=torch.uint8, generator=g))
        tgs.append((torch.rand(240, 240, generator=g) > 0.5).to(torch.uint8) * 255)
    o = P.predict_raw(torch.stack(imgs).to(dev), torch.stack(tgs).to(dev))
    m.update(o["counts"])
res = m.reduce(device=dev).compute()
assert dist.get_backend() == "nccl"
if rank == 0:
    print("RESULT " + json.dumps(res))
dist.destroy_process_group()
"""


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(600)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sharded_evaluation_nccl_world2(tmp_path):
    import yolo_u_b200 as ysp
    from yolo_u_b200.synth import synth_state_dicts
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", str(_free_port()), str(w), ROOT, str(N_VOL), str(PER), str(BATCH)],
                         capture_output=True, text=True, env=env, timeout=540)
    assert out.returncode == 0, out.stderr[-3000:]
    got = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    # single process over all slices, same data
    det_sd, seg_sd = synth_state_dicts(0)
    P = ysp.Predictor(det_sd, seg_sd, device="cuda:0", mode="tc32")
    m = ysp.SegMetrics()
    g = torch.Generator()
    for a, b in ysp.batches(0, N_VOL * PER, BATCH):
        imgs, tgs = [], []
        for s in range(a, b):
            g.manual_seed(500 + s)
            imgs.append(torch.randint(0, 256, (240, 240, 4), dtype=torch.uint8, generator=g))
            tgs.append((torch.rand(240, 240, generator=g) > 0.5).to(torch.uint8) * 255)
        m.update(P.predict_raw(torch.stack(imgs).cuda(), torch.stack(tgs).cuda())["counts"])
    want = m.compute()
    assert got["slices"] == want["slices"] == N_VOL * PER
    assert (got["TP"], got["FP"], got["FN"]) == (want["TP"], want["FP"], want["FN"])        # integer counters: bit-exact
    assert got["dice"] == pytest.approx(want["dice"], abs=1e-12)


TRAIN_WORKER = r"""
import json, os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from oracle.model import build_models, synth_inputs
from yolo_u_b200.trainer import SegHeadTrainer
world, rank, local = (int(os.environ[k]) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
_, seg = build_models(0)
B, S = 2, 64
x, lg, tg = (t.to(dev) for t in synth_inputs(B, S, seed=10 + rank))        # this rank's shard
tr = SegHeadTrainer(seg.state_dict(), batch_size=B, image_size=S, lr=1e-3, epochs=10, device=dev)
for _ in range(2):
    tr.step(x, tg, lg)                                                       # ONE NCCL all-reduce of the flat gradient buffer per step
gathered = [torch.zeros_like(tr.params) for _ in range(world)]
dist.all_gather(gathered, tr.params)
same = max((g - tr.params).abs().max().item() for g in gathered)
stats = [torch.zeros_like(tr.stats) for _ in range(world)]
dist.all_gather(stats, tr.stats)
assert dist.get_backend() == "nccl"
if rank == 0:
    torch.save(tr.params.cpu(), sys.argv[2])
    print("RESULT " + json.dumps({"same": same, "stat_diff": (stats[0] - stats[1]).abs().max().item()}))
dist.destroy_process_group()
"""


@pytest.mark.timeout(600)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_data_parallel_training_nccl_world2(tmp_path):
    """BASELINE cfg 4 over NCCL (train.py:302-331 per rank + one gradient all-reduce): two ranks on two GPUs end with
    bit-identical parameters, equal to a single-process emulation with averaged gradients; BN running statistics stay
    per-rank.  The gloo twin (both ranks on one GPU) is tests/test_gpu_train_ddp.py."""
    from oracle.model import build_models, synth_inputs
    from yolo_u_b200._lib import check, lib
    from yolo_u_b200.trainer import SegHeadTrainer
    w = tmp_path / "train_worker.py"
    w.write_text(TRAIN_WORKER)
    out_pt = tmp_path / "params.pt"
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", str(_free_port()), str(w), ROOT, str(out_pt)],
                         capture_output=True, text=True, env=env, timeout=540)
    assert out.returncode == 0, out.stderr[-3000:]
    got = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert got["same"] == 0.0 and got["stat_diff"] > 0.0
    params = torch.load(out_pt)
    # single-process emulation: both shards' gradients, averaged, same optimiser
    _, seg = build_models(0)
    sd = seg.state_dict()
    B, S = 2, 64
    shards = [synth_inputs(B, S, seed=10 + r) for r in range(2)]
    mk = lambda: SegHeadTrainer(sd, batch_size=B, image_size=S, lr=1e-3, epochs=10, device="cuda:0")
    emu = [mk() for _ in range(2)]
    for _ in range(2):
        for r, e in enumerate(emu):
            ex, elg, etg = (t.cuda() for t in shards[r])
            e.forward_backward(ex, etg, elg)
        avg = sum(e.grads for e in emu) / 2
        for e in emu:
            e.grads.copy_(avg)
            e.step_count += 1
            check(lib().ysp_adamw(e.params.data_ptr(), e.grads.data_ptr(), e.adam_m.data_ptr(), e.adam_v.data_ptr(),
                                  e.params.numel(), e.lr, 0.9, 0.999, e.eps, e.weight_decay, e.step_count, 1.0, 0.0,
                                  e._scratch.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    start = mk().params
    diff = ((params.cuda() - emu[0].params).norm() / (emu[0].params - start).norm()).item()
    assert diff <= 2e-2, diff
