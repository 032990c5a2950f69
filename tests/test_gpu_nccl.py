"""NCCL, world_size 2, one rank per GPU: the evaluation of BASELINE cfg 3 (evaluate_model.py:134-187) sharded by volume gives
the same reduced metrics as one process over all slices -- SegMetrics.reduce() over NCCL (the gloo twin is
tests/test_dist_cpu.py).  Needs 2 GPUs (`gpurun --gpus 2`); skipped on a 1-GPU box."""
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
