"""The reference's evaluation loop (evaluate_model.py:134-187) on this library, with synthetic volumes.

    python examples/evaluate_synthetic.py [--volumes 4] [--slices 155] [--batch 256] [--mode tc32]
    torchrun --nproc-per-node N examples/evaluate_synthetic.py ...      # contiguous-by-volume shards, ONE metric all-reduce

Per batch: detector -> sigmoid(P3 class logit) bottleneck -> NMS -> YOLO-Seg++ -> sigmoid > 0.5 -> Dice / TP / FP / FN,
all inside `Predictor.predict_raw` (one ysp_pipeline call); the host only accumulates the [B,3] integer counters."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import yolo_u_b200 as ysp
from yolo_u_b200.synth import calibrate, synth_state_dicts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--volumes", type=int, default=4)
    ap.add_argument("--slices", type=int, default=155)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--mode", default="tc32", choices=["tc32", "bf16", "fp32"])
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    det_sd, seg_sd = calibrate(*synth_state_dicts(0), device=dev)          # stand-in for best.pt / best.pth
    P = ysp.Predictor(det_sd, seg_sd, device=dev, mode=args.mode)
    lo, hi = ysp.shard_slices(args.volumes, args.slices, world, rank)       # this rank's slices, whole volumes
    metrics = ysp.SegMetrics()
    g = torch.Generator().manual_seed(0)
    for a, b in ysp.batches(lo, hi, args.batch):
        g.manual_seed(a)                                                    # slice data depends on the global index only
        img = torch.randint(0, 256, (b - a, 240, 240, 4), dtype=torch.uint8, generator=g).to(dev)   # decoded PNGs (BGRA)
        mask = (torch.rand(b - a, 1, 240, 240, generator=g) > 0.5).float().to(dev)
        out = P.predict_raw(img, mask)
        metrics.update(out["counts"])
    res = metrics.reduce(device=dev).compute()
    if rank == 0:
        print({k: (round(v, 6) if isinstance(v, float) else v) for k, v in res.items()})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
