"""The reference's seg-head training loop body (train.py:302-331, non-AMP branch) on this library, synthetic data.

    python examples/train_head_synthetic.py [--steps 50] [--batch 32] [--size 240] [--loss dice]
    torchrun --nproc-per-node N examples/train_head_synthetic.py ...   # data parallel: ONE gradient all-reduce per step

Frozen detector encoder -> decoder in train() mode (BatchNorm batch statistics) -> monai-style Dice loss -> hand-derived
backward -> AdamW, all in libysp (`SegHeadTrainer.step`).  `state_dict()` is a reference-compatible `best.pth`."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from yolo_u_b200.synth import synth_state_dicts
from yolo_u_b200.trainer import SegHeadTrainer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=240)
    ap.add_argument("--loss", default="dice", choices=["dice", "dice_bce"])
    ap.add_argument("--save", default="")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _, seg_sd = synth_state_dicts(0)                                        # stand-in for the pretrained detector + fresh head
    tr = SegHeadTrainer(seg_sd, batch_size=args.batch, image_size=args.size, lr=1e-3, epochs=10, loss=args.loss, device=dev)
    S, B = args.size, args.batch
    g = torch.Generator().manual_seed(100 + rank)
    img = torch.rand(B, 4, S, S, generator=g).to(dev)
    heat = torch.sigmoid(torch.randn(B, 1, S // 8, S // 8, generator=g)).to(dev)
    mask = torch.zeros(B, 1, S, S)
    mask[:, :, S // 4:3 * S // 4, S // 3:2 * S // 3] = 1.0                  # a blob to segment
    mask = mask.to(dev)
    for it in range(args.steps):
        loss, _ = tr.step(img, mask, heat)
        if rank == 0 and (it % 10 == 0 or it == args.steps - 1):
            print(f"step {it:4d}  loss {loss[0].item():.4f}  lr {tr.lr:.2e}")
        if (it + 1) % 25 == 0:
            tr.scheduler_step()                                             # the reference steps the cosine schedule per epoch
    if rank == 0 and args.save:
        torch.save(tr.state_dict(), args.save)                              # loads with the reference's load_state_dict
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
