#!/usr/bin/env python
"""Sampling entry point: the role of /root/reference/sample_uncond.py.
    torchrun --nproc-per-node N scripts/sample_uncond.py --cfg configs/cifar10/ddm_uncond_const_uncond_unet.yaml [--out samples.pt]
The sample_num images are sharded over the ranks (no communication); each rank writes its shard."""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adm_b200.trainer import Sampler  # noqa: E402
from scripts.train_uncond_dpm import build_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", required=True)
    ap.add_argument("--out", default=None)
    ap.add_argument("--num", type=int, default=None)
    args = ap.parse_args()
    cfg = yaml.safe_load(open(args.cfg))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=device)
    rank = dist.get_rank() if dist.is_initialized() else 0
    torch.manual_seed(99 + rank)
    s = cfg["sampler"]
    model = build_model(cfg, device).eval()
    ckpt = s.get("ckpt_path")
    sampler = Sampler(model, batch_size=s["batch_size"], sample_num=args.num or s["sample_num"],
                      ckpt_path=ckpt if ckpt and os.path.isfile(ckpt) else None, use_ema=s.get("use_ema", True), cfg=cfg)
    imgs = sampler.sample()
    if args.out:
        torch.save(imgs.float().cpu(), f"{args.out}.rank{rank}" if dist.is_initialized() else args.out)
    print(f"rank {rank}: sampled {tuple(imgs.shape)} in [{float(imgs.min()):.3f}, {float(imgs.max()):.3f}]", flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
