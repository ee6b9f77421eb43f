#!/usr/bin/env python
"""Training entry point: the role of /root/reference/train_uncond_dpm.py (`accelerate launch train_uncond_dpm.py --cfg`),
without accelerate:   torchrun --nproc-per-node N scripts/train_uncond_dpm.py --cfg configs/cifar10/ddm_uncond_const_uncond_unet.yaml
Builds the UNet and the DDPM module from the YAML with construct_class_by_name (train_uncond_dpm.py:36-46), then runs
adm_b200.trainer.Trainer."""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adm_b200.ddm.utils import construct_class_by_name  # noqa: E402
from adm_b200.trainer import Trainer  # noqa: E402


def synthetic_loader(batch, image_size, seed):
    g = torch.Generator().manual_seed(seed)
    while True:
        yield {"image": (2 * torch.rand(batch, 3, *image_size, generator=g) - 1).pin_memory()}


def build_model(cfg, device):
    model_cfg = dict(cfg["model"])
    unet_cfg = dict(model_cfg.pop("unet"))
    unet = construct_class_by_name(**unet_cfg)
    model_cfg.pop("class_name_unet", None)
    cls = model_cfg.pop("class_name")
    return construct_class_by_name(class_name=cls, model=unet, cfg=model_cfg, **model_cfg).to(device)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", required=True)
    ap.add_argument("--steps", type=int, default=None, help="override trainer.train_num_steps")
    args = ap.parse_args()
    cfg = yaml.safe_load(open(args.cfg))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=device)
    rank = dist.get_rank() if dist.is_initialized() else 0
    torch.manual_seed(1234)  # identical initial weights on every rank (TrainStep also broadcasts rank 0's arena)
    model = build_model(cfg, device)
    torch.manual_seed(1234 + rank)  # per-rank streams for t, noise, dropout and augmentation from here on
    t, d = cfg["trainer"], cfg["data"]
    if d.get("class_name", "synthetic") == "synthetic":
        dl = synthetic_loader(d["batch_size"], d["image_size"], seed=rank)
    else:
        ds = construct_class_by_name(**{k: v for k, v in d.items() if k not in ("batch_size", "num_workers")})
        dl = torch.utils.data.DataLoader(ds, batch_size=d["batch_size"], shuffle=True, pin_memory=True,
                                         num_workers=d.get("num_workers", 0), drop_last=True)
    trainer = Trainer(model, dl, train_batch_size=d["batch_size"], gradient_accumulate_every=t["gradient_accumulate_every"],
                      train_lr=t["lr"], train_num_steps=args.steps or t["train_num_steps"],
                      save_and_sample_every=t["save_and_sample_every"], results_folder=t["results_folder"],
                      amp=t.get("amp", False), fp16=t.get("fp16", False), log_freq=t["log_freq"],
                      resume_milestone=t.get("resume_milestone", 0), cfg=cfg)
    trainer.train()
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
