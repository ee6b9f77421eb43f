#!/usr/bin/env python
"""Latent training entry point: the role of /root/reference/train_uncond_ldm.py (`accelerate launch train_uncond_ldm.py
--cfg ...`), without accelerate:

    torchrun --nproc-per-node N scripts/train_uncond_ldm.py --cfg configs/celebahq/celeb_uncond_ddm_const_uncond_unet_ldm.yaml

Builds the frozen first stage, the UNet and the LatentDiffusion module from the YAML with construct_class_by_name
(train_uncond_ldm.py:40-57) and runs adm_b200.trainer.Trainer.  What the reference does inside
``LatentDiffusion.training_step`` (ddm_const_2.py:494-524: frozen AE encode under no_grad, x scale_factor) is done here on
the batch before it reaches the fused DDM step (``TrainStep``: forward, backward, clip, AdamW, data-parallel all-reduce),
exactly as ``bench.py --config celebahq`` measures it.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adm_b200.ddm.utils import construct_class_by_name  # noqa: E402
from adm_b200.trainer import Trainer  # noqa: E402
from scripts.train_uncond_dpm import synthetic_loader  # noqa: E402


def build_model(cfg, device):
    """unet, first stage and the diffusion module from the reference-schema YAML (train_uncond_ldm.py:40-57)."""
    model_cfg = dict(cfg["model"])
    first_stage = construct_class_by_name(**dict(model_cfg.pop("first_stage")))
    unet = construct_class_by_name(**dict(model_cfg.pop("unet")))
    cls = model_cfg.pop("class_name")
    return construct_class_by_name(class_name=cls, model=unet, auto_encoder=first_stage, cfg=model_cfg,
                                   **model_cfg).to(device)


class LatentLoader:
    """Images -> scaled latents of the frozen first stage (ddm_const_2.py:494-524) in front of the DDM step; re-iterable,
    so a finite DataLoader cycles the way the Trainer expects."""

    def __init__(self, ldm, dl, device):
        self.ldm, self.dl, self.device, self.first = ldm, dl, device, True

    def __iter__(self):
        ldm = self.ldm
        for batch in self.dl:
            batch = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}
            with torch.no_grad():
                if self.first:
                    ldm.on_train_batch_start(batch)  # std-rescaling from the first batch unless default_scale (:473-491)
                    self.first = False
                z, *_ = ldm.get_input(batch)
                if ldm.scale_by_softsign:
                    z = torch.nn.functional.softsign(z)
                elif ldm.scale_by_std:
                    z = ldm.scale_factor * z
            yield {"image": z}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", required=True)
    ap.add_argument("--steps", type=int, default=None, help="override trainer.train_num_steps")
    ap.add_argument("--batch", type=int, default=None, help="override data.batch_size (per GPU)")
    args = ap.parse_args()
    cfg = yaml.safe_load(open(args.cfg))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=device)
    rank = dist.get_rank() if dist.is_initialized() else 0
    torch.manual_seed(1234)  # identical initial weights on every rank (TrainStep also broadcasts rank 0's arena)
    ldm = build_model(cfg, device)
    torch.manual_seed(1234 + rank)  # per-rank streams for t, noise and dropout from here on
    t, d = cfg["trainer"], cfg["data"]
    bs = args.batch or d["batch_size"]
    if d.get("class_name", "synthetic") == "synthetic":
        dl = synthetic_loader(bs, d["image_size"], seed=rank)
    else:
        ds = construct_class_by_name(**{k: v for k, v in d.items() if k not in ("batch_size", "num_workers")})
        dl = torch.utils.data.DataLoader(ds, batch_size=bs, shuffle=True, pin_memory=True,
                                         num_workers=d.get("num_workers", 0), drop_last=True)
    trainer = Trainer(ldm, LatentLoader(ldm, dl, device), train_batch_size=bs,
                      gradient_accumulate_every=t["gradient_accumulate_every"], train_lr=t["lr"],
                      train_num_steps=args.steps or t["train_num_steps"],
                      save_and_sample_every=t["save_and_sample_every"], results_folder=t["results_folder"],
                      amp=t.get("amp", False), fp16=t.get("fp16", False), log_freq=t["log_freq"],
                      resume_milestone=t.get("resume_milestone", 0), cfg=cfg)
    trainer.train()
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
