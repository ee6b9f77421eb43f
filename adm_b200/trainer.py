"""Trainer / Sampler shells around the hot path: mirror of ``Trainer`` in /root/reference/train_uncond_dpm.py:117-365 and
``Sampler`` in sample_uncond.py:90-191, without HF accelerate (one process per GPU under torchrun; gradients are reduced
by ``adm_b200.train.TrainStep``).  Same constructor arguments, same loop (micro-batches, clip 1.0, AdamW + warm-up /
polynomial LambdaLR :169-182, EMA on the main process :184-189, save-and-sample cadence :311-334) and the same checkpoint
dict ``{'step','model','opt','lr_scheduler','ema','scaler'}`` under ``results_folder/model-{k}.pt`` (:205-231).
``'opt'`` holds the fused optimizer's flat moments (not a torch.optim state_dict); ``'model'`` and ``'ema'`` are
interchangeable with the reference's checkpoints.
"""
from __future__ import annotations

import math
import os
from pathlib import Path

import torch
import torch.distributed as dist

from .ddm.ema import EMA
from .train import TrainStep


def cycle(dl):
    while True:
        for data in dl:
            yield data


def has_int_squareroot(num):
    return (math.sqrt(num) ** 2) == num


def _get(cfg, key, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class Trainer(object):
    def __init__(self, model, data_loader, train_batch_size=16, gradient_accumulate_every=1, train_lr=1e-4,
                 train_wd=1e-4, train_num_steps=100000, save_and_sample_every=1000, num_samples=25,
                 results_folder="./results", amp=False, fp16=False, split_batches=True, log_freq=20, resume_milestone=0,
                 cfg={}):
        assert has_int_squareroot(num_samples), "number of samples must have an integer square root"
        self.model = model
        self.num_samples, self.save_and_sample_every = num_samples, save_and_sample_every
        self.batch_size, self.gradient_accumulate_every = train_batch_size, gradient_accumulate_every
        self.log_freq, self.train_num_steps = log_freq, train_num_steps
        self.image_size = model.image_size
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.is_main = self.rank == 0
        self.dl = cycle(data_loader)
        tcfg = _get(cfg, "trainer", {})
        warmup_iter, min_lr = _get(tcfg, "warmup_iter", 5000), _get(tcfg, "min_lr", 1e-6)

        def warm_up_lr(it):  # train_uncond_dpm.py:169-177
            if it <= warmup_iter:
                return (it + 1) / warmup_iter
            return max((1 - (it - warmup_iter) / train_num_steps) ** 0.96, min_lr / train_lr)

        self.step_fn = TrainStep(model, lr=train_lr, weight_decay=train_wd, max_grad_norm=1.0,
                                 grad_accum=gradient_accumulate_every, lr_schedule=warm_up_lr)
        self.results_folder = Path(results_folder)
        if self.is_main:
            self.results_folder.mkdir(exist_ok=True, parents=True)
            self.ema = EMA(model, ema_model=None, beta=0.9996, update_after_step=_get(tcfg, "ema_update_after_step", 100),
                           update_every=_get(tcfg, "ema_update_every", 10))
        self.step = 0
        self.last_loss = None
        if os.path.isfile(str(self.results_folder / f"model-{resume_milestone}.pt")):
            self.load(resume_milestone)

    # ------------------------------------------------------------------------------------------ checkpoints
    def save(self, milestone):
        if not self.is_main:
            return
        s = self.step_fn
        data = {"step": self.step, "model": self.model.state_dict(),
                "opt": {"m": s.m, "v": s.v, "step_count": s.step_count, "format": "adm_b200 flat AdamW moments",
                        # dropout-mask stream position: device step counter + the engine's per-forward seed index
                        "seed_counter": int(s.seed_counter.item()) if s.seed_counter is not None else 0,
                        "seed_iter": s.engine.seed_position()},
                "lr_scheduler": {"last_epoch": s.step_count}, "ema": self.ema.state_dict(), "scaler": None}
        torch.save(data, str(self.results_folder / f"model-{milestone}.pt"))

    def load(self, milestone):
        data = torch.load(str(self.results_folder / f"model-{milestone}.pt"), map_location="cpu", weights_only=False)
        self.model.load_state_dict(data["model"])
        self.step = data["step"]
        s = self.step_fn
        opt = data.get("opt") or {}
        if isinstance(opt, dict) and "m" in opt and opt["m"].numel() == s.m.numel():
            s.m.copy_(opt["m"])
            s.v.copy_(opt["v"])
            s.step_count = int(opt.get("step_count", self.step))
            if s.seed_counter is not None:
                s.seed_counter.fill_(int(opt.get("seed_counter", 0)))
            s.engine.seed_position(int(opt.get("seed_iter", 0)))
        else:  # a reference checkpoint: torch.optim state is not transferable to the flat arena; restart the moments
            s.step_count = self.step
        if self.is_main and "ema" in data:
            self.ema.load_state_dict(data["ema"])
            self.ema.invalidate_engines()
        s.sync_params()  # every rank ends up with rank 0's weights and moments (only rank 0 is guaranteed the file)

    # ------------------------------------------------------------------------------------------ the loop
    def train_one_step(self):
        dev = next(self.model.parameters()).device
        batches = []
        for _ in range(self.gradient_accumulate_every):
            batch = next(self.dl)
            x = batch["image"] if isinstance(batch, dict) else batch
            batches.append(x.to(dev, non_blocking=True))
        loss = self.step_fn(batches)
        self.step += 1
        if self.is_main:
            self.ema.update()
        self.last_loss = loss
        return loss

    def train(self, on_milestone=None):
        while self.step < self.train_num_steps:
            loss = self.train_one_step()
            if self.is_main and self.step % self.log_freq == 0:
                print(f"[Train Step] {self.step}/{self.train_num_steps}: loss {loss.item():.4f}", flush=True)
            if self.step != 0 and self.step % self.save_and_sample_every == 0:
                milestone = self.step // self.save_and_sample_every
                self.save(milestone)
                if self.is_main and on_milestone is not None:
                    on_milestone(self, milestone)
        if self.is_main:
            print("training complete", flush=True)


class Sampler(object):
    """sample_uncond.py:90-191: loads ``model-{k}.pt`` (EMA weights when use_ema) and draws ``sample_num`` images in
    batches; returns them as a tensor (saving PNGs / FID are outside the hot path)."""

    def __init__(self, model, batch_size=128, sample_num=1000, results_folder="./results", ckpt_path=None, use_ema=True,
                 cfg={}):
        self.model, self.batch_size, self.sample_num = model, batch_size, sample_num
        self.results_folder = Path(results_folder)
        if ckpt_path is not None:
            data = torch.load(ckpt_path, map_location="cpu", weights_only=False)
            if use_ema and "ema" in data:
                sd = {k[len("ema_model."):]: v for k, v in data["ema"].items() if k.startswith("ema_model.")}
            else:
                sd = data["model"]
            self.model.load_state_dict(sd)

    @torch.no_grad()
    def sample(self):
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        mine = (self.sample_num + world - 1 - rank) // world  # the batch is sharded by rank, no communication
        out = []
        while sum(o.shape[0] for o in out) < mine:
            b = min(self.batch_size, mine - sum(o.shape[0] for o in out))
            out.append(self.model.sample(batch_size=b))
        return torch.cat(out) if out else torch.empty(0)
