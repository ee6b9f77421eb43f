#!/usr/bin/env python
"""Conditional latent training entry point: the role of /root/reference/train_cond_ldm.py (`accelerate launch
train_cond_ldm.py --cfg ...`) on one GPU, without accelerate:

    python scripts/train_cond_ldm.py --cfg configs/super-resolution/div2k_cond_ddm_const_ldm.yaml [--graph]

Builds the frozen first stage, the conditional UNet and the LatentDiffusion module from the YAML with
construct_class_by_name (train_cond_ldm.py:37-62) and runs the reference's loop (:212-330): micro-batches through
``LatentDiffusion.training_step`` (frozen AE encode + DDM-const step on the sm_100a kernels under torch autograd), clip 1.0,
torch AdamW, warm-up / polynomial LambdaLR, EMA, ``model-{k}.pt`` checkpoints with the reference's keys (the optimizer here IS
torch.optim.AdamW, so 'opt' and 'lr_scheduler' interchange with the reference's files too).  ``--graph`` replays the whole
step (AE encode, forward, autograd backward, clip, AdamW) as one CUDA graph, the way ``bench.py --config div2k`` measures it.
"""
import argparse
import os
import sys
from pathlib import Path

import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adm_b200.ddm.ema import EMA  # noqa: E402
from scripts.train_uncond_ldm import build_model  # noqa: E402


def synthetic_loader(batch, image_size, down, seed):
    """{'image': [B,3,H,W], 'cond': [B,3,H/down,W/down]} in [-1, 1]: the super-resolution pairs of the DIV2K config."""
    g = torch.Generator().manual_seed(seed)
    while True:
        yield {"image": (2 * torch.rand(batch, 3, *image_size, generator=g) - 1).pin_memory(),
               "cond": (2 * torch.rand(batch, 3, image_size[0] // down, image_size[1] // down, generator=g) - 1).pin_memory()}


class CondTrainer:
    """train_cond_ldm.py:96-330 without accelerate / tensorboard."""

    def __init__(self, model, data_loader, gradient_accumulate_every=1, train_lr=1e-4, train_wd=1e-4,
                 train_num_steps=100000, save_and_sample_every=1000, results_folder="./results", log_freq=20,
                 resume_milestone=0, cfg=None, use_graph=False):
        cfg = cfg or {}
        self.model, self.dl = model, iter(data_loader)
        self.accum, self.train_num_steps = gradient_accumulate_every, train_num_steps
        self.save_and_sample_every, self.log_freq, self.train_lr = save_and_sample_every, log_freq, train_lr
        tcfg = cfg.get("trainer", {})
        warmup_iter, min_lr = tcfg.get("warmup_iter", 5000), tcfg.get("min_lr", 1e-6)

        def warm_up_lr(it):  # train_cond_ldm.py:139-146
            if it <= warmup_iter:
                return (it + 1) / warmup_iter
            return max((1 - (it - warmup_iter) / train_num_steps) ** 0.96, min_lr / train_lr)
        self.lr_lambda = warm_up_lr
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.device = self.params[0].device
        self.use_graph = use_graph and gradient_accumulate_every == 1
        # the learning rate lives in a device scalar so that a captured step replays with the scheduler's current value
        self.lr_t = torch.tensor(train_lr * warm_up_lr(0), device=self.device, dtype=torch.float32)
        self.opt = torch.optim.AdamW(self.params, lr=self.lr_t, weight_decay=train_wd, fused=True,
                                     capturable=self.use_graph)
        self.results_folder = Path(results_folder)
        self.results_folder.mkdir(exist_ok=True, parents=True)
        self.ema = EMA(model, ema_model=None, beta=0.9996, update_after_step=tcfg.get("ema_update_after_step", 100),
                       update_every=tcfg.get("ema_update_every", 10))
        self.step, self.graph, self.static = 0, None, None
        if os.path.isfile(str(self.results_folder / f"model-{resume_milestone}.pt")):
            self.load(resume_milestone)

    # ------------------------------------------------------------------------------------------ checkpoints (:176-210)
    def save(self, milestone):
        data = {"step": self.step, "model": self.model.state_dict(), "opt": self.opt.state_dict(),
                "lr_scheduler": {"last_epoch": self.step}, "ema": self.ema.state_dict(), "scaler": None}
        torch.save(data, str(self.results_folder / f"model-{milestone}.pt"))

    def load(self, milestone):
        data = torch.load(str(self.results_folder / f"model-{milestone}.pt"), map_location="cpu", weights_only=False)
        self.model.load_state_dict(data["model"])
        self.step = data["step"]
        self.opt.load_state_dict(data["opt"])
        self.opt.param_groups[0]["lr"] = self.lr_t  # keep the device scalar (load_state_dict restores a copy)
        if "ema" in data:
            self.ema.load_state_dict(data["ema"])

    # ------------------------------------------------------------------------------------------ the loop (:212-330)
    def _to_device(self, batch):
        return {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}

    def _one(self, batches):
        self.opt.zero_grad(set_to_none=True)
        total = None
        for batch in batches:
            loss, _ = self.model.training_step(batch)
            (loss / len(batches)).backward()
            total = loss.detach() if total is None else total + loss.detach()
        torch.nn.utils.clip_grad_norm_(self.params, 1.0)
        self.opt.step()
        return total / len(batches)

    def train_one_step(self):
        self.lr_t.fill_(self.train_lr * self.lr_lambda(self.step))
        batches = [self._to_device(next(self.dl)) for _ in range(self.accum)]
        if self.use_graph:
            if self.graph is None and self.step >= 3:  # three eager steps first (allocator warm-up, lazy initialisation)
                self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batches[0].items()}
                torch.cuda.synchronize()
                self.opt.zero_grad(set_to_none=True)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=torch.cuda.current_stream()):
                    self.static_loss = self._one([self.static])
            if self.graph is not None:
                for k, v in batches[0].items():
                    if torch.is_tensor(v):
                        self.static[k].copy_(v, non_blocking=True)
                self.graph.replay()
                loss = self.static_loss
            else:
                loss = self._one(batches)
        else:
            loss = self._one(batches)
        self.step += 1
        self.ema.update()
        return loss

    def train(self):
        while self.step < self.train_num_steps:
            loss = self.train_one_step()
            if self.step % self.log_freq == 0:
                print(f"[Train Step] {self.step}/{self.train_num_steps}: loss {loss.item():.4f} lr {self.lr_t.item():.3e}",
                      flush=True)
            if self.step != 0 and self.step % self.save_and_sample_every == 0:
                self.save(self.step // self.save_and_sample_every)
        print("training complete", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", required=True)
    ap.add_argument("--steps", type=int, default=None, help="override trainer.train_num_steps")
    ap.add_argument("--batch", type=int, default=None, help="override data.batch_size")
    ap.add_argument("--graph", action="store_true", help="replay the step as one CUDA graph (gradient_accumulate_every 1)")
    ap.add_argument("--log-freq", type=int, default=None)
    args = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("train_cond_ldm.py runs on one GPU: the conditional UNet trains through torch autograd and a torch "
                         "optimizer; data parallel training is wired for the TrainStep path (train_uncond_dpm / _ldm)")
    cfg = yaml.safe_load(open(args.cfg))
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    # autograd's AccumulateGrad nodes must not be bound to the default stream if the step is to be captured: run on a side one
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    torch.cuda.set_stream(side)
    torch.manual_seed(1234)
    ldm = build_model(cfg, device)
    ldm.train()
    t, d = cfg["trainer"], cfg["data"]
    bs = args.batch or d["batch_size"]
    if d.get("class_name", "synthetic") == "synthetic":
        dl = synthetic_loader(bs, d["image_size"], ldm.first_stage_model.down_ratio, seed=0)
    else:
        from adm_b200.ddm.utils import construct_class_by_name
        from adm_b200.trainer import cycle
        ds = construct_class_by_name(**{k: v for k, v in d.items() if k not in ("batch_size", "num_workers")})
        dl = cycle(torch.utils.data.DataLoader(ds, batch_size=bs, shuffle=True, pin_memory=True,
                                               num_workers=d.get("num_workers", 0), drop_last=True))
    trainer = CondTrainer(ldm, dl, gradient_accumulate_every=1 if args.graph else t["gradient_accumulate_every"],
                          train_lr=t["lr"], train_num_steps=args.steps or t["train_num_steps"],
                          save_and_sample_every=t["save_and_sample_every"], results_folder=t["results_folder"],
                          log_freq=args.log_freq or t["log_freq"], resume_milestone=t.get("resume_milestone", 0), cfg=cfg,
                          use_graph=args.graph)
    trainer.train()


if __name__ == "__main__":
    main()
