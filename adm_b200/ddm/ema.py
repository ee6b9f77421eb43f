"""Exponential moving average of the model weights: mirror of /root/reference/ddm/ema.py (``EMA`` :21-191) — same
constructor arguments, warm-up decay ``1 - (1 + epoch / inv_gamma) ** -power`` clamped to [min_value, beta] (:132-139),
``update()`` cadence (:141-156) and state_dict layout (``online_model.*``, ``ema_model.*``, ``initted``, ``step``).

Differences in execution only: the ~1600 per-tensor ``lerp_`` launches of the reference become ONE multi-tensor launch
(``torch._foreach_lerp_``), and because the moving average is written through ``.data`` (which does not bump tensor
versions) every fused UNet engine inside the EMA copy is told to re-derive its cached bf16 operands.
"""
from __future__ import annotations

import copy

import torch
from torch import nn


def exists(val):
    return val is not None


def clamp(value, min_value=None, max_value=None):
    assert exists(min_value) or exists(max_value)
    if exists(min_value):
        value = max(value, min_value)
    if exists(max_value):
        value = min(value, max_value)
    return value


def _engines(model):
    return [m.engine for m in model.modules() if hasattr(m, "engine") and hasattr(m.engine, "invalidate")]


def _detached_copy(model):
    """deepcopy that leaves training-arena plumbing behind: CUDA graphs are not copyable, and parameters re-homed in a
    TrainStep arena carry bf16-shadow attributes that must not follow the copy."""
    stash = [(m, m.__dict__.pop("_sample_graphs")) for m in model.modules() if "_sample_graphs" in m.__dict__]
    hooks = [(e, e.grad_hook, e.affine_pack, e._cache) for e in _engines(model)]
    for e, *_ in hooks:
        e.grad_hook, e.affine_pack, e._cache = None, None, {}
    try:
        new = copy.deepcopy(model)
    finally:
        for m, g in stash:
            m.__dict__["_sample_graphs"] = g
        for e, h, a, c in hooks:
            e.grad_hook, e.affine_pack, e._cache = h, a, c
    for p in new.parameters():
        p.__dict__.pop("_adm_pack", None)
        p.data = p.data.clone(memory_format=torch.contiguous_format)
        p.grad = None
    return new


class EMA(nn.Module):
    def __init__(self, model, ema_model=None, beta=0.9999, update_after_step=100, update_every=10, inv_gamma=1.0,
                 power=2 / 3, min_value=0.0, param_or_buffer_names_no_ema=set(), ignore_names=set(),
                 ignore_startswith_names=set(), include_online_model=True):
        super().__init__()
        self.beta = beta
        self.include_online_model = include_online_model
        if include_online_model:
            self.online_model = model
        else:
            self.online_model = [model]  # not registered as a sub-module
        self.ema_model = ema_model if exists(ema_model) else _detached_copy(model)
        self.ema_model.requires_grad_(False)
        self.parameter_names = {n for n, p in self.ema_model.named_parameters() if p.dtype == torch.float}
        self.buffer_names = {n for n, b in self.ema_model.named_buffers() if b.dtype == torch.float}
        self.update_every, self.update_after_step = update_every, update_after_step
        self.inv_gamma, self.power, self.min_value = inv_gamma, power, min_value
        assert isinstance(param_or_buffer_names_no_ema, (set, list))
        self.param_or_buffer_names_no_ema = param_or_buffer_names_no_ema
        self.ignore_names = ignore_names
        self.ignore_startswith_names = ignore_startswith_names
        self.register_buffer("initted", torch.Tensor([False]))
        self.register_buffer("step", torch.tensor([0]))

    @property
    def model(self):
        return self.online_model if self.include_online_model else self.online_model[0]

    def restore_ema_model_device(self):
        self.ema_model.to(self.initted.device)

    def get_params_iter(self, model):
        for name, param in model.named_parameters():
            if name in self.parameter_names:
                yield name, param

    def get_buffers_iter(self, model):
        for name, buffer in model.named_buffers():
            if name in self.buffer_names:
                yield name, buffer

    def _invalidate(self):
        for e in _engines(self.ema_model):
            e.invalidate()

    @torch.no_grad()
    def copy_params_from_model_to_ema(self):
        for (_, ma), (_, cur) in zip(self.get_params_iter(self.ema_model), self.get_params_iter(self.model)):
            ma.data.copy_(cur.data)
        for (_, ma), (_, cur) in zip(self.get_buffers_iter(self.ema_model), self.get_buffers_iter(self.model)):
            ma.data.copy_(cur.data)
        self._invalidate()

    def get_current_decay(self):
        epoch = clamp(self.step.item() - self.update_after_step - 1, min_value=0.)
        value = 1 - (1 + epoch / self.inv_gamma) ** - self.power
        if epoch <= 0:
            return 0.
        return clamp(value, min_value=self.min_value, max_value=self.beta)

    def update(self):
        step = self.step.item()
        self.step += 1
        if (step % self.update_every) != 0:
            return
        if step <= self.update_after_step:
            self.copy_params_from_model_to_ema()
            return
        if not self.initted.item():
            self.copy_params_from_model_to_ema()
            self.initted.data.copy_(torch.Tensor([True]))
        self.update_moving_average(self.ema_model, self.model)

    @torch.no_grad()
    def update_moving_average(self, ma_model, current_model):
        decay = self.get_current_decay()
        lerp_ma, lerp_cur = [], []
        for it in (self.get_params_iter, self.get_buffers_iter):
            for (name, cur), (_, ma) in zip(it(current_model), it(ma_model)):
                if name in self.ignore_names or any(name.startswith(p) for p in self.ignore_startswith_names):
                    continue
                if name in self.param_or_buffer_names_no_ema:
                    ma.data.copy_(cur.data)
                    continue
                lerp_ma.append(ma.data)
                lerp_cur.append(cur.data if cur.data.is_contiguous() == ma.data.is_contiguous() else cur.data.contiguous())
        if lerp_ma:
            torch._foreach_lerp_(lerp_ma, lerp_cur, 1. - decay)  # one multi-tensor launch
        self._invalidate()

    def __call__(self, *args, **kwargs):
        return self.ema_model(*args, **kwargs)
