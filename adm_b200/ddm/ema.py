"""Exponential moving average of the model weights as ONE flat lerp over the parameter arena.

Role of /root/reference/ddm/ema.py (``EMA`` :21-191) for the Trainer / Sampler shells: same constructor keywords, the
same warm-up decay ``1 - (1 + epoch / inv_gamma) ** -power`` clamped to [min_value, beta] (:132-139), the same
``update()`` cadence (:141-156: hard copies until ``update_after_step``, then a lerp every ``update_every`` calls) and
the same state_dict layout (``online_model.*``, ``ema_model.*``, ``initted``, ``step``) so checkpoints interchange.

Execution is different.  When the online parameters live in a training arena (``adm_b200.train.ParamArena``: every
parameter is a view of one flat fp32 buffer) the averaged copy is laid out as a second flat buffer with identical
offsets and strides, and an update is a single HBM-bound kernel ``ema += w * (online - ema)`` over the whole arena
(``adm_lerp_f32``, 12 B per parameter) instead of ~1600 per-tensor launches.  Parameters outside an arena (host tests,
float buffers) take per-tensor ``lerp_``.  The moving average is written through raw pointers, so the fused UNet
engines inside the averaged copy are told to re-derive their cached bf16 operands.
"""
from __future__ import annotations

import copy

import torch
from torch import nn


def ema_decay(step, update_after_step=100, inv_gamma=1.0, power=2 / 3, min_value=0.0, beta=0.9999):
    """Decay used by the update that happens when the step counter reads ``step`` (ddm/ema.py:132-139)."""
    epoch = max(step - update_after_step - 1, 0.)
    if epoch <= 0:
        return 0.
    return min(max(1 - (1 + epoch / inv_gamma) ** -power, min_value), beta)


def _engines(model):
    return [m.engine for m in model.modules() if hasattr(m, "engine") and hasattr(m.engine, "invalidate")]


def _detached_copy(model):
    """deepcopy that leaves training-arena plumbing behind: CUDA graphs are not copyable, and parameters re-homed in a
    TrainStep arena carry bf16-shadow attributes that must not follow the copy."""
    stash = [(m, k, m.__dict__.pop(k)) for m in model.modules() for k in ("_sample_graphs", "_latent_graphs")
             if k in m.__dict__]
    hooks = [(e, e.grad_hook, e.affine_pack, e._cache, e.seed_counter) for e in _engines(model)]
    for e, *_ in hooks:
        e.grad_hook, e.affine_pack, e._cache, e.seed_counter = None, None, {}, None
    try:
        new = copy.deepcopy(model)
    finally:
        for m, k, g in stash:
            m.__dict__[k] = g
        for e, h, a, c, s in hooks:
            e.grad_hook, e.affine_pack, e._cache, e.seed_counter = h, a, c, s
    for p in new.parameters():
        p.__dict__.pop("_adm_pack", None)
        p.__dict__.pop("_adm_qkv", None)
        p.grad = None
    return new


def _shared_flat(params):
    """(flat fp32 tensor covering the storage, True) when every tensor is a view of ONE fp32 CUDA storage."""
    if not params or any((not p.is_cuda) or p.dtype != torch.float32 for p in params):
        return None
    first = params[0].untyped_storage()
    if any(p.untyped_storage().data_ptr() != first.data_ptr() for p in params):
        return None
    n = first.nbytes() // 4
    return torch.empty(0, device=params[0].device, dtype=torch.float32).set_(first, 0, (n,), (1,))


class EMA(nn.Module):
    def __init__(self, model, ema_model=None, beta=0.9999, update_after_step=100, update_every=10, inv_gamma=1.0,
                 power=2 / 3, min_value=0.0, param_or_buffer_names_no_ema=(), ignore_names=(),
                 ignore_startswith_names=(), include_online_model=True):
        super().__init__()
        self.beta, self.update_every, self.update_after_step = beta, update_every, update_after_step
        self.inv_gamma, self.power, self.min_value = inv_gamma, power, min_value
        self.include_online_model = include_online_model
        if include_online_model:
            self.online_model = model
        else:
            self._online = [model]  # kept out of the module tree (and of the state_dict)
        self.ema_model = ema_model if ema_model is not None else _detached_copy(model)
        self.ema_model.requires_grad_(False)
        self._copy_only = set(param_or_buffer_names_no_ema)
        self._skip = (set(ignore_names), tuple(ignore_startswith_names))
        self.register_buffer("initted", torch.Tensor([False]))
        self.register_buffer("step", torch.tensor([0]))
        self._plan = None

    @property
    def model(self):
        return self.online_model if self.include_online_model else self._online[0]

    # ------------------------------------------------------------------------------------------ layout
    def _float_tensors(self, module):
        out = {n: p for n, p in module.named_parameters() if p.dtype == torch.float}
        out.update({n: b for n, b in module.named_buffers() if b.dtype == torch.float})
        return out

    def _build_plan(self):
        """Pairs (ema, online) by name; when the online parameters are views of one arena, re-home the averaged copy
        into a flat buffer with the same offsets/strides so that ONE kernel updates all of them."""
        on, av = self._float_tensors(self.model), self._float_tensors(self.ema_model)
        names = [n for n in av if n in on and n not in self._skip[0] and not n.startswith(self._skip[1] or ("\0",))]
        lerp = [n for n in names if n not in self._copy_only]
        flat_on = _shared_flat([on[n].data for n in lerp if isinstance(on[n], nn.Parameter)])
        flat_ema, loose = None, lerp
        if flat_on is not None and not self._copy_only:
            flat_ema = torch.empty_like(flat_on)
            flat_ema.copy_(flat_on)
            homed = []
            for n in lerp:
                src = on[n]
                if isinstance(src, nn.Parameter):
                    view = flat_ema.as_strided(src.shape, src.stride(), src.storage_offset())
                    view.copy_(av[n].data)
                    av[n].data = view
                    homed.append(n)
            loose = [n for n in lerp if n not in set(homed)]
        self._plan = dict(flat=(flat_ema, flat_on), loose=[(av[n], on[n]) for n in loose],
                          copy=[(av[n], on[n]) for n in names if n in self._copy_only])

    def invalidate_engines(self):
        for e in _engines(self.ema_model):
            e.invalidate()

    _invalidate = invalidate_engines

    # ------------------------------------------------------------------------------------------ updates
    @torch.no_grad()
    def _blend(self, weight):
        """ema <- ema + weight * (online - ema); weight == 1 is a hard copy."""
        key = tuple(p.data_ptr() for p in self.model.parameters())  # re-homed parameters (a new arena) re-plan
        if self._plan is None or self._plan["key"] != key:
            self._build_plan()
            self._plan["key"] = key
        flat_ema, flat_on = self._plan["flat"]
        if flat_ema is not None:
            if weight >= 1.:
                flat_ema.copy_(flat_on)
            else:
                from .. import ops
                ops.lerp_f32(flat_ema, flat_on, weight)
        for ma, cur in self._plan["loose"]:
            if weight >= 1.:
                ma.data.copy_(cur.data)
            else:
                ma.data.lerp_(cur.data.to(ma.dtype), weight)
        for ma, cur in self._plan["copy"]:
            ma.data.copy_(cur.data)
        self.invalidate_engines()

    def copy_params_from_model_to_ema(self):
        self._blend(1.)

    def get_current_decay(self):
        return ema_decay(self.step.item(), self.update_after_step, self.inv_gamma, self.power, self.min_value,
                         self.beta)

    def update(self):
        step = int(self.step.item())
        self.step += 1
        if step % self.update_every:
            return
        if step <= self.update_after_step:
            self._blend(1.)
            return
        if not bool(self.initted.item()):
            self._blend(1.)
            self.initted.fill_(1.)
        self._blend(1. - self.get_current_decay())

    def load_state_dict(self, state_dict, strict=True, assign=False):
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self.invalidate_engines()
        return out

    def forward(self, *args, **kwargs):
        return self.ema_model(*args, **kwargs)
