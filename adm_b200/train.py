"""Data-parallel training step around the DDPM module: flat parameter/gradient arenas, bucketed NCCL all-reduce
overlapped with the engine's backward, fused clip + AdamW over the arena.

Mirrors the step body of the reference Trainer (/root/reference/train_uncond_dpm.py:247-310): micro-batch loop,
``clip_grad_norm_(1.0)`` (:292), AdamW(lr, weight_decay=1e-4) (:179-180), warm-up + polynomial LambdaLR (:169-182).
Semantics we define explicitly (the reference's DDP wiring is accidental): gradients are the MEAN over
ranks of per-rank gradients, reduced once per optimizer step (not per micro-batch).
"""
from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist

from . import ops


def completion_order(net):
    """Parameters of EDMPrecond in the order their gradients are finalised by UNetEngine.backward."""
    m = net.model
    order = []

    def block_params(blk):
        names = []
        if blk.num_heads:
            names += ["proj", "qkv", "norm2"]
        names += ["conv1", "norm1", "conv0", "skip", "norm0"]
        for nme in names:
            sub = getattr(blk, nme, None)
            if sub is not None:
                order.extend(p for p in sub.parameters())

    for dec, dname, norm, oconv in ((m.dec2, "decouple2", m.out_norm2, m.out_conv2),
                                    (m.dec, "decouple1", m.out_norm, m.out_conv)):
        order.extend(oconv.parameters())
        order.extend(norm.parameters())
        for blk in reversed(list(dec.values())):
            block_params(blk)
        order.extend(getattr(m, dname).parameters())
    for blk in reversed(list(m.enc.values())):
        if hasattr(blk, "affine"):
            block_params(blk)
        else:
            order.extend(blk.parameters())
    # every block's affine (time-embedding) weight, then every bias, back to back in the engine's block order:
    # their gradients all come from ONE grouped GEMM at the end of backward, and the contiguous layout lets that GEMM
    # (and the forward one) address them as a single [sum 2*Cout, emb] matrix inside the arena.
    affine = [blk.affine for sec in (m.enc, m.dec, m.dec2) for blk in sec.values() if hasattr(blk, "affine")]
    order.extend(a.weight for a in affine)
    order.extend(a.bias for a in affine)
    order.extend(m.map_layer1.parameters())
    order.extend(m.map_layer0.parameters())
    if m.map_augment is not None:
        order.extend(m.map_augment.parameters())
    seen, out = set(), []
    for p in order:
        if id(p) not in seen:
            seen.add(id(p))
            out.append(p)
    rest = [p for p in net.parameters() if id(p) not in seen]
    return out + rest


class ParamArena:
    """Re-homes every parameter (and its .grad) of a module into two flat fp32 buffers, in gradient-completion order."""

    def __init__(self, module, order=None, align=8, channels_last=()):
        """channels_last: conv weights [Cout, Cin, k, k] to store physically as [Cout][k][k][Cin] — the packed operand
        order of the GEMM engine.  They keep their logical (reference) shape as a permuted view, so state_dict /
        load_state_dict are unaffected, and on CUDA they get a bf16 shadow (written by the fused AdamW kernel) plus a
        packed view of their gradient: ``p._adm_pack = (bf16 [Cout, k*k, Cin], fp32 grad [Cout, k*k, Cin], version,
        fp32 flat slice)``."""
        params = order if order is not None else list(module.parameters())
        dev = params[0].device
        cl = {id(p) for p in channels_last}
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + align - 1) // align * align
        self.params = params
        self.offsets = offs
        self.numel = total
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grads = torch.zeros(total, device=dev, dtype=torch.float32)
        self.shadow = torch.zeros(total, device=dev, dtype=torch.bfloat16) if (cl and dev.type == "cuda") else None
        self._grad_views = []
        for p, o in zip(params, offs):
            n = p.numel()
            if id(p) in cl:
                co, ci, kh, kw = p.shape
                view = self.flat[o:o + n].view(co, kh, kw, ci).permute(0, 3, 1, 2)
                gview = self.grads[o:o + n].view(co, kh, kw, ci).permute(0, 3, 1, 2)
            else:
                view = self.flat[o:o + n].view(p.shape)
                gview = self.grads[o:o + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = gview
            self._grad_views.append(gview)
            if id(p) in cl and self.shadow is not None:
                p._adm_pack = (self.shadow[o:o + n].view(co, kh * kw, ci), self.grads[o:o + n].view(co, kh * kw, ci),
                               p._version, self.flat[o:o + n])
        self.refresh_shadow()

    def refresh_shadow(self):
        """Re-derive the bf16 shadow from the fp32 masters (after construction or an out-of-band parameter write)."""
        if self.shadow is not None:
            ops.cast_bf16_into(self.flat, self.shadow)
            for p in self.params:
                pk = getattr(p, "_adm_pack", None)
                if pk is not None:
                    p._adm_pack = (pk[0], pk[1], p._version, pk[3])

    def slice_of(self, params):
        """(offset, numel) of a run of parameters that sit back to back in the arena, else None."""
        idx = {id(p): i for i, p in enumerate(self.params)}
        first = idx.get(id(params[0]))
        if first is None:
            return None
        o = self.offsets[first]
        end = o
        for j, p in enumerate(params):
            i = idx.get(id(p))
            if i != first + j or self.offsets[i] != end:
                return None
            end += p.numel()
        return o, end - o

    def rebind_grads(self):
        """If someone set .grad to None (optimizer.zero_grad(set_to_none=True)), point it back into the arena."""
        for p, o, gv in zip(self.params, self.offsets, self._grad_views):
            g = p.grad
            if g is None or g.data_ptr() != self.grads.data_ptr() + 4 * o:
                p.grad = gv

    def zero_grad(self):
        self.grads.zero_()


def lr_lambda(step, train_num_steps, lr, min_lr, warmup=5000):
    """train_uncond_dpm.py:169-177: linear warm-up to 1 over `warmup` steps, then (1 - iter/total)^0.96 with a floor."""
    if step < warmup:
        return (step + 1) / warmup
    return max(min_lr / lr, (1 - (step - warmup) / max(1, train_num_steps - warmup)) ** 0.96)


class TrainStep:
    """One data-parallel optimizer step of DDM-const training on this rank's GPU."""

    def __init__(self, dpm, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=1.0,
                 grad_accum=1, bucket_mb=64, process_group=None, lr_schedule=None):
        self.dpm = dpm
        self.net = dpm.model
        self.engine = self.net.model.engine
        self.arena = ParamArena(self.net, completion_order(self.net), channels_last=self.engine.packable_params())
        self._bind_affine()
        self.m = torch.zeros_like(self.arena.flat)
        self.v = torch.zeros_like(self.arena.flat)
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.max_grad_norm = max_grad_norm
        self.grad_accum = grad_accum
        self.step_count = 0
        self.lr_schedule = lr_schedule
        dev = self.arena.flat.device
        self.sqnorm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.hyper = torch.zeros(3, device=dev, dtype=torch.float32)
        self.hyper_host = torch.zeros(3, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(3)
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.pg = process_group
        # gradient buckets in completion order
        n = self.arena.numel
        per = max(1, bucket_mb * (1 << 20) // 4)
        self.buckets = [(s, min(n, s + per)) for s in range(0, n, per)]
        self.comm_stream = torch.cuda.Stream(device=dev) if (self.world > 1 and dev.type == "cuda") else None
        self._param_end = {}
        for p, o in zip(self.arena.params, self.arena.offsets):
            self._param_end[id(p)] = o + p.numel()
        self.engine.invalidate()
        # device-resident step counter mixed into every dropout seed (fresh masks per CUDA-graph replay)
        self.seed_counter = torch.zeros(1, device=dev, dtype=torch.int64) if dev.type == "cuda" else None
        if self.seed_counter is not None:
            ops.set_seed_counter(self.seed_counter)
        self.graph = None

    def _bind_affine(self):
        """Hand the engine ONE [sum 2*Cout, emb] view of all affine weights (bf16 shadow), biases and their gradients."""
        a, eng = self.arena, self.engine
        eng.affine_pack = None
        if a.shadow is None:
            return
        mods = [m.affine for _, m, _ in eng.block_list]
        ws, bs = a.slice_of([m.weight for m in mods]), a.slice_of([m.bias for m in mods])
        if ws is None or bs is None:
            return
        emb = mods[0].in_features
        t = ws[1] // emb
        assert t == eng.affine_total and bs[1] == t
        eng.affine_pack = (a.shadow[ws[0]:ws[0] + ws[1]].view(t, emb), a.flat[bs[0]:bs[0] + t],
                           a.grads[ws[0]:ws[0] + ws[1]].view(t, emb), a.grads[bs[0]:bs[0] + t],
                           a.flat[ws[0]:ws[0] + ws[1]], tuple(m.weight._version for m in mods))

    def refresh(self):
        """Call after parameters were written outside the fused optimizer (load_state_dict, manual edits)."""
        self.arena.refresh_shadow()
        self.engine.invalidate()

    # ------------------------------------------------------------------------------------------ gradient reduction
    # The arena is laid out in gradient-completion order, so "everything below offset X is final" grows monotonically
    # during backward.  Each bucket is all-reduced (NCCL, sum) on a side stream the moment it is entirely final, which
    # overlaps the transfer over NVLink with the rest of the backward pass.  1/world is folded into the AdamW kernel.
    def _arm(self):
        self._done_upto, self._next_bucket = 0, 0
        self.engine.grad_hook = self._on_grads if self.world > 1 else None

    def _on_grads(self, params):
        for p in params:
            e = self._param_end.get(id(p))
            if e is not None and e > self._done_upto:
                self._done_upto = e
        self._flush(self._done_upto)

    def _flush(self, upto):
        if self._next_bucket >= len(self.buckets) or self.buckets[self._next_bucket][1] > upto:
            return
        if self.comm_stream is None:  # host tensors (gloo tests): reduce synchronously
            while self._next_bucket < len(self.buckets) and self.buckets[self._next_bucket][1] <= upto:
                s, e = self.buckets[self._next_bucket]
                dist.all_reduce(self.arena.grads[s:e], op=dist.ReduceOp.SUM, group=self.pg)
                self._next_bucket += 1
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            while self._next_bucket < len(self.buckets) and self.buckets[self._next_bucket][1] <= upto:
                s, e = self.buckets[self._next_bucket]
                dist.all_reduce(self.arena.grads[s:e], op=dist.ReduceOp.SUM, group=self.pg)
                self._next_bucket += 1

    def _allreduce(self):
        """Reduce whatever the overlapped path has not launched yet, then join the side stream."""
        if self.world == 1:
            return
        self._flush(self.arena.numel)
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.engine.grad_hook = None

    # ------------------------------------------------------------------------------------------ the step
    def _fill_hyper_host(self):
        step = self.step_count
        lr = self.lr * (self.lr_schedule(step - 1) if self.lr_schedule is not None else 1.0)
        self.hyper_host[0] = lr
        self.hyper_host[1] = 1.0 - self.betas[0] ** step
        self.hyper_host[2] = math.sqrt(1.0 - self.betas[1] ** step)

    def optimizer_step(self):
        self.step_count += 1
        self._fill_hyper_host()
        self.hyper.copy_(self.hyper_host, non_blocking=True)
        if self.seed_counter is not None:
            self.seed_counter.add_(1)
        self.device_update()

    # ------------------------------------------------------------------------------------------ CUDA graph
    def capture(self, x_example, warmup=2):
        """Captures one full step (forward, backward, [all-reduce], clip + AdamW, weight re-pack) on `x_example`'s
        shape into a CUDA graph.  Per-step scalars (lr, bias corrections, dropout seed counter) live in device memory."""
        assert self.grad_accum == 1, "graph capture covers the single-micro-batch step"
        self.static_x = x_example.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.micro_step(self.static_x)
                self.optimizer_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.step_count += 1
        self._fill_hyper_host()
        g = torch.cuda.CUDAGraph()
        from . import _lib
        l0 = _lib.load().adm_launch_count()
        with torch.cuda.graph(g):
            self.hyper.copy_(self.hyper_host, non_blocking=True)
            self.seed_counter.add_(1)
            self.static_loss = self.micro_step(self.static_x)
            self.device_update()
        self.launches_per_step = _lib.load().adm_launch_count() - l0  # kernels of ours inside one replay
        self.graph = g
        return g

    def replay(self, x=None):
        """One optimizer step through the captured graph; `x` (device or pinned host) is copied into the static input."""
        if x is not None:
            self.static_x.copy_(x, non_blocking=True)
        self.step_count += 1
        self._fill_hyper_host()
        self.graph.replay()
        return self.static_loss

    def device_update(self):
        """Everything after backward that runs on the device (capturable in a CUDA graph)."""
        a = self.arena
        self._allreduce()
        gscale = 1.0 / (self.world * self.grad_accum)
        self.sqnorm.zero_()
        ops.sq_norm(a.grads, self.sqnorm)
        ops.adamw(a.flat, a.grads, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                  max(1, self.step_count), grad_scale=gscale, max_norm=self.max_grad_norm, sqnorm=self.sqnorm,
                  hyper_dev=self.hyper, p_bf16=a.shadow)
        a.zero_grad()
        self.engine.invalidate()

    def micro_step(self, x, t=None, noise=None, last=True, **kw):
        """forward + backward of one micro-batch; returns the (detached) loss tensor.  Gradient reduction is armed only
        on the last micro-batch of an accumulation window."""
        self.arena.rebind_grads()
        if last:
            self._arm()
        else:
            self.engine.grad_hook = None
        if t is None:
            loss, _ = self.dpm(x, **kw)
        else:
            loss, _ = self.dpm.p_losses(x, t, noise=noise, **kw)
        loss.backward()
        return loss.detach()

    def __call__(self, batches):
        """batches: list of `grad_accum` image tensors (this rank's shard).  Returns the mean loss tensor."""
        total = None
        for i, x in enumerate(batches):
            l = self.micro_step(x, last=(i == len(batches) - 1))
            total = l if total is None else total + l
        self.optimizer_step()
        return total / len(batches)
