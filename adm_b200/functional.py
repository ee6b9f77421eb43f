"""Differentiable tensor ops (torch.autograd.Function) over the C-ABI kernels, for module graphs that torch autograd
drives (the conditional UNet, unet/cond_unet.py).  Activations are NHWC bf16 tensors [B, H, W, C]; parameters are the
fp32 masters in the reference layout.  Every forward and backward here is one or a few sm_100a kernels; there is no
CPU / eager fallback (non-CUDA tensors raise in ``ops``).
"""
from __future__ import annotations

import torch

from . import ops
from .ops import BF16, F32


def _nhwc(x):
    if x.dtype != BF16:
        x = x.to(BF16)
    return x if x.is_contiguous() else x.contiguous()


class _Conv2dFn(torch.autograd.Function):
    """k x k convolution (k = 1, 3, 5, 7; stride 1, padding k // 2) as tcgen05 implicit GEMM; ws=True applies weight
    standardisation (K11) on the fly.
    Replaces F.conv2d in cond_unet.py: WeightStandardizedConv2d.forward :349-358, nn.Conv2d of Block / res_conv /
    to_qkv / to_out / Upsample / final_conv."""

    @staticmethod
    def forward(ctx, x, weight, bias, ws, ws_eps):
        x = _nhwc(x)
        cout, cin, k, _ = weight.shape
        assert k in (1, 3, 5, 7) and x.shape[-1] == cin, (weight.shape, x.shape)
        w = weight.detach()
        if ws:
            wpk, stats = ops.ws_pack(w, ws_eps)
        else:
            wpk, stats = ops.pack_conv_weight(w.contiguous()), None
        y = ops.conv_fprop(x, wpk, bias=bias.detach().float() if bias is not None else None)
        ctx.save_for_backward(x, weight, wpk, stats)
        ctx.ws, ctx.has_bias = ws, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, wpk, stats = ctx.saved_tensors
        cout, cin, k, _ = weight.shape
        if cout % 8:  # TMA needs 16-byte pixel strides: view the gradient inside a zero-padded buffer
            dy_full = torch.zeros(*dy.shape[:-1], (cout + 7) // 8 * 8, device=dy.device, dtype=BF16)
            dy_full[..., :cout] = dy
            dy = dy_full[..., :cout]
        else:
            dy = dy_full = _nhwc(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.conv_dgrad(dy, wpk, n_valid=cin)
            if not dx.is_contiguous():
                dx = dx.contiguous()
        if ctx.needs_input_grad[1]:
            dwp = ops.conv_wgrad(dy, x, ntaps=k * k)
            if ctx.ws:
                dw = ops.ws_pack_bwd(dwp, weight.detach(), stats)
            else:
                dw = ops.unpack_conv_wgrad(dwp, cin, 0, k)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(dy_full.shape[-1], device=dy.device, dtype=F32)
            ops.col_sums(dy_full, db)
            db = db[:cout]
        return dx, dw, db, None, None


def conv2d(x, weight, bias=None, ws=False, ws_eps=1e-5):
    return _Conv2dFn.apply(x, weight, bias, ws, ws_eps)


class _GroupNormActFn(torch.autograd.Function):
    """GroupNorm (+ adaptive (1 + scale), shift) (+ SiLU) in one fused kernel each way.  cond_unet.py Block.forward
    :434-443 (norm, scale_shift, act) and the bare nn.GroupNorm uses (:657-660, :741-748)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, scale_shift, act):
        x = _nhwc(x)
        params = scale_shift.detach().float().contiguous() if scale_shift is not None else None
        coef, y = ops.gn_forward(x, None, gamma.detach().float(), beta.detach().float(), groups, eps, params=params,
                                 act=act)
        ctx.save_for_backward(x, gamma, beta, coef, params if params is not None else torch.empty(0))
        ctx.groups, ctx.act, ctx.has_params = groups, act, params is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, coef, params = ctx.saved_tensors
        dy = _nhwc(dy)
        c = x.shape[-1]
        dgamma = torch.zeros(c, device=x.device, dtype=F32)
        dbeta = torch.zeros(c, device=x.device, dtype=F32)
        dparams = torch.empty(x.shape[0], 2 * c, device=x.device, dtype=F32) if ctx.has_params else None
        dx, _ = ops.gn_bwd(dy, x, None, coef, gamma.detach().float(), beta.detach().float(), ctx.groups,
                           params=params if ctx.has_params else None, act=ctx.act, dgamma=dgamma, dbeta=dbeta,
                           dparams=dparams)
        return dx, dgamma, dbeta, None, None, dparams, None


def group_norm_act(x, gamma, beta, groups, eps=1e-5, scale_shift=None, act=True):
    """scale_shift: [B, 2C] = (scale | shift) or None."""
    return _GroupNormActFn.apply(x, gamma, beta, groups, eps, scale_shift, act)


class _LinearAttentionFn(torch.autograd.Function):
    """K12: cond_unet.py LinearAttention.forward :516-531 between to_qkv and to_out."""

    @staticmethod
    def forward(ctx, qkv, heads, scale):
        qkv = _nhwc(qkv)
        out, cx, kstat = ops.linattn_fwd(qkv, heads, scale)
        ctx.save_for_backward(qkv, cx, kstat)
        ctx.heads, ctx.scale = heads, scale
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, cx, kstat = ctx.saved_tensors
        return ops.linattn_bwd(_nhwc(dout), qkv, cx, kstat, ctx.heads, ctx.scale), None, None


def linear_attention(qkv, heads, scale):
    return _LinearAttentionFn.apply(qkv, heads, scale)


class _AttentionFn(torch.autograd.Function):
    """Softmax attention over pixels (cond_unet.py Attention.forward :544-555) on qkv laid out (q | k | v) x head x d."""

    @staticmethod
    def forward(ctx, qkv, heads, scale):
        qkv = _nhwc(qkv)
        a, aux = ops.attention_fwd(qkv, heads, scale)  # aux: log-sum-exp (fused kernel) or probabilities (unfused)
        ctx.save_for_backward(qkv, aux, a)
        ctx.heads, ctx.scale = heads, scale
        return a

    @staticmethod
    def backward(ctx, da):
        qkv, aux, a = ctx.saved_tensors
        return ops.attention_bwd(_nhwc(da), qkv, aux, ctx.heads, ctx.scale, a=a), None, None


def attention(qkv, heads, scale):
    return _AttentionFn.apply(qkv, heads, scale)


class _SpatialAttFn(torch.autograd.Function):
    """out = softsign(softmax(q k^T) att) * h + res with att = h . w_map + b (cond_unet.py SpatialAtt :119-137 and the
    residual `x + decouple(x)` :871-872).  scalars = (b_map, wq, bq, wk, bk)."""

    @staticmethod
    def forward(ctx, h, res, w_map, scalars):
        h, res = _nhwc(h), _nhwc(res)
        wm, sc = w_map.detach().float().contiguous(), scalars.detach().float().contiguous()
        out, att, o = ops.spatial_att_fwd(h, res, wm, sc)
        ctx.save_for_backward(h, wm, sc, att, o)
        return out

    @staticmethod
    def backward(ctx, dy):
        h, wm, sc, att, o = ctx.saved_tensors
        dy = _nhwc(dy)
        dw = torch.zeros_like(wm)
        dsc = torch.zeros_like(sc)
        dh = ops.spatial_att_bwd(dy, h, wm, sc, att, o, dw, dsc)
        return dh, dy, dw, dsc


def spatial_att(h, res, w_map, scalars):
    return _SpatialAttFn.apply(h, res, w_map, scalars)


class _ChanLayerNormFn(torch.autograd.Function):
    """Per-pixel LayerNorm over channels with a gain and no bias (cond_unet.py LayerNorm :360-369): one HBM-bound kernel each
    way on NHWC bf16; the backward re-derives the statistics from x."""

    @staticmethod
    def forward(ctx, x, g, eps):
        x = _nhwc(x)
        gf = g.detach().reshape(-1).float().contiguous()
        ctx.save_for_backward(x, gf)
        ctx.eps, ctx.g_shape = eps, g.shape
        return ops.chan_layernorm_fwd(x, gf, eps)

    @staticmethod
    def backward(ctx, dy):
        x, gf = ctx.saved_tensors
        dx, dg = ops.chan_layernorm_bwd(_nhwc(dy), x, gf, ctx.eps)
        return dx, dg.reshape(ctx.g_shape), None


def channel_layer_norm(x, g, eps=1e-5):
    return _ChanLayerNormFn.apply(x, g, eps)


class _RelationTailFn(torch.autograd.Function):
    """Tail of a relation layer (cond_unet.py:236-251): GroupNorm(x + y) + bilinear(z) with x the trunk features, y the
    concat_conv output and z the out_conv of the pooled tokens.  The sum, the statistics and the normalised value live in
    fp32 registers (the reference's fp32 residual stream) and one bf16 tensor is written; the backward re-derives them."""

    @staticmethod
    def forward(ctx, x, y, z, gamma, beta, groups, eps):
        x, y = _nhwc(x), _nhwc(y)
        zf = z.detach().float().contiguous()
        gf, bf = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        out, stats = ops.rel_gn_fwd(x, y, zf, gf, bf, groups, eps)
        ctx.save_for_backward(x, y, gf, stats)
        ctx.groups, ctx.z_shape, ctx.z_dtype = groups, z.shape, z.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y, gf, stats = ctx.saved_tensors
        dout = _nhwc(dout)
        dpre, dgamma, dbeta = ops.rel_gn_bwd(dout, x, y, gf, stats, ctx.groups)
        dz = ops.bilinear_bwd(dout, ctx.z_shape[1:3]).to(ctx.z_dtype) if ctx.needs_input_grad[2] else None
        return dpre, dpre, dz, dgamma, dbeta, None, None


def relation_tail(x, y, z, gamma, beta, groups, eps=1e-5):
    return _RelationTailFn.apply(x, y, z, gamma, beta, groups, eps)


class _BilinearFn(torch.autograd.Function):
    """F.interpolate(mode='bilinear', align_corners=True) on an NHWC bf16 map (cond_unet.py:184, :248)."""

    @staticmethod
    def forward(ctx, x, size):
        x = _nhwc(x)
        ctx.size_in = x.shape[1:3]
        return ops.bilinear_fwd(x, size)

    @staticmethod
    def backward(ctx, dy):
        return ops.bilinear_bwd(_nhwc(dy), ctx.size_in).to(BF16), None


def bilinear_resize(x, size):
    return _BilinearFn.apply(x, tuple(int(v) for v in size))


class _AvgPoolFn(torch.autograd.Function):
    """F.pad to a multiple of the window + nn.AvgPool2d(window) on an NHWC bf16 map (cond_unet.py:190-200)."""

    @staticmethod
    def forward(ctx, x, window):
        x = _nhwc(x)
        ctx.shape, ctx.window = tuple(x.shape), window
        return ops.avgpool_fwd(x, window)

    @staticmethod
    def backward(ctx, dy):
        return ops.avgpool_bwd(_nhwc(dy), ctx.shape, ctx.window), None


def avg_pool_window(x, window):
    window = (int(window[0]), int(window[1]))
    return x if window == (1, 1) else _AvgPoolFn.apply(x, window)
