"""B200-native mirror of ``SongUNet`` (/root/reference/unet/uncond_unet.py:253-441): the DDPM++ / NCSN++ two-decoder UNet
that ``EDMPrecond(model_type='SongUNet')`` instantiates (:612).  SURVEY section 8 row f-4.

Same constructor arguments, attribute names and ``state_dict`` layout as the reference (``map_noise.freqs`` buffer of the
Fourier embedding, ``enc / dec / dec2`` ModuleDicts with the ``{res}x{res}_{conv,down,block{i},aux_*,in0,in1,up}`` keys,
``resample_filter`` buffers of the up / down convs).  Unlike ``DhariwalUNet`` — whose arithmetic lives in the hand-written
``UNetEngine`` — this network is an ordinary module graph under torch autograd, like ``cond_unet.Unet``: activations are
NHWC bf16 and every hot op is an ``autograd.Function`` over the C-ABI kernels (``adm_b200/functional.py``):

* 3x3 / 1x1 convolutions: tcgen05 implicit GEMM (fprop, dgrad, wgrad), bias in the epilogue;
* GroupNorm (+ SiLU) (+ the adaptive (1 + scale), shift of ``adaptive_scale=True``): the fused GroupNorm kernels;
* attention (``num_heads=1``: one head as wide as the block, :204-208): the attention kernels on (q | k | v)-ordered
  projections (the reference interleaves (head, d, {q,k,v}) rows, :205 — permuted in the weights, not in the activations);
* 2x2 box down / nearest up of ``resample_filter=[1,1]`` (DDPM++): the resample kernel, both directions.

What stays torch glue (small tensors, or shapes the implicit-GEMM tiling does not take): the embedding MLP and the per-block
``affine`` Linear on ``[B, emb]``; the per-sample bias add / ``skip_scale`` multiplies / dropout (elementwise on bf16); the
depthwise FIR resampling of ``resample_filter=[1,3,3,1]`` (NCSN++) and the fused-resample conv with padding 2 that comes
with it (``F.conv2d`` on the channels-last view).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as AF
from .. import ops
from .uncond_unet import PositionalEmbedding, SpatialAtt, weight_init

BF16 = torch.bfloat16


def silu(x):
    return F.silu(x)


def _conv_ok(h, w):
    """Image sizes the implicit-GEMM kernels tile (128-pixel output boxes, 64-pixel K boxes of the weight gradient)."""
    def box(pixels):
        if w >= pixels:
            return w % pixels == 0
        if pixels % w:
            return False
        rows = pixels // w
        return h % rows == 0 if h >= rows else rows % h == 0
    return box(128) and box(64)


class _ResampleFn(torch.autograd.Function):
    """2x2 box down (mode 1) / nearest x2 up (mode 2) on NHWC bf16: the [1,1] resample filter of Conv2d.forward (:105-108)."""

    @staticmethod
    def forward(ctx, x, mode):
        ctx.mode = mode
        return ops.resample(x.contiguous(), mode)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        if ctx.mode == 1:  # d(avg 2x2) = nearest-up / 4
            return ops.resample(dy, 2) * 0.25, None
        return ops.resample(dy, 1) * 4.0, None  # d(nearest up) = sum over the 2x2 patch = 4 * avg


class FourierEmbedding(nn.Module):
    """uncond_unet.py:236-244."""

    def __init__(self, num_channels, scale=16):
        super().__init__()
        self.register_buffer("freqs", torch.randn(num_channels // 2) * scale)

    def forward(self, x):
        x = x.unsqueeze(1) * (2 * np.pi * self.freqs).to(x.dtype).unsqueeze(0)
        return torch.cat([x.cos(), x.sin()], dim=1)


class Linear(nn.Module):
    """uncond_unet.py:53-66 on [B, features] (host-side glue: the embedding MLP and the per-block affine)."""

    def __init__(self, in_features, out_features, bias=True, init_mode="kaiming_normal", init_weight=1, init_bias=0):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        kw = dict(mode=init_mode, fan_in=in_features, fan_out=out_features)
        self.weight = nn.Parameter(weight_init([out_features, in_features], **kw) * init_weight)
        self.bias = nn.Parameter(weight_init([out_features], **kw) * init_bias) if bias else None

    def forward(self, x):
        return F.linear(x, self.weight.to(x.dtype), self.bias.to(x.dtype) if self.bias is not None else None)


class Conv2d(nn.Module):
    """uncond_unet.py:72-113 on NHWC bf16: optional up / down resampling around a 3x3 / 1x1 conv (or alone, kernel = 0)."""

    def __init__(self, in_channels, out_channels, kernel, bias=True, up=False, down=False, resample_filter=[1, 1],
                 fused_resample=False, init_mode="kaiming_normal", init_weight=1, init_bias=0):
        assert not (up and down)
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.up, self.down, self.fused_resample, self.kernel = up, down, fused_resample, kernel
        kw = dict(mode=init_mode, fan_in=in_channels * kernel * kernel, fan_out=out_channels * kernel * kernel)
        self.weight = nn.Parameter(weight_init([out_channels, in_channels, kernel, kernel], **kw) * init_weight) \
            if kernel else None
        self.bias = nn.Parameter(weight_init([out_channels], **kw) * init_bias) if kernel and bias else None
        f = torch.as_tensor(resample_filter, dtype=torch.float32)
        f = f.ger(f).unsqueeze(0).unsqueeze(1) / f.sum().square()
        self.register_buffer("resample_filter", f if up or down else None)
        self._box = list(resample_filter) == [1, 1]

    # ---- resampling: the kernel for the [1,1] box filter, depthwise torch convs for longer FIR filters
    def _up(self, x, pad):
        if self._box and pad == 0 and x.shape[-1] % 8 == 0:
            return _ResampleFn.apply(x, 2)
        c = x.shape[-1]
        f = self.resample_filter.to(x.dtype).mul(4).tile([c, 1, 1, 1])
        y = F.conv_transpose2d(x.permute(0, 3, 1, 2), f, groups=c, stride=2, padding=pad)
        return y.permute(0, 2, 3, 1).contiguous()

    def _down(self, x, pad):
        if self._box and pad == 0 and x.shape[-1] % 8 == 0 and x.shape[1] % 2 == 0 and x.shape[2] % 2 == 0:
            return _ResampleFn.apply(x, 1)
        c = x.shape[-1]
        f = self.resample_filter.to(x.dtype).tile([c, 1, 1, 1])
        y = F.conv2d(x.permute(0, 3, 1, 2), f, groups=c, stride=2, padding=pad)
        return y.permute(0, 2, 3, 1).contiguous()

    def _conv(self, x, pad, bias):
        w = self.weight
        k = w.shape[-1]
        if pad == k // 2 and _conv_ok(x.shape[1], x.shape[2]):
            cin = w.shape[1]
            if cin % 8:  # RGB / auxiliary-image inputs: zero-pad activation and weight to 8 channels
                p = (-cin) % 8
                x, w = F.pad(x, (0, p)), F.pad(w, (0, 0, 0, 0, 0, p))
            return AF.conv2d(x, w, bias)
        # paddings / image sizes the implicit-GEMM tiling does not take (the fused-resample conv of NCSN++ pads by 2)
        y = F.conv2d(x.permute(0, 3, 1, 2), w.to(x.dtype), bias.to(x.dtype) if bias is not None else None, padding=pad)
        return y.permute(0, 2, 3, 1).contiguous()

    def forward(self, x):
        w, b = self.weight, self.bias
        f = self.resample_filter
        w_pad = w.shape[-1] // 2 if w is not None else 0
        f_pad = (f.shape[-1] - 1) // 2 if f is not None else 0
        if self.fused_resample and self.up and w is not None:
            x = self._up(x, max(f_pad - w_pad, 0))
            return self._conv(x, max(w_pad - f_pad, 0), b)
        if self.fused_resample and self.down and w is not None:
            x = self._conv(x, w_pad + f_pad, None)
            x = self._down(x, 0)
            return x + b.to(x.dtype) if b is not None else x
        if self.up:
            x = self._up(x, f_pad)
        if self.down:
            x = self._down(x, f_pad)
        if w is not None:
            x = self._conv(x, w_pad, b)
        return x


class GroupNorm(nn.Module):
    """uncond_unet.py:119-129; forward = the fused GroupNorm (+ SiLU) (+ adaptive scale / shift) kernel on NHWC bf16."""

    def __init__(self, num_channels, num_groups=32, min_channels_per_group=4, eps=1e-5):
        super().__init__()
        self.num_groups = min(num_groups, num_channels // min_channels_per_group)
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))

    def forward(self, x, act=False, scale_shift=None):
        return AF.group_norm_act(x, self.weight, self.bias, self.num_groups, self.eps, scale_shift=scale_shift, act=act)


class UNetBlock(nn.Module):
    """uncond_unet.py:157-211 in every flavour the reference builds (adaptive_scale on / off, skip_scale, resample_proj)."""

    def __init__(self, in_channels, out_channels, emb_channels, up=False, down=False, attention=False, num_heads=None,
                 channels_per_head=64, dropout=0, skip_scale=1, eps=1e-5, resample_filter=[1, 1], resample_proj=False,
                 adaptive_scale=True, init=dict(), init_zero=dict(init_weight=0), init_attn=None):
        super().__init__()
        self.in_channels, self.out_channels, self.emb_channels = in_channels, out_channels, emb_channels
        self.num_heads = 0 if not attention else num_heads if num_heads is not None else out_channels // channels_per_head
        self.dropout, self.skip_scale, self.adaptive_scale = dropout, skip_scale, adaptive_scale
        self.norm0 = GroupNorm(num_channels=in_channels, eps=eps)
        self.conv0 = Conv2d(in_channels, out_channels, kernel=3, up=up, down=down, resample_filter=resample_filter, **init)
        self.affine = Linear(emb_channels, out_channels * (2 if adaptive_scale else 1), **init)
        self.norm1 = GroupNorm(num_channels=out_channels, eps=eps)
        self.conv1 = Conv2d(out_channels, out_channels, kernel=3, **init_zero)
        self.skip = None
        if out_channels != in_channels or up or down:
            kernel = 1 if resample_proj or out_channels != in_channels else 0
            self.skip = Conv2d(in_channels, out_channels, kernel=kernel, up=up, down=down,
                               resample_filter=resample_filter, **init)
        if self.num_heads:
            self.norm2 = GroupNorm(num_channels=out_channels, eps=eps)
            self.qkv = Conv2d(out_channels, out_channels * 3, kernel=1, **(init_attn if init_attn is not None else init))
            self.proj = Conv2d(out_channels, out_channels, kernel=1, **init_zero)

    def _attention(self, x):
        """:204-208 with the projection rows re-ordered to (q | k | v) x head x d and every head zero-padded to a multiple
        of 64 channels inside the weights (q.k and p.v are unchanged by zero columns)."""
        c, heads = self.out_channels, self.num_heads
        d = c // heads
        dp = (d + 63) // 64 * 64
        wq = self.qkv.weight.reshape(heads, d, 3, c).permute(2, 0, 1, 3)          # [3, heads, d, C]
        bq = self.qkv.bias.reshape(heads, d, 3).permute(2, 0, 1)
        wo = self.proj.weight.reshape(c, heads, d)
        if dp != d:
            wq, bq, wo = F.pad(wq, (0, 0, 0, dp - d)), F.pad(bq, (0, dp - d)), F.pad(wo, (0, dp - d))
        qkv = AF.conv2d(self.norm2(x), wq.reshape(3 * heads * dp, c, 1, 1).contiguous(), bq.reshape(-1).contiguous())
        a = AF.attention(qkv, heads, d ** -0.5)
        return AF.conv2d(a, wo.reshape(c, heads * dp, 1, 1).contiguous(), self.proj.bias)

    def forward(self, x, emb):
        orig = x
        x = self.conv0(self.norm0(x, act=True))
        params = self.affine(emb)  # [B, Cout] or [B, 2 Cout] = (scale | shift), fp32
        if self.adaptive_scale:
            x = self.norm1(x, act=True, scale_shift=params)
        else:  # the per-sample bias is added in fp32 (it is an fp32 quantity in the reference) and rounded once
            x = self.norm1((x.float() + params[:, None, None, :]).to(BF16), act=True)
        x = self.conv1(F.dropout(x, p=self.dropout, training=self.training))
        # (h + skip) * skip_scale summed in fp32: one bf16 rounding of the block output instead of three
        x = ((x.float() + (self.skip(orig) if self.skip is not None else orig).float()) * self.skip_scale).to(BF16)
        if self.num_heads:
            x = ((self._attention(x).float() + x.float()) * self.skip_scale).to(BF16)
        return x


class SongUNet(nn.Module):
    """uncond_unet.py:253-441."""

    def __init__(self, img_resolution, in_channels, out_channels, label_dim=0, augment_dim=0, model_channels=128,
                 channel_mult=[1, 2, 2, 2], channel_mult_emb=4, num_blocks=4, attn_resolutions=[16], dropout=0.10,
                 label_dropout=0, embedding_type="fourier", channel_mult_noise=2, encoder_type="residual",
                 decoder_type="standard", resample_filter=[1, 3, 3, 1], **unused):
        assert embedding_type in ["fourier", "positional"]
        assert encoder_type in ["standard", "skip", "residual"]
        assert decoder_type in ["standard", "skip"]
        super().__init__()
        self.label_dropout = label_dropout
        self.img_resolution, self.in_channels, self.out_channels = img_resolution, in_channels, out_channels
        emb_channels = model_channels * channel_mult_emb
        noise_channels = model_channels * channel_mult_noise
        init = dict(init_mode="xavier_uniform")
        init_zero = dict(init_mode="xavier_uniform", init_weight=1e-5)
        init_attn = dict(init_mode="xavier_uniform", init_weight=np.sqrt(0.2))
        block_kwargs = dict(emb_channels=emb_channels, num_heads=1, dropout=dropout, skip_scale=np.sqrt(0.5), eps=1e-6,
                            resample_filter=resample_filter, resample_proj=True, adaptive_scale=False, init=init,
                            init_zero=init_zero, init_attn=init_attn)
        self.map_noise = PositionalEmbedding(num_channels=noise_channels, endpoint=True) \
            if embedding_type == "positional" else FourierEmbedding(num_channels=noise_channels)
        self.map_label = Linear(label_dim, noise_channels, **init) if label_dim else None
        self.map_augment = Linear(augment_dim, noise_channels, bias=False, **init) if augment_dim else None
        self.map_layer0 = Linear(noise_channels, emb_channels, **init)
        self.map_layer1 = Linear(emb_channels, emb_channels, **init)

        self.enc = nn.ModuleDict()
        cout = in_channels
        caux = in_channels
        for level, mult in enumerate(channel_mult):
            res = img_resolution >> level
            if level == 0:
                cin, cout = cout, model_channels
                self.enc[f"{res}x{res}_conv"] = Conv2d(cin, cout, kernel=3, **init)
            else:
                self.enc[f"{res}x{res}_down"] = UNetBlock(cout, cout, down=True, **block_kwargs)
                if encoder_type == "skip":
                    self.enc[f"{res}x{res}_aux_down"] = Conv2d(caux, caux, kernel=0, down=True,
                                                               resample_filter=resample_filter)
                    self.enc[f"{res}x{res}_aux_skip"] = Conv2d(caux, cout, kernel=1, **init)
                if encoder_type == "residual":
                    self.enc[f"{res}x{res}_aux_residual"] = Conv2d(caux, cout, kernel=3, down=True,
                                                                   resample_filter=resample_filter, fused_resample=True,
                                                                   **init)
                    caux = cout
            for idx in range(num_blocks):
                cin, cout = cout, model_channels * mult
                self.enc[f"{res}x{res}_block{idx}"] = UNetBlock(cin, cout, attention=(res in attn_resolutions),
                                                                **block_kwargs)
        skips = [block.out_channels for name, block in self.enc.items() if "aux" not in name]

        self.decouple1 = nn.Sequential(nn.Conv2d(cout, cout, 3, 1, 1), SpatialAtt(cout))
        self.decouple2 = nn.Sequential(nn.Conv2d(cout, cout, 3, 1, 1), SpatialAtt(cout))

        def make_decoder():
            dec = nn.ModuleDict()
            c = cout
            sk = list(skips)
            for level, mult in reversed(list(enumerate(channel_mult))):
                res = img_resolution >> level
                if level == len(channel_mult) - 1:
                    dec[f"{res}x{res}_in0"] = UNetBlock(c, c, attention=True, **block_kwargs)
                    dec[f"{res}x{res}_in1"] = UNetBlock(c, c, **block_kwargs)
                else:
                    dec[f"{res}x{res}_up"] = UNetBlock(c, c, up=True, **block_kwargs)
                for idx in range(num_blocks + 1):
                    cin = c + sk.pop()
                    c = model_channels * mult
                    attn = idx == num_blocks and res in attn_resolutions
                    dec[f"{res}x{res}_block{idx}"] = UNetBlock(cin, c, attention=attn, **block_kwargs)
                if decoder_type == "skip" or level == 0:
                    if decoder_type == "skip" and level < len(channel_mult) - 1:
                        dec[f"{res}x{res}_aux_up"] = Conv2d(out_channels, out_channels, kernel=0, up=True,
                                                            resample_filter=resample_filter)
                    dec[f"{res}x{res}_aux_norm"] = GroupNorm(num_channels=c, eps=1e-6)
                    dec[f"{res}x{res}_aux_conv"] = Conv2d(c, out_channels, kernel=3, **init_zero)
            return dec

        self.dec = make_decoder()
        self.dec2 = make_decoder()

    # ------------------------------------------------------------------------------------------ pieces of the forward
    @staticmethod
    def _decouple(seq, x):
        """decouple{1,2} (:325-332, applied at :391-392 as ``decouple(x) + x``): 3x3 conv + SpatialAtt, residual fused."""
        conv, sa = seq[0], seq[1]
        h = AF.conv2d(x, conv.weight, conv.bias)
        scalars = torch.cat([sa.map.bias, sa.q_conv.weight.reshape(-1), sa.q_conv.bias, sa.k_conv.weight.reshape(-1),
                             sa.k_conv.bias])
        return AF.spatial_att(h, x, sa.map.weight.reshape(-1), scalars)

    @staticmethod
    def _decoder(dec, x, skips, emb):
        aux = tmp = None
        for name, block in dec.items():
            if "aux_up" in name:
                aux = block(aux)
            elif "aux_norm" in name:
                tmp = block(x, act=True)   # silu(aux_norm(x)) fused (:402-404)
            elif "aux_conv" in name:
                tmp = block(tmp)
                aux = tmp if aux is None else tmp + aux
            else:
                if x.shape[-1] != block.in_channels:
                    x = torch.cat([x, skips.pop()], dim=-1)
                x = block(x, emb)
        return aux

    def forward(self, x, noise_labels, class_labels=None, augment_labels=None, **kwargs):
        """x NCHW float (already scaled by c_in), noise_labels [B].  Returns (F_x, F_y) NCHW fp32 (:359-425)."""
        if not x.is_cuda:
            raise RuntimeError("adm_b200: the UNet runs on CUDA (sm_100a) only; there is no CPU fallback")
        emb = self.map_noise(noise_labels.to(torch.float32).reshape(-1))
        emb = emb.reshape(emb.shape[0], 2, -1).flip(1).reshape(*emb.shape)  # swap sin / cos
        if self.map_label is not None:
            tmp = class_labels
            if self.training and self.label_dropout:
                tmp = tmp * (torch.rand([x.shape[0], 1], device=x.device) >= self.label_dropout).to(tmp.dtype)
            emb = emb + self.map_label(tmp * np.sqrt(self.map_label.in_features))
        if self.map_augment is not None and augment_labels is not None:
            emb = emb + self.map_augment(augment_labels.to(torch.float32))
        emb = silu(self.map_layer0(emb))
        emb = silu(self.map_layer1(emb))

        x = x.permute(0, 2, 3, 1).to(BF16).contiguous()  # NHWC bf16
        # The fork keeps a second skip list for decoder 2 (:371, :388) that the aux branches do NOT update
        # (`x = skips[-1] = ...`, :381-384, touches the first list only): decoder 2 sees the pre-aux tensors.  Kept as is.
        skips, skips2 = [], []
        aux = x
        for name, block in self.enc.items():
            if "aux_down" in name:
                aux = block(aux)
            elif "aux_skip" in name:
                x = skips[-1] = (x.float() + block(aux).float()).to(BF16)
            elif "aux_residual" in name:
                x = skips[-1] = aux = ((x.float() + block(aux).float()) * float(1 / np.sqrt(2))).to(BF16)
            else:
                x = block(x, emb) if isinstance(block, UNetBlock) else block(x)
                skips.append(x)
                skips2.append(x)
        x1 = self._decouple(self.decouple1, x)
        x2 = self._decouple(self.decouple2, x)
        f1 = self._decoder(self.dec, x1, skips, emb)
        f2 = self._decoder(self.dec2, x2, skips2, emb)
        return f1.permute(0, 3, 1, 2).float(), f2.permute(0, 3, 1, 2).float()
