"""B200-native mirror of the frozen first stage of the latent configs: ``AutoencoderKL`` of
/root/reference/ddm/encoder_decoder.py (:894-969; Encoder :386-471, Decoder :474-586, ResnetBlock :100-159, AttnBlock
:168-220, Downsample / Upsample :60-97, DiagonalGaussianDistribution :854-892).  SURVEY §8 row f-1.

Same constructor (``AutoencoderKL(ddconfig, lossconfig, embed_dim, ckpt_path=None, ...)``), same ``encode`` / ``decode`` /
``down_ratio`` surface and the same ``encoder.* / decoder.* / quant_conv.* / post_quant_conv.*`` state_dict keys; the
training-only ``loss.*`` sub-module (LPIPS + PatchGAN, needs a VGG download) is not built — checkpoints load with
``strict=False`` exactly as the reference's own ``init_from_ckpt`` does (:935).

It is used frozen and without gradients (ddm_const_2.py:438-442, 494-503), so only the forward exists here: activations
are NHWC bf16, every GroupNorm(32, eps 1e-6)+swish is the fused GroupNorm kernel, every 3x3 / 1x1 stride-1 conv is the
tcgen05 implicit GEMM, nearest x2 is the resample kernel.  The stride-2 3x3 convs of the encoder (:78-97: pad right /
bottom by one, stride 2, no other padding) run through the same implicit-GEMM kernel: such a conv equals the pad-1
stride-1 conv sampled at the odd pixels, o[i, j] = full[2i + 1, 2j + 1].  The single-head mid attention (:168-220,
d = 512, H*W up to 4096 tokens) is q.k^T -> softmax kernel -> p.v on the tcgen05 batched GEMMs with the three 1x1
projections fused into one conv; beyond 4096 tokens the N x N score matrix is not worth materialising and
``scaled_dot_product_attention`` takes over.  Only the 1x1 (quant) convs on 3-6 channels stay ATen.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as AF
from .. import ops


def _nhwc(x):
    return x.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2)


class Normalize(nn.GroupNorm):
    """encoder_decoder.py:56-57; forward = fused GroupNorm (+ swish) on NHWC bf16."""

    def __init__(self, in_channels, num_groups=32):
        super().__init__(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)

    def forward(self, x, act=True):
        return AF.group_norm_act(x, self.weight, self.bias, self.num_groups, self.eps, act=act)


class _Conv(nn.Conv2d):
    """3x3 (pad 1) / 1x1 stride-1 conv on NHWC bf16 through the implicit-GEMM engine; input channels that are not a
    multiple of 8 (RGB / latent inputs) are zero-padded on both the activation and the weight."""

    def forward(self, x):
        cin = self.weight.shape[1]
        if cin % 8:
            pad = (-cin) % 8
            x = F.pad(x, (0, pad))
            w = F.pad(self.weight, (0, 0, 0, 0, 0, pad))
            return AF.conv2d(x, w, self.bias)
        return AF.conv2d(x, self.weight, self.bias)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = _Conv(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        x = ops.resample(x, 2)  # nearest x2
        return self.conv(x) if self.with_conv else x


class Downsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def forward(self, x):
        if not self.with_conv:
            return ops.resample(x, 1)  # 2x2 average
        if x.shape[1] % 2 or x.shape[2] % 2:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = self.conv(F.pad(_nchw(x), (0, 1, 0, 1), mode="constant", value=0))
            return _nhwc(y)
        # pad (0, 1, 0, 1) + stride 2 + no padding reads rows 2i .. 2i+2: the window of the pad-1 stride-1 conv centred on
        # the odd pixel (2i+1, 2j+1); its zero row / column beyond the image is the reference's explicit padding
        full = AF.conv2d(x, self.conv.weight, self.conv.bias)
        return full[:, 1::2, 1::2, :].contiguous()


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels, self.use_conv_shortcut = in_channels, out_channels, conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = _Conv(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if temb_channels > 0:
            self.temb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = _Conv(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = _Conv(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
            else:
                self.nin_shortcut = _Conv(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, x, temb=None):
        h = self.conv1(self.norm1(x))
        if temb is not None:
            h = h + self.temb_proj(F.silu(temb))[:, None, None, :].to(h.dtype)
        h = self.conv2(self.dropout(self.norm2(h)))
        if self.in_channels != self.out_channels:
            x = self.conv_shortcut(x) if self.use_conv_shortcut else self.nin_shortcut(x)
        return x + h


class AttnBlock(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = _Conv(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = _Conv(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = _Conv(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = _Conv(in_channels, in_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, x):
        b, h, w, c = x.shape
        hn = self.norm(x, act=False)
        if h * w > 4096 or c % 64:
            q, k, v = (f(hn).reshape(b, 1, h * w, c) for f in (self.q, self.k, self.v))
            a = F.scaled_dot_product_attention(q, k, v, scale=int(c) ** (-0.5))  # softmax(q k^T / sqrt(c)) v
            return x + self.proj_out(a.reshape(b, h, w, c))
        # one conv for the three projections, output laid out (q | k | v) as the attention kernels read it
        wqkv = torch.cat([self.q.weight, self.k.weight, self.v.weight], dim=0)
        bqkv = torch.cat([self.q.bias, self.k.bias, self.v.bias], dim=0)
        qkv = AF.conv2d(hn, wqkv, bqkv)
        a, _ = ops.attention_fwd(qkv.contiguous(), 1, scale=int(c) ** (-0.5), need_p=False, fused=False)
        return x + self.proj_out(a)


def make_attn(in_channels, attn_type="vanilla"):
    assert attn_type in ["vanilla", "none"], f"adm_b200: attn_type {attn_type} is not built (the configs use 'vanilla')"
    return AttnBlock(in_channels) if attn_type == "vanilla" else nn.Identity()


class Encoder(nn.Module):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, use_linear_attn=False,
                 attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.conv_in = _Conv(in_channels, ch, kernel_size=3, stride=1, padding=1)
        curr_res = tuple(resolution) if not isinstance(resolution, int) else (resolution, resolution)
        in_ch_mult = (1,) + tuple(ch_mult)
        self.in_ch_mult = in_ch_mult
        attn_res = [tuple(r) if not isinstance(r, int) else (r, r) for r in attn_resolutions]
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_res:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res = (curr_res[0] // 2, curr_res[1] // 2)
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.conv_out = _Conv(block_in, 2 * z_channels if double_z else z_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        """x NCHW float -> moments NCHW fp32."""
        h = self.conv_in(_nhwc(x))
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block](h)
                if len(self.down[i_level].attn) > 0:
                    h = self.down[i_level].attn[i_block](h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample(h)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        return _nchw(self.conv_out(self.norm_out(h))).float()


class Decoder(nn.Module):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.give_pre_end, self.tanh_out = give_pre_end, tanh_out
        res = tuple(resolution) if not isinstance(resolution, int) else (resolution, resolution)
        attn_res = [tuple(r) if not isinstance(r, int) else (r, r) for r in attn_resolutions]
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = (res[0] // 2 ** (self.num_resolutions - 1), res[1] // 2 ** (self.num_resolutions - 1))
        self.z_shape = (1, z_channels, curr_res[0], curr_res[1])
        self.conv_in = _Conv(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_res:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = (curr_res[0] * 2, curr_res[1] * 2)
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = _Conv(block_in, out_ch, kernel_size=3, stride=1, padding=1)

    def forward(self, z):
        """z NCHW float -> image NCHW fp32."""
        self.last_z_shape = z.shape
        h = self.conv_in(_nhwc(z))
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block](h)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block](h)
            if i_level != 0:
                h = self.up[i_level].upsample(h)
        if self.give_pre_end:
            return _nchw(h).float()
        h = _nchw(self.conv_out(self.norm_out(h))).float()
        return torch.tanh(h) if self.tanh_out else h


class DiagonalGaussianDistribution(object):
    """encoder_decoder.py:854-892."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self):
        return self.mean + self.std * torch.randn(self.mean.shape, device=self.parameters.device)

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.])
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0
                               - self.logvar + other.logvar, dim=[1, 2, 3])

    def nll(self, sample, dims=[1, 2, 3]):
        if self.deterministic:
            return torch.Tensor([0.])
        return 0.5 * torch.sum(np.log(2.0 * np.pi) + self.logvar + torch.pow(sample - self.mean, 2) / self.var, dim=dims)

    def mode(self):
        return self.mean


class AutoencoderKL(nn.Module):
    def __init__(self, ddconfig, lossconfig=None, embed_dim=3, ckpt_path=None, ignore_keys=[], image_key="image",
                 colorize_nlabels=None, monitor=None, **kwargs):
        super().__init__()
        ddconfig = dict(ddconfig)
        self.image_key = image_key
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.down_ratio = 2 ** (len(ddconfig["ch_mult"]) - 1)
        assert ddconfig["double_z"]
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim = embed_dim
        if colorize_nlabels is not None:
            assert type(colorize_nlabels) == int
            self.register_buffer("colorize", torch.randn(3, colorize_nlabels, 1, 1))
        if monitor is not None:
            self.monitor = monitor
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)

    def init_from_ckpt(self, path, ignore_keys=list(), use_ema=True):
        sd = torch.load(path, map_location="cpu")
        if "ema" in sd and use_ema:
            sd = {k[10:]: v for k, v in sd["ema"].items() if k.startswith("ema_model.")}
        elif "model" in sd:
            sd = sd["model"]
        elif "state_dict" in sd:
            sd = sd["state_dict"]
        else:
            raise ValueError("")
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        msg = self.load_state_dict(sd, strict=False)  # the reference's loss.* (LPIPS / discriminator) keys are unused
        print(f"Restored from {path}")
        print("==>Load AutoEncoder Info: ", msg)

    @torch.no_grad()
    def encode(self, x):
        if not x.is_cuda:
            raise RuntimeError("adm_b200.AutoencoderKL runs on CUDA (sm_100a) only; there is no CPU fallback")
        return DiagonalGaussianDistribution(self.quant_conv(self.encoder(x)))

    @torch.no_grad()
    def decode(self, z):
        if not z.is_cuda:
            raise RuntimeError("adm_b200.AutoencoderKL runs on CUDA (sm_100a) only; there is no CPU fallback")
        return self.decoder(self.post_quant_conv(z.float()))

    def forward(self, input, sample_posterior=True):
        posterior = self.encode(input)
        z = posterior.sample() if sample_posterior else posterior.mode()
        return self.decode(z), posterior
