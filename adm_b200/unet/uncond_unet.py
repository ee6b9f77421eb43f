"""B200-native mirror of /root/reference/unet/uncond_unet.py (EDMPrecond -> DhariwalUNet, the CIFAR-10 / CelebAHQ
denoiser of DDM).  Same class names, constructor arguments, attributes and ``state_dict`` layout as the reference:

    EDMPrecond(img_resolution, img_channels, label_dim=0, use_fp16=False, sigma_min=0, sigma_max=inf, sigma_data=0.5,
               model_type='DhariwalUNet' | 'SongUNet', precondition=True, **model_kwargs)        uncond_unet.py:588-612
    forward(x, sigma, class_labels=None, force_fp32=False, **{'augment_labels': ...}) -> (D_x, D_y)   :614-635

The modules below only *own parameters* (fp32, reference names and shapes).  All arithmetic of the network runs in
``adm_b200.unet.engine.UNetEngine``: NHWC bf16 activations, tcgen05 implicit-GEMM convolutions, fused GroupNorm kernels,
hand-written backward — entered through one autograd node so that ``loss.backward()`` works like with the reference.
There is no PyTorch / CPU fallback: calling the network on a non-CUDA tensor raises.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


# ------------------------------------------------------------------------------------------------ init (uncond_unet.py:42-47)
def weight_init(shape, mode, fan_in, fan_out):
    if mode == "xavier_uniform":
        return np.sqrt(6 / (fan_in + fan_out)) * (torch.rand(*shape) * 2 - 1)
    if mode == "xavier_normal":
        return np.sqrt(2 / (fan_in + fan_out)) * torch.randn(*shape)
    if mode == "kaiming_uniform":
        return np.sqrt(3 / fan_in) * (torch.rand(*shape) * 2 - 1)
    if mode == "kaiming_normal":
        return np.sqrt(1 / fan_in) * torch.randn(*shape)
    raise ValueError(f'Invalid init mode "{mode}"')


class _EngineOnly(nn.Module):
    """Parameter container; the arithmetic lives in UNetEngine."""

    def forward(self, *a, **k):
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container in adm_b200: run it through EDMPrecond / DhariwalUNet "
            "(the fused sm_100a engine); there is no stand-alone PyTorch forward")


class Linear(_EngineOnly):
    """uncond_unet.py:53-66."""

    def __init__(self, in_features, out_features, bias=True, init_mode="kaiming_normal", init_weight=1, init_bias=0):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        kw = dict(mode=init_mode, fan_in=in_features, fan_out=out_features)
        self.weight = nn.Parameter(weight_init([out_features, in_features], **kw) * init_weight)
        self.bias = nn.Parameter(weight_init([out_features], **kw) * init_bias) if bias else None


class Conv2d(_EngineOnly):
    """uncond_unet.py:72-113 (resample_filter [1,1]: 2x2 box down / nearest up)."""

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
        if (up or down) and list(resample_filter) != [1, 1]:
            raise NotImplementedError("adm_b200 implements the [1,1] resample filter of DhariwalUNet only")


class GroupNorm(_EngineOnly):
    """uncond_unet.py:119-129."""

    def __init__(self, num_channels, num_groups=32, min_channels_per_group=4, eps=1e-5):
        super().__init__()
        self.num_groups = min(num_groups, num_channels // min_channels_per_group)
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))


class UNetBlock(_EngineOnly):
    """uncond_unet.py:157-211."""

    def __init__(self, in_channels, out_channels, emb_channels, up=False, down=False, attention=False, num_heads=None,
                 channels_per_head=64, dropout=0, skip_scale=1, eps=1e-5, resample_filter=[1, 1], resample_proj=False,
                 adaptive_scale=True, init=dict(), init_zero=dict(init_weight=0), init_attn=None):
        super().__init__()
        self.in_channels, self.out_channels, self.emb_channels = in_channels, out_channels, emb_channels
        self.num_heads = 0 if not attention else num_heads if num_heads is not None else out_channels // channels_per_head
        self.dropout, self.skip_scale, self.adaptive_scale = dropout, skip_scale, adaptive_scale
        self.up, self.down = up, down
        if not adaptive_scale or skip_scale != 1 or resample_proj:
            raise NotImplementedError("adm_b200 implements the DhariwalUNet block flavour (adaptive_scale, skip_scale=1)")
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


class PositionalEmbedding(nn.Module):
    """uncond_unet.py:217-230 — host glue on a [B] vector (kept in PyTorch)."""

    def __init__(self, num_channels, max_positions=10000, endpoint=False):
        super().__init__()
        self.num_channels, self.max_positions, self.endpoint = num_channels, max_positions, endpoint

    def forward(self, x):
        freqs = torch.arange(start=0, end=self.num_channels // 2, dtype=torch.float32, device=x.device)
        freqs = freqs / (self.num_channels // 2 - (1 if self.endpoint else 0))
        freqs = (1 / self.max_positions) ** freqs
        x = x.unsqueeze(1) * freqs.to(x.dtype).unsqueeze(0)  # == x.ger(freqs), as one elementwise kernel (ger runs an SGEMM)
        return torch.cat([x.cos(), x.sin()], dim=1)


class SpatialAtt(_EngineOnly):
    """uncond_unet.py:19-37."""

    def __init__(self, in_dim):
        super().__init__()
        self.map = nn.Conv2d(in_dim, 1, 1)
        self.q_conv = nn.Conv2d(1, 1, 1)
        self.k_conv = nn.Conv2d(1, 1, 1)
        self.activation = nn.Softsign()


class DhariwalUNet(nn.Module):
    """uncond_unet.py:450-581 (two decoders sharing the encoder skips)."""

    def __init__(self, img_resolution, in_channels, out_channels, label_dim=0, augment_dim=0, model_channels=192,
                 channel_mult=[1, 2, 3, 4], channel_mult_emb=4, num_blocks=3, attn_resolutions=[32, 16, 8],
                 dropout=0.10, label_dropout=0, out_mul=1, **kwargs):
        super().__init__()
        if label_dim:
            raise NotImplementedError("class-conditional DhariwalUNet (label_dim > 0) is outside the DDM hot path")
        self.label_dropout = label_dropout
        self.img_resolution, self.in_channels, self.out_channels = img_resolution, in_channels, out_channels
        self.model_channels, self.augment_dim = model_channels, augment_dim
        emb_channels = model_channels * channel_mult_emb
        self.emb_channels = emb_channels
        init = dict(init_mode="kaiming_uniform", init_weight=np.sqrt(1 / 3), init_bias=np.sqrt(1 / 3))
        init_zero = dict(init_mode="kaiming_uniform", init_weight=0, init_bias=0)
        init_one = dict(init_mode="kaiming_uniform", init_weight=1, init_bias=0)
        block_kwargs = dict(emb_channels=emb_channels, channels_per_head=64, dropout=dropout, init=init,
                            init_zero=init_zero)
        self.map_noise = PositionalEmbedding(num_channels=model_channels)
        self.map_augment = Linear(augment_dim, model_channels, bias=False, **init_zero) if augment_dim else None
        self.map_layer0 = Linear(model_channels, emb_channels, **init)
        self.map_layer1 = Linear(emb_channels, emb_channels, **init)
        self.map_label = None

        self.enc = nn.ModuleDict()
        cout = in_channels
        for level, mult in enumerate(channel_mult):
            res = img_resolution >> level
            if level == 0:
                cin, cout = cout, model_channels * mult
                self.enc[f"{res}x{res}_conv"] = Conv2d(cin, cout, kernel=3, **init)
            else:
                self.enc[f"{res}x{res}_down"] = UNetBlock(cout, cout, down=True, **block_kwargs)
            for idx in range(num_blocks):
                cin, cout = cout, model_channels * mult
                self.enc[f"{res}x{res}_block{idx}"] = UNetBlock(cin, cout, attention=(res in attn_resolutions),
                                                                **block_kwargs)
        skips = [block.out_channels for block in self.enc.values()]

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
                    dec[f"{res}x{res}_block{idx}"] = UNetBlock(cin, c, attention=(res in attn_resolutions),
                                                               **block_kwargs)
            return dec, c

        self.dec, c1 = make_decoder()
        self.out_norm = GroupNorm(num_channels=c1)
        self.out_conv = Conv2d(c1, out_channels * out_mul, kernel=3, **init_one)
        self.dec2, c2 = make_decoder()
        self.out_norm2 = GroupNorm(num_channels=c2)
        self.out_conv2 = Conv2d(c2, out_channels, kernel=3, **init_one)
        self._engine = None

    @property
    def engine(self):
        if self._engine is None:
            from .engine import UNetEngine
            self._engine = UNetEngine(self)
        return self._engine

    def forward(self, x, noise_labels, class_labels=None, augment_labels=None, **kwargs):
        """(F_x, F_y) for an already pre-scaled input (uncond_unet.py:544-581); inference only — training enters through
        EDMPrecond so that the preconditioning edges stay fused."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("call DhariwalUNet under torch.no_grad(), or train through EDMPrecond")
        return self.engine.forward_raw(x, noise_labels, augment_labels)


class EDMPrecond(nn.Module):
    """uncond_unet.py:588-638."""

    def __init__(self, img_resolution, img_channels, label_dim=0, use_fp16=False, sigma_min=0, sigma_max=float("inf"),
                 sigma_data=0.5, model_type="DhariwalUNet", precondition=True, **model_kwargs):
        super().__init__()
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.self_condition = None
        self.precondition = precondition
        self.channels = img_channels
        self.label_dim, self.use_fp16 = label_dim, use_fp16
        self.sigma_min, self.sigma_max, self.sigma_data = sigma_min, sigma_max, sigma_data
        if model_type not in ("DhariwalUNet", "SongUNet"):
            raise NotImplementedError(f"adm_b200 implements model_type 'DhariwalUNet' and 'SongUNet' (got {model_type!r})")
        model_kwargs.pop("class_name", None)
        self.model_type = model_type
        if model_type == "SongUNet":
            from .song_unet import SongUNet
            self.model = SongUNet(img_resolution=img_resolution, in_channels=img_channels, out_channels=img_channels,
                                  label_dim=label_dim, **model_kwargs)
            return
        if not precondition:
            raise NotImplementedError("adm_b200 implements precondition=True for DhariwalUNet (the configured DDM path)")
        self.model = DhariwalUNet(img_resolution=img_resolution, in_channels=img_channels, out_channels=img_channels,
                                  label_dim=label_dim, **model_kwargs)

    def forward(self, x, sigma, class_labels=None, force_fp32=False, *args, **model_kwargs):
        """x: [B, C, H, W] (any float dtype, NCHW), sigma = t: [B] or 0-dim.  Returns (D_x, D_y) fp32 NCHW."""
        if self.model_type == "SongUNet":
            return self._forward_module_graph(x, sigma, class_labels, **model_kwargs)
        from .engine import unet_apply
        return unet_apply(self.model.engine, x, sigma, model_kwargs.get("augment_labels"))

    def _forward_module_graph(self, x, sigma, class_labels=None, **model_kwargs):
        """uncond_unet.py:614-635 around a network that is a torch-autograd module graph (SongUNet): the per-sample
        preconditioning scalars and the two output mixes are a handful of elementwise ops on 3-channel images."""
        x = x.to(torch.float32)
        sigma = sigma.to(torch.float32).reshape(-1, 1, 1, 1)
        if self.label_dim == 0:
            class_labels = None
        elif class_labels is None:
            class_labels = torch.zeros([1, self.label_dim], device=x.device)
        else:
            class_labels = class_labels.to(torch.float32).reshape(-1, self.label_dim)
        den = sigma ** 2 - sigma + 1
        c_skip1, c_skip2 = (sigma - 1) / den, sigma.sqrt() / den
        c_out1, c_out2 = torch.sqrt(sigma / den), (1 - sigma) / den.sqrt()
        c_in = 1 / torch.sqrt((1 - sigma) ** 2 + sigma)
        c_noise = sigma.log()
        f_x, f_y = self.model(c_in * x, c_noise.flatten().expand(x.shape[0]), class_labels=class_labels, **model_kwargs)
        if not self.precondition:
            return f_x, f_y
        return c_skip1 * x + c_out1 * f_x, c_skip2 * x + c_out2 * f_y

    def round_sigma(self, sigma):
        return torch.as_tensor(sigma)


def create_model(cfg):
    """uncond_unet.py:640-656."""
    return EDMPrecond(img_resolution=cfg.img_resolution, img_channels=cfg.img_channels, sigma_data=cfg.sigma_data,
                      model_type=cfg.model_type, model_channels=cfg.model_channels, channel_mult=cfg.channel_mult,
                      channel_mult_emb=cfg.channel_mult_emb, num_blocks=cfg.num_blocks,
                      attn_resolutions=cfg.attn_resolutions, dropout=cfg.dropout, label_dropout=cfg.label_dropout,
                      augment_dim=cfg.augment_dim, out_mul=cfg.get("out_mul", 1))
