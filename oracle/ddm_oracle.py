"""ORACLE — test infrastructure only.  Never import this from the product (adm_b200/), only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

A CPU restatement, in plain functional PyTorch (fp32 / fp64), of the reference's DDM-const hot path:

* EDMPrecond + DhariwalUNet forward  — /root/reference/unet/uncond_unet.py:588-635, :450-581, :157-211, :19-37,
  :53-66, :72-113, :119-129, :217-230.  Parameters are taken from a flat ``state_dict`` with the reference's key
  grammar (``model.enc.32x32_block0.conv0.weight`` ...), so the restatement shares no module code with either the
  reference or the product.
* EDMPrecond + SongUNet forward      — /root/reference/unet/uncond_unet.py:253-441 (row f-4), every Conv2d / UNetBlock
  branch it uses (:91-113, :189-211); pinned by tests/golden/make_golden_song.py (song_unet.pt).
* DDM-const math                     — /root/reference/ddm/ddm_const.py:274-287 (t, q_sample), :290-303 (x0, x_{t-s}),
  :305-364 (loss; ``loss_main_func`` = MSE_Loss, ddm/loss.py:300-312; LPIPS term = 0 because the reference cannot build
  LPIPS offline), :425-476 (deterministic sampler), :381-422 (stochastic sampler).

Parity pin: tests/golden/*.json were produced by tests/golden/make_golden.py, which imports the *unmodified* reference
``unet.uncond_unet.EDMPrecond`` (and ``ddm.ddm_const_2.DDPM`` for the plumbing cross-check) from /root/reference in the
build container and records its outputs on seeded inputs; tests/test_oracle.py checks this file against them.
The DDM-const math is pinned DIRECTLY: tests/golden/make_golden_ddm.py imports the reference's own
``ddm/ddm_const.py`` with its absent import-only packages (ldm, cldm, pytorch_lightning, ...) stubbed and records what
the unmodified ``DDPM.q_sample / pred_x0_from_xt / pred_xtms_from_xt / p_losses / sample_fn_d / sample_fn_s`` (and the
sibling's ``LatentDiffusion.p_losses``) return on seeded inputs with a closed-form denoiser (tests/golden/ddm_math.pt);
tests/test_oracle.py checks every function below against it.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------- config / shapes
def unet_config(img_resolution=32, img_channels=3, model_channels=192, channel_mult=(1, 2, 2, 2), channel_mult_emb=4,
                num_blocks=3, attn_resolutions=(16, 8), dropout=0.10, augment_dim=9, label_dim=0, **_):
    return dict(img_resolution=img_resolution, img_channels=img_channels, model_channels=model_channels,
                channel_mult=list(channel_mult), channel_mult_emb=channel_mult_emb, num_blocks=num_blocks,
                attn_resolutions=list(attn_resolutions), dropout=dropout, augment_dim=augment_dim, label_dim=label_dim)


def _block_shapes(prefix, cin, cout, emb, up, down, attention, out):
    """UNetBlock parameters, uncond_unet.py:173-187 (channels_per_head=64 -> heads = cout // 64)."""
    out[f"{prefix}.norm0.weight"] = (cin,)
    out[f"{prefix}.norm0.bias"] = (cin,)
    out[f"{prefix}.conv0.weight"] = (cout, cin, 3, 3)
    out[f"{prefix}.conv0.bias"] = (cout,)
    if up or down:
        out[f"{prefix}.conv0.resample_filter"] = (1, 1, 2, 2)
    out[f"{prefix}.affine.weight"] = (2 * cout, emb)
    out[f"{prefix}.affine.bias"] = (2 * cout,)
    out[f"{prefix}.norm1.weight"] = (cout,)
    out[f"{prefix}.norm1.bias"] = (cout,)
    out[f"{prefix}.conv1.weight"] = (cout, cout, 3, 3)
    out[f"{prefix}.conv1.bias"] = (cout,)
    if cout != cin or up or down:
        if cout != cin:  # kernel = 1 (resample_proj is False for DhariwalUNet)
            out[f"{prefix}.skip.weight"] = (cout, cin, 1, 1)
            out[f"{prefix}.skip.bias"] = (cout,)
        if up or down:
            out[f"{prefix}.skip.resample_filter"] = (1, 1, 2, 2)
    if attention:
        out[f"{prefix}.norm2.weight"] = (cout,)
        out[f"{prefix}.norm2.bias"] = (cout,)
        out[f"{prefix}.qkv.weight"] = (3 * cout, cout, 1, 1)
        out[f"{prefix}.qkv.bias"] = (3 * cout,)
        out[f"{prefix}.proj.weight"] = (cout, cout, 1, 1)
        out[f"{prefix}.proj.bias"] = (cout,)


def unet_layout(cfg):
    """Walks DhariwalUNet.__init__ (uncond_unet.py:467-542).  Returns (shapes, plan): ``shapes`` maps every state_dict
    key (prefix ``model.``) to its shape; ``plan`` lists the blocks in execution order per section."""
    mc, mult, nb = cfg["model_channels"], cfg["channel_mult"], cfg["num_blocks"]
    res0, cimg, attn_res = cfg["img_resolution"], cfg["img_channels"], cfg["attn_resolutions"]
    emb = mc * cfg["channel_mult_emb"]
    shapes: Dict[str, tuple] = {}
    plan = dict(enc=[], dec=[], dec2=[])
    if cfg["augment_dim"]:
        shapes["model.map_augment.weight"] = (mc, cfg["augment_dim"])
    shapes["model.map_layer0.weight"] = (emb, mc)
    shapes["model.map_layer0.bias"] = (emb,)
    shapes["model.map_layer1.weight"] = (emb, emb)
    shapes["model.map_layer1.bias"] = (emb,)
    cout = cimg
    skips = []
    for level, m in enumerate(mult):
        res = res0 >> level
        if level == 0:
            cin, cout = cout, mc * m
            shapes[f"model.enc.{res}x{res}_conv.weight"] = (cout, cin, 3, 3)
            shapes[f"model.enc.{res}x{res}_conv.bias"] = (cout,)
            plan["enc"].append(dict(name=f"{res}x{res}_conv", kind="conv", cin=cin, cout=cout))
        else:
            _block_shapes(f"model.enc.{res}x{res}_down", cout, cout, emb, False, True, False, shapes)
            plan["enc"].append(dict(name=f"{res}x{res}_down", kind="block", cin=cout, cout=cout, down=True, up=False,
                                    attention=False))
        skips.append(cout)
        for idx in range(nb):
            cin, cout = cout, mc * m
            att = res in attn_res
            _block_shapes(f"model.enc.{res}x{res}_block{idx}", cin, cout, emb, False, False, att, shapes)
            plan["enc"].append(dict(name=f"{res}x{res}_block{idx}", kind="block", cin=cin, cout=cout, down=False,
                                    up=False, attention=att))
            skips.append(cout)
    for d in ("decouple1", "decouple2"):
        shapes[f"model.{d}.0.weight"] = (cout, cout, 3, 3)
        shapes[f"model.{d}.0.bias"] = (cout,)
        shapes[f"model.{d}.1.map.weight"] = (1, cout, 1, 1)
        shapes[f"model.{d}.1.map.bias"] = (1,)
        for qk in ("q_conv", "k_conv"):
            shapes[f"model.{d}.1.{qk}.weight"] = (1, 1, 1, 1)
            shapes[f"model.{d}.1.{qk}.bias"] = (1,)
    cbott = cout
    for dec, suffix in (("dec", ""), ("dec2", "2")):
        c = cbott
        sk = list(skips)
        for level, m in reversed(list(enumerate(mult))):
            res = res0 >> level
            if level == len(mult) - 1:
                _block_shapes(f"model.{dec}.{res}x{res}_in0", c, c, emb, False, False, True, shapes)
                plan[dec].append(dict(name=f"{res}x{res}_in0", kind="block", cin=c, cout=c, down=False, up=False,
                                      attention=True))
                _block_shapes(f"model.{dec}.{res}x{res}_in1", c, c, emb, False, False, False, shapes)
                plan[dec].append(dict(name=f"{res}x{res}_in1", kind="block", cin=c, cout=c, down=False, up=False,
                                      attention=False))
            else:
                _block_shapes(f"model.{dec}.{res}x{res}_up", c, c, emb, True, False, False, shapes)
                plan[dec].append(dict(name=f"{res}x{res}_up", kind="block", cin=c, cout=c, down=False, up=True,
                                      attention=False))
            for idx in range(nb + 1):
                cin = c + sk.pop()
                c = mc * m
                att = res in attn_res
                _block_shapes(f"model.{dec}.{res}x{res}_block{idx}", cin, c, emb, False, False, att, shapes)
                plan[dec].append(dict(name=f"{res}x{res}_block{idx}", kind="block", cin=cin, cout=c, down=False,
                                      up=False, attention=att))
        shapes[f"model.out_norm{suffix}.weight"] = (c,)
        shapes[f"model.out_norm{suffix}.bias"] = (c,)
        shapes[f"model.out_conv{suffix}.weight"] = (cimg, c, 3, 3)
        shapes[f"model.out_conv{suffix}.bias"] = (cimg,)
    return shapes, plan


def make_state_dict(cfg, seed=0, std=0.02, dtype=torch.float32) -> StateDict:
    """Deterministic synthetic weights in the reference layout (every tensor N(0, std^2) except norms ~ 1 + N and the
    fixed resample filters), so that zero-initialised layers (conv1, proj, map_augment) are exercised."""
    shapes, _ = unet_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("resample_filter"):
            sd[k] = torch.full(shp, 0.25, dtype=dtype)
        elif ".norm" in k and k.endswith("weight") or "out_norm" in k and k.endswith("weight"):
            sd[k] = (1.0 + 0.1 * torch.randn(shp, generator=g)).to(dtype)
        elif k.endswith("bias"):
            sd[k] = (0.05 * torch.randn(shp, generator=g)).to(dtype)
        else:
            fan_in = int(np.prod(shp[1:])) if len(shp) > 1 else shp[0]
            sd[k] = (torch.randn(shp, generator=g) / math.sqrt(fan_in)).to(dtype)
    return sd


# ----------------------------------------------------------------------------------------------- UNet forward
def _gn(x, sd, p, eps=1e-5):
    c = x.shape[1]
    return F.group_norm(x, min(32, c // 4), sd[p + ".weight"], sd[p + ".bias"], eps)  # uncond_unet.py:122,128


def _conv(x, sd, p, up=False, down=False):
    """Conv2d.forward, uncond_unet.py:91-113 (non-fused resample branch, filter [1,1])."""
    w = sd.get(p + ".weight")
    b = sd.get(p + ".bias")
    c = x.shape[1]
    if up:
        f = torch.ones(c, 1, 2, 2, dtype=x.dtype, device=x.device)  # (f/4)*4
        x = F.conv_transpose2d(x, f, groups=c, stride=2, padding=0)
    if down:
        f = torch.full((c, 1, 2, 2), 0.25, dtype=x.dtype, device=x.device)
        x = F.conv2d(x, f, groups=c, stride=2, padding=0)
    if w is not None:
        x = F.conv2d(x, w, padding=w.shape[-1] // 2)
    if b is not None:
        x = x + b.reshape(1, -1, 1, 1)
    return x


def _block(x, emb, sd, p, blk, dropout_mask=None):
    """UNetBlock.forward, uncond_unet.py:189-211 (adaptive_scale=True, skip_scale=1, channels_per_head=64)."""
    orig = x
    x = _conv(F.silu(_gn(x, sd, p + ".norm0")), sd, p + ".conv0", up=blk["up"], down=blk["down"])
    params = (emb @ sd[p + ".affine.weight"].t() + sd[p + ".affine.bias"]).unsqueeze(2).unsqueeze(3)
    scale, shift = params.chunk(2, dim=1)
    x = F.silu(torch.addcmul(shift, _gn(x, sd, p + ".norm1"), scale + 1))
    if dropout_mask is not None:  # mask already holds 1/(1-p) scaling
        x = x * dropout_mask
    x = _conv(x, sd, p + ".conv1")
    if blk["cout"] != blk["cin"] or blk["up"] or blk["down"]:
        x = x + _conv(orig, sd, p + ".skip", up=blk["up"], down=blk["down"])
    else:
        x = x + orig
    if blk["attention"]:
        heads = blk["cout"] // 64
        qkv = _conv(_gn(x, sd, p + ".norm2"), sd, p + ".qkv")
        q, k, v = qkv.reshape(x.shape[0] * heads, x.shape[1] // heads, 3, -1).unbind(2)
        w = torch.einsum("ncq,nck->nqk", q, k / np.sqrt(k.shape[1])).softmax(dim=2)
        a = torch.einsum("nqk,nck->ncq", w, v)
        x = _conv(a.reshape(*x.shape), sd, p + ".proj") + x
    return x


def _spatial_att(x, sd, p):
    """SpatialAtt.forward, uncond_unet.py:27-37."""
    b, _, h, w = x.shape
    att = F.conv2d(x, sd[p + ".map.weight"], sd[p + ".map.bias"])
    q = F.conv2d(att, sd[p + ".q_conv.weight"], sd[p + ".q_conv.bias"]).reshape(b, 1, h * w).transpose(1, 2)
    k = F.conv2d(att, sd[p + ".k_conv.weight"], sd[p + ".k_conv.bias"]).reshape(b, 1, h * w)
    a = att.reshape(b, 1, h * w).transpose(1, 2)
    a = F.softmax(q @ k, dim=-1) @ a
    return F.softsign(a.reshape(b, 1, h, w)) * x


def positional_embedding(x, num_channels, max_positions=10000):
    """uncond_unet.py:224-230 (endpoint=False)."""
    freqs = torch.arange(0, num_channels // 2, dtype=torch.float32, device=x.device)
    freqs = freqs / (num_channels // 2)
    freqs = (1 / max_positions) ** freqs
    x = x.ger(freqs.to(x.dtype))
    return torch.cat([x.cos(), x.sin()], dim=1)


def dhariwal_forward(sd: StateDict, cfg, x, noise_labels, augment_labels=None, dropout_masks=None):
    """DhariwalUNet.forward, uncond_unet.py:544-581.  Returns (F_x, F_y)."""
    _, plan = unet_layout(cfg)
    emb = positional_embedding(noise_labels, cfg["model_channels"])
    if cfg["augment_dim"] and augment_labels is not None:
        emb = emb + augment_labels @ sd["model.map_augment.weight"].t()
    emb = F.silu(emb @ sd["model.map_layer0.weight"].t() + sd["model.map_layer0.bias"])
    emb = emb @ sd["model.map_layer1.weight"].t() + sd["model.map_layer1.bias"]
    emb = F.silu(emb)
    dm = dropout_masks or {}
    skips = []
    for blk in plan["enc"]:
        p = "model.enc." + blk["name"]
        x = _conv(x, sd, p) if blk["kind"] == "conv" else _block(x, emb, sd, p, blk, dm.get(p))
        skips.append(x)
    outs = []
    for dec, dname, suffix in (("dec", "decouple1", ""), ("dec2", "decouple2", "2")):
        h = F.conv2d(x, sd[f"model.{dname}.0.weight"], sd[f"model.{dname}.0.bias"], padding=1)
        h = _spatial_att(h, sd, f"model.{dname}.1") + x
        sk = list(skips)
        for blk in plan[dec]:
            p = f"model.{dec}." + blk["name"]
            if h.shape[1] != blk["cin"]:
                h = torch.cat([h, sk.pop()], dim=1)
            h = _block(h, emb, sd, p, blk, dm.get(p))
        h = _conv(F.silu(_gn(h, sd, f"model.out_norm{suffix}")), sd, f"model.out_conv{suffix}")
        outs.append(h)
    return outs[0], outs[1]


def edm_precond_forward(sd: StateDict, cfg, x, sigma, augment_labels=None, dropout_masks=None):
    """EDMPrecond.forward, uncond_unet.py:614-635 (precondition=True).  Returns (D_x, D_y) = (C_pred, eps_pred)."""
    x = x.to(torch.float32)
    sigma = sigma.to(torch.float32).reshape(-1, 1, 1, 1)
    q = sigma ** 2 - sigma + 1
    c_skip1 = (sigma - 1) / q
    c_skip2 = sigma.sqrt() / q
    c_out1 = torch.sqrt(sigma / q)
    c_out2 = (1 - sigma) / q.sqrt()
    c_in = 1 / torch.sqrt((1 - sigma) ** 2 + sigma)
    c_noise = sigma.log()
    f_x, f_y = dhariwal_forward(sd, cfg, c_in * x, c_noise.flatten(), augment_labels, dropout_masks)
    return c_skip1 * x + c_out1 * f_x, c_skip2 * x + c_out2 * f_y


# ----------------------------------------------------------------------------------------------- SongUNet (row f-4)
def song_config(img_resolution=32, img_channels=3, model_channels=128, channel_mult=(1, 2, 2, 2), channel_mult_emb=4,
                num_blocks=4, attn_resolutions=(16,), dropout=0.10, augment_dim=0, embedding_type="fourier",
                channel_mult_noise=2, encoder_type="residual", decoder_type="standard", resample_filter=(1, 3, 3, 1), **_):
    """Constructor defaults of SongUNet, uncond_unet.py:254-272."""
    return dict(img_resolution=img_resolution, img_channels=img_channels, model_channels=model_channels,
                channel_mult=list(channel_mult), channel_mult_emb=channel_mult_emb, num_blocks=num_blocks,
                attn_resolutions=list(attn_resolutions), dropout=dropout, augment_dim=augment_dim,
                embedding_type=embedding_type, channel_mult_noise=channel_mult_noise, encoder_type=encoder_type,
                decoder_type=decoder_type, resample_filter=list(resample_filter))


def _song_conv(x, sd, p, filt, up=False, down=False, fused=False):
    """Conv2d.forward, uncond_unet.py:91-113, every branch (general resample filter, fused resample)."""
    w = sd.get(p + ".weight")
    b = sd.get(p + ".bias")
    f = None
    if up or down:
        f1 = torch.as_tensor(filt, dtype=torch.float32)
        f = (f1.ger(f1) / f1.sum().square()).reshape(1, 1, len(filt), len(filt)).to(x.dtype)
    w_pad = w.shape[-1] // 2 if w is not None else 0
    f_pad = (f.shape[-1] - 1) // 2 if f is not None else 0
    cin = x.shape[1]
    if fused and up and w is not None:
        x = F.conv_transpose2d(x, f.mul(4).tile([cin, 1, 1, 1]), groups=cin, stride=2, padding=max(f_pad - w_pad, 0))
        x = F.conv2d(x, w, padding=max(w_pad - f_pad, 0))
    elif fused and down and w is not None:
        x = F.conv2d(x, w, padding=w_pad + f_pad)
        x = F.conv2d(x, f.tile([w.shape[0], 1, 1, 1]), groups=w.shape[0], stride=2)
    else:
        if up:
            x = F.conv_transpose2d(x, f.mul(4).tile([cin, 1, 1, 1]), groups=cin, stride=2, padding=f_pad)
        if down:
            x = F.conv2d(x, f.tile([cin, 1, 1, 1]), groups=cin, stride=2, padding=f_pad)
        if w is not None:
            x = F.conv2d(x, w, padding=w_pad)
    if b is not None:
        x = x + b.reshape(1, -1, 1, 1)
    return x


def _song_block(x, emb, sd, p, filt, up=False, down=False, dropout_mask=None):
    """UNetBlock.forward, uncond_unet.py:189-211 as SongUNet configures it (:284-288): adaptive_scale=False, num_heads=1,
    skip_scale=sqrt(0.5), eps=1e-6, resample_proj=True.  Which sub-modules exist is read off the state_dict."""
    skip_scale = math.sqrt(0.5)
    gn = lambda t, q: F.group_norm(t, min(32, t.shape[1] // 4), sd[q + ".weight"], sd[q + ".bias"], 1e-6)
    orig = x
    x = _song_conv(F.silu(gn(x, p + ".norm0")), sd, p + ".conv0", filt, up=up, down=down)
    params = (emb @ sd[p + ".affine.weight"].t() + sd[p + ".affine.bias"]).unsqueeze(2).unsqueeze(3)
    x = F.silu(gn(x + params, p + ".norm1"))
    if dropout_mask is not None:
        x = x * dropout_mask
    x = _song_conv(x, sd, p + ".conv1", filt)
    has_skip = (p + ".skip.weight") in sd or up or down
    x = x + (_song_conv(orig, sd, p + ".skip", filt, up=up, down=down) if has_skip else orig)
    x = x * skip_scale
    if (p + ".qkv.weight") in sd:
        qkv = _song_conv(gn(x, p + ".norm2"), sd, p + ".qkv", filt)
        q, k, v = qkv.reshape(x.shape[0], x.shape[1], 3, -1).unbind(2)  # num_heads = 1
        w = torch.einsum("ncq,nck->nqk", q, k / np.sqrt(k.shape[1])).softmax(dim=2)
        a = torch.einsum("nqk,nck->ncq", w, v)
        x = (_song_conv(a.reshape(*x.shape), sd, p + ".proj", filt) + x) * skip_scale
    return x


def song_forward(sd: StateDict, cfg, x, noise_labels, augment_labels=None):
    """SongUNet.forward, uncond_unet.py:359-425 (the fork's two-decoder variant).  Returns (F_x, F_y).  The module order of
    enc / dec / dec2 is the key order of the state_dict (= construction order of the ModuleDicts)."""
    filt = cfg["resample_filter"]
    nc = cfg["model_channels"] * cfg["channel_mult_noise"]
    if cfg["embedding_type"] == "positional":  # PositionalEmbedding(endpoint=True), :217-230
        freqs = torch.arange(0, nc // 2, dtype=torch.float32) / (nc // 2 - 1)
        freqs = (1 / 10000) ** freqs
        e = noise_labels.ger(freqs.to(noise_labels.dtype))
    else:  # FourierEmbedding, :236-244
        e = noise_labels.ger((2 * np.pi * sd["model.map_noise.freqs"]).to(noise_labels.dtype))
    emb = torch.cat([e.cos(), e.sin()], dim=1)
    emb = emb.reshape(emb.shape[0], 2, -1).flip(1).reshape(*emb.shape)
    if cfg["augment_dim"] and augment_labels is not None:
        emb = emb + augment_labels @ sd["model.map_augment.weight"].t()
    emb = F.silu(emb @ sd["model.map_layer0.weight"].t() + sd["model.map_layer0.bias"])
    emb = F.silu(emb @ sd["model.map_layer1.weight"].t() + sd["model.map_layer1.bias"])

    def modules(section):
        seen = []
        for k in sd:
            if k.startswith(f"model.{section}."):
                name = k.split(".")[2]
                if name not in seen:
                    seen.append(name)
        return seen

    # aux_down / aux_up (kernel 0) own no parameter: only their resample_filter buffer shows up in the state_dict, which
    # is enough to list them
    # The fork keeps a second skip list for decoder 2 (:371, :388) that is appended to but NOT updated by the aux branches
    # (`x = skips[-1] = ...` touches the first list only, :381-384): decoder 2 sees the pre-aux tensors.  Restated as is.
    skips, skips2, aux = [], [], x
    for name in modules("enc"):
        p = "model.enc." + name
        if "aux_down" in name:
            aux = _song_conv(aux, sd, p, filt, down=True)
        elif "aux_skip" in name:
            x = skips[-1] = x + _song_conv(aux, sd, p, filt)
        elif "aux_residual" in name:
            x = skips[-1] = aux = (x + _song_conv(aux, sd, p, filt, down=True, fused=True)) / np.sqrt(2)
        else:
            x = _song_conv(x, sd, p, filt) if name.endswith("_conv") else \
                _song_block(x, emb, sd, p, filt, down=name.endswith("_down"))
            skips.append(x)
            skips2.append(x)
    outs = []
    for dec, dname, src in (("dec", "decouple1", skips), ("dec2", "decouple2", skips2)):
        h = F.conv2d(x, sd[f"model.{dname}.0.weight"], sd[f"model.{dname}.0.bias"], padding=1)
        h = _spatial_att(h, sd, f"model.{dname}.1") + x
        sk = list(src)
        aux_o = tmp = None
        for name in modules(dec):
            p = f"model.{dec}." + name
            if "aux_up" in name:
                aux_o = _song_conv(aux_o, sd, p, filt, up=True)
            elif "aux_norm" in name:
                tmp = F.group_norm(h, min(32, h.shape[1] // 4), sd[p + ".weight"], sd[p + ".bias"], 1e-6)
            elif "aux_conv" in name:
                tmp = _song_conv(F.silu(tmp), sd, p, filt)
                aux_o = tmp if aux_o is None else tmp + aux_o
            else:
                cin = sd[p + ".norm0.weight"].shape[0]
                if h.shape[1] != cin:
                    h = torch.cat([h, sk.pop()], dim=1)
                h = _song_block(h, emb, sd, p, filt, up=name.endswith("_up"))
        outs.append(aux_o)
    return outs[0], outs[1]


def song_precond_forward(sd: StateDict, cfg, x, sigma, augment_labels=None):
    """EDMPrecond.forward (uncond_unet.py:614-635) around SongUNet."""
    x = x.to(torch.float32)
    sigma = sigma.to(torch.float32).reshape(-1, 1, 1, 1)
    q = sigma ** 2 - sigma + 1
    c_in = 1 / torch.sqrt((1 - sigma) ** 2 + sigma)
    f_x, f_y = song_forward(sd, cfg, c_in * x, sigma.log().flatten(), augment_labels)
    return (sigma - 1) / q * x + torch.sqrt(sigma / q) * f_x, sigma.sqrt() / q * x + (1 - sigma) / q.sqrt() * f_y


# ----------------------------------------------------------------------------------------------- DDM-const math
def q_sample(x_start, noise, t):
    """ddm_const.py:284-287 with C = -x_start (:319)."""
    c = -1 * x_start
    time = t.reshape(c.shape[0], *((1,) * (c.dim() - 1)))
    return x_start + c * time + torch.sqrt(time) * noise


def pred_x0_from_xt(xt, noise, c, t):
    """ddm_const.py:290-293."""
    time = t.reshape(c.shape[0], *((1,) * (c.dim() - 1)))
    return xt - c * time - torch.sqrt(time) * noise


def pred_xtms_from_xt(xt, noise, c, t, s, z):
    """ddm_const.py:296-303 with the Gaussian draw passed in (z replaces randn_like, :300)."""
    time = t.reshape(c.shape[0], *((1,) * (c.dim() - 1)))
    s = s.reshape(c.shape[0], *((1,) * (c.dim() - 1)))
    mean = xt + c * (time - s) - c * time - s / torch.sqrt(time) * noise
    sigma = torch.sqrt(s * (time - s) / time)
    return mean + sigma * z


def ddm_loss(c_pred, noise_pred, x_start, noise, t, eps=1e-4, weighting=True, use_l1=False):
    """ddm_const.py:335-358 with loss_main_func = MSE_Loss(reduction='sum') and loss_vlb = 0.
    Returns (loss, loss_simple_per_sample)."""
    c = -1 * x_start
    if weighting:
        w1 = (t ** 2 - t + 1) / t
        w2 = (t ** 2 - t + 1) / (1 - t + eps)
    else:
        w1 = w2 = 1
    ls = w1 * F.mse_loss(c_pred, c, reduction="none").sum(dim=[1, 2, 3]) + \
        w2 * F.mse_loss(noise_pred, noise, reduction="none").sum(dim=[1, 2, 3])
    if use_l1:
        ls = ls + w1 * (c_pred - c).abs().mean([1, 2, 3]) + w2 * (noise_pred - noise).abs().mean([1, 2, 3])
        ls = ls / 2
    return ls.sum() / c.shape[0], ls


def p_losses(model_fn: Callable, x_start, t, noise, eps=1e-4, weighting=True, use_l1=False, **model_kwargs):
    """ddm_const.py:305-364 with explicit (t, noise); augmentation is applied by the caller (it is data glue)."""
    x_noisy = q_sample(x_start, noise, t)
    c_pred, noise_pred = model_fn(x_noisy, t, **model_kwargs)
    loss, ls = ddm_loss(c_pred, noise_pred, x_start, noise, t, eps, weighting, use_l1)
    n = x_start[0].numel() * x_start.shape[0]
    # ddm_const.py:359-363 — note the reference divides the (already batch-averaged) loss by B*C*H*W again
    return loss, {"train/loss_simple": ls.detach().sum() / n, "train/loss_vlb": torch.zeros(()),
                  "train/loss": loss.detach() / n}


def t_steps_deterministic(n, sigma_min=1e-2, sigma_max=1.0):
    """ddm_const.py:429-436.  n == 1 is NaN in the reference (0/0); we define it as [sigma_max, 0]."""
    smin = sigma_min ** 2
    idx = torch.arange(n, dtype=torch.float64)
    if n == 1:
        ts = torch.tensor([float(sigma_max)], dtype=torch.float64)
    else:
        ts = sigma_max + idx / (n - 1) * (smin - sigma_max)
    return torch.cat([ts, torch.zeros_like(ts[:1])])


def sample_fn_d(model_fn: Callable, x_T, n_steps, sigma_min=1e-2, sigma_max=1.0, scale_input=1.0, clip_x_start=True,
                unnormalize=True, return_trajectory=False):
    """ddm_const.py:425-476.  x_T: fp64 standard normal noise of the sample shape."""
    ts = t_steps_deterministic(n_steps, sigma_min, sigma_max).to(x_T.device)
    x_next = x_T.to(torch.float64) * ts[0]
    traj = [x_next]
    for t_cur, t_next in zip(ts[:-1], ts[1:]):
        x_cur = x_next
        c, noise = model_fn(x_cur, t_cur)[:2]
        c, noise = c.to(torch.float64), noise.to(torch.float64)
        x0 = x_cur - c * t_cur - noise * t_cur.sqrt()
        if clip_x_start:
            x0 = x0.clamp(-1. * scale_input, 1. * scale_input)
        x_next = x0 + c * t_next + noise * t_next.sqrt()
        traj.append(x_next)
    x_next = x_next.clamp(-1. * scale_input, 1. * scale_input)
    if scale_input != 1:
        x_next = x_next / scale_input
    if unnormalize:
        x_next = (x_next + 1) * 0.5
    return (x_next, traj) if return_trajectory else x_next


def sample_fn_s(model_fn: Callable, x_T, z_list, n_steps, sigma_min=1e-2, sigma_max=1.0, scale_input=1.0,
                clip_x_start=True, unnormalize=True):
    """ddm_const.py:381-422 with the per-step Gaussian draws passed in (z_list[i] replaces randn_like, :300)."""
    dev = x_T.device
    idx = torch.arange(n_steps, dtype=torch.float64, device=dev)
    ts = (sigma_max ** 2) + idx / (n_steps - 1) * (sigma_min ** 2 - sigma_max ** 2)
    ts = torch.cat((ts, torch.tensor([0.0], dtype=torch.float64, device=dev)))
    time_steps = -torch.diff(ts)
    img = x_T.to(torch.float32)
    batch = img.shape[0]
    cur_time = torch.ones((batch,), dtype=torch.float64, device=dev)
    for i, time_step in enumerate(time_steps):
        s = torch.full((batch,), float(time_step), dtype=torch.float64, device=dev)
        if i == time_steps.shape[0] - 1:
            s = cur_time
        c, noise = model_fn(img, cur_time)[:2]
        time = cur_time.reshape(batch, 1, 1, 1)
        x0 = img - c * time - torch.sqrt(time) * noise
        if clip_x_start:
            x0 = x0.clamp(-1. * scale_input, 1. * scale_input)
        c = -1 * x0
        sv = s.reshape(batch, 1, 1, 1)
        mean = img + c * (time - sv) - c * time - sv / torch.sqrt(time) * noise
        sigma = torch.sqrt(sv * (time - sv) / time)
        img = mean + sigma * z_list[i]
        cur_time = cur_time - s
    img = img.clamp(-1. * scale_input, 1. * scale_input)
    if scale_input != 1:
        img = img / scale_input
    if unnormalize:
        img = (img + 1) * 0.5
    return img


# ----------------------------------------------------------------------------------------------- latent variant
def p_losses_latent(model_fn: Callable, x_start, t, noise, eps=1e-4, weighting=True, use_l1=True, variant="const",
                    **model_kwargs):
    """LatentDiffusion.p_losses: structure of ddm_const_2.py:527-588 (L1 as a SUM over CHW :561-564, reconstruction term
    -log(t)/2 * sum|x_rec - x0| :566-568, use_disloss off).
    variant "const"  : the sqrt(t) schedule's formulas — q_sample ddm_const.py:284-287, x_rec = pred_x0_from_xt
                       ddm_const.py:290-293, weights ddm_const.py:336-338 (the fork's own latent p_losses, :716-784, is a
                       nuScenes segmentation objective and unusable here);
    variant "const_2": the sibling file verbatim — t-linear noise (ddm_const_2.py:175,181) and weights
                       ((t-1)/t)^2+1, (t/(1-t+eps))^2+1 (:553-555)."""
    b = x_start.shape[0]
    if variant == "const_2":
        tt = t.reshape(b, 1, 1, 1)
        x_noisy = x_start + (-x_start) * tt + tt * noise
        c_pred, noise_pred = model_fn(x_noisy, t, **model_kwargs)[:2]
        x_rec = x_noisy - c_pred * tt - tt * noise_pred
        c = -x_start
        w1 = ((t - 1) / t) ** 2 + 1 if weighting else torch.ones_like(t)
        w2 = (t / (1 - t + eps)) ** 2 + 1 if weighting else torch.ones_like(t)
        ls = w1 * ((c_pred - c) ** 2).sum([1, 2, 3]) + w2 * ((noise_pred - noise) ** 2).sum([1, 2, 3])
        if use_l1:
            ls = (ls + w1 * (c_pred - c).abs().sum([1, 2, 3]) + w2 * (noise_pred - noise).abs().sum([1, 2, 3])) / 2
        vlb = (x_rec - x_start).abs().sum([1, 2, 3]) * (-torch.log(t.reshape(b, 1)) / 2)  # [B, B], as the reference
        loss = ls.sum() / b + vlb.sum() / b
        n = float(x_start.numel())
        return loss, {"train/loss_simple": ls.detach().sum() / n, "train/loss_vlb": vlb.detach().sum() / n,
                      "train/loss": loss.detach() / n}
    x_noisy = q_sample(x_start, noise, t)
    c_pred, noise_pred = model_fn(x_noisy, t, **model_kwargs)[:2]
    tt = t.reshape(b, 1, 1, 1)
    x_rec = x_noisy - c_pred * tt - torch.sqrt(tt) * noise_pred
    c = -x_start
    if weighting:
        w1 = (t ** 2 - t + 1) / t
        w2 = (t ** 2 - t + 1) / (1 - t + eps)
    else:
        w1 = w2 = torch.ones_like(t)
    ls = w1 * ((c_pred - c) ** 2).sum([1, 2, 3]) + w2 * ((noise_pred - noise) ** 2).sum([1, 2, 3])
    if use_l1:
        ls = ls + w1 * (c_pred - c).abs().sum([1, 2, 3]) + w2 * (noise_pred - noise).abs().sum([1, 2, 3])
        ls = ls / 2
    # rec_weight has shape [B, 1] in the reference (ddm_const_2.py:565): [B] * [B, 1] broadcasts to the [B, B] outer
    # product, whose sum is (sum_i a_i) * (sum_j w_j).  Reproduced as is.
    vlb = (x_rec - x_start).abs().sum([1, 2, 3]) * (-torch.log(t.reshape(b, 1)) / 2)
    loss = ls.sum() / b + vlb.sum() / b
    n = float(x_start.numel())
    return loss, {"train/loss_simple": ls.detach().sum() / n, "train/loss_vlb": vlb.detach().sum() / n,
                  "train/loss": loss.detach() / n}


def sample_fn_latent(model_fn: Callable, x_T, n_steps, sigma_min=1e-2, sigma_max=1.0):
    """ddm_const.py:868-888 (fp64 state, no clamp)."""
    ts = t_steps_deterministic(n_steps, sigma_min, sigma_max).to(x_T.device)
    x = x_T.to(torch.float64) * ts[0]
    for t_cur, t_next in zip(ts[:-1], ts[1:]):
        c, noise = model_fn(x, t_cur)[:2]
        c, noise = c.to(torch.float64), noise.to(torch.float64)
        x = x + (t_next - t_cur) * (c + noise / (t_cur.sqrt() + t_next.sqrt()))
    return x
