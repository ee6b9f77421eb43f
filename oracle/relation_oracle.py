"""ORACLE — test infrastructure only.  Never import this from the product (adm_b200/); only tests/ may.

A CPU restatement, in plain functional PyTorch (fp32 / fp64), of one relation layer of the reference's conditional UNet
and of the algorithms its sm_100a kernels (adm_b200/csrc/relation_ops.cu) implement:

* ``relation_layer``            — /root/reference/unet/cond_unet.py:192-252 (``BasicAttetnionLayer.forward``) with
  ``PositionEmbeddingSine.forward`` (:36-66, normalize=False, temperature 10000) and ``Mlp.forward`` (:146-152, eval mode)
  restated inline.  Parameters come from a flat state_dict with the reference's key names.  ``commute_out_conv=True``
  applies ``out_conv`` to the pooled tokens BEFORE the bilinear resize — the order the fused kernel path uses; the two
  orders are the same linear map (a 1x1 conv acts per pixel, the align_corners resize weights sum to one).
* ``bilinear_bwd_separable``    — the transpose of ``F.interpolate(mode='bilinear', align_corners=True)`` the way
  ``adm_bilinear_bwd`` computes it: rows then columns, the (i0, i1, lambda) triplets re-derived from
  ``src = dst * (n_in - 1) / (n_out - 1)`` exactly as the forward derives them.
* ``window_pool``               — ``F.pad`` to a multiple of the window + ``AvgPool2d`` (:200-215), NHWC.
* ``relation_tail`` / ``relation_tail_bwd`` — GroupNorm(x + y) * gamma + beta + resize(z) and its hand-derived backward
  (d pre = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dout * gamma), the formulas of adm_rel_gn_fwd / bwd.

Parity pin: tests/golden/relation_layer.pt is recorded from the UNMODIFIED reference module by
tests/golden/make_golden_relation.py; tests/test_oracle.py checks ``relation_layer`` (both orders) against it and the
three algorithm restatements against torch autograd.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def position_embedding_sine(x, temperature=10000.0):
    """cond_unet.py:36-66 for x [b, h, w, d] (normalize=False): pos [b, h, w, d] = (pos_y | pos_x)."""
    b, h, w, d = x.shape
    npf = d // 2
    y_embed = torch.arange(1, h + 1, dtype=torch.float32).reshape(1, h, 1).expand(b, h, w)
    x_embed = torch.arange(1, w + 1, dtype=torch.float32).reshape(1, 1, w).expand(b, h, w)
    dim_t = torch.arange(npf, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / npf)
    pos_x = x_embed[..., None] / dim_t
    pos_y = y_embed[..., None] / dim_t
    pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).to(x.dtype)


def _pad_to(x, win):  # NCHW, zero pad right / bottom to a multiple of the window (:200-209)
    ph = (win[0] - x.shape[2] % win[0]) % win[0]
    pw = (win[1] - x.shape[3] % win[1]) % win[1]
    return F.pad(x, (0, pw, 0, ph))


def relation_layer(sd, x1, x2, nhead, window_size1, window_size2, groups=8, eps=1e-5, commute_out_conv=False):
    """x1 (condition features, queries) and x2 (trunk features, keys / values): NCHW.  Returns NCHW."""
    dt = x1.dtype
    p = {k: v.to(dt) for k, v in sd.items()}
    b, c1, h1, w1 = x1.shape
    _, c2, h2, w2 = x2.shape
    up = F.interpolate(x1, size=(h2, w2), mode="bilinear", align_corners=True)
    shortcut = x2 + F.conv2d(torch.cat([up, x2], dim=1), p["concat_conv.weight"], p["concat_conv.bias"])
    shortcut = F.group_norm(shortcut, groups, p["gn.weight"], p["gn.bias"], eps)
    x1p, x2p = _pad_to(x1, window_size1), _pad_to(x2, window_size2)
    x1_s = F.avg_pool2d(x1p, tuple(window_size1))
    qg = x1_s.permute(0, 2, 3, 1)
    qg = (qg + position_embedding_sine(qg)).reshape(b, -1, c2)
    kg = F.avg_pool2d(x2p, tuple(window_size2)).permute(0, 2, 3, 1)
    kg = (kg + position_embedding_sine(kg)).reshape(b, -1, c1)
    nq, nk, hd = qg.shape[1], kg.shape[1], c1 // nhead
    q = F.linear(qg, p["q_lin.weight"], p["q_lin.bias"]).reshape(b, nq, nhead, hd).permute(0, 2, 1, 3)
    k = F.linear(kg, p["k_lin.weight"], p["k_lin.bias"]).reshape(b, nk, nhead, hd).permute(0, 2, 1, 3)
    v = F.linear(kg, p["v_lin.weight"], p["v_lin.bias"]).reshape(b, nk, nhead, hd).permute(0, 2, 1, 3)
    attn = torch.softmax(q @ k.transpose(-2, -1), dim=-1)  # no 1 / sqrt(d), as the reference
    o = (attn @ v).transpose(1, 2).reshape(b, nq, c1)
    o = o.transpose(1, 2).reshape(b, c1, x1_s.shape[2], x1_s.shape[3])
    x1_s = x1_s + o
    hid = F.relu(F.conv2d(x1_s, p["mlp.fc1.weight"], p["mlp.fc1.bias"]))
    x1_s = x1_s + F.conv2d(hid, p["mlp.fc2.weight"], p["mlp.fc2.bias"])
    if commute_out_conv:
        z = F.conv2d(x1_s, p["out_conv.weight"], p["out_conv.bias"])
        return shortcut + F.interpolate(z, size=(h2, w2), mode="bilinear", align_corners=True)
    x1_s = F.interpolate(x1_s, size=(h2, w2), mode="bilinear", align_corners=True)
    return shortcut + F.conv2d(x1_s, p["out_conv.weight"], p["out_conv.bias"])


# ------------------------------------------------------------------------------------------------ kernel algorithms
def _lerp_coord(o, n_in, n_out):
    """(i0, i1, weight of i1) of output index o, fp32 arithmetic as the kernels (and ATen) do it."""
    scale = torch.tensor((n_in - 1) / (n_out - 1) if n_out > 1 else 0.0, dtype=torch.float32)
    s = scale * torch.tensor(float(o), dtype=torch.float32)
    i0 = min(int(s.item()), n_in - 1)
    i1 = i0 + (1 if i0 < n_in - 1 else 0)
    return i0, i1, float(s.item() - i0)


def _axis_matrix(n_in, n_out, dtype):
    """[n_out, n_in] interpolation matrix of one axis."""
    m = torch.zeros(n_out, n_in, dtype=dtype)
    for o in range(n_out):
        i0, i1, l1 = _lerp_coord(o, n_in, n_out)
        m[o, i0] += 1.0 - l1
        m[o, i1] += l1
    return m


def bilinear_fwd(x, size):
    """x [B, h, w, C] -> [B, H, W, C], separable form of the 4-tap gather of bilinear_fwd_kernel."""
    mh, mw = _axis_matrix(x.shape[1], size[0], x.dtype), _axis_matrix(x.shape[2], size[1], x.dtype)
    return torch.einsum("oh,bhwc,pw->bopc", mh, x, mw)


def bilinear_bwd_separable(dy, size_in):
    """dy [B, H, W, C] -> dx [B, h, w, C]: rows first (tmp [B, h, W, C]), then columns — adm_bilinear_bwd."""
    mh, mw = _axis_matrix(size_in[0], dy.shape[1], dy.dtype), _axis_matrix(size_in[1], dy.shape[2], dy.dtype)
    tmp = torch.einsum("oh,bowc->bhwc", mh, dy)
    return torch.einsum("pw,bhpc->bhwc", mw, tmp)


def window_pool(x, window):
    """x [B, H, W, C] -> [B, ceil(H / kh), ceil(W / kw), C]; pixels past the edge count as zero (F.pad + AvgPool2d)."""
    b, h, w, c = x.shape
    kh, kw = window
    ho, wo = -(-h // kh), -(-w // kw)
    xp = torch.zeros(b, ho * kh, wo * kw, c, dtype=x.dtype)
    xp[:, :h, :w] = x
    return xp.reshape(b, ho, kh, wo, kw, c).sum(dim=(2, 4)) / (kh * kw)


def relation_tail(x, y, z, gamma, beta, groups, eps=1e-5):
    """out = GroupNorm(x + y) * gamma + beta + resize(z); x, y [B, H, W, C], z [B, hq, wq, C].  Returns (out, mean, rstd)
    with the statistics per (sample, group) over H * W * (C / groups) elements, biased variance."""
    b, h, w, c = x.shape
    pre = (x + y).reshape(b, h * w, groups, c // groups)
    mean = pre.mean(dim=(1, 3), keepdim=True)
    var = ((pre - mean) ** 2).mean(dim=(1, 3), keepdim=True)
    rstd = (var + eps).rsqrt()
    xhat = ((pre - mean) * rstd).reshape(b, h, w, c)
    return xhat * gamma + beta + bilinear_fwd(z, (h, w)), mean.reshape(b, groups), rstd.reshape(b, groups)


def relation_tail_bwd(dout, x, y, gamma, groups, eps=1e-5):
    """(d pre, d gamma, d beta, d z-resize input) of relation_tail by the closed-form GroupNorm backward."""
    b, h, w, c = x.shape
    cpg = c // groups
    pre = (x + y).reshape(b, h * w, groups, cpg)
    mean = pre.mean(dim=(1, 3), keepdim=True)
    rstd = (((pre - mean) ** 2).mean(dim=(1, 3), keepdim=True) + eps).rsqrt()
    xhat = (pre - mean) * rstd
    d = dout.reshape(b, h * w, groups, cpg)
    g = d * gamma.reshape(1, 1, groups, cpg)
    m2 = g.mean(dim=(1, 3), keepdim=True)
    m1 = (g * xhat).mean(dim=(1, 3), keepdim=True)
    dpre = (rstd * (g - m2 - xhat * m1)).reshape(b, h, w, c)
    dgamma = (d * xhat).sum(dim=(0, 1)).reshape(c)
    dbeta = d.sum(dim=(0, 1)).reshape(c)
    return dpre, dgamma, dbeta
