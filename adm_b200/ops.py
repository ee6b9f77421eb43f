"""Tensor-level wrappers around the C-ABI kernels.  torch is used for device memory and streams only.

Conventions: activations are NHWC bf16 tensors of shape [N, H, W, C] whose last dim is contiguous and whose pixel
stride ``ld = x.stride(2)`` may exceed C (channel-slice views).  Every function enqueues on torch's current stream.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import GemmDesc, Operand, check

BF16 = torch.bfloat16
F32 = torch.float32


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("adm_b200 kernels need CUDA tensors (there is no CPU fallback)")


def _ptr(t):
    return None if t is None else t.data_ptr()


def pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def _nhwc(x):
    """(ptr, C, ld, N, H, W) of an NHWC bf16 view."""
    assert x.dim() == 4 and x.dtype == BF16 and x.stride(3) == 1, (x.shape, x.dtype, x.stride())
    n, h, w, c = x.shape
    ld = x.stride(2)
    assert x.stride(1) == ld * w and x.stride(0) == ld * w * h, "NHWC view must be dense over pixels"
    return x.data_ptr(), c, ld, n, h, w


# ------------------------------------------------------------------------------------------------ DDM elementwise
def qsample(x0, noise, t):
    _need_cuda(x0, noise, t)
    x0, noise, t = x0.contiguous(), noise.contiguous(), t.contiguous().float()
    out = torch.empty_like(x0)
    b = x0.shape[0]
    check(_lib.load().adm_qsample(_ptr(x0), _ptr(noise), _ptr(t), _ptr(out), b, x0.numel() // b, _stream()), "qsample")
    return out


def ddm_loss(c_pred, eps_pred, x0, noise, t, eps, weighting, use_l1=False, grad_scale=1.0, need_grad=True):
    """Returns (loss_per_sample [B], dC_pred, dEps_pred); grads are d(mean_b loss_b * grad_scale)."""
    _need_cuda(c_pred, eps_pred, x0, noise, t)
    c_pred, eps_pred, x0, noise = (a.contiguous() for a in (c_pred, eps_pred, x0, noise))
    b = x0.shape[0]
    loss = torch.empty(b, device=x0.device, dtype=F32)
    dc = torch.empty_like(c_pred) if need_grad else None
    de = torch.empty_like(eps_pred) if need_grad else None
    check(_lib.load().adm_ddm_loss(_ptr(c_pred), _ptr(eps_pred), _ptr(x0), _ptr(noise), _ptr(t.contiguous().float()),
                                   float(eps), int(bool(weighting)), int(bool(use_l1)), float(grad_scale), _ptr(loss),
                                   _ptr(dc), _ptr(de), b, x0.numel() // b, _stream()), "ddm_loss")
    return loss, dc, de


def sampler_step(x, c_pred, eps_pred, t_cur, t_next, clip=1.0, do_clip=True, last=False, scale_input=1.0):
    _need_cuda(x, c_pred, eps_pred)
    assert x.dtype in (torch.float64, torch.float32)
    x = x.contiguous()
    out = torch.empty_like(x)
    check(_lib.load().adm_sampler_step(_ptr(x), _ptr(c_pred.contiguous()), _ptr(eps_pred.contiguous()), _ptr(out),
                                       float(t_cur), float(t_next), float(clip), int(do_clip), int(last),
                                       float(scale_input), int(x.dtype == torch.float64), x.numel(), _stream()),
          "sampler_step")
    return out


def sampler_step_stochastic(x, c_pred, eps_pred, z, t_cur, s, clip=1.0, do_clip=True):
    _need_cuda(x, c_pred, eps_pred, z)
    x = x.contiguous().float()
    out = torch.empty_like(x)
    check(_lib.load().adm_sampler_step_stochastic(_ptr(x), _ptr(c_pred.contiguous()), _ptr(eps_pred.contiguous()),
                                                  _ptr(z.contiguous().float()), _ptr(out), float(t_cur), float(s),
                                                  float(clip), int(do_clip), x.numel(), _stream()),
          "sampler_step_stochastic")
    return out


# ------------------------------------------------------------------------------------------------ EDM edges
def unet_input(x_nchw, sigma, ld_out=8):
    _need_cuda(x_nchw, sigma)
    n, c, h, w = x_nchw.shape
    out = torch.empty(n, h, w, ld_out, device=x_nchw.device, dtype=BF16)
    sigma = sigma.reshape(-1).float().contiguous()
    check(_lib.load().adm_unet_input(_ptr(x_nchw.contiguous().float()), _ptr(sigma), int(sigma.numel() == 1),
                                     _ptr(out), n, c, h, w, ld_out, _stream()), "unet_input")
    return out


def unet_output(f1, f2, x_nchw, sigma):
    """f1, f2: fp32 NHWC [N,H,W,ldf]; returns (D1, D2) NCHW fp32."""
    n, c, h, w = x_nchw.shape
    ldf = f1.shape[-1]
    d1 = torch.empty(n, c, h, w, device=x_nchw.device, dtype=F32)
    d2 = torch.empty_like(d1)
    sigma = sigma.reshape(-1).float().contiguous()
    check(_lib.load().adm_unet_output(_ptr(f1), _ptr(f2), ldf, _ptr(x_nchw.contiguous().float()), _ptr(sigma),
                                      int(sigma.numel() == 1), _ptr(d1), _ptr(d2), n, c, h, w, _stream()),
          "unet_output")
    return d1, d2


def unet_output_bwd(dd1, dd2, sigma, ld_out=8):
    n, c, h, w = dd1.shape
    df1 = torch.empty(n, h, w, ld_out, device=dd1.device, dtype=BF16)
    df2 = torch.empty_like(df1)
    sigma = sigma.reshape(-1).float().contiguous()
    if sigma.numel() == 1:
        sigma = sigma.expand(n).contiguous()
    check(_lib.load().adm_unet_output_bwd(_ptr(dd1.contiguous()), _ptr(dd2.contiguous()), _ptr(sigma), _ptr(df1),
                                          _ptr(df2), n, c, h, w, ld_out, _stream()), "unet_output_bwd")
    return df1, df2


# ------------------------------------------------------------------------------------------------ GEMM engine
def pack_conv_weight(w, c1=None, c2=0):
    """fp32 [cout, c1+c2, k, k] -> bf16 [cout, k*k, pad64(c1)+pad64(c2)]."""
    _need_cuda(w)
    cout, cin, k, _ = w.shape
    c1 = cin if c1 is None else c1
    assert c1 + c2 == cin
    kpad = pad64(c1) + (pad64(c2) if c2 else 0)
    out = torch.empty(cout, k * k, kpad, device=w.device, dtype=BF16)
    check(_lib.load().adm_pack_conv_weight(_ptr(w.contiguous().float()), _ptr(out), cout, c1, c2, k, _stream()),
          "pack_conv_weight")
    return out


def unpack_conv_wgrad(dw_packed, c1, c2, k, out=None, accumulate=False):
    cout = dw_packed.shape[0]
    if out is None:
        out = torch.empty(cout, c1 + c2, k, k, device=dw_packed.device, dtype=F32)
        accumulate = False
    check(_lib.load().adm_unpack_conv_wgrad(_ptr(dw_packed), _ptr(out), cout, c1, c2, k, int(accumulate), _stream()),
          "unpack_conv_wgrad")
    return out


def cast_bf16(x):
    _need_cuda(x)
    x = x.contiguous().float()
    out = torch.empty(x.shape, device=x.device, dtype=BF16)
    check(_lib.load().adm_cast_f32_bf16(_ptr(x), _ptr(out), x.numel(), _stream()), "cast_f32_bf16")
    return out


def conv_fprop(x1, wpk, x2=None, bias=None, residual=None, alpha=1.0, out_dtype=BF16, out=None, nout=None):
    """Implicit-GEMM conv (3x3 pad 1 or 1x1) on NHWC bf16.  wpk: [nout, taps, kpad] bf16."""
    _need_cuda(x1, wpk)
    p1, c1, ld1, n, h, w = _nhwc(x1)
    p2, c2, ld2 = None, 0, 0
    if x2 is not None:
        p2, c2, ld2, n2, h2, w2 = _nhwc(x2)
        assert (n2, h2, w2) == (n, h, w)
    nout_w, ntaps, kpad = wpk.shape
    nout = nout_w if nout is None else nout
    assert kpad == pad64(c1) + (pad64(c2) if x2 is not None else 0), (kpad, c1, c2)
    if out is None:
        ldc = (nout + 7) // 8 * 8 if out_dtype == BF16 else (nout + 3) // 4 * 4
        out = torch.empty(n, h, w, ldc, device=x1.device, dtype=out_dtype)
        if ldc != nout:
            out.zero_()
    ldc = out.stride(2)
    ldr = 0
    if residual is not None:
        assert residual.dtype == BF16 and residual.stride(3) == 1
        ldr = residual.stride(2)
    check(_lib.load().adm_conv_fprop(p1, c1, ld1, p2, c2, ld2, n, h, w, _ptr(wpk), nout, ntaps, _ptr(out),
                                     0 if out.dtype == BF16 else 1, ldc, _ptr(bias), _ptr(residual), ldr, float(alpha),
                                     _stream()), "conv_fprop")
    return out[..., :nout] if out.shape[-1] != nout else out


def conv_dgrad(dy, wpk, n_valid=None, residual=None, alpha=1.0, out=None):
    """dx = conv^T(dy) using the fprop-packed weights.  Returns NHWC bf16 [N,H,W,n_valid]."""
    _need_cuda(dy, wpk)
    pd, cout, ldd, n, h, w = _nhwc(dy)
    nout_w, ntaps, kpad = wpk.shape
    assert cout <= nout_w
    n_valid = kpad if n_valid is None else n_valid
    if out is None:
        out = torch.empty(n, h, w, (n_valid + 7) // 8 * 8, device=dy.device, dtype=BF16)
    ldr = residual.stride(2) if residual is not None else 0
    check(_lib.load().adm_conv_dgrad(pd, cout, ldd, n, h, w, _ptr(wpk), kpad, ntaps, _ptr(out), n_valid,
                                     out.stride(2), _ptr(residual), ldr, float(alpha), _stream()), "conv_dgrad")
    return out[..., :n_valid] if out.shape[-1] != n_valid else out


def conv_wgrad(dy, x1, x2=None, ntaps=9, out=None):
    """Packed weight gradient fp32 [cout, taps, kpad] (accumulated into `out` when given)."""
    _need_cuda(dy, x1)
    pd, cout, ldd, n, h, w = _nhwc(dy)
    p1, c1, ld1, n1, h1, w1 = _nhwc(x1)
    assert (n1, h1, w1) == (n, h, w)
    p2, c2, ld2 = None, 0, 0
    if x2 is not None:
        p2, c2, ld2, _, _, _ = _nhwc(x2)
    kpad = pad64(c1) + (pad64(c2) if x2 is not None else 0)
    if out is None:
        out = torch.zeros(cout, ntaps, kpad, device=dy.device, dtype=F32)
    check(_lib.load().adm_conv_wgrad(pd, cout, ldd, p1, c1, ld1, p2, c2, ld2, n, h, w, ntaps, _ptr(out), _stream()),
          "conv_wgrad")
    return out


def _operand(t, mn_major, dims, strides, c0=0, c0_lo=0, c1=0, c1_lo=0, bhi=0, blo=0):
    return Operand(t.data_ptr(), int(mn_major), dims[0], dims[1], dims[2], strides[0], strides[1], c0, c0_lo, c1,
                   c1_lo, bhi, blo)


def gemm_nt(a, b, bias=None, out_dtype=F32, alpha=1.0, out=None):
    """C[M,N] = alpha * A[M,K] @ B[N,K]^T + bias.  a, b: bf16 row-major (last dim contiguous)."""
    _need_cuda(a, b)
    assert a.dtype == BF16 and b.dtype == BF16 and a.stride(1) == 1 and b.stride(1) == 1
    m, k = a.shape
    n, k2 = b.shape
    assert k == k2
    if out is None:
        out = torch.empty(m, n, device=a.device, dtype=out_dtype)
    d = GemmDesc()
    d.a = _operand(a, 0, (k, m, 1), (a.stride(0), a.stride(0) * m))
    d.b = _operand(b, 0, (k, n, 1), (b.stride(0), b.stride(0) * n))
    d.m, d.n, d.k, d.batches, d.bdiv, d.splits = m, n, k, 1, 1, 1
    d.c, d.out_mode, d.ldc = out.data_ptr(), 0 if out.dtype == BF16 else 1, out.stride(0)
    d.bias, d.residual, d.alpha = _ptr(bias), None, float(alpha)
    check(_lib.load().adm_gemm_batched(d, _stream()), "gemm_nt")
    return out


def gemm_nn(a, b, out_dtype=F32, alpha=1.0, out=None):
    """C[M,N] = alpha * A[M,K] @ B[K,N]; B row-major (N contiguous) consumed as an MN-major operand."""
    _need_cuda(a, b)
    assert a.dtype == BF16 and b.dtype == BF16 and a.stride(1) == 1 and b.stride(1) == 1
    m, k = a.shape
    k2, n = b.shape
    assert k == k2
    if out is None:
        out = torch.empty(m, n, device=a.device, dtype=out_dtype)
    d = GemmDesc()
    d.a = _operand(a, 0, (k, m, 1), (a.stride(0), a.stride(0) * m))
    d.b = _operand(b, 1, (n, k, 1), (b.stride(0), b.stride(0) * k))
    d.m, d.n, d.k, d.batches, d.bdiv, d.splits = m, n, k, 1, 1, 1
    d.c, d.out_mode, d.ldc = out.data_ptr(), 0 if out.dtype == BF16 else 1, out.stride(0)
    d.alpha = float(alpha)
    check(_lib.load().adm_gemm_batched(d, _stream()), "gemm_nn")
    return out


def gemm_tn(a, b, out_dtype=F32, alpha=1.0, out=None, splits=1):
    """C[M,N] = alpha * A[K,M]^T @ B[K,N]; both row-major, both consumed MN-major (the wgrad shape)."""
    _need_cuda(a, b)
    assert a.dtype == BF16 and b.dtype == BF16 and a.stride(1) == 1 and b.stride(1) == 1
    k, m = a.shape
    k2, n = b.shape
    assert k == k2
    if out is None:
        out = (torch.zeros if splits > 1 else torch.empty)(m, n, device=a.device, dtype=F32 if splits > 1 else out_dtype)
    d = GemmDesc()
    d.a = _operand(a, 1, (m, k, 1), (a.stride(0), a.stride(0) * k))
    d.b = _operand(b, 1, (n, k, 1), (b.stride(0), b.stride(0) * k))
    d.m, d.n, d.k, d.batches, d.bdiv, d.splits = m, n, k, 1, 1, splits
    d.c, d.out_mode, d.ldc = out.data_ptr(), (2 if splits > 1 else (0 if out.dtype == BF16 else 1)), out.stride(0)
    d.alpha = float(alpha)
    check(_lib.load().adm_gemm_batched(d, _stream()), "gemm_tn")
    return out
