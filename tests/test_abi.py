"""CPU: the C-ABI shared library loads and exports every symbol include/adm_b200.h declares; the ctypes table matches."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "adm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    names = _declared()
    assert len(names) >= 30
    for must in ("adm_qsample", "adm_ddm_loss", "adm_sampler_step", "adm_conv_fprop", "adm_conv_dgrad",
                 "adm_conv_wgrad", "adm_gemm_batched", "adm_gn_stats", "adm_gn_apply", "adm_gn_bwd", "adm_adamw"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from adm_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in adm_b200.h but not exported"
    assert lib.adm_version() >= 100
    assert lib.adm_last_error() is not None


def test_ctypes_table_covers_header():
    from adm_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: the product path fails loudly without a CUDA tensor."""
    import torch
    from adm_b200 import ops
    from adm_b200.unet.uncond_unet import EDMPrecond
    x = torch.zeros(2, 3, 8, 8)
    with pytest.raises(RuntimeError):
        ops.qsample(x, x, torch.ones(2))
    net = EDMPrecond(img_resolution=16, img_channels=3, model_channels=64, channel_mult=[1, 2], num_blocks=1,
                     attn_resolutions=[8], augment_dim=9)
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            net(torch.zeros(1, 3, 16, 16), torch.ones(1))
