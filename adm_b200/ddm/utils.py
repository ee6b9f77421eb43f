"""Config -> object boundary of the reference (ddm/utils.py:67-68, :94-161): the YAML files name classes by dotted
path (``ddm.ddm_const.DDPM``, ``unet.uncond_unet.EDMPrecond``); those names resolve to this package's implementations."""
from __future__ import annotations

import importlib

# reference module path -> adm_b200 module path
ALIASES = {
    "ddm.ddm_const": "adm_b200.ddm.ddm_const",
    "ddm.utils": "adm_b200.ddm.utils",
    "unet.uncond_unet": "adm_b200.unet.uncond_unet",
    "unet.cond_unet": "adm_b200.unet.cond_unet",
    "ddm.encoder_decoder": "adm_b200.ddm.encoder_decoder",
    "ddm.ema": "adm_b200.ddm.ema",
}


def exists(x):
    return x is not None


def default(val, d):
    if exists(val):
        return val
    return d() if callable(d) else d


def identity(t, *args, **kwargs):
    return t


def normalize_to_neg_one_to_one(img):
    return img * 2 - 1


def unnormalize_to_zero_to_one(t):
    return (t + 1) * 0.5


def get_obj_by_name(name: str):
    parts = name.split(".")
    for i in range(len(parts) - 1, 0, -1):
        mod_name, attr = ".".join(parts[:i]), parts[i:]
        mod_name = ALIASES.get(mod_name, mod_name)
        try:
            obj = importlib.import_module(mod_name)
        except ImportError:
            continue
        try:
            for a in attr:
                obj = getattr(obj, a)
            return obj
        except AttributeError:
            continue
    raise ImportError(name)


def call_func_by_name(*args, func_name: str = None, **kwargs):
    assert func_name is not None
    fn = get_obj_by_name(func_name)
    assert callable(fn)
    return fn(*args, **kwargs)


def construct_class_by_name(*args, class_name: str = None, **kwargs):
    """ddm/utils.py:159-161."""
    return call_func_by_name(*args, func_name=class_name, **kwargs)
