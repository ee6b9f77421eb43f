"""Records one augmented batch of the UNMODIFIED reference AugmentPipe (ddm/augment.py:115-328) with the arguments the
reference's DDM module uses (ddm/ddm_const.py:179-180), on CPU under a fixed global seed (build container only).

    python tests/golden/make_golden_augment.py      ->  tests/golden/augment.pt
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SEEDS = (123, 7)
KW = dict(p=0.15, xflip=1e8, yflip=1, scale=1, rotate_frac=1, aniso=1, translate_frac=1)
KW_HOT = dict(KW, p=1.0)  # every transform fires: exercises the full warp on every sample


def inputs(seed, n=16):
    g = torch.Generator().manual_seed(1000 + seed)
    return 2 * torch.rand(n, 3, 32, 32, generator=g) - 1


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference not mounted; golden files can only be regenerated in the build container")
    sys.path.insert(0, REF)
    adm = types.ModuleType("ADM")
    adm.__path__ = [REF]
    sys.modules["ADM"] = adm
    from ddm.augment import AugmentPipe
    out = {}
    for name, kw in (("cfg", KW), ("hot", KW_HOT)):
        for seed in SEEDS:
            x = inputs(seed)
            torch.manual_seed(seed)
            y, lab = AugmentPipe(**kw)(x)
            out[f"{name}_{seed}"] = {"images": y.clone(), "labels": lab.clone()}
    torch.save(out, os.path.join(HERE, "augment.pt"))
    print({k: (tuple(v["images"].shape), tuple(v["labels"].shape)) for k, v in out.items()})


if __name__ == "__main__":
    main()
