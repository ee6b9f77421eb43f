"""Generates tests/golden/ae_small.pt by running the UNMODIFIED reference ``ddm.encoder_decoder.AutoencoderKL`` from
/root/reference (build container only; CPU fp32, eval).  Its training-only loss module (LPIPS + discriminator: needs a
VGG download) is stubbed with nn.Identity; weights come from OUR mirror built under torch.manual_seed(SEED)
and are loaded into the reference (strict over encoder / decoder / quant convs).

    python tests/golden/make_golden_ae.py
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

SEED = 17
DDCONFIG = dict(double_z=True, z_channels=3, resolution=[64, 64], in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 4],
                num_res_blocks=1, attn_resolutions=[], dropout=0.0)


def build_ours():
    from adm_b200.ddm.encoder_decoder import AutoencoderKL
    torch.manual_seed(SEED)
    ae = AutoencoderKL(ddconfig=DDCONFIG, lossconfig={}, embed_dim=3).eval()
    # default inits give near-identical channels; spread the norm / conv parameters a little so that errors show
    g = torch.Generator().manual_seed(SEED + 1)
    with torch.no_grad():
        for n, p in ae.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return ae


def inputs():
    g = torch.Generator().manual_seed(3)
    return 2 * torch.rand(2, 3, 64, 64, generator=g) - 1


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference not mounted; golden files can only be regenerated in the build container")
    sys.path.insert(0, REF)
    adm = types.ModuleType("ADM")
    adm.__path__ = [REF]
    sys.modules["ADM"] = adm
    import ddm.encoder_decoder as ed
    ed.LPIPSWithDiscriminator = lambda **kw: torch.nn.Identity()
    ours = build_ours()
    sd = ours.state_dict()
    ref = ed.AutoencoderKL(ddconfig=DDCONFIG, lossconfig={}, embed_dim=3).eval()
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("loss.") for k in missing), (missing, unexpected)
    x = inputs()
    with torch.no_grad():
        post = ref.encode(x)
        dec = ref.decode(post.mode())
    torch.save({"keys": {k: list(v.shape) for k, v in sd.items()}, "mean": post.mean, "logvar": post.logvar, "dec": dec},
               os.path.join(HERE, "ae_small.pt"))
    print("mean", post.mean.norm().item(), "logvar", post.logvar.norm().item(), "dec", dec.norm().item(), "keys", len(sd))


if __name__ == "__main__":
    main()
