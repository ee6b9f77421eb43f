"""Golden vectors of the UNMODIFIED reference ``EDMPrecond(model_type='SongUNet')`` (unet/uncond_unet.py:253-441, :588-635)
for two small configurations: DDPM++ (positional embedding, standard encoder, [1,1] resampling) and NCSN++ (Fourier
embedding, residual encoder with the fused-resample aux convs, [1,3,3,1] resampling) — build container only.

    python tests/golden/make_golden_song.py       ->  tests/golden/song_unet.pt

Weights are NOT stored: ``state_dict_for`` derives every parameter from a seeded generator by key order and shape, so the
test re-creates the same tensors and loads them (strict) into the module under test.  Recorded per config: the state_dict
key -> shape map, D_x / D_y on seeded inputs, the gradient norm of every parameter and the full gradient of a spread of
tensors (those of at most 20 000 elements, to keep the fixture small) for the scalar  sum(D_x * g1) + sum(D_y * g2).
"""
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

CONFIGS = {
    "ddpmpp": dict(img_resolution=16, img_channels=3, sigma_data=1.0, model_type="SongUNet", model_channels=32,
                   channel_mult=[1, 2], channel_mult_emb=4, num_blocks=1, attn_resolutions=[8], dropout=0.0, augment_dim=9,
                   embedding_type="positional", channel_mult_noise=1, encoder_type="standard", decoder_type="standard",
                   resample_filter=[1, 1]),
    "ncsnpp": dict(img_resolution=16, img_channels=3, sigma_data=1.0, model_type="SongUNet", model_channels=32,
                   channel_mult=[1, 2], channel_mult_emb=4, num_blocks=1, attn_resolutions=[8], dropout=0.0, augment_dim=0,
                   embedding_type="fourier", channel_mult_noise=2, encoder_type="residual", decoder_type="standard",
                   resample_filter=[1, 3, 3, 1]),
}
BATCH = 4


def state_dict_for(shapes, seed=0):
    """Every parameter (and the Fourier frequencies) from a seeded generator, by key order; resample_filter buffers are the
    module's own constants and are left alone.  shapes: ordered {key: shape}."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("resample_filter"):
            continue
        shp = tuple(shp)
        if k.endswith("map_noise.freqs"):
            sd[k] = torch.randn(shp, generator=g) * 2.0
        elif ".norm" in k or "aux_norm" in k:
            sd[k] = (1.0 + 0.1 * torch.randn(shp, generator=g)) if k.endswith("weight") else 0.1 * torch.randn(shp, generator=g)
        elif len(shp) > 1:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[k] = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        else:
            sd[k] = 0.1 * torch.randn(shp, generator=g)
    return sd


def inputs(seed=3):
    g = torch.Generator().manual_seed(seed)
    x = 2 * torch.rand(BATCH, 3, 16, 16, generator=g) - 1
    t = torch.rand(BATCH, generator=g) * (1 - 1e-4) + 1e-4
    aug = torch.randn(BATCH, 9, generator=g) * 0.5
    g1 = torch.randn(BATCH, 3, 16, 16, generator=g)
    g2 = torch.randn(BATCH, 3, 16, 16, generator=g)
    return x, t, aug, g1, g2


def probe_keys(keys):
    """A spread of tensors whose full gradients are recorded: every 7th parameter plus the first / last few."""
    keys = [k for k in keys if not k.endswith("resample_filter") and not k.endswith("freqs")]
    return sorted(set(keys[::7] + keys[:4] + keys[-4:]), key=keys.index)


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference not mounted; golden files can only be regenerated in the build container")
    sys.path.insert(0, REF)
    from unet import uncond_unet as R
    out = {}
    for name, cfg in CONFIGS.items():
        torch.manual_seed(0)
        net = R.EDMPrecond(**cfg).eval()
        shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
        missing, unexpected = net.load_state_dict(state_dict_for(shapes), strict=False)
        assert not unexpected and all(k.endswith("resample_filter") for k in missing), (missing, unexpected)
        x, t, aug, g1, g2 = inputs()
        kw = {"augment_labels": aug} if cfg["augment_dim"] else {}
        d_x, d_y = net(x, t, **kw)
        ((d_x * g1).sum() + (d_y * g2).sum()).backward()
        grads = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
        out[name] = dict(keys=shapes, d_x=d_x.detach(), d_y=d_y.detach(),
                         grad_norms={k: float(v.norm()) for k, v in grads.items()},
                         grads={k: grads[k].clone() for k in probe_keys(list(shapes))
                                if k in grads and grads[k].numel() <= 20000})
        print(name, len(shapes), "keys; |D_x|", float(d_x.norm()), "|D_y|", float(d_y.norm()), "probes", len(out[name]["grads"]))
    torch.save(out, os.path.join(HERE, "song_unet.pt"))


if __name__ == "__main__":
    main()
