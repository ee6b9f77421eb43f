"""Generates tests/golden/*.json|*.pt by running the UNMODIFIED reference from /root/reference (build container only).

    python tests/golden/make_golden.py

What is recorded (all on CPU, fp32 unless noted):
  * unet_tiny.json   — reference EDMPrecond (tiny config) : state_dict key->shape map, output checksums + probe values,
                       loss of the DDM-const step, gradient norms of selected parameters, 5-step sampler output stats.
  * unet_cifar.json  — same for the CIFAR-10 config of configs/cifar10/ddm_uncond_const_uncond_unet.yaml at B=2
                       (forward + loss only) plus the full 829-key state_dict layout.
  * const2_step.json — the sibling ddm.ddm_const_2.DDPM (the importable class with the upstream API) run through its own
                       training_step() plumbing with LPIPS stubbed to 0, against the restated step with const_2's
                       three formulas.
  * sample_tiny.pt   — the fp64 trajectory end point of the tiny model (8x3x16x16).
Weights are regenerated from a seed by oracle.ddm_oracle.make_state_dict, so no weight tensors are stored.
"""
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference():
    """Put the reference on sys.path (also as package `ADM`, needed by ddm/augment.py:15-16)."""
    if not os.path.isdir(REF):
        raise SystemExit("reference not mounted; golden files can only be regenerated in the build container")
    sys.path.insert(0, REF)
    adm = types.ModuleType("ADM")
    adm.__path__ = [REF]
    sys.modules["ADM"] = adm
    import unet.uncond_unet as uu
    return uu


def checksum(t):
    t = t.detach().double()
    flat = t.flatten()
    idx = torch.linspace(0, flat.numel() - 1, 7).long()
    return dict(sum=t.sum().item(), abssum=t.abs().sum().item(), sqsum=(t * t).sum().item(),
                probes=[flat[i].item() for i in idx])


TINY = dict(img_resolution=16, img_channels=3, model_channels=64, channel_mult=[1, 2], channel_mult_emb=4,
            num_blocks=1, attn_resolutions=[8], dropout=0.0, augment_dim=9, label_dim=0)
CIFAR = dict(img_resolution=32, img_channels=3, model_channels=192, channel_mult=[1, 2, 2, 2], channel_mult_emb=4,
             num_blocks=3, attn_resolutions=[16, 8], dropout=0.1, augment_dim=9, label_dim=0)


def build_ref(uu, cfg, sd):
    kw = {k: v for k, v in cfg.items() if k not in ("img_resolution", "img_channels", "label_dim")}
    net = uu.EDMPrecond(img_resolution=cfg["img_resolution"], img_channels=cfg["img_channels"], sigma_data=1.0,
                        model_type="DhariwalUNet", **kw)
    missing, unexpected = net.load_state_dict(sd, strict=True)
    net.eval()  # dropout off; parity tests inject masks explicitly
    return net


def inputs(cfg, b, seed):
    g = torch.Generator().manual_seed(seed)
    r = cfg["img_resolution"]
    x = 2 * torch.rand(b, 3, r, r, generator=g) - 1
    t = torch.rand(b, generator=g) * (1 - 1e-4) + 1e-4
    noise = torch.randn(b, 3, r, r, generator=g)
    aug = 0.5 * torch.randn(b, 9, generator=g)
    return x, t, noise, aug


def run_case(uu, cfg, b, with_grad, n_sample_steps):
    from oracle import ddm_oracle as O
    sd = O.make_state_dict(cfg, seed=0)
    net = build_ref(uu, cfg, sd)
    ref_shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    x, t, noise, aug = inputs(cfg, b, 1)
    out = dict(cfg=cfg, batch=b, state_dict_shapes=ref_shapes)
    # --- reference forward of the DDM-const step (math restated from ddm_const.py, UNet = the reference's)
    x_noisy = O.q_sample(x, noise, t)
    params = [p for p in net.parameters()]
    for p in params:
        p.requires_grad_(with_grad)
    c_pred, e_pred = net(x_noisy, t, augment_labels=aug)
    loss, ls = O.ddm_loss(c_pred, e_pred, x, noise, t, 1e-4, True, False)
    out["c_pred"] = checksum(c_pred)
    out["eps_pred"] = checksum(e_pred)
    out["loss"] = loss.item()
    out["loss_per_sample"] = ls.tolist()
    if with_grad:
        loss.backward()
        names = [n for n, _ in net.named_parameters()]
        pick = names[::max(1, len(names) // 24)]
        out["grad_norms"] = {n: dict(net.named_parameters())[n].grad.double().norm().item() for n in pick}
        out["grad_probe"] = {n: checksum(dict(net.named_parameters())[n].grad) for n in pick[:6]}
    # --- oracle on the same inputs must agree with the reference module
    with torch.no_grad():
        oc, oe = O.edm_precond_forward(sd, cfg, x_noisy, t, augment_labels=aug)
        out["oracle_vs_ref_maxabs"] = max((oc - c_pred).abs().max().item(), (oe - e_pred).abs().max().item())
        if n_sample_steps:
            g = torch.Generator().manual_seed(7)
            r = cfg["img_resolution"]
            x_T = torch.randn(b, 3, r, r, generator=g, dtype=torch.float64)
            fn = lambda xx, tt: net(xx, tt)
            img = O.sample_fn_d(fn, x_T, n_sample_steps)
            out["sample"] = checksum(img)
            out["sample_steps"] = n_sample_steps
            out["_sample_tensor"] = img
    return out


def run_const2(uu):
    """Cross-check the step plumbing against the importable sibling class ddm.ddm_const_2.DDPM."""
    from oracle import ddm_oracle as O
    import torch.nn as nn
    # stubs for packages the sibling imports but the CIFAR path never calls
    import ddm.ddm_const_2 as c2

    class ZeroLPIPS(nn.Module):
        def forward(self, a, b):
            return torch.zeros(a.shape[0], 1, 1, 1)

    c2.LPIPS = lambda: ZeroLPIPS()
    cfg = TINY
    sd = O.make_state_dict(cfg, seed=0)
    net = build_ref(uu, cfg, sd)
    model_cfg = dict(image_size=[16, 16], sampling_timesteps=5, loss_type="l2", start_dist="normal",
                     perceptual_weight=1.0, eps=1e-4, sigma_max=1, sigma_min=0.01, weighting_loss=True, use_l1=False,
                     use_augment=False)
    dpm = c2.DDPM(model=net, cfg=model_cfg, **model_cfg)
    x, t, noise, _ = inputs(cfg, 4, 3)
    torch.manual_seed(11)
    loss_ref, ld = dpm.training_step({"image": x})
    # replay the same RNG draws: forward() draws t (:167), p_losses draws noise (:199)
    torch.manual_seed(11)
    t2 = torch.rand(4) * (1. - dpm.eps) + dpm.eps
    noise2 = torch.randn_like(x)
    time = t2.reshape(4, 1, 1, 1)
    x_noisy = x + (-x) * time + time * noise2  # const_2 schedule (:175)
    c_pred, e_pred = net(x_noisy, t2)
    w1 = ((t2 - 1) / t2) ** 2 + 1  # const_2 weights (:228-230)
    w2 = (t2 / (1 - t2 + dpm.eps)) ** 2 + 1
    ls = w1 * ((c_pred + x) ** 2).sum([1, 2, 3]) + w2 * ((e_pred - noise2) ** 2).sum([1, 2, 3])
    loss2 = ls.sum() / 4
    return dict(loss_ref=loss_ref.item(), loss_restated=loss2.item(), state_dict_keys=len(dpm.state_dict()),
                loss_dict={k: float(v) for k, v in ld.items()})


def main():
    uu = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(8)
    tiny = run_case(uu, TINY, 8, True, 5)
    torch.save(tiny.pop("_sample_tensor"), os.path.join(HERE, "sample_tiny.pt"))
    json.dump(tiny, open(os.path.join(HERE, "unet_tiny.json"), "w"), indent=1)
    print("tiny: loss", tiny["loss"], "oracle_vs_ref", tiny["oracle_vs_ref_maxabs"])
    cifar = run_case(uu, CIFAR, 2, False, 0)
    json.dump(cifar, open(os.path.join(HERE, "unet_cifar.json"), "w"), indent=1)
    print("cifar: loss", cifar["loss"], "oracle_vs_ref", cifar["oracle_vs_ref_maxabs"], "keys",
          len(cifar["state_dict_shapes"]))
    try:
        c2 = run_const2(uu)
        json.dump(c2, open(os.path.join(HERE, "const2_step.json"), "w"), indent=1)
        print("const_2:", c2)
    except Exception as e:  # the sibling needs taming/LPIPS imports; record why if it cannot run
        print("const_2 cross-check unavailable:", repr(e))


if __name__ == "__main__":
    main()
