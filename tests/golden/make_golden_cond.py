"""Generates tests/golden/cond_unet_small.{json,pt} by running the UNMODIFIED reference ``unet.cond_unet.Unet`` from
/root/reference (build container only; CPU fp32, eval mode).

    python tests/golden/make_golden_cond.py

The reference module needs two absent packages only for an import and a base class (``fvcore``, ``pytorch_lightning``;
SURVEY §8c) — they are stubbed — and its Swin factory is called with weights=None instead of downloading ImageNet
weights.  Weights: our mirror is constructed under torch.manual_seed(SEED) and its state_dict is loaded STRICTLY into the
reference module (which also pins the 1158-key layout); the GPU test rebuilds the same weights from the seed.
Recorded: key -> shape map, output probes of (C_pred, eps_pred), the DDM-const loss (image-space weights) and the latent
loss (L1 sum + reconstruction term), gradient norms of every parameter group probed and full gradients of a few small
tensors, and a 3-step latent sampler end point.
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

SEED = 11
CFG = dict(dim=32, dim_mults=[1, 2, 4, 4], cond_in_dim=3, cond_dim=32, cond_dim_mults=[], channels=3, out_mul=1,
           cond_net="swin", fix_bb=False, window_sizes1=[[8, 8], [4, 4], [2, 2], [1, 1]],
           window_sizes2=[[4, 4], [2, 2], [1, 1], [1, 1]], fourier_scale=16, cond_pe=False, num_pos_feats=128,
           cond_feature_size=[128, 128])
BATCH, LAT, COND = 2, 32, 64
GRAD_KEYS = ["downs.0.0.block1.proj.weight", "downs.0.0.block1.norm.weight", "downs.0.0.mlp.1.weight",
             "downs.1.2.fn.fn.to_qkv.weight", "downs.1.2.fn.fn.to_out.1.g", "mid_attn.fn.fn.to_qkv.weight",
             "mid_attn.fn.fn.to_out.bias", "decouple1.1.weight", "decouple2.2.map.weight", "ups.0.0.block1.proj.weight",
             "ups2.3.3.weight", "ups.1.3.1.weight", "final_conv.weight", "final_conv2.bias", "time_mlp.1.weight",
             "projects.1.weight", "relation_layers_down.0.attentions.0.q_lin.weight", "downs.0.3.weight",
             "init_conv.0.weight", "final_res_block2.block2.proj.weight"]


def build_ours():
    from adm_b200.unet.cond_unet import Unet
    torch.manual_seed(SEED)
    return Unet(**CFG).eval()


def inputs(seed=5):
    g = torch.Generator().manual_seed(seed)
    x = 2 * torch.rand(BATCH, 3, LAT, LAT, generator=g) - 1
    t = torch.rand(BATCH, generator=g) * 0.8 + 0.1
    noise = torch.randn(BATCH, 3, LAT, LAT, generator=g)
    cond = 2 * torch.rand(BATCH, 3, COND, COND, generator=g) - 1
    return x, t, noise, cond


def import_reference_cond():
    if not os.path.isdir(REF):
        raise SystemExit("reference not mounted; golden files can only be regenerated in the build container")
    sys.path.insert(0, REF)
    fv = types.ModuleType("fvcore"); fvc = types.ModuleType("fvcore.common"); fvcc = types.ModuleType("fvcore.common.config")
    fvcc.CfgNode = dict
    fv.common = fvc; fvc.config = fvcc
    sys.modules.update({"fvcore": fv, "fvcore.common": fvc, "fvcore.common.config": fvcc})
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = torch.nn.Module
    sys.modules["pytorch_lightning"] = pl
    import unet.swin_transformer as st
    orig = st.swin_b
    st.swin_b = lambda weights=None, **kw: orig(weights=None, **kw)
    import unet.cond_unet as cu
    return cu


def main():
    from oracle import ddm_oracle as O
    ours = build_ours()
    sd = ours.state_dict()
    cu = import_reference_cond()
    ref = cu.Unet(**CFG).eval()
    ref.load_state_dict(sd, strict=True)  # identical key set and shapes, or this raises
    x, t, noise, cond = inputs()
    fn = lambda xx, tt: ref(xx, tt, cond)
    out = {"seed": SEED, "cfg": CFG, "batch": BATCH, "n_keys": len(sd),
           "keys": {k: list(v.shape) for k, v in sd.items() if not k.startswith("init_conv_mask.")},
           "n_swin_keys": sum(k.startswith("init_conv_mask.") for k in sd)}
    with torch.no_grad():
        xt = O.q_sample(x, noise, t)
        c_pred, e_pred = fn(xt, t)
    out["c_pred_probe"] = c_pred.flatten()[::97][:32].tolist()
    out["e_pred_probe"] = e_pred.flatten()[::97][:32].tolist()
    out["c_pred_norm"], out["e_pred_norm"] = c_pred.norm().item(), e_pred.norm().item()
    loss_img, _ = O.p_losses(fn, x, t, noise)
    out["loss_image_space"] = loss_img.item()
    ref.zero_grad()
    loss_lat, ld = O.p_losses_latent(fn, x, t, noise, use_l1=True, weighting=True)
    loss_lat.backward()
    out["loss_latent"] = loss_lat.item()
    out["loss_latent_vlb"] = ld["train/loss_vlb"].item()
    grads = dict(ref.named_parameters())
    out["grad_norms"] = {k: grads[k].grad.norm().item() for k in GRAD_KEYS}
    small = {k: grads[k].grad.clone() for k in GRAD_KEYS if grads[k].numel() <= 40000}
    with torch.no_grad():
        g = torch.Generator().manual_seed(9)
        x_T = torch.randn(BATCH, 3, LAT, LAT, generator=g, dtype=torch.float64)
        z = O.sample_fn_latent(fn, x_T, 3)
    torch.save({"grads": small, "c_pred": c_pred, "e_pred": e_pred, "sample_latent": z},
               os.path.join(HERE, "cond_unet_small.pt"))
    json.dump(out, open(os.path.join(HERE, "cond_unet_small.json"), "w"), indent=0)
    print("loss_image_space", out["loss_image_space"], "loss_latent", out["loss_latent"], "keys", out["n_keys"])


if __name__ == "__main__":
    main()
