"""Pins the DDM-const MATH directly against the reference's own functions (build container only).

    python tests/golden/make_golden_ddm.py        ->  tests/golden/ddm_math.pt

/root/reference/ddm/ddm_const.py cannot be imported as is (it needs ldm, cldm, pytorch_lightning, nuScenesSegDataset,
tools: ddm_const.py:10-20).  None of those take part in the arithmetic of the hot path, so they are replaced by empty
stub modules and the UNMODIFIED methods of ``ddm.ddm_const.DDPM`` are then called as plain functions on a small
namespace object that carries the attributes they read:

  * q_sample :284-287, pred_x0_from_xt :290-293, pred_xtms_from_xt :296-303 (its randn_like draw is recorded);
  * p_losses :305-364 with loss_main_func = ddm.loss.MSE_Loss (the sibling's default, ddm_const_2.py:98) and the LPIPS
    term replaced by zeros (the reference cannot build LPIPS offline and crashes with perceptual_weight = 0),
    weighting on/off, use_l1 on/off;
  * sample_fn_d :425-476 and sample_fn_s :381-422 (their randn draws are recorded) at N = 2, 5, 10;
  * LatentDiffusion.p_losses of the sibling ddm/ddm_const_2.py:527-588 (its own weights ((t-1)/t)^2+1, (t/(1-t+eps))^2+1).

The denoiser is a closed-form stand-in ``toy_model`` (no UNet), restated in tests/test_oracle.py, so the fixture pins
exactly the diffusion arithmetic.  Everything is CPU fp32/fp64 on seeded inputs.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

SHAPE = (4, 3, 8, 8)


def toy_model(x, t, *args, **kwargs):
    """Deterministic two-output 'denoiser' (C_pred, eps_pred) of (x, t); t is [B] or 0-dim."""
    t = torch.as_tensor(t, dtype=x.dtype, device=x.device)
    tt = t.reshape(-1, 1, 1, 1) if t.dim() else t
    return torch.tanh(0.7 * x - tt), 0.5 * torch.sin(x + 2.0 * tt)


def import_reference_ddm():
    if not os.path.isdir(REF):
        raise SystemExit("reference not mounted; golden files can only be regenerated in the build container")
    sys.path.insert(0, REF)
    adm = types.ModuleType("ADM")
    adm.__path__ = [REF]
    sys.modules["ADM"] = adm

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Any:  # placeholder classes / callables the import statements bind
        def __init__(self, *a, **k):
            pass

    stub("ldm"); stub("ldm.modules"); stub("ldm.modules.distributions")
    stub("ldm.modules.distributions.distributions", DiagonalGaussianDistribution=_Any)
    stub("ldm.util", instantiate_from_config=_Any)
    stub("ldm.modules.ema", LitEma=_Any)
    stub("cldm"); stub("cldm.loss", compute_layer_weights=_Any, SegmentationLoss=_Any)
    stub("pytorch_lightning", LightningModule=torch.nn.Module)
    stub("nuScenesSegDataset", nuScenesSegDataset=_Any)
    stub("tools"); stub("tools.training_log_analysis", parse_csv_and_plot=_Any)
    import ddm.ddm_const as const
    import ddm.loss as rloss
    return const, rloss


class _Recorder:
    """Records every torch.randn / randn_like draw made inside the reference functions."""

    def __enter__(self):
        self.draws = []
        self._randn, self._randn_like = torch.randn, torch.randn_like

        def randn(*a, **k):
            out = self._randn(*a, **k)
            self.draws.append(out.clone())
            return out

        def randn_like(*a, **k):
            out = self._randn_like(*a, **k)
            self.draws.append(out.clone())
            return out
        torch.randn, torch.randn_like = randn, randn_like
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like


def main():
    const, rloss = import_reference_ddm()
    D = const.DDPM
    g = torch.Generator().manual_seed(11)
    x = 2 * torch.rand(SHAPE, generator=g) - 1
    t = torch.rand(SHAPE[0], generator=g) * (1 - 1e-4) + 1e-4
    noise = torch.randn(SHAPE, generator=g)
    s = t * torch.rand(SHAPE[0], generator=g) * 0.9
    out = {"x": x, "t": t, "noise": noise, "s": s}
    ns = types.SimpleNamespace()
    C = -x
    out["q_sample"] = D.q_sample(ns, x_start=x, noise=noise, t=t, C=C)
    xt = out["q_sample"]
    out["pred_x0_from_xt"] = D.pred_x0_from_xt(ns, xt, noise, C, t)
    torch.manual_seed(5)
    with _Recorder() as rec:
        out["pred_xtms_from_xt"] = D.pred_xtms_from_xt(ns, xt, noise, C, t, s)
    out["pred_xtms_z"] = rec.draws[0]

    # ---- p_losses (image space), four flag combinations
    for weighting in (True, False):
        for use_l1 in (False, True):
            ns = types.SimpleNamespace(
                start_dist="normal", use_augment=False, model=toy_model, weighting_loss=weighting, use_l1=use_l1,
                eps=torch.tensor(1e-4), loss_main_func=rloss.MSE_Loss(), perceptual_weight=1.,
                perceptual_loss=lambda a, b: torch.zeros(a.shape[0], 1, 1, 1))
            ns.q_sample = lambda **kw: D.q_sample(ns, **kw)
            torch.manual_seed(21)
            with _Recorder() as rec:
                loss, ld = D.p_losses(ns, x.clone(), t)
            key = f"p_losses_w{int(weighting)}_l1{int(use_l1)}"
            out[key] = {"loss": loss, "noise": rec.draws[0], **{k: v for k, v in ld.items()}}

    # ---- samplers
    for n in (2, 5, 10):
        ns = types.SimpleNamespace(eps=torch.tensor(1e-4), sampling_timesteps=n, sigma_min=1e-2, sigma_max=1,
                                   model=toy_model, clip_x_start=True, scale_input=1, start_dist="normal")
        ns.pred_x0_from_xt = lambda *a: D.pred_x0_from_xt(ns, *a)
        ns.pred_xtms_from_xt = lambda *a: D.pred_xtms_from_xt(ns, *a)
        torch.manual_seed(31 + n)
        with _Recorder() as rec:
            img = D.sample_fn_d.__wrapped__(ns, SHAPE) if hasattr(D.sample_fn_d, "__wrapped__") else D.sample_fn_d(ns, SHAPE)
        out[f"sample_fn_d_{n}"] = {"x_T": rec.draws[0], "img": img}
        torch.manual_seed(41 + n)
        with _Recorder() as rec:
            img = D.sample_fn_s(ns, SHAPE)
        out[f"sample_fn_s_{n}"] = {"x_T": rec.draws[0], "z": torch.stack(rec.draws[1:]), "img": img}

    # ---- the sibling's latent loss (const_2 weights), ddm_const_2.py:527-588
    sys.modules.pop("ddm.ddm_const", None)
    import ddm.ddm_const_2 as c2
    L = c2.LatentDiffusion
    for use_l1 in (False, True):
        ns = types.SimpleNamespace(start_dist="normal", model=toy_model, weighting_loss=True, use_l1=use_l1,
                                   eps=torch.tensor(1e-4), loss_main_func=rloss.MSE_Loss(), cfg={}, perceptual_weight=0.)
        ns.q_sample = lambda **kw: c2.DDPM.q_sample(ns, **kw)
        ns.pred_x0_from_xt = lambda *a: c2.DDPM.pred_x0_from_xt(ns, *a)
        torch.manual_seed(51)
        with _Recorder() as rec:
            loss, ld = L.p_losses(ns, x.clone(), t)
        out[f"latent2_p_losses_l1{int(use_l1)}"] = {"loss": loss, "noise": rec.draws[0], **{k: v for k, v in ld.items()}}

    def strip(o):
        if isinstance(o, dict):
            return {k: strip(v) for k, v in o.items()}
        return o.detach().clone() if torch.is_tensor(o) else o
    torch.save(strip(out), os.path.join(HERE, "ddm_math.pt"))
    print("wrote ddm_math.pt:", sorted(out))


if __name__ == "__main__":
    main()
