"""Generates tests/golden/relation_layer.pt by running the UNMODIFIED reference ``unet.cond_unet.BasicAttetnionLayer``
(/root/reference/unet/cond_unet.py:153-252; build container only; CPU fp32, eval mode) on seeded inputs:

    python tests/golden/make_golden_relation.py

Recorded: the layer's state_dict (drawn after the reference's own init, with the GroupNorm affine and the conv biases
perturbed so that they are exercised), the inputs x1 (condition features) / x2 (trunk features), the output, and the
gradients of sum(output * probe) with respect to both inputs and every parameter.  Two cases: a low-resolution condition
map (4x4 -> 16x16, the DIV2K config's geometry: Swin features are resized up to the trunk resolution) and equal
resolutions with ragged windows (the zero padding of F.pad + AvgPool2d).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.make_golden_cond import import_reference_cond  # noqa: E402

CASES = {
    "lowres_cond": dict(embed_dim=64, nhead=8, ffn_dim=128, window_size1=[2, 2], window_size2=[4, 4], b=2,
                        hw1=(4, 4), hw2=(16, 16)),
    "ragged_windows": dict(embed_dim=64, nhead=8, ffn_dim=128, window_size1=[4, 4], window_size2=[3, 3], b=1,
                           hw1=(10, 10), hw2=(10, 10)),
}


def make(cu, spec, seed):
    torch.manual_seed(seed)
    layer = cu.BasicAttetnionLayer(embed_dim=spec["embed_dim"], nhead=spec["nhead"], ffn_dim=spec["ffn_dim"],
                                   window_size1=spec["window_size1"], window_size2=spec["window_size2"]).eval()
    with torch.no_grad():
        layer.gn.weight.add_(0.3 * torch.randn_like(layer.gn.weight))
        layer.gn.bias.add_(0.2 * torch.randn_like(layer.gn.bias))
        for m in (layer.concat_conv, layer.out_conv, layer.mlp.fc1, layer.mlp.fc2, layer.q_lin, layer.k_lin, layer.v_lin):
            m.bias.add_(0.1 * torch.randn_like(m.bias))
    c, b = spec["embed_dim"], spec["b"]
    x1 = torch.randn(b, c, *spec["hw1"]).requires_grad_(True)
    x2 = (torch.randn(b, c, *spec["hw2"]) * 1.3 + 0.2).requires_grad_(True)
    probe = torch.randn(b, c, *spec["hw2"])
    out = layer(x1, x2)
    (out * probe).sum().backward()
    return dict(spec=spec, state_dict={k: v.detach().clone() for k, v in layer.state_dict().items()},
                x1=x1.detach(), x2=x2.detach(), probe=probe, out=out.detach(), dx1=x1.grad.clone(), dx2=x2.grad.clone(),
                grads={k: p.grad.clone() for k, p in layer.named_parameters()})


def main():
    cu = import_reference_cond()
    rec = {name: make(cu, spec, 100 + i) for i, (name, spec) in enumerate(CASES.items())}
    path = os.path.join(HERE, "relation_layer.pt")
    torch.save(rec, path)
    print("wrote", path, {k: tuple(v["out"].shape) for k, v in rec.items()}, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
