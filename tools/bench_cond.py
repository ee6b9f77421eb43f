"""Throughput of the conditional latent path at the DIV2K config's full size (configs/super-resolution/
div2k_cond_ddm_const_ldm.yaml: cond_unet dim 128, mults 1-2-4-4, latent 128x128x3, condition 3x128x128, Swin-B):
LatentDiffusion.p_losses forward + backward (identity first stage: the AutoencoderKL is row f-1), and a 10-step sample.
Usage: python tools/bench_cond.py [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adm_b200.unet.cond_unet import Unet
from adm_b200.ddm.ddm_const import LatentDiffusion

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
CFG = dict(dim=128, dim_mults=[1, 2, 4, 4], cond_in_dim=3, cond_dim=128, cond_dim_mults=[], channels=3, out_mul=1,
           cond_net="swin", fix_bb=False, window_sizes1=[[8, 8], [4, 4], [2, 2], [1, 1]],
           window_sizes2=[[4, 4], [2, 2], [1, 1], [1, 1]], fourier_scale=16, cond_pe=False, num_pos_feats=128,
           cond_feature_size=[128, 128])


class AE(torch.nn.Module):
    down_ratio = 1
    def encode(self, x): return x
    def decode(self, z): return z


# everything runs on a side stream: autograd's AccumulateGrad nodes must not be bound to the default stream if the step
# is to be captured into a CUDA graph later
_side = torch.cuda.Stream()
torch.cuda.set_stream(_side)
torch.manual_seed(0)
net = Unet(**CFG).cuda()
mcfg = dict(image_size=[128, 128], sampling_timesteps=10, eps=1e-4, sigma_max=1, sigma_min=0.01, weighting_loss=True,
            use_l1=True, scale_factor=0.195, scale_by_std=True, default_scale=True)
ldm = LatentDiffusion(auto_encoder=AE(), model=net, cfg=mcfg, **mcfg).cuda()
print(f"cond Unet parameters: {sum(p.numel() for p in net.parameters())/1e6:.1f} M")
x = 2 * torch.rand(B, 3, 128, 128, device="cuda") - 1
cond = 2 * torch.rand(B, 3, 128, 128, device="cuda") - 1
net.eval()  # BatchNorm running stats / no dropout in RelationNet: deterministic timing


def step():
    net.zero_grad(set_to_none=True)
    loss, _ = ldm.training_step({"image": x, "cond": cond})
    loss.backward()
    return loss


for _ in range(3):
    l = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 5
for _ in range(n):
    l = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"[eager] train fwd+bwd B={B}: {ms:.1f} ms/step  {B / ms * 1000:.2f} img/s  {1153.9 * B / ms:.1f} TFLOP/s (1153.9 GFLOP/img)  loss {l.item():.1f}")
# the same step captured into ONE CUDA graph (whole-network capture: forward, loss, autograd backward)
try:
    torch.cuda.synchronize()
    net.zero_grad(set_to_none=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        static_loss = step()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"[graph] train fwd+bwd B={B}: {ms:.1f} ms/step  {B / ms * 1000:.2f} img/s  {1153.9 * B / ms:.1f} TFLOP/s  loss {static_loss.item():.1f}")
except Exception as e:
    print("graph capture failed:", repr(e)[:300])
torch.cuda.synchronize()
with torch.no_grad():
    ldm.sample(cond=cond)
    torch.cuda.synchronize()
    t0 = time.time()
    img = ldm.sample(cond=cond)
    torch.cuda.synchronize()
    dt = time.time() - t0
print(f"10-step sample B={B}: {dt * 1000:.1f} ms  {B / dt:.2f} img/s  {384.6 * 10 * B / dt / 1000:.1f} TFLOP/s; out {tuple(img.shape)}")

if os.environ.get("ADM_PROFILE"):
    import collections
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    tot = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            nm = ev.name.replace("void at::native::(anonymous namespace)::", "at::").replace("void at::native::", "at::")[:110]
            tot[nm][0] += 1
            tot[nm][1] += ev.device_time_total
    allt = sum(v[1] for v in tot.values())
    print(f"kernel time of one training step: {allt / 1000:.1f} ms")
    for nm, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{us / 1000:8.2f} ms {100 * us / allt:5.1f}%  n={c:5d}  {nm}")
