"""One profiled training step (cudaProfilerStart/Stop) after warm-up; run under `ncu --profile-from-start off`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from adm_b200.train import TrainStep

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda", 0)
dpm = bench.build_model(dev)
dpm.train()
step = TrainStep(dpm)
x = 2 * torch.rand(B, 3, 32, 32, device=dev) - 1
for _ in range(3):
    step.micro_step(x)
    step.optimizer_step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
if mode == "train":
    step.micro_step(x)
    step.optimizer_step()
else:
    dpm.eval()
    dpm.sampling_timesteps = 2
    dpm.sample(batch_size=B)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one", mode, "step")
