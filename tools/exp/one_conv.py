import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from adm_b200 import ops
w = ops.pack_conv_weight(torch.randn(384, 384, 3, 3, device="cuda") / 60)
x = torch.randn(8, 16, 16, 384, device="cuda").bfloat16()
out = torch.empty(8, 16, 16, 384, device="cuda", dtype=torch.bfloat16)
ops.conv_fprop(x, w, out=out)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
