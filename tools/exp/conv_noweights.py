"""Upper bound of what weight-tile multicast could buy the 3x3 halo conv: time it with the weight TMA loads skipped
(ADM_GEMM_DEBUG=2: the MMAs run on whatever the ring holds, results are garbage) against the real kernel."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from adm_b200 import ops
    from tools.bench_convs_lib import timeit
    for cin, cout, res in [(384, 384, 16), (192, 192, 32), (768, 384, 16), (384, 384, 32)]:
        x = torch.randn(128, res, res, cin, device="cuda").bfloat16()
        w = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda") / (3 * cin ** 0.5))
        out = torch.empty(128, res, res, cout, device="cuda", dtype=torch.bfloat16)
        t = timeit(lambda: ops.conv_fprop(x, w, out=out))
        fl = 2.0 * 128 * res * res * cin * cout * 9
        print(f"  {cin}->{cout} @{res}: {t*1000:7.1f} us {fl/t/1e9:7.1f} TF/s", flush=True)
    sys.exit(0)
for dbg in ("0", "2", "0", "2"):
    env = dict(os.environ, ADM_GEMM_DEBUG=dbg)
    print(f"ADM_GEMM_DEBUG={dbg} ({'weight loads skipped' if dbg == '2' else 'real kernel'})", flush=True)
    subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, cwd=ROOT)
