"""Samples SM clock / power with nvidia-smi (20 ms period) while the sampler or the training step runs for ~3 s."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from adm_b200.train import TrainStep
mode = sys.argv[1] if len(sys.argv) > 1 else "sample"
dev = torch.device("cuda", 0)
dpm = bench.build_model(dev)
step = TrainStep(dpm)
x = 2 * torch.rand(128, 3, 32, 32, device=dev) - 1
if mode == "train":
    dpm.train(); step.capture(x); run = lambda: step.replay(x)
else:
    dpm.eval(); dpm.sample(batch_size=128); run = lambda: dpm.sample(batch_size=128)
torch.cuda.synchronize()
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active",
                         "--format=csv,noheader,nounits", "-i", "0", "-lms", "20"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()
time.sleep(0.3)
t0 = time.time()
while time.time() - t0 < 3.0:
    run()
torch.cuda.synchronize()
t1 = time.time()
time.sleep(0.2)
proc.terminate()
print(f"{mode}: load from {0:.2f} to {t1 - t0:.2f} s")
for t, l in rows[::4]:
    print(f"{t - t0:6.2f}s  {l}")
