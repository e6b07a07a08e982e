"""Where does a sharded training step (mc.MCTrainStep + Adam, BBBAlexNet B=512 LRT) spend its time: device kernel time
per step (torch profiler) against the wall time per step, top kernels, launch count."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from bench import build_net
from pytorch_bayesiancnn_b200 import mc

torch.set_num_threads(1)
dev = torch.device("cuda:0")
B = 512
net = build_net("lrt", 10, dev, os.environ.get("BBB_B200_MATH", "bf16"))
x = torch.randn(B, 3, 32, 32, device=dev)
labels = torch.randint(0, 10, (B,), device=dev)
ts = mc.MCTrainStep(net, x, 1, train_size=50000.0, seed=1)
opt = torch.optim.Adam(net.parameters(), lr=1e-5)
for _ in range(3):
    ts(x, labels, 0.1); opt.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    ts(x, labels, 0.1); opt.step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 10 * 1e3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        ts(x, labels, 0.1); opt.step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows) / 5
n = sum(e.count for e in rows if e.device_time_total > 0) / 5
print(f"training step: wall {wall:.2f} ms; device kernels {tot / 1e3:.2f} ms in {n:.0f} launches per step")
for e in rows[:16]:
    print(f"{e.device_time_total / 5:9.1f} us  x{e.count / 5:5.1f}  {e.key[:100]}")
