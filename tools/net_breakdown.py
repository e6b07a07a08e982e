"""Kernel-time breakdown of one eager forward (engine kernels + the aten glue between the layers):
    python tools/net_breakdown.py [3conv3fc|lenet|alexnet] [lrt|bbb] [batch] [inputs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from bench import build_net

net_type = sys.argv[1] if len(sys.argv) > 1 else "3conv3fc"
variant = sys.argv[2] if len(sys.argv) > 2 else "lrt"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
inputs = int(sys.argv[4]) if len(sys.argv) > 4 else (1 if net_type == "3conv3fc" else 3)
dev = torch.device("cuda:0")
net = build_net(variant, 10, dev, os.environ.get("BBB_B200_MATH", "bf16"), net_type, inputs)
x = torch.randn(B, inputs, 32, 32, device=dev)
with torch.no_grad():
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            net(x)
        torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"{net_type} {variant} B={B}: {tot / 5:.0f} us of kernels per forward")
for e in rows[:18]:
    print(f"{e.device_time_total / 5:9.1f} us  x{e.count // 5:3d}  {e.key[:110]}")
