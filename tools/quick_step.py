"""Event-timed MC step (bench.py's headline step, nothing else): python tools/quick_step.py [lrt|bbb] [batch] [num_ens]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pytorch_bayesiancnn_b200 as bbb
from pytorch_bayesiancnn_b200 import mc
from bench import build_net

variant = sys.argv[1] if len(sys.argv) > 1 else "lrt"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
net = build_net(variant, 10, dev, "bf16")
xs = [torch.randn(B, 3, 32, 32, device=dev) for _ in range(24)]   # 24 x 6 MB: inputs never L2-warm
bbb.manual_seed(1)
eng = mc.MCForward(net, xs[0], S, seed=1, static_inputs=xs, overlap=os.environ.get("BBB_B200_MC_OVERLAP", "0") == "1",
                   inflight=int(os.environ.get("BBB_B200_MC_INFLIGHT", "1")))
for k in range(10):
    eng(slot=k % 24)
win = []
for w in range(15):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(20):
        eng(slot=(w * 20 + k) % 24)
    eng.wait()
    e1.record()
    torch.cuda.synchronize()
    win.append(e0.elapsed_time(e1) * 1e3 / 20)
print(f"BBBAlexNet {variant} B={B} S={S}: {statistics.median(win):.1f} us/step (min {min(win):.1f}, max {max(win):.1f})")
