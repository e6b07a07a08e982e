"""Where does a graph-replayed BBBAlexNet forward spend its time?

Every engine kernel of the fused chain records [first CTA entry, last CTA exit] in %globaltimer ns into
debug slots fixed at capture time (bbb_debug_set_timeline); this tool replays the captured forward and
prints the median timeline relative to the first kernel's start, next to the event-timed step.

    python tools/timeline.py [lrt|bbb] [batch]
"""
import ctypes as C
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pytorch_bayesiancnn_b200 as bbb
from pytorch_bayesiancnn_b200 import _lib as L
from bench import build_net

variant = sys.argv[1] if len(sys.argv) > 1 else "lrt"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
lib = C.CDLL(L.LIB_PATH)
lib.bbb_debug_set_timeline.argtypes = [C.c_void_p, C.c_int]
lib.bbb_debug_timeline_name.restype = C.c_char_p
lib.bbb_debug_timeline_name.argtypes = [C.c_int]

net = build_net(variant, 10, dev, "bf16")
xs = [torch.randn(B, 3, 32, 32, device=dev) for _ in range(24)]
bbb.manual_seed(1)
MC = os.environ.get("TIMELINE_MC", "1") != "0"                  # the public MC step (mc.MCForward) instead of the bare forward
from pytorch_bayesiancnn_b200 import mc
if MC:
    mc.MCForward(net, xs[0], 1, seed=1)                          # warm-up + plan/workspace creation, untraced
else:
    bbb.GraphedForward(net, xs[0])
CAP = 64
slots = torch.zeros(CAP, 4, dtype=torch.int64, device=dev)
lib.bbb_debug_set_timeline(C.c_void_p(slots.data_ptr()), CAP)
if MC:
    class _G:                                                    # capture only (the eager warm-up steps inside consume slots too:
        pass                                                     # the LAST launches belong to the captured graph)
    eng = mc.MCForward(net, xs[0], 1, seed=1, static_inputs=xs[:1])
    g = _G(); g.inputs = eng.inputs; g.__class__.__call__ = lambda self: eng()
else:
    g = bbb.GraphedForward(net, xs[0], warmup=0, static_inputs=xs[:1])   # capture only: slot k <-> k-th instrumented launch
n = lib.bbb_debug_timeline_count()
names = [lib.bbb_debug_timeline_name(k).decode() for k in range(n)]
lib.bbb_debug_set_timeline(None, 0)

init = torch.tensor([[2 ** 62, 0, 2 ** 62, 0]] * CAP, dtype=torch.int64, device=dev)
runs, ev_us = [], []
for rep in range(30):
    g.inputs[0].copy_(xs[rep % 24])                              # fresh (L2-cold) input each replay, like bench.py
    slots.copy_(init)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g(); e1.record()
    torch.cuda.synchronize()
    ev_us.append(e0.elapsed_time(e1) * 1e3)
    t = slots[:n].cpu()
    t0 = int(t[:, 0].min())
    runs.append(((t[:, 0] - t0).tolist(), (t[:, 1] - t0).tolist(), (torch.minimum(t[:, 2], t[:, 1]) - t0).tolist()))
print(f"BBBAlexNet {variant} B={B}: event-timed replay median {statistics.median(ev_us):.1f} us "
      f"(min {min(ev_us):.1f}); {n} instrumented launches per replay")
med = lambda k, j: statistics.median(r[j][k] for r in runs[5:]) / 1e3
order = sorted(range(n), key=lambda k: med(k, 0))
print(f"{'kernel':44s} {'start us':>9s} {'end us':>9s} {'dur us':>8s} {'deps ok':>8s} {'work us':>8s}")
order = [k for k in order if med(k, 1) > 0 and med(k, 0) < 1e6]     # slots the replay really wrote
for k in order:
    dep = med(k, 2) if med(k, 2) < 1e6 else med(k, 0)
    print(f"{names[k]:44s} {med(k, 0):9.1f} {med(k, 1):9.1f} {med(k, 1) - med(k, 0):8.1f} {dep:8.1f} {med(k, 1) - dep:8.1f}")
gemm = [k for k in order if "gemm" in names[k] or "conv_s4 " in names[k]]
print("GEMM critical path: " + "  ".join(
    f"[{names[k].split()[0]} {med(k, 1) - med(k, 0):.1f}]" + (f" gap {med(gemm[i + 1], 0) - med(k, 1):.1f}" if i + 1 < len(gemm) else "")
    for i, k in enumerate(gemm)))
print(f"first kernel start -> last GEMM end: {med(gemm[-1], 1):.1f} us; sum of GEMM durations "
      f"{sum(med(k, 1) - med(k, 0) for k in gemm):.1f} us; first GEMM starts at {med(gemm[0], 0):.1f} us")
