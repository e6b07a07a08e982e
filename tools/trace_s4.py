"""clock64 checkpoints of conv_s4_kernel (first layer of the fused chain):  python tools/trace_s4.py [lrt|bbb]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_bayesiancnn_b200 as bbb
from pytorch_bayesiancnn_b200 import fused, _lib as L
from bench import build_net
variant = sys.argv[1] if len(sys.argv) > 1 else "lrt"
dev = torch.device("cuda:0")
net = build_net(variant, 10, dev, "bf16")
xs = [torch.randn(512, 3, 32, 32, device=dev) for _ in range(30)]     # > L2: every run reads a cold batch
steps = fused.plan(list(net.children()), tuple(xs[0].shape))
fn = C.CDLL(L.LIB_PATH).bbb_debug_set_trace
fn.argtypes = [C.c_void_p]
trace = torch.zeros(4096 * 128, dtype=torch.int64, device=dev)
names = ["entry", "setup", "staged", "noise", "mma_issued", "accum", "epi_end", "exit"]
with torch.no_grad():
    for rep in range(30):
        trace.zero_()
        fn(C.c_void_p(trace.data_ptr()))
        torch.cuda.synchronize()
        fused.run_step(steps[0], steps[1].layer, xs[rep], None, 0)
        torch.cuda.synchronize()
    t = trace.view(-1, 128).cpu()
    t = t[t[:, 0] != 0]
    rel = (t - t[:, :1]).double()
    print(f"{t.shape[0]} CTAs; mean cycles since entry:", {n: int(rel[:, k].mean()) for k, n in enumerate(names)}, "max exit", int(rel[:, 7].max()))
    one = rel[0]
    print("CTA0 full[r] passed:", [int(one[8 + r]) for r in range(11)])
    print("CTA0 row r issued  :", [int(one[24 + r]) for r in range(11)])
    print("staging checkpoints (mean): zero-fill done, then per batch [loads issued, noise slice done, stored]:", [int(rel[:, 40 + k].mean()) for k in range(7)])
    t0 = t[:, 0].min()
    print("CTA entry spread (cycles):", int((t[:, 0] - t0).double().mean()), int((t[:, 0] - t0).max()), " last exit:", int((t[:, 7] - t0).max()))
fn(None)
