"""Debug aid: per-CTA clock64 checkpoints of tap_gemm_kernel for each fused AlexNet layer.
    python tools/trace_tapgemm.py [lrt|bbb]"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_bayesiancnn_b200 as bbb
from pytorch_bayesiancnn_b200 import fused, _lib as L
from pytorch_bayesiancnn_b200.models import BBBAlexNet

variant = sys.argv[1] if len(sys.argv) > 1 else "lrt"
priors = {"prior_mu": 0, "prior_sigma": 0.1, "posterior_mu_initial": (0, 0.1), "posterior_rho_initial": (-5, 0.1)}
dev = torch.device("cuda:0")
net = BBBAlexNet(10, 3, priors, variant, "softplus").to(dev).train()
net.set_flag("math", "bf16")
x = torch.randn(512, 3, 32, 32, device=dev)
steps = fused.plan(list(net.children()), tuple(x.shape))
lib = L.lib()
lib._handle  # noqa
fn = C.CDLL(L.LIB_PATH).bbb_debug_set_trace
fn.argtypes = [C.c_void_p]
trace = torch.zeros(4096 * 128, dtype=torch.int64, device=dev)
names = ["entry", "setup", "tma0|kb0", "full0|prod_end", "lastmma", "accum", "epi_end", "exit"]
with torch.no_grad():
    for rep in range(3):
        cur, sq, pitch = x, None, 0
        for i, st in enumerate(steps):
            nxt = steps[i + 1].layer if i + 1 < len(steps) else None
            trace.zero_()
            fn(C.c_void_p(trace.data_ptr()))
            torch.cuda.synchronize()
            cur, sq, pitch = fused.run_step(st, nxt, cur, sq, pitch)
            torch.cuda.synchronize()
            if rep == 2:
                t = trace.view(-1, 128).cpu()
                t = t[t[:, 0] != 0]
                rel = (t - t[:, :1]).double()
                print(f"layer {i}: {t.shape[0]} CTAs; mean cycles since entry:",
                      {n: int(rel[:, k].mean()) for k, n in enumerate(names)},
                      "max exit", int(rel[:, 7].max()))
                if i > 0:
                    one = rel[0]
                    nst = int((t[0, 8:40] != 0).sum())
                    print("   CTA0 per step: mma_full", [int(one[8 + k]) for k in range(nst)])
                    print("   CTA0 per step: tma_empty", [int(one[48 + k]) for k in range(nst)])
                    print("   CTA0 per step: tma_issued", [int(one[88 + k]) for k in range(nst)])
                    print("   CTA0 issue->full latency", [int(one[8 + k] - one[88 + k]) for k in range(nst)])
    fn(None)
