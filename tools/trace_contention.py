"""Experiment: does tap_gemm's per-step time depend on how many CTAs share the same A tile?
conv4 of AlexNet with Cout in {64, 256, 1024}: n-tiles per m-tile = 4, 16, 64."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_bayesiancnn_b200 import fused, _lib as L, models as M
variant = sys.argv[1] if len(sys.argv) > 1 else "bbb"
priors = {"prior_mu": 0, "prior_sigma": 0.1, "posterior_mu_initial": (0, 0.1), "posterior_rho_initial": (-5, 0.1)}
dev = torch.device("cuda:0")
fn = C.CDLL(L.LIB_PATH).bbb_debug_set_trace
fn.argtypes = [C.c_void_p]
trace = torch.zeros(8192 * 128, dtype=torch.int64, device=dev)
for cout4 in (64, 256, 1024):
    arch = list(M._ARCH["alexnet"])
    arch[8] = ("conv4", "c", (cout4, 3, 1, 1))
    M._ARCH["tmp"] = tuple(arch)
    class Net(M._TableNet):
        _key = "tmp"
    net = Net(10, 3, priors, variant, "softplus").to(dev).train()
    net.set_flag("math", "bf16")
    x = torch.randn(512, 3, 32, 32, device=dev)
    steps = fused.plan(list(net.children()), tuple(x.shape))
    with torch.no_grad():
        for rep in range(3):
            cur, sq, pitch = x, None, 0
            for i, st in enumerate(steps):
                nxt = steps[i + 1].layer if i + 1 < len(steps) else None
                trace.zero_(); fn(C.c_void_p(trace.data_ptr()) if i == 3 else None); torch.cuda.synchronize()
                cur, sq, pitch = fused.run_step(st, nxt, cur, sq, pitch)
                torch.cuda.synchronize()
                if i == 3 and rep == 2:
                    t = trace.view(-1, 128).cpu(); t = t[t[:, 0] != 0]; rel = (t - t[:, :1]).double()
                    full = rel[:, 8:32]
                    print(f"cout={cout4}: {t.shape[0]} CTAs, mean step interval (cycles):",
                          int((full[:, 23] - full[:, 3]).mean() / 20), " exit", int(rel[:, 7].mean()))
fn(None)
