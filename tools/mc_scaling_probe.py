"""Per-rank view of the sharded MC step (run under torchrun): step period seen by every rank, and -- from the engine's
device timestamps (%globaltimer, per GPU) -- when the exchange kernel starts / has its inputs / ends relative to the end of
the layer chain, and when the NEXT step's first kernel starts.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mc_scaling_probe.py
"""
import ctypes as C
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from pytorch_bayesiancnn_b200 import _lib as L, mc
from bench import build_net, pin_to_gpu_numa_node

pin_to_gpu_numa_node(local)
overlap = os.environ.get("BBB_B200_MC_OVERLAP", "1") == "1"
B = 512
net = build_net("lrt", 10, dev, "bf16")
xs = [torch.randn(B, 3, 32, 32, device=dev) for _ in range(24)]
lib = C.CDLL(L.LIB_PATH)
lib.bbb_debug_set_timeline.argtypes = [C.c_void_p, C.c_int]
lib.bbb_debug_timeline_name.restype = C.c_char_p
lib.bbb_debug_timeline_name.argtypes = [C.c_int]

# (1) step period, uninstrumented
eng = mc.MCForward(net, xs[0], world, seed=1, static_inputs=xs, overlap=overlap, inflight=int(os.environ.get("BBB_B200_MC_INFLIGHT", "1")))
for k in range(10):
    eng(slot=k % 24)
per = []
for w in range(15):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(40):
        eng(slot=(w * 40 + k) % 24)
    eng.wait()
    e1.record()
    torch.cuda.synchronize()
    per.append(e0.elapsed_time(e1) * 1e3 / 40)
msg = [f"rank {rank}/{world} overlap={int(overlap)}: {statistics.median(per):.1f} us/step (min {min(per):.1f})"]
eng.close()

# (2) instrumented: one resident input, slots fixed at capture
CAP = 96
slots = torch.zeros(CAP, 4, dtype=torch.int64, device=dev)
lib.bbb_debug_set_timeline(C.c_void_p(slots.data_ptr()), CAP)
mtrace = torch.zeros(64, 8, dtype=torch.int64, device=dev)
lib.bbb_debug_set_mcx_trace.argtypes = [C.c_void_p]
lib.bbb_debug_set_mcx_trace(C.c_void_p(mtrace.data_ptr()))
eng = mc.MCForward(net, xs[0], world, seed=1, static_inputs=xs[:1], overlap=overlap)
lib.bbb_debug_set_mcx_trace(None)
n = lib.bbb_debug_timeline_count()
names = [lib.bbb_debug_timeline_name(k).decode() for k in range(n)]
lib.bbb_debug_set_timeline(None, 0)
torch.cuda.synchronize()
ex = [k for k in range(n) if names[k].startswith("mc_exchange")]
per_step = (n - 0) // (4 if overlap else 3)        # 2 eager warm-ups + 1 (2 with overlap) captured steps
cap0 = n - per_step * (2 if overlap else 1)        # first slot of the captured step (parity 0)
init = torch.tensor([[2 ** 62, 0, 2 ** 62, 0]] * CAP, dtype=torch.int64, device=dev)
rows, hs = [], []
for rep in range(60):
    if rep % 2 == 0:
        slots.copy_(init)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
    eng()
    if rep % 2 == 1:
        eng.wait()
        torch.cuda.synchronize()
        t = slots[:n].cpu()
        m = mtrace.cpu()
        if world > 1:
            hs.append([float((m[:, 1] - m[:, 0]).double().median()), float((m[:, 5] - m[:, 1]).double().median()), float(m[:, 0].max() - m[:, 0].min())])
        if overlap:
            a, b = cap0, cap0 + per_step       # parity-0 step then parity-1 step
        else:
            continue
        first0, x0 = a, a + per_step - 1
        fc_end0 = int(t[a:x0, 1].max())
        rows.append(dict(chain=fc_end0 - int(t[a:x0, 0].min()), x_start=int(t[x0, 0]) - fc_end0, x_dep=int(t[x0, 2]) - fc_end0,
                         x_end=int(t[x0, 1]) - fc_end0, next_start=int(t[b:b + per_step - 1, 0].min()) - fc_end0,
                         next_chain=int(t[b:b + per_step - 1, 1].max()) - int(t[b:b + per_step - 1, 0].min())))
if rows:
    med = {k: statistics.median(r[k] for r in rows[3:]) / 1e3 for k in rows[0]}
    msg.append("   (us, relative to the end of step t's last layer kernel) " + "  ".join(f"{k} {v:.1f}" for k, v in med.items()))
if hs:
    med = [statistics.median(h[i] for h in hs[3:]) / 1e3 for i in range(3)]
    msg.append("   exchange kernel, median CTA (us): partials+push %.1f  wait+finish %.1f ; CTA start spread %.1f" % tuple(med))
for r in range(world):
    if r == rank:
        print("\n".join(msg), flush=True)
    if world > 1:
        dist.barrier()
eng.close()
if world > 1:
    dist.destroy_process_group()
