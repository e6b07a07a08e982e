"""Why does the end-to-end arm move between ~2 and ~4 M img/s?  Raw pinned H2D bandwidth, the NUMA node the pinned pages
really live on (move_pages query), and the e2e loop of bench.py with / without the NVML clock sampler thread."""
import ctypes as C
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import ClockSampler, build_net, pin_to_gpu_numa_node

numa = pin_to_gpu_numa_node(0) if os.environ.get("PROBE_PIN", "1") == "1" else None
torch.cuda.set_device(0)
dev = torch.device("cuda:0")
from pytorch_bayesiancnn_b200 import mc


def page_nodes(t, n=64):
    libc = C.CDLL("libc.so.6", use_errno=True)
    step = max(4096, (t.numel() * t.element_size() // n) // 4096 * 4096)
    pages = (C.c_void_p * n)(*[t.data_ptr() // 4096 * 4096 + i * step for i in range(n)])
    status = (C.c_int * n)()
    rc = libc.syscall(279, 0, C.c_ulong(n), pages, None, status, 0)
    return rc, sorted(set(status))


B = 512
x_host = [torch.randn(B, 3, 32, 32).pin_memory() for _ in range(4)]
print("numa pin:", numa, " OMP threads:", torch.get_num_threads(), " pinned pages on nodes:", page_nodes(x_host[0]), flush=True)
staging = [torch.empty(B, 3, 32, 32, device=dev) for _ in range(2)]
st = torch.cuda.Stream()
for rep in range(2):
    with torch.cuda.stream(st):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(100):
            staging[i & 1].copy_(x_host[i % 4], non_blocking=True)
        e1.record(st)
    torch.cuda.synchronize()
    print(f"raw H2D 6.29 MB x100: {100 * 6.291456e6 / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s", flush=True)

net = build_net("lrt", 10, dev, "bf16")
eng = mc.MCForward(net, staging[0], 1, seed=1, static_inputs=staging, overlap=True, inflight=int(os.environ.get("BBB_B200_MC_INFLIGHT", "4")))
out_host = torch.empty(B, 10).pin_memory()
kl_host = torch.empty(1).pin_memory()
main = torch.cuda.current_stream()
copy_stream = torch.cuda.Stream()
ready = [torch.cuda.Event() for _ in range(2)]
consumed = [torch.cuda.Event() for _ in range(2)]


def e2e_steps(nsteps):
    copy_stream.wait_stream(main)
    with torch.cuda.stream(copy_stream):
        staging[0].copy_(x_host[0], non_blocking=True)
        ready[0].record(copy_stream)
    for i in range(nsteps):
        s = i & 1
        if i + 1 < nsteps:
            with torch.cuda.stream(copy_stream):
                if i >= 1:
                    copy_stream.wait_event(consumed[s ^ 1])
                staging[s ^ 1].copy_(x_host[(i + 1) % 4], non_blocking=True)
                ready[s ^ 1].record(copy_stream)
        main.wait_event(ready[s])
        out = eng(slot=s)
        consumed[s] = eng.input_consumed()
        with torch.cuda.stream(eng.result_stream):
            out_host.copy_(out["log_outputs"], non_blocking=True)
            kl_host.copy_(out["kl"].reshape(1), non_blocking=True)
    main.wait_stream(eng.result_stream)


def timed(tag, nsteps=40, reps=7):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(main); e2e_steps(nsteps); e1.record(main)
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e1) * 1e3 / nsteps, t_enq * 1e6 / nsteps))
    print(f"{tag}: {statistics.median(t[0] for t in ts):.1f} us/step on the device ({B / statistics.median(t[0] for t in ts):.2f} M img/s), "
          f"host enqueue {statistics.median(t[1] for t in ts):.1f} us/step", flush=True)


timed("e2e, no sampler")
smp = ClockSampler(0); smp.start()
timed("e2e, NVML sampler thread running")
print(smp.stop())
timed("e2e, sampler stopped again")
