#!/usr/bin/env python
"""bench.py -- BBBAlexNet forward + KL images/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one synthetic batch: BBBAlexNet
(CIFAR-10 shape, 3x32x32, batch 512), ONE Monte-Carlo weight sample per GPU,
all six Bayesian layers + the model file's own activation/pool/flatten modules +
the summed KL scalar.  With N GPUs the num_ens MC loop (main_bayesian.py:46-49)
is the shard axis: rank r runs sample r of the SAME batch and one NCCL all-reduce
combines sum_j softmax_j and the KL -> weak scaling, value =
B * N * K / t.

Printed JSON (rank 0, one line): the driver contract + `roofline`, `cpu_baseline`,
`e2e`, `clocks`, `gpu_launches`, `per_layer`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

PRIORS = {"prior_mu": 0, "prior_sigma": 0.1, "posterior_mu_initial": (0, 0.1),
          "posterior_rho_initial": (-5, 0.1)}          # config_bayesian.py:4-9
METRIC = "BBBAlexNet fwd+KL images/sec"
L2_FLUSH_BYTES = 256 << 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------- #
# clocks
# --------------------------------------------------------------------------- #
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled every ~2 ms from a thread
    (the timed region lasts tens of milliseconds, too short for `nvidia-smi -lms`)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        self.index, self.sm, self.mask, self.max_mhz, self.err = index, [], 0, None, None
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    except Exception as e:          # keep sampling what we can
                        self.err = str(e)
                    time.sleep(0.002)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception as e:
            self.err = str(e)

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=1)
        reasons = sorted(n for n, bit in self.REASONS.items() if self.mask & bit)
        out = {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
               "reasons": reasons, "samples": len(self.sm)}
        if self.err:
            out["note"] = self.err
        return out


# --------------------------------------------------------------------------- #
# workload
# --------------------------------------------------------------------------- #
def layer_table(batch, classes=10):
    """Algorithmic FLOPs / bytes per Bayesian layer."""
    spec = [("conv1", 3, 32, 64, 11, 4, 5), ("conv2", 64, 4, 192, 5, 1, 2), ("conv3", 192, 2, 384, 3, 1, 1),
            ("conv4", 384, 2, 256, 3, 1, 1), ("conv5", 256, 2, 128, 3, 1, 1), ("classifier", 128, 1, classes, 1, 1, 0)]
    rows = []
    for name, cin, hin, cout, k, s, p in spec:
        ho = (hin + 2 * p - k) // s + 1
        K = cin * k * k
        rows.append({"name": name, "M": batch * ho * ho, "N": cout, "K": K,
                     "flops_mean": 2.0 * batch * ho * ho * cout * K,
                     "x_elems": batch * cin * hin * hin, "y_elems": batch * cout * ho * ho,
                     "params": cout * K + cout})
    return rows


def algorithmic(row, variant, act_bytes=4):
    v = 2.0 if variant == "lrt" else 1.0
    flops = row["flops_mean"] * v
    byts = row["x_elems"] * act_bytes + 2 * row["params"] * 4 + row["y_elems"] * act_bytes + 4
    return flops, byts


def build_net(variant, classes, device, math):
    import pytorch_bayesiancnn_b200 as bbb  # noqa: F401
    from pytorch_bayesiancnn_b200.models import BBBAlexNet
    torch.manual_seed(123)
    net = BBBAlexNet(classes, 3, PRIORS, variant, "softplus")
    with torch.no_grad():                       # identical params on every rank, drawn on the CPU generator
        g = torch.Generator().manual_seed(123)
        for name, p in net.named_parameters():
            mean = -5.0 if name.endswith("rho") else 0.0
            p.copy_(torch.empty(p.shape).normal_(mean, 0.1, generator=g))
    net = net.to(device).train()
    net.set_flag("math", math)
    return net


# --------------------------------------------------------------------------- #
# our arm
# --------------------------------------------------------------------------- #
def run_ours(args):
    import pytorch_bayesiancnn_b200 as bbb
    from pytorch_bayesiancnn_b200 import functional as Fn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        # The one collective is a 61 KB all-reduce (latency bound).  Every NCCL channel is a CTA that occupies an SM
        # while the forward kernels fill all 148 SMs at 1-2 CTAs each: keep the collective to two channels.
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "2")
        os.environ.setdefault("NCCL_MIN_NCHANNELS", "1")
        # NCCL prints its version banner on stdout at communicator creation: keep stdout clean for the one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            # (the collective itself runs on the process group's internal stream: make that one high priority too)
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
            dist.init_process_group("nccl", device_id=dev, pg_options=opts)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    B, C = args.batch, args.classes
    pk = peaks()

    net = build_net(args.variant, C, dev, args.math)
    gx = torch.Generator().manual_seed(0)
    n_inputs = 4                                 # pinned host batches (e2e arm)
    n_dev_inputs = 24                            # device-resident arm: 24 x 6.3 MB = 151 MB of inputs > 126 MB L2
    x_host = [torch.randn(B, 3, 32, 32, generator=gx).pin_memory() for _ in range(n_inputs)]
    x_dev = [torch.randn(B, 3, 32, 32, device=dev) for _ in range(n_dev_inputs)]
    bbb.manual_seed(2024)
    # rank r owns MC sample r: its Philox streams start at r << 32 (functional.begin_sample)
    # one captured forward per resident batch: the graph reads x_dev[k] in place (no staging copy in the step)
    graphed = bbb.GraphedForward(net, x_dev[0], first_stream=rank << 32, static_inputs=x_dev)
    # e2e arm: two more captures whose static inputs are the targets of the double-buffered host->device copies
    staging = [torch.empty_like(x_dev[0]) for _ in range(2)]
    graphed_e2e = bbb.GraphedForward(net, x_dev[0], first_stream=(rank << 32) + (1 << 30), static_inputs=staging)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    main = torch.cuda.current_stream(dev)

    # Multi-GPU combine: per-rank reduce of its sample by the engine's MC-combine kernel
    # (softmax / softmax^2 / logit sums), then ONE NCCL all-reduce of [3*B*C + 1] floats per step.  The collective
    # runs on its own stream, double buffered, so step i's all-reduce overlaps step i+1's forward.
    NSLOT = 2
    # High priority: the all-reduce's two CTAs must not queue behind the forward's kernels, which (chained by PDL)
    # hand every freed SM straight to the next GEMM; a late collective stalls the two-slot ring below.
    comm = torch.cuda.Stream(device=dev, priority=-1) if dist is not None else None
    comb = [torch.zeros(3 * B * C + 1, dtype=torch.float32, device=dev) for _ in range(NSLOT)]
    outs = [torch.empty(B, C, dtype=torch.float32, device=dev) for _ in range(NSLOT)]
    lo_scratch = torch.empty(B, C, dtype=torch.float32, device=dev)
    ev_packed = [torch.cuda.Event() for _ in range(NSLOT)]
    ev_reduced = [torch.cuda.Event() for _ in range(NSLOT)]
    extra_launches = [0]

    # The per-step combine is captured too (one graph per slot and stream), so a step costs the host three graph
    # launches and a few event calls instead of ~10 eager launches (which made N>1 host-bound).
    pack_graphs, comm_graphs = {}, []
    if dist is not None:
        from pytorch_bayesiancnn_b200 import _lib as L_

        def pack(k, out=None):
            logits, kl = out if out is not None else graphed.outputs[0]
            rc = L_.lib().bbb_mc_combine(Fn._ptr(logits), 1, B, C, Fn._ptr(lo_scratch), Fn._ptr(comb[k]), Fn._stream(dev))
            L_.check(rc, "bbb_mc_combine")
            comb[k][3 * B * C:].copy_(kl.reshape(1))

        def reduce_(k):
            dist.all_reduce(comb[k])
            torch.log(comb[k][:B * C] / world, out=outs[k].view(-1))

        for k in range(NSLOT):                       # warm up eagerly (NCCL communicator, allocator), then capture
            pack(k)
            with torch.cuda.stream(comm):
                comm.wait_stream(main)
                reduce_(k)
            main.wait_stream(comm)
        torch.cuda.synchronize(dev)
        dist.barrier()
        for gobj in (graphed, graphed_e2e):              # one pack graph per captured forward (its outputs are private)
            for slot, out in enumerate(gobj.outputs):
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    pack(slot % NSLOT, out)
                pack_graphs[(id(gobj), slot)] = g1
        for k in range(NSLOT):
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, stream=comm):
                reduce_(k)
            comm_graphs.append(g2)
        torch.cuda.synchronize(dev)

    def step(i, gobj, slot):
        """Step i on static input `slot` of `gobj` (slot parity == i parity, so comm slot k pairs with it)."""
        logits, kl = gobj(slot=slot)            # forward + KL: one graph launch
        if dist is None:
            return logits, kl
        k = i % NSLOT
        assert k == slot % NSLOT
        if i >= NSLOT:
            main.wait_event(ev_reduced[k])      # slot k is free again
        pack_graphs[(id(gobj), slot)].replay()  # engine MC-combine kernel + KL into comb[k]
        extra_launches[0] += 1
        ev_packed[k].record(main)
        with torch.cuda.stream(comm):
            comm.wait_event(ev_packed[k])
            comm_graphs[k].replay()             # the ONE all-reduce + log
            ev_reduced[k].record(comm)
        return outs[k], comb[k][3 * B * C]

    def join():
        if comm is not None:
            main.wait_stream(comm)

    def sync_all():
        join()
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident throughput: K steps back to back, inputs rotate through 151 MB (> L2), one event pair ----
    for i in range(args.warmup):
        step(i, graphed, i % n_dev_inputs)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0, x0 = graphed.replays, extra_launches[0]
    assert n_dev_inputs % NSLOT == 0
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    e_start.record(main)
    for i in range(args.steps):
        step(i, graphed, i % n_dev_inputs)      # replays the graph captured on resident batch i % 24
    join()
    e_stop.record(main)
    sync_all()
    wall = time.perf_counter() - wall0
    launches = (graphed.replays - l0) * graphed.kernels_per_replay + (extra_launches[0] - x0)   # engine kernels in the timed steps
    t_ms = torch.tensor([e_start.elapsed_time(e_stop)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms = float(t_ms.item())
    value = B * world * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API: pinned host input -> H2D -> forward -> D2H ----
    out_host = torch.empty(B, C, dtype=torch.float32).pin_memory()
    kl_host = torch.empty(1, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    def e2e_run(nsteps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record(main)
        copy_stream.wait_event(e0)
        with torch.cuda.stream(copy_stream):
            staging[0].copy_(x_host[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(nsteps):
            s = i & 1
            if i + 1 < nsteps:                  # prefetch the next batch while this one computes
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(consumed[s ^ 1])
                    staging[s ^ 1].copy_(x_host[(i + 1) % n_inputs], non_blocking=True)
                    ready[s ^ 1].record(copy_stream)
            main.wait_event(ready[s])
            lo, kl = step(i, graphed_e2e, s)
            consumed[s].record(main)
            if comm is not None:
                main.wait_event(ev_reduced[i % NSLOT])   # the combined result comes from the collective's stream
            out_host.copy_(lo, non_blocking=True)
            kl_host.copy_(kl.reshape(1), non_blocking=True)
        join()
        e1.record(main)
        sync_all()
        return e0.elapsed_time(e1)

    e2e_run(max(3, args.warmup))
    e_ms = torch.tensor([e2e_run(args.steps)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / (float(e_ms.item()) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    # ---- optional extra figure: S independent forwards in flight on S streams (not the headline) ----
    streams_fig = None
    if args.streams > 1 and dist is None:
        S = args.streams
        strs = [torch.cuda.Stream(device=dev) for _ in range(S)]
        gs = [bbb.GraphedForward(net, x_dev[0], first_stream=(rank << 32) + ((s + 1) << 26),
                                 static_inputs=x_dev[s::S], ws_slot=s + 1) for s in range(S)]

        def streams_run(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record(main)
            for st in strs:
                st.wait_event(e0)
            for i in range(n):
                s = i % S
                with torch.cuda.stream(strs[s]):
                    gs[s](slot=(i // S) % len(gs[s].inputs))
            for st in strs:
                main.wait_stream(st)
            e1.record(main)
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1)

        streams_run(max(args.warmup, 3) * S)
        t_s = streams_run(args.steps)
        streams_fig = {"streams": S, "ms_per_step": t_s / args.steps, "value": B * args.steps / (t_s * 1e-3),
                       "unit": "images/s",
                       "note": "separate figure, not the headline: independent forwards of different resident batches "
                               "overlap on S streams (per-stream layer workspaces and noise bases)"}

    # ---- per-layer kernel timing + roofline of the dominant kernel (rank 0) ----
    per_layer, roof = [], None
    if rank == 0:
        per_layer, roof = layer_rooflines(net, x_dev[0], args, pk, flush)

    mc_batched = None
    if rank == 0 and world == 1 and args.variant == "lrt" and args.mc_batch > 1:
        # S Monte-Carlo samples folded into ONE launch: for LRT the samples differ only in the per-activation noise, so
        # S samples of a batch == one batch of S*B rows (what uncertainty_estimation.py:38-41 does); KL computed once.
        S = args.mc_batch
        xb = x_dev[0].repeat(S, 1, 1, 1)
        gb = bbb.GraphedForward(net, xb, first_stream=1 << 40)
        for _ in range(3):
            gb()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            gb()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        mc_batched = {"mc_samples": S, "rows_per_launch": S * B, "ms_per_launch": ms,
                      "value": S * B / (ms * 1e-3), "unit": "sample-images/s",
                      "note": "separate figure, not the headline: S samples of the same 512 images per launch (LRT only)"}
        del gb, xb

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(args, seconds=args.cpu_seconds)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.math == "fp32" else "bf16 operands, f32 accumulate",
            "data": "synthetic (randn inputs, random-init params N(0,0.1)/rho N(-5,0.1))",
            "config": {"workload": f"BBBAlexNet-{C} CIFAR-10 shape 3x32x32, batch {B}, {args.variant} layers, "
                                   f"softplus, 1 MC sample per GPU per step (MC samples sharded over GPUs)",
                       "named_config": ("BASELINE.json configs[2] (BBBAlexNet CIFAR-10 batch 512 bf16, 10 MC samples, 1xB200, "
                                        "BBB_LRT): a step is ONE MC sample of the batch and the metric counts image-samples "
                                        "(B*S/t, SURVEY 8d), so the 10-sample loop has this same throughput; the single-launch "
                                        "S=10 figure is reported separately under mc_batched")
                       if (args.variant, args.math, B, C) == ("lrt", "bf16", 512, 10) else None,
                       "batch": B, "variant": args.variant, "math": args.math, "mc_samples_total": world,
                       "parallelism": f"mc{world}", "l2": "no flush: inputs rotate through 24 resident batches = 151 MB > 126 MB L2",
                       "launch": "CUDA graph replay of the full forward (noise advance, per-layer prep + GEMM kernels, KL sum); one captured graph per resident input batch, read in place"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * 32 * 32 * 4,
                    "d2h_bytes_per_step": B * C * 4 + 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "per_layer": per_layer,
            "streams": streams_fig,
            "cpu_baseline": cpu,
            "mc_batched": mc_batched,
            "wall_s_timed_loop": wall,
        }
        print(json.dumps(out), flush=True)
    if dist is not None:
        # leave without tearing NCCL down under live CUDA graphs (observed to hang at exit); every rank has
        # synchronised and rank 0 has printed
        sys.stdout.flush(); sys.stderr.flush()
        torch.cuda.synchronize(dev)
        dist.barrier()
        os._exit(0)


def layer_rooflines(net, x, args, pk, flush, reps=20):
    """Time each Bayesian layer call alone: the call (prep + GEMM kernels; fused chain
    step when the net runs fused) is captured in its own CUDA graph so host launch
    overhead stays out, replayed with CUDA events on the launching stream, L2 flushed
    (untimed) before every replay."""
    import pytorch_bayesiancnn_b200 as bbb
    from pytorch_bayesiancnn_b200 import fused
    rows = layer_table(args.batch, args.classes)
    steps = fused.plan(list(net.children()), tuple(x.shape)) if getattr(net, "fuse", True) else None
    calls = []
    with torch.no_grad():
        if steps is not None:
            cur, cur_sq, pitch = x.contiguous().float(), None, 0
            for i, st in enumerate(steps):
                nxt = steps[i + 1].layer if i + 1 < len(steps) else None
                calls.append((lambda st=st, nxt=nxt, a=cur, b=cur_sq, c=pitch, ph=0: fused.run_step(st, nxt, a, b, c, phase=ph)))
                cur, cur_sq, pitch = fused.run_step(st, nxt, cur, cur_sq, pitch)
        else:
            h = x
            for name, m in net.named_children():
                if hasattr(m, "W_mu"):
                    calls.append((lambda m=m, a=h.contiguous(): m(a)))
                h = m(h)
    def timed(call):
        with torch.no_grad():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                call(); call()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                call()
            times = []
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); gr.replay(); e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
        return statistics.median(times)

    from pytorch_bayesiancnn_b200 import _lib as L
    out = []
    for row, call in zip(rows, calls):
        ms = timed(call)
        # the two kernels of a fused-chain layer timed alone: parameter-only prep, and the GEMM kernel
        ms_prep = timed(lambda: call(ph=L.FUSED_PREP_ONLY)) if steps is not None else None
        ms_gemm = timed(lambda: call(ph=L.FUSED_SKIP_PREP)) if steps is not None else None
        fl, by = algorithmic(row, args.variant)
        t_tc = fl / (pk["tf_burst"] * 1e12)
        t_hbm = by / (pk["hbm_gbs"] * 1e9)
        bound = "tensor" if t_tc >= t_hbm else "hbm"
        out.append({"name": row["name"], "gemm": [row["M"], row["N"], row["K"]], "ms": ms,
                    "gflop": fl / 1e9, "mbytes": by / 1e6, "bound": bound,
                    "tflops": fl / (ms * 1e-3) / 1e12, "gbs": by / (ms * 1e-3) / 1e9,
                    "frac": max(t_tc, t_hbm) / (ms * 1e-3), "fused": steps is not None,
                    "ms_prep_kernel": ms_prep, "ms_gemm_kernel": ms_gemm})
    # dominant kernel = the longest single kernel: the GEMM kernel of a layer when the chain runs fused (the
    # layer's flops all execute there; its bytes are the layer's minus the fp32 mu/rho the prep kernel reads,
    # plus the bf16 operand tiles it reads instead), else the one fused fp32 layer kernel
    kt = (lambda r: r["ms_gemm_kernel"]) if steps is not None else (lambda r: r["ms"])
    top = max(out, key=kt)
    t_k = kt(top) * 1e-3
    if top["bound"] == "tensor":
        roof = {"kernel": top["name"] + (" GEMM kernel" if steps is not None else ""), "bound": "tensor",
                "achieved": top["gflop"] / 1e3 / t_k, "peak": pk["tf_burst"],
                "unit": "TFLOP/s", "frac": top["gflop"] / 1e3 / t_k / pk["tf_burst"], "traffic": None,
                "peak_source": pk["source"] + ", burst bf16 (kernel timed alone)",
                "kernel_us": t_k * 1e6, "layer_us_prep_plus_gemm": top["ms"] * 1e3, "layer_frac": top["frac"]}
    else:
        roof = {"kernel": top["name"] + (" GEMM kernel" if steps is not None else ""), "bound": "hbm",
                "achieved": top["mbytes"] / 1e3 / t_k, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": top["mbytes"] / 1e3 / t_k / pk["hbm_gbs"], "traffic": None,
                "peak_source": pk["source"],
                "kernel_us": t_k * 1e6, "layer_us_prep_plus_gemm": top["ms"] * 1e3, "layer_frac": top["frac"]}
    tp = os.path.join(ROOT, "profiles", "r1_ncu_full_gemm_final_traffic.json")      # dram__bytes_read+write of that kernel, one ncu --set full capture
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("variant") == args.variant and tj.get("batch") == args.batch and top["name"] in tj["layers"]:
            roof["traffic"] = tj["layers"][top["name"]]["dram_bytes"]
            roof["traffic_source"] = tj["source"]
    t_roof = sum(max(algorithmic(r, args.variant)[0] / (pk["tf_burst"] * 1e12),
                     algorithmic(r, args.variant)[1] / (pk["hbm_gbs"] * 1e9)) for r in rows)
    roof["net_t_roof_us"] = t_roof * 1e6
    roof["net_layer_kernels_us"] = sum(r["ms"] for r in out) * 1e3
    return out, roof


# --------------------------------------------------------------------------- #
# CPU reference arm (the oracle port of the reference's CPU path)
# --------------------------------------------------------------------------- #
def pick_threads(one):
    """The reference arm gets the thread count that serves it best on this host: oneDNN
    on 100+ threads is often slower than on a few dozen for convs this small."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (ncpu, ncpu // 2, 64, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        one()
        t0 = time.perf_counter(); one(); one()
        dt = (time.perf_counter() - t0) / 2
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_step_fn(args):
    from oracle import bbb_oracle as O               # bench's cpu_baseline leg may use the oracle
    params = O.init_params("alexnet", args.classes, 3, PRIORS, seed=123)
    x = torch.randn(args.batch, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    shapes = O.eps_shapes("alexnet", args.classes, 3, args.variant, args.batch)

    def one():
        with torch.no_grad():
            # the reference draws eps on the CPU generator inside every forward (BBB/BBBConv.py:63)
            eps = [torch.empty(s).normal_(0, 1) for s in shapes]
            logits, kl = O.net_forward("alexnet", params, x, eps, args.variant, "softplus", 0.0, 0.1, args.classes)
            return float(kl) + float(logits[0, 0])
    return one


def cpu_reference(args, seconds=10.0):
    """cpu_baseline leg: the reference arm in a fresh process (no CUDA context, no
    clock sampler competing for cores), bounded to ~`seconds` of CPU work."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "12", "--warmup", "3",
           "--variant", args.variant, "--batch", str(args.batch), "--classes", str(args.classes)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT).stdout.strip().splitlines()
        d = json.loads(out[-1])
        return d["cpu_baseline"]
    except Exception as e:                              # a baseline that cannot be taken is reported, not invented
        return {"value": None, "unit": "images/s", "cores": None, "kind": "port", "sample": f"failed: {e}"}


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    one = cpu_step_fn(args)
    steps = min(args.steps, 200)
    cores = pick_threads(one)
    for _ in range(args.warmup):
        one()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); one(); ts.append(time.perf_counter() - t0)
    dt = sum(ts)
    val = args.batch * steps / dt
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
           "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"BBBAlexNet-{args.classes} CIFAR-10 shape 3x32x32, batch {args.batch}, "
                                  f"{args.variant} layers, softplus, 1 MC sample per step", "batch": args.batch,
                      "variant": args.variant},
           "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                            "sample": f"{steps} forwards of the full batch-{args.batch} workload (median "
                                      f"{statistics.median(ts) * 1e3:.1f} ms, min {min(ts) * 1e3:.1f} ms); torch-CPU "
                                      f"restatement of the reference incl. its per-forward CPU eps draws; threads "
                                      f"picked as the fastest of a probe",
                            "cpu_model": cpu_model()},
           "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="lrt", choices=["lrt", "bbb"])
    ap.add_argument("--math", default=os.environ.get("BBB_B200_MATH", "bf16"), choices=["fp32", "bf16", "auto"])
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mc-batch", type=int, default=10, help="also report S MC samples folded into one launch (LRT; 0 = skip)")
    ap.add_argument("--streams", type=int, default=1,
                    help="extra figure (single GPU): that many forward graphs in flight on separate streams; 1 = off")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback); use --impl reference for the CPU arm")
        run_ours(args)


if __name__ == "__main__":
    main()
