"""CUDA-graph capture of a whole forward(+KL): the six layer kernels and the
interleaved aten activation/pool kernels become one graph launch, which is what
the problem size needs (the whole BBBAlexNet forward is ~10-20 us of
roofline time, i.e. the cost of its own kernel launches).

Noise under replay: kernel arguments are frozen at capture, so the Philox stream
is taken relative to a device scalar that a captured bbb_noise_advance kernel
moves forward at the head of every replay.
"""
from __future__ import annotations

import torch

from . import _lib
from . import functional as Fn

_STRIDE = 1 << 20      # stream ids one replay may consume (>= Bayesian layer calls per forward)


class GraphedForward:
    """logits, kl = GraphedForward(net, example_x)(x).  Replay r draws Philox streams
    first_stream + r*2^20 + (0, 1, 2, ...) -- reproducible from (seed, first_stream)."""

    def __init__(self, net, example_x: torch.Tensor, first_stream: int = 0, warmup: int = 2, post=None):
        """post(logits, kl) -> outputs is captured behind the forward (e.g. the multi-GPU combine with
        its NCCL all-reduce), so a whole step is one graph launch."""
        assert example_x.is_cuda
        self.net = net
        dev = example_x.device
        self.x = example_x.clone()
        self.base = torch.zeros(1, dtype=torch.int64, device=dev)
        self.first_stream = int(first_stream)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                with Fn.stream_base(self.base):
                    out = net(self.x)
                if post is not None:
                    post(*out)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph), torch.no_grad():
            Fn.noise_advance(self.base, _STRIDE)
            with Fn.stream_base(self.base):
                self.logits, self.kl = net(self.x)
            if post is not None:
                self.logits, self.kl = post(self.logits, self.kl)
        self.kernels_per_replay = _lib.launch_count() - n0     # engine kernels captured in the graph
        self.replays = 0
        self.reset(self.first_stream)

    def reset(self, first_stream: int = 0):
        """Next replay uses streams first_stream + (0, 1, ...)."""
        self.base.fill_(int(first_stream) - _STRIDE)

    def __call__(self, x: torch.Tensor | None = None, non_blocking: bool = True):
        if x is not None:
            self.x.copy_(x, non_blocking=non_blocking)
        self.graph.replay()
        self.replays += 1
        return self.logits, self.kl
