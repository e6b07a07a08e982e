"""Monte-Carlo sample sharding and the one exchange of the forward path.

The reference's ``num_ens`` loop (main_bayesian.py:46-53, validate :75-80;
uncertainty_estimation.py:70-78) runs S independent weight samples of the SAME batch and
combines them with logmeanexp of log-softmax.  Samples only differ in their noise, so
rank r of R takes the global sample ids {j : j mod R == r} (Philox stream ``j << 32``:
results do not depend on R) and ONE all-reduce of [3*B*C + 1] floats carries
sum_j softmax_j, sum_j softmax_j^2, sum_j logits_j and sum_j KL_j.

``forward_fn(x, sample_id) -> (logits [B,C], kl scalar)`` is whatever runs one sample
(the CUDA engine in production; tests drive the host logic with a CPU stand-in over gloo).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


def local_samples(num_ens: int, world: int, rank: int):
    """Global sample ids owned by `rank` (round-robin; C4: 25 samples over 8 ranks -> 4,3,3,...)."""
    return list(range(rank, num_ens, world))


def mc_forward(forward_fn: Callable, x: torch.Tensor, num_ens: int, group=None, want_uncertainty: bool = False):
    """Returns (log_outputs [B,C], kl) like main_bayesian.py:46-53, and optionally
    (pred, epistemic, aleatoric, entropy) like uncertainty_estimation.py:70-96."""
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    acc = None
    for j in local_samples(num_ens, world, rank):
        logits, kl = forward_fn(x, j)
        p = torch.softmax(logits.float(), dim=1)
        part = torch.cat([p.reshape(-1), (p * p).reshape(-1), logits.float().reshape(-1),
                          torch.as_tensor(kl, dtype=torch.float32, device=logits.device).reshape(1)])
        acc = part if acc is None else acc + part
        shape = logits.shape
    if acc is None:                                   # a rank with no sample (num_ens < world) still joins the collective
        probe, _ = forward_fn(x, 0)
        shape = probe.shape
        acc = torch.zeros(3 * probe.numel() + 1, dtype=torch.float32, device=probe.device)
    if distributed and world > 1:
        dist.all_reduce(acc, group=group)             # the ONE collective of the forward path
    n = shape[0] * shape[1]
    S = float(num_ens)
    p_bar = (acc[:n] / S).view(shape)
    log_outputs = torch.log(p_bar)                    # == logmeanexp_j log_softmax_j (utils.py:14-22)
    kl = acc[3 * n] / S                               # main_bayesian.py:51
    if not want_uncertainty:
        return log_outputs, kl
    p2 = (acc[n:2 * n] / S).view(shape)
    pred = (acc[2 * n:3 * n] / S).view(shape)
    epistemic = p2 - p_bar * p_bar                    # diag((p-pbar)^T (p-pbar))/T  (uncertainty_estimation.py:89-91)
    aleatoric = p_bar - p2                            # diag(diag(pbar) - p^T p / T)  (:94-95)
    entropy = -(p_bar * torch.log(p_bar.clamp_min(1e-38))).sum(1)      # H[pbar]; no reference
    return log_outputs, kl, (pred, epistemic, aleatoric, entropy)


def engine_forward_fn(net) -> Callable:
    """forward_fn for a net built on the engine: positions the Philox stream at sample j."""
    from . import functional as Fn

    def fn(x, j):
        Fn.begin_sample(j)
        with torch.no_grad():
            return net(x)
    return fn
