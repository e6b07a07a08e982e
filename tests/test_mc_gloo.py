"""world_size-2 gloo test of the MC-sample sharding / combine host logic (CPU).
The per-sample forward is a CPU stand-in (the oracle) -- only the exchange is under test."""
import os
import socket
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

from tests.util import CFG_PRIORS


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _forward_fn():
    from oracle import bbb_oracle as O
    params = O.init_params("lenet", 10, 3, CFG_PRIORS, seed=5)
    shapes = O.eps_shapes("lenet", 10, 3, "lrt", 6)

    def fn(x, j):
        eps = O.draw_eps_like_reference(shapes, seed=1000 + j)      # noise keyed by the GLOBAL sample id
        return O.net_forward("lenet", params, x, eps, "lrt", "softplus", 0.0, 0.1, 10)
    return fn


def _worker(rank, world, port, num_ens, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from pytorch_bayesiancnn_b200 import mc
    x = torch.randn(6, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    out, kl, unc = mc.mc_forward(_forward_fn(), x, num_ens, want_uncertainty=True)
    if rank == 0:
        torch.save((out, kl, unc), out_path)
    dist.destroy_process_group()


@pytest.mark.parametrize("num_ens", [1, 5])
def test_sharded_mc_equals_single_process(num_ens):
    from oracle import bbb_oracle as O
    from pytorch_bayesiancnn_b200 import mc
    torch.set_num_threads(1)
    x = torch.randn(6, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    fn = _forward_fn()
    logits = [fn(x, j) for j in range(num_ens)]
    ref = O.mc_combine([l for l, _ in logits])                         # main_bayesian.py:46-53 restated
    pred, epi, ale, ent = O.uncertainty([l for l, _ in logits])
    single, kl1 = mc.mc_forward(fn, x, num_ens)
    assert torch.allclose(single, ref, atol=1e-5)
    ctx = mp.get_context("spawn")
    port = _free_port()
    out_path = os.path.join(tempfile.mkdtemp(), "rank0.pt")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_ens, out_path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    out, kl, unc = torch.load(out_path)
    assert torch.allclose(out, ref, atol=1e-5)
    assert abs(float(kl) - float(logits[0][1])) <= 1e-6 * abs(float(kl))   # KL identical on every sample
    assert torch.allclose(unc[0].double(), pred, atol=1e-5)
    assert torch.allclose(unc[1].double(), epi, atol=1e-6) and torch.allclose(unc[2].double(), ale, atol=1e-6)
    assert torch.allclose(unc[3].double(), ent, atol=1e-5)


def test_local_samples_partition():
    from pytorch_bayesiancnn_b200 import mc
    assert [len(mc.local_samples(25, 8, r)) for r in range(8)] == [4, 3, 3, 3, 3, 3, 3, 3]     # C4
    assert [len(mc.local_samples(100, 8, r)) for r in range(8)] == [13, 13, 13, 13, 12, 12, 12, 12]  # C5
    assert sorted(sum((mc.local_samples(7, 3, r) for r in range(3)), [])) == list(range(7))
