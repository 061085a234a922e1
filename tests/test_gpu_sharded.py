"""Row-block sharded MK_MMD over NCCL.  Each rank holds a slice, the loss must equal the single-device loss of the
gathered problem and each rank's gradient must be its rows of the global gradient (numpy oracle, fp64).  The world-size-1
cases run on any GPU box (a 1-rank process group: `torchrun --nproc 1`, or evaluation under no_grad); the others use
every GPU the box has, up to 8, and are skipped below 2."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpu_util import have_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")]
need2 = pytest.mark.skipif(not have_gpu() or torch.cuda.device_count() < 2, reason="needs >= 2 CUDA devices")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nl, d, prec):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import edrl_b200
    from oracle import edrl_oracle as O
    xs, ys = [], []
    for r in range(world):
        g = torch.Generator().manual_seed(2000 + r)
        xs.append(torch.randn(nl, d, generator=g, dtype=torch.float64))
        ys.append(torch.randn(nl, d, generator=g, dtype=torch.float64) * 1.25 + 0.1)
    x = xs[rank].float().cuda().requires_grad_(True)
    y = ys[rank].float().cuda().requires_grad_(True)
    with torch.no_grad():                                       # loss-only evaluation: the tile-sharded forward kernel
        l0 = edrl_b200.sharded_MK_MMD(x, y, precision=prec)
    loss = edrl_b200.sharded_MK_MMD(x, y, precision=prec)
    (2.0 * loss).backward()
    xa, ya = torch.cat(xs).numpy(), torch.cat(ys).numpy()
    ref, _, dx, dy = O.mk_mmd_grad(xa, ya, grad_out=2.0)
    ltol, gtol = (1e-4, 1e-4) if prec == "3xtf32" else (1e-3, 2e-3)
    assert np.isclose(loss.item(), ref, rtol=ltol, atol=1e-6), (loss.item(), ref)
    assert np.isclose(l0.item(), ref, rtol=ltol, atol=1e-6), (l0.item(), ref)
    gmax = max(np.abs(dx).max(), np.abs(dy).max())
    assert np.abs(x.grad.cpu().numpy() - dx[rank * nl:(rank + 1) * nl]).max() <= gtol * gmax
    assert np.abs(y.grad.cpu().numpy() - dy[rank * nl:(rank + 1) * nl]).max() <= gtol * gmax
    # and against the single-device kernel on the gathered problem
    single = edrl_b200.MK_MMD(torch.tensor(xa).float().cuda(), torch.tensor(ya).float().cuda(), precision=prec)
    assert np.isclose(single.item(), loss.item(), rtol=1e-5)
    dist.destroy_process_group()


@need2
@pytest.mark.parametrize("prec", ["tf32", "3xtf32", "f16s"])
def test_sharded_mk_mmd_nccl(prec):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 320, 96, prec), nprocs=world, join=True)


@need2
@pytest.mark.parametrize("prec", ["tf32", "f16s"])
def test_sharded_mk_mmd_nccl_quad_kernel_all_gpus(prec):
    """d = 1100 takes the 4-CTA quad sweep (the kernel behind BASELINE configs[3]); every GPU of the box, up to 8."""
    world = min(torch.cuda.device_count(), 8)
    mp.spawn(_worker, args=(world, _free_port(), 200, 1100, prec), nprocs=world, join=True)


@pytest.mark.parametrize("prec", ["tf32", "3xtf32"])
@pytest.mark.parametrize("d", [96, 1100])
def test_sharded_mk_mmd_world_size_one(prec, d):
    """A 1-rank process group (fused and non-fused precisions, with and without gradients) is the plain problem."""
    mp.spawn(_worker, args=(1, _free_port(), 300, d, prec), nprocs=1, join=True)
