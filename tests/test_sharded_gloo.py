"""World-size-2 gloo test of the sharded-MMD host logic (RowBlockPlan + collectives) on CPU.
The per-rank tile work is computed by the numpy oracle here (checker), standing in for the CUDA
kernel: what is under test is the partitioning, the all-gather order and the partial-sum reduce."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ns_local, nt_local, d, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import edrl_b200
    from edrl_b200.sharded import RowBlockPlan, TILE, gather_rows, reduce_partials
    from oracle import edrl_oracle as O
    g = torch.Generator().manual_seed(2000 + rank)
    x_loc = torch.randn(ns_local, d, generator=g, dtype=torch.float64)
    y_loc = torch.randn(nt_local, d, generator=g, dtype=torch.float64) * 1.25 + 0.1
    plan = RowBlockPlan(rank, world, ns_local, nt_local)
    x_all = gather_rows(x_loc)
    y_all = gather_rows(y_loc)
    assert torch.equal(x_all[rank * ns_local:(rank + 1) * ns_local], x_loc)
    z = torch.cat([x_all, y_all]).numpy()
    n = plan.n
    kmat = O.gaussian_kernel(x_all.numpy(), y_all.numpy())
    a = np.concatenate([np.full(plan.n_s, 1.0 / plan.n_s), np.full(plan.n_t, -1.0 / plan.n_t)])
    part = 0.0
    for (i, j) in plan.tiles():
        blk = (a[i * TILE:(i + 1) * TILE, None] * a[None, j * TILE:(j + 1) * TILE]
               * kmat[i * TILE:(i + 1) * TILE, j * TILE:(j + 1) * TILE]).sum()
        part += blk if i == j else 2.0 * blk
    partial = torch.tensor([part, 0.0], dtype=torch.float64)
    reduce_partials(partial)
    loss_ref, m, dx, dy = O.mk_mmd_grad(x_all.numpy(), y_all.numpy())
    assert np.isclose(abs(partial[0].item()), loss_ref, rtol=1e-10), (partial, loss_ref)
    # local gradient rows are exactly this rank's slices of the global gradient: no exchange needed
    r0, cnt = plan.source_rows()
    t0, tcnt = plan.target_rows()
    assert (r0, cnt) == (rank * ns_local, ns_local) and (t0, tcnt) == (plan.n_s + rank * nt_local, nt_local)
    dz = np.concatenate([dx, dy])
    np.save(os.path.join(out, f"g{rank}.npy"), np.concatenate([dz[r0:r0 + cnt], dz[t0:t0 + tcnt]]))
    # all tiles are covered exactly once across ranks
    counts = torch.tensor([len(plan.tiles())])
    dist.all_reduce(counts)
    assert counts.item() == plan.num_tiles()
    # the product path refuses CPU tensors even when a process group exists
    try:
        edrl_b200.sharded_MK_MMD(x_loc.float(), y_loc.float())
        raise AssertionError("expected RuntimeError")
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    dist.destroy_process_group()


def test_sharded_plan_and_collectives_world2(tmp_path):
    world, ns_local, nt_local, d = 2, 150, 170, 12
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ns_local, nt_local, d, str(tmp_path)), nprocs=world, join=True)
    g0 = np.load(tmp_path / "g0.npy")
    g1 = np.load(tmp_path / "g1.npy")
    assert g0.shape == (ns_local + nt_local, d) and g1.shape == g0.shape and not np.allclose(g0, g1)
