"""Time the fused sweep on ONE rank's rows of the configs[3] workload (N = 65536 per side, d = 1024) without the other
ranks: `edrl_mmd_forward_grad` + `edrl_mmd_apply_grad` over the row ranges rank `r` of `w` owns, and the whole workload on
one GPU through the public API.  For A/B runs of launch plans (EDRL_MMD_HYBRID, EDRL_MMD_QUAD, EDRL_MMD_SLABS)."""
import sys
import torch
sys.path.insert(0, ".")
import edrl_b200
from edrl_b200 import _lib, mmd as _mmd
from edrl_b200.sharded import RowBlockPlan
from edrl_b200.mmd import Workspace, NUM_STATS

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
rank = world // 2 - 1 if world > 1 else 0
g = torch.Generator(device="cuda").manual_seed(7)
x = torch.randn(N, d, device="cuda", generator=g)
y = torch.randn(N, d, device="cuda", generator=g) * 1.25 + 0.1
lib = _lib.load()
flags = _mmd._flags("tf32")
plan = RowBlockPlan(rank, world, N // world, N // world)
ws = Workspace(plan.n_s, plan.n_t, d, flags, x.device)
(r0, c0), (r1, c1) = plan.source_rows(), plan.target_rows()
slabs = _mmd._grad_slabs(plan.n_s, plan.n_t, d, flags, c0, c1, x.device.index)
u = torch.empty(slabs, c0 + c1, d, device="cuda")
dz = torch.empty(c0 + c1, d, device="cuda")
partial = torch.zeros(2, dtype=torch.float64, device="cuda")
loss = torch.empty((), device="cuda")
stats = torch.empty(NUM_STATS, device="cuda")
gout = torch.ones((), device="cuda")
stream = _lib.stream_and_device(x)


def shard_step():
    partial.zero_()
    _lib.check(lib.edrl_mmd_forward_grad(x.data_ptr(), y.data_ptr(), N, N, d, 2.0, 5, flags, r0, c0, r1, c1, 0, None, None,
                                         partial.data_ptr(), u.data_ptr(), ws.ptr, ws.nbytes, stream))
    _lib.check(lib.edrl_mmd_finalize(partial.data_ptr(), N, N, 2.0, 5, loss.data_ptr(), stats.data_ptr(), ws.ptr, ws.nbytes,
                                     stream))
    _lib.check(lib.edrl_mmd_apply_grad(N, N, d, flags, stats.data_ptr(), gout.data_ptr(), u.data_ptr(), r0, c0, r1, c1,
                                       dz.data_ptr(), ws.ptr, ws.nbytes, stream))


def timeit(fn, n):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


mn, md = timeit(shard_step, 6)
print(f"shard rank {rank}/{world} rows {c0}+{c1} slabs {slabs}: min {mn:.3f} ms median {md:.3f} ms  dz|sum| {dz.abs().sum().item():.6e}", flush=True)
del u, dz, ws
if world > 1 and (len(sys.argv) <= 4 or sys.argv[4] != "noanchor"):
    x.requires_grad_(True); y.requires_grad_(True)

    def full():
        x.grad = None; y.grad = None
        edrl_b200.MK_MMD(x, y).backward()
    mn, md = timeit(full, 3)
    print(f"whole workload, 1 GPU: min {mn:.3f} ms median {md:.3f} ms  loss-grad |sum| {x.grad.abs().sum().item():.6e}", flush=True)
