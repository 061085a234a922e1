"""Row-block sharded MK_MMD across the GPUs of one node.

Every rank holds a slice ``X_p [N/P, d]``, ``Y_p [N/P, d]`` of the two sample sets.  One exchange
step: the feature rows are all-gathered (NCCL over NVLink), each rank evaluates its share of the
upper-triangular Gram tile list with the fused tcgen05 forward kernel, the two partial block sums
are all-reduced (2 doubles), and every rank finalises the same loss.  Backward needs no exchange:
rank p recomputes only the kernel rows of its own samples (G is symmetric).

The collective plumbing (``RowBlockPlan`` + ``gather_rows`` / ``reduce_partials``) is separate from
the kernel calls so the N>1 logic is testable with the gloo backend on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _lib
from . import mmd as _mmd
from .mmd import FLAG_TF32, NUM_STATS, Workspace, _flags

TILE = 128   # Gram tile edge of csrc/mmd.cu (BM = BN)


@dataclass(frozen=True)
class RowBlockPlan:
    """Who owns what in a P-way row-block sharded evaluation with equal local slices."""
    rank: int
    world: int
    ns_local: int
    nt_local: int

    @property
    def n_s(self) -> int:
        return self.ns_local * self.world

    @property
    def n_t(self) -> int:
        return self.nt_local * self.world

    @property
    def n(self) -> int:
        return self.n_s + self.n_t

    def source_rows(self):
        """Row range of this rank's source samples inside Z = [X_all; Y_all]."""
        return self.rank * self.ns_local, self.ns_local

    def target_rows(self):
        return self.n_s + self.rank * self.nt_local, self.nt_local

    def num_tiles(self) -> int:
        nb = (self.n + TILE - 1) // TILE
        return nb * (nb + 1) // 2

    def tiles(self):
        """Upper-triangular (I, J) tiles this rank evaluates: tile t = rank + world * q (the kernel's map)."""
        nb = (self.n + TILE - 1) // TILE
        out = []
        t = 0
        for i in range(nb):
            for j in range(i, nb):
                if t % self.world == self.rank:
                    out.append((i, j))
                t += 1
        return out


def gather_rows(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather equal-sized row slices into one [P * rows, d] tensor (rank order)."""
    world = dist.get_world_size(group)
    out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def reduce_partials(partial: torch.Tensor, group=None) -> torch.Tensor:
    dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


class _ShardedMKMMDFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, y_local, kernel_mul, kernel_num, flags, group):
        lib = _lib.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        plan = RowBlockPlan(rank, world, x_local.shape[0], y_local.shape[0])
        d = x_local.shape[1]
        x_all = gather_rows(x_local, group)
        y_all = gather_rows(y_local, group)
        stream = _lib.stream_and_device(x_all)
        ws = Workspace(plan.n_s, plan.n_t, d, flags, x_all.device)
        partial = torch.zeros(2, dtype=torch.float64, device=x_all.device)
        loss = torch.empty((), dtype=torch.float32, device=x_all.device)
        stats = torch.empty(NUM_STATS, dtype=torch.float32, device=x_all.device)
        ctx.U = None
        if _mmd._fused(flags, d) and any(ctx.needs_input_grad[:2]):
            # fused pass over this rank's rows (source rows, then target rows): partial sums + gradient part U
            (r0, c0), (r1, c1) = plan.source_rows(), plan.target_rows()
            slabs = _mmd._grad_slabs(plan.n_s, plan.n_t, d, flags, c0, c1, x_all.device.index)
            u = torch.empty(slabs, c0 + c1, d, dtype=torch.float32, device=x_all.device)
            _lib.check(lib.edrl_mmd_forward_grad(x_all.data_ptr(), y_all.data_ptr(), plan.n_s, plan.n_t, d,
                                                 float(kernel_mul), int(kernel_num), flags, r0, c0, r1, c1, 0, None,
                                                 None, partial.data_ptr(), u.data_ptr(), ws.ptr, ws.nbytes, stream))
            ctx.U = u
        else:
            # (a 1-rank group: the C entry point finalises by itself and wants the outputs; edrl_mmd_finalize below
            #  rewrites the same values from the partial sums)
            one = world == 1
            _lib.check(lib.edrl_mmd_forward(x_all.data_ptr(), y_all.data_ptr(), plan.n_s, plan.n_t, d,
                                            float(kernel_mul), int(kernel_num), flags, rank, world,
                                            loss.data_ptr() if one else None, stats.data_ptr() if one else None,
                                            partial.data_ptr(), ws.ptr, ws.nbytes, stream))
        reduce_partials(partial, group)
        _lib.check(lib.edrl_mmd_finalize(partial.data_ptr(), plan.n_s, plan.n_t, float(kernel_mul), int(kernel_num),
                                         loss.data_ptr(), stats.data_ptr(), ws.ptr, ws.nbytes, stream))
        ctx.plan, ctx.ws, ctx.stats, ctx.d = plan, ws, stats, d
        ctx.hyper = (float(kernel_mul), int(kernel_num), flags)
        ctx.keep = (x_all, y_all)      # the workspace holds the operands; keep the gathered rows alive with it
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        lib = _lib.load()
        plan, d = ctx.plan, ctx.d
        mul, num, flags = ctx.hyper
        g = grad_out.to(torch.float32).contiguous()
        stream = _lib.stream_and_device(g)
        if ctx.U is not None:
            (r0, c0), (r1, c1) = plan.source_rows(), plan.target_rows()
            dz = torch.empty_like(ctx.U[0])
            _lib.check(lib.edrl_mmd_apply_grad(plan.n_s, plan.n_t, d, flags, ctx.stats.data_ptr(), g.data_ptr(),
                                               ctx.U.data_ptr(), r0, c0, r1, c1, dz.data_ptr(), ctx.ws.ptr,
                                               ctx.ws.nbytes, stream))
            return (dz[:c0] if ctx.needs_input_grad[0] else None, dz[c0:] if ctx.needs_input_grad[1] else None,
                    None, None, None, None)
        outs = []
        for need, (r0, cnt) in ((ctx.needs_input_grad[0], plan.source_rows()),
                                (ctx.needs_input_grad[1], plan.target_rows())):
            if not need:
                outs.append(None)
                continue
            dz = torch.empty(cnt, d, dtype=torch.float32, device=g.device)
            _lib.check(lib.edrl_mmd_backward(plan.n_s, plan.n_t, d, mul, num, flags, ctx.stats.data_ptr(),
                                             g.data_ptr(), r0, cnt, dz.data_ptr(), ctx.ws.ptr, ctx.ws.nbytes, stream))
            outs.append(dz)
        return outs[0], outs[1], None, None, None, None


def sharded_MK_MMD(source_local, target_local, kernel_mul=2.0, kernel_num=5, precision=None, group=None):
    """MK_MMD over the union of all ranks' rows.  Returns the (replicated) loss of the *global*
    problem; its gradient w.r.t. the local slices is d loss / d (local rows), no averaging."""
    if not dist.is_initialized():
        raise RuntimeError("sharded_MK_MMD needs an initialised torch.distributed process group")
    _lib.require_cuda(source_local, target_local)
    if source_local.dim() != 2 or target_local.dim() != 2 or source_local.shape[1] != target_local.shape[1]:
        raise RuntimeError("sharded_MK_MMD expects [n_local, d] slices with the same d")
    return _ShardedMKMMDFunction.apply(source_local.to(torch.float32), target_local.to(torch.float32),
                                       kernel_mul, kernel_num, _flags(precision), group)
