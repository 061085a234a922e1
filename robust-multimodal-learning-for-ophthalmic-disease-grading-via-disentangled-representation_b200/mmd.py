"""Part A host side: ``MK_MMD`` / ``gaussian_kernel`` with the reference's signatures
(reference: code/MMD.py:3-74) on top of the sm_100a kernels in ``csrc/mmd.cu``.

``MK_MMD(source, target, kernel_mul=2.0, kernel_num=5)`` is a ``torch.autograd.Function`` whose
forward is one fused tcgen05 pass (the n x n kernel matrix never reaches HBM) and whose backward
recomputes the kernel tiles.  Arithmetic: Gram on the tensor cores in TF32 (``"tf32"``) or as a
hi/lo split with three MMAs per product (``"3xtf32"``, fp32-level accuracy; the same fused
forward+gradient sweep, one Gram pass per 512 feature columns); everything else fp32 with fp64
block accumulators.
"""
from __future__ import annotations

import functools
import os

import torch

from . import _lib

FLAG_TF32 = 0
FLAG_3XTF32 = 1
FLAG_TF32H = 2
# "tf32h": TF32 Gram like "tf32"; the G.Z product of the fused training sweep reads G and Z as power-of-two-scaled
# binary16 -- the same 11-bit significands as their TF32 roundings, half the bytes through shared memory.
FLAG_F16S = 4
# "f16s": like "tf32h", and the Gram of the fused sweep reads a binary16 copy of the same TF32-rounded operand (one
# power-of-two scale for the whole matrix): identical significands and exact products, both contractions on kind::f16.
_PRECISIONS = {"tf32": FLAG_TF32, "3xtf32": FLAG_3XTF32, "tf32h": FLAG_TF32H, "f16s": FLAG_F16S}


def _fused(flags, d):
    """Does a training step (gradients requested) of this precision and width take the fused forward+gradient sweep?"""
    return _FUSED        # (every precision mode; 3xTF32 beyond d = 768 sweeps the Gram once per 512-column pass)
_default_precision = os.environ.get("EDRL_MMD_PRECISION", "tf32").lower()
NUM_STATS = 8
# TF32 training steps use the fused pass (forward sums + gradient in one sweep over the Gram tiles);
# EDRL_MMD_FUSED=0 falls back to the separate forward and tile-recomputing backward kernels.
_FUSED = os.environ.get("EDRL_MMD_FUSED", "1") != "0"


def set_default_precision(name: str) -> None:
    """``"tf32"`` (default: one MMA per product, loss within 1e-3 / gradients within 2e-3 |g|_inf of the
    fp32 reference) or ``"3xtf32"`` (hi/lo split, three MMAs per product, fp32-level accuracy)."""
    global _default_precision
    if name.lower() not in _PRECISIONS:
        raise ValueError(f"unknown MMD precision {name!r}; expected one of {sorted(_PRECISIONS)}")
    _default_precision = name.lower()


def get_default_precision() -> str:
    return _default_precision


def _flags(precision) -> int:
    p = (_default_precision if precision is None else precision).lower()
    if p not in _PRECISIONS:
        raise ValueError(f"unknown MMD precision {precision!r}; expected one of {sorted(_PRECISIONS)}")
    return _PRECISIONS[p]


def _check_inputs(source, target):
    if source.dim() != 2 or target.dim() != 2:
        raise RuntimeError(f"MK_MMD expects [n, d] inputs, got {tuple(source.shape)} and {tuple(target.shape)}")
    if source.shape[1] != target.shape[1]:
        # torch.cat in the reference (code/MMD.py:21) raises RuntimeError for this
        raise RuntimeError(f"Sizes of tensors must match except in dimension 0: {tuple(source.shape)} vs "
                           f"{tuple(target.shape)}")
    _lib.require_cuda(source, target)
    if source.device != target.device:
        raise RuntimeError("source and target must be on the same CUDA device")


@functools.lru_cache(maxsize=256)
def _workspace_bytes(n_s, n_t, d, flags):
    return int(_lib.load().edrl_mmd_workspace_bytes(n_s, n_t, d, flags))


@functools.lru_cache(maxsize=256)
def _grad_slabs(n_s, n_t, d, flags, row_count, row_count2, device_index):
    # (host arithmetic on the shape and the device's SM count: cached per shape and device -- small problems are bound by
    #  host time, every ctypes call counts)
    return int(_lib.load().edrl_mmd_grad_slabs(n_s, n_t, d, flags, row_count, row_count2))


class Workspace:
    """1024-byte aligned scratch for one loss evaluation (shared by forward and backward)."""

    def __init__(self, n_s, n_t, d, flags, device):
        self.nbytes = _workspace_bytes(n_s, n_t, d, flags)
        if self.nbytes == 0:
            raise ValueError(f"MK_MMD: empty input (n_s={n_s} n_t={n_t} d={d})")
        self.buf = torch.empty(self.nbytes + 1024, dtype=torch.uint8, device=device)
        base = self.buf.data_ptr()
        self.ptr = (base + 1023) // 1024 * 1024


class _MKMMDFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, source, target, kernel_mul, kernel_num, flags):
        lib = _lib.load()
        x = source.contiguous()
        y = target.contiguous()
        n_s, d = x.shape
        n_t = y.shape[0]
        stream = _lib.stream_and_device(x)
        ws = Workspace(n_s, n_t, d, flags, x.device)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        stats = torch.empty(NUM_STATS, dtype=torch.float32, device=x.device)
        ctx.U = None
        if _fused(flags, d) and any(ctx.needs_input_grad[:2]):
            # one sweep over the Gram tiles: forward block sums + the bandwidth-independent gradient part U
            slabs = _grad_slabs(n_s, n_t, d, flags, n_s + n_t, 0, x.device.index)
            u = torch.empty(slabs, n_s + n_t, d, dtype=torch.float32, device=x.device)
            _lib.check(lib.edrl_mmd_forward_grad(x.data_ptr(), y.data_ptr(), n_s, n_t, d, float(kernel_mul),
                                                 int(kernel_num), flags, 0, n_s + n_t, 0, 0, 1, loss.data_ptr(),
                                                 stats.data_ptr(), None, u.data_ptr(), ws.ptr, ws.nbytes, stream))
            ctx.U = u
        else:
            _lib.check(lib.edrl_mmd_forward(x.data_ptr(), y.data_ptr(), n_s, n_t, d, float(kernel_mul),
                                            int(kernel_num), flags, 0, 1, loss.data_ptr(), stats.data_ptr(), None,
                                            ws.ptr, ws.nbytes, stream))
        ctx.ws = ws
        ctx.stats = stats
        ctx.shape = (n_s, n_t, d)
        ctx.hyper = (float(kernel_mul), int(kernel_num), flags)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        lib = _lib.load()
        n_s, n_t, d = ctx.shape
        mul, num, flags = ctx.hyper
        if not (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            return None, None, None, None, None
        g = grad_out.to(torch.float32).contiguous()
        stream = _lib.stream_and_device(g)
        if ctx.U is not None:
            dz = torch.empty_like(ctx.U[0])
            _lib.check(lib.edrl_mmd_apply_grad(n_s, n_t, d, flags, ctx.stats.data_ptr(), g.data_ptr(),
                                               ctx.U.data_ptr(), 0, n_s + n_t, 0, 0, dz.data_ptr(), ctx.ws.ptr,
                                               ctx.ws.nbytes, stream))
            return (dz[:n_s] if ctx.needs_input_grad[0] else None,
                    dz[n_s:] if ctx.needs_input_grad[1] else None, None, None, None)
        # rows of Z = [X; Y]: only compute the row range that needs a gradient
        r0 = 0 if ctx.needs_input_grad[0] else n_s
        r1 = n_s + n_t if ctx.needs_input_grad[1] else n_s
        dz = torch.empty(r1 - r0, d, dtype=torch.float32, device=g.device)
        _lib.check(lib.edrl_mmd_backward(n_s, n_t, d, mul, num, flags, ctx.stats.data_ptr(), g.data_ptr(), r0,
                                         r1 - r0, dz.data_ptr(), ctx.ws.ptr, ctx.ws.nbytes, stream))
        dx = dz[: n_s - r0] if ctx.needs_input_grad[0] else None
        dy = dz[n_s - r0:] if ctx.needs_input_grad[1] else None
        return dx, dy, None, None, None


def MK_MMD(source, target, kernel_mul=2.0, kernel_num=5, precision=None):
    """Multi-kernel MMD loss ``|XX + YY - XY - YX|`` -- drop-in for code/MMD.py:46-74.

    ``source`` [n_s, d] and ``target`` [n_t, d] are CUDA tensors; the result is a 0-d tensor on
    the same device, differentiable w.r.t. both (bandwidth not detached, like the reference).
    Non-fp32 inputs are computed in fp32 and the result is cast back.
    """
    _check_inputs(source, target)
    dt = source.dtype
    out = _MKMMDFunction.apply(source.to(torch.float32), target.to(torch.float32), kernel_mul, kernel_num,
                               _flags(precision))
    return out if dt == torch.float32 else out.to(dt)


def mk_mmd_with_stats(source, target, kernel_mul=2.0, kernel_num=5, precision=None):
    """Forward only: (loss, stats[8]) -- stats slots are the EDRL_MMD_STAT_* of the C header."""
    _check_inputs(source, target)
    lib = _lib.load()
    flags = _flags(precision)
    x = source.detach().to(torch.float32).contiguous()
    y = target.detach().to(torch.float32).contiguous()
    stream = _lib.stream_and_device(x)
    ws = Workspace(x.shape[0], y.shape[0], x.shape[1], flags, x.device)
    loss = torch.empty((), dtype=torch.float32, device=x.device)
    stats = torch.empty(NUM_STATS, dtype=torch.float32, device=x.device)
    _lib.check(lib.edrl_mmd_forward(x.data_ptr(), y.data_ptr(), x.shape[0], y.shape[0], x.shape[1],
                                    float(kernel_mul), int(kernel_num), flags, 0, 1, loss.data_ptr(),
                                    stats.data_ptr(), None, ws.ptr, ws.nbytes, stream))
    return loss, stats


class _GaussianKernelFunction(torch.autograd.Function):
    """Materialised kernel matrix (API completeness, code/MMD.py:3-44).  Forward is the tcgen05 tile
    kernel writing K; backward is composed from torch ops on the already-materialised n x n
    matrices -- this function is not on the hot path (MK_MMD never calls it)."""

    @staticmethod
    def forward(ctx, source, target, kernel_mul, kernel_num, flags):
        lib = _lib.load()
        x = source.contiguous()
        y = target.contiguous()
        n_s, d = x.shape
        n_t = y.shape[0]
        n = n_s + n_t
        stream = _lib.stream_and_device(x)
        ws = Workspace(n_s, n_t, d, flags, x.device)
        k = torch.empty(n, n, dtype=torch.float32, device=x.device)
        _lib.check(lib.edrl_mmd_kernel_matrix(x.data_ptr(), y.data_ptr(), n_s, n_t, d, float(kernel_mul),
                                              int(kernel_num), flags, k.data_ptr(), ws.ptr, ws.nbytes, stream))
        ctx.save_for_backward(x, y)
        ctx.hyper = (float(kernel_mul), int(kernel_num))
        return k

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gk):
        x, y = ctx.saved_tensors
        mul, num = ctx.hyper
        z = torch.cat([x, y], 0)
        z = z - z.mean(0, keepdim=True)
        n = z.shape[0]
        r = (z * z).sum(1, keepdim=True)
        l_raw = r + r.t() - 2.0 * (z @ z.t())
        l2 = l_raw.clamp(min=0.0)
        half = mul ** (num // 2)
        sigma0 = l2.sum() / (n * n - n) / half
        a = torch.zeros_like(l2)
        dsig = torch.zeros((), dtype=l2.dtype, device=l2.device)
        for k in range(num):
            sk = sigma0 * mul ** k
            e = torch.exp(-l2 / sk)
            a -= e / sk
            dsig += (gk * e * l2).sum() / (sk * sigma0)
        gl = (gk * a + dsig / ((n * n - n) * half)) * (l_raw >= 0)
        h = gl + gl.t()
        dz = 2.0 * (h.sum(1, keepdim=True) * z - h @ z)
        return dz[: x.shape[0]], dz[x.shape[0]:], None, None, None


def gaussian_kernel(source, target, kernel_mul=2.0, kernel_num=5, precision=None):
    """[n, n] summed multi-bandwidth kernel matrix -- drop-in for code/MMD.py:3-44."""
    _check_inputs(source, target)
    dt = source.dtype
    out = _GaussianKernelFunction.apply(source.to(torch.float32), target.to(torch.float32), kernel_mul, kernel_num,
                                        _flags(precision))
    return out if dt == torch.float32 else out.to(dt)


def compute_kl_divergence(p, m):
    """code/MMD.py:92-95 (pass-through torch; not on the hot path, call site commented out upstream)."""
    return torch.sum(p * torch.log(p / m), dim=1).mean()


def compute_js_divergence(p, q):
    """code/MMD.py:76-90 (pass-through torch)."""
    m = 0.5 * (p + q)
    return 0.5 * (compute_kl_divergence(p, m) + compute_kl_divergence(q, m))
