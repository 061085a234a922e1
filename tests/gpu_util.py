"""Helpers shared by the GPU parity tests (not collected by pytest)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def have_gpu() -> bool:
    return torch.cuda.is_available()


def tf32_round(x: np.ndarray) -> np.ndarray:
    """cvt.rna.tf32.f32: round to nearest (ties away) onto a 10-bit mantissa."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).copy()
    u = (u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return u.view(np.float32)


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype)


class RawMMD:
    """Direct C-ABI calls (tests go through the same entry points the host package uses)."""

    def __init__(self):
        import edrl_b200
        self.pkg = edrl_b200
        self.lib = edrl_b200._lib.load()
        self.L = edrl_b200._lib

    def workspace(self, n_s, n_t, d, flags):
        from edrl_b200.mmd import Workspace
        return Workspace(n_s, n_t, d, flags, torch.device("cuda", torch.cuda.current_device()))

    def kernel_matrix(self, x, y, mul=2.0, num=5, flags=0):
        n = x.shape[0] + y.shape[0]
        ws = self.workspace(x.shape[0], y.shape[0], x.shape[1], flags & 0xff)
        k = torch.empty(n, n, device="cuda")
        st = self.L.stream_and_device(x)
        self.L.check(self.lib.edrl_mmd_kernel_matrix(x.data_ptr(), y.data_ptr(), x.shape[0], y.shape[0], x.shape[1],
                                                     mul, num, flags, k.data_ptr(), ws.ptr, ws.nbytes, st))
        torch.cuda.synchronize()
        return k

    def forward(self, x, y, mul=2.0, num=5, flags=0, tile_rank=0, tile_world=1, ws=None):
        ws = ws or self.workspace(x.shape[0], y.shape[0], x.shape[1], flags)
        loss = torch.zeros((), device="cuda")
        stats = torch.zeros(8, device="cuda")
        partial = torch.zeros(2, dtype=torch.float64, device="cuda")
        st = self.L.stream_and_device(x)
        self.L.check(self.lib.edrl_mmd_forward(x.data_ptr(), y.data_ptr(), x.shape[0], y.shape[0], x.shape[1], mul,
                                               num, flags, tile_rank, tile_world, loss.data_ptr(), stats.data_ptr(),
                                               partial.data_ptr(), ws.ptr, ws.nbytes, st))
        return loss, stats, partial, ws

    def finalize(self, partial, n_s, n_t, ws, mul=2.0, num=5):
        loss = torch.zeros((), device="cuda")
        stats = torch.zeros(8, device="cuda")
        st = self.L.stream_and_device(partial)
        self.L.check(self.lib.edrl_mmd_finalize(partial.data_ptr(), n_s, n_t, mul, num, loss.data_ptr(),
                                                stats.data_ptr(), ws.ptr, ws.nbytes, st))
        return loss, stats

    def backward(self, n_s, n_t, d, stats, ws, row_begin, row_count, grad_out=1.0, mul=2.0, num=5, flags=0):
        g = torch.full((), float(grad_out), device="cuda")
        dz = torch.empty(row_count, d, device="cuda")
        st = self.L.stream_and_device(dz)
        self.L.check(self.lib.edrl_mmd_backward(n_s, n_t, d, mul, num, flags, stats.data_ptr(), g.data_ptr(),
                                                row_begin, row_count, dz.data_ptr(), ws.ptr, ws.nbytes, st))
        return dz


def raw_forward_grad(raw, x, y, r0, c0, r1=0, c1=0, finalize=1, mul=2.0, num=5, ws=None, flags=0):
    """edrl_mmd_forward_grad through the C-ABI: returns (loss, stats, partial, U, ws)."""
    ws = ws or raw.workspace(x.shape[0], y.shape[0], x.shape[1], flags)
    loss = torch.zeros((), device="cuda")
    stats = torch.zeros(8, device="cuda")
    partial = torch.zeros(2, dtype=torch.float64, device="cuda")
    slabs = int(raw.lib.edrl_mmd_grad_slabs(x.shape[0], y.shape[0], x.shape[1], flags, c0, c1))
    u = torch.empty(slabs, c0 + c1, x.shape[1], device="cuda")
    st = raw.L.stream_and_device(x)
    raw.L.check(raw.lib.edrl_mmd_forward_grad(x.data_ptr(), y.data_ptr(), x.shape[0], y.shape[0], x.shape[1], mul, num,
                                              flags, r0, c0, r1, c1, finalize, loss.data_ptr(), stats.data_ptr(),
                                              partial.data_ptr(), u.data_ptr(), ws.ptr, ws.nbytes, st))
    return loss, stats, partial, u, ws


def raw_apply_grad(raw, n_s, n_t, d, stats, u, ws, r0, c0, r1=0, c1=0, grad_out=1.0, flags=0):
    g = torch.full((), float(grad_out), device="cuda")
    dz = torch.empty_like(u[0])
    st = raw.L.stream_and_device(u)
    raw.L.check(raw.lib.edrl_mmd_apply_grad(n_s, n_t, d, flags, stats.data_ptr(), g.data_ptr(), u.data_ptr(), r0, c0, r1, c1,
                                            dz.data_ptr(), ws.ptr, ws.nbytes, st))
    return dz
