"""DILR Barlow-Twins cross-correlation loss on the sm_100a kernels of ``csrc/dilr.cu``.

``bt_loss_cross(self, z1, z2, common_dim)`` has the signature and return tuple of the reference method
``DILR.bt_loss_cross`` (code/fusion_net.py:656-677) and reads the same state from ``self`` -- ``self.bn1`` / ``self.bn2``
(``nn.BatchNorm1d(2048, affine=False)``: training flag, eps, momentum, running statistics, ``num_batches_tracked``) and
``self.args.batch_size`` -- so that the drop-in binds it over the reference class
(``<pkg>/dropin/fusion_net.py``: ``DILR.bt_loss_cross = edrl_b200.dilr.bt_loss_cross``).  The 2048 x 2048 correlation
matrix is never formed; forward is 2 kernels, backward 2 (the reference: ~30 launches each way).
"""
from __future__ import annotations

import functools

import torch

from . import _lib


@functools.lru_cache(maxsize=64)
def _ws_bytes(B, D):
    return int(_lib.load().edrl_dilr_workspace_bytes(B, D))


class _BTLossCross(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, common_dim, batch_size, eps, training, momentum, rm1, rv1, rm2, rv2):
        lib = _lib.load()
        a = z1.contiguous()
        b = z2.contiguous()
        B, D = a.shape
        ws = torch.empty(_ws_bytes(B, D) + 256, dtype=torch.uint8, device=a.device)
        ptr = (ws.data_ptr() + 255) // 256 * 256
        out = torch.empty(6, dtype=torch.float32, device=a.device)
        st = _lib.stream_and_device(a)
        _lib.check(lib.edrl_dilr_bt_loss_fwd(a.data_ptr(), b.data_ptr(), B, D, int(common_dim), int(batch_size), float(eps),
                                             int(training), float(momentum), _lib.ptr(rm1), _lib.ptr(rv1), _lib.ptr(rm2),
                                             _lib.ptr(rv2), out.data_ptr(), ptr, _ws_bytes(B, D), st))
        ctx.ws, ctx.ptr = ws, ptr
        ctx.cfg = (B, D, int(common_dim), int(batch_size), int(training))
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        lib = _lib.load()
        B, D, dc, bs, training = ctx.cfg
        g = gout.to(torch.float32).contiguous()
        dz1 = torch.empty(B, D, dtype=torch.float32, device=g.device) if ctx.needs_input_grad[0] else None
        dz2 = torch.empty(B, D, dtype=torch.float32, device=g.device) if ctx.needs_input_grad[1] else None
        if dz1 is None and dz2 is None:
            return (None,) * 11
        st = _lib.stream_and_device(g)
        _lib.check(lib.edrl_dilr_bt_loss_bwd(B, D, dc, bs, training, g.data_ptr(), _lib.ptr(dz1), _lib.ptr(dz2), ctx.ptr,
                                             _ws_bytes(B, D), st))
        return (dz1, dz2) + (None,) * 9


def bt_loss_cross_values(z1, z2, common_dim, batch_size, eps=1e-5, training=True, momentum=0.1, running=None):
    """The six outputs as one tensor [6] = (loss_c, on_diag_c, off_diag_c, loss_u, on_diag_u, off_diag_u).
    ``running`` = (run_mean1, run_var1, run_mean2, run_var2) fp32 device tensors, updated in place when training."""
    _lib.require_cuda(z1, z2)
    if z1.dim() != 2 or z1.shape != z2.shape:
        raise RuntimeError(f"bt_loss_cross expects two [B, D] tensors, got {tuple(z1.shape)} and {tuple(z2.shape)}")
    rm1, rv1, rm2, rv2 = running if running is not None else (None, None, None, None)
    if not training and running is None:
        raise RuntimeError("bt_loss_cross in eval mode needs the BatchNorm running statistics")
    return _BTLossCross.apply(z1.to(torch.float32), z2.to(torch.float32), common_dim, batch_size, eps, training, momentum,
                              rm1, rv1, rm2, rv2)


def bt_loss_cross(self, z1, z2, common_dim):
    """Drop-in for ``DILR.bt_loss_cross`` (code/fusion_net.py:656-677): same arguments, same 6-tuple, same side effects on
    ``self.bn1`` / ``self.bn2`` (running statistics and ``num_batches_tracked`` advance once per call in train mode)."""
    bn1, bn2 = self.bn1, self.bn2
    training = bn1.training or bn1.running_mean is None
    running = None
    if bn1.running_mean is not None:
        running = (bn1.running_mean, bn1.running_var, bn2.running_mean, bn2.running_var)
        if training:
            bn1.num_batches_tracked += 1
            bn2.num_batches_tracked += 1
    momentum = 0.1 if bn1.momentum is None else bn1.momentum
    out = bt_loss_cross_values(z1, z2, int(common_dim), self.args.batch_size, bn1.eps, training, momentum, running)
    return tuple(out.unbind(0))
