"""The small losses that close ``MedFusion.forward`` as one launch each way (reference:
code/fusion_net.py:929-942 with ``KL_between_normals`` / ``get_KL_loss``, :390-402, 838-850)."""
from __future__ import annotations

import torch

from . import _lib


class _HeadLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, y, mu_f, sig_f, mu_o, sig_o, smoothing, num_classes):
        lib = _lib.load()
        p = pred.to(torch.float32)
        if p.stride(-1) != 1:
            p = p.contiguous()
        ts = [t.to(torch.float32).contiguous() for t in (mu_f, sig_f, mu_o, sig_o)]
        yl = y.to(torch.int64).contiguous()
        B, C = p.shape[0], int(num_classes)
        Cm, F = ts[0].shape[1], ts[0].shape[2]
        out = torch.empty(3, dtype=torch.float32, device=p.device)
        st = _lib.stream_and_device(p)
        _lib.check(lib.edrl_head_losses_fwd(p.data_ptr(), p.stride(0), yl.data_ptr(), B, C, float(smoothing), ts[0].data_ptr(),
                                            ts[1].data_ptr(), ts[2].data_ptr(), ts[3].data_ptr(), Cm, F, out.data_ptr(), st))
        ctx.save_for_backward(p, yl, *ts)
        ctx.cfg = (B, C, Cm, F, float(smoothing), pred.shape[1])
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        p, yl, mu_f, sig_f, mu_o, sig_o = ctx.saved_tensors
        B, C, Cm, F, smoothing, width = ctx.cfg
        g = g.to(torch.float32).contiguous()
        need = ctx.needs_input_grad
        dpred = torch.empty(B, C, dtype=torch.float32, device=g.device) if need[0] else None
        outs = [torch.empty_like(t) if need[2 + i] else None for i, t in enumerate((mu_f, sig_f, mu_o, sig_o))]
        st = _lib.stream_and_device(g)
        _lib.check(lib.edrl_head_losses_bwd(p.data_ptr(), p.stride(0), yl.data_ptr(), B, C, smoothing, mu_f.data_ptr(),
                                            sig_f.data_ptr(), mu_o.data_ptr(), sig_o.data_ptr(), Cm, F, g.data_ptr(),
                                            _lib.ptr(dpred), *[_lib.ptr(o) for o in outs], st))
        if dpred is not None and width != C:                 # the reference slices pred[:, :C]: the other columns get zero
            full = torch.zeros(B, width, dtype=torch.float32, device=g.device)
            full[:, :C] = dpred
            dpred = full
        return (dpred, None, *outs, None, None)


def head_losses(pred, y, mu_fundus, sigma_fundus, mu_oct, sigma_oct, smoothing=0.1, num_classes=2):
    """(loss1, kl_fundus, kl_oct): the label-smoothed cross-entropy over ``pred[:, :num_classes]`` and the two
    ``get_KL_loss`` terms of code/fusion_net.py:929-942, differentiable w.r.t. pred and the four proxy tensors."""
    _lib.require_cuda(pred, mu_fundus, sigma_fundus, mu_oct, sigma_oct)
    out = _HeadLosses.apply(pred, y, mu_fundus, sigma_fundus, mu_oct, sigma_oct, smoothing, num_classes)
    return out[0], out[1], out[2]
