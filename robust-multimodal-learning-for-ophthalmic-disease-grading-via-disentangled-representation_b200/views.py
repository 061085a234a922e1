"""The loader's clean / noisy views generated on the device (reference: the per-item numpy code of
code/data_harvard.py:698-783).  ``noise_views(x)`` returns ``(low, high)`` with the reference's semantics:
``low = clip(x, 0, 1)``, ``high = clip(x + N(0, sigma), 0, 1)``; uint8 input gets the ``/ 255`` of :694-695 fused."""
from __future__ import annotations

import torch

from . import _lib


def noise_views(x, sigma=0.5, seed=11, shared_field=True, noise=None):
    """x: [items, ...] CUDA tensor, float32 in [0, 1] or uint8.  ``seed`` defaults to the reference's ``seed_idx = 11``
    (code/fusion_train.py:545); ``shared_field=True`` gives every item the same noise field, as the reference's per-item
    reseeding does.  ``noise`` (float32, x's shape): add this field instead of drawing one (parity mode)."""
    _lib.require_cuda(x)
    if x.dtype not in (torch.float32, torch.uint8):
        x = x.to(torch.float32)
    xc = x.contiguous()
    items = xc.shape[0] if xc.dim() > 1 else 1
    per_item = xc.numel() // max(items, 1)
    low = torch.empty(xc.shape, dtype=torch.float32, device=xc.device)
    high = torch.empty_like(low)
    nz = None
    if noise is not None:
        if noise.shape != xc.shape:
            raise RuntimeError(f"noise must have x's shape {tuple(xc.shape)}, got {tuple(noise.shape)}")
        nz = noise.to(device=xc.device, dtype=torch.float32).contiguous()
    st = _lib.stream_and_device(xc)
    _lib.check(_lib.load().edrl_noise_views(xc.data_ptr(), int(xc.dtype == torch.uint8), per_item, items, float(sigma),
                                            int(seed) & 0xFFFFFFFFFFFFFFFF, int(bool(shared_field)), _lib.ptr(nz),
                                            low.data_ptr(), high.data_ptr(), st))
    return low, high
