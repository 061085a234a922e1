"""Stand-in for ``DR_2.fusion_net`` (code/fusion_train.py:731 imports ``MedFusion`` from this unpublished package): the
reference's published ``code/fusion_net.py`` with the two broken statements of SURVEY.md F6 neutralised, loaded through
``<pkg>/dropin/fusion_net.py`` -- with ``EPRL`` rebound to the sm_100a class (swapped arm) or left as the reference's own
(reference arm, ``EDRL_SWAP_EPRL=0``).  ``MedFusion`` is subclassed only to put CUDA events around each forward call for
the step-time report of ``examples/run_reference_driver.py``; arithmetic and parameters are the reference's."""
import os
import time

import torch

os.environ.setdefault("EDRL_PATCH_MEDFUSION", "1")

import fusion_net as _fn  # noqa: E402   (<pkg>/dropin/fusion_net.py: first on sys.path in both arms)

if not _fn.REFERENCE_LOADED:
    raise ImportError(f"the reference's fusion_net.py could not be loaded: {_fn.REFERENCE_ERROR}")

STATS = {"forward_host_t": [], "events": [], "models": [], "training": []}


class MedFusion(_fn.MedFusion):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        STATS["models"].append(self)

    def forward(self, X, y, epoch):
        cuda = torch.cuda.is_available()
        if cuda:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        STATS["forward_host_t"].append(time.perf_counter())
        STATS["training"].append(self.training)
        out = super().forward(X, y, epoch)
        if cuda:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            STATS["events"].append((e0, e1))
        return out


EPRL = _fn.EPRL
