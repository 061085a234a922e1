"""Import the UNMODIFIED reference modules in the build container (TEST INFRASTRUCTURE).

``/root/reference`` exists only in the build container, never on the GPU box, so
this loader is used by ``oracle/gen_golden.py`` (fixture generation) and by the
``-m "not gpu"`` tests that cross-check the numpy oracle against the live
reference when it is present.  Nothing in the product imports it.

Recipe: ``code/MMD.py`` needs only torch.  ``code/fusion_net.py``
imports ``ot``, ``matplotlib`` and the unpublished ``Models`` package at import
time, and ``EPRL`` hard-codes ``.cuda()``; stub the former in ``sys.modules`` and,
on a GPU-less host, make ``Tensor.cuda`` the identity for the duration of a call.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EDRL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "code", "MMD.py"))


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_mmd():
    """The reference's ``MMD`` module (code/MMD.py), unmodified."""
    return _load("_edrl_reference_MMD", os.path.join(REFERENCE_ROOT, "code", "MMD.py"))


def load_reference_fusion_net():
    """The reference's ``fusion_net`` module (code/fusion_net.py), unmodified, with
    import-time stubs for modules that are absent in this image."""
    stubs = {}
    for name in ("ot", "matplotlib", "matplotlib.pyplot", "Models",
                 "Models.fundus_swin_network", "Models.unetr"):
        if name not in sys.modules:
            stubs[name] = types.ModuleType(name)
    if "Models.fundus_swin_network" in stubs:
        stubs["Models.fundus_swin_network"].build_model = lambda *a, **k: None
    if "Models.unetr" in stubs:
        stubs["Models.unetr"].UNETR_base_3DNet = lambda *a, **k: None
    if "matplotlib" in stubs and "matplotlib.pyplot" in stubs:
        stubs["matplotlib"].pyplot = stubs["matplotlib.pyplot"]
    sys.modules.update(stubs)
    try:
        return _load("_edrl_reference_fusion_net",
                     os.path.join(REFERENCE_ROOT, "code", "fusion_net.py"))
    finally:
        for name in stubs:
            sys.modules.pop(name, None)


@contextlib.contextmanager
def cuda_identity_if_no_gpu():
    """EPRL calls ``.cuda()`` on freshly drawn CPU tensors (code/fusion_net.py:107,
    110,187,190,228,230).  Without a GPU make that a no-op while the block runs."""
    import torch
    if torch.cuda.is_available():
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig
