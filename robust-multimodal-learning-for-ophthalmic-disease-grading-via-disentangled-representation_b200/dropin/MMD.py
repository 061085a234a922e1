"""Drop-in replacement for the reference's ``code/MMD.py``: put this directory in front of the
reference's on ``sys.path`` (``PYTHONPATH=<repo>/<pkg>/dropin:...``) and
``from MMD import MK_MMD`` / ``from MMD import compute_js_divergence`` (code/fusion_train.py:11-12,
code/fusion_test.py:11-12) resolve to the sm_100a kernels.  Same names, same signatures."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from edrl_b200 import (MK_MMD, gaussian_kernel, compute_js_divergence, compute_kl_divergence)  # noqa: E402,F401

__all__ = ["MK_MMD", "gaussian_kernel", "compute_js_divergence", "compute_kl_divergence"]
