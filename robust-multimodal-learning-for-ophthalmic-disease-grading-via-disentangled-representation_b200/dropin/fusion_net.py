"""Drop-in replacement for the reference's ``fusion_net`` module.

``EPRL`` is the sm_100a-backed class (same constructor, ``state_dict`` keys and forward contract,
code/fusion_net.py:63-255).  Everything else of the reference module (``MedFusion``, ``PoE``,
``DILR`` ...) is caller code this project does not re-implement: when the reference source is
available (``EDRL_REFERENCE_ROOT``, default ``/root/reference``) and its imports resolve, it is
executed into this module's namespace and only the name ``EPRL`` is rebound, so
``MedFusion.__init__`` (code/fusion_net.py:817-821) instantiates the accelerated class unchanged.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from edrl_b200 import EPRL as _EPRL_B200  # noqa: E402

_ref = os.path.join(os.environ.get("EDRL_REFERENCE_ROOT", "/root/reference"), "code", "fusion_net.py")
REFERENCE_LOADED = False
if os.path.isfile(_ref):
    try:
        with open(_ref) as _f:
            exec(compile(_f.read(), _ref, "exec"), globals())   # the unmodified caller code
        REFERENCE_LOADED = True
    except ImportError:
        # the reference imports unpublished packages (Models.*, ot, ...): callers are unavailable here
        REFERENCE_LOADED = False

EPRL = _EPRL_B200
