"""Drop-in replacement for the reference's ``fusion_net`` module.

``EPRL`` is the sm_100a-backed class (same constructor, ``state_dict`` keys and forward contract,
code/fusion_net.py:63-255).  Everything else of the reference module (``MedFusion``, ``PoE``,
``DILR`` ...) is caller code this project does not re-implement: when the reference source is
available (``EDRL_REFERENCE_ROOT`` or the ``oracle/_ref`` copy) and its imports resolve, it is executed
into a scratch namespace, copied into this module on success, and only the name ``EPRL`` is rebound, so
``MedFusion.__init__`` (code/fusion_net.py:817-821) instantiates the accelerated class unchanged.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from edrl_b200 import EPRL as _EPRL_B200  # noqa: E402

# Where the caller code comes from: EDRL_REFERENCE_ROOT/code/fusion_net.py (a checkout of the reference), else the copy
# `oracle/build_ref.py` makes under oracle/_ref/ (git-ignored; it travels to the GPU box).  Never a hard-coded path.
_candidates = []
if os.environ.get("EDRL_REFERENCE_ROOT"):
    _candidates.append(os.path.join(os.environ["EDRL_REFERENCE_ROOT"], "code", "fusion_net.py"))
_candidates.append(os.path.join(_ROOT, "oracle", "_ref", "fusion_net.py"))
_ref = next((p for p in _candidates if os.path.isfile(p)), None)
REFERENCE_LOADED = False
REFERENCE_ERROR = None

# `MedFusion.forward` as published cannot execute -- it calls a method the class does not have
# (code/fusion_net.py:905-906; the result `eps` is never used) and `DILR` applies Linear(1024, ..) projectors to the
# 256-wide guided features (:642-643 against :730-731).  EDRL_PATCH_MEDFUSION=1 rewrites exactly those two statements
# in the source text at exec time so that the reference's own drivers can run; the file on disk is untouched.
_PATCHES = (
    ("        eps = self.gaussian_noise(samples=(16, self.sample_num), k=dim,\n"
     "                                  seed=self.seed)  # eps torch.Size([8, 50, 2])\n", ""),
    ("self.guided_features_projector1 = nn.Linear(1024,", "self.guided_features_projector1 = nn.Linear(256,"),
    ("self.guided_features_projector2 = nn.Linear(1024,", "self.guided_features_projector2 = nn.Linear(256,"),
)


def _patch_medfusion(src):
    for old, new in _PATCHES:
        if src.count(old) != 1:
            raise RuntimeError("EDRL_PATCH_MEDFUSION: the reference source does not look like the surveyed revision "
                               f"(pattern {old[:50]!r} found {src.count(old)} times)")
        src = src.replace(old, new)
    return src


if _ref is not None:
    _ns = {"__name__": __name__, "__file__": _ref, "__builtins__": __builtins__}
    try:
        with open(_ref) as _f:
            _src = _f.read()
        if os.environ.get("EDRL_PATCH_MEDFUSION") == "1":
            _src = _patch_medfusion(_src)
        exec(compile(_src, _ref, "exec"), _ns)             # the reference's caller code, into a scratch namespace
    except ImportError as _e:
        # the reference imports unpublished packages (Models.*, ot, ...): the callers are unavailable here, and nothing
        # half-defined leaks into this module
        REFERENCE_ERROR = f"{type(_e).__name__}: {_e}"
    else:
        REFERENCE_EPRL = _ns["EPRL"]                       # the reference's own class, for A/B runs
        if os.environ.get("EDRL_SWAP_EPRL", "1") != "0":
            _ns["EPRL"] = _EPRL_B200                       # the classes' global scope: MedFusion.__init__ looks EPRL up here
        if os.environ.get("EDRL_SWAP_DILR", os.environ.get("EDRL_SWAP_EPRL", "1")) != "0":
            # the Barlow-Twins cross-correlation loss of DILR on the fused kernels (same method signature)
            from edrl_b200.dilr import bt_loss_cross as _bt_loss_cross_b200
            REFERENCE_BT_LOSS_CROSS = _ns["DILR"].bt_loss_cross
            _ns["DILR"].bt_loss_cross = _bt_loss_cross_b200
        globals().update({k: v for k, v in _ns.items() if not k.startswith("__")})
        REFERENCE_LOADED = True

if os.environ.get("EDRL_SWAP_EPRL", "1") != "0" or not REFERENCE_LOADED:
    EPRL = _EPRL_B200
