"""ctypes binding of ``lib/libedrl_b200.so`` (the C-ABI declared in ``include/edrl_b200.h``).

There is no CPU fallback and no other backend: if the library has not been built (run
``python __graft_entry__.py`` or ``python <package>/build.py``) or a tensor is not on a CUDA
device, the call raises.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libedrl_b200.so")

c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes).  Every symbol of include/edrl_b200.h; tests check the two agree.
PROTOTYPES = {
    "edrl_abi_version": (c_int, []),
    "edrl_last_error": (ctypes.c_char_p, []),
    "edrl_launch_count": (ctypes.c_uint64, []),
    "edrl_set_device": (c_int, [c_int]),
    "edrl_mmd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "edrl_mmd_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "edrl_mmd_finalize": (c_int, [c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                  c_void_p]),
    "edrl_mmd_backward": (c_int, [c_int, c_int, c_int, c_float, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                  c_void_p, c_void_p, c_size_t, c_void_p]),
    "edrl_mmd_grad_slabs": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "edrl_mmd_sweep_plan": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_int)]),
    "edrl_mmd_forward_grad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                      c_void_p]),
    "edrl_mmd_apply_grad": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "edrl_mmd_kernel_matrix": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p,
                                       c_void_p, c_size_t, c_void_p]),
    "edrl_token_stats_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "edrl_token_stats_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                     c_void_p]),
    "edrl_token_featmean": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "edrl_proxy_normalize_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                         c_void_p]),
    "edrl_proxy_normalize_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_void_p, c_void_p, c_void_p]),
    "edrl_score_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "edrl_score_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "edrl_topk_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "edrl_select_topk_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "edrl_proxy_loss_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "edrl_select_loss_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p]),
    "edrl_essence_saved_floats": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "edrl_essence_scratch_floats": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "edrl_essence_train_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p]),
    "edrl_essence_train_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "edrl_gather_rows_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "edrl_gather_rows_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "edrl_noise_views": (c_int, [c_void_p, c_int, ctypes.c_longlong, c_int, c_float, ctypes.c_ulonglong, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    "edrl_head_losses_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int, c_int, c_void_p, c_void_p]),
    "edrl_head_losses_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "edrl_dilr_workspace_bytes": (c_size_t, [c_int, c_int]),
    "edrl_dilr_bt_loss_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_float, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "edrl_dilr_bt_loss_bwd": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                      c_void_p]),
}

ABI_VERSION = 1
_lock = threading.Lock()
_lib = None


class EdrlLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library once; raise loudly when it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise EdrlLibraryError(
                f"{LIB_PATH} is missing: build the sm_100a kernels first (python __graft_entry__.py). "
                "This package has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)        # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.edrl_abi_version() != ABI_VERSION:
            raise EdrlLibraryError(f"libedrl_b200.so has ABI {lib.edrl_abi_version()}, expected {ABI_VERSION}")
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().edrl_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def launch_count() -> int:
    return int(load().edrl_launch_count())


def check(rc: int, exc=RuntimeError):
    """Status 1 = argument error (ValueError-like in the reference's torch), others = CUDA/runtime."""
    if rc == 0:
        return
    msg = last_error()
    if rc == 1:
        raise (exc if exc is not RuntimeError else ValueError)(msg)
    raise RuntimeError(msg)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "edrl_b200 kernels run on a CUDA (sm_100a) device only; got a tensor on "
                f"'{t.device}'. There is no CPU fallback.")


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream_and_device(t: torch.Tensor):
    """Current torch stream handle for t's device; names that device to the library (each C-ABI call binds the thread
    to it for its own duration and restores torch's current device before returning)."""
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    check(load().edrl_set_device(dev))
    return torch.cuda.current_stream(dev).cuda_stream
