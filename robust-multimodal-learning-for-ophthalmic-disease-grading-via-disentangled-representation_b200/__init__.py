"""edrl_b200 -- B200-native (sm_100a) implementation of the EDRL hot path:
multi-bandwidth Gaussian MMD (``MK_MMD``) and Essence-Point scoring / top-k selection (``EPRL``),
behind the reference's own Python signatures.  Import as ``edrl_b200`` via the repo-root shim
(the directory name is not a valid Python identifier)."""
from . import _lib
from .mmd import (MK_MMD, gaussian_kernel, compute_js_divergence, compute_kl_divergence, mk_mmd_with_stats,
                  set_default_precision, get_default_precision)
from .sharded import sharded_MK_MMD, RowBlockPlan
from . import dilr
from .views import noise_views
from .head import head_losses
from .dilr import bt_loss_cross, bt_loss_cross_values
from .eprl import (EPRL, essence_scores, essence_select_loss, essence_train_loss, topk_rows, gather_rows,
                   select_gather)

__all__ = ["MK_MMD", "gaussian_kernel", "compute_js_divergence", "compute_kl_divergence", "mk_mmd_with_stats",
           "set_default_precision", "get_default_precision", "EPRL", "essence_scores", "essence_select_loss", "essence_train_loss",
           "topk_rows", "gather_rows", "select_gather", "launch_count", "sharded_MK_MMD", "RowBlockPlan", "bt_loss_cross",
           "bt_loss_cross_values", "noise_views", "head_losses"]


def launch_count() -> int:
    """Kernels launched by libedrl_b200.so since it was loaded."""
    return _lib.launch_count()
