"""pranet_v2_b200 -- B200-native DSRA decoder head + dual-supervision losses of PraNet-V2.

Host side is Python/PyTorch and mirrors the reference's module API; all hot-path computation happens in
libpranetv2_b200.so (hand-written sm_100a CUDA, C ABI in include/pv2.h).  There is no CPU fallback.
"""
from . import _lib, engine, ops  # noqa: F401
from .engine import get_precision, set_precision  # noqa: F401
from .heads import DSRAStages  # noqa: F401
from .losses import mc_dual_loss, structure_loss, structure_loss_lowres, structure_loss_multi  # noqa: F401
from .models import PVT_PraNet, PVT_PraNet_V2, PraNet, PraNet_V2  # noqa: F401
from .multiclass import CAM, CASCADE_Add_dual, EMCAD_dual, EMCADNet  # noqa: F401
from .ops import dsra_fuse, interpolate_bilinear, ra_v1_scale  # noqa: F401

__version__ = "0.1.0"
