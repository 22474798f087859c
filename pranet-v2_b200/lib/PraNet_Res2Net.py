"""Same import surface as the reference's binary_seg/lib/PraNet_Res2Net.py (PraNet-V1)."""
from ..heads import BasicConv2d, RFB_modified, aggregation  # noqa: F401
from ..models import PVT_PraNet, PraNet  # noqa: F401
