"""Same import surface as the reference's binary_seg/lib/pranet.py."""
from ..heads import BasicConv2d, RFB_modified, aggregation  # noqa: F401
from ..models import PVT_PraNet_V2, PraNet_V2  # noqa: F401
