"""Loss entry points with the reference's call signatures.

`structure_loss(pred, pred_bg, mask_fg, mask_bg)` is binary_seg/MyTrain_med.py:19; the multiclass dual
loss is the block at EMCAD/trainer.py:123-140 (== MERIT/train_ACDC.py:259-284 == MIST/trainer.py:112-129).
"""
from __future__ import annotations

from .ops import mc_dual_loss, structure_loss, structure_loss_lowres, structure_loss_multi  # noqa: F401
