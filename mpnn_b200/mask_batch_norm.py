"""Drop-in for the reference's top-level `mask_batch_norm` module (models/mask_batch_norm.py)."""
from .modules import MaskBatchNorm, MaskBatchNorm1d  # noqa: F401
