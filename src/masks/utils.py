"""src.masks.utils -> jepa_b200.tensors.apply_masks (kernel-backed on CUDA tensors)."""
from jepa_b200.tensors import apply_masks  # noqa: F401
