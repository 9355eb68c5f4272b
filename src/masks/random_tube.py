"""src.masks.random_tube -> jepa_b200.masks."""
from jepa_b200.masks import RandomTubeMaskCollator as MaskCollator  # noqa: F401
from jepa_b200.masks import _RandomTubeGenerator as _MaskGenerator  # noqa: F401
