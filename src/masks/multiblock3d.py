"""src.masks.multiblock3d -> jepa_b200.masks."""
from jepa_b200.masks import MultiBlock3DMaskCollator as MaskCollator  # noqa: F401
from jepa_b200.masks import _MultiBlock3DGenerator as _MaskGenerator  # noqa: F401
