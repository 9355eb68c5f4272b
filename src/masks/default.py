"""src.masks.default -> jepa_b200.masks."""
from jepa_b200.masks import DefaultCollator  # noqa: F401
