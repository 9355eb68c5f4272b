"""src.models.vision_transformer (reference: src/models/vision_transformer.py) -> jepa_b200.models."""
from jepa_b200.models import (  # noqa: F401
    VisionTransformer, vit_tiny, vit_small, vit_base, vit_large, vit_huge, vit_giant, vit_gigantic, VIT_EMBED_DIMS)
