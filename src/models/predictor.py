"""src.models.predictor (reference: src/models/predictor.py) -> jepa_b200.models."""
from jepa_b200.models import VisionTransformerPredictor, vit_predictor  # noqa: F401
