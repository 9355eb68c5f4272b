"""src.models.utils.multimask -> jepa_b200.models."""
from jepa_b200.models import MultiMaskWrapper, PredictorMultiMaskWrapper  # noqa: F401
