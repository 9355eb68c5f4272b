"""src.models.utils.modules (reference: src/models/utils/modules.py:13-182) -> jepa_b200.models / jepa_b200.pooler."""
from jepa_b200.models import MLP, Attention, Block  # noqa: F401
from jepa_b200.pooler import CrossAttention, CrossAttentionBlock  # noqa: F401
