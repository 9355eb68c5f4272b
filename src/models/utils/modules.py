"""src.models.utils.modules (reference: src/models/utils/modules.py:13-120) -> jepa_b200.models.
CrossAttention / CrossAttentionBlock (eval probes only) are outside the pre-training path."""
from jepa_b200.models import MLP, Attention, Block  # noqa: F401
