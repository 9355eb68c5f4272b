"""Import path of the reference (`from src.models.attentive_pooler import AttentiveClassifier`, evals/*/eval.py:32)."""
from jepa_b200.pooler import AttentiveClassifier, AttentivePooler  # noqa: F401
