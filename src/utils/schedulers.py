"""src.utils.schedulers -> jepa_b200.schedulers."""
from jepa_b200.schedulers import WarmupCosineSchedule, CosineWDSchedule  # noqa: F401
