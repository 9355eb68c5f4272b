"""src.utils.logging -> jepa_b200.logging_utils."""
from jepa_b200.logging_utils import (  # noqa: F401
    gpu_timer, get_logger, CSVLogger, AverageMeter, grad_logger, adamw_logger, LOG_FORMAT, DATE_FORMAT)
