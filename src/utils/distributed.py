"""src.utils.distributed -> jepa_b200.distributed."""
from jepa_b200.distributed import init_distributed, AllGather, AllReduceSum, AllReduce, DistributedDataParallel, FlatGradSync  # noqa: F401
