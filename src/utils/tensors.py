"""src.utils.tensors -> jepa_b200.tensors."""
from jepa_b200.tensors import trunc_normal_, repeat_interleave_batch  # noqa: F401
from jepa_b200.tensors import apply_masks as _apply_masks


def apply_masks(x, masks):
    """Two-argument variant (src/utils/tensors.py:53-62): always concatenates along batch."""
    return _apply_masks(x, masks, concat=True)
