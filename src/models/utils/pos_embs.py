"""src.models.utils.pos_embs -> jepa_b200.pos_embs."""
from jepa_b200.pos_embs import (  # noqa: F401
    get_3d_sincos_pos_embed, get_2d_sincos_pos_embed, get_1d_sincos_pos_embed, get_1d_sincos_pos_embed_from_grid)
