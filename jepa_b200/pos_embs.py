"""Fixed sin-cos positional tables (float64 numpy, like src/models/utils/pos_embs.py:11-99).

Kept on the host in float64 so the fp32 tables are bit-identical to the reference's
(checked against sha256 fixtures in tests/golden/).
"""
import numpy as np


def _sincos_1d(dim, positions):
    """positions (any shape) -> [M, dim] = [sin(pos * w) | cos(pos * w)], w_k = 10000^(-2k/dim)."""
    assert dim % 2 == 0
    freq = np.arange(dim // 2, dtype=float)
    freq /= dim / 2.
    freq = 1. / 10000 ** freq
    angle = np.einsum('m,d->md', positions.reshape(-1), freq)
    return np.concatenate([np.sin(angle), np.cos(angle)], axis=1)


def get_1d_sincos_pos_embed_from_grid(embed_dim, pos):
    return _sincos_1d(embed_dim, pos)


def get_3d_sincos_pos_embed(embed_dim, grid_size, grid_depth, cls_token=False, uniform_power=False):
    """[grid_depth*grid_size*grid_size (+1), embed_dim]; token order (d, h, w) row-major."""
    d = np.arange(grid_depth, dtype=float)
    h = np.arange(grid_size, dtype=float)
    w = np.arange(grid_size, dtype=float)
    # meshgrid argument order decides which axis varies fastest: result indexes as [d, h, w]
    gh, gd, gw = np.meshgrid(h, d, w)
    if uniform_power:
        dim_d = dim_h = dim_w = int(np.ceil(embed_dim / 6) * 2)
    else:
        dim_h = dim_w = embed_dim // 4
        dim_d = embed_dim // 2
    table = np.concatenate([_sincos_1d(dim_d, gd), _sincos_1d(dim_h, gh), _sincos_1d(dim_w, gw)], axis=1)
    table = table[:, :embed_dim]
    if cls_token:
        table = np.concatenate([np.zeros([1, embed_dim]), table], axis=0)
    return table


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    h = np.arange(grid_size, dtype=float)
    w = np.arange(grid_size, dtype=float)
    gw, gh = np.meshgrid(w, h)
    table = np.concatenate([_sincos_1d(embed_dim // 2, gh), _sincos_1d(embed_dim // 2, gw)], axis=1)
    if cls_token:
        table = np.concatenate([np.zeros([1, embed_dim]), table], axis=0)
    return table


def get_1d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    table = _sincos_1d(embed_dim, np.arange(grid_size, dtype=float))
    if cls_token:
        table = np.concatenate([np.zeros([1, embed_dim]), table], axis=0)
    return table
