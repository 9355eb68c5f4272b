"""Small tensor helpers with the reference's semantics (src/utils/tensors.py)."""
import math

import torch


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    """In-place truncated normal via inverse-CDF sampling (tensors.py:17-50).

    NB: a/b are ABSOLUTE bounds, so with std=0.02 the truncation at +-2 is ~100 sigma - the same
    quirk as the reference; the RNG call sequence (uniform_ then erfinv_) is kept so seeded inits agree.
    """
    def cdf(x):
        return (1. + math.erf(x / math.sqrt(2.))) / 2.

    with torch.no_grad():
        lo, hi = cdf((a - mean) / std), cdf((b - mean) / std)
        tensor.uniform_(2 * lo - 1, 2 * hi - 1)
        tensor.erfinv_()
        tensor.mul_(std * math.sqrt(2.))
        tensor.add_(mean)
        tensor.clamp_(min=a, max=b)
    return tensor


def repeat_interleave_batch(x, B, repeat):
    """[c0 | c1 | ...] chunks of B rows -> every chunk repeated `repeat` times in place (tensors.py:65-71)."""
    n_chunks = len(x) // B
    if repeat == 1:
        return x[:n_chunks * B] if n_chunks * B != len(x) else x
    pieces = []
    for i in range(n_chunks):
        pieces.extend([x[i * B:(i + 1) * B]] * repeat)
    return torch.cat(pieces, dim=0)


def apply_masks(x, masks, concat=True):
    """Keep-index gather (src/masks/utils.py:11-23): x [B,N,D], masks list of int64 [B,K] -> [len*B,K,D].

    CUDA tensors go through the vj_gather_rows kernel (no [B,K,D] int64 index is materialised); CPU
    tensors (dataloader-side uses) use torch.gather.  Both are bit-exact row copies.
    """
    outs = []
    for m in masks:
        if x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and (x.shape[-1] * x.element_size()) % 16 == 0 \
                and not (torch.is_grad_enabled() and x.requires_grad):
            from . import kernels as K
            outs.append(K.gather_rows(x.contiguous(), m.contiguous()))
        else:
            outs.append(torch.gather(x, dim=1, index=m.unsqueeze(-1).expand(-1, -1, x.size(-1))))
    if not concat:
        return outs
    return torch.cat(outs, dim=0)
