"""Tensor-level wrappers over the C ABI: marshal torch CUDA tensors into raw pointers + sizes.

PyTorch is only the allocator / stream provider here; every op below is one of our own sm_100a
kernels.  All functions enqueue on torch's current stream and never synchronise.
"""
import torch

from . import _lib

EPI_NONE, EPI_GELU, EPI_ADD, EPI_DGELU, EPI_MUL, EPI_GELU_GRAD = 0, 1, 2, 3, 4, 5
BF16, F32 = torch.bfloat16, torch.float32


def _s():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise _lib.VJError(f"{name} must live on a CUDA device (the hot path has no CPU fallback)")
    if not t.is_contiguous():
        raise _lib.VJError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise _lib.VJError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def _isf32(t):
    if t.dtype == F32:
        return 1
    if t.dtype == BF16:
        return 0
    raise _lib.VJError(f"unsupported dtype {t.dtype}")


def gemm(a, b, out, *, a_mn=False, b_mn=False, bias=None, alpha=1.0, epi=EPI_NONE, aux=None, aux_rowmap=None,
         aux_period=0, aux_out=None, split_k=1, accumulate=False):
    """out[M,N] = epi(alpha * A @ B^T).  a: [M,K] (or [K,M] if a_mn); b: [N,K] (or [K,N] if b_mn)."""
    _chk(a, BF16, "a"); _chk(b, BF16, "b"); _chk(out, None, "out")
    if a_mn:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    if K != Kb or tuple(out.shape) != (M, N):
        raise _lib.VJError(f"gemm shape mismatch a={tuple(a.shape)} b={tuple(b.shape)} out={tuple(out.shape)}")
    if bias is not None:
        _chk(bias, F32, "bias")
    if aux is not None:
        _chk(aux, None, "aux")
    if aux_rowmap is not None:
        _chk(aux_rowmap, torch.int32, "aux_rowmap")
    if aux_out is not None:
        _chk(aux_out, BF16, "aux_out")
    _lib.call("vj_gemm", _p(a), a.stride(0), int(a_mn), _p(b), b.stride(0), int(b_mn), _p(out), out.stride(0),
              _isf32(out), M, N, K, _p(bias), float(alpha), int(epi), _p(aux),
              aux.stride(0) if aux is not None else 0, _isf32(aux) if aux is not None else 0, _p(aux_rowmap),
              int(aux_period), _p(aux_out), aux_out.stride(0) if aux_out is not None else 0, int(split_k),
              int(accumulate), _s())
    return out


def attn_fwd(qkv, out, lse2, cu_seqlens, nseq, max_len, H, HD, scale):
    _chk(qkv, BF16, "qkv"); _chk(out, BF16, "out"); _chk(lse2, F32, "lse2"); _chk(cu_seqlens, torch.int32, "cu_seqlens")
    T = qkv.shape[0]
    _lib.call("vj_attn_fwd", _p(qkv), _p(out), _p(lse2), _p(cu_seqlens), nseq, max_len, H, HD, T, float(scale), _s())
    return out


def attn_bwd(qkv, out, dout, lse2, delta_ws, dqkv, cu_seqlens, nseq, max_len, H, HD, scale, dq_acc_ws=None):
    for t, n in ((qkv, "qkv"), (out, "out"), (dout, "dout"), (dqkv, "dqkv")):
        _chk(t, BF16, n)
    _chk(lse2, F32, "lse2"); _chk(delta_ws, F32, "delta_ws")
    T = qkv.shape[0]
    if dq_acc_ws is not None:
        _chk(dq_acc_ws, F32, "dq_acc_ws")
    _lib.call("vj_attn_bwd", _p(qkv), _p(out), _p(dout), _p(lse2), _p(delta_ws), _p(dqkv), _p(dq_acc_ws), _p(cu_seqlens),
              nseq, max_len, H, HD, T, float(scale), _s())
    return dqkv


def layernorm_fwd(x, y, gamma, beta, eps, mean=None, rstd=None):
    _chk(x, None, "x"); _chk(y, None, "y"); _chk(gamma, F32, "gamma"); _chk(beta, F32, "beta")
    T, D = x.shape
    _lib.call("vj_layernorm_fwd", _p(x), _isf32(x), _p(y), _isf32(y), _p(gamma), _p(beta), _p(mean), _p(rstd), T, D,
              float(eps), _s())
    return y


_ln_ws = {}


def layernorm_bwd(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta):
    _chk(dy, BF16, "dy"); _chk(x, None, "x"); _chk(dx, x.dtype, "dx")
    _chk(dgamma, F32, "dgamma"); _chk(dbeta, F32, "dbeta")
    T, D = x.shape
    need = _lib.load().vj_layernorm_bwd_workspace(T, D)
    key = (x.device.index, torch.cuda.current_stream().cuda_stream)
    ws = _ln_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        _ln_ws[key] = ws
    _lib.call("vj_layernorm_bwd", _p(dy), _p(x), _isf32(x), _p(gamma), _p(mean), _p(rstd), _p(dres), _p(dx),
              _p(dgamma), _p(dbeta), _p(ws), ws.numel(), T, D, _s())
    return dx


def colsum(x, out, period=0, lo=0, hi=0):
    _chk(x, None, "x"); _chk(out, F32, "out")
    T, N = x.shape
    _lib.call("vj_colsum", _p(x), _isf32(x), _p(out), T, N, x.stride(0), period, lo, hi, _s())
    return out


def im2col_tubelets(clips, patches, idx, tubelet, patch):
    _chk(clips, F32, "clips"); _chk(patches, BF16, "patches")
    B, C, T, H, W = clips.shape
    K = 0
    if idx is not None:
        _chk(idx, torch.int64, "idx")
        K = idx.shape[1]
    _lib.call("vj_im2col_tubelets", _p(clips), _p(patches), _p(idx), B, C, T, H, W, tubelet, patch, K, _s())
    return patches


def gather_rows(x, idx, out=None):
    """apply_masks for one mask: x [B,N,D], idx int64 [B,K] -> [B,K,D] (bit-exact row copy)."""
    _chk(x, None, "x"); _chk(idx, torch.int64, "idx")
    B, N, D = x.shape
    K = idx.shape[1]
    if out is None:
        out = torch.empty(B, K, D, dtype=x.dtype, device=x.device)
    _lib.call("vj_gather_rows", _p(x), _p(out), _p(idx), B, N, K, D * x.element_size(), _s())
    return out


def scatter_rows_add(dy, dx, idx):
    _chk(dy, None, "dy"); _chk(dx, dy.dtype, "dx"); _chk(idx, torch.int64, "idx")
    B, N, D = dx.shape
    K = idx.shape[1]
    _lib.call("vj_scatter_rows_add", _p(dy), _p(dx), _p(idx), B, N, K, D, _isf32(dx), _s())
    return dx


def target_ln_gather(x, idx, gamma, beta, eps_norm, eps_target, out=None):
    _chk(x, BF16, "x"); _chk(idx, torch.int64, "idx")
    B, N, D = x.shape
    K = idx.shape[1]
    if out is None:
        out = torch.empty(B, K, D, dtype=F32, device=x.device)
    _lib.call("vj_target_ln_gather", _p(x), _p(out), _p(idx), _p(gamma), _p(beta), B, N, K, D, float(eps_norm),
              float(eps_target), _s())
    return out


def pred_assemble_fwd(emb, pos, mask_token, idx_ctx, idx_tgt, x, B, Ke, Kp, Dp):
    _chk(emb, BF16, "emb"); _chk(pos, F32, "pos"); _chk(mask_token, F32, "mask_token"); _chk(x, None, "x")
    _lib.call("vj_pred_assemble_fwd", _p(emb), _p(pos), _p(mask_token), _p(idx_ctx), _p(idx_tgt), _p(x), _isf32(x),
              B, Ke, Kp, Dp, _s())
    return x


def pred_assemble_bwd(dx, demb, dmask_token, B, Ke, Kp, Dp):
    _chk(dx, None, "dx"); _chk(demb, BF16, "demb"); _chk(dmask_token, F32, "dmask_token")
    _lib.call("vj_pred_assemble_bwd", _p(dx), _isf32(dx), _p(demb), _p(dmask_token), B, Ke, Kp, Dp, _s())


def seq_slice(src, dst, B, Ke, Kp, D, scatter=False, zero_ctx=False):
    _chk(src, None, "src"); _chk(dst, src.dtype, "dst")
    _lib.call("vj_seq_slice", _p(src), _p(dst), _isf32(src), B, Ke, Kp, D, int(scatter), int(zero_ctx), _s())
    return dst


def l1_loss_fwd(z, h, loss_sum, weight):
    _chk(z, BF16, "z"); _chk(h, F32, "h"); _chk(loss_sum, F32, "loss_sum")
    _lib.call("vj_l1_loss_fwd", _p(z), _p(h), _p(loss_sum), z.numel(), float(weight), _s())


def l1_loss_bwd(z, h, grad_scale, scale, dz):
    _chk(z, BF16, "z"); _chk(h, F32, "h"); _chk(dz, BF16, "dz")
    _lib.call("vj_l1_loss_bwd", _p(z), _p(h), _p(grad_scale), float(scale), _p(dz), z.numel(), _s())
    return dz


def lp_loss_fwd(z, h, loss_sum, weight, p):
    _chk(z, BF16, "z"); _chk(h, F32, "h"); _chk(loss_sum, F32, "loss_sum")
    _lib.call("vj_lp_loss_fwd", _p(z), _p(h), _p(loss_sum), z.numel(), float(weight), float(p), _s())


def lp_loss_bwd(z, h, grad_scale, scale, dz, p):
    _chk(z, BF16, "z"); _chk(h, F32, "h"); _chk(dz, BF16, "dz")
    _lib.call("vj_lp_loss_bwd", _p(z), _p(h), _p(grad_scale), float(scale), _p(dz), z.numel(), float(p), _s())
    return dz


def token_std_bwd(z, pstd_total, grad_scale, scale, dz, weight, eps=1e-4):
    _chk(z, BF16, "z"); _chk(pstd_total, F32, "pstd_total"); _chk(dz, BF16, "dz")
    B, K, D = z.shape
    _lib.call("vj_token_std_bwd", _p(z), _p(pstd_total), _p(grad_scale), float(scale), _p(dz), B, K, D, float(eps),
              float(weight), _s())
    return dz


def cross_attn_fwd(q, kv, out, B, nq, S, H, hd, scale):
    """softmax(q k^T scale) v for nq query tokens per clip over S keys (attentive probe, modules.py:138-153)."""
    _chk(q, BF16, "q"); _chk(kv, BF16, "kv"); _chk(out, BF16, "out")
    _lib.call("vj_cross_attn_fwd", _p(q), _p(kv), _p(out), B, nq, S, H, hd, float(scale), _s())
    return out


def token_std_accum(z, pstd, weight, eps=1e-4):
    _chk(z, BF16, "z"); _chk(pstd, F32, "pstd")
    B, K, D = z.shape
    _lib.call("vj_token_std_accum", _p(z), _p(pstd), B, K, D, float(eps), float(weight), _s())


def cast_f32_bf16(src, dst):
    _chk(src, F32, "src"); _chk(dst, BF16, "dst")
    _lib.call("vj_cast_f32_bf16", _p(src), _p(dst), src.numel(), _s())
    return dst


def head_pad(src, dst, outer, G, hd, hdp, inner, unpad_add=False):
    _chk(src, None, "src"); _chk(dst, None, "dst")
    _lib.call("vj_head_pad", _p(src), _isf32(src), _p(dst), _isf32(dst), outer, G, hd, hdp, inner, int(unpad_add), _s())
    return dst


def ema_update(k_flat, q_flat, m):
    _chk(k_flat, F32, "k"); _chk(q_flat, F32, "q")
    _lib.call("vj_ema_update", _p(k_flat), _p(q_flat), k_flat.numel(), float(m), float(1.0 - m), _s())


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, wd, step, inv_scale=None, found_inf=None):
    _lib.call("vj_adamw_step", _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), float(beta1), float(beta2),
              float(eps), float(wd), int(step), _p(inv_scale), _p(found_inf), _s())


def sumsq(x, out):
    _chk(x, F32, "x"); _chk(out, F32, "out")
    _lib.call("vj_sumsq", _p(x), x.numel(), _p(out), _s())


def ema_update_shadow(k_flat, q_flat, m, shadow):
    _chk(k_flat, F32, "k"); _chk(q_flat, F32, "q"); _chk(shadow, BF16, "shadow")
    _lib.call("vj_ema_update_shadow", _p(k_flat), _p(q_flat), k_flat.numel(), float(m), float(1.0 - m), _p(shadow), _s())


def grad_unscale_stats(gflat, seg, sumsq_out, inv_scale=None, found_inf=None, write_back=True):
    """One pass over a flat fp32 gradient buffer: optional in-place unscale, non-finite flag, per-tensor sum of squares."""
    _chk(gflat, F32, "gflat"); _chk(seg, torch.uint16, "seg"); _chk(sumsq_out, F32, "sumsq_out")
    _lib.call("vj_grad_unscale_stats", _p(gflat), _p(seg), gflat.numel(), _p(inv_scale), _p(found_inf), _p(sumsq_out),
              int(bool(write_back)), _s())


def seg_abs_sum(x, seg, out):
    _chk(x, F32, "x"); _chk(seg, torch.uint16, "seg"); _chk(out, F32, "out")
    _lib.call("vj_seg_abs_sum", _p(x), _p(seg), x.numel(), _p(out), _s())


def clip_coef(sumsq, max_norm, total_norm_out, coef_out):
    _chk(sumsq, F32, "sumsq")
    _lib.call("vj_clip_coef", _p(sumsq), sumsq.numel(), float(max_norm), _p(total_norm_out), _p(coef_out), _s())


def scale_flat(x, coef_dev):
    _chk(x, F32, "x"); _chk(coef_dev, F32, "coef")
    _lib.call("vj_scale_flat", _p(x), x.numel(), _p(coef_dev), _s())
