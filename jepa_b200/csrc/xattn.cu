// Cross-attention of a few learned query tokens over a long key / value sequence: the attentive probe of the frozen-
// encoder evaluations (src/models/utils/modules.py:122-153 CrossAttention.forward, used by AttentivePooler,
// src/models/attentive_pooler.py:96-102).  SURVEY section 8 row f4 (inference side).
//
//   q   bf16 [B*nq, H*hd]          (the q Linear's output; row b*nq + j = query j of clip b)
//   kv  bf16 [B*S, 2*H*hd]         (the kv Linear's output: k | v halves, head-major inside a half)
//   out bf16 [B*nq, H*hd] = softmax(q k^T * scale) v      per (clip, head, query)
//
// nq is 1 for the probe, S is 1568 .. 9216 encoder tokens: the work is a GEMV-shaped streaming pass over K and V
// (2 * S * hd * 2 bytes per (clip, head)), i.e. HBM / L2 bound, so there is nothing for the tensor cores to do: one CTA per
// (head, clip*query), eight warps split the keys, a group of hd/8 lanes owns one key at a time (16-byte loads), online
// softmax per group, groups and warps are merged through shared memory at the end.
#include "common.cuh"
#include "vjepa_b200.h"

namespace vj {

constexpr int kXattnThreads = 256;

template <int LPK>   // lanes per key = hd / 8
__global__ void __launch_bounds__(kXattnThreads) xattn_fwd_kernel(const __nv_bfloat16* __restrict__ q,
                                                                  const __nv_bfloat16* __restrict__ kv,
                                                                  __nv_bfloat16* __restrict__ out, int nq, int S, int H,
                                                                  float scale_log2) {
  constexpr int HD = LPK * 8;
  constexpr int KPW = 32 / LPK;                 // keys handled per warp iteration
  constexpr int NGROUPS = (kXattnThreads / 32) * KPW;
  const int head = blockIdx.x, bq = blockIdx.y;
  const int b = bq / nq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / LPK, gl = lane % LPK;   // key slot inside the warp, 8-element slice of the head dim
  const bool active = grp < KPW;
  const long long D = (long long)H * HD;

  float qf[8];
  {
    const uint4 u = *reinterpret_cast<const uint4*>(q + (long long)bq * D + head * HD + (active ? gl : 0) * 8);
    qf[0] = bf16_lo(u.x); qf[1] = bf16_hi(u.x); qf[2] = bf16_lo(u.y); qf[3] = bf16_hi(u.y);
    qf[4] = bf16_lo(u.z); qf[5] = bf16_hi(u.z); qf[6] = bf16_lo(u.w); qf[7] = bf16_hi(u.w);
  }
  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  const __nv_bfloat16* kbase = kv + (long long)b * S * 2 * D + head * HD + gl * 8;
  for (int j0 = warp * KPW; j0 < S; j0 += (kXattnThreads / 32) * KPW) {
    const int j = j0 + grp;
    const bool ok = active && j < S;
    float part = 0.f;
    uint4 vv = make_uint4(0, 0, 0, 0);
    if (ok) {
      const uint4 ku = *reinterpret_cast<const uint4*>(kbase + (long long)j * 2 * D);
      vv = *reinterpret_cast<const uint4*>(kbase + (long long)j * 2 * D + D);
      part = qf[0] * bf16_lo(ku.x) + qf[1] * bf16_hi(ku.x) + qf[2] * bf16_lo(ku.y) + qf[3] * bf16_hi(ku.y) +
             qf[4] * bf16_lo(ku.z) + qf[5] * bf16_hi(ku.z) + qf[6] * bf16_lo(ku.w) + qf[7] * bf16_hi(ku.w);
    }
    // score of key j = sum of the LPK partial dot products of its lane group (rotation inside the group)
    float s = part;
#pragma unroll
    for (int i = 1; i < LPK; ++i) s += __shfl_sync(0xffffffffu, part, grp * LPK + (gl + i) % LPK);
    if (ok) {
      s *= scale_log2;
      const float mn = fmaxf(m, s);
      const float corr = ex2_approx(m - mn);      // 0 on the first key (m = -inf)
      const float p = ex2_approx(s - mn);
      l = l * corr + p;
      o[0] = o[0] * corr + p * bf16_lo(vv.x); o[1] = o[1] * corr + p * bf16_hi(vv.x);
      o[2] = o[2] * corr + p * bf16_lo(vv.y); o[3] = o[3] * corr + p * bf16_hi(vv.y);
      o[4] = o[4] * corr + p * bf16_lo(vv.z); o[5] = o[5] * corr + p * bf16_hi(vv.z);
      o[6] = o[6] * corr + p * bf16_lo(vv.w); o[7] = o[7] * corr + p * bf16_hi(vv.w);
      m = mn;
    }
  }
  // merge the NGROUPS partial softmaxes: every group publishes (m, l, o[HD]); thread (gl) of group 0 / warp 0 combines
  __shared__ float sm_m[NGROUPS], sm_l[NGROUPS], sm_o[NGROUPS][HD];
  if (active) {
    const int g = warp * KPW + grp;
    if (gl == 0) { sm_m[g] = m; sm_l[g] = l; }
#pragma unroll
    for (int i = 0; i < 8; ++i) sm_o[g][gl * 8 + i] = o[i];
  }
  __syncthreads();
  if (threadIdx.x < HD) {
    const int d = threadIdx.x;
    float M = -INFINITY;
    for (int g = 0; g < NGROUPS; ++g) M = fmaxf(M, sm_m[g]);
    float L = 0.f, acc = 0.f;
    for (int g = 0; g < NGROUPS; ++g) {
      const float w = sm_m[g] == -INFINITY ? 0.f : ex2_approx(sm_m[g] - M);
      L += w * sm_l[g];
      acc += w * sm_o[g][d];
    }
    out[(long long)bq * D + head * HD + d] = __float2bfloat16(acc / L);
  }
}

}  // namespace vj

using namespace vj;

extern "C" int vj_cross_attn_fwd(const void* q, const void* kv, void* out, int B, int nq, int S, int H, int HD, float scale,
                                 void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  VJ_CHECK_ARG(q && kv && out, "vj_cross_attn_fwd: null pointer");
  VJ_CHECK_ARG(B > 0 && nq > 0 && S > 0 && H > 0, "vj_cross_attn_fwd: empty problem");
  VJ_CHECK_ARG((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(kv) & 15) == 0,
               "vj_cross_attn_fwd: pointers must be 16-byte aligned");
  dim3 grid(H, B * nq);
  const float sl2 = scale * 1.4426950408889634f;
#define VJ_XATTN(LPK)                                                                                                    \
  xattn_fwd_kernel<LPK><<<grid, kXattnThreads, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(q),                        \
                                                       reinterpret_cast<const __nv_bfloat16*>(kv),                       \
                                                       reinterpret_cast<__nv_bfloat16*>(out), nq, S, H, sl2)
  switch (HD) {
    case 32: VJ_XATTN(4); break;
    case 64: VJ_XATTN(8); break;
    case 80: VJ_XATTN(10); break;
    case 128: VJ_XATTN(16); break;
    default: set_error("vj_cross_attn_fwd: head dim %d unsupported (32 / 64 / 80 / 128)", HD); return -1;
  }
#undef VJ_XATTN
  VJ_CUDA(cudaGetLastError());
  vj::count_launch(1);
  return 0;
}
