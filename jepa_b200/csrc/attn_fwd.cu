// Dense (non-causal) variable-length flash attention forward on tcgen05 / TMEM (sm_100a).
//
// Replaces F.scaled_dot_product_attention at src/models/utils/modules.py:66-69 (the `mask`
// argument there is ignored - "masked attention" is attention over the gathered token subset,
// so every sequence is dense and only its length varies).
//
// Layout: qkv bf16 [T, 3*H*HD] (row = token, q|k|v packed, head-major inside each third - the
// layout the qkv GEMM epilogue writes), O bf16 [T, H*HD], lse2 fp32 [H, T] (log2 domain:
// m*scale*log2e + log2(l)).  Sequences are row ranges [cu[s], cu[s+1]).
//
// One CTA = one 128-row query tile of one (sequence, head):
//   warp 0 : TMA producer (Q once, then K_j / V_j tiles of 128 keys)
//   warp 1 : MMA issuer   S = Q K_j^T  (128x128xHD)  and  O_j = P_j V_j (128xHDx128), both into TMEM
//   warps 2-5 : one thread per query row: online softmax straight out of TMEM (tcgen05.ld), P_j
//               written as bf16 into a 128B-swizzled K-major smem tile, O accumulated in registers.
// Two CTAs are resident per SM (80 KB smem, 256 TMEM columns each for HD=64), so one CTA's
// softmax overlaps the other's MMAs.
#include "attn_common.cuh"
#include "vjepa_b200.h"

namespace vj {

struct AttnFwdParams {
  const int* cu_seqlens;
  __nv_bfloat16* out;
  float* lse2;
  int H, T;
  long long ld_out;
  float scale_log2;
};

template <int HD>
__global__ void __launch_bounds__(kAttnThreads, HD <= 64 ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwdParams p) {
  using C = AttnCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int seq = blockIdx.y, head = blockIdx.z;
  const int row_begin = p.cu_seqlens[seq];
  const int len = p.cu_seqlens[seq + 1] - row_begin;
  const int q0 = blockIdx.x * 128;
  if (q0 >= len) return;
  const int n_kv = (len + 127) / 128;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  const uint32_t bar_q = smem_u32(bars + 0);
  const uint32_t bar_k = smem_u32(bars + 1);
  const uint32_t bar_v = smem_u32(bars + 2);
  const uint32_t bar_kfree = smem_u32(bars + 3);
  const uint32_t bar_vfree = smem_u32(bars + 4);
  const uint32_t bar_s = smem_u32(bars + 5);
  const uint32_t bar_p = smem_u32(bars + 6);
  const uint32_t bar_o = smem_u32(bars + 7);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1);
    mbar_init(bar_kfree, 1); mbar_init(bar_vfree, 1);
    mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  const uint32_t sQ = smem_u32(smem + C::Q_OFF), sK = smem_u32(smem + C::K_OFF);
  const uint32_t sV = smem_u32(smem + C::V_OFF), sP = smem_u32(smem + C::P_OFF);
  const int HHD = p.H * HD;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_q, C::TILE_BYTES);
#pragma unroll
      for (int b = 0; b < C::NBOX; ++b)
        tma_load_2d(sQ + b * C::BOX_BYTES, &tmQKV, bar_q, head * HD + b * C::BOX_INNER, row_begin + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int kr = row_begin + j * 128;
        mbar_wait(bar_kfree, (j & 1) ^ 1);
        mbar_expect_tx(bar_k, C::TILE_BYTES);
#pragma unroll
        for (int b = 0; b < C::NBOX; ++b)
          tma_load_2d(sK + b * C::BOX_BYTES, &tmQKV, bar_k, HHD + head * HD + b * C::BOX_INNER, kr);
        mbar_wait(bar_vfree, (j & 1) ^ 1);
        mbar_expect_tx(bar_v, C::TILE_BYTES);
#pragma unroll
        for (int b = 0; b < C::NBOX; ++b)
          tma_load_2d(sV + b * C::BOX_BYTES, &tmQKV, bar_v, 2 * HHD + head * HD + b * C::BOX_INNER, kr);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);
      mbar_wait(bar_q, 0);
      for (int j = 0; j < n_kv; ++j) {
        // S = Q K_j^T
        mbar_wait(bar_k, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          umma_f16(tmem_S, kmajor_desc<HD>(sQ, kk), kmajor_desc<HD>(sK, kk), idesc_s, kk > 0);
        umma_commit(bar_kfree);
        umma_commit(bar_s);
        // O_j = P_j V_j
        mbar_wait(bar_p, j & 1);
        mbar_wait(bar_v, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16(tmem_O, ptile_desc(sP, kk), mnmajor_desc<HD>(sV, kk), idesc_o, kk > 0);
        umma_commit(bar_vfree);
        umma_commit(bar_o);
      }
    }
    __syncwarp();
  } else {
    const int qd = warp & 3;                 // TMEM lane quarter
    const int r = qd * 32 + lane;            // query row inside the tile
    const uint32_t lane_addr = uint32_t(qd * 32) << 16;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    uint8_t* prow = smem + C::P_OFF + r * 128;
    for (int j = 0; j < n_kv; ++j) {
      const int valid = min(128, len - j * 128);
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      // pass 1: row max
      float mx = m;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float alpha = exp2f((m - mx) * p.scale_log2);  // m = -inf on the first block -> 0
      const float moff = mx * p.scale_log2;
      m = mx;
      l *= alpha;
      // pass 2: p = exp2(s*scale - m*scale) -> bf16 -> swizzled smem
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + c * 32, v);
        tmem_wait_ld();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2, -moff));
          pv[i] = (c * 32 + i < valid) ? e : 0.f;
          l += pv[i];
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(pv[8 * g + 0], pv[8 * g + 1]);
          u.y = pack_bf16x2(pv[8 * g + 2], pv[8 * g + 3]);
          u.z = pack_bf16x2(pv[8 * g + 4], pv[8 * g + 5]);
          u.w = pack_bf16x2(pv[8 * g + 6], pv[8 * g + 7]);
          const int col8 = c * 4 + g;  // 16-byte chunk index along the 128 kv columns
          *reinterpret_cast<uint4*>(prow + (col8 >> 3) * 16384 + (((col8 & 7) ^ (r & 7)) << 4)) = u;
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(bar_p);
      // O accumulate
      mbar_wait(bar_o, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_O + lane_addr + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, __uint_as_float(v[i]));
      }
      tc_fence_before();
    }
    // epilogue: O / l -> bf16, staged through the (now idle) P tile for coalesced stores
    const float inv = 1.0f / l;
    const bool row_ok = q0 + r < len;
    if (row_ok) p.lse2[(long long)head * p.T + row_begin + q0 + r] = m * p.scale_log2 + log2f(l);
    constexpr int ORB = HD * 2;               // bytes per output row
    constexpr int CH = ORB / 16;              // 16-byte chunks per row
    uint8_t* stage = smem + C::P_OFF + (warp - 2) * (32 * ORB);
#pragma unroll
    for (int g = 0; g < CH; ++g) {
      uint4 u;
      u.x = pack_bf16x2(o[8 * g + 0] * inv, o[8 * g + 1] * inv);
      u.y = pack_bf16x2(o[8 * g + 2] * inv, o[8 * g + 3] * inv);
      u.z = pack_bf16x2(o[8 * g + 4] * inv, o[8 * g + 5] * inv);
      u.w = pack_bf16x2(o[8 * g + 6] * inv, o[8 * g + 7] * inv);
      *reinterpret_cast<uint4*>(stage + lane * ORB + ((g ^ (lane & (CH - 1))) << 4)) = u;
    }
    __syncwarp();
    // coalesced write-out: CH lanes cover one row
    constexpr int ROWS_PER_IT = 32 / CH;
#pragma unroll
    for (int it = 0; it < CH; ++it) {
      const int rr = it * ROWS_PER_IT + lane / CH;
      const int g = lane % CH;
      const int grow = q0 + qd * 32 + rr;
      if (grow < len) {
        const uint4 u = *reinterpret_cast<const uint4*>(stage + rr * ORB + ((g ^ (rr & (CH - 1))) << 4));
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) +
                                  ((long long)(row_begin + grow) * p.ld_out + head * HD) * 2 + g * 16) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int HD>
static int launch_attn_fwd(const void* qkv, void* out, float* lse2, const int* cu, int nseq, int max_len, int H, int T,
                           float scale, cudaStream_t s) {
  using C = AttnCfg<HD>;
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, qkv, 0, (uint64_t)3 * H * HD, T, (uint64_t)3 * H * HD * 2, C::BOX_INNER, 128,
                        C::TMAP_SWIZZLE);
  if (rc) return rc;
  auto kern = attn_fwd_kernel<HD>;
  static bool configured = false;
  if (!configured) {
    VJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  AttnFwdParams p;
  p.cu_seqlens = cu; p.out = reinterpret_cast<__nv_bfloat16*>(out); p.lse2 = lse2;
  p.H = H; p.T = T; p.ld_out = (long long)H * HD;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((max_len + 127) / 128, nseq, H);
  kern<<<grid, kAttnThreads, C::SMEM_BYTES, s>>>(tm, p);
  VJ_CUDA(cudaGetLastError());
  vj::count_launch(1);
  return 0;
}

}  // namespace vj

extern "C" int vj_attn_fwd(const void* qkv, void* out, float* lse2, const int* cu_seqlens, int nseq, int max_len,
                           int H, int HD, int T, float scale, void* stream_) {
  using namespace vj;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  VJ_CHECK_ARG(qkv && out && lse2 && cu_seqlens, "vj_attn_fwd: null pointer");
  VJ_CHECK_ARG(nseq > 0 && max_len > 0 && H > 0 && T > 0, "vj_attn_fwd: empty problem");
  VJ_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "vj_attn_fwd: pointers must be 16-byte aligned");
  switch (HD) {
    case 32: return launch_attn_fwd<32>(qkv, out, lse2, cu_seqlens, nseq, max_len, H, T, scale, s);
    case 64: return launch_attn_fwd<64>(qkv, out, lse2, cu_seqlens, nseq, max_len, H, T, scale, s);
    case 128: return launch_attn_fwd<128>(qkv, out, lse2, cu_seqlens, nseq, max_len, H, T, scale, s);
    default: set_error("vj_attn_fwd: head dim %d unsupported (32/64/128; pad 24->32 in the weights)", HD); return -1;
  }
}
