// Dense var-len flash attention forward, fourth variant (head dims 32 / 64): small, fully serial CTAs - one 128-row
// query tile x 64-key KV tiles, 128 TMEM columns - FOUR of them resident per SM (sm_100a, tcgen05 / TMEM / TMA).
//
// Replaces F.scaled_dot_product_attention at src/models/utils/modules.py:66-69; contract identical to attn_fwd.cu
// (qkv bf16 [T, 3*H*HD] q|k|v thirds head-major, out bf16 [T, H*HD], lse2 fp32 [H, T] log2 domain, cu_seqlens rows).
//
// Why (profiles/r02_ncu_attn_fwd{1,3}.txt): at head dims <= 64 the binding pipe is MUFU (16 exp/clk/SM; the tensor pipe
// needs half that time), and inside a CTA the work is one serial chain  S = Q K^T -> softmax -> O += P V -> next S:
// whatever the softmax organisation (one thread per row, two threads per row with early S hand-back), two CTAs per SM
// left MUFU 59 % busy because all softmax warps of a CTA are phase-locked by the CTA's own barriers, so only two
// independent phases exist per SM.  Here the CTA is as simple and as small as possible and the SM runs FOUR of them: the
// exp passes of the other three CTAs fill the S / PV / barrier latency of each chain.
//   * KV tile = 64 keys: S is 64 fp32 columns, P (bf16 pairs, 32 columns) overwrites the first half of S once every
//     thread has its score row in registers, O takes HD columns -> 128 TMEM columns per CTA (4 x 128 = 512);
//   * one MMA group per iteration: PV_j (A = P from tensor memory) then Q K_{j+1}^T, one commit;
//   * K and V rings (2 stages of 8 KB at hd 64, 4 stages of 4 KB at hd 32), 48 KB of shared memory per CTA;
//   * 192 threads: warp 0 TMA, warp 1 MMA, warps 2-5 softmax (one thread per query row, 64 live scores), 80 registers;
//   * packed fp32x2 scale / subtract and row sums (FFMA2 / FADD2), lazy rescale of the TMEM-resident O (2^8 head-room).
#include <stdlib.h>

#include "attn_common.cuh"
#include "vjepa_b200.h"

namespace vj {

constexpr int kFwd4Threads = 192;
constexpr int kFwd4KT = 64;          // keys per KV tile
constexpr int kFwd4DefaultPoly = 0;  // set from measurements (profiles/)

struct AttnFwd4Params {
  const int* cu_seqlens;
  __nv_bfloat16* out;
  float* lse2;
  int H, T;
  long long ld_out;
  float scale_log2;
};

template <int HD, int NCTA>
struct Fwd4Cfg {
  using A = AttnCfg<HD>;
  static_assert(HD == 32 || HD == 64 || HD == 128, "attn_fwd4: head dims 32 / 64 / 128");
  static constexpr int ST = HD == 32 ? 4 : (HD == 64 && NCTA == 3 ? 3 : 2);   // K / V ring depth
  static constexpr int Q_BYTES = A::TILE_BYTES;                     // [128 x HD]
  static constexpr int KV_BOX_BYTES = kFwd4KT * A::ROW_BYTES;       // one TMA box of a K / V tile: 64 rows x <= 64 columns
  static constexpr int KV_BYTES = A::NBOX * KV_BOX_BYTES;           // [64 x HD]
  static constexpr int Q_OFF = 0;                                   // also the output staging area
  static constexpr int K_OFF = Q_BYTES;
  static constexpr int V_OFF = K_OFF + ST * KV_BYTES;
  static constexpr int BAR_OFF = V_OFF + ST * KV_BYTES;
  static constexpr int NBARS = 3 + 4 * ST;
  static constexpr int SMEM_BYTES = BAR_OFF + NBARS * 8 + 16 + 1024;
  static constexpr int TM_S = 0, TM_P = 0, TM_O = 64;               // P aliases S columns 0..31
  static constexpr int TMEM_COLS = HD <= 64 ? 128 : 256;
  // byte offset of reduction step kk (16 head-dim elements) inside a K-major [64 x HD] K tile
  __host__ __device__ static constexpr uint32_t k_koff(int kk) {
    return uint32_t((kk / (A::BOX_INNER / 16)) * KV_BOX_BYTES + (kk % (A::BOX_INNER / 16)) * 32);
  }
};

template <int HD, int NCTA, int POLY>
__global__ void __launch_bounds__(kFwd4Threads, NCTA)
attn_fwd4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnFwd4Params p) {
  using C = AttnCfg<HD>;
  using F = Fwd4Cfg<HD, NCTA>;
  constexpr int ST = F::ST, KT = kFwd4KT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int seq = blockIdx.y, head = blockIdx.z;
  const int row_begin = p.cu_seqlens[seq];
  const int len = p.cu_seqlens[seq + 1] - row_begin;
  const int q0 = blockIdx.x * 128;
  if (q0 >= len) return;
  const int n_kv = (len + KT - 1) / KT;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F::BAR_OFF);
  const uint32_t bar_q = smem_u32(bars + 0);
  const uint32_t bar_s = smem_u32(bars + 1);        // MMA group j retired: S_j readable (and PV_{j-1} done)
  const uint32_t bar_p = smem_u32(bars + 2);        // P_j written by the four softmax warps
  const uint32_t k_full0 = smem_u32(bars + 3), k_free0 = k_full0 + 8 * ST;
  const uint32_t v_full0 = k_free0 + 8 * ST, v_free0 = v_full0 + 8 * ST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + F::NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 4);
    for (int i = 0; i < ST; ++i) {
      mbar_init(k_full0 + 8 * i, 1); mbar_init(k_free0 + 8 * i, 1);
      mbar_init(v_full0 + 8 * i, 1); mbar_init(v_free0 + 8 * i, 1);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV); }
  if (warp == 1) tmem_alloc<F::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sQ = smem_u32(smem + F::Q_OFF), sK = smem_u32(smem + F::K_OFF), sV = smem_u32(smem + F::V_OFF);
  const int HHD = p.H * HD;

  if (warp == 0) {
    // ------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(bar_q, F::Q_BYTES);
#pragma unroll
      for (int b = 0; b < C::NBOX; ++b)
        tma_load_2d(sQ + b * C::BOX_BYTES, &tmQ, bar_q, head * HD + b * C::BOX_INNER, row_begin + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int kr = row_begin + j * KT;
        const int st = j % ST;
        const uint32_t ph = uint32_t(j / ST) & 1;
        mbar_wait(k_free0 + 8 * st, ph ^ 1);
        mbar_expect_tx(k_full0 + 8 * st, F::KV_BYTES);
#pragma unroll
        for (int b = 0; b < C::NBOX; ++b)
          tma_load_2d(sK + st * F::KV_BYTES + b * F::KV_BOX_BYTES, &tmKV, k_full0 + 8 * st, HHD + head * HD + b * C::BOX_INNER, kr);
        mbar_wait(v_free0 + 8 * st, ph ^ 1);
        mbar_expect_tx(v_full0 + 8 * st, F::KV_BYTES);
#pragma unroll
        for (int b = 0; b < C::NBOX; ++b)
          tma_load_2d(sV + st * F::KV_BYTES + b * F::KV_BOX_BYTES, &tmKV, v_full0 + 8 * st, 2 * HHD + head * HD + b * C::BOX_INNER,
                      kr);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, KT, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);
      const uint64_t dQ = kmajor_base<HD>(sQ), dK0 = kmajor_base<HD>(sK);
      const uint64_t dV0 = make_smem_desc(sV, F::KV_BOX_BYTES, C::SBO, C::LAYOUT);   // MN-major: LBO = stride of the 64-column boxes
      const uint32_t tS = tmem_base + F::TM_S, tO = tmem_base + F::TM_O, tP = tmem_base + F::TM_P;
      auto issue_qk = [&](int j) {
        const int st = j % ST;
        mbar_wait(k_full0 + 8 * st, uint32_t(j / ST) & 1);
        tc_fence_after();
        const uint64_t dk = desc_advance(dK0, uint32_t(st) * F::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          umma_f16(tS, desc_advance(dQ, kmajor_koff<HD>(kk)), desc_advance(dk, F::k_koff(kk)), idesc_s, kk > 0);
        umma_commit(k_free0 + 8 * st);
      };
      mbar_wait(bar_q, 0);
      issue_qk(0);
      umma_commit(bar_s);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % ST;
        mbar_wait(bar_p, j & 1);
        mbar_wait(v_full0 + 8 * st, uint32_t(j / ST) & 1);
        tc_fence_after();
        const uint64_t dv = desc_advance(dV0, uint32_t(st) * F::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk)
          umma_f16_ts(tO, tP + kk * 8, desc_advance(dv, mnmajor_koff<HD>(kk)), idesc_o, (j > 0 || kk > 0));
        umma_commit(v_free0 + 8 * st);
        if (j + 1 < n_kv) issue_qk(j + 1);   // executes after PV_j (issue order): P_j is consumed before S_{j+1} lands on it
        umma_commit(bar_s);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------ softmax: one thread per query row
    const int qd = warp & 3;                 // TMEM lane quarter
    const int r = qd * 32 + lane;            // query row inside the tile
    const uint32_t lane_addr = uint32_t(qd * 32) << 16;
    const uint32_t tS = tmem_base + F::TM_S + lane_addr;
    const uint32_t tO = tmem_base + F::TM_O + lane_addr;
    const uint64_t scale2 = pk2(p.scale_log2, p.scale_log2);
    float m_ref = -INFINITY;                 // reference max the accumulators are expressed against
    uint64_t lsum = pk2(0.f, 0.f);
    for (int j = 0; j < n_kv; ++j) {
      const int valid = min(KT, len - j * KT);
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld32(tS, s0);
      tmem_ld32(tS + 32, s1);
      tmem_wait_ld();
      float mx = -INFINITY;
      if (valid == KT) {
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          mx = fmaxf(fmaxf(mx, __uint_as_float(s0[i])),
                     fmaxf(__uint_as_float(s0[i + 1]), fmaxf(__uint_as_float(s1[i]), __uint_as_float(s1[i + 1]))));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < valid) mx = fmaxf(mx, __uint_as_float(s0[i]));
          if (32 + i < valid) mx = fmaxf(mx, __uint_as_float(s1[i]));
        }
      }
      // ---- lazy rescale (bar_s of this iteration was committed after PV_{j-1}: O is stable)
      const bool grow = (mx - m_ref) * p.scale_log2 > 8.0f;   // true on the first tile (m_ref = -inf)
      if (__any_sync(0xffffffffu, grow)) {
        if (j > 0) {
          const float alpha = grow ? ex2_approx((m_ref - mx) * p.scale_log2) : 1.0f;
#pragma unroll
          for (int c = 0; c < HD / 16; ++c) {
            uint32_t o[16];
            tmem_ld16(tO + c * 16, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tO + c * 16, o);
          }
          lsum = mul2(lsum, pk2(alpha, alpha));
        }
        if (grow) m_ref = mx;
      }
      const float nmoff = -m_ref * p.scale_log2;
      const uint64_t nmoff2 = pk2(nmoff, nmoff);
      // ---- p = 2^(s*scale - m) -> bf16 pairs over the first 32 columns of S (every score of this row has been seen).
      // Columns 32..63 are not kept in registers across the first half of the pass: P only overwrites columns 0..31, so
      // they are simply read again from TMEM (issued half-way through the first 32 exponentials, landed by their end) -
      // at 80 registers per thread (four CTAs per SM) 64 live scores spill.
      if (valid == KT) {
        exp_store32<false, POLY, 0, 2>(s0, 0, valid, scale2, nmoff2, lsum, tS);
        tmem_ld32(tS + 32, s1);
        exp_store32<false, POLY, 2, 4>(s0, 0, valid, scale2, nmoff2, lsum, tS);
        tmem_wait_ld();
        exp_store32<false, POLY>(s1, 32, valid, scale2, nmoff2, lsum, tS + 16);
      } else {
        exp_store32<true, POLY, 0, 2>(s0, 0, valid, scale2, nmoff2, lsum, tS);
        tmem_ld32(tS + 32, s1);
        exp_store32<true, POLY, 2, 4>(s0, 0, valid, scale2, nmoff2, lsum, tS);
        tmem_wait_ld();
        exp_store32<true, POLY>(s1, 32, valid, scale2, nmoff2, lsum, tS + 16);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    // ---- epilogue: O / l -> bf16 -> staging (the Q tile, dead once the last Q K^T retired) -> coalesced 16-byte stores
    mbar_wait(bar_s, n_kv & 1);
    tc_fence_after();
    float la, lb;
    upk2(lsum, la, lb);
    const float l = la + lb;
    const float inv = 1.0f / l;
    if (q0 + r < len) p.lse2[(long long)head * p.T + row_begin + q0 + r] = m_ref * p.scale_log2 + log2f(l);
    constexpr int ORB = HD * 2;               // bytes per output row
    constexpr int CH = ORB / 16;              // 16-byte chunks per row
    const uint32_t stage = sQ + qd * (32 * ORB);
#pragma unroll
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t o[16];
      tmem_ld16(tO + c * 16, o);
      tmem_wait_ld();
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(o[8 * h2 + 0]) * inv, __uint_as_float(o[8 * h2 + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(o[8 * h2 + 2]) * inv, __uint_as_float(o[8 * h2 + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(o[8 * h2 + 4]) * inv, __uint_as_float(o[8 * h2 + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(o[8 * h2 + 6]) * inv, __uint_as_float(o[8 * h2 + 7]) * inv);
        const int g = 2 * c + h2;
        sts128(stage + lane * ORB + ((g ^ (lane & (CH - 1))) << 4), u);
      }
    }
    tc_fence_before();
    __syncwarp();
    constexpr int ROWS_PER_IT = 32 / CH;
#pragma unroll
    for (int it = 0; it < CH; ++it) {
      const int rr = it * ROWS_PER_IT + lane / CH;
      const int g = lane % CH;
      const int grow_ = q0 + qd * 32 + rr;
      if (grow_ < len) {
        const uint4 u = lds128(stage + rr * ORB + ((g ^ (rr & (CH - 1))) << 4));
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) +
                                  ((long long)(row_begin + grow_) * p.ld_out + head * HD) * 2 + g * 16) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<F::TMEM_COLS>(tmem_base);
}

template <int HD, int NCTA>
int launch_attn_fwd4(const void* qkv, void* out, float* lse2, const int* cu, int nseq, int max_len, int H, int T,
                     float scale, cudaStream_t s) {
  using C = AttnCfg<HD>;
  using F = Fwd4Cfg<HD, NCTA>;
  CUtensorMap tmq, tmkv;
  int rc = make_tmap_2d(&tmq, qkv, 0, (uint64_t)3 * H * HD, T, (uint64_t)3 * H * HD * 2, C::BOX_INNER, 128, C::TMAP_SWIZZLE);
  if (rc) return rc;
  rc = make_tmap_2d(&tmkv, qkv, 0, (uint64_t)3 * H * HD, T, (uint64_t)3 * H * HD * 2, C::BOX_INNER, kFwd4KT, C::TMAP_SWIZZLE);
  if (rc) return rc;
  // VJ_ATTN_POLY = exponentials out of every 8 that run on the FMA pipe instead of MUFU (0 or 2)
  static int poly = -1;
  if (poly < 0) { const char* e = getenv("VJ_ATTN_POLY"); poly = e ? atoi(e) : kFwd4DefaultPoly; }
  void (*kern)(const CUtensorMap, const CUtensorMap, const AttnFwd4Params) = attn_fwd4_kernel<HD, NCTA, 0>;
  if (poly == 2) kern = attn_fwd4_kernel<HD, NCTA, 2>;
  static void* configured = nullptr;
  if (configured != reinterpret_cast<void*>(kern)) {
    VJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM_BYTES));
    configured = reinterpret_cast<void*>(kern);
  }
  AttnFwd4Params p;
  p.cu_seqlens = cu; p.out = reinterpret_cast<__nv_bfloat16*>(out); p.lse2 = lse2;
  p.H = H; p.T = T; p.ld_out = (long long)H * HD;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((max_len + 127) / 128, nseq, H);
  kern<<<grid, kFwd4Threads, F::SMEM_BYTES, s>>>(tmq, tmkv, p);
  VJ_CUDA(cudaGetLastError());
  vj::count_launch(1);
  return 0;
}

template int launch_attn_fwd4<128, 2>(const void*, void*, float*, const int*, int, int, int, int, float, cudaStream_t);
template int launch_attn_fwd4<32, 4>(const void*, void*, float*, const int*, int, int, int, int, float, cudaStream_t);
template int launch_attn_fwd4<64, 4>(const void*, void*, float*, const int*, int, int, int, int, float, cudaStream_t);
template int launch_attn_fwd4<32, 3>(const void*, void*, float*, const int*, int, int, int, int, float, cudaStream_t);
template int launch_attn_fwd4<64, 3>(const void*, void*, float*, const int*, int, int, int, int, float, cudaStream_t);

}  // namespace vj
