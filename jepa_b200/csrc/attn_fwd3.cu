// Dense var-len flash attention forward, third variant: the first-generation CTA (one 128-row query tile, two CTAs per
// SM, attn_fwd.cu) with EIGHT softmax warps - every query row is shared by two threads, each owning 64 of the 128 score
// columns of a KV tile (sm_100a, tcgen05 / TMEM / TMA).
//
// Replaces F.scaled_dot_product_attention at src/models/utils/modules.py:66-69; contract identical to attn_fwd.cu
// (qkv bf16 [T, 3*H*HD] q|k|v thirds head-major, out bf16 [T, H*HD], lse2 fp32 [H, T] log2 domain, cu_seqlens rows).
//
// Why (ncu, profiles/r02_ncu_attn_fwd*.txt): with one thread per row a softmax warp issues ~0.28 instructions per clock
// inside its exp pass - the FFMA -> MUFU -> FADD / F2FP chains of ONE warp per scheduler cannot cover the MUFU latency, so
// the 16 exp/clk/SM pipe that bounds attention at head dims <= 64 sits at 50-59 %.  Splitting the row
//   * doubles the warps per scheduler that are in an exp pass (4 with the co-resident CTA),
//   * halves the live score registers per thread (64): the whole S_j tile is in registers after ONE round of tcgen05.ld,
//     so S is handed back to the MMA warp BEFORE the exp pass (gen 1 re-read the last 32 columns from TMEM three quarters
//     into the pass) and S_{j+1} = Q K_{j+1}^T fully overlaps the exponentials,
//   * halves the serial per-thread work between "S ready" and "P ready" (the chain the PV MMA waits for).
// The two threads of a row agree on the row max through shared memory (one float each way + a 64-thread named barrier);
// the lazy-rescale decision is a function of that common max, so both take it together and each rescales its half of the
// O columns.  Scale / subtract and the row sums use packed fp32x2 arithmetic (FFMA2 / FADD2).
//
//   warp 0    : TMA producer (Q once; K double-buffered, V single-buffered 128-key tiles)
//   warp 1    : MMA issuer (S = Q K^T; O += P V with P read from TENSOR MEMORY)
//   warps 2-9 : softmax; warp w owns TMEM lane quarter w & 3 and score columns ((w-2)>>2)*64 .. +63
#include <stdlib.h>

#include "attn_common.cuh"
#include "vjepa_b200.h"

namespace vj {

constexpr int kFwd3Threads = 320;

struct AttnFwd3Params {
  const int* cu_seqlens;
  __nv_bfloat16* out;
  float* lse2;
  int H, T;
  long long ld_out;
  float scale_log2;
};

template <int HD>
struct Fwd3Cfg {
  using A = AttnCfg<HD>;
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = A::TILE_BYTES;                // 2 stages
  static constexpr int V_OFF = 3 * A::TILE_BYTES;
  static constexpr int STG_OFF = 4 * A::TILE_BYTES;          // output staging: 8 warps x [32 rows x HD/2 bf16]
  static constexpr int XCH_OFF = STG_OFF + 128 * HD * 2;     // [2 buffers][2 halves][128 rows] fp32: row max / row sum exchange
  static constexpr int BAR_OFF = XCH_OFF + 2048;
  static constexpr int SMEM_BYTES = BAR_OFF + 128 + 1024;
  static constexpr int TMEM_COLS = (128 + HD + 64) <= 256 ? 256 : 512;   // S (fp32) | O (fp32) | P (bf16 pairs)
};

// 64-thread named barrier of the two warps that share TMEM lane quarter qd.  Immediate barrier ids: with a register id
// ptxas reserves all 16 hardware barriers for the CTA and a second CTA no longer fits on the SM.
VJ_DEVINL void pair_sync(int qd) {
  switch (qd) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

template <int HD>
__global__ void __launch_bounds__(kFwd3Threads, HD <= 64 ? 2 : 1)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwd3Params p) {
  using C = AttnCfg<HD>;
  using F = Fwd3Cfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int seq = blockIdx.y, head = blockIdx.z;
  const int row_begin = p.cu_seqlens[seq];
  const int len = p.cu_seqlens[seq + 1] - row_begin;
  const int q0 = blockIdx.x * 128;
  if (q0 >= len) return;
  const int n_kv = (len + 127) / 128;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F::BAR_OFF);
  const uint32_t bar_q = smem_u32(bars + 0);
  const uint32_t bar_k0 = smem_u32(bars + 1);      // 2 stages: +0, +8
  const uint32_t bar_kfree0 = smem_u32(bars + 3);  // 2 stages
  const uint32_t bar_v = smem_u32(bars + 5);
  const uint32_t bar_vfree = smem_u32(bars + 6);   // also "PV_j retired": P columns reusable, O readable
  const uint32_t bar_s = smem_u32(bars + 7);
  const uint32_t bar_sfree = smem_u32(bars + 8);
  const uint32_t bar_p = smem_u32(bars + 9);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_k0, 1); mbar_init(bar_k0 + 8, 1);
    mbar_init(bar_kfree0, 1); mbar_init(bar_kfree0 + 8, 1);
    mbar_init(bar_v, 1); mbar_init(bar_vfree, 1);
    mbar_init(bar_s, 1); mbar_init(bar_sfree, 8); mbar_init(bar_p, 8);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 1) tmem_alloc<F::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;
  const uint32_t tmem_P = tmem_base + 128 + HD;   // P_j as bf16 pairs: the A operand of the PV MMA, never in smem

  const uint32_t sQ = smem_u32(smem + F::Q_OFF), sK = smem_u32(smem + F::K_OFF), sV = smem_u32(smem + F::V_OFF);
  const int HHD = p.H * HD;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_q, C::TILE_BYTES);
#pragma unroll
      for (int b = 0; b < C::NBOX; ++b)
        tma_load_2d(sQ + b * C::BOX_BYTES, &tmQKV, bar_q, head * HD + b * C::BOX_INNER, row_begin + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int kr = row_begin + j * 128;
        const int st = j & 1;
        const uint32_t use = uint32_t(j >> 1) & 1;   // per-stage phase
        mbar_wait(bar_kfree0 + 8 * st, use ^ 1);
        mbar_expect_tx(bar_k0 + 8 * st, C::TILE_BYTES);
#pragma unroll
        for (int b = 0; b < C::NBOX; ++b)
          tma_load_2d(sK + st * C::TILE_BYTES + b * C::BOX_BYTES, &tmQKV, bar_k0 + 8 * st,
                      HHD + head * HD + b * C::BOX_INNER, kr);
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
      const uint64_t dQ = kmajor_base<HD>(sQ), dK0 = kmajor_base<HD>(sK), dV = mnmajor_base<HD>(sV);
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        mbar_wait(bar_k0 + 8 * st, uint32_t(j >> 1) & 1);
        tc_fence_after();
        const uint64_t dk = desc_advance(dK0, uint32_t(st) * C::TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          umma_f16(tmem_S, desc_advance(dQ, kmajor_koff<HD>(kk)), desc_advance(dk, kmajor_koff<HD>(kk)), idesc_s, kk > 0);
        umma_commit(bar_kfree0 + 8 * st);
        umma_commit(bar_s);
      };
      mbar_wait(bar_q, 0);
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          mbar_wait(bar_sfree, j & 1);   // S_j is in the softmax threads' registers (before their exp pass)
          issue_qk(j + 1);
        }
        mbar_wait(bar_p, j & 1);
        mbar_wait(bar_v, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16_ts(tmem_O, tmem_P + kk * 8, desc_advance(dV, mnmajor_koff<HD>(kk)), idesc_o, (j > 0 || kk > 0));
        umma_commit(bar_vfree);
      }
    }
    __syncwarp();
  } else {
    const int qd = warp & 3;                 // TMEM lane quarter
    const int half = (warp - 2) >> 2;        // which 64 score columns of a KV tile (and which half of the O columns)
    const int r = qd * 32 + lane;            // query row inside the tile
    const int col0 = half * 64;
    const uint32_t lane_addr = uint32_t(qd * 32) << 16;
    const uint32_t tS = tmem_S + lane_addr + col0;
    const uint32_t tP = tmem_P + lane_addr + half * 32;
    const uint32_t xch = smem_u32(smem + F::XCH_OFF);
    const uint64_t scale2 = pk2(p.scale_log2, p.scale_log2);
    float m_ref = -INFINITY;                 // reference max the accumulators are expressed against (common to the pair)
    uint64_t lsum = pk2(0.f, 0.f);           // this thread's share of the row sum (two partial sums)
    for (int j = 0; j < n_kv; ++j) {
      const int valid = min(128, len - j * 128);
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      // ---- this thread's 64 scores -> registers, then S_j goes straight back to the MMA warp
      uint32_t s0[32], s1[32];
      tmem_ld32(tS, s0);
      tmem_ld32(tS + 32, s1);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sfree);
      float mx = -INFINITY;
      if (valid == 128) {
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          mx = fmaxf(fmaxf(mx, __uint_as_float(s0[i])),
                     fmaxf(__uint_as_float(s0[i + 1]), fmaxf(__uint_as_float(s1[i]), __uint_as_float(s1[i + 1]))));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (col0 + i < valid) mx = fmaxf(mx, __uint_as_float(s0[i]));
          if (col0 + 32 + i < valid) mx = fmaxf(mx, __uint_as_float(s1[i]));
        }
      }
      // ---- row max of the whole 128-column tile: exchange with the thread that owns the other half of this row
      const uint32_t xb = xch + (j & 1) * 1024;
      sts32f(xb + half * 512 + r * 4, mx);
      pair_sync(qd);
      mx = fmaxf(mx, lds32f(xb + (half ^ 1) * 512 + r * 4));
      // ---- lazy rescale: only move the reference max when it would overflow the 2^8 head-room
      const bool grow = (mx - m_ref) * p.scale_log2 > 8.0f;   // true on the first block (m_ref = -inf)
      if (__any_sync(0xffffffffu, grow)) {   // same rows, same maxima -> same decision in both warps of the pair
        if (j > 0) {
          mbar_wait(bar_vfree, (j - 1) & 1);   // PV_{j-1} retired: O is stable
          tc_fence_after();
          const float alpha = grow ? ex2_approx((m_ref - mx) * p.scale_log2) : 1.0f;
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t o[16];
            tmem_ld16(tmem_O + lane_addr + half * (HD / 2) + c * 16, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tmem_O + lane_addr + half * (HD / 2) + c * 16, o);
          }
          tmem_wait_st();
          lsum = mul2(lsum, pk2(alpha, alpha));
        }
        if (grow) m_ref = mx;
      }
      const float nmoff = -m_ref * p.scale_log2;
      const uint64_t nmoff2 = pk2(nmoff, nmoff);
      // ---- p = 2^(s*scale - m) -> bf16 pairs in TMEM; the P columns are free once PV_{j-1} retired
      if (j > 0) { mbar_wait(bar_vfree, (j - 1) & 1); tc_fence_after(); }
      if (valid == 128) {
        exp_store32<false>(s0, col0, valid, scale2, nmoff2, lsum, tP);
        exp_store32<false>(s1, col0 + 32, valid, scale2, nmoff2, lsum, tP + 16);
      } else {
        exp_store32<true>(s0, col0, valid, scale2, nmoff2, lsum, tP);
        exp_store32<true>(s1, col0 + 32, valid, scale2, nmoff2, lsum, tP + 16);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    // ---- epilogue: row sum of the pair, O / l -> bf16 -> per-warp staging -> coalesced 16-byte stores
    float la, lb;
    upk2(lsum, la, lb);
    float l = la + lb;
    const uint32_t xb = xch + (n_kv & 1) * 1024;   // the buffer the last iteration did not use
    sts32f(xb + half * 512 + r * 4, l);
    pair_sync(qd);
    l += lds32f(xb + (half ^ 1) * 512 + r * 4);
    mbar_wait(bar_vfree, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    if (half == 0 && q0 + r < len) p.lse2[(long long)head * p.T + row_begin + q0 + r] = m_ref * p.scale_log2 + log2f(l);
    constexpr int ORB = HD;                   // bytes per staged row: HD/2 bf16
    constexpr int CH = ORB / 16;              // 16-byte chunks per staged row (2 / 4 / 8)
    const uint32_t stage = smem_u32(smem + F::STG_OFF) + (warp - 2) * (32 * ORB);
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t o[16];
      tmem_ld16(tmem_O + lane_addr + half * (HD / 2) + c * 16, o);
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
                                  ((long long)(row_begin + grow_) * p.ld_out + head * HD) * 2 + half * HD + g * 16) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<F::TMEM_COLS>(tmem_base);
}

template <int HD>
int launch_attn_fwd3(const void* qkv, void* out, float* lse2, const int* cu, int nseq, int max_len, int H, int T,
                     float scale, cudaStream_t s) {
  using C = AttnCfg<HD>;
  using F = Fwd3Cfg<HD>;
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, qkv, 0, (uint64_t)3 * H * HD, T, (uint64_t)3 * H * HD * 2, C::BOX_INNER, 128,
                        C::TMAP_SWIZZLE);
  if (rc) return rc;
  auto kern = attn_fwd3_kernel<HD>;
  static bool configured = false;
  if (!configured) {
    VJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM_BYTES));
    configured = true;
  }
  AttnFwd3Params p;
  p.cu_seqlens = cu; p.out = reinterpret_cast<__nv_bfloat16*>(out); p.lse2 = lse2;
  p.H = H; p.T = T; p.ld_out = (long long)H * HD;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((max_len + 127) / 128, nseq, H);
  kern<<<grid, kFwd3Threads, F::SMEM_BYTES, s>>>(tm, p);
  VJ_CUDA(cudaGetLastError());
  vj::count_launch(1);
  return 0;
}

template int launch_attn_fwd3<32>(const void*, void*, float*, const int*, int, int, int, int, float, cudaStream_t);
template int launch_attn_fwd3<64>(const void*, void*, float*, const int*, int, int, int, int, float, cudaStream_t);
template int launch_attn_fwd3<128>(const void*, void*, float*, const int*, int, int, int, int, float, cudaStream_t);

}  // namespace vj
