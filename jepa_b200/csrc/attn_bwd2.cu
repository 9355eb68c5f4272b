// Flash-attention backward, second generation (head dims <= 32: the predictor's 24 -> 32 heads, 12 layers x S ~ 1190 -
// the largest single attention cost of a V-JEPA step): one PERSISTENT CTA per SM that owns TWO 128-key tiles at a time.
//
// Backward of F.scaled_dot_product_attention (src/models/utils/modules.py:66-69), same math and layouts as attn_bwd.cu:
//   P = softmax(scale Q K^T), dV = P^T dO, dP = dO V^T, dS = P o (dP - delta), dQ = scale dS K, dK = scale dS^T Q.
// Per key tile kt in {0,1} the CTA walks the query tiles i of the sequence:
//   S^T = K Q_i^T -> P^T = 2^(S^T scale - lse[q]) -> dV += P^T dO_i ; dP^T = V dO_i^T -> dS^T = P^T o (dP^T - delta[q])
//   -> dK += dS^T Q_i ; dQ_i += dS K   (both key tiles accumulate into ONE dQ partial, TMA reduce-added once per i)
// What changed against the first generation (one key tile per CTA, two CTAs per SM, 3400 clk per 128x128 tile):
//   * the two key-tile chains run in one CTA in an enforced ping-pong (named-barrier token around the exp pass): exactly
//     one of the two 8-warp softmax groups exponentiates at a time while the other does its dS pass / waits for MMAs,
//     instead of two co-resident CTAs drifting into the same phase and contending for MUFU, then idling together;
//   * P^T and dS^T are handed to the dV / dK MMAs through TENSOR MEMORY (A operand, bf16 pairs written in place over the
//     columns of S^T / dP^T each warp has just read): a [128 x 128] bf16 A tile read from shared memory costs 32 KB of
//     shared-memory bandwidth per MMA, 2.5x the tensor time of an N = 32 MMA.  Only dS^T still goes to shared memory
//     (the fused dQ MMA needs it M-major);
//   * Q / dO tiles, the softmax statistics and the dQ partial (one reduce-add per query tile instead of two) are shared
//     by the two key tiles; K / V of the next work item are prefetched while the current one finishes;
//   * query tiles past the end of the sequence tail skip their exponentials and dV / dK reduction steps.
//
//   warps 0-7     softmax group of key tile 0: warp (qd, half) = key rows qd*32.. x query columns half*64..+64
//   warps 8-15    softmax group of key tile 1
//   warp 16       TMA producer (+ stages lse2 / delta rows of every query tile in a 2-deep smem ring)
//   warp 17       MMA issuer: one thread issues all ~56 MMAs of a tile pair (most of them N = 32, i.e. 16 tensor clocks
//                 each) - it owns the highest warp id (first pick of its scheduler) and only adds compile-time offsets to
//                 per-tile descriptors; every k-step loop is fully unrolled with a predicate
// TMEM (512 columns): S^T/dP^T of kt 0 | of kt 1 | dV0 dK0 dV1 dK1 | dQ[2] (ping-pong over query tiles).
#include <stdlib.h>

#include "attn_common.cuh"
#include "vjepa_b200.h"

namespace vj {

constexpr int kBwd2Threads = 64 + 16 * 32;

struct AttnBwd2Params {
  const int* cu_seqlens;
  const float* lse2;
  const float* delta;
  __nv_bfloat16* dqkv;
  int H, T, nseq, kpairs, n_items;
  int pingpong;     // 1: the two groups take turns in the exp pass; 0: free running
  int debug;        // VJ_BWD2_DEBUG bring-up switches
  float scale, scale_log2;
};

template <int HD>
struct Bwd2Cfg {
  using A = AttnCfg<HD>;
  static_assert(HD <= 32, "attn_bwd2: TMEM budget is laid out for head dims <= 32");
  static constexpr int TILE = A::TILE_BYTES;
  static constexpr int KV_OFF = 0;                            // [2 buffers][2 key tiles][K, V]
  static constexpr int QDO_OFF = KV_OFF + 8 * TILE;           // [2 stages][Q, dO]
  static constexpr int DS_OFF = QDO_OFF + 4 * TILE;           // [2 key tiles] dS^T tiles, 32 KB each
  static constexpr int STAT_OFF = DS_OFF + 2 * A::P_BYTES;    // [2 slots][lse 128 | delta 128] floats
  static constexpr int DQS_OFF = STAT_OFF + 2048;             // [2 groups][4 warps] x [32 rows x 32 fp32] dQ staging (each
                                                              // warp only waits for ITS OWN earlier TMA reduce-add)
  static constexpr int BAR_OFF = DQS_OFF + 8 * 4096;
  static constexpr int NBARS = 34;
  static constexpr int SMEM_BYTES = BAR_OFF + NBARS * 8 + 16 + 1024;
  static_assert(SMEM_BYTES <= 232448, "attn_bwd2 shared memory budget exceeded");
  static constexpr int TM_ST0 = 0, TM_ST1 = 128, TM_DV0 = 256, TM_DK0 = 256 + HD, TM_DV1 = 256 + 2 * HD,
                       TM_DK1 = 256 + 3 * HD, TM_DQ0 = 256 + 4 * HD, TM_DQ1 = 256 + 5 * HD;
  static_assert(TM_DQ1 + HD <= 512, "TMEM overflow");
};

VJ_DEVINL void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
VJ_DEVINL void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int HD>
__global__ void __launch_bounds__(kBwd2Threads, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmDQ, const AttnBwd2Params p) {
  using C = AttnCfg<HD>;
  using B = Bwd2Cfg<HD>;
  constexpr int TILE = B::TILE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B::BAR_OFF);
  const uint32_t b0 = smem_u32(bars);
  auto kv_full = [&](int buf, int kt) { return b0 + 8 * (buf * 2 + kt); };            // 0..3
  auto kv_free = [&](int buf, int kt) { return b0 + 8 * (4 + buf * 2 + kt); };        // 4..7
  auto qdo_full = [&](int st) { return b0 + 8 * (8 + st); };                          // 8..9
  auto qdo_free = [&](int st) { return b0 + 8 * (10 + st); };                         // 10..11
  auto stat_full = [&](int sb) { return b0 + 8 * (12 + sb); };                        // 12..13
  auto stat_free = [&](int g, int sb) { return b0 + 8 * (14 + g * 2 + sb); };         // 14..17
  auto s_full = [&](int kt) { return b0 + 8 * (18 + kt); };
  auto p_full = [&](int kt) { return b0 + 8 * (20 + kt); };
  auto dp_full = [&](int kt) { return b0 + 8 * (22 + kt); };
  auto ds_full = [&](int kt) { return b0 + 8 * (24 + kt); };
  auto ps_free = [&](int kt) { return b0 + 8 * (26 + kt); };
  auto dq_done = [&](int b) { return b0 + 8 * (28 + b); };
  auto dq_free = [&](int b) { return b0 + 8 * (30 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B::NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(kv_full(0, 0) + 8 * i, 1); mbar_init(kv_free(0, 0) + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(qdo_full(i), 1); mbar_init(qdo_free(i), 1); mbar_init(stat_full(i), 1);
      mbar_init(stat_free(0, i), 8); mbar_init(stat_free(1, i), 8);
      mbar_init(s_full(i), 1); mbar_init(p_full(i), 8); mbar_init(dp_full(i), 1); mbar_init(ds_full(i), 8);
      mbar_init(ps_free(i), 1); mbar_init(dq_done(i), 1); mbar_init(dq_free(i), 4);
    }
    fence_mbar_init();
  }
  if (warp == 16 && lane == 0) { tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmDQ); }
  if (warp == 17) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sKV = smem_u32(smem + B::KV_OFF), sQDO = smem_u32(smem + B::QDO_OFF), sDS = smem_u32(smem + B::DS_OFF);
  const int HHD = p.H * HD;

  struct Item { int row_begin, len, kv0, n_kt, n_q, head; };
  auto decode = [&](int item, Item& it) -> bool {
    const int kp = item % p.kpairs;
    const int rest = item / p.kpairs;
    const int seq = rest % p.nseq;
    it.head = rest / p.nseq;
    it.row_begin = p.cu_seqlens[seq];
    it.len = p.cu_seqlens[seq + 1] - it.row_begin;
    it.kv0 = kp * 256;
    if (it.kv0 >= it.len) return false;
    it.n_kt = (it.kv0 + 128 < it.len) ? 2 : 1;
    it.n_q = (it.len + 127) >> 7;
    return true;
  };

  if (warp == 16) {
    // ------------------------------------------------------------------------------ TMA producer + statistics
    int uk0 = 0, uk1 = 0;          // K/V loads so far per key tile
    int qi = 0;                    // query tiles streamed so far (Q/dO stage = qi & 1, statistics slot = qi & 1)
    int slot_nkt[2] = {0, 0};      // how many groups consumed each statistics slot at its previous use
    int slot_use[2][2] = {{0, 0}, {0, 0}};   // [group][slot] completed uses (phase of stat_free)
    const uint32_t stats_w = smem_u32(smem + B::STAT_OFF);
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      Item it;
      if (!decode(item, it)) continue;
      if (lane == 0) {
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          if (kt >= it.n_kt) continue;
          const int uk = kt ? uk1 : uk0;
          const int buf = uk & 1;
          mbar_wait(kv_free(buf, kt), (uint32_t(uk >> 1) & 1) ^ 1);
          mbar_expect_tx(kv_full(buf, kt), 2 * TILE);
          const uint32_t dst = sKV + (buf * 2 + kt) * 2 * TILE;
#pragma unroll
          for (int b = 0; b < C::NBOX; ++b) {
            tma_load_2d(dst + b * C::BOX_BYTES, &tmQKV, kv_full(buf, kt), HHD + it.head * HD + b * C::BOX_INNER,
                        it.row_begin + it.kv0 + kt * 128);
            tma_load_2d(dst + TILE + b * C::BOX_BYTES, &tmQKV, kv_full(buf, kt), 2 * HHD + it.head * HD + b * C::BOX_INNER,
                        it.row_begin + it.kv0 + kt * 128);
          }
        }
      }
      if (it.n_kt > 0) uk0++;
      if (it.n_kt > 1) uk1++;
      for (int i = 0; i < it.n_q; ++i, ++qi) {
        const int st = qi & 1;
        if (lane == 0) {   // Q_i / dO_i
          mbar_wait(qdo_free(st), (uint32_t(qi >> 1) & 1) ^ 1);
          mbar_expect_tx(qdo_full(st), 2 * TILE);
          const uint32_t dst = sQDO + st * 2 * TILE;
#pragma unroll
          for (int b = 0; b < C::NBOX; ++b) {
            tma_load_2d(dst + b * C::BOX_BYTES, &tmQKV, qdo_full(st), it.head * HD + b * C::BOX_INNER, it.row_begin + i * 128);
            tma_load_2d(dst + TILE + b * C::BOX_BYTES, &tmDO, qdo_full(st), it.head * HD + b * C::BOX_INNER,
                        it.row_begin + i * 128);
          }
        }
        // per-query-row softmax statistics of tile i -> smem slot (qi & 1); +inf LSE for rows past the sequence end makes
        // their probabilities exactly 0 without predicates.  The slot is free once every group that read it at its
        // previous use has said so.
        const int sb = qi & 1;
#pragma unroll
        for (int g = 0; g < 2; ++g)
          if (g < slot_nkt[sb]) mbar_wait(stat_free(g, sb), uint32_t(slot_use[g][sb] - 1) & 1);
        float lv[4], dv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int qrow = i * 128 + k * 32 + lane;
          const bool ok = qrow < it.len;
          const long long gidx = (long long)it.head * p.T + it.row_begin + qrow;
          lv[k] = ok ? p.lse2[gidx] : INFINITY;
          dv[k] = ok ? p.delta[gidx] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sts32f(stats_w + sb * 1024 + 4 * (k * 32 + lane), lv[k]);
          sts32f(stats_w + sb * 1024 + 512 + 4 * (k * 32 + lane), dv[k]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(stat_full(sb));
        slot_nkt[sb] = it.n_kt;
        slot_use[0][sb] += 1;
        if (it.n_kt > 1) slot_use[1][sb] += 1;
      }
    }
    __syncwarp();
  } else if (warp == 17) {
    // ------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_128 = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_hd = make_idesc_bf16(128, HD, 0, 1);
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, HD, 1, 1);
      const uint64_t dQk0 = kmajor_base<HD>(sQDO), dQmn0 = mnmajor_base<HD>(sQDO);   // stage 0 of the Q / dO ring
      const uint64_t dDS0 = make_smem_desc(sDS, 16384, 1024, 2);
      int uk0 = 0, uk1 = 0, qi = 0, c0 = 0, c1 = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        Item it;
        if (!decode(item, it)) continue;
        uint64_t dKk[2] = {0, 0}, dVk[2] = {0, 0}, dKmn[2] = {0, 0};   // K / V tiles as K-major A operands, K as MN-major B
        int kbuf[2] = {0, 0};
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          if (kt >= it.n_kt) continue;
          const int uk = kt ? uk1 : uk0;
          kbuf[kt] = uk & 1;
          mbar_wait(kv_full(kbuf[kt], kt), uint32_t(uk >> 1) & 1);
          const uint32_t ka = sKV + (kbuf[kt] * 2 + kt) * 2 * TILE;
          dKk[kt] = kmajor_base<HD>(ka);
          dVk[kt] = kmajor_base<HD>(ka + TILE);
          dKmn[kt] = mnmajor_base<HD>(ka);
        }
        if (it.n_kt > 0) uk0++;
        if (it.n_kt > 1) uk1++;
        auto issue_S = [&](int kt, int st) {   // S^T = K Q^T (M = keys, N = queries); overwrites the dS^T columns
          const uint32_t tST = tmem_base + (kt ? B::TM_ST1 : B::TM_ST0);
          const uint64_t dq = desc_advance(dQk0, uint32_t(st) * 2 * TILE);
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_f16(tST, desc_advance(dKk[kt], kmajor_koff<HD>(kk)), desc_advance(dq, kmajor_koff<HD>(kk)), idesc_128, kk > 0);
          umma_commit(s_full(kt));
        };
        mbar_wait(qdo_full(qi & 1), uint32_t(qi >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
          if (kt < it.n_kt) issue_S(kt, qi & 1);
        for (int i = 0; i < it.n_q; ++i, ++qi) {
          const int st = qi & 1;
          const uint64_t dQk = desc_advance(dQk0, uint32_t(st) * 2 * TILE), dQmn = desc_advance(dQmn0, uint32_t(st) * 2 * TILE);
          const uint64_t dDOk = desc_advance(dQk, TILE), dDOmn = desc_advance(dQmn, TILE);
          const int qv = min(128, it.len - i * 128);
          const int qsteps = (qv + 15) >> 4;          // reduction steps over the valid queries of this tile
          const uint32_t tDQ = tmem_base + ((qi & 1) ? B::TM_DQ1 : B::TM_DQ0);
#pragma unroll
          for (int kt = 0; kt < 2; ++kt) {
            if (kt >= it.n_kt) continue;
            const int cc = kt ? c1 : c0;
            const uint32_t tST = tmem_base + (kt ? B::TM_ST1 : B::TM_ST0);
            const uint32_t tDV = tmem_base + (kt ? B::TM_DV1 : B::TM_DV0);
            const uint32_t tDK = tmem_base + (kt ? B::TM_DK1 : B::TM_DK0);
            const int kvalid = min(128, it.len - it.kv0 - kt * 128);
            const int ksteps = (kvalid + 15) >> 4;
            // ---- P^T written: dV += P^T dO_i (A = P^T from TMEM, in place over the S^T columns: queries 0..63 sit in
            // columns 0..31, 64..127 in columns 64..95), then dP^T = V dO_i^T over the same columns - tcgen05.mma
            // executes in issue order, so P^T has been consumed before it is overwritten
            mbar_wait(p_full(kt), uint32_t(cc) & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (kk < qsteps)
                umma_f16_ts(tDV, tST + (kk >> 2) * 64 + (kk & 3) * 8, desc_advance(dDOmn, mnmajor_koff<HD>(kk)), idesc_hd,
                            (i > 0 || kk > 0));
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk)
              umma_f16(tST, desc_advance(dVk[kt], kmajor_koff<HD>(kk)), desc_advance(dDOk, kmajor_koff<HD>(kk)), idesc_128, kk > 0);
            umma_commit(dp_full(kt));
            // ---- dS^T written (TMEM in place over dP^T, and shared memory): dK += dS^T Q_i ; S^T of the next query
            // tile (after dK in issue order: it overwrites dS^T) ; dQ_i (+)= dS K
            mbar_wait(ds_full(kt), uint32_t(cc) & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (kk < qsteps)
                umma_f16_ts(tDK, tST + (kk >> 2) * 64 + (kk & 3) * 8, desc_advance(dQmn, mnmajor_koff<HD>(kk)), idesc_hd,
                            (i > 0 || kk > 0));
            if (i + 1 < it.n_q) {
              if (kt == 0) { mbar_wait(qdo_full((qi + 1) & 1), uint32_t((qi + 1) >> 1) & 1); tc_fence_after(); }
              issue_S(kt, (qi + 1) & 1);
            }
            if (kt == 0 && qi >= 2) { mbar_wait(dq_free(qi & 1), uint32_t((qi >> 1) - 1) & 1); tc_fence_after(); }
            {
              const uint64_t dds = desc_advance(dDS0, uint32_t(kt) * C::P_BYTES);   // dS^T tile read M-major (queries contiguous)
              if (p.debug == 1) {   // bring-up: descriptors rebuilt per k-step exactly as attn_bwd.cu does
                const uint32_t ds_tile = sDS + kt * C::P_BYTES;
                const uint32_t ka = sKV + (kbuf[kt] * 2 + kt) * 2 * TILE;
                for (int kk = 0; kk < ksteps; ++kk)
                  umma_f16(tDQ, make_smem_desc(ds_tile + kk * 2048, 16384, 1024, 2), mnmajor_desc<HD>(ka, kk), idesc_dq,
                           (kt > 0 || kk > 0));
              } else {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)
                if (kk < ksteps)
                  umma_f16(tDQ, desc_advance(dds, uint32_t(kk) * 2048), desc_advance(dKmn[kt], mnmajor_koff<HD>(kk)), idesc_dq,
                           (kt > 0 || kk > 0));
              }
            }
            umma_commit(ps_free(kt));
            if (kt == it.n_kt - 1) { umma_commit(dq_done(qi & 1)); umma_commit(qdo_free(st)); }
            if (kt) ++c1; else ++c0;
          }
        }
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
          if (kt < it.n_kt) umma_commit(kv_free(kbuf[kt], kt));   // every MMA that reads this K / V buffer has retired
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------ softmax groups
    const int g = warp >> 3;                  // group = key tile
    const int wg = warp & 7;                  // warp inside the group
    const int qd = warp & 3;                  // TMEM lane quarter
    const int half = wg >> 2;                 // query-column half (needs wg's low bits to cover all four quarters)
    const int r = qd * 32 + lane;             // key row inside the tile
    const uint32_t lane_addr = uint32_t(qd * 32) << 16;
    const uint32_t stats = smem_u32(smem + B::STAT_OFF);
    const uint32_t ps = sDS + g * C::P_BYTES;
    const uint32_t dqs = smem_u32(smem + B::DQS_OFF) + (g * 4 + qd) * 4096;
    const uint32_t tST = tmem_base + (g ? B::TM_ST1 : B::TM_ST0) + lane_addr;
    const int col0 = half * 64;
    const int my_bar = 2 + g, other_bar = 3 - g;
    if (g == 1 && p.pingpong) bar_arrive_n(2, 512);   // group 0 runs the first exp pass
    int c = 0;                                // iterations done by this group (phases of s/p/dp/ds/ps_free)
    int qi = 0;                               // query tiles seen by the CTA (statistics slot, dQ buffer)
    auto drain_dq = [&](const Item& it, int qtile, int q_index) {   // dQ partial of a finished query tile -> global fp32
      const int b = q_index & 1;
      mbar_wait(dq_done(b), uint32_t(q_index >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + (b ? B::TM_DQ1 : B::TM_DQ0) + lane_addr, v);
      tmem_wait_ld();
      tc_fence_before();
      if (lane == 0) tma_wait_group_read<0>();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(dqs + lane * 128 + ((j ^ (lane & 7)) << 4), make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(dq_free(b));              // the TMEM partial is in registers / smem: the MMA warp may reuse the buffer
        tma_reduce_add_2d(&tmDQ, dqs, it.head * HD, it.row_begin + qtile * 128 + qd * 32);
        tma_commit_group();
      }
    };
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      Item it;
      if (!decode(item, it)) continue;
      if (g >= it.n_kt) { qi += it.n_q; continue; }
      const bool pingpong = it.n_kt == 2 && p.pingpong;
      const bool drainer = (g == it.n_kt - 1) && half == 0;   // the group whose dQ MMA completes the partial drains it
      const int kv_row0 = it.kv0 + g * 128;
      const uint32_t kvmask = (kv_row0 + r < it.len) ? 0xFFFFFFFFu : 0u;
      const bool kv_partial = kv_row0 + 128 > it.len;
      for (int i = 0; i < it.n_q; ++i, ++c, ++qi) {
        const int sb = qi & 1;
        const uint32_t lse_s = stats + sb * 1024 + 4 * col0;
        const uint32_t del_s = lse_s + 512;
        const int qv = min(128, it.len - i * 128);
        mbar_wait(stat_full(sb), uint32_t(qi >> 1) & 1);
        mbar_wait(s_full(g), uint32_t(c) & 1);
        tc_fence_after();
        if (pingpong) bar_sync_n(my_bar, 512);
        // ---- P^T = 2^(S^T scale - lse[q]) for this warp's 32 key rows x 64 query columns
        uint32_t pk[32];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (col0 + cc * 32 < qv) {
            uint32_t v[32];
            tmem_ld32(tST + col0 + cc * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 L = lds128f(lse_s + 4 * (cc * 32 + e));
              const float a0 = ex2_approx(fmaf(__uint_as_float(v[e]), p.scale_log2, -L.x));
              const float a1 = ex2_approx(fmaf(__uint_as_float(v[e + 1]), p.scale_log2, -L.y));
              const float a2 = ex2_approx(fmaf(__uint_as_float(v[e + 2]), p.scale_log2, -L.z));
              const float a3 = ex2_approx(fmaf(__uint_as_float(v[e + 3]), p.scale_log2, -L.w));
              pk[cc * 16 + e / 2] = pack_bf16x2(a0, a1);
              pk[cc * 16 + e / 2 + 1] = pack_bf16x2(a2, a3);
            }
            // the fused dQ sums over key rows: rows past the sequence end must carry P = dS = 0
            if (kv_partial) {
#pragma unroll
              for (int e = 0; e < 16; ++e) pk[cc * 16 + e] &= kvmask;
            }
          } else {   // query columns past the end of the sequence: probabilities are exactly 0, no exponentials
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[cc * 16 + e] = 0u;
          }
        }
        {   // P^T -> TMEM in place (both 32-column chunks of this warp have been read): A operand of the dV MMA
          uint32_t lo[16], hi[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) { lo[e] = pk[e]; hi[e] = pk[16 + e]; }
          tmem_st16(tST + col0, lo);
          tmem_st16(tST + col0 + 16, hi);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
        if (pingpong) bar_arrive_n(other_bar, 512);
        // the dQ partial of the previous query tile is complete (or about to be): drain it off the MMA warp's critical path
        if (drainer && i > 0) drain_dq(it, i - 1, qi - 1);
        mbar_wait(dp_full(g), uint32_t(c) & 1);
        tc_fence_after();
        // ---- dS^T = P^T o (dP^T - delta[q]) in packed bf16
        uint32_t dsa[32];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (col0 + cc * 32 < qv) {
            uint32_t v[32];
            tmem_ld32(tST + col0 + cc * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 Dl = lds128f(del_s + 4 * (cc * 32 + e));
              dsa[cc * 16 + e / 2] = mul_bf16x2(pk[cc * 16 + e / 2],
                                                pack_bf16x2(__uint_as_float(v[e]) - Dl.x, __uint_as_float(v[e + 1]) - Dl.y));
              dsa[cc * 16 + e / 2 + 1] = mul_bf16x2(pk[cc * 16 + e / 2 + 1],
                                                    pack_bf16x2(__uint_as_float(v[e + 2]) - Dl.z, __uint_as_float(v[e + 3]) - Dl.w));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) dsa[cc * 16 + e] = 0u;
          }
        }
        // shared-memory copy for the dQ MMA (M-major A operand); the tile is free once the previous dQ MMA retired
        if (i > 0) mbar_wait(ps_free(g), uint32_t(c - 1) & 1);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            ptile_store(ps, r, half * 8 + cc * 4 + q4,
                        make_uint4(dsa[cc * 16 + 4 * q4], dsa[cc * 16 + 4 * q4 + 1], dsa[cc * 16 + 4 * q4 + 2], dsa[cc * 16 + 4 * q4 + 3]));
        {   // TMEM copy in place over the dP^T columns this warp has read: A operand of the dK MMA
          uint32_t lo[16], hi[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) { lo[e] = dsa[e]; hi[e] = dsa[16 + e]; }
          tmem_st16(tST + col0, lo);
          tmem_st16(tST + col0 + 16, hi);
        }
        tmem_wait_st();
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(ds_full(g)); mbar_arrive(stat_free(g, sb)); }
      }
      // ---- item epilogue: every MMA of this key tile has retired once ps_free fired for the last query tile
      mbar_wait(ps_free(g), uint32_t(c - 1) & 1);
      tc_fence_after();
      if (drainer) drain_dq(it, it.n_q - 1, qi - 1);
      // half-0 warps store dV, half-1 warps dK (x scale) -> bf16 -> dqkv[:, v / k third]
      const int rows_valid = max(0, min(32, it.len - kv_row0 - qd * 32));
      const uint32_t stage = ps + wg * 2048;
      const uint32_t tsrc = tmem_base + lane_addr + (half == 0 ? (g ? B::TM_DV1 : B::TM_DV0) : (g ? B::TM_DK1 : B::TM_DK0));
      const float mul = half == 0 ? 1.0f : p.scale;
      __nv_bfloat16* gbase = p.dqkv + (half == 0 ? 2 : 1) * HHD + it.head * HD;
      {
        uint32_t v[32];
        tmem_ld32(tsrc, v);
        tmem_wait_ld();
        float acc[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] = __uint_as_float(v[e]);
        store_rows_bf16<32>(stage, acc, mul, lane, gbase, 3LL * HHD, it.row_begin + kv_row0 + qd * 32, rows_valid);
      }
      tc_fence_before();
    }
    if (half == 0 && lane == 0) tma_wait_group<0>();   // dQ reduce-adds of this warp have landed
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc<512>(tmem_base);
}

// small helper kernels shared with the first generation (defined in attn_bwd.cu)
int launch_attn_delta(const void* out, const void* dout, float* delta, int T, int H, int HD, cudaStream_t s);
int launch_attn_dq_convert(const float* dq_acc, void* dqkv, long long T, int HHD, float scale, cudaStream_t s);

template <int HD>
int launch_attn_bwd2(const void* qkv, const void* out, const void* dout, const float* lse2, float* delta, void* dqkv,
                     float* dq_acc, const int* cu, int nseq, int max_len, int H, int T, float scale, cudaStream_t s) {
  using C = AttnCfg<HD>;
  using B = Bwd2Cfg<HD>;
  CUtensorMap tq, tdo, tdq;
  int rc = make_tmap_2d(&tq, qkv, 0, (uint64_t)3 * H * HD, T, (uint64_t)3 * H * HD * 2, C::BOX_INNER, 128, C::TMAP_SWIZZLE);
  if (rc) return rc;
  rc = make_tmap_2d(&tdo, dout, 0, (uint64_t)H * HD, T, (uint64_t)H * HD * 2, C::BOX_INNER, 128, C::TMAP_SWIZZLE);
  if (rc) return rc;
  rc = make_tmap_2d(&tdq, dq_acc, 1, (uint64_t)H * HD, T, (uint64_t)H * HD * 4, 32, 32, 3);
  if (rc) return rc;
  auto kern = attn_bwd2_kernel<HD>;
  static bool configured = false;
  if (!configured) {
    VJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, B::SMEM_BYTES));
    configured = true;
  }
  VJ_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)T * H * HD * sizeof(float), s));
  rc = launch_attn_delta(out, dout, delta, T, H, HD, s);
  if (rc) return rc;
  AttnBwd2Params p;
  p.cu_seqlens = cu; p.lse2 = lse2; p.delta = delta; p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.H = H; p.T = T; p.nseq = nseq; p.kpairs = (max_len + 255) / 256;
  p.n_items = p.kpairs * nseq * H;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  static int persist = -1, pingpong = -1;
  if (persist < 0) { const char* e = getenv("VJ_ATTN_PERSIST"); persist = (e && e[0] == '0') ? 0 : 1; }
  if (pingpong < 0) { const char* e = getenv("VJ_ATTN_PINGPONG"); pingpong = (e && e[0] == '0') ? 0 : 1; }
  p.pingpong = pingpong;
  { const char* e = getenv("VJ_BWD2_DEBUG"); p.debug = e ? atoi(e) : 0; }
  const int grid = (persist && p.n_items > sm_budget()) ? sm_budget() : p.n_items;
  kern<<<grid, kBwd2Threads, B::SMEM_BYTES, s>>>(tq, tdo, tdq, p);
  VJ_CUDA(cudaGetLastError());
  vj::count_launch(1);
  return launch_attn_dq_convert(dq_acc, dqkv, T, H * HD, scale, s);
}

template int launch_attn_bwd2<32>(const void*, const void*, const void*, const float*, float*, void*, float*, const int*, int,
                                  int, int, int, float, cudaStream_t);

}  // namespace vj
