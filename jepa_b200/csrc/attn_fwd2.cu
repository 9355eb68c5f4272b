// Dense var-len flash attention forward, second generation: one PERSISTENT CTA per SM working on TWO 128-row query
// tiles at a time with two softmax warpgroups in an explicit ping-pong (sm_100a, tcgen05 / TMEM / TMA).
//
// Replaces F.scaled_dot_product_attention at src/models/utils/modules.py:66-69 (same contract as attn_fwd.cu: qkv bf16
// [T, 3*H*HD] q|k|v thirds head-major, out bf16 [T, H*HD], lse2 fp32 [H, T] in the log2 domain, sequences = row ranges
// of cu_seqlens).
//
// Why: softmax at head dim 64 / 32 needs 16384 exponentials per 128x128 score tile = 1024 MUFU cycles per SM, twice the
// tensor time.  With one softmax group per CTA and two independent CTAs per SM (attn_fwd.cu) the two groups drift
// into the same phase: both fight for the MUFU pipe during their exp pass and both leave it idle while they load scores
// / wait for MMAs (measured: MUFU 59 % busy, tensor 30 %).  Here the two groups live in ONE CTA and hand a token back
// and forth (named barriers): exactly one group is in its exp pass at any time, the other one meanwhile pulls its next
// score tile out of TMEM, takes the row max and (rarely) rescales - the MUFU pipe never idles, K / V tiles are loaded
// once for both query tiles, and the MMA warp always has the next QK^T in flight.
//
//   warps 0-3   softmax group 0 (query tile 0), one thread per query row, TMEM lane quarter = warp & 3
//   warps 4-7   softmax group 1 (query tile 1)
//   warp 8      TMA producer : Q tiles (double-buffered across work items when HD <= 64), K ring, V ring
//   warp 9      MMA issuer   : S_t = Q_t K_j^T (tile t in {0,1}), O_t += P_t V_j with P_t read from TENSOR MEMORY.  One
//               thread issues every tcgen05.mma of the CTA: it has the highest warp id (the schedulers pick the highest
//               eligible warp first) and adds compile-time offsets to per-tile descriptors instead of rebuilding them
// TMEM (512 columns): S0 | S1 | O0 | O1 | P0 | P1 for HD <= 64; for HD = 128 the bf16 P tile overwrites the first 64
// columns of its own S tile (every thread has its whole score row in registers before it writes P).
// The last KV tile of a sequence only runs ceil(valid/16) reduction steps of the PV MMA and ceil(valid/16)*16 score
// columns; fully masked 32-column chunks are never exponentiated.
// Work items (sequence, head, pair of query tiles) are processed persistently: the next item's Q / K / V loads and
// first QK^T overlap the previous item's epilogue.
#include <stdlib.h>

#include "attn_common.cuh"
#include "vjepa_b200.h"

namespace vj {

constexpr int kFwd2Threads = 320;
constexpr int kDefaultPoly = 0;   // set from measurements (profiles/): exponentials out of 8 moved to the FMA pipe

struct AttnFwd2Params {
  const int* cu_seqlens;
  __nv_bfloat16* out;
  float* lse2;
  int H, T, nseq, qpairs, n_items;
  int pingpong;     // 1: the two softmax groups take turns in the exp pass (named-barrier token); 0: free running
  long long ld_out;
  float scale_log2;
};

template <int HD>
struct Fwd2Cfg {
  using A = AttnCfg<HD>;
  static constexpr int TILE = A::TILE_BYTES;
  static constexpr bool ALIAS_P = HD > 64;
  static constexpr int QBUF = HD <= 64 ? 2 : 1;
  static constexpr int KST = HD <= 64 ? 3 : 2;
  static constexpr int VST = HD <= 64 ? 3 : 2;
  static constexpr int Q_OFF = 0;                                   // [QBUF][2] tiles
  static constexpr int K_OFF = Q_OFF + QBUF * 2 * TILE;
  static constexpr int V_OFF = K_OFF + KST * TILE;
  static constexpr int STG_OFF = QBUF == 2 ? V_OFF + VST * TILE : Q_OFF;   // O staging (aliases Q when Q is single-buffered)
  static constexpr int BAR_OFF = V_OFF + VST * TILE + (QBUF == 2 ? 2 * TILE : 0);
  static constexpr int NBARS = 4 * QBUF + 2 * KST + 2 * VST + 8;
  static constexpr int SMEM_USED = BAR_OFF + NBARS * 8 + 16 + 1024;
  static constexpr int SMEM_BYTES = SMEM_USED < 120 * 1024 ? 120 * 1024 : SMEM_USED;   // one CTA per SM (512 TMEM columns)
  static_assert(SMEM_BYTES <= 232448, "attn_fwd2 shared memory budget exceeded");
  static constexpr int OW = HD <= 64 ? 64 : 128;                    // TMEM columns reserved per O tile
  static constexpr int TM_S0 = 0, TM_S1 = 128, TM_O0 = 256, TM_O1 = 256 + OW;
  static constexpr int TM_P0 = ALIAS_P ? 0 : 384, TM_P1 = ALIAS_P ? 128 : 448;
};

VJ_DEVINL void bar_sync_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
VJ_DEVINL void bar_arrive_named(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// 32 scores -> 32 probabilities p = 2^(s*scale - moff), row-sum accumulated, packed to 16 bf16 pairs.
// MASK: columns >= valid (relative to this chunk's first column `base`) produce exactly 0.
// POLY of every 8 exponentials run on the FMA pipe (ex2_poly) instead of MUFU.
template <bool MASK, int POLY>
VJ_DEVINL void exp_chunk(const uint32_t (&s)[32], int base, int valid, float scale_log2, float moff, float& l0, float& l1,
                         uint32_t (&pk)[16]) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float xa = fmaf(__uint_as_float(s[i]), scale_log2, -moff);
    const float xb = fmaf(__uint_as_float(s[i + 1]), scale_log2, -moff);
    float a = ((i & 7) < POLY) ? ex2_poly(xa) : ex2_approx(xa);
    float b = (((i + 1) & 7) < POLY) ? ex2_poly(xb) : ex2_approx(xb);
    if (MASK) {
      a = (base + i < valid) ? a : 0.f;
      b = (base + i + 1 < valid) ? b : 0.f;
    }
    l0 += a;
    l1 += b;
    pk[i >> 1] = pack_bf16x2(a, b);
  }
}

template <int HD, int POLY>
__global__ void __launch_bounds__(kFwd2Threads, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwd2Params p) {
  using C = AttnCfg<HD>;
  using F = Fwd2Cfg<HD>;
  constexpr int QBUF = F::QBUF, KST = F::KST, VST = F::VST, TILE = F::TILE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F::BAR_OFF);
  const uint32_t b0 = smem_u32(bars);
  // barrier map (8 bytes each)
  auto q_full = [&](int buf, int t) { return b0 + 8 * (buf * 2 + t); };
  auto q_free = [&](int buf, int t) { return b0 + 8 * (2 * QBUF + buf * 2 + t); };
  const uint32_t k_full0 = b0 + 8 * (4 * QBUF), k_free0 = k_full0 + 8 * KST;
  const uint32_t v_full0 = k_free0 + 8 * KST, v_free0 = v_full0 + 8 * VST;
  const uint32_t s_full0 = v_free0 + 8 * VST;     // [2]
  const uint32_t s_free0 = s_full0 + 16;          // [2]
  const uint32_t p_full0 = s_free0 + 16;          // [2]
  const uint32_t o_done0 = p_full0 + 16;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + F::NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * QBUF; ++i) { mbar_init(q_full(0, 0) + 8 * i, 1); mbar_init(q_free(0, 0) + 8 * i, 4); }
    for (int i = 0; i < KST; ++i) { mbar_init(k_full0 + 8 * i, 1); mbar_init(k_free0 + 8 * i, 1); }
    for (int i = 0; i < VST; ++i) { mbar_init(v_full0 + 8 * i, 1); mbar_init(v_free0 + 8 * i, 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(s_full0 + 8 * t, 1); mbar_init(s_free0 + 8 * t, 4);
      mbar_init(p_full0 + 8 * t, 4); mbar_init(o_done0 + 8 * t, 1);
    }
    fence_mbar_init();
  }
  if (warp == 8 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 9) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sQ = smem_u32(smem + F::Q_OFF), sK = smem_u32(smem + F::K_OFF), sV = smem_u32(smem + F::V_OFF);
  const int HHD = p.H * HD;

  // work item decode, identical in every role
  struct Item { int row_begin, len, q0, n_qt, n_kv, head; };
  auto decode = [&](int item, Item& it) -> bool {
    const int qp = item % p.qpairs;
    const int rest = item / p.qpairs;
    const int seq = rest % p.nseq;
    it.head = rest / p.nseq;
    it.row_begin = p.cu_seqlens[seq];
    it.len = p.cu_seqlens[seq + 1] - it.row_begin;
    it.q0 = qp * 256;
    if (it.q0 >= it.len) return false;
    it.n_qt = (it.q0 + 128 < it.len) ? 2 : 1;
    it.n_kv = (it.len + 127) >> 7;
    return true;
  };

  if (warp == 8) {
    // ------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int kv_it = 0;            // KV tiles loaded so far (ring position)
      int uq[2] = {0, 0};       // Q loads so far per tile
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        Item it;
        if (!decode(item, it)) continue;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t >= it.n_qt) continue;
          const int buf = uq[t] % QBUF;
          const uint32_t ph = uint32_t(uq[t] / QBUF) & 1;
          mbar_wait(q_free(buf, t), ph ^ 1);
          mbar_expect_tx(q_full(buf, t), TILE);
#pragma unroll
          for (int b = 0; b < C::NBOX; ++b)
            tma_load_2d(sQ + (buf * 2 + t) * TILE + b * C::BOX_BYTES, &tmQKV, q_full(buf, t),
                        it.head * HD + b * C::BOX_INNER, it.row_begin + it.q0 + t * 128);
          ++uq[t];
        }
        for (int j = 0; j < it.n_kv; ++j, ++kv_it) {
          const int kr = it.row_begin + j * 128;
          const int ks = kv_it % KST, vs = kv_it % VST;
          mbar_wait(k_free0 + 8 * ks, (uint32_t(kv_it / KST) & 1) ^ 1);
          mbar_expect_tx(k_full0 + 8 * ks, TILE);
#pragma unroll
          for (int b = 0; b < C::NBOX; ++b)
            tma_load_2d(sK + ks * TILE + b * C::BOX_BYTES, &tmQKV, k_full0 + 8 * ks, HHD + it.head * HD + b * C::BOX_INNER, kr);
          mbar_wait(v_free0 + 8 * vs, (uint32_t(kv_it / VST) & 1) ^ 1);
          mbar_expect_tx(v_full0 + 8 * vs, TILE);
#pragma unroll
          for (int b = 0; b < C::NBOX; ++b)
            tma_load_2d(sV + vs * TILE + b * C::BOX_BYTES, &tmQKV, v_full0 + 8 * vs,
                        2 * HHD + it.head * HD + b * C::BOX_INNER, kr);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);
      constexpr uint32_t idesc_full = make_idesc_bf16(128, 128, 0, 0);
      const uint64_t dK0 = kmajor_base<HD>(sK), dV0 = mnmajor_base<HD>(sV);
      const uint32_t tmS[2] = {tmem_base + F::TM_S0, tmem_base + F::TM_S1};
      const uint32_t tmO[2] = {tmem_base + F::TM_O0, tmem_base + F::TM_O1};
      const uint32_t tmP[2] = {tmem_base + F::TM_P0, tmem_base + F::TM_P1};
      int kv_it = 0;
      int uq[2] = {0, 0};
      int c[2] = {0, 0};        // KV iterations completed per query tile (barrier phases of s/p/o)
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        Item it;
        if (!decode(item, it)) continue;
        uint64_t dQ[2] = {0, 0};
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t >= it.n_qt) continue;
          const int buf = uq[t] % QBUF;
          mbar_wait(q_full(buf, t), uint32_t(uq[t] / QBUF) & 1);
          dQ[t] = kmajor_base<HD>(sQ + (buf * 2 + t) * TILE);
          ++uq[t];
        }
        const int last_valid = it.len - (it.n_kv - 1) * 128;           // keys in the last KV tile (1..128)
        // the tail tile only produces ceil16(valid) score columns and runs ceil16(valid)/16 reduction steps of P V
        const uint32_t idesc_tail = make_idesc_bf16(128, (last_valid + 15) & ~15, 0, 0);
        auto issue_qk = [&](int t, int j, int kvi) {   // S_t = Q_t K_j^T
          const uint64_t dk = desc_advance(dK0, uint32_t(kvi % KST) * TILE);
          const uint32_t idesc_s = (j == it.n_kv - 1) ? idesc_tail : idesc_full;
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_f16(tmS[t], desc_advance(dQ[t], kmajor_koff<HD>(kk)), desc_advance(dk, kmajor_koff<HD>(kk)), idesc_s, kk > 0);
          umma_commit(s_full0 + 8 * t);
        };
        // first score tiles of the item: the S buffers were handed back by the softmax groups in their last iteration
        mbar_wait(k_full0 + 8 * (kv_it % KST), uint32_t(kv_it / KST) & 1);
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t >= it.n_qt) continue;
          if (!F::ALIAS_P && c[t] > 0) { mbar_wait(s_free0 + 8 * t, uint32_t(c[t] - 1) & 1); tc_fence_after(); }
          issue_qk(t, 0, kv_it);
        }
        umma_commit(k_free0 + 8 * (kv_it % KST));
        for (int j = 0; j < it.n_kv; ++j, ++kv_it) {
          const bool more = j + 1 < it.n_kv;
          const int vs = kv_it % VST;
          const int ksteps = (j == it.n_kv - 1) ? (last_valid + 15) >> 4 : 8;   // reduction steps of the PV MMA
          const uint64_t dv = desc_advance(dV0, uint32_t(vs) * TILE);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (t >= it.n_qt) continue;
            auto issue_pv = [&]() {
              mbar_wait(p_full0 + 8 * t, uint32_t(c[t]) & 1);
              if (t == 0) mbar_wait(v_full0 + 8 * vs, uint32_t(kv_it / VST) & 1);
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)
                if (kk < ksteps)
                  umma_f16_ts(tmO[t], tmP[t] + kk * 8, desc_advance(dv, mnmajor_koff<HD>(kk)), idesc_o, (j > 0 || kk > 0));
              umma_commit(o_done0 + 8 * t);
              if (t == it.n_qt - 1) umma_commit(v_free0 + 8 * vs);
            };
            auto issue_next_qk = [&]() {
              if (!more) return;
              if (!F::ALIAS_P) mbar_wait(s_free0 + 8 * t, uint32_t(c[t]) & 1);   // S_t(j) is in registers
              if (t == 0) mbar_wait(k_full0 + 8 * ((kv_it + 1) % KST), uint32_t((kv_it + 1) / KST) & 1);
              tc_fence_after();
              issue_qk(t, j + 1, kv_it + 1);
              if (t == it.n_qt - 1) umma_commit(k_free0 + 8 * ((kv_it + 1) % KST));
            };
            if (F::ALIAS_P) { issue_pv(); issue_next_qk(); }   // P lives in S's columns: PV must be issued first
            else            { issue_next_qk(); issue_pv(); }
            ++c[t];
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------ softmax groups
    const int t = warp >> 2;                 // query tile / group
    const int qd = warp & 3;                 // TMEM lane quarter
    const int r = qd * 32 + lane;            // row inside the tile
    const uint32_t lane_addr = uint32_t(qd * 32) << 16;
    const uint32_t tmS = tmem_base + (t ? F::TM_S1 : F::TM_S0) + lane_addr;
    const uint32_t tmO = tmem_base + (t ? F::TM_O1 : F::TM_O0) + lane_addr;
    const uint32_t tmP = tmem_base + (t ? F::TM_P1 : F::TM_P0) + lane_addr;
    const uint32_t s_full = s_full0 + 8 * t, s_free = s_free0 + 8 * t, p_full = p_full0 + 8 * t, o_done = o_done0 + 8 * t;
    const int my_bar = 2 + t, other_bar = 3 - t;   // named barriers 2 / 3: "group t may run its exp pass"
    if (t == 1 && p.pingpong) bar_arrive_named(2, 256);   // group 0 goes first
    int c = 0;        // KV iterations done by this group
    int uq = 0;       // items done by this group
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      Item it;
      if (!decode(item, it)) continue;
      if (t >= it.n_qt) continue;
      const bool pingpong = it.n_qt == 2 && p.pingpong;
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < it.n_kv; ++j, ++c) {
        const int valid = min(128, it.len - j * 128);
        const int nch = (valid + 31) >> 5;
        mbar_wait(s_full, uint32_t(c) & 1);
        tc_fence_after();
        // ---- score row -> registers (only the chunks that hold valid keys), row max.  Columns 96..127 are only scanned
        // for the max here and re-read from TMEM in the middle of the exp pass (prefetched behind chunks 1 and 2): with ten
        // warps per CTA an SM sub-partition hosts three of them, i.e. 168 registers per thread, and 128 live scores plus
        // the packed probabilities do not fit.
        uint32_t s0[32], s1[32], s2[32];
        float mx = -INFINITY;
        {
          uint32_t t3[32];
          tmem_ld32(tmS + 0, s0);
          if (nch > 1) tmem_ld32(tmS + 32, s1);
          if (nch > 2) tmem_ld32(tmS + 64, s2);
          if (nch > 3) tmem_ld32(tmS + 96, t3);
          tmem_wait_ld();
          if (valid == 128) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              mx = fmaxf(fmaxf(mx, __uint_as_float(s0[i])),
                         fmaxf(__uint_as_float(s1[i]), fmaxf(__uint_as_float(s2[i]), __uint_as_float(t3[i]))));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i < valid) mx = fmaxf(mx, __uint_as_float(s0[i]));
              if (32 + i < valid) mx = fmaxf(mx, __uint_as_float(s1[i]));
              if (64 + i < valid) mx = fmaxf(mx, __uint_as_float(s2[i]));
              if (96 + i < valid) mx = fmaxf(mx, __uint_as_float(t3[i]));
            }
          }
        }
        // ---- lazy rescale: move the reference max only when it would overflow the 2^8 head-room
        const bool grow = (mx - m_ref) * p.scale_log2 > 8.0f;   // true on the first tile (m_ref = -inf)
        if (__any_sync(0xffffffffu, grow)) {
          if (j > 0) {
            mbar_wait(o_done, uint32_t(c - 1) & 1);   // PV(j-1) retired: O_t is stable and not being accumulated into
            tc_fence_after();
            const float alpha = grow ? ex2_approx((m_ref - mx) * p.scale_log2) : 1.0f;
#pragma unroll
            for (int cc = 0; cc < HD / 16; ++cc) {
              uint32_t o[16];
              tmem_ld16(tmO + cc * 16, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(tmO + cc * 16, o);
            }
            tmem_wait_st();
            l *= alpha;
          }
          if (grow) m_ref = mx;
        }
        const float moff = m_ref * p.scale_log2;
        // ---- exp pass: only one group at a time (MUFU is the binding pipe)
        if (pingpong) bar_sync_named(my_bar, 256);
        if (j > 0) { mbar_wait(o_done, uint32_t(c - 1) & 1); tc_fence_after(); }   // PV(j-1) finished reading P_t
        float l0 = 0.f, l1 = 0.f;
        uint32_t pk[16];
        auto release_s = [&]() {   // every score of S_t(j) has left TMEM: the next QK^T may overwrite it
          if (!F::ALIAS_P) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);
          }
        };
        if (valid == 128) {
          exp_chunk<false, POLY>(s0, 0, valid, p.scale_log2, moff, l0, l1, pk);  tmem_st16(tmP + 0, pk);
          tmem_ld32(tmS + 96, s0);     // chunk 3 again, into the registers chunk 0 just left; lands behind chunks 1, 2
          exp_chunk<false, POLY>(s1, 32, valid, p.scale_log2, moff, l0, l1, pk); tmem_st16(tmP + 16, pk);
          exp_chunk<false, POLY>(s2, 64, valid, p.scale_log2, moff, l0, l1, pk); tmem_st16(tmP + 32, pk);
          tmem_wait_ld();
          release_s();
          exp_chunk<false, POLY>(s0, 96, valid, p.scale_log2, moff, l0, l1, pk); tmem_st16(tmP + 48, pk);
        } else {
          exp_chunk<true, POLY>(s0, 0, valid, p.scale_log2, moff, l0, l1, pk);  tmem_st16(tmP + 0, pk);
          if (nch > 3) tmem_ld32(tmS + 96, s0);
          if (nch > 1) { exp_chunk<true, POLY>(s1, 32, valid, p.scale_log2, moff, l0, l1, pk); tmem_st16(tmP + 16, pk); }
          if (nch > 2) { exp_chunk<true, POLY>(s2, 64, valid, p.scale_log2, moff, l0, l1, pk); tmem_st16(tmP + 32, pk); }
          tmem_wait_ld();
          release_s();
          if (nch > 3) { exp_chunk<true, POLY>(s0, 96, valid, p.scale_log2, moff, l0, l1, pk); tmem_st16(tmP + 48, pk); }
        }
        l += l0 + l1;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // hand the token over (the very last hand-over of the kernel is harmlessly left pending)
        if (pingpong) bar_arrive_named(other_bar, 256);
      }
      // ---- epilogue: O / l -> bf16 -> swizzled smem staging -> coalesced global stores; LSE
      mbar_wait(o_done, uint32_t(c - 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / l;
      const bool row_ok = it.q0 + t * 128 + r < it.len;
      if (row_ok) p.lse2[(long long)it.head * p.T + it.row_begin + it.q0 + t * 128 + r] = m_ref * p.scale_log2 + log2f(l);
      constexpr int ORB = HD * 2;               // bytes per output row
      constexpr int CH = ORB / 16;              // 16-byte chunks per row
      const int buf = uq % QBUF;
      const uint32_t stage = smem_u32(smem + F::STG_OFF) + (QBUF == 2 ? t * TILE : (buf * 2 + t) * TILE) + qd * (32 * ORB);
#pragma unroll
      for (int cc = 0; cc < HD / 16; ++cc) {
        uint32_t o[16];
        tmem_ld16(tmO + cc * 16, o);
        tmem_wait_ld();
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * h2 + 0]) * inv, __uint_as_float(o[8 * h2 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[8 * h2 + 2]) * inv, __uint_as_float(o[8 * h2 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[8 * h2 + 4]) * inv, __uint_as_float(o[8 * h2 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[8 * h2 + 6]) * inv, __uint_as_float(o[8 * h2 + 7]) * inv);
          const int g = 2 * cc + h2;
          sts128(stage + lane * ORB + ((g ^ (lane & (CH - 1))) << 4), u);
        }
      }
      tc_fence_before();
      __syncwarp();
      constexpr int ROWS_PER_IT = 32 / CH;
      const int tile_row0 = it.q0 + t * 128 + qd * 32;
#pragma unroll
      for (int k2 = 0; k2 < CH; ++k2) {
        const int rr = k2 * ROWS_PER_IT + lane / CH;
        const int g = lane % CH;
        if (tile_row0 + rr < it.len) {
          const uint4 u = lds128(stage + rr * ORB + ((g ^ (rr & (CH - 1))) << 4));
          *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) +
                                    ((long long)(it.row_begin + tile_row0 + rr) * p.ld_out + it.head * HD) * 2 + g * 16) = u;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(q_free(buf, t));   // Q buffer (and, if aliased, the staging area) may be reloaded
      ++uq;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

template <int HD>
int launch_attn_fwd2(const void* qkv, void* out, float* lse2, const int* cu, int nseq, int max_len, int H, int T,
                     float scale, cudaStream_t s) {
  using C = AttnCfg<HD>;
  using F = Fwd2Cfg<HD>;
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, qkv, 0, (uint64_t)3 * H * HD, T, (uint64_t)3 * H * HD * 2, C::BOX_INNER, 128, C::TMAP_SWIZZLE);
  if (rc) return rc;
  // VJ_ATTN_POLY = exponentials out of every 8 that run on the FMA pipe (0, 2 or 3; head dims <= 64 only)
  static int poly = -1;
  if (poly < 0) { const char* e = getenv("VJ_ATTN_POLY"); poly = e ? atoi(e) : kDefaultPoly; }
  void (*kern)(const CUtensorMap, const AttnFwd2Params) = attn_fwd2_kernel<HD, 0>;
  if (HD <= 64 && poly == 2) kern = attn_fwd2_kernel<(HD <= 64 ? HD : 32), 2>;
  if (HD <= 64 && poly == 3) kern = attn_fwd2_kernel<(HD <= 64 ? HD : 32), 3>;
  static void* configured[4] = {nullptr, nullptr, nullptr, nullptr};
  if (configured[poly & 3] != reinterpret_cast<void*>(kern)) {
    VJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM_BYTES));
    configured[poly & 3] = reinterpret_cast<void*>(kern);
  }
  AttnFwd2Params p;
  p.cu_seqlens = cu; p.out = reinterpret_cast<__nv_bfloat16*>(out); p.lse2 = lse2;
  p.H = H; p.T = T; p.nseq = nseq; p.qpairs = (max_len + 255) / 256;
  p.n_items = p.qpairs * nseq * H;
  p.ld_out = (long long)H * HD;
  p.scale_log2 = scale * 1.4426950408889634f;
  static int persist = -1, pingpong = -1;
  if (persist < 0) { const char* e = getenv("VJ_ATTN_PERSIST"); persist = (e && e[0] == '0') ? 0 : 1; }
  if (pingpong < 0) { const char* e = getenv("VJ_ATTN_PINGPONG"); pingpong = (e && e[0] == '0') ? 0 : 1; }
  p.pingpong = pingpong;
  const int grid = (persist && p.n_items > sm_budget()) ? sm_budget() : p.n_items;
  kern<<<grid, kFwd2Threads, F::SMEM_BYTES, s>>>(tm, p);
  VJ_CUDA(cudaGetLastError());
  vj::count_launch(1);
  return 0;
}

template int launch_attn_fwd2<32>(const void*, void*, float*, const int*, int, int, int, int, float, cudaStream_t);
template int launch_attn_fwd2<64>(const void*, void*, float*, const int*, int, int, int, int, float, cudaStream_t);
template int launch_attn_fwd2<128>(const void*, void*, float*, const int*, int, int, int, int, float, cudaStream_t);

}  // namespace vj
