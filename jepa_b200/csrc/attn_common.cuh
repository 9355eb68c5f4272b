// Shared tile geometry and smem-descriptor helpers of the attention kernels (fwd and bwd).
#pragma once
#include "common.cuh"

namespace vj {

constexpr int kAttnThreads = 192;

template <int HD>
struct AttnCfg {
  static constexpr int BOX_INNER = HD >= 64 ? 64 : HD;          // elements per TMA box row
  static constexpr int NBOX = HD / BOX_INNER;                    // boxes per [128 x HD] tile
  static constexpr int ROW_BYTES = BOX_INNER * 2;                // 128 (SW128) or 64 (SW64)
  static constexpr int BOX_BYTES = 128 * ROW_BYTES;
  static constexpr int TILE_BYTES = NBOX * BOX_BYTES;            // Q / K / V tile
  static constexpr int LAYOUT = HD >= 64 ? 2 : 4;                // smem descriptor swizzle type
  static constexpr int TMAP_SWIZZLE = HD >= 64 ? 3 : 2;
  static constexpr int SBO = 8 * ROW_BYTES;                      // 8-row group stride
  static constexpr int MN_KSTEP = 16 * ROW_BYTES;                // 16 reduction rows of an MN-major tile
  static constexpr int P_BYTES = 128 * 128 * 2;
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = TILE_BYTES;
  static constexpr int V_OFF = 2 * TILE_BYTES;
  static constexpr int P_OFF = 3 * TILE_BYTES;
  static constexpr int BAR_OFF = P_OFF + P_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 128 + 1024;
  static constexpr int TMEM_COLS = (128 + HD) <= 256 ? 256 : 512;
};


// K-major operand descriptor for k-step kk (16 elements) of a [128 x HD] tile
template <int HD>
VJ_DEVINL uint64_t kmajor_desc(uint32_t tile, int kk) {
  using C = AttnCfg<HD>;
  constexpr int steps_per_box = C::BOX_INNER / 16;
  const uint32_t addr = tile + (kk / steps_per_box) * C::BOX_BYTES + (kk % steps_per_box) * 32;
  return make_smem_desc(addr, 16, C::SBO, C::LAYOUT);
}
// MN-major operand descriptor (N = HD contiguous, reduction = rows) for k-step kk (16 rows)
template <int HD>
VJ_DEVINL uint64_t mnmajor_desc(uint32_t tile, int kk) {
  using C = AttnCfg<HD>;
  return make_smem_desc(tile + kk * C::MN_KSTEP, C::BOX_BYTES, C::SBO, C::LAYOUT);
}
// Byte offsets of k-step kk inside a tile (compile-time when kk is an unrolled loop index): the MMA-issuing thread builds
// one descriptor per operand tile and only ADDS these (>> 4) to its address field - a tcgen05.mma then costs a couple of
// integer instructions to issue instead of a descriptor construction (a single thread issues every MMA of the CTA).
template <int HD>
VJ_DEVINL constexpr uint32_t kmajor_koff(int kk) {
  using C = AttnCfg<HD>;
  return uint32_t((kk / (C::BOX_INNER / 16)) * C::BOX_BYTES + (kk % (C::BOX_INNER / 16)) * 32);
}
template <int HD>
VJ_DEVINL constexpr uint32_t mnmajor_koff(int kk) { return uint32_t(kk * AttnCfg<HD>::MN_KSTEP); }
template <int HD>
VJ_DEVINL uint64_t kmajor_base(uint32_t tile) { return make_smem_desc(tile, 16, AttnCfg<HD>::SBO, AttnCfg<HD>::LAYOUT); }
template <int HD>
VJ_DEVINL uint64_t mnmajor_base(uint32_t tile) {
  return make_smem_desc(tile, AttnCfg<HD>::BOX_BYTES, AttnCfg<HD>::SBO, AttnCfg<HD>::LAYOUT);
}
VJ_DEVINL uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + uint64_t(bytes >> 4); }

// P / dS tile: [128 rows x 128 reduction] bf16, K-major, two 128B-swizzled atoms of 64 columns
VJ_DEVINL uint64_t ptile_desc(uint32_t tile, int kk) {
  return make_smem_desc(tile + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024, 2);
}


// write 8 packed bf16 (16 bytes) of row r, 16-byte column chunk col8 (0..15) into a [128x128] K-major tile
VJ_DEVINL void ptile_store(uint32_t tile, int r, int col8, const uint4& u) {
  sts128(tile + (col8 >> 3) * 16384 + r * 128 + (((col8 & 7) ^ (r & 7)) << 4), u);
}

// coalesced store of a per-warp staged [32 rows x HD] bf16 block to global rows
template <int HD>
VJ_DEVINL void store_rows_bf16(uint32_t stage, const float (&vals)[HD], float mul, int lane, __nv_bfloat16* gbase,
                               long long ld, int row_first, int rows_valid) {
  constexpr int ORB = HD * 2, CH = ORB / 16, ROWS_PER_IT = 32 / CH;
#pragma unroll
  for (int g = 0; g < CH; ++g) {
    uint4 u;
    u.x = pack_bf16x2(vals[8 * g + 0] * mul, vals[8 * g + 1] * mul);
    u.y = pack_bf16x2(vals[8 * g + 2] * mul, vals[8 * g + 3] * mul);
    u.z = pack_bf16x2(vals[8 * g + 4] * mul, vals[8 * g + 5] * mul);
    u.w = pack_bf16x2(vals[8 * g + 6] * mul, vals[8 * g + 7] * mul);
    sts128(stage + lane * ORB + ((g ^ (lane & (CH - 1))) << 4), u);
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < CH; ++it) {
    const int rr = it * ROWS_PER_IT + lane / CH;
    const int g = lane % CH;
    if (rr < rows_valid) {
      const uint4 u = lds128(stage + rr * ORB + ((g ^ (rr & (CH - 1))) << 4));
      *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(gbase) + ((long long)(row_first + rr) * ld) * 2 + g * 16) = u;
    }
  }
  __syncwarp();
}


// 32 scores -> p = 2^(s*scale - moff), packed fp32x2 scale/subtract and row sums, bf16 pairs stored to TMEM four columns
// (eight probabilities) at a time.  MASK: columns >= valid (absolute index base + i) produce exactly 0.
// POLY of every 8 exponentials run on the FMA / ALU pipes (ex2_poly) instead of MUFU, the pipe that bounds softmax.
// Only the 8-column groups G0 <= g < G1 are processed (lets the caller interleave a tcgen05.ld between two halves).
template <bool MASK, int POLY = 0, int G0 = 0, int G1 = 4>
VJ_DEVINL void exp_store32(const uint32_t (&sv)[32], int base, int valid, uint64_t scale2, uint64_t nmoff2, uint64_t& lsum,
                           uint32_t tmem_p) {
#pragma unroll
  for (int g = G0; g < G1; ++g) {
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 8 * g + 2 * k;
      float xa, xb;
      upk2(fma2(pk2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), scale2, nmoff2), xa, xb);
      float a = ((i & 7) >= 8 - POLY) ? ex2_poly(xa) : ex2_approx(xa);
      float b = (((i + 1) & 7) >= 8 - POLY) ? ex2_poly(xb) : ex2_approx(xb);
      if (MASK) {
        a = (base + i < valid) ? a : 0.f;
        b = (base + i + 1 < valid) ? b : 0.f;
      }
      lsum = add2(lsum, pk2(a, b));
      o[k] = pack_bf16x2(a, b);
    }
    tmem_st4(tmem_p + g * 4, make_uint4(o[0], o[1], o[2], o[3]));
  }
}


}  // namespace vj
