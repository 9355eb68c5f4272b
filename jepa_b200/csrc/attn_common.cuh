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
// P / dS tile: [128 rows x 128 reduction] bf16, K-major, two 128B-swizzled atoms of 64 columns
VJ_DEVINL uint64_t ptile_desc(uint32_t tile, int kk) {
  return make_smem_desc(tile + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024, 2);
}


}  // namespace vj
