// LayerNorm forward / backward (nn.LayerNorm(eps=1e-6) at src/models/utils/modules.py:115,119,
// src/models/vision_transformer.py:192-193, src/models/predictor.py:233).  HBM-bound: one warp per row,
// 16-byte vector accesses, the row lives in registers between the two passes.  Kernels are templated on
// NV = ceil(D / 256) (8-element chunks per lane) so no dead registers are carried and >= 4 CTAs of 8
// warps stay resident per SM; the grid is a multiple of the SM count.
#include "common.cuh"
#include "vjepa_b200.h"

namespace vj {

template <bool F32>
VJ_DEVINL void ld8(const void* base, long long off, float (&v)[8]) {
  if (F32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    const float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
    v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
    v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
  }
}
template <bool F32>
VJ_DEVINL void st8(void* base, long long off, const float (&v)[8]) {
  if (F32) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = u;
  }
}

template <int NV, bool IN_F32, bool OUT_F32>
__global__ void __launch_bounds__(256, NV <= 2 ? 4 : (NV <= 4 ? 3 : 1))
ln_fwd_kernel(const void* __restrict__ x, void* __restrict__ y, const float* __restrict__ gamma,
              const float* __restrict__ beta, float* __restrict__ mean_out, float* __restrict__ rstd_out, int T, int D,
              float eps) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nvec = D >> 3;
  const float invD = 1.0f / D;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < T; row += (long long)gridDim.x * wpb) {
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        ld8<IN_F32>(x, row * D + c * 8, v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    }
    const float mean = warp_sum(s) * invD;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          ss += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) * invD + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float o[8], g[8], b[8];   // gamma / beta come from L1 (same 4 KB for every row)
        ld8<true>(gamma, c * 8, g);
        ld8<true>(beta, c * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
        st8<OUT_F32>(y, row * D + c * 8, o);
      }
    }
  }
}

// dx = dres + rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)).  The row (x, dy) stays in registers in
// its storage format between the two passes; dgamma/dbeta partial sums live in per-warp shared-memory
// slices (no atomics, no persistent registers), reduced per block into [gridDim.x, D] partials.
template <int NV, bool X_F32>
// (register caps chosen so nothing spills: with ~220 KB of the SM given to shared memory L1 is tiny and every
// local-memory access is an L2 round trip)
__global__ void __launch_bounds__(256, NV <= 2 ? 3 : (NV <= 4 ? 2 : 1))
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const void* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, const void* __restrict__ dres,
              void* __restrict__ dx, float* __restrict__ part_dgamma, float* __restrict__ part_dbeta, int T, int D) {
  extern __shared__ float sm[];  // [warps][2][D]
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int nvec = D >> 3;
  const float invD = 1.0f / D;
  float* my_dg = sm + (size_t)wib * 2 * D;
  float* my_db = my_dg + D;
  for (int i = lane; i < 2 * D; i += 32) my_dg[i] = 0.f;
  __syncwarp();
  for (long long row = (long long)blockIdx.x * wpb + wib; row < T; row += (long long)gridDim.x * wpb) {
    const float mu = mean[row], rs = rstd[row];
    float xv[NV][X_F32 ? 8 : 1];
    uint4 xp[NV], dyp[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float xf[8], dyf[8], g[8];
        if (X_F32) {
          ld8<true>(x, row * D + c * 8, xf);
#pragma unroll
          for (int j = 0; j < (X_F32 ? 8 : 1); ++j) xv[i][j] = xf[j];
        } else {
          xp[i] = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + row * D + c * 8);
          xf[0] = bf16_lo(xp[i].x); xf[1] = bf16_hi(xp[i].x); xf[2] = bf16_lo(xp[i].y); xf[3] = bf16_hi(xp[i].y);
          xf[4] = bf16_lo(xp[i].z); xf[5] = bf16_hi(xp[i].z); xf[6] = bf16_lo(xp[i].w); xf[7] = bf16_hi(xp[i].w);
        }
        dyp[i] = *reinterpret_cast<const uint4*>(dy + row * D + c * 8);
        dyf[0] = bf16_lo(dyp[i].x); dyf[1] = bf16_hi(dyp[i].x); dyf[2] = bf16_lo(dyp[i].y); dyf[3] = bf16_hi(dyp[i].y);
        dyf[4] = bf16_lo(dyp[i].z); dyf[5] = bf16_hi(dyp[i].z); dyf[6] = bf16_lo(dyp[i].w); dyf[7] = bf16_hi(dyp[i].w);
        ld8<true>(gamma, c * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gy = g[j] * dyf[j];
          s1 += gy;
          s2 += gy * ((xf[j] - mu) * rs);
        }
      }
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float xf[8], dyf[8], g[8], o[8];
        if (X_F32) {
#pragma unroll
          for (int j = 0; j < (X_F32 ? 8 : 1); ++j) xf[j] = xv[i][j];
        } else {
          xf[0] = bf16_lo(xp[i].x); xf[1] = bf16_hi(xp[i].x); xf[2] = bf16_lo(xp[i].y); xf[3] = bf16_hi(xp[i].y);
          xf[4] = bf16_lo(xp[i].z); xf[5] = bf16_hi(xp[i].z); xf[6] = bf16_lo(xp[i].w); xf[7] = bf16_hi(xp[i].w);
        }
        dyf[0] = bf16_lo(dyp[i].x); dyf[1] = bf16_hi(dyp[i].x); dyf[2] = bf16_lo(dyp[i].y); dyf[3] = bf16_hi(dyp[i].y);
        dyf[4] = bf16_lo(dyp[i].z); dyf[5] = bf16_hi(dyp[i].z); dyf[6] = bf16_lo(dyp[i].w); dyf[7] = bf16_hi(dyp[i].w);
        ld8<true>(gamma, c * 8, g);
        float4* pg = reinterpret_cast<float4*>(my_dg + c * 8);
        float4* pb = reinterpret_cast<float4*>(my_db + c * 8);
        float4 g0 = pg[0], g1 = pg[1], b0 = pb[0], b1 = pb[1];
        float xh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[j] = (xf[j] - mu) * rs;
          o[j] = rs * (g[j] * dyf[j] - s1 - xh[j] * s2);
        }
        g0.x += dyf[0] * xh[0]; g0.y += dyf[1] * xh[1]; g0.z += dyf[2] * xh[2]; g0.w += dyf[3] * xh[3];
        g1.x += dyf[4] * xh[4]; g1.y += dyf[5] * xh[5]; g1.z += dyf[6] * xh[6]; g1.w += dyf[7] * xh[7];
        b0.x += dyf[0]; b0.y += dyf[1]; b0.z += dyf[2]; b0.w += dyf[3];
        b1.x += dyf[4]; b1.y += dyf[5]; b1.z += dyf[6]; b1.w += dyf[7];
        pg[0] = g0; pg[1] = g1; pb[0] = b0; pb[1] = b1;
        if (dres) {
          float r[8];
          ld8<X_F32>(dres, row * D + c * 8, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        st8<X_F32>(dx, row * D + c * 8, o);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < wpb; ++w) {
      a += sm[(size_t)w * 2 * D + i];
      b += sm[(size_t)w * 2 * D + D + i];
    }
    part_dgamma[(long long)blockIdx.x * D + i] = a;
    part_dbeta[(long long)blockIdx.x * D + i] = b;
  }
}

// out_a[c] += sum_r a[r,c]; out_b[c] += sum_r b[r,c].  grid = (C/32, row chunks): 32 columns x 8 row groups per
// block, coalesced 128-byte row segments, one atomicAdd per (column, row chunk).
__global__ void __launch_bounds__(256) partial_reduce_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                             float* __restrict__ out_a, float* __restrict__ out_b,
                                                             int R, int C, int rows_per_block) {
  __shared__ float sm[2][8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rg = threadIdx.x >> 5;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  float sa = 0.f, sb = 0.f;
  if (col < C) {
    for (int r = r0 + rg; r < r1; r += 8) {
      sa += a[(long long)r * C + col];
      sb += b[(long long)r * C + col];
    }
  }
  sm[0][rg][threadIdx.x & 31] = sa;
  sm[1][rg][threadIdx.x & 31] = sb;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int which = threadIdx.x >> 5, cc = threadIdx.x & 31;
    const int c = blockIdx.x * 32 + cc;
    if (c < C) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += sm[which][i][cc];
      atomicAdd(&(which ? out_b : out_a)[c], s);
    }
  }
}

static int ln_grid(int T) {
  long long g = (T + 7) / 8;
  const long long cap = (long long)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return int(g);
}
static int ln_bwd_grid(int T) {
  long long g = (T + 7) / 8;
  const long long cap = (long long)num_sms() * 4;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return int(g);
}

template <int NV>
static void launch_ln_fwd(const void* x, int x_f32, void* y, int y_f32, const float* gamma, const float* beta, float* mean,
                          float* rstd, int T, int D, float eps, cudaStream_t s) {
  const int grid = ln_grid(T);
  if (x_f32 && y_f32) ln_fwd_kernel<NV, true, true><<<grid, 256, 0, s>>>(x, y, gamma, beta, mean, rstd, T, D, eps);
  else if (x_f32) ln_fwd_kernel<NV, true, false><<<grid, 256, 0, s>>>(x, y, gamma, beta, mean, rstd, T, D, eps);
  else if (y_f32) ln_fwd_kernel<NV, false, true><<<grid, 256, 0, s>>>(x, y, gamma, beta, mean, rstd, T, D, eps);
  else ln_fwd_kernel<NV, false, false><<<grid, 256, 0, s>>>(x, y, gamma, beta, mean, rstd, T, D, eps);
}

template <int NV>
static void launch_ln_bwd(const void* dy, const void* x, int x_f32, const float* gamma, const float* mean,
                          const float* rstd, const void* dres, void* dx, float* pg, float* pb, int grid, int T, int D,
                          cudaStream_t s) {
  const size_t smem = (size_t)8 * 2 * D * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(ln_bwd_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 2048 * 4);
    cudaFuncSetAttribute(ln_bwd_kernel<NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 2048 * 4);
    configured = true;
  }
  if (x_f32)
    ln_bwd_kernel<NV, true><<<grid, 256, smem, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy), x, gamma, mean, rstd,
                                                    dres, dx, pg, pb, T, D);
  else
    ln_bwd_kernel<NV, false><<<grid, 256, smem, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy), x, gamma, mean, rstd,
                                                     dres, dx, pg, pb, T, D);
}

}  // namespace vj

using namespace vj;

#define VJ_LN_DISPATCH(D, CALL)                   \
  do {                                            \
    const int nv_ = ((D) + 255) / 256;            \
    if (nv_ <= 1) { CALL(1); }                    \
    else if (nv_ == 2) { CALL(2); }               \
    else if (nv_ <= 4) { CALL(4); }               \
    else if (nv_ == 5) { CALL(5); }               \
    else { CALL(8); }                             \
  } while (0)

extern "C" int vj_layernorm_fwd(const void* x, int x_f32, void* y, int y_f32, const float* gamma, const float* beta,
                                float* mean, float* rstd, int T, int D, float eps, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  if (T <= 0) return 0;
  VJ_CHECK_ARG(x && y && gamma && beta, "vj_layernorm_fwd: null pointer");
  VJ_CHECK_ARG(D % 8 == 0 && D <= 2048, "vj_layernorm_fwd: D=%d unsupported (multiple of 8, <= 2048)", D);
#define VJ_CALL(NV) launch_ln_fwd<NV>(x, x_f32, y, y_f32, gamma, beta, mean, rstd, T, D, eps, s)
  VJ_LN_DISPATCH(D, VJ_CALL);
#undef VJ_CALL
  VJ_CUDA(cudaGetLastError());
  vj::count_launch(1);
  return 0;
}

extern "C" size_t vj_layernorm_bwd_workspace(int T, int D) {
  (void)T;
  return (size_t)4 * num_sms() * 2 * D * sizeof(float);
}

extern "C" int vj_layernorm_bwd(const void* dy, const void* x, int x_f32, const float* gamma, const float* mean,
                                const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                                void* workspace, size_t ws_bytes, int T, int D, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  if (T <= 0) return 0;
  VJ_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace, "vj_layernorm_bwd: null pointer");
  VJ_CHECK_ARG(D % 8 == 0 && D <= 2048, "vj_layernorm_bwd: D=%d unsupported (multiple of 8, <= 2048)", D);
  const int grid = ln_bwd_grid(T);
  VJ_CHECK_ARG(ws_bytes >= (size_t)grid * 2 * D * sizeof(float), "vj_layernorm_bwd: workspace too small");
  float* pg = reinterpret_cast<float*>(workspace);
  float* pb = pg + (size_t)grid * D;
#define VJ_CALL(NV) launch_ln_bwd<NV>(dy, x, x_f32, gamma, mean, rstd, dres, dx, pg, pb, grid, T, D, s)
  VJ_LN_DISPATCH(D, VJ_CALL);
#undef VJ_CALL
  VJ_CUDA(cudaGetLastError());
  {
    const int rows_per_block = 64;
    dim3 rgrid((D + 31) / 32, (grid + rows_per_block - 1) / rows_per_block);
    partial_reduce_kernel<<<rgrid, 256, 0, s>>>(pg, pb, dgamma, dbeta, grid, D, rows_per_block);
  }
  VJ_CUDA(cudaGetLastError());
  vj::count_launch(2);
  return 0;
}
