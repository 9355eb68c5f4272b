// Microbenchmark: tcgen05.ld throughput per SM as a function of resident warps (is TMEM read a shared
// 64 B/clk port or per-scheduler?).  Prints bytes/clk/SM.  Not a correctness test.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../jepa_b200/csrc/common.cuh"

using namespace vj;

template <int X>
__global__ void tmem_ld_bench(long long* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<256>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (X == 32) {
      uint32_t v[32];
      tmem_ld32(base + ((i & 3) * 32), v);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; ++e) acc ^= v[e];
    } else {
      uint32_t v[16];
      tmem_ld16(base + ((i & 7) * 16), v);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 16; ++e) acc ^= v[e];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (lane == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
  if (acc == 0x12345678u) out[1000] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(slot);
}

// two loads in flight per warp before the wait
__global__ void tmem_ld_bench2(long long* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<256>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i += 2) {
    uint32_t v[32], w[32];
    tmem_ld32(base, v);
    tmem_ld32(base + 32, w);
    tmem_wait_ld();
#pragma unroll
    for (int e = 0; e < 32; ++e) acc ^= v[e] ^ w[e];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (lane == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
  if (acc == 0x12345678u) out[1000] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(slot);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8192 * 8);
  const int iters = 2000;
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int warps : {1, 4, 8, 16}) {
      if (ctas * warps > 32) continue;
      for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(d, 0, 8192 * 8);
        const int grid = 148 * ctas;   // ctas per SM co-resident (256 TMEM columns each)
        if (mode == 0) tmem_ld_bench<32><<<grid, warps * 32>>>(d, iters);
        else if (mode == 1) tmem_ld_bench<16><<<grid, warps * 32>>>(d, iters);
        else tmem_ld_bench2<<<grid, warps * 32>>>(d, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[32];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
        const double bytes_per_instr = mode == 1 ? 2048.0 : 4096.0;
        const double bpc = bytes_per_instr * iters * warps * ctas / (double)mx;
        printf("ctas/SM=%d warps/CTA=%2d %s: %.1f clk per ld per warp, %.1f B/clk/SM\n", ctas, warps,
               mode == 0 ? "x32       " : (mode == 1 ? "x16       " : "x32 2-deep"), (double)mx / iters, bpc);
      }
    }
  return 0;
}
