// GPU input pipeline: decoded uint8 frames -> random-resized crop (bilinear) -> horizontal flip ->
// per-channel normalisation -> network input layout, in one pass, so that uint8 frames (not fp32 clips, 4x the bytes)
// cross PCIe and no CPU core touches a pixel.
//
// Replaces, for the non-auto-augment path every shipped pre-training config uses (configs/pretrain/*.yaml):
//   app/vjepa/transforms.py:86-117   VideoTransform.__call__: float conversion, permute, spatial transform, flip, normalise
//   src/datasets/utils/video/transforms.py:545-577  random_resized_crop = crop [i:i+h, j:j+w] +
//       F.interpolate(mode='bilinear', align_corners=False)  (no antialiasing)
//   src/datasets/utils/video/transforms.py:160-190  horizontal_flip (applied AFTER the resize)
//   app/vjepa/transforms.py:140-153  _tensor_normalize_inplace: (x - 255 mean_c) / (255 std_c)
// The random parameters (crop box, flip) are drawn on the host in the reference's RNG call order
// (jepa_b200/transforms.py) and handed over as a small table; the kernel is deterministic.
//
// in  : uint8 frames of clip b at src + src_off[b], layout [T, H_b, W_b, 3] (what decord / the reference's loader yields)
// out : [B, 3, T, S, S] fp32 or bf16 (the Conv3d / vj_im2col_tubelets input layout)
// Bilinear sampling follows ATen's upsample_bilinear2d (align_corners = False): src = max(0, (dst + 0.5) * in/out - 0.5),
// the upper neighbour is clamped to the last row / column of the CROP.
#include "common.cuh"
#include "vjepa_b200.h"

namespace vj {

struct ClipParam {     // one per clip, 8 x int32 + 1 x int64 offset (host-built, device-resident)
  long long src_off;   // byte offset of the clip's first frame in `src`
  int H, W;            // decoded frame size
  int i, j, h, w;      // crop box: rows [i, i+h), columns [j, j+w)
  int flip;            // 1: mirror the OUTPUT horizontally
  int pad;
};

template <typename TO>
__global__ void __launch_bounds__(256) clip_preprocess_kernel(const uint8_t* __restrict__ src, const ClipParam* __restrict__ prm,
                                                              TO* __restrict__ out, int T, int S, float3 mean255,
                                                              float3 std255) {
  const int b = blockIdx.z, t = blockIdx.y;
  const ClipParam cp = prm[b];
  const uint8_t* frame = src + cp.src_off + (long long)t * cp.H * cp.W * 3;
  const float sh = float(cp.h) / float(S), sw = float(cp.w) / float(S);
  const long long plane = (long long)S * S;
  TO* ob = out + ((long long)b * 3 * T + t) * plane;      // channel c at + c * T * plane
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < S * S; idx += gridDim.x * blockDim.x) {
    const int y = idx / S, xo = idx - y * S;
    const int x = cp.flip ? (S - 1 - xo) : xo;              // output column xo shows resized column x
    float fy = fmaxf((float(y) + 0.5f) * sh - 0.5f, 0.f);
    float fx = fmaxf((float(x) + 0.5f) * sw - 0.5f, 0.f);
    const int y0 = min(int(fy), cp.h - 1), x0 = min(int(fx), cp.w - 1);
    const int y1 = min(y0 + 1, cp.h - 1), x1 = min(x0 + 1, cp.w - 1);
    const float ly = fy - float(y0), lx = fx - float(x0);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const uint8_t* r0 = frame + ((long long)(cp.i + y0) * cp.W + cp.j) * 3;
    const uint8_t* r1 = frame + ((long long)(cp.i + y1) * cp.W + cp.j) * 3;
    const float mean[3] = {mean255.x, mean255.y, mean255.z};
    const float sdev[3] = {std255.x, std255.y, std255.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float p00 = float(r0[x0 * 3 + c]), p01 = float(r0[x1 * 3 + c]);
      const float p10 = float(r1[x0 * 3 + c]), p11 = float(r1[x1 * 3 + c]);
      const float v = hy * (hx * p00 + lx * p01) + ly * (hx * p10 + lx * p11);
      ob[(long long)c * T * plane + idx] = TO(__fdiv_rn(v - mean[c], sdev[c]));   // sub_ then div_, as the reference
    }
  }
}

}  // namespace vj

extern "C" int vj_clip_preprocess(const void* src_u8, const void* params, void* out, int out_f32, int B, int T, int S,
                                  const float* mean3, const float* std3, void* stream_) {
  using namespace vj;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  VJ_CHECK_ARG(src_u8 && params && out && mean3 && std3, "vj_clip_preprocess: null pointer");
  VJ_CHECK_ARG(B > 0 && T > 0 && S > 0, "vj_clip_preprocess: empty problem");
  VJ_CHECK_ARG((reinterpret_cast<uintptr_t>(params) & 7) == 0, "vj_clip_preprocess: params must be 8-byte aligned");
  const float3 mean255 = make_float3(mean3[0] * 255.f, mean3[1] * 255.f, mean3[2] * 255.f);       // host arrays
  const float3 istd = make_float3(std3[0] * 255.f, std3[1] * 255.f, std3[2] * 255.f);
  dim3 grid((S * S + 255) / 256, T, B);
  if (grid.x > 64) grid.x = 64;
  if (out_f32)
    clip_preprocess_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(src_u8),
                                                       reinterpret_cast<const ClipParam*>(params),
                                                       reinterpret_cast<float*>(out), T, S, mean255, istd);
  else
    clip_preprocess_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(src_u8),
                                                               reinterpret_cast<const ClipParam*>(params),
                                                               reinterpret_cast<__nv_bfloat16*>(out), T, S, mean255, istd);
  VJ_CUDA(cudaGetLastError());
  vj::count_launch(1);
  return 0;
}
