// Final projections of the prediction heads, kept in fp32 (a 0.5 px keypoint budget at crop sides up to ~1750 px
// leaves ~3e-4 in normalised units, so the last layer and the sigmoid never see reduced precision):
//   logits = cls_embed(hs)                         Linear 256 -> 12     (reference: RV/models/detr_speed.py:50, :83)
//   points = sigmoid(point_embed.layers[2](h2))    Linear 256 -> 2      (RV/models/detr_speed.py:16-29, :52, :84)
//   log_sigma = sigma_embed.layers[2](s2)          Linear 256 -> 1, duplicated to (x, y)
//                                                  (SA/src/zoo/rtdetr/rtdetr_decoder.py:295-297, :367)
// The two hidden 256 -> 256 layers of each MLP run on the tensor-core GEMM; this kernel is one warp per query row.
#include "spe_internal.h"
#include "profile.h"
#include "spe_ptx.cuh"
#include <cuda_bf16.h>

namespace spe {

namespace {

template <typename T> __device__ __forceinline__ void load_row8(const T* p, float (&x)[8]);
template <> __device__ __forceinline__ void load_row8<float>(const float* p, float (&x)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
template <> __device__ __forceinline__ void load_row8<__nv_bfloat16>(const __nv_bfloat16* p, float (&x)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float2 f = __bfloat1622float2(h[u]);
    x[2 * u] = f.x;
    x[2 * u + 1] = f.y;
  }
}

__device__ __forceinline__ float warp_dot(const float (&x)[8], const float* __restrict__ w, int lane) {
  const float4 a = *reinterpret_cast<const float4*>(w + lane * 8);
  const float4 b = *reinterpret_cast<const float4*>(w + lane * 8 + 4);
  float s = x[0] * a.x;
  s = fmaf(x[1], a.y, s); s = fmaf(x[2], a.z, s); s = fmaf(x[3], a.w, s);
  s = fmaf(x[4], b.x, s); s = fmaf(x[5], b.y, s); s = fmaf(x[6], b.z, s); s = fmaf(x[7], b.w, s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

template <typename T>
__global__ void __launch_bounds__(256)
head_final_kernel(const T* __restrict__ hs, const T* __restrict__ h2, const T* __restrict__ s2, long long rows,
                  const float* __restrict__ Wc, const float* __restrict__ bc, const float* __restrict__ W3,
                  const float* __restrict__ b3, const float* __restrict__ Ws3, const float* __restrict__ bs3,
                  float* __restrict__ logits, float* __restrict__ points, float* __restrict__ logsig) {
  pdl_wait();
  pdl_launch();
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float x[8];
  load_row8(hs + row * 256 + lane * 8, x);
  for (int c = 0; c < 12; ++c) {
    const float d = warp_dot(x, Wc + c * 256, lane);
    if (lane == 0) logits[row * 12 + c] = d + bc[c];
  }
  load_row8(h2 + row * 256 + lane * 8, x);
  for (int c = 0; c < 2; ++c) {
    const float d = warp_dot(x, W3 + c * 256, lane) + b3[c];
    if (lane == 0) points[row * 2 + c] = 1.0f / (1.0f + expf(-d));
  }
  if (s2 != nullptr) {
    load_row8(s2 + row * 256 + lane * 8, x);
    const float d = warp_dot(x, Ws3, lane) + bs3[0];
    if (lane == 0) { logsig[row * 2] = d; logsig[row * 2 + 1] = d; }
  }
}

}  // namespace

std::string launch_head_final(Dtype dt, const void* hs, const void* h2, const void* s2, long long rows,
                              const float* Wc, const float* bc, const float* W3, const float* b3,
                              const float* Ws3, const float* bs3, float* logits, float* points, float* logsig,
                              cudaStream_t s) {
  if (rows <= 0) return "";
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  ProfScope ps(kFamHeads, s);
  if (dt == kTF32) {
    SPE_CUDA_TRY(launch_pdl(head_final_kernel<float>, dim3(blocks), dim3(256), 0, s,
                            reinterpret_cast<const float*>(hs), reinterpret_cast<const float*>(h2),
                            reinterpret_cast<const float*>(s2), rows, Wc, bc, W3, b3, Ws3, bs3, logits, points,
                            logsig));
  } else {
    SPE_CUDA_TRY(launch_pdl(head_final_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, s,
                            reinterpret_cast<const __nv_bfloat16*>(hs), reinterpret_cast<const __nv_bfloat16*>(h2),
                            reinterpret_cast<const __nv_bfloat16*>(s2), rows, Wc, bc, W3, b3, Ws3, bs3, logits, points,
                            logsig));
  }
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

}  // namespace spe
