// Memory-bound layout / pooling / normalisation kernels around the tensor-core GEMMs.  All activations are NHWC
// (channels contiguous) so every access below is a 16-byte vector along C and warps touch contiguous memory.
//
//   stem_im2col      NCHW fp32 image -> [pixels, 192] patch matrix for the 7x7/s2 stem (RV backbone conv1)
//   im2col_nhwc      patch matrix for the few strided convolutions (layerN.0.conv2 3x3/s2, downsample 1x1/s2)
//   maxpool3x3s2     torchvision resnet stem max-pool
//   upsample2x       nn.UpsamplingBilinear2d(scale_factor=2) == bilinear, align_corners=True
//                    (reference: RV/models/backbone.py:125, :140)
//   layernorm        nn.LayerNorm(256), eps 1e-5, fp32 statistics (RV/models/transformer.py:159-166)
#include "spe_internal.h"
#include "profile.h"
#include "spe_ptx.cuh"
#include <cuda_bf16.h>

namespace spe {

namespace {

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// 16-byte vector of T.  fp32 storage feeds kind::tf32 MMAs (10-bit mantissa): values are rounded to nearest on the
// way in, so the tensor core sees them exactly (idempotent for already-rounded data such as copies and max-pools).
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  float4 v;
  __device__ __forceinline__ float get(int i) const { return reinterpret_cast<const float*>(&v)[i]; }
  __device__ __forceinline__ void set(int i, float x) { reinterpret_cast<float*>(&v)[i] = rna_tf32(x); }
  __device__ __forceinline__ void set_exact(int i, float x) { reinterpret_cast<float*>(&v)[i] = x; }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  uint4 v;
  __device__ __forceinline__ float get(int i) const {
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&v)[i]);
  }
  __device__ __forceinline__ void set(int i, float x) {
    reinterpret_cast<__nv_bfloat16*>(&v)[i] = __float2bfloat16_rn(x);
  }
  __device__ __forceinline__ void set_exact(int i, float x) { set(i, x); }
};
template <typename T> __device__ __forceinline__ Vec<T> vload(const T* p) {
  Vec<T> r;
  r.v = *reinterpret_cast<const decltype(r.v)*>(p);
  return r;
}
template <typename T> __device__ __forceinline__ void vstore(T* p, const Vec<T>& x) {
  *reinterpret_cast<decltype(x.v)*>(p) = x.v;
}

// ---------------------------------------------------------------------------------------------------------------
// stem im2col: out[(n, oy, ox), (r*7 + s)*3 + c] = in[n, c, 2*oy - 3 + r, 2*ox - 3 + s]  (0 outside), K padded to 192
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStemK = 147, kStemKPad = 192;

template <typename T>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ in, int Hin, int Win, int Ho, int Wo, long long total_vec,
                   T* __restrict__ out) {
  constexpr int VN = Vec<T>::N;
  constexpr int VPR = kStemKPad / VN;  // vectors per output row
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total_vec) return;
  const int kv = static_cast<int>(idx % VPR);
  const long long pix = idx / VPR;
  const int ox = static_cast<int>(pix % Wo);
  const int oy = static_cast<int>((pix / Wo) % Ho);
  const long long n = pix / (static_cast<long long>(Wo) * Ho);
  const float* img = in + n * 3 * Hin * Win;
  Vec<T> r;
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    const int kk = kv * VN + e;
    float x = 0.f;
    if (kk < kStemK) {
      const int c = kk % 3, tap = kk / 3;
      const int fr = tap / 7, fs = tap % 7;
      const int iy = 2 * oy - 3 + fr, ix = 2 * ox - 3 + fs;
      if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) x = __ldg(img + (static_cast<long long>(c) * Hin + iy) * Win + ix);
    }
    r.set(e, x);
  }
  vstore(out + pix * kStemKPad + kv * VN, r);
}

// ---------------------------------------------------------------------------------------------------------------
// stem input re-layout: NCHW fp32 [B,3,H,W] -> zero-bordered NHWC [B, H+6, W+6, Cp] with Cp = 16 B / sizeof(T)
// (3 real channels + zero padding), so that 8 consecutive pixels form one 128-byte K-block that TMA can fetch as an
// overlapping window (GEMM "stem" mode) -- replaces the 600 MB im2col patch matrix.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
stem_pad_kernel(const float* __restrict__ in, int H, int W, long long total_px, T* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  constexpr int VN = Vec<T>::N;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total_px) return;
  const int Wp = W + 6, Hp = H + 6;
  const int x = static_cast<int>(idx % Wp) - 3;
  const int y = static_cast<int>((idx / Wp) % Hp) - 3;
  const long long n = idx / (static_cast<long long>(Wp) * Hp);
  Vec<T> r;
#pragma unroll
  for (int e = 0; e < VN; ++e) r.set(e, 0.f);
  if (x >= 0 && x < W && y >= 0 && y < H) {
    const float* px = in + (n * 3 * H + y) * W + x;
    const long long plane = static_cast<long long>(H) * W;
    r.set(0, __ldg(px));
    r.set(1, __ldg(px + plane));
    r.set(2, __ldg(px + 2 * plane));
  }
  vstore(out + idx * VN, r);
}

// ---------------------------------------------------------------------------------------------------------------
// generic NHWC im2col (used only for stride-2 convolutions)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
im2col_nhwc_kernel(const T* __restrict__ in, int H, int W, int C, int R, int S, int stride, int pad, int Ho, int Wo,
                   long long total_vec, T* __restrict__ out) {
  constexpr int VN = Vec<T>::N;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total_vec) return;
  const int cv = C / VN;
  const int c0 = static_cast<int>(idx % cv) * VN;
  long long rest = idx / cv;
  const int tap = static_cast<int>(rest % (R * S));
  const long long pix = rest / (R * S);
  const int ox = static_cast<int>(pix % Wo);
  const int oy = static_cast<int>((pix / Wo) % Ho);
  const long long n = pix / (static_cast<long long>(Wo) * Ho);
  const int fr = tap / S, fs = tap % S;
  const int iy = oy * stride - pad + fr, ix = ox * stride - pad + fs;
  Vec<T> r;
  if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
    r = vload(in + ((n * H + iy) * W + ix) * C + c0);
  } else {
#pragma unroll
    for (int e = 0; e < VN; ++e) r.set(e, 0.f);
  }
  vstore(out + (pix * (R * S) + tap) * C + c0, r);
}

// ---------------------------------------------------------------------------------------------------------------
// max-pool 3x3 stride 2 pad 1 (padding never wins: torch pads with -inf)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const T* __restrict__ in, int H, int W, int C, int Ho, int Wo, long long total_vec,
                    T* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  constexpr int VN = Vec<T>::N;
  // grid (x: vectors of one output row, y: output row, z: image): one 32-bit division per thread -- the flat 64-bit
  // index decode kept the XU pipe 18 % and the issue slots 72 % busy in a kernel that should only move bytes (ncu r02)
  (void)total_vec;
  const int cv = C / VN;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= static_cast<unsigned>(Wo * cv)) return;
  const int ox = static_cast<int>(t / static_cast<unsigned>(cv));
  const int c0 = static_cast<int>(t - static_cast<unsigned>(ox) * cv) * VN;
  const int oy = blockIdx.y;
  const long long n = blockIdx.z;
  const long long pix = (n * Ho + oy) * Wo + ox;
  float m[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) m[e] = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = 2 * oy - 1 + dy;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = 2 * ox - 1 + dx;
      if (ix < 0 || ix >= W) continue;
      const Vec<T> x = vload(in + ((n * H + iy) * W + ix) * C + c0);
#pragma unroll
      for (int e = 0; e < VN; ++e) m[e] = fmaxf(m[e], x.get(e));
    }
  }
  Vec<T> r;
#pragma unroll
  for (int e = 0; e < VN; ++e) r.set(e, m[e]);
  vstore(out + pix * C + c0, r);
}

// ---------------------------------------------------------------------------------------------------------------
// bilinear x2 upsample, align_corners=True (torch upsample_bilinear2d arithmetic: fp32 scale = (in-1)/(out-1))
// ---------------------------------------------------------------------------------------------------------------
// One CTA per TILE x TILE block of output pixels (all index math and the four bilinear weights are block-uniform);
// threads walk the channel vectors, so every load and store is a contiguous 16 bytes per lane.  With one pixel per CTA
// every output pixel pulled its four input pixels through L2 on its own: 4 x 205 MB of L2 -> SM traffic for a 51 MB
// input at B = 64 (93 us, 2.4 x the 39 us its HBM bytes need).  A 4 x 4 output tile touches at most 3 x 3 input pixels
// (36 KB for 1024 fp32 channels), so fifteen of its sixteen visits hit L1.
constexpr int kUpTile = 4;
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_kernel(const T* __restrict__ in, int H, int W, int C, T* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  constexpr int VN = Vec<T>::N;
  const int Ho = 2 * H, Wo = 2 * W;
  const long long n = blockIdx.z;
  const float sh = Ho > 1 ? static_cast<float>(H - 1) / static_cast<float>(Ho - 1) : 0.f;
  const float sw = Wo > 1 ? static_cast<float>(W - 1) / static_cast<float>(Wo - 1) : 0.f;
  const T* base = in + n * H * W * C;
  for (int dy = 0; dy < kUpTile; ++dy) {
    const int oy = blockIdx.y * kUpTile + dy;
    if (oy >= Ho) break;
    const float fy = sh * oy;
    const int y0 = static_cast<int>(fy);
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0);
    const float ly1 = fy - y0, ly0 = 1.f - ly1;
    for (int dx = 0; dx < kUpTile; ++dx) {
      const int ox = blockIdx.x * kUpTile + dx;
      if (ox >= Wo) break;
      const float fx = sw * ox;
      const int x0 = static_cast<int>(fx);
      const int x1 = x0 + (x0 < W - 1 ? 1 : 0);
      const float lx1 = fx - x0, lx0 = 1.f - lx1;
      const T* pa = base + (static_cast<long long>(y0) * W + x0) * C;
      const T* pb = base + (static_cast<long long>(y0) * W + x1) * C;
      const T* pc = base + (static_cast<long long>(y1) * W + x0) * C;
      const T* pd = base + (static_cast<long long>(y1) * W + x1) * C;
      T* po = out + ((n * Ho + oy) * Wo + ox) * C;
      for (int c0 = threadIdx.x * VN; c0 < C; c0 += blockDim.x * VN) {
        const Vec<T> a = vload(pa + c0), b = vload(pb + c0), c = vload(pc + c0), d = vload(pd + c0);
        Vec<T> r;
#pragma unroll
        for (int e = 0; e < VN; ++e)
          r.set(e, ly0 * (lx0 * a.get(e) + lx1 * b.get(e)) + ly1 * (lx0 * c.get(e) + lx1 * d.get(e)));
        vstore(po + c0, r);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (dim = 32 * 8 = 256): one warp per row, two-pass fp32 statistics in registers
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
layernorm256_kernel(const T* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                    long long rows, T* __restrict__ out, int exact, const float* __restrict__ gamma2,
                    const float* __restrict__ beta2, T* __restrict__ out2, float* __restrict__ out32) {
  pdl_wait();
  pdl_launch();
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  constexpr int VN = Vec<T>::N;
  constexpr int NV = 8 / VN;  // vectors per lane
  float x[8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const Vec<T> vv = vload(in + row * 256 + (i * 32 + lane) * VN);
#pragma unroll
    for (int e = 0; e < VN; ++e) x[i * VN + e] = vv.get(e);
  }
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) s += x[e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / 256.f);
  float q = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) { const float d = x[e] - mean; q += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.f / 256.f) + 1e-5f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * VN;
    Vec<T> r;
    float yv[VN];
    if (exact == 2) {
      // 3xTF32 operand layout [hi | lo | hi] (row stride 768): the consumer is a plain K = 768 GEMM against
      // [W_hi | W_hi | W_lo], i.e. A_hi W_hi + A_lo W_hi + A_hi W_lo without any in-kernel operand split
      Vec<T> hi, lo;
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const float y = (x[i * VN + e] - mean) * rstd * gamma[c + e] + beta[c + e];
        const float h = rna_tf32(y);
        hi.set_exact(e, h);
        lo.set(e, y - h);
      }
      vstore(out + row * 768 + c, hi);
      vstore(out + row * 768 + 256 + c, lo);
      vstore(out + row * 768 + 512 + c, hi);
      continue;
    }
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      const float y = (x[i * VN + e] - mean) * rstd * gamma[c + e] + beta[c + e];
      if (exact) r.set_exact(e, y); else r.set(e, y);
      x[i * VN + e] = r.get(e);          // what a second kernel would read back
      if (out32 != nullptr) yv[e] = rna_tf32(y);
    }
    vstore(out + row * 256 + c, r);
    if (out32 != nullptr) {
      // the same rows as fp32 holding TF32 values (bf16 storage: operand + residual of the fused feed-forward kernel,
      // which is built for fp32 / TF32 -- two more mantissa bits than the bf16 copy)
#pragma unroll
      for (int e = 0; e < VN; e += 4)
        *reinterpret_cast<float4*>(out32 + row * 256 + c + e) = make_float4(yv[e], yv[e + 1], yv[e + 2], yv[e + 3]);
    }
  }
  if (gamma2 == nullptr) {
    if (out2 != nullptr) {
      // second copy of the SAME rows rounded to the storage precision of a tensor-core operand (fp32 storage: TF32)
      // while `out` stays exact: operand and residual of the layer that follows
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        Vec<T> r;
#pragma unroll
        for (int e = 0; e < VN; ++e) r.set(e, x[i * VN + e]);
        vstore(out2 + row * 256 + (i * 32 + lane) * VN, r);
      }
    }
    return;
  }
  // chained second normalisation of the row just written (decoder: norm3 followed by the shared decoder norm,
  // RV/models/transformer.py:116-118), stored unrounded -- one launch instead of two, same values
  float s2 = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) s2 += x[e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  const float mean2 = s2 * (1.f / 256.f);
  float q2 = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) { const float d = x[e] - mean2; q2 += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, o);
  const float rstd2 = rsqrtf(q2 * (1.f / 256.f) + 1e-5f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * VN;
    Vec<T> r;
#pragma unroll
    for (int e = 0; e < VN; ++e) r.set_exact(e, (x[i * VN + e] - mean2) * rstd2 * gamma2[c + e] + beta2[c + e]);
    vstore(out2 + row * 256 + c, r);
  }
}

// The large plain LayerNorms (encoder: 50 k rows per batch of 64): the same arithmetic, in the same order, on FOUR rows
// per warp with all loads issued before the first use -- one warp per row keeps two 16-byte loads in flight per lane and
// measured ~3.4 TB/s on a tensor that fits L2; results are bit-identical to layernorm256_kernel.
template <typename T>
__global__ void __launch_bounds__(256)
layernorm256_rows4_kernel(const T* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                          long long rows, T* __restrict__ out, int exact) {
  pdl_wait();
  pdl_launch();
  constexpr int RW = 4;
  const long long row0 = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RW;
  if (row0 >= rows) return;
  const int lane = threadIdx.x & 31;
  constexpr int VN = Vec<T>::N;
  constexpr int NV = 8 / VN;
  Vec<T> raw[RW][NV];
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const long long row = row0 + r < rows ? row0 + r : rows - 1;     // tail rows re-read the last row (not stored)
#pragma unroll
    for (int i = 0; i < NV; ++i) raw[r][i] = vload(in + row * 256 + (i * 32 + lane) * VN);
  }
  float g[8], b[8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      g[i * VN + e] = gamma[(i * 32 + lane) * VN + e];
      b[i * VN + e] = beta[(i * 32 + lane) * VN + e];
    }
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    float x[8];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int e = 0; e < VN; ++e) x[i * VN + e] = raw[r][i].get(e);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += x[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / 256.f);
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float d = x[e] - mean; q += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / 256.f) + 1e-5f);
    if (row0 + r < rows) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        Vec<T> o;
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const float y = (x[i * VN + e] - mean) * rstd * g[i * VN + e] + b[i * VN + e];
          if (exact) o.set_exact(e, y); else o.set(e, y);
        }
        vstore(out + (row0 + r) * 256 + (i * 32 + lane) * VN, o);
      }
    }
  }
}

inline unsigned blocks_for(long long n, int per) { return static_cast<unsigned>((n + per - 1) / per); }

}  // namespace

#define DISPATCH_T(dt, ...)                                    \
  if ((dt) == kTF32) { using T = float; __VA_ARGS__; }         \
  else { using T = __nv_bfloat16; __VA_ARGS__; }

std::string launch_stem_im2col(Dtype dt, const float* nchw, int NB, int Hin, int Win, void* out, cudaStream_t s) {
  const int Ho = (Hin + 6 - 7) / 2 + 1, Wo = (Win + 6 - 7) / 2 + 1;
  ProfScope ps(kFamElementwise, s);
  DISPATCH_T(dt, {
    const long long total = static_cast<long long>(NB) * Ho * Wo * (kStemKPad / Vec<T>::N);
    stem_im2col_kernel<T><<<blocks_for(total, 256), 256, 0, s>>>(nchw, Hin, Win, Ho, Wo, total,
                                                                 reinterpret_cast<T*>(out));
  });
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_stem_pad(Dtype dt, const float* nchw, int NB, int Hin, int Win, void* out, cudaStream_t s) {
  ProfScope ps(kFamElementwise, s);
  DISPATCH_T(dt, {
    const long long total = static_cast<long long>(NB) * (Hin + 6) * (Win + 6);
    SPE_CUDA_TRY(launch_pdl(stem_pad_kernel<T>, dim3(blocks_for(total, 256)), dim3(256), 0, s, nchw, Hin, Win, total,
                            reinterpret_cast<T*>(out)));
  });
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_im2col_nhwc(Dtype dt, const void* in, int NB, int H, int W, int C, int R, int S, int stride,
                               int pad, void* out, cudaStream_t s) {
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
  ProfScope ps(kFamElementwise, s);
  DISPATCH_T(dt, {
    if (C % Vec<T>::N) return "im2col: C must be a multiple of the 16-byte vector";
    const long long total = static_cast<long long>(NB) * Ho * Wo * R * S * (C / Vec<T>::N);
    im2col_nhwc_kernel<T><<<blocks_for(total, 256), 256, 0, s>>>(reinterpret_cast<const T*>(in), H, W, C, R, S,
                                                                 stride, pad, Ho, Wo, total,
                                                                 reinterpret_cast<T*>(out));
  });
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_maxpool3x3s2(Dtype dt, const void* in, int NB, int H, int W, int C, void* out, cudaStream_t s) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  ProfScope ps(kFamElementwise, s);
  DISPATCH_T(dt, {
    const long long total = static_cast<long long>(NB) * Ho * Wo * (C / Vec<T>::N);
    const int row_vecs = Wo * (C / Vec<T>::N);
    const int threads = row_vecs >= 256 ? 256 : ((row_vecs + 31) / 32) * 32;
    if (Ho > 65535 || NB > 65535) return "maxpool: extent outside the launch grid";
    SPE_CUDA_TRY(launch_pdl(maxpool3x3s2_kernel<T>, dim3((row_vecs + threads - 1) / threads, Ho, NB), dim3(threads), 0, s,
                            reinterpret_cast<const T*>(in), H, W, C, Ho, Wo, total, reinterpret_cast<T*>(out)));
  });
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_upsample2x(Dtype dt, const void* in, int NB, int H, int W, int C, void* out, cudaStream_t s) {
  ProfScope ps(kFamElementwise, s);
  DISPATCH_T(dt, {
    const int nvec = C / Vec<T>::N;
    const int threads = nvec >= 256 ? 256 : ((nvec + 31) / 32) * 32;
    SPE_CUDA_TRY(launch_pdl(upsample2x_kernel<T>,
                            dim3((2 * W + kUpTile - 1) / kUpTile, (2 * H + kUpTile - 1) / kUpTile, NB), dim3(threads), 0, s,
                            reinterpret_cast<const T*>(in), H, W, C, reinterpret_cast<T*>(out)));
  });
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

// ---------------------------------------------------------------------------------------------------------------
// s16_latern at the resolution of its input.  The reference computes conv3x3(upsample2x(x)) (RV/models/backbone.py:
// 140-141, bilinear, align_corners=True).  Both are linear and the channel mixing commutes with the spatial operators:
//   conv3x3(U x)[p] = sum_taps W_tap (U x)[p + tap] = sum_taps (U (W_tap x))[p + tap]
// so the nine 1024 -> 256 tap matrices are applied to the 14 x 14 map (one GEMM, M = 196 per image instead of 784:
// a quarter of the multiply-adds) and this kernel gathers:  out[p, o] = sum over the taps that fall inside the 28 x 28
// map (zero padding) of the bilinear interpolation of Y_tap at p + tap.
// One CTA = one image x kTapCh output channels; its slice of Y ([H*W positions][9 taps][kTapCh]) is staged in shared
// memory once.  The bilinear interpolation is separable, and the tap index is (r, q) = (row shift, column shift), so the
// gather runs in two passes through shared memory:
//   columns:  Z[y, j, r] = sum_q  sum_{two source columns x of upsampled column j + q - 1}  cw * Y[y, x, (r, q)]
//   rows:     out[i, j]  = sum_r  sum_{two source rows y of upsampled row i + r - 1}        rw * Z[y, j, r]
// 6 + 6 multiply-adds per output element and channel instead of 36 -- and, what bounds this kernel, 2.4 x fewer
// shared-memory reads.  History (B = 64, profiles/r02*_launches_summary.md): v1 recomputed the interpolation indices per
// channel and tap, 269 us (the top launch of the step); v2, one thread per output position with 16 channels in
// registers, 126 us at its shared-memory-bandwidth floor; this is v3.  Interpolation tables (source index pair +
// weights of every upsampled row / column, zero weights outside the map = the convolution's zero padding) are built
// once per CTA.  Z rows are padded to kZStride words so the eight lanes of an LDS.128 fall on different banks.
constexpr int kTapCh = 8;
constexpr int kYStride = 9 * kTapCh;                  // words per source position
constexpr int kZStride = 3 * kTapCh + 4;              // words per (y, j): three tap rows + padding
template <typename TO>
__global__ void __launch_bounds__(256, 2)
upsample_tapsum_kernel(const float* __restrict__ Y, int H, int W, int Cout, TO* __restrict__ out, int out_ld, int round_tf32) {
  pdl_wait();
  pdl_launch();
  extern __shared__ __align__(16) float sy[];         // [H*W][kYStride] | Z [H][2W][kZStride] | tables
  const int b = blockIdx.x, c0 = blockIdx.y * kTapCh;
  const int HW = H * W, Ho = 2 * H, Wo = 2 * W;
  float* sz = sy + HW * kYStride;
  int* tab_i = reinterpret_cast<int*>(sz + H * Wo * kZStride);    // [Wo + 2] column source index, then [Ho + 2] row
  float* tab_w = reinterpret_cast<float*>(tab_i + (Wo + 2) + (Ho + 2));   // weight of the SECOND source (first = 1 - w), or -1 = outside
  const float* yb = Y + static_cast<long long>(b) * HW * 9 * Cout;
  // eight independent 16-byte loads in flight per thread
  const int nvec = HW * 9 * (kTapCh / 4);
  for (int u0 = threadIdx.x; u0 < nvec; u0 += 8 * blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int u = u0 + k * blockDim.x;
      if (u < nvec)
        v[k] = __ldg(reinterpret_cast<const float4*>(yb + static_cast<long long>(u / (kTapCh / 4)) * Cout + c0 + (u % (kTapCh / 4)) * 4));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int u = u0 + k * blockDim.x;
      if (u < nvec) *reinterpret_cast<float4*>(sy + (u / (kTapCh / 4)) * kTapCh + (u % (kTapCh / 4)) * 4) = v[k];
    }
  }
  // interpolation tables: entry t <-> upsampled coordinate t - 1 (nn.UpsamplingBilinear2d: align_corners=True)
  for (int t = threadIdx.x; t < (Wo + 2) + (Ho + 2); t += blockDim.x) {
    const bool col = t < Wo + 2;
    const int u = (col ? t : t - (Wo + 2)) - 1, n_out = col ? Wo : Ho, n_in = col ? W : H;
    if (u < 0 || u >= n_out) { tab_i[t] = 0; tab_w[t] = -1.f; continue; }
    const float f = (static_cast<float>(n_in - 1) / static_cast<float>(n_out - 1)) * u;
    const int i0 = static_cast<int>(f);
    tab_i[t] = i0;
    tab_w[t] = i0 < n_in - 1 ? f - i0 : 0.f;          // at the last source the second neighbour is the first again
  }
  __syncthreads();
  // ---- columns
  for (int it = threadIdx.x; it < H * Wo * 3; it += blockDim.x) {
    const int r = it % 3, yj = it / 3, j = yj % Wo, y = yj / Wo;
    float acc[kTapCh];
#pragma unroll
    for (int c = 0; c < kTapCh; ++c) acc[c] = 0.f;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float w1 = tab_w[j + q];
      if (w1 < 0.f) continue;                         // outside the 2W-wide map: zero padding
      const int x0 = tab_i[j + q];
      const int x1 = x0 + (x0 < W - 1 ? 1 : 0);
      const float w0 = 1.f - w1;
      const float* s0 = sy + (y * W + x0) * kYStride + (r * 3 + q) * kTapCh;
      const float* s1 = sy + (y * W + x1) * kYStride + (r * 3 + q) * kTapCh;
#pragma unroll
      for (int c = 0; c < kTapCh; c += 4) {
        const float4 a = *reinterpret_cast<const float4*>(s0 + c), bb = *reinterpret_cast<const float4*>(s1 + c);
        acc[c] += w0 * a.x + w1 * bb.x; acc[c + 1] += w0 * a.y + w1 * bb.y;
        acc[c + 2] += w0 * a.z + w1 * bb.z; acc[c + 3] += w0 * a.w + w1 * bb.w;
      }
    }
    float* z = sz + yj * kZStride + r * kTapCh;
#pragma unroll
    for (int c = 0; c < kTapCh; c += 4) *reinterpret_cast<float4*>(z + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
  }
  __syncthreads();
  // ---- rows
  const int* rtab_i = tab_i + (Wo + 2);
  const float* rtab_w = tab_w + (Wo + 2);
  for (int p = threadIdx.x; p < Ho * Wo; p += blockDim.x) {
    const int i = p / Wo, j = p % Wo;
    float acc[kTapCh];
#pragma unroll
    for (int c = 0; c < kTapCh; ++c) acc[c] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float w1 = rtab_w[i + r];
      if (w1 < 0.f) continue;
      const int y0 = rtab_i[i + r];
      const int y1 = y0 + (y0 < H - 1 ? 1 : 0);
      const float w0 = 1.f - w1;
      const float* s0 = sz + (y0 * Wo + j) * kZStride + r * kTapCh;
      const float* s1 = sz + (y1 * Wo + j) * kZStride + r * kTapCh;
#pragma unroll
      for (int c = 0; c < kTapCh; c += 4) {
        const float4 a = *reinterpret_cast<const float4*>(s0 + c), bb = *reinterpret_cast<const float4*>(s1 + c);
        acc[c] += w0 * a.x + w1 * bb.x; acc[c + 1] += w0 * a.y + w1 * bb.y;
        acc[c + 2] += w0 * a.z + w1 * bb.z; acc[c + 3] += w0 * a.w + w1 * bb.w;
      }
    }
    const long long o = (static_cast<long long>(b) * Ho * Wo + p) * out_ld + c0;
    if (sizeof(TO) == 4) {
      float* op = reinterpret_cast<float*>(out) + o;
#pragma unroll
      for (int c = 0; c < kTapCh; c += 4) {
        float4 v = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        if (round_tf32) {
          v.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xffffe000u);
          v.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xffffe000u);
          v.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xffffe000u);
          v.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xffffe000u);
        }
        *reinterpret_cast<float4*>(op + c) = v;
      }
    } else {
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out) + o;
      uint4 o8;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o8);
#pragma unroll
      for (int u = 0; u < 4; ++u) h[u] = __floats2bfloat162_rn(acc[2 * u], acc[2 * u + 1]);
      *reinterpret_cast<uint4*>(op) = o8;
    }
  }
}

std::string launch_upsample_tapsum(Dtype dt, const float* Y, int NB, int H, int W, int Cout, void* out, int out_ld,
                                   cudaStream_t s) {
  if (NB <= 0) return "";
  if (Cout % kTapCh) return "upsample_tapsum: output channels must be a multiple of 8";
  const size_t smem = (static_cast<size_t>(H) * W * kYStride + static_cast<size_t>(H) * 2 * W * kZStride) * sizeof(float) +
                      static_cast<size_t>(2 * W + 2 + 2 * H + 2) * 8;
  if (smem > 200 * 1024) return "upsample_tapsum: feature map too large for the shared-memory slice";
  ProfScope ps(kFamElementwise, s);
  if (dt == kTF32) {
    static bool attr = false;
    if (!attr) {
      SPE_CUDA_TRY(cudaFuncSetAttribute(upsample_tapsum_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = true;
    }
    SPE_CUDA_TRY(launch_pdl(upsample_tapsum_kernel<float>, dim3(NB, Cout / kTapCh), dim3(256), smem, s, Y, H, W, Cout,
                            reinterpret_cast<float*>(out), out_ld, 1));
  } else {
    static bool attr = false;
    if (!attr) {
      SPE_CUDA_TRY(cudaFuncSetAttribute(upsample_tapsum_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        200 * 1024));
      attr = true;
    }
    SPE_CUDA_TRY(launch_pdl(upsample_tapsum_kernel<__nv_bfloat16>, dim3(NB, Cout / kTapCh), dim3(256), smem, s, Y, H, W, Cout,
                            reinterpret_cast<__nv_bfloat16*>(out), out_ld, 0));
  }
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_layernorm(Dtype dt, const void* in, const float* gamma, const float* beta, long long rows,
                             int dim, void* out, cudaStream_t s, int exact, const float* gamma2, const float* beta2,
                             void* out2, float* out_f32) {
  if (dim != 256) return "layernorm: only hidden_dim 256 is built";
  if (rows <= 0) return "";
  ProfScope ps(kFamElementwise, s);
  static const int rows4 = getenv("SPE_LN_ROWS4") ? atoi(getenv("SPE_LN_ROWS4")) : 1;
  if (rows4 && rows >= 8192 && exact != 2 && gamma2 == nullptr && out2 == nullptr && out_f32 == nullptr) {
    DISPATCH_T(dt, {
      SPE_CUDA_TRY(launch_pdl(layernorm256_rows4_kernel<T>, dim3(blocks_for(rows, 32)), dim3(256), 0, s,
                              reinterpret_cast<const T*>(in), gamma, beta, rows, reinterpret_cast<T*>(out), exact));
    });
    SPE_CUDA_TRY(cudaGetLastError());
    return "";
  }
  DISPATCH_T(dt, {
    SPE_CUDA_TRY(launch_pdl(layernorm256_kernel<T>, dim3(blocks_for(rows, 8)), dim3(256), 0, s,
                            reinterpret_cast<const T*>(in), gamma, beta, rows, reinterpret_cast<T*>(out), exact, gamma2,
                            beta2, reinterpret_cast<T*>(out2), out_f32));
  });
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

}  // namespace spe
