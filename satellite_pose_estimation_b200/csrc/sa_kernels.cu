// Memory-bound kernels of the SA (RT-DETR) keypoint predictor, SURVEY.md section 8f rank 2 -- everything of
// PResNet-50-vd / HybridEncoder / RTDETRTransformer that is not a tensor-core GEMM, an attention or a LayerNorm
// (those reuse gemm_tc.cu / attention.cu / elementwise.cu):
//
//   sa_stem_im2col_kernel     3x3 / stride-2 patches of the NCHW input image -> [pixels, 32] rows (27 taps + 5 zeros):
//                             conv1_1 of the "vd" stem as a plain GEMM        SA/nn/backbone/presnet.py:166-172
//   avgpool2x2_kernel         AvgPool2d(2, 2) of the variant-d shortcut        SA/nn/backbone/presnet.py:93-104
//   act_rows_kernel           SiLU / GELU (+ optional second operand) between GEMMs whose epilogue only knows ReLU
//                             SA/src/zoo/rtdetr/hybrid_encoder.py:47-53, :119-123, :164
//   upsample_nearest2x_kernel F.interpolate(scale_factor=2, mode="nearest")    hybrid_encoder.py:377-379
//   bicubic_half_kernel       F.interpolate(scale_factor=0.5, mode="bicubic")  hybrid_encoder.py:394
//   small_linear_kernel       Linear layers with <= 16 outputs (class scores, last layer of the keypoint MLP) in fp32
//                             SA/src/zoo/rtdetr/rtdetr_decoder.py:631-635
//   query_pos_hidden_kernel   first layer of query_pos_head (2 -> 512, ReLU)   rtdetr_decoder.py:323, :458
//   sa_head_kernel            per decoder layer: class logits, refined keypoint sigmoid(delta + inverse_sigmoid(ref)),
//                             log-sigma                                        rtdetr_decoder.py:338-367
//
// Activations are NHWC fp32; values that feed a kind::tf32 MMA are rounded to TF32 by their producer (round = 1).
#include "spe_internal.h"
#include "profile.h"

#include <math.h>

namespace spe {

namespace {

__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

inline unsigned blocks_for(long long total, int threads) {
  return static_cast<unsigned>((total + threads - 1) / threads);
}

// out[(b, ho, wo), (r*3 + s)*3 + c] = img[b, c, 2 ho - 1 + r, 2 wo - 1 + s]  (zero outside), columns 27..31 zero
__global__ void sa_stem_im2col_kernel(const float* __restrict__ img, int H, int W, int Ho, int Wo, long long total,
                                      int round, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // one (pixel, tap) per thread
  if (i >= total) return;
  const int tap = static_cast<int>(i % 11);          // taps 0..8, 9 and 10 write the zero tail
  const long long pix = i / 11;
  const int wo = static_cast<int>(pix % Wo);
  const int ho = static_cast<int>((pix / Wo) % Ho);
  const long long b = pix / (static_cast<long long>(Wo) * Ho);
  float* o = out + pix * 32;
  if (tap >= 9) {
    if (tap == 9) { o[27] = 0.f; o[28] = 0.f; o[29] = 0.f; }
    else { o[30] = 0.f; o[31] = 0.f; }
    return;
  }
  const int r = tap / 3, s = tap % 3;
  const int y = 2 * ho - 1 + r, x = 2 * wo - 1 + s;
  const bool in = y >= 0 && y < H && x >= 0 && x < W;
  const float* p = img + (b * 3 * H + y) * static_cast<long long>(W) + x;
  const long long plane = static_cast<long long>(H) * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[tap * 3 + c] = in ? (round ? rna_tf32(p[c * plane]) : p[c * plane]) : 0.f;
}

// NHWC 2x2 / stride-2 average (H, W even), float4 per thread
__global__ void avgpool2x2_kernel(const float4* __restrict__ in, int H, int W, int C4, long long total, int round,
                                  float4* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Ho = H / 2, Wo = W / 2;
  const int c = static_cast<int>(i % C4);
  const long long pix = i / C4;
  const int wo = static_cast<int>(pix % Wo);
  const int ho = static_cast<int>((pix / Wo) % Ho);
  const long long b = pix / (static_cast<long long>(Wo) * Ho);
  const float4* p = in + ((b * H + 2 * ho) * W + 2 * wo) * C4 + c;
  const float4 a = p[0], bq = p[C4], cq = p[static_cast<long long>(W) * C4], d = p[static_cast<long long>(W) * C4 + C4];
  float4 r;
  r.x = (a.x + bq.x + cq.x + d.x) * 0.25f;
  r.y = (a.y + bq.y + cq.y + d.y) * 0.25f;
  r.z = (a.z + bq.z + cq.z + d.z) * 0.25f;
  r.w = (a.w + bq.w + cq.w + d.w) * 0.25f;
  if (round) { r.x = rna_tf32(r.x); r.y = rna_tf32(r.y); r.z = rna_tf32(r.z); r.w = rna_tf32(r.w); }
  out[i] = r;
}

__device__ __forceinline__ float act_apply(float v, int kind) {
  // SiLU: ex2.approx + one approximate division (relative error ~4e-7, far below the 3xTF32 products around it); the
  // IEEE expf + division made this memory-bound pass issue-bound (ncu r02: SM throughput 80 %, 45 us for 93 MB)
  if (kind == 1) return __fdividef(v, 1.f + __expf(-v));
  if (kind == 2) return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));  // GELU (erf form, nn.GELU default)
  if (kind == 3) return 1.f / (1.f + expf(-v));                             // sigmoid
  return v;
}

// out[r, c] = act(in[r, c]) (+ add[r, c]), rows with independent strides (so that results land in channel slices of
// a concatenated tensor); C a multiple of 4
__global__ void act_rows_kernel(const float* __restrict__ in, int in_ld, const float* __restrict__ add, int add_ld,
                                float* __restrict__ out, int out_ld, long long rows, int C4, int kind, int round) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * C4) return;
  const long long r = i / C4;
  const int c = static_cast<int>(i % C4) * 4;
  float4 v = *reinterpret_cast<const float4*>(in + r * in_ld + c);
  v.x = act_apply(v.x, kind); v.y = act_apply(v.y, kind); v.z = act_apply(v.z, kind); v.w = act_apply(v.w, kind);
  if (add != nullptr) {
    const float4 a = *reinterpret_cast<const float4*>(add + r * add_ld + c);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  if (round) { v.x = rna_tf32(v.x); v.y = rna_tf32(v.y); v.z = rna_tf32(v.z); v.w = rna_tf32(v.w); }
  *reinterpret_cast<float4*>(out + r * out_ld + c) = v;
}

// out[b, y, x, :] = in[b, y / 2, x / 2, :]   (nearest, exact scale 2)
__global__ void upsample_nearest2x_kernel(const float* __restrict__ in, int in_ld, int H, int W, int C4, long long total,
                                          float* __restrict__ out, int out_ld) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C4);
  const long long pix = i / C4;
  const int x = static_cast<int>(pix % (2 * W));
  const int y = static_cast<int>((pix / (2 * W)) % (2 * H));
  const long long b = pix / (4ll * W * H);
  const float4 v = *reinterpret_cast<const float4*>(in + ((b * H + y / 2) * W + x / 2) * in_ld + c * 4);
  *reinterpret_cast<float4*>(out + pix * out_ld + c * 4) = v;
}

// Bicubic x0.5 as torch computes it (upsample_bicubic2d, align_corners=False, A = -0.75, no antialias): the source
// coordinate of output o is 2 o + 0.5, so the four taps 2o-1 .. 2o+2 always carry the weights of t = 0.5,
// (-0.09375, 0.59375, 0.59375, -0.09375), with indices clamped to the map.
__global__ void bicubic_half_kernel(const float* __restrict__ in, int H, int W, int C4, long long total, int round,
                                    float* __restrict__ out, int out_ld) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Ho = H / 2, Wo = W / 2;
  const int c = static_cast<int>(i % C4);
  const long long pix = i / C4;
  const int xo = static_cast<int>(pix % Wo);
  const int yo = static_cast<int>((pix / Wo) % Ho);
  const long long b = pix / (static_cast<long long>(Wo) * Ho);
  // cubic_convolution2(x + 1, A), cubic_convolution1(x, A), ... at x = 0.5, in torch's evaluation order
  const float A = -0.75f, t = 0.5f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = (1.f - t) + 1.f;
  float wgt[4];
  wgt[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  wgt[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  wgt[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  wgt[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
  const float4* base = reinterpret_cast<const float4*>(in) + b * H * W * C4 + c;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int y = min(max(2 * yo - 1 + r, 0), H - 1);
    float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int x = min(max(2 * xo - 1 + s, 0), W - 1);
      const float4 v = base[(static_cast<long long>(y) * W + x) * C4];
      row.x += v.x * wgt[s]; row.y += v.y * wgt[s]; row.z += v.z * wgt[s]; row.w += v.w * wgt[s];
    }
    acc.x += row.x * wgt[r]; acc.y += row.y * wgt[r]; acc.z += row.z * wgt[r]; acc.w += row.w * wgt[r];
  }
  if (round) { acc.x = rna_tf32(acc.x); acc.y = rna_tf32(acc.y); acc.z = rna_tf32(acc.z); acc.w = rna_tf32(acc.w); }
  *reinterpret_cast<float4*>(out + pix * out_ld + c * 4) = acc;
}

// out[r, n] = x[r, :] . W[n, :] + b[n] (+ addend[r % add_mod, n]);  one warp per row, N <= 16, K a multiple of 32
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ x, int ldx, long long rows, int K, const float* __restrict__ Wt,
                    const float* __restrict__ b, int N, float* __restrict__ out, int ldo,
                    const float* __restrict__ addend, int add_mod) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  float acc[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) acc[n] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float xv = xr[k];
#pragma unroll
    for (int n = 0; n < 16; ++n)
      if (n < N) acc[n] = fmaf(xv, Wt[static_cast<long long>(n) * K + k], acc[n]);
  }
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    if (n < N) {                       // warp-uniform
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    }
  }
  if (lane < N) {
    float v = 0.f;
#pragma unroll
    for (int n = 0; n < 16; ++n)
      if (n == lane) v = acc[n];
    v += b ? b[lane] : 0.f;
    if (addend) v += addend[(row % add_mod) * N + lane];
    out[row * ldo + lane] = v;
  }
}

// The same for the two shapes the SA predictor uses (K = 256; N = 12 class scores, N = 2 keypoint logits) with the
// weights held in registers and a warp walking `rows_per_warp` rows: the generic kernel re-reads N x 1 KB of weights
// from L1 for every row (12 KB per 1 KB row at N = 12).  Same summation order per row as the generic kernel.
template <int N>
__global__ void __launch_bounds__(256)
small_linear256_kernel(const float* __restrict__ x, int ldx, long long rows, const float* __restrict__ Wt,
                       const float* __restrict__ b, float* __restrict__ out, int ldo, const float* __restrict__ addend,
                       int add_mod, int rows_per_warp) {
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  float w[N][8];
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int j = 0; j < 8; ++j) w[n][j] = Wt[n * 256 + lane + 32 * j];
  const float bias = lane < N ? (b ? b[lane] : 0.f) : 0.f;
  const long long r0 = warp * rows_per_warp;
  for (long long row = r0; row < r0 + rows_per_warp && row < rows; ++row) {
    const float* xr = x + row * ldx;
    float xv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xv[j] = xr[lane + 32 * j];
    float mine = 0.f;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) a = fmaf(xv[j], w[n][j], a);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (n == lane) mine = a;
    }
    if (lane < N) {
      float v = mine + bias;
      if (addend) v += addend[(row % add_mod) * N + lane];
      out[row * ldo + lane] = v;
    }
  }
}

// out[r, j] = relu(W0[j, 0] ref[r, 0] + W0[j, 1] ref[r, 1] + b0[j]),  j < H
__global__ void query_pos_hidden_kernel(const float* __restrict__ ref, const float* __restrict__ W0,
                                        const float* __restrict__ b0, int Hd, long long total, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long r = i / Hd;
  const int j = static_cast<int>(i % Hd);
  const float v = fmaf(W0[2 * j + 1], ref[2 * r + 1], fmaf(W0[2 * j], ref[2 * r], b0[j]));
  out[i] = fmaxf(v, 0.f);
}

__device__ __forceinline__ float inverse_sigmoid_dev(float x) {   // SA/src/zoo/rtdetr/utils.py:10-12, eps 1e-5
  x = fminf(fmaxf(x, 0.f), 1.f);
  return logf(fmaxf(x, 1e-5f) / fmaxf(1.f - x, 1e-5f));
}

struct SaHeadParams {
  const float* tgt;      // [rows, 256] decoder layer output
  const float* h2;       // [rows, 256] second hidden layer of dec_bbox_head[i]
  const float* g2;       // [rows, 256] second hidden layer of sigma_embed[i]
  const float* ref_in;   // [rows, 2] reference points of this layer (sigmoid domain)
  const float *Wc, *bc;  // dec_score_head[i]: [12, 256], [12]
  const float *Wb, *bb;  // dec_bbox_head[i].layers.2: [2, 256], [2]
  const float *Ws, *bs;  // sigma_embed[i].layers.2: [1, 256], [1]
  float* logits;         // [rows, 12]
  float* pts;            // [rows, 2]  sigmoid(delta + inverse_sigmoid(ref_in))
  float* logsig;         // [rows, 2]  (the scalar repeated, rtdetr_decoder.py:367)
  float* ref_out;        // [rows, 2]  = pts (next layer's reference points)
  long long rows;
};

// one warp per (image, query) row, E = 256
__global__ void __launch_bounds__(256) sa_head_kernel(const SaHeadParams p) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= p.rows) return;
  float t[8], h[8], g[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    t[j] = p.tgt[row * 256 + lane + 32 * j];
    h[j] = p.h2[row * 256 + lane + 32 * j];
    g[j] = p.g2[row * 256 + lane + 32 * j];
  }
  float acc[15];
#pragma unroll
  for (int n = 0; n < 15; ++n) {
    const float* w = n < 12 ? p.Wc + n * 256 : (n < 14 ? p.Wb + (n - 12) * 256 : p.Ws);
    const float* src = n < 12 ? t : (n < 14 ? h : g);
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) a = fmaf(src[j], w[lane + 32 * j], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    acc[n] = a;
  }
  if (lane < 12) {
    float v = 0.f;
#pragma unroll
    for (int n = 0; n < 12; ++n)
      if (n == lane) v = acc[n];
    p.logits[row * 12 + lane] = v + p.bc[lane];
  } else if (lane < 14) {
    const int c = lane - 12;
    const float d = (c == 0 ? acc[12] : acc[13]) + p.bb[c];
    const float v = 1.f / (1.f + expf(-(d + inverse_sigmoid_dev(p.ref_in[row * 2 + c]))));
    p.pts[row * 2 + c] = v;
    p.ref_out[row * 2 + c] = v;
  } else if (lane < 16) {
    p.logsig[row * 2 + (lane - 14)] = acc[14] + p.bs[0];
  }
}

__global__ void add_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out, long long n4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 x = a[i], y = b[i];
  out[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}

}  // namespace

std::string launch_sa_stem_im2col(const float* nchw, int NB, int H, int W, float* out, int round, cudaStream_t s) {
  if (H % 2 || W % 2) return "sa stem: input extent must be even";
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(NB) * Ho * Wo * 11;
  ProfScope ps(kFamElementwise, s);
  sa_stem_im2col_kernel<<<blocks_for(total, 256), 256, 0, s>>>(nchw, H, W, Ho, Wo, total, round, out);
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_avgpool2x2(const float* in, int NB, int H, int W, int C, float* out, int round, cudaStream_t s) {
  if (H % 2 || W % 2 || C % 4) return "avgpool2x2: extent must be even and C a multiple of 4";
  const long long total = static_cast<long long>(NB) * (H / 2) * (W / 2) * (C / 4);
  ProfScope ps(kFamElementwise, s);
  avgpool2x2_kernel<<<blocks_for(total, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(in), H, W, C / 4, total, round,
                                                           reinterpret_cast<float4*>(out));
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_act_rows(const float* in, int in_ld, const float* add, int add_ld, float* out, int out_ld,
                            long long rows, int C, int kind, int round, cudaStream_t s) {
  if (C % 4 || in_ld % 4 || out_ld % 4 || (add && add_ld % 4)) return "act_rows: strides must be multiples of 4";
  if (rows <= 0) return "";
  ProfScope ps(kFamElementwise, s);
  act_rows_kernel<<<blocks_for(rows * (C / 4), 256), 256, 0, s>>>(in, in_ld, add, add_ld, out, out_ld, rows, C / 4, kind, round);
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_upsample_nearest2x(const float* in, int in_ld, int NB, int H, int W, int C, float* out, int out_ld,
                                      cudaStream_t s) {
  if (C % 4 || out_ld % 4 || in_ld % 4) return "upsample_nearest2x: C and the row strides must be multiples of 4";
  const long long total = static_cast<long long>(NB) * 4 * H * W * (C / 4);
  ProfScope ps(kFamElementwise, s);
  upsample_nearest2x_kernel<<<blocks_for(total, 256), 256, 0, s>>>(in, in_ld, H, W, C / 4, total, out, out_ld);
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_bicubic_half(const float* in, int NB, int H, int W, int C, float* out, int out_ld, int round,
                                cudaStream_t s) {
  if (H % 2 || W % 2 || C % 4 || out_ld % 4) return "bicubic_half: extent must be even and C a multiple of 4";
  const long long total = static_cast<long long>(NB) * (H / 2) * (W / 2) * (C / 4);
  ProfScope ps(kFamElementwise, s);
  bicubic_half_kernel<<<blocks_for(total, 256), 256, 0, s>>>(in, H, W, C / 4, total, round, out, out_ld);
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_small_linear(const float* x, int ldx, long long rows, int K, const float* Wt, const float* b, int N,
                                float* out, int ldo, const float* addend, int add_mod, cudaStream_t s) {
  if (N <= 0 || N > 16) return "small_linear: 1..16 outputs";
  if (rows <= 0) return "";
  ProfScope ps(kFamHeads, s);
  if (K == 256 && (N == 12 || N == 2) && rows >= 4096) {
    const int rpw = 16;
    const long long warps = (rows + rpw - 1) / rpw;
    if (N == 12)
      small_linear256_kernel<12><<<blocks_for(warps * 32, 256), 256, 0, s>>>(x, ldx, rows, Wt, b, out, ldo, addend,
                                                                             add_mod > 0 ? add_mod : 1, rpw);
    else
      small_linear256_kernel<2><<<blocks_for(warps * 32, 256), 256, 0, s>>>(x, ldx, rows, Wt, b, out, ldo, addend,
                                                                            add_mod > 0 ? add_mod : 1, rpw);
    SPE_CUDA_TRY(cudaGetLastError());
    return "";
  }
  small_linear_kernel<<<blocks_for(rows * 32, 256), 256, 0, s>>>(x, ldx, rows, K, Wt, b, N, out, ldo, addend,
                                                                 add_mod > 0 ? add_mod : 1);
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_query_pos_hidden(const float* ref, const float* W0, const float* b0, int Hd, long long rows, float* out,
                                    cudaStream_t s) {
  if (rows <= 0) return "";
  ProfScope ps(kFamHeads, s);
  query_pos_hidden_kernel<<<blocks_for(rows * Hd, 256), 256, 0, s>>>(ref, W0, b0, Hd, rows * Hd, out);
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_sa_head(const float* tgt, const float* h2, const float* g2, const float* ref_in, long long rows,
                           const float* Wc, const float* bc, const float* Wb, const float* bb, const float* Ws,
                           const float* bs, float* logits, float* pts, float* logsig, float* ref_out, cudaStream_t s) {
  if (rows <= 0) return "";
  SaHeadParams p{tgt, h2, g2, ref_in, Wc, bc, Wb, bb, Ws, bs, logits, pts, logsig, ref_out, rows};
  ProfScope ps(kFamHeads, s);
  sa_head_kernel<<<blocks_for(rows * 32, 256), 256, 0, s>>>(p);
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_add(const float* a, const float* b, float* out, long long n, cudaStream_t s) {
  if (n % 4) return "add: element count must be a multiple of 4";
  if (n <= 0) return "";
  ProfScope ps(kFamElementwise, s);
  add_kernel<<<blocks_for(n / 4, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                    reinterpret_cast<float4*>(out), n / 4);
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

}  // namespace spe
