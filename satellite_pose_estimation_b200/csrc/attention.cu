// Multi-head attention core: softmax(Q K^T * scale) V per (image, head), head_dim = 32, never materialising the
// attention weights (the reference's nn.MultiheadAttention returns head-averaged weights that every caller on the
// hot path discards: RV/models/transformer.py:157-158, :227-235).
//
// Flash-style single pass with an online softmax.  One CTA = 64 query rows (4 warps x 16 rows) of one head; K/V are
// staged tile by tile in padded shared memory (bank-conflict free for the fragment loads below).  Both products run
// on the tensor cores via mma.sync.m16n8k8 TF32 with fp32 accumulation and fp32 softmax statistics.  The P operand
// of P.V is fed straight from the QK^T accumulator registers: accumulator columns (2t, 2t+1) are mapped onto
// A-fragment columns (t, t+4) and the V rows are fetched with the same permutation, so no shuffle is needed.
//
// Encoder self-attention (784 x 784), decoder self-attention (Q x Q) and decoder cross-attention (Q x 784) all go
// through this kernel; ragged tails are masked.
#include "spe_internal.h"
#include "profile.h"
#include "spe_ptx.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace spe {

namespace {

constexpr int HD = 32;        // head dim
constexpr int KLD = HD + 4;   // padded smem row (floats)
constexpr int QT = 64;        // query rows per CTA

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ float fast_exp2(float x) {  // MUFU.EX2: exp2(-inf) = 0, ~2 ulp
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
        "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

template <typename T> __device__ __forceinline__ void load8(const T* p, float (&o)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&o)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float2 f = __bfloat1622float2(h[u]);
    o[2 * u] = f.x;
    o[2 * u + 1] = f.y;
  }
}
template <typename T> __device__ __forceinline__ float load1(const T* p);
template <> __device__ __forceinline__ float load1<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float load1<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void store2(T* p, float a, float b, int exact);
template <> __device__ __forceinline__ void store2<float>(float* p, float a, float b, int exact) {
  // rounded to TF32 when the consumer is a plain kind::tf32 GEMM; kept exact for the 3xTF32 decoder GEMMs
  *reinterpret_cast<float2*>(p) = exact ? make_float2(a, b) : make_float2(to_tf32(a), to_tf32(b));
}
template <> __device__ __forceinline__ void store2<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, int) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// X3 (fp32 storage only): error-compensated products, S = Q_hi K_hi + Q_lo K_hi + Q_hi K_lo and likewise P V, with
// x_hi = rna_tf32(x), x_lo = rna_tf32(x - x_hi): fp32-grade scores whatever their magnitude.  Used by the decoder, whose
// learned query embeddings can drive the self-attention logits into the hundreds, where plain TF32 operands are off by
// tenths of a logit.
template <typename T, int KT, bool X3>
__global__ void __launch_bounds__(128)
attention_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ out,
                 int ldq, int ldk, int ldv, int ldo, long long bsq, long long bsk, long long bsv, long long bso,
                 int Lq, int Lk, float scale_log2e, int exact_out) {
  pdl_wait();
  pdl_launch();
  __shared__ __align__(16) float Ks[KT * KLD];
  __shared__ __align__(16) float Vs[KT * KLD];
  __shared__ __align__(16) float Kl[X3 ? KT * KLD : 4];   // low parts
  __shared__ __align__(16) float Vl[X3 ? KT * KLD : 4];

  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * QT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  const T* qb = q + b * bsq + h * HD;
  const T* kb = k + b * bsk + h * HD;
  const T* vb = v + b * bsv + h * HD;

  // Q fragments (pre-scaled by scale * log2(e) so the softmax uses exp2 directly)
  const int r0 = q0 + warp * 16 + g;
  const int r1 = r0 + 8;
  float qa[4][4];
  float ql[X3 ? 4 : 1][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int c = ks * 8 + t;
    float x[4];
    x[0] = r0 < Lq ? load1(qb + static_cast<long long>(r0) * ldq + c) * scale_log2e : 0.f;
    x[1] = r1 < Lq ? load1(qb + static_cast<long long>(r1) * ldq + c) * scale_log2e : 0.f;
    x[2] = r0 < Lq ? load1(qb + static_cast<long long>(r0) * ldq + c + 4) * scale_log2e : 0.f;
    x[3] = r1 < Lq ? load1(qb + static_cast<long long>(r1) * ldq + c + 4) * scale_log2e : 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      qa[ks][e] = to_tf32(x[e]);
      if constexpr (X3) ql[ks][e] = to_tf32(x[e] - qa[ks][e]);
    }
  }

  float o[4][4];
#pragma unroll
  for (int d = 0; d < 4; ++d) { o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int kv0 = 0; kv0 < Lk; kv0 += KT) {
    __syncthreads();  // previous tile fully consumed
    for (int u = threadIdx.x; u < KT * 4; u += 128) {
      const int row = u >> 2, c8 = (u & 3) * 8;
      float kk[8], vv[8];
      if (kv0 + row < Lk) {
        load8(kb + static_cast<long long>(kv0 + row) * ldk + c8, kk);
        load8(vb + static_cast<long long>(kv0 + row) * ldv + c8, vv);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) { kk[e] = 0.f; vv[e] = 0.f; }
      }
      float* kd = Ks + row * KLD + c8;
      float* vd = Vs + row * KLD + c8;
      float kh[8], vh[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { kh[e] = to_tf32(kk[e]); vh[e] = to_tf32(vv[e]); }
      *reinterpret_cast<float4*>(kd) = make_float4(kh[0], kh[1], kh[2], kh[3]);
      *reinterpret_cast<float4*>(kd + 4) = make_float4(kh[4], kh[5], kh[6], kh[7]);
      *reinterpret_cast<float4*>(vd) = make_float4(vh[0], vh[1], vh[2], vh[3]);
      *reinterpret_cast<float4*>(vd + 4) = make_float4(vh[4], vh[5], vh[6], vh[7]);
      if constexpr (X3) {
        float* kld = Kl + row * KLD + c8;
        float* vld = Vl + row * KLD + c8;
        *reinterpret_cast<float4*>(kld) = make_float4(to_tf32(kk[0] - kh[0]), to_tf32(kk[1] - kh[1]), to_tf32(kk[2] - kh[2]), to_tf32(kk[3] - kh[3]));
        *reinterpret_cast<float4*>(kld + 4) = make_float4(to_tf32(kk[4] - kh[4]), to_tf32(kk[5] - kh[5]), to_tf32(kk[6] - kh[6]), to_tf32(kk[7] - kh[7]));
        *reinterpret_cast<float4*>(vld) = make_float4(to_tf32(vv[0] - vh[0]), to_tf32(vv[1] - vh[1]), to_tf32(vv[2] - vh[2]), to_tf32(vv[3] - vh[3]));
        *reinterpret_cast<float4*>(vld + 4) = make_float4(to_tf32(vv[4] - vh[4]), to_tf32(vv[5] - vh[5]), to_tf32(vv[6] - vh[6]), to_tf32(vv[7] - vh[7]));
      }
    }
    __syncthreads();

    // S = Q K^T for this tile: KT/8 accumulator tiles of 16 x 8
    float s[KT / 8][4];
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      const float* kr = Ks + (j * 8 + g) * KLD + t;
      if constexpr (X3) {
        const float* kl = Kl + (j * 8 + g) * KLD + t;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {     // small terms first
          mma_tf32(s[j], ql[ks], kr[ks * 8], kr[ks * 8 + 4]);
          mma_tf32(s[j], qa[ks], kl[ks * 8], kl[ks * 8 + 4]);
        }
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_tf32(s[j], qa[ks], kr[ks * 8], kr[ks * 8 + 4]);
    }
    // mask the ragged tail and take the running row maxima
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
      const int kc = kv0 + j * 8 + 2 * t;
      if (kc >= Lk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (kc + 1 >= Lk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: every tile holds >= 1 valid key
    const float a0 = fast_exp2(m0 - mn0), a1 = fast_exp2(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
      s[j][0] = fast_exp2(s[j][0] - mn0); s[j][1] = fast_exp2(s[j][1] - mn0);
      s[j][2] = fast_exp2(s[j][2] - mn1); s[j][3] = fast_exp2(s[j][3] - mn1);
      ps0 += s[j][0] + s[j][1];
      ps1 += s[j][2] + s[j][3];
    }
    l0 = l0 * a0 + ps0;  // per-thread partial sums; reduced across the quad once at the end
    l1 = l1 * a1 + ps1;
#pragma unroll
    for (int d = 0; d < 4; ++d) { o[d][0] *= a0; o[d][1] *= a0; o[d][2] *= a1; o[d][3] *= a1; }

    // O += P V ; A-fragment column t <-> key 2t, column t+4 <-> key 2t+1 (same permutation on the V rows)
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
      float pa[4];
      pa[0] = to_tf32(s[j][0]); pa[1] = to_tf32(s[j][2]); pa[2] = to_tf32(s[j][1]); pa[3] = to_tf32(s[j][3]);
      const float* vr0 = Vs + (j * 8 + 2 * t) * KLD + g;
      const float* vr1 = vr0 + KLD;
      if constexpr (X3) {
        float pl[4];
        pl[0] = to_tf32(s[j][0] - pa[0]); pl[1] = to_tf32(s[j][2] - pa[1]);
        pl[2] = to_tf32(s[j][1] - pa[2]); pl[3] = to_tf32(s[j][3] - pa[3]);
        const float* vl0 = Vl + (j * 8 + 2 * t) * KLD + g;
        const float* vl1 = vl0 + KLD;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          mma_tf32(o[d], pl, vr0[d * 8], vr1[d * 8]);
          mma_tf32(o[d], pa, vl0[d * 8], vl1[d * 8]);
        }
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) mma_tf32(o[d], pa, vr0[d * 8], vr1[d * 8]);
    }
  }

  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  T* ob = out + b * bso + h * HD;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int c = d * 8 + 2 * t;
    if (r0 < Lq) store2(ob + static_cast<long long>(r0) * ldo + c, o[d][0] * i0, o[d][1] * i0, exact_out);
    if (r1 < Lq) store2(ob + static_cast<long long>(r1) * ldo + c, o[d][2] * i1, o[d][3] * i1, exact_out);
  }
}

template <typename T>
std::string launch_attn_t(const AttnDesc& d, cudaStream_t s) {
  const float sl2 = d.scale * 1.4426950408889634f;
  dim3 grid((d.Lq + QT - 1) / QT, d.heads, d.B);
  const T* q = reinterpret_cast<const T*>(d.q);
  const T* k = reinterpret_cast<const T*>(d.k);
  const T* v = reinterpret_cast<const T*>(d.v);
  T* o = reinterpret_cast<T*>(d.out);
  ProfScope ps(kFamAttention, s);
  if constexpr (sizeof(T) == 4) {
    if (d.x3) {
      // the decoder's attention: compensated products; half-size tiles keep the four operand
      // tiles within the 48 KB of static shared memory
      if (d.Lk % 56 == 0) {
        SPE_CUDA_TRY(launch_pdl(attention_kernel<T, 56, true>, grid, dim3(128), 0, s, q, k, v, o, d.ldq, d.ldk, d.ldv, d.ldo,
                                d.bsq, d.bsk, d.bsv, d.bso, d.Lq, d.Lk, sl2, d.exact_out));
      } else {
        SPE_CUDA_TRY(launch_pdl(attention_kernel<T, 64, true>, grid, dim3(128), 0, s, q, k, v, o, d.ldq, d.ldk, d.ldv, d.ldo,
                                d.bsq, d.bsk, d.bsv, d.bso, d.Lq, d.Lk, sl2, d.exact_out));
      }
      SPE_CUDA_TRY(cudaGetLastError());
      return "";
    }
  }
  if (d.Lk % 112 == 0) {
    SPE_CUDA_TRY(launch_pdl(attention_kernel<T, 112, false>, grid, dim3(128), 0, s, q, k, v, o, d.ldq, d.ldk, d.ldv, d.ldo, d.bsq,
                            d.bsk, d.bsv, d.bso, d.Lq, d.Lk, sl2, d.exact_out));
  } else {
    SPE_CUDA_TRY(launch_pdl(attention_kernel<T, 64, false>, grid, dim3(128), 0, s, q, k, v, o, d.ldq, d.ldk, d.ldv, d.ldo, d.bsq,
                            d.bsk, d.bsv, d.bso, d.Lq, d.Lk, sl2, d.exact_out));
  }
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

}  // namespace

std::string launch_attention(Dtype dt, const AttnDesc& d, cudaStream_t s) {
  if (d.B <= 0 || d.Lq <= 0 || d.Lk <= 0) return "attention: empty problem";
  const int vec = 8;
  if (d.ldk % vec || d.ldv % vec || d.ldq % 1) return "attention: K/V row strides must be multiples of 8 elements";
  static const bool no_tc = getenv("SPE_ATTN_LEGACY") != nullptr;
  if (!no_tc && !d.x3 && attention_tc_supported(dt, d)) return launch_attention_tc(d, s);
  if (d.mixed) return "attention: fp32 Q/K/V with bf16 output needs the tcgen05 kernel (shape not supported)";
  if (dt == kTF32) return launch_attn_t<float>(d, s);
  return launch_attn_t<__nv_bfloat16>(d, s);
}

}  // namespace spe
