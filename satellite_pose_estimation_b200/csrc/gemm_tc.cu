// Tensor-core GEMM / implicit-GEMM convolution for sm_100a.
//
//   out[m, n] = act( scale[n] * sum_k A[m, k] * Wt[n, k] + bias[n] + residual[m', n] )
//
// Replaces the cuDNN / cuBLAS library calls the reference dispatches for every convolution and linear layer of
// the keypoint-set predictor (reference: RV/models/backbone.py:133-149 conv stack, RV/models/transformer.py:154-239
// linear layers, RV/models/detr_speed.py:50-55 projections).
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: A and W tiles -> 128B-swizzled shared memory ring (mbarrier full/empty pairs)
//   warp 1      owns TMEM; one lane issues tcgen05.mma (M=128, N=BN, K=32 bytes) into a double-buffered accumulator
//   warps 2..5  epilogue: tcgen05.ld accumulator -> scale/bias/residual/ReLU -> global (overlaps next tile's MMAs)
//
// Storage dtype float  -> kind::tf32 (BK = 32 elements), storage dtype bf16 -> kind::f16 (BK = 64 elements).
// Convolution mode reads the NHWC activation through a 4-D tensor map: for filter tap (r, s) the A tile is the box
// {BK channels, W, hrows, 1 image} shifted by (s - pad, r - pad); TMA zero-fills outside the image, which is
// exactly the convolution's zero padding, so no im2col buffer is ever materialised.
#include "spe_internal.h"
#include <algorithm>
#include "profile.h"
#include "spe_ptx.cuh"

#include <mutex>
#include <stdlib.h>

namespace spe {

namespace {

constexpr int BM = 128;

template <typename T> struct GemmTraits;
template <> struct GemmTraits<float> {
  static constexpr int BK = 32;
  static constexpr bool kTf32 = true;
  static constexpr int kFmt = 2;
};
template <> struct GemmTraits<__nv_bfloat16> {
  static constexpr int BK = 64;
  static constexpr bool kTf32 = false;
  static constexpr int kFmt = 1;
};

// X3 = error-compensated "3xTF32": A = A_hi + A_lo (split in shared memory by dedicated warps), W = W_hi + W_lo
// (split once at weight load, stored as [N, 2K] = [W_hi | W_lo]); D = A_hi W_hi + A_lo W_hi + A_hi W_lo in one TMEM
// accumulator.  Used where TF32's 10-bit mantissa is not enough (decoder + heads).
// EPI8 = eight epilogue warps instead of four (two per TMEM lane quarter, interleaved over the 32-column chunks).
// Small-K GEMMs are bound by the epilogue's instruction latency and by how many residual loads it keeps in flight,
// not by the tensor pipe; they trade smem stages for the wider epilogue.
template <int BN, bool X3, bool EPI8> struct StageCfg {
  static_assert(!(X3 && EPI8), "the 3xTF32 variant already spends its extra warps on the operand split");
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = X3 ? 2 * (A_BYTES + B_BYTES) : (A_BYTES + B_BYTES);
  static constexpr int EPI_WARPS = EPI8 ? 8 : 4;
  static constexpr int STAGES = X3 ? (BN <= 64 ? 4 : 3)
                                   : (EPI8 ? (BN <= 64 ? 7 : (BN <= 128 ? 5 : 3)) : ((BN <= 64) ? 8 : (BN <= 128 ? 6 : 4)));
  static constexpr int THREADS = (X3 || EPI8) ? 320 : 192;
  static constexpr int TMEM_COLS = 2 * BN;  // power of two for BN in {64,128,256}
  static constexpr int STAGING_BYTES = EPI_WARPS * 4096;  // one 32-row x 128-byte transpose buffer per epilogue warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 2 * 2 * BN * 4 /*scale,bias x2*/ +
                                    (3 * STAGES + 4) * 8 /*barriers*/ + 16 /*tmem ptr*/ + 1024 /*align slack*/;
};

struct GemmKParams {
  int mode;
  int M, N;            // M = number of valid output rows in total
  int num_kb;          // k-blocks per tile
  int num_m_tiles, num_n_tiles;
  uint32_t a_bytes;    // bytes one A TMA box delivers
  // conv
  int HW, W, S, pad, hrows, tiles_per_img, kb_per_tap, H, cstride;   // H, W = OUTPUT extent
  int tiles_per_row;   // stem mode: 128-wide column tiles per output row
  // epilogue
  const float* scale;
  const float* bias;
  const void* residual;
  int res_ld, res_mod, res_f32, relu;
  void* out;
  int out_ld;
  int round_out;       // fp32 storage: round results to TF32 (consumer is a kind::tf32 MMA)
  int out_f32;         // bf16 storage: write this GEMM's output as fp32 (rounded to TF32), e.g. Q|K|V for the tcgen05
                       // attention kernel, which takes fp32 / TF32 operands
  int K;               // X3: column offset of W_lo inside the [N, 2K] weight matrix
  int remap_wp;        // tap-reuse 3x3 kernel: accumulator row m' = h * remap_wp + w of a (W + 2)-wide padded grid
  int sub_rows;        // ... image rows per 128-row accumulator
  int stem_im2col;     // mode 2 through an im2col map: tile = 128 consecutive output pixels
  uint32_t a_stage;    // ... bytes of one halo-tile stage
  int kb_split;        // mode 0 with a second operand source: k-blocks [kb_split, num_kb) come from tmA2 (else = num_kb)
  int a2_im2col;       // ... through an im2col map (1x1 / stride 2 sampling; uses HW, W, cstride) instead of a plain one
  int dbg;             // SPE_GEMM_DBG bit 0: skip the output stores (profiling experiments only)
  long long* tdbg;     // -DSPE_GEMM_TIMING + SPE_GEMM_TDBG=1: per-CTA wait-cycle counters of gemm_tc_kernel (bring-up only)
};

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// fp32 -> TF32, round to nearest, ties away from zero: what cvt.rna.tf32.f32 computes, on the integer ALUs.  The
// conversion instruction runs on the 16-lane XU pipe (8 issue cycles per warp, 32 of them per 32-column chunk of the
// epilogue); IEEE floats are sign-magnitude, so adding half a TF32 ulp to the bit pattern and clearing the low 13 bits
// rounds the magnitude for either sign (carries into the exponent, up to infinity, like the instruction).  NaNs pass.
__device__ __forceinline__ float rna_tf32(float x) {
  const uint32_t r = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  return x != x ? x : __uint_as_float(r);     // one FSETP + a predicated LOP3
}
// the same for a value that cannot be NaN (fmaxf(x, 0) never is): add + mask only
__device__ __forceinline__ float rna_tf32_finite(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

#ifdef SPE_GEMM_TIMING
#define GEMM_CLOCK() clock64()
#else
#define GEMM_CLOCK() 0ll
#endif
__device__ __forceinline__ void mbar_wait_g(uint64_t* bar, uint32_t parity, int tag, long long& acc) {
#ifdef SPE_GEMM_TIMING
  const long long t0 = clock64();
  mbar_wait(bar, parity, tag);
  acc += clock64() - t0;
#else
  mbar_wait(bar, parity, tag);
#endif
}

// Last step of one 32-column chunk in the coalesced domain: ReLU, rounding, 16-byte stores.  RELU / ROUND / FULL are
// compile-time so the per-row loop carries no uniform branches or predicate reloads (they were a third of the
// epilogue's issue slots, and the epilogue's issue rate is what bounds the short-K GEMMs).
template <typename T, bool RELU, bool ROUND, bool FULL, int NIT, int G>
__device__ __forceinline__ void store_chunk(float (&f)[NIT][G], uint8_t* gp, const long long out_step, const int crow,
                                            const int rows_here, const int rpi, const bool out_f32 = false) {
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    if (RELU) {
#pragma unroll
      for (int u = 0; u < G; ++u) f[i][u] = fmaxf(f[i][u], 0.0f);
    }
    if (FULL || i * rpi < rows_here - crow) {
      if (sizeof(T) == 4) {
        // fp32 storage feeds kind::tf32 MMAs, which drop the low 13 mantissa bits: round to nearest here
        // so the next layer's products are exact and the error stays unbiased
        float4 o4 = make_float4(f[i][0], f[i][1], f[i][2], f[i][3]);
        if (ROUND && RELU)
          o4 = make_float4(rna_tf32_finite(o4.x), rna_tf32_finite(o4.y), rna_tf32_finite(o4.z), rna_tf32_finite(o4.w));
        else if (ROUND) o4 = make_float4(rna_tf32(o4.x), rna_tf32(o4.y), rna_tf32(o4.z), rna_tf32(o4.w));
        *reinterpret_cast<float4*>(gp) = o4;
      } else if (out_f32) {
        // bf16 storage, fp32 output (TF32-rounded): this lane's eight columns are 32 contiguous bytes
        *reinterpret_cast<float4*>(gp) = make_float4(rna_tf32(f[i][0]), rna_tf32(f[i][1]), rna_tf32(f[i][2]),
                                                     rna_tf32(f[i][3]));
        *reinterpret_cast<float4*>(gp + 16) = make_float4(rna_tf32(f[i][4 % G]), rna_tf32(f[i][5 % G]),
                                                          rna_tf32(f[i][6 % G]), rna_tf32(f[i][7 % G]));
      } else {
        uint4 o8;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o8);
#pragma unroll
        for (int u = 0; u < 4; ++u) h[u] = __floats2bfloat162_rn(f[i][(2 * u) % G], f[i][(2 * u + 1) % G]);
        *reinterpret_cast<uint4*>(gp) = o8;
      }
    }
    gp += out_step;
  }
}

// One 128 x BN accumulator tile of one epilogue warp: rows [q*32, q*32+32) of the tile (its TMEM lane quarter), column
// chunks half, half + CSTEP, ...  `taddr` = TMEM address of the warp's lane quarter in the accumulator buffer.
// ACT: the instantiation also knows SiLU (relu == 2) and GELU (relu == 3) -- only the 3xTF32 kernels carry it (the SA
// predictor's ConvNormLayer / AIFI activations), so the plain-TF32 / bf16 epilogues keep their register budget.
template <typename T, int BN, int CSTEP, bool REMAP = false, bool ACT = false>
__device__ __forceinline__ void epilogue_tile(const GemmKParams& p, uint8_t* stg, const int lane, const int q,
                                              const int half, const uint32_t taddr, const int valid_rows,
                                              const long long m_base, const int n0, const float* s_scale,
                                              const float* s_bias, long long* tc = nullptr) {
  // tcgen05.ld hands thread t accumulator row t, but a warp-wide access "32 rows x 16 bytes" touches 32 different
  // 128-byte lines.  So each 32x32 fp32 accumulator chunk is transposed ONCE through a per-warp XOR-swizzled
  // 32 x 128-byte staging tile (raw accumulators in, conflict-free both ways), and everything else -- BN
  // scale/bias, residual add, ReLU, rounding, store -- happens in the coalesced domain where a warp instruction
  // covers RPI whole rows x 16 bytes per lane: global memory only sees full contiguous row segments, the
  // residual needs no staging at all, and each thread keeps one fixed group of G columns (one scale/bias fetch).
  constexpr int G = 16 / static_cast<int>(sizeof(T));        // columns per lane: 4 (fp32) / 8 (bf16)
  constexpr int CPR = 32 / G;                                // lanes per row: 8 / 4
  constexpr int RPI = 32 / CPR;                              // rows per warp instruction: 4 / 8
  constexpr int NIT = 32 / RPI;                              // instructions per chunk: 8 / 4
  const int crow = lane / CPR, cseg = lane % CPR;
  const int rows_here = valid_rows - q * 32;                 // valid rows in this warp's 32-row slab (may be <= 0)
  const long long slab0 = m_base + q * 32;                   // first global row of the slab
  const bool resid = p.residual != nullptr && rows_here > 0;
  const int res_es = (sizeof(T) == 4 || p.res_f32) ? 4 : 2;  // residual element size (fp32 addends in bf16 mode)
  // per-thread row pointers advance by a constant stride; the batch-broadcast addend wraps with one compare
  const bool out_f32 = sizeof(T) == 2 && p.out_f32;             // bf16 storage, fp32 output
  const long long oes = out_f32 ? 4 : static_cast<long long>(sizeof(T));   // output element size
  const long long out_step = static_cast<long long>(RPI) * p.out_ld * oes;
  uint8_t* out0 = reinterpret_cast<uint8_t*>(p.out) + ((slab0 + crow) * p.out_ld) * oes + cseg * G * oes;
  int rr0 = 0;
  if (resid) rr0 = p.res_mod > 0 ? static_cast<int>(static_cast<unsigned>(slab0 + crow) % static_cast<unsigned>(p.res_mod))
                                 : 0;
  const uint8_t* res_base = reinterpret_cast<const uint8_t*>(p.residual);
  // which store_chunk instance this tile uses (uniform over the warp)
  const int variant = (p.relu == 1 ? 1 : 0) | ((sizeof(T) == 4 && p.round_out) ? 2 : 0) | (rows_here >= 32 ? 4 : 0);
  // this lane's staging addresses (byte offsets): its own row for the transposing write, its column group for reads
  const uint32_t stg_w = smem_u32(stg) + lane * 128;
  const int wsw = lane & 7;
  // coalesced-domain reads: row r = i * RPI + crow, 16-byte chunk k of it sits at (k ^ (r & 7)).  fp32: k = cseg and
  // r & 7 alternates between crow and crow + 4 with i; bf16: k = 2 * cseg + {0, 1} and r & 7 = crow.  Two base
  // addresses either way, the rest is an immediate offset.
  const uint8_t* stg_r0 = stg + crow * 128 + (sizeof(T) == 4 ? ((cseg ^ crow) * 16) : (((2 * cseg) ^ crow) * 16));
  const uint8_t* stg_r1 = stg + crow * 128 + (sizeof(T) == 4 ? ((cseg ^ (crow + 4)) * 16) : (((2 * cseg + 1) ^ crow) * 16));

  // residual values of this lane for one chunk: NIT rows x G columns (prefetched one chunk ahead).  The row pointer
  // advances by a constant stride and the batch-broadcast addend wraps with one compare: the epilogue warp's issue
  // rate is what bounds the short-K GEMMs, and a 64-bit multiply per row was 22 instructions per load.
  uint4 rx[NIT][2];
#pragma unroll
  for (int i = 0; i < NIT; ++i) rx[i][0] = rx[i][1] = make_uint4(0u, 0u, 0u, 0u);
  const long long res_step = static_cast<long long>(RPI) * p.res_ld * res_es;
  const long long res_wrap = static_cast<long long>(p.res_mod) * p.res_ld * res_es;
  const int wrap_at = p.res_mod > 0 ? p.res_mod : 0x7fffffff;
  const uint8_t* res0 = res_base + ((p.res_mod > 0 ? static_cast<long long>(rr0) : slab0 + crow) * p.res_ld + cseg * G) *
                                       static_cast<long long>(res_es);
  const int rows_left = rows_here - crow;                    // row i * RPI + crow exists iff i * RPI < rows_left
  auto fetch_residual = [&](int ncol_) {
    const uint8_t* gp = res0 + static_cast<long long>(ncol_) * res_es;
    const int lim = ncol_ < p.N ? rows_left : 0;
    int w = rr0;
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      if (i * RPI < lim) {
        rx[i][0] = *reinterpret_cast<const uint4*>(gp);
        if (res_es * G > 16) rx[i][1] = *reinterpret_cast<const uint4*>(gp + 16);   // 8 fp32 addends (bf16 mode)
      }
      gp += res_step;
      w += RPI;
      if (w >= wrap_at) { w -= wrap_at; gp -= res_wrap; }
    }
  };
  if (resid) fetch_residual(n0 + half * 32);
  uint32_t v[32];
  tmem_ld_32x32(taddr + static_cast<uint32_t>(half * 32), v);
#pragma unroll 1
  for (int c = half; c < BN / 32; c += CSTEP) {
    const int ncol = n0 + c * 32;
    const bool col_ok = ncol < p.N;
    const long long k0 = GEMM_CLOCK();
    tmem_wait_ld();
    const long long k1 = GEMM_CLOCK();
    // own row -> staging (raw fp32 accumulators)
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_w + ((k ^ wsw) * 16)), "r"(v[4 * k]),
                   "r"(v[4 * k + 1]), "r"(v[4 * k + 2]), "r"(v[4 * k + 3])
                   : "memory");
    __syncwarp();
    const long long k2 = GEMM_CLOCK();
    // this lane's fixed column group: scale / bias
    float sc[G], bi[G];
#pragma unroll
    for (int u = 0; u < G; u += 4) {
      const float4 s4 = *reinterpret_cast<const float4*>(s_scale + c * 32 + cseg * G + u);
      const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c * 32 + cseg * G + u);
      sc[u] = s4.x; sc[u + 1] = s4.y; sc[u + 2] = s4.z; sc[u + 3] = s4.w;
      bi[u] = b4.x; bi[u + 1] = b4.y; bi[u + 2] = b4.z; bi[u + 3] = b4.w;
    }
    float f[NIT][G];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
#pragma unroll
      for (int u = 0; u < G; u += 4) {
        const uint8_t* sp = (sizeof(T) == 4 ? ((i & 1) ? stg_r1 : stg_r0) : (u ? stg_r1 : stg_r0)) + i * RPI * 128;
        const float4 a4 = *reinterpret_cast<const float4*>(sp);
        f[i][u] = fmaf(a4.x, sc[u], bi[u]);
        f[i][u + 1] = fmaf(a4.y, sc[u + 1], bi[u + 1]);
        f[i][u + 2] = fmaf(a4.z, sc[u + 2], bi[u + 2]);
        f[i][u + 3] = fmaf(a4.w, sc[u + 3], bi[u + 3]);
      }
    }
    __syncwarp();   // staging free for the next chunk
    // The next chunk's accumulators travel TMEM -> registers while this one is finished below.  Issued here and not
    // right behind the staging writes: tcgen05.ld overwrites the registers those st.shared still read, and waiting
    // for them to drain was the largest single stall of the loop (ncu source view, r01n).
    if (c + CSTEP < BN / 32) tmem_ld_32x32(taddr + static_cast<uint32_t>((c + CSTEP) * 32), v);
    const long long k3 = GEMM_CLOCK();
    if (resid && col_ok) {
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        if (res_es == 4) {
          f[i][0] += __uint_as_float(rx[i][0].x); f[i][1] += __uint_as_float(rx[i][0].y);
          f[i][2] += __uint_as_float(rx[i][0].z); f[i][3] += __uint_as_float(rx[i][0].w);
          if (G == 8) {
            f[i][G - 4] += __uint_as_float(rx[i][1].x); f[i][G - 3] += __uint_as_float(rx[i][1].y);
            f[i][G - 2] += __uint_as_float(rx[i][1].z); f[i][G - 1] += __uint_as_float(rx[i][1].w);
          }
        } else {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rx[i][0]);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 ff = __bfloat1622float2(h[u]);
            f[i][(2 * u) % G] += ff.x;
            f[i][(2 * u + 1) % G] += ff.y;
          }
        }
      }
    }
    const long long k4 = GEMM_CLOCK();
    if (resid && c + CSTEP < BN / 32) fetch_residual(ncol + 32 * CSTEP);   // next chunk, in flight during the stores
    const long long k5 = GEMM_CLOCK();
    if constexpr (ACT) {
      if (p.relu == 2) {                                   // SiLU
#pragma unroll
        for (int i = 0; i < NIT; ++i)
#pragma unroll
          for (int u = 0; u < G; ++u) f[i][u] = f[i][u] / (1.f + expf(-f[i][u]));
      } else if (p.relu == 3) {                            // GELU, erf form (nn.GELU default)
#pragma unroll
        for (int i = 0; i < NIT; ++i)
#pragma unroll
          for (int u = 0; u < G; ++u) f[i][u] = 0.5f * f[i][u] * (1.f + erff(f[i][u] * 0.70710678118654752f));
      }
    }
    if constexpr (REMAP) {
      // accumulator row r is pixel (h, w) = (r / Wp, r % Wp) of the padded grid; columns w >= W are the halo (their
      // values are meaningless) and valid_rows counts the image rows of this sub-tile that exist
      if (col_ok) {
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          const int r = q * 32 + i * RPI + crow;
          const int hl = r / p.remap_wp;
          const int w = r - hl * p.remap_wp;
          if (w < p.W && hl < valid_rows) {
            if (p.relu == 1) {
#pragma unroll
              for (int u = 0; u < G; ++u) f[i][u] = fmaxf(f[i][u], 0.0f);
            }
            uint8_t* gp = reinterpret_cast<uint8_t*>(p.out) +
                          ((m_base + static_cast<long long>(hl) * p.W + w) * p.out_ld + ncol) *
                              static_cast<long long>(sizeof(T)) + cseg * 16;
            if (sizeof(T) == 4) {
              float4 o4 = make_float4(f[i][0], f[i][1], f[i][2], f[i][3]);
              if (p.round_out) o4 = make_float4(rna_tf32(o4.x), rna_tf32(o4.y), rna_tf32(o4.z), rna_tf32(o4.w));
              *reinterpret_cast<float4*>(gp) = o4;
            } else {
              uint4 o8;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o8);
#pragma unroll
              for (int u = 0; u < 4; ++u) h[u] = __floats2bfloat162_rn(f[i][(2 * u) % G], f[i][(2 * u + 1) % G]);
              *reinterpret_cast<uint4*>(gp) = o8;
            }
          }
        }
      }
      continue;
    }
    if (col_ok && rows_here > 0 && !(p.dbg & 1)) {
      uint8_t* gp = out0 + static_cast<long long>(ncol) * oes;
      switch (variant) {
        case 0: store_chunk<T, false, false, false, NIT, G>(f, gp, out_step, crow, rows_here, RPI, out_f32); break;
        case 1: store_chunk<T, true, false, false, NIT, G>(f, gp, out_step, crow, rows_here, RPI, out_f32); break;
        case 2: store_chunk<T, false, true, false, NIT, G>(f, gp, out_step, crow, rows_here, RPI, out_f32); break;
        case 3: store_chunk<T, true, true, false, NIT, G>(f, gp, out_step, crow, rows_here, RPI, out_f32); break;
        case 4: store_chunk<T, false, false, true, NIT, G>(f, gp, out_step, crow, rows_here, RPI, out_f32); break;
        case 5: store_chunk<T, true, false, true, NIT, G>(f, gp, out_step, crow, rows_here, RPI, out_f32); break;
        case 6: store_chunk<T, false, true, true, NIT, G>(f, gp, out_step, crow, rows_here, RPI, out_f32); break;
        default: store_chunk<T, true, true, true, NIT, G>(f, gp, out_step, crow, rows_here, RPI, out_f32); break;
      }
    }
#ifdef SPE_GEMM_TIMING
    if (tc) {
      const long long k6 = GEMM_CLOCK();
      tc[0] += k1 - k0; tc[1] += k2 - k1; tc[2] += k3 - k2; tc[3] += k4 - k3; tc[4] += k5 - k4; tc[5] += k6 - k5;
    }
#else
    (void)k0; (void)k1; (void)k2; (void)k3; (void)k4; (void)k5; (void)tc;
#endif
  }
}

template <typename T, int BN, bool X3, bool EPI8>
__global__ void __launch_bounds__((StageCfg<BN, X3, EPI8>::THREADS), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const GemmKParams p) {
  using Tr = GemmTraits<T>;
  using Cfg = StageCfg<BN, X3, EPI8>;
  constexpr int EPI_WARPS = Cfg::EPI_WARPS;
  constexpr int CSTEP = EPI_WARPS / 4;   // chunk stride of one epilogue warp
  static_assert(!X3 || sizeof(T) == 4, "3xTF32 needs fp32 storage");
  constexpr int BK = Tr::BK;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment of the shared-space address (required by the 128B swizzle atom)
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint8_t* sm_staging = smem + STAGES * Cfg::STAGE_BYTES;                          // [4 warps][4096], 1024-aligned
  float* sm_scale = reinterpret_cast<float*>(sm_staging + Cfg::STAGING_BYTES);   // [2][BN]
  float* sm_bias = sm_scale + 2 * BN;                                            // [2][BN]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sm_bias + 2 * BN);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* split_bar = empty_bar + STAGES;   // X3 only: A tile has been split into hi / lo
  uint64_t* tfull_bar = split_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.kb_split < p.num_kb) tma_prefetch_desc(&tmA2);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
      mbar_init(&split_bar[i], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();     // everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (warp-uniform loops, one elected lane issues: see elect_one_sync in spe_ptx.cuh)
    {
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0;
      const long long t_begin = GEMM_CLOCK();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.num_n_tiles;
        const int n0 = (tile - m_tile * p.num_n_tiles) * BN;
        int img = 0, h0 = 0, x0 = 0;
        if (p.mode == 1) {
          img = m_tile / p.tiles_per_img;
          h0 = (m_tile - img * p.tiles_per_img) * p.hrows;
        } else if (p.mode == 3 || (p.mode == 0 && p.a2_im2col)) {   // im2col map: 128 consecutive output pixels
          const int m0 = m_tile * BM;
          img = m0 / p.HW;
          const int rem = m0 - img * p.HW;
          h0 = rem / p.W;
          x0 = rem - h0 * p.W;
        } else if (p.mode == 2 && p.stem_im2col) {
          const int m0 = m_tile * BM;
          img = m0 / p.HW;
          const int rem = m0 - img * p.HW;
          h0 = rem / p.W;
          x0 = rem - h0 * p.W;
        } else if (p.mode == 2) {
          img = m_tile / p.tiles_per_img;
          const int rem = m_tile - img * p.tiles_per_img;
          h0 = rem / p.tiles_per_row;                    // output row
          x0 = (rem - h0 * p.tiles_per_row) * BM;        // first output column of the tile
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait_g(&empty_bar[stage], phase ^ 1u, 1, w_empty);
          if (elect_one_sync()) {
            mbar_expect_tx(&full_bar[stage], p.a_bytes + (X3 ? 2 : 1) * Cfg::B_BYTES);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + (X3 ? 2 : 1) * Cfg::A_BYTES;
            if (p.mode == 0) {
              if (kb < p.kb_split) tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_tile * BM);
              else if (!p.a2_im2col) tma_load_2d(sa, &tmA2, &full_bar[stage], (kb - p.kb_split) * BK, m_tile * BM);
              else tma_load_im2col_4d(sa, &tmA2, &full_bar[stage], (kb - p.kb_split) * BK, x0 * p.cstride, h0 * p.cstride,
                                      img, 0, 0);
            } else if (p.mode == 2) {
              // k-block kb = filter row kb: 8 consecutive padded pixels x Cp channels per output pixel, windows of
              // neighbouring output pixels overlap (dim-1 stride = 2 pixels)
              if (p.stem_im2col)
                tma_load_im2col_4d(sa, &tmA, &full_bar[stage], 0, x0, 2 * h0, img, 0, static_cast<uint16_t>(kb));
              else
                tma_load_4d(sa, &tmA, &full_bar[stage], 0, x0, 2 * h0 + kb, img);
            } else {
              const int tap = kb / p.kb_per_tap;
              const int c0 = (kb - tap * p.kb_per_tap) * BK;
              const int r = tap / p.S;
              const int s = tap - r * p.S;
              if (p.mode == 3)
                tma_load_im2col_4d(sa, &tmA, &full_bar[stage], c0, x0 * p.cstride - p.pad, h0 * p.cstride - p.pad, img,
                                   static_cast<uint16_t>(s), static_cast<uint16_t>(r));
              else
                tma_load_4d(sa, &tmA, &full_bar[stage], c0, s - p.pad, h0 * p.cstride + r - p.pad, img);
            }
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
            if constexpr (X3) tma_load_2d(sb + Cfg::B_BYTES, &tmB, &full_bar[stage], p.K + kb * BK, n0);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (p.tdbg && lane == 0) { p.tdbg[blockIdx.x * 16 + 0] = GEMM_CLOCK() - t_begin; p.tdbg[blockIdx.x * 16 + 1] = w_empty; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform, one elected lane)
    {
      constexpr uint32_t idesc = umma_idesc(Tr::kFmt, BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long w_full = 0, w_tempty = 0;
      const long long t_begin = GEMM_CLOCK();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t use_par = static_cast<uint32_t>(it >> 1) & 1u;
        mbar_wait_g(&tempty_bar[buf], use_par ^ 1u, 2, w_tempty);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait_g(X3 ? &split_bar[stage] : &full_bar[stage], phase, 3, w_full);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint64_t adesc = umma_desc_sw128(sa);
            const uint64_t bdesc = umma_desc_sw128(sa + (X3 ? 2 : 1) * Cfg::A_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // advance 32 bytes along K inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
              umma_ss<Tr::kTf32>(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            if constexpr (X3) {
              const uint64_t alo = umma_desc_sw128(sa + Cfg::A_BYTES);
              const uint64_t blo = umma_desc_sw128(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_ss<true>(tmem_d, alo + 2u * k, bdesc + 2u * k, idesc, 1u);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_ss<true>(tmem_d, adesc + 2u * k, blo + 2u * k, idesc, 1u);
            }
            tc_commit(&empty_bar[stage]);                        // smem slot reusable once these MMAs retire
            if (kb == p.num_kb - 1) tc_commit(&tfull_bar[buf]);  // accumulator complete
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (p.tdbg && lane == 0) {
        p.tdbg[blockIdx.x * 16 + 2] = GEMM_CLOCK() - t_begin; p.tdbg[blockIdx.x * 16 + 3] = w_full;
        p.tdbg[blockIdx.x * 16 + 4] = w_tempty;
      }
    }
    __syncwarp();
  } else if (X3 && warp >= 6) {
    // ------------------------------------------------------------------ X3: split the A tile in place (4 warps)
    const int st = threadIdx.x - 192;  // 0..127
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase, 5);
        float4* a = reinterpret_cast<float4*>(smem + stage * Cfg::STAGE_BYTES);
        float4* lo = a + Cfg::A_BYTES / 16;
#pragma unroll
        for (int i = 0; i < Cfg::A_BYTES / 16 / 128; ++i) {
          const float4 v = a[st + i * 128];
          // (no NaN guard on the rounding: for a NaN input the remainder v - h below is NaN whatever h came out as,
          // and the product sum carries it)
          const float4 h = make_float4(rna_tf32_finite(v.x), rna_tf32_finite(v.y), rna_tf32_finite(v.z),
                                       rna_tf32_finite(v.w));
          a[st + i * 128] = h;                                                   // A_hi: exactly TF32
          lo[st + i * 128] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);  // A_lo: the remainder
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // row inside the 128-row tile
    const int et = threadIdx.x - 64;        // 0 .. 32 * EPI_WARPS - 1
    const int half = (warp - 2) >> 2;       // which interleaved set of column chunks this warp handles
    int it = 0;
    long long w_tfull = 0, t_epi = 0, t_pre = 0;
    long long tcs[6] = {0, 0, 0, 0, 0, 0};
    const long long t_begin = GEMM_CLOCK();
    constexpr int NJ = (BN + 32 * EPI_WARPS - 1) / (32 * EPI_WARPS);   // scale / bias columns per thread and tile
    float nsc[NJ], nbi[NJ];
    auto fetch_scale_bias = [&](int tile_) {
      const int n0_ = (tile_ % p.num_n_tiles) * BN;
#pragma unroll
      for (int u = 0; u < NJ; ++u) {
        const int j = et + u * 32 * EPI_WARPS;
        const bool ok = j < BN && (n0_ + j) < p.N;
        nsc[u] = (p.scale != nullptr && ok) ? p.scale[n0_ + j] : 1.0f;
        nbi[u] = (p.bias != nullptr && ok) ? p.bias[n0_ + j] : 0.0f;
      }
    };
    if (static_cast<int>(blockIdx.x) < num_tiles) fetch_scale_bias(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const long long c0 = GEMM_CLOCK();
      const int buf = it & 1;
      const uint32_t use_par = static_cast<uint32_t>(it >> 1) & 1u;
      const int m_tile = tile / p.num_n_tiles;
      const int n0 = (tile - m_tile * p.num_n_tiles) * BN;
      long long m_base;
      int valid_rows;
      if (p.mode == 0 || p.mode == 3 || (p.mode == 2 && p.stem_im2col)) {
        m_base = static_cast<long long>(m_tile) * BM;
        const long long rem = static_cast<long long>(p.M) - m_base;
        valid_rows = rem < BM ? static_cast<int>(rem) : BM;
      } else if (p.mode == 2) {
        const int img = m_tile / p.tiles_per_img;
        const int rem = m_tile - img * p.tiles_per_img;
        const int oy = rem / p.tiles_per_row;
        const int x0 = (rem - oy * p.tiles_per_row) * BM;
        m_base = static_cast<long long>(img) * p.HW + static_cast<long long>(oy) * p.W + x0;
        valid_rows = (p.W - x0) < BM ? (p.W - x0) : BM;
      } else {
        const int img = m_tile / p.tiles_per_img;
        const int h0 = (m_tile - img * p.tiles_per_img) * p.hrows;
        m_base = static_cast<long long>(img) * p.HW + static_cast<long long>(h0) * p.W;
        const int hr = (p.H - h0) < p.hrows ? (p.H - h0) : p.hrows;
        valid_rows = hr * p.W;
      }
      float* s_scale = sm_scale + buf * BN;
      float* s_bias = sm_bias + buf * BN;
#pragma unroll
      for (int u = 0; u < NJ; ++u) {
        const int j = et + u * 32 * EPI_WARPS;
        if (j < BN) { s_scale[j] = nsc[u]; s_bias[j] = nbi[u]; }
      }
      named_bar_sync(1, 32 * EPI_WARPS);
      // the next tile's scale / bias travel to registers while this tile is drained (their L2 round trip used to sit
      // in front of every tile of the warps that bound the short-K GEMMs)
      if (tile + static_cast<int>(gridDim.x) < num_tiles) fetch_scale_bias(tile + gridDim.x);
      const long long c1 = GEMM_CLOCK();
      mbar_wait_g(&tfull_bar[buf], use_par, 4, w_tfull);
      const long long c2 = GEMM_CLOCK();
      tc_fence_after();

      epilogue_tile<T, BN, CSTEP, false, X3>(p, sm_staging + (warp - 2) * 4096, lane, q, half,
                                  tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN),
                                  valid_rows, m_base, n0, s_scale, s_bias, tcs);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      t_pre += c1 - c0; t_epi += GEMM_CLOCK() - c2;
    }
    if (p.tdbg && warp == 2 && lane == 0) {
      p.tdbg[blockIdx.x * 16 + 5] = GEMM_CLOCK() - t_begin; p.tdbg[blockIdx.x * 16 + 6] = w_tfull;
      p.tdbg[blockIdx.x * 16 + 7] = t_epi; p.tdbg[blockIdx.x * 16 + 8] = t_pre; p.tdbg[blockIdx.x * 16 + 9] = it;
      for (int k = 0; k < 6; ++k) p.tdbg[blockIdx.x * 16 + 10 + k] = tcs[k];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one 256 x 256 tile per cluster of two CTAs.  Each CTA loads its own 128 rows of A
// and HALF of the 256-row weight tile; a single tcgen05.mma.cta_group::2 (M = 256, issued by the leader CTA) reads
// both CTAs' shared memory and writes 128 accumulator rows into each CTA's TMEM.  Per 128x256x32 MACs a CTA now
// pulls 32 KB through L2 instead of 48 KB -- fp32 operands make L2->SM bandwidth (~13 TB/s measured) the ceiling of
// the big-K layers, not the tensor pipe.
// ------------------------------------------------------------------------------------------------------------------
template <bool EPI8>
struct Cg2Cfg {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = (BN / 2) * 128;       // this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARPS = EPI8 ? 8 : 4;
  static constexpr int STAGES = EPI8 ? 5 : 6;   // the 8-warp epilogue needs 32 KB of staging
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int TMEM_COLS = 512;
  static constexpr int STAGING_BYTES = EPI_WARPS * 4096;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 2 * 2 * BN * 4 + (2 * STAGES + 4) * 8 + 16 + 1024;
};

template <typename T, bool EPI8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cg2Cfg<EPI8>::THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const GemmKParams p) {
  using Tr = GemmTraits<T>;
  using Cfg = Cg2Cfg<EPI8>;
  constexpr int CSTEP = EPI8 ? 2 : 1;
  constexpr int BK = Tr::BK;
  constexpr int BN = Cfg::BN;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);   // same offset in both CTAs of the pair
  uint8_t* sm_staging = smem + STAGES * Cfg::STAGE_BYTES;
  float* sm_scale = reinterpret_cast<float*>(sm_staging + Cfg::STAGING_BYTES);
  float* sm_bias = sm_scale + 2 * BN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sm_bias + 2 * BN);   // used in the leader CTA only
  uint64_t* empty_bar = full_bar + STAGES;                              // per CTA (multicast commit)
  uint64_t* tfull_bar = empty_bar + STAGES;                             // per CTA (multicast commit)
  uint64_t* tempty_bar = tfull_bar + 2;                                 // leader CTA only: both epilogues arrive
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_tiles = ((p.num_m_tiles + 1) >> 1) * p.num_n_tiles;     // 256-row pair tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * Cfg::EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits and the TMEM allocation are visible in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();     // everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs; warp-uniform)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int pm = tile / p.num_n_tiles;
        const int n0 = (tile - pm * p.num_n_tiles) * BN;
        const int m_tile = 2 * pm + static_cast<int>(rank);    // may be one past the end: TMA zero-fills, nothing stored
        int img = 0, h0 = 0, w0 = 0;
        if (p.mode == 1) {
          img = m_tile / p.tiles_per_img;
          h0 = (m_tile - img * p.tiles_per_img) * p.hrows;
        } else if (p.mode == 3) {
          // im2col map: the tile is 128 CONSECUTIVE output pixels (rows and images wrap inside the TMA unit), no padded
          // rows; a tile past the end re-reads tile 0 (nothing of it is stored)
          const int m0 = m_tile < p.num_m_tiles ? m_tile * BM : 0;
          img = m0 / p.HW;
          const int rem = m0 - img * p.HW;
          h0 = rem / p.W;
          w0 = rem - h0 * p.W;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u, 21);
          if (elect_one_sync()) {
            // both CTAs' bytes are accounted on the leader's barrier
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * (p.a_bytes + Cfg::B_BYTES));
            const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + Cfg::A_BYTES;
            if (p.mode == 0) {
              tma_load_2d_cg2(sa, &tmA, lead_bar, kb * BK, m_tile * BM);
            } else {
              const int tap = kb / p.kb_per_tap;
              const int c0 = (kb - tap * p.kb_per_tap) * BK;
              const int r = tap / p.S;
              const int s = tap - r * p.S;
              if (p.mode == 3)
                tma_load_im2col_4d_cg2(sa, &tmA, lead_bar, c0, w0 * p.cstride - p.pad, h0 * p.cstride - p.pad, img,
                                       static_cast<uint16_t>(s), static_cast<uint16_t>(r));
              else
                tma_load_4d_cg2(sa, &tmA, lead_bar, c0, s - p.pad, h0 * p.cstride + r - p.pad, img);
            }
            tma_load_2d_cg2(sb, &tmB, lead_bar, kb * BK, n0 + static_cast<int>(rank) * (BN / 2));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: leader CTA, one elected lane
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc(Tr::kFmt, 2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long w_full = 0, w_tempty = 0;
      const long long t_begin = GEMM_CLOCK();
      for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
        const int buf = it & 1;
        const uint32_t use_par = static_cast<uint32_t>(it >> 1) & 1u;
        mbar_wait_g(&tempty_bar[buf], use_par ^ 1u, 22, w_tempty);   // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait_g(&full_bar[stage], phase, 23, w_full);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint64_t adesc = umma_desc_sw128(sa);
            const uint64_t bdesc = umma_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss_cg2<Tr::kTf32>(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            tc_commit_cg2(&empty_bar[stage], 3);                        // frees the slot in both CTAs
            if (kb == p.num_kb - 1) tc_commit_cg2(&tfull_bar[buf], 3);  // accumulator complete in both CTAs
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (p.tdbg && lane == 0) {
        p.tdbg[blockIdx.x * 16 + 2] = GEMM_CLOCK() - t_begin; p.tdbg[blockIdx.x * 16 + 3] = w_full;
        p.tdbg[blockIdx.x * 16 + 4] = w_tempty;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (each CTA its own 128 rows)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;   // with 8 warps the two halves take alternate column chunks
    const int et = threadIdx.x - 64;
    int it = 0;
    long long w_tfull = 0, t_epi = 0, t_arr = 0;
    long long tcs[6] = {0, 0, 0, 0, 0, 0};
    const long long t_begin = GEMM_CLOCK();
    for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
      const int buf = it & 1;
      const uint32_t use_par = static_cast<uint32_t>(it >> 1) & 1u;
      const int pm = tile / p.num_n_tiles;
      const int n0 = (tile - pm * p.num_n_tiles) * BN;
      const int m_tile = 2 * pm + static_cast<int>(rank);
      long long m_base = 0;
      int valid_rows = 0;
      if (m_tile < p.num_m_tiles) {
        if (p.mode == 0 || p.mode == 3) {
          m_base = static_cast<long long>(m_tile) * BM;
          const long long rem = static_cast<long long>(p.M) - m_base;
          valid_rows = rem < BM ? static_cast<int>(rem) : BM;
        } else {
          const int img = m_tile / p.tiles_per_img;
          const int h0 = (m_tile - img * p.tiles_per_img) * p.hrows;
          m_base = static_cast<long long>(img) * p.HW + static_cast<long long>(h0) * p.W;
          const int hr = (p.H - h0) < p.hrows ? (p.H - h0) : p.hrows;
          valid_rows = hr * p.W;
        }
      }
      float* s_scale = sm_scale + buf * BN;
      float* s_bias = sm_bias + buf * BN;
      for (int j = et; j < BN; j += 32 * Cfg::EPI_WARPS) {
        const bool ok = (n0 + j) < p.N;
        s_scale[j] = (p.scale != nullptr && ok) ? p.scale[n0 + j] : 1.0f;
        s_bias[j] = (p.bias != nullptr && ok) ? p.bias[n0 + j] : 0.0f;
      }
      named_bar_sync(1, 32 * Cfg::EPI_WARPS);
      mbar_wait_g(&tfull_bar[buf], use_par, 24, w_tfull);
      const long long c2 = GEMM_CLOCK();
      tc_fence_after();
      epilogue_tile<T, BN, CSTEP>(p, sm_staging + (warp - 2) * 4096, lane, q, half,
                              tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN),
                              valid_rows, m_base, n0, s_scale, s_bias, tcs);
      const long long c3 = GEMM_CLOCK();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {   // leader's barrier; relaxed: only TMEM reads are handed over (SPE_GEMM_RELEASE_ARRIVE for A/B)
#ifdef SPE_GEMM_RELEASE_ARRIVE
        mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[buf]), 0));
#else
        mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tempty_bar[buf]), 0));
#endif
      }
      t_epi += c3 - c2; t_arr += GEMM_CLOCK() - c3;
    }
    if (p.tdbg && warp == 2 && lane == 0) {
      p.tdbg[blockIdx.x * 16 + 5] = GEMM_CLOCK() - t_begin; p.tdbg[blockIdx.x * 16 + 6] = w_tfull;
      p.tdbg[blockIdx.x * 16 + 7] = t_epi; p.tdbg[blockIdx.x * 16 + 8] = t_arr; p.tdbg[blockIdx.x * 16 + 9] = it;
      for (int kq = 0; kq < 6; ++kq) p.tdbg[blockIdx.x * 16 + 10 + kq] = tcs[kq];
    }
  }

  tc_fence_before();
  cluster_sync_all();   // no CTA of the pair may exit (or free TMEM) while its partner still uses it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
  }
}

template <typename T, bool EPI8>
std::string launch_cg2(const GemmKParams& kp_in, const CUtensorMap& tmA, const CUtensorMap& tmB, int num_sms,
                       cudaStream_t stream) {
  GemmKParams kp = kp_in;
  static const bool tdbg_on = getenv("SPE_GEMM_TDBG") != nullptr;
  static long long* tdbg_dev = nullptr;
  if (tdbg_on && tdbg_dev == nullptr) SPE_CUDA_TRY(cudaMalloc(&tdbg_dev, 256 * 16 * sizeof(long long)));
  if (tdbg_on) SPE_CUDA_TRY(cudaMemsetAsync(tdbg_dev, 0, 256 * 16 * sizeof(long long), stream));
  kp.tdbg = tdbg_on ? tdbg_dev : nullptr;
  static bool attr_set = false;
  using Cfg = Cg2Cfg<EPI8>;
  auto kfn = gemm_tc2_kernel<T, EPI8>;
  if (!attr_set) {
    SPE_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int pair_tiles = ((kp.num_m_tiles + 1) / 2) * kp.num_n_tiles;
  int pairs = num_sms / 2;
  if (pairs > pair_tiles) pairs = pair_tiles;
  {
    ProfScope ps(kFamGemm, stream);
    SPE_CUDA_TRY(launch_pdl(kfn, dim3(2 * pairs), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, kp));
  }
  SPE_CUDA_TRY(cudaGetLastError());
  if (tdbg_on) {   // only meaningful in a -DSPE_GEMM_TIMING build
    static long long host[256 * 16];
    SPE_CUDA_TRY(cudaStreamSynchronize(stream));
    SPE_CUDA_TRY(cudaMemcpy(host, tdbg_dev, sizeof(host), cudaMemcpyDeviceToHost));
    for (int cta : {0, 1, 146}) {
      const long long* h = host + cta * 16;
      fprintf(stderr, "[gemm2 dbg] epi8 %d cta %3d: mma.total=%lld mma.wait_full=%lld mma.wait_tempty=%lld epi.total=%lld "
              "epi.wait_tfull=%lld epi.work=%lld epi.arrive=%lld tiles=%lld | wait_ld=%lld sts=%lld lds=%lld res=%lld fetch=%lld "
              "store=%lld\n", static_cast<int>(EPI8), cta, h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11], h[12],
              h[13], h[14], h[15]);
    }
  }
  return "";
}

// ------------------------------------------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 convolution with tap reuse (Cout = BN <= 128).  The generic implicit GEMM above fetches the
// shifted input window once per filter tap: nine TMA loads of (almost) the same pixels per channel block, and with
// N <= 128 also one pass over the whole filter bank per 112 output pixels -- the 3x3 layers of layer1 / layer2 ran at
// 2.7-4.3x their ideal, bound by L2 -> SM traffic.  Here a CTA loads, per 32-channel block, ONE halo tile of
// (2 * hrows + 2) x (W + 2) pixels and feeds the tensor core nine row-shifted views of it: in the flattened padded grid
// output pixel m' = h * (W + 2) + w needs input pixel m' + r * (W + 2) + s for tap (r, s), i.e. the same 128-byte rows of
// shared memory starting r * (W + 2) + s rows further down (the descriptor's start address may be any 128-byte row:
// the swizzle is a function of absolute address bits).  The two halo columns per image row become two meaningless accumulator rows that are simply not stored.  Each
// filter k-block is used for two 128-row accumulators (2 * hrows image rows), which halves the filter traffic too.
// L2 -> SM bytes per 112 output pixels: 396 KB -> 118 KB (64 -> 64 channels at 56 x 56), 1080 KB -> 365 KB (128 -> 128
// channels at 28 x 28).
// ------------------------------------------------------------------------------------------------------------------
template <int BN> struct Conv3Cfg {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int A_STAGES = 2;
  static constexpr int THREADS = 192;
  static constexpr int TMEM_COLS = 4 * BN;        // two accumulators per tile, double-buffered
  static constexpr int STAGING_BYTES = 4 * 4096;
  static constexpr int FIXED_BYTES = STAGING_BYTES + 2 * 2 * BN * 4 + 64 * 8 + 16 + 1024;
};

template <typename T, int BN>
__global__ void __launch_bounds__(Conv3Cfg<BN>::THREADS, 1)
conv3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const GemmKParams p, const int b_stages) {
  using Tr = GemmTraits<T>;
  using Cfg = Conv3Cfg<BN>;
  constexpr int BK = Tr::BK;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sA = smem;                                               // [2][a_stage] halo tiles
  uint8_t* sB = sA + Cfg::A_STAGES * p.a_stage;                     // [b_stages][BN x 128 B]
  uint8_t* sm_staging = sB + b_stages * Cfg::B_BYTES;
  float* sm_scale = reinterpret_cast<float*>(sm_staging + Cfg::STAGING_BYTES);
  float* sm_bias = sm_scale + 2 * BN;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sm_bias + 2 * BN);
  uint64_t* a_empty = a_full + 2;
  uint64_t* tfull_bar = a_empty + 2;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* b_full = tempty_bar + 2;                                // [b_stages]
  uint64_t* b_empty = b_full + 24;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_empty + 24);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Wp = p.remap_wp;
  const int hrows = p.sub_rows;                // image rows per 128-row accumulator
  const int CB = p.kb_per_tap;                 // 32-channel blocks
  const int num_tiles = p.num_m_tiles;         // tiles of 2 * hrows image rows

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    for (int i = 0; i < b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();     // everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform, one lane issues)
    {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int img = tile / p.tiles_per_img;
        const int h0 = (tile - img * p.tiles_per_img) * (2 * hrows);
        for (int cb = 0; cb < CB; ++cb) {
          mbar_wait(&a_empty[as], aph ^ 1u, 41);
          if (elect_one_sync()) {
            mbar_expect_tx(&a_full[as], p.a_bytes);
            // halo tile: image rows h0-1 .. h0+2*hrows, columns -1 .. W (out-of-bounds = zero = the padding)
            tma_load_4d(sA + as * p.a_stage, &tmA, &a_full[as], cb * BK, -1, h0 - 1, img);
          }
          __syncwarp();
          if (++as == Cfg::A_STAGES) { as = 0; aph ^= 1u; }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[bs], bph ^ 1u, 42);
            if (elect_one_sync()) {
              mbar_expect_tx(&b_full[bs], Cfg::B_BYTES);
              tma_load_2d(sB + bs * Cfg::B_BYTES, &tmB, &b_full[bs], (tap * CB + cb) * BK, 0);
            }
            __syncwarp();
            if (++bs == b_stages) { bs = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform, one elected lane)
    {
      constexpr uint32_t idesc = umma_idesc(Tr::kFmt, BM, BN);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t use_par = static_cast<uint32_t>(it >> 1) & 1u;
        mbar_wait(&tempty_bar[buf], use_par ^ 1u, 43);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * 2 * BN);
        for (int cb = 0; cb < CB; ++cb) {
          mbar_wait(&a_full[as], aph, 44);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + as * p.a_stage);
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_full[bs], bph, 45);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + bs * Cfg::B_BYTES));
              const int r = tap / 3, sx = tap - 3 * r;
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const uint64_t adesc = umma_desc_sw128(a0 + static_cast<uint32_t>(((u * hrows + r) * Wp + sx) * 128));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_ss<Tr::kTf32>(tmem_d + static_cast<uint32_t>(u * BN), adesc + 2u * k, bdesc + 2u * k, idesc,
                                     (cb | tap | k) != 0 ? 1u : 0u);
              }
              tc_commit(&b_empty[bs]);
              if (tap == 8) {
                tc_commit(&a_empty[as]);
                if (cb == CB - 1) tc_commit(&tfull_bar[buf]);
              }
            }
            __syncwarp();
            if (++bs == b_stages) { bs = 0; bph ^= 1u; }
          }
          if (++as == Cfg::A_STAGES) { as = 0; aph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    const int et = threadIdx.x - 64;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use_par = static_cast<uint32_t>(it >> 1) & 1u;
      const int img = tile / p.tiles_per_img;
      const int h0 = (tile - img * p.tiles_per_img) * (2 * hrows);
      float* s_scale = sm_scale + buf * BN;
      float* s_bias = sm_bias + buf * BN;
      for (int j = et; j < BN; j += 128) {
        s_scale[j] = p.scale != nullptr ? p.scale[j] : 1.0f;
        s_bias[j] = p.bias != nullptr ? p.bias[j] : 0.0f;
      }
      named_bar_sync(1, 128);
      mbar_wait(&tfull_bar[buf], use_par, 46);
      tc_fence_after();
#pragma unroll 1
      for (int u = 0; u < 2; ++u) {
        const int hu = h0 + u * hrows;                                   // first image row of this accumulator
        int hv = p.H - hu;
        hv = hv < 0 ? 0 : (hv > hrows ? hrows : hv);
        const long long m_base = static_cast<long long>(img) * p.HW + static_cast<long long>(hu) * p.W;
        epilogue_tile<T, BN, 1, true>(p, sm_staging + (warp - 2) * 4096, lane, q, 0,
                                      tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                          static_cast<uint32_t>(buf * 2 * BN + u * BN),
                                      hv, m_base, 0, s_scale, s_bias);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn(std::string* err) {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  static std::string once_err;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
      once_err = std::string("cuTensorMapEncodeTiled not available: ") + cudaGetErrorString(e);
    } else {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  if (!fn && err) *err = once_err;
  return fn;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

// NHWC activation as an im2col tensor map for an R x S convolution with symmetric padding: the bounding box of base pixels
// is [-pad, W - 1 + pad - (S - 1)] x [-pad, H - 1 + pad - (R - 1)], walked with the convolution stride, i.e. exactly the
// output grid; one load delivers `pixels` consecutive output pixels x `channels` channels, shifted by the tap offset
// given at issue time
EncodeIm2colFn get_im2col_fn();

// stem: the overlapping-window view of the padded image ({8 pixels x 4 channels} per output pixel, W stride = 2 input
// pixels) as an im2col map: base rows 0, 2, 4, ... (traversal stride 2 in H), filter row = im2col offset in H
std::string encode_map_im2col_stem(CUtensorMap* m, Dtype dt, const void* base, int BKe, int Wo, int Hp, int Wp, int Cp,
                                   int NB, int pixels) {
  EncodeIm2colFn fn = get_im2col_fn();
  if (!fn) return "cuTensorMapEncodeIm2col not available";
  const long long es = static_cast<long long>(dtype_size(dt));
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(BKe), static_cast<cuuint64_t>(Wo), static_cast<cuuint64_t>(Hp),
                        static_cast<cuuint64_t>(NB)};
  cuuint64_t str[3] = {static_cast<cuuint64_t>(2 * Cp * es), static_cast<cuuint64_t>(static_cast<long long>(Wp) * Cp * es),
                       static_cast<cuuint64_t>(static_cast<long long>(Hp) * Wp * Cp * es)};
  int lower[2] = {0, 0};
  int upper[2] = {0, -6};
  cuuint32_t estr[4] = {1, 1, 2, 1};
  CUresult r = fn(m, dt == kTF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                  const_cast<void*>(base), dims, str, lower, upper, static_cast<cuuint32_t>(BKe),
                  static_cast<cuuint32_t>(pixels), estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return "cuTensorMapEncodeIm2col (stem) failed (" + std::to_string(static_cast<int>(r)) + ")";
  return "";
}

std::string encode_map_im2col(CUtensorMap* m, Dtype dt, const void* base, int C, int W, int H, int NB, int R, int S,
                              int pad, int stride, int channels, int pixels, int c_ld = 0) {
  EncodeIm2colFn fn = get_im2col_fn();
  if (!fn) return "cuTensorMapEncodeIm2col not available";
  const long long es = static_cast<long long>(dtype_size(dt));
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(NB)};
  const long long ld = c_ld > 0 ? c_ld : C;    // elements between consecutive pixels (a channel slice of a wider tensor)
  cuuint64_t str[3] = {static_cast<cuuint64_t>(ld * es), static_cast<cuuint64_t>(W * ld * es),
                       static_cast<cuuint64_t>(static_cast<long long>(H) * W * ld * es)};
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad - (S - 1), pad - (R - 1)};
  // traversal stride of the base pixel = convolution stride
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  CUresult r = fn(m, dt == kTF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                  const_cast<void*>(base), dims, str, lower, upper, static_cast<cuuint32_t>(channels),
                  static_cast<cuuint32_t>(pixels), estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return "cuTensorMapEncodeIm2col failed (" + std::to_string(static_cast<int>(r)) + ")";
  return "";
}

EncodeIm2colFn get_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  });
  return fn;
}

std::string encode_map(CUtensorMap* m, Dtype dt, int rank, const void* base, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes, const cuuint32_t* box, int spatial_stride = 1,
                       bool swizzle_atom32 = false) {
  std::string err;
  EncodeTiledFn fn = get_encode_fn(&err);
  if (!fn) return err;
  // traversal stride on the two spatial dims of an NHWC map = the convolution stride (TMA then delivers
  // ceil(box / stride) elements per dim)
  cuuint32_t estr[5] = {1, static_cast<cuuint32_t>(spatial_stride), static_cast<cuuint32_t>(spatial_stride), 1, 1};
  if (rank < 4) estr[1] = estr[2] = 1;
  CUresult r = fn(m, dt == kTF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf),
             "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] base %p",
             static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
             rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
    return buf;
  }
  return "";
}

template <typename T, int BN, bool X3, bool EPI8>
std::string launch_t(const GemmDesc& d, const GemmKParams& kp_in, const CUtensorMap& tmA, const CUtensorMap& tmB,
                     const CUtensorMap& tmA2, int num_sms, cudaStream_t stream) {
  GemmKParams kp = kp_in;
  static const bool tdbg_on = getenv("SPE_GEMM_TDBG") != nullptr;
  static long long* tdbg_dev = nullptr;
  if (tdbg_on && tdbg_dev == nullptr) SPE_CUDA_TRY(cudaMalloc(&tdbg_dev, 256 * 16 * sizeof(long long)));
  if (tdbg_on) SPE_CUDA_TRY(cudaMemsetAsync(tdbg_dev, 0, 256 * 16 * sizeof(long long), stream));
  kp.tdbg = tdbg_on ? tdbg_dev : nullptr;
  using Cfg = StageCfg<BN, X3, EPI8>;
  static bool attr_set = false;
  auto kfn = gemm_tc_kernel<T, BN, X3, EPI8>;
  if (!attr_set) {
    SPE_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = kp.num_m_tiles * kp.num_n_tiles;
  const int grid = tiles < num_sms ? tiles : num_sms;
  {
    ProfScope ps(kFamGemm, stream);
    SPE_CUDA_TRY(launch_pdl(kfn, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, tmA2, kp));
  }
  SPE_CUDA_TRY(cudaGetLastError());
  (void)d;
  if (tdbg_on) {   // only meaningful in a -DSPE_GEMM_TIMING build
    static long long host[256 * 16];
    SPE_CUDA_TRY(cudaStreamSynchronize(stream));
    SPE_CUDA_TRY(cudaMemcpy(host, tdbg_dev, sizeof(host), cudaMemcpyDeviceToHost));
    for (int cta : {0, 73, 147}) {
      const long long* h = host + cta * 16;
      fprintf(stderr, "[gemm dbg] BN %d epi8 %d cta %3d: prod.total=%lld prod.wait_empty=%lld mma.total=%lld mma.wait_full=%lld "
              "mma.wait_tempty=%lld epi.total=%lld epi.wait_tfull=%lld epi.work=%lld epi.pre=%lld tiles=%lld\n", BN,
              static_cast<int>(EPI8), cta, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
      fprintf(stderr, "[gemm dbg]   per-chunk phases (sum): wait_ld=%lld sts+ldtm=%lld lds+fma=%lld residual_add=%lld "
              "fetch_next=%lld store=%lld\n", h[10], h[11], h[12], h[13], h[14], h[15]);
    }
  }
  return "";
}

}  // namespace

std::string encode_tmap_2d(void* map, Dtype dt, const void* base, long long dim0, long long dim1,
                           long long stride1_bytes, int box0, int box1, bool swizzle_atom32) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(dim0), static_cast<cuuint64_t>(dim1)};
  cuuint64_t str[1] = {static_cast<cuuint64_t>(stride1_bytes)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1)};
  return encode_map(reinterpret_cast<CUtensorMap*>(map), dt, 2, base, dims, str, box, 1, swizzle_atom32);
}

// tap-reuse 3x3 / stride 1 / pad 1 convolution (conv3_tc_kernel); returns "skip" when the shape does not qualify
static std::string launch_conv3(Dtype dt, const GemmDesc& d, int num_sms, cudaStream_t stream) {
  static const int env_on = getenv("SPE_CONV3_REUSE") ? atoi(getenv("SPE_CONV3_REUSE")) : 1;
  const int es = static_cast<int>(dtype_size(dt));
  const int BK = 128 / es;
  const int cs = d.conv_stride > 0 ? d.conv_stride : 1;
  if (!env_on || d.mode != 1 || d.R != 3 || d.S != 3 || d.pad != 1 || cs != 1 || d.x3 || d.residual != nullptr ||
      (d.c_ld > 0 && d.c_ld != d.C))
    return "skip";
  if ((d.N != 64 && d.N != 128) || d.C % BK != 0 || d.W + 2 > 128 || d.H < 1) return "skip";
  const int Wp = d.W + 2;
  const int hrows = BM / Wp;
  const int dbl = 2 * hrows;
  if (dbl + 2 > 256) return "skip";
  if (d.H < dbl) return "skip";   // maps smaller than one double-row tile (8 x 8 of the SA predictor): generic path
  const int need_rows = std::max(Wp * (dbl + 2), (hrows + 2) * Wp + 2 + BM);
  const uint32_t a_stage = (static_cast<uint32_t>(need_rows) * 128u + 1023u) & ~1023u;
  const int b_bytes = d.N * 128;
  const int fixed = d.N == 64 ? Conv3Cfg<64>::FIXED_BYTES : Conv3Cfg<128>::FIXED_BYTES;
  int b_stages = (232448 - fixed - 2 * static_cast<int>(a_stage)) / b_bytes;
  if (b_stages > 12) b_stages = 12;
  if (b_stages < 3) return "skip";
  const int smem_bytes = fixed + 2 * static_cast<int>(a_stage) + b_stages * b_bytes;

  GemmKParams kp{};
  kp.mode = 1;
  kp.M = d.NB * d.H * d.W;
  kp.N = d.N;
  kp.tiles_per_img = (d.H + dbl - 1) / dbl;
  kp.num_m_tiles = kp.tiles_per_img * d.NB;
  kp.num_n_tiles = 1;
  kp.HW = d.H * d.W;
  kp.H = d.H;
  kp.W = d.W;
  kp.kb_per_tap = d.C / BK;
  kp.a_bytes = static_cast<uint32_t>(Wp * (dbl + 2) * 128);
  kp.remap_wp = Wp;
  kp.sub_rows = hrows;
  kp.a_stage = a_stage;
  kp.scale = d.scale; kp.bias = d.bias; kp.residual = nullptr; kp.relu = d.relu;
  kp.out = d.out; kp.out_ld = d.out_ld; kp.round_out = d.round_out;
  static const int dbg_flags = getenv("SPE_GEMM_DBG") ? atoi(getenv("SPE_GEMM_DBG")) : 0;
  kp.dbg = dbg_flags;

  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d.C), static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H),
                          static_cast<cuuint64_t>(d.NB)};
    cuuint64_t str[3] = {static_cast<cuuint64_t>(d.C) * es, static_cast<cuuint64_t>(d.W) * d.C * es,
                         static_cast<cuuint64_t>(d.H) * d.W * d.C * es};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(Wp), static_cast<cuuint32_t>(dbl + 2), 1};
    std::string err = encode_map(&tmA, dt, 4, d.A, dims, str, box);
    if (!err.empty()) return "conv3: " + err;
  }
  {
    const int K = 9 * d.C;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(d.N)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(K) * es};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(d.N)};
    std::string err = encode_map(&tmB, dt, 2, d.Wt, dims, str, box);
    if (!err.empty()) return "conv3: " + err;
  }
  const int grid = kp.num_m_tiles < num_sms ? kp.num_m_tiles : num_sms;
#define SPE_CONV3(TT, BNV)                                                                                        \
  do {                                                                                                            \
    auto kfn = conv3_tc_kernel<TT, BNV>;                                                                          \
    SPE_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));                 \
    ProfScope ps(kFamGemm, stream);                                                                               \
    SPE_CUDA_TRY(launch_pdl(kfn, dim3(grid), dim3(Conv3Cfg<BNV>::THREADS), smem_bytes, stream, tmA, tmB, kp,      \
                            b_stages));                                                                           \
  } while (0)
  if (dt == kTF32) {
    if (d.N == 64) SPE_CONV3(float, 64); else SPE_CONV3(float, 128);
  } else {
    if (d.N == 64) SPE_CONV3(__nv_bfloat16, 64); else SPE_CONV3(__nv_bfloat16, 128);
  }
#undef SPE_CONV3
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_gemm(Dtype dt, const GemmDesc& d, int num_sms, cudaStream_t stream) {
  const int es = static_cast<int>(dtype_size(dt));
  const int BK = 128 / es;
  if (d.N % 32 != 0) return "gemm: N must be a multiple of 32";
  if (d.out_ld % (16 / es) != 0 && !d.out_f32) return "gemm: out_ld must keep rows 16-byte aligned";
  if (d.x3 && dt != kTF32) return "gemm: 3xTF32 needs fp32 storage";
  if (d.x3 && d.mode == 2) return "gemm: 3xTF32 is not built for the windowed stem";
  if (d.relu < 0 || d.relu > 3 || (d.relu > 1 && !d.x3)) return "gemm: SiLU / GELU epilogues are only built for 3xTF32";
  if (d.mode < 0 || d.mode > 2) return "gemm: bad mode";
  if (d.A2 != nullptr && (d.mode != 0 || d.x3)) return "gemm: a second operand source needs a plain, uncompensated GEMM";
  if (d.mode == 1) {
    const std::string c3 = launch_conv3(dt, d, num_sms, stream);
    if (c3 != "skip") return c3;
  }
  const int cs_est = d.conv_stride > 1 ? d.conv_stride : 1;
  const long long m_tiles_est = d.mode == 0 ? (d.M + BM - 1) / BM
                                            : static_cast<long long>(d.NB) * (((d.H / cs_est) * (d.W / cs_est) + BM - 1) / BM);
  int BN = d.N <= 64 ? 64 : 128;
  // wide tiles halve the A re-reads (L2 -> SM bandwidth is what bounds fp32-operand GEMMs) when there are still
  // enough tiles to fill the machine
  if (!d.x3 && d.N % 256 == 0 && m_tiles_est * (d.N / 256) >= num_sms) BN = 256;
  // CTA pairs (256 x 256 tiles) for the deep-K layers, where operand traffic through L2 is the limiter
  static const int cg2_force = getenv("SPE_GEMM_CG2") ? atoi(getenv("SPE_GEMM_CG2")) : -1;
  static const int cg2_min_kb = getenv("SPE_GEMM_CG2_MINKB") ? atoi(getenv("SPE_GEMM_CG2_MINKB")) : 16;
  const int k_est = d.mode == 0 ? d.K : d.R * d.S * d.C;
  bool cg2 = !d.x3 && d.mode != 2 && d.N % 256 == 0 && k_est / BK >= cg2_min_kb && ((m_tiles_est + 1) / 2) * (d.N / 256) >= 48;
  if (cg2_force >= 0) cg2 = cg2 && cg2_force != 0;
  if (d.A2 != nullptr) cg2 = false;   // the two-source producer is built into gemm_tc_kernel only
  if (cg2) BN = 256;
  static const int bn_force = getenv("SPE_GEMM_BN") ? atoi(getenv("SPE_GEMM_BN")) : 0;   // experiments only
  if (!cg2 && (bn_force == 64 || bn_force == 128) && bn_force <= d.N) BN = bn_force;
  // few-row GEMMs (the decoder: M = 40 queries x batch): 128-wide tiles leave most SMs without a tile while every CTA
  // serialises the whole K loop -- 64-wide tiles double the CTAs that share it
  static const int small_bn = getenv("SPE_GEMM_SMALL_BN") ? atoi(getenv("SPE_GEMM_SMALL_BN")) : 1;
  if (small_bn && !cg2 && bn_force == 0 && BN == 128 && d.N % 64 == 0 &&
      m_tiles_est * ((d.N + 127) / 128) * 2 <= num_sms)
    BN = 64;

  GemmKParams kp{};
  kp.round_out = d.round_out;
  static const int dbg_flags = getenv("SPE_GEMM_DBG") ? atoi(getenv("SPE_GEMM_DBG")) : 0;
  kp.dbg = dbg_flags;
  kp.mode = d.mode;
  kp.N = d.N;
  kp.scale = d.scale;
  kp.bias = d.bias;
  kp.residual = d.residual;
  kp.res_ld = d.res_ld;
  kp.res_mod = d.res_mod;
  kp.res_f32 = d.res_f32;
  kp.relu = d.relu;
  kp.out = d.out;
  kp.out_ld = d.out_ld;
  kp.out_f32 = (dt != kTF32 && d.out_f32) ? 1 : 0;
  if (kp.out_f32 && d.mode != 0) return "gemm: fp32 output from bf16 storage is only built for plain matrices";
  kp.num_n_tiles = (d.N + BN - 1) / BN;

  CUtensorMap tmA, tmB, tmA2;
  std::string err;
  int K;
  kp.kb_split = 1 << 30;
  if (d.mode == 0) {
    K = d.K;
    if (K % BK != 0) return "gemm: K must be a multiple of the 128-byte k-block";
    if ((static_cast<long long>(d.lda) * es) % 16 != 0) return "gemm: lda must keep rows 16-byte aligned";
    kp.M = static_cast<int>(d.M);
    kp.num_m_tiles = static_cast<int>((d.M + BM - 1) / BM);
    kp.a_bytes = BM * 128;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(d.M)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(d.lda) * es};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), BM};
    err = encode_map(&tmA, dt, 2, d.A, dims, str, box);
    if (!err.empty()) return err;
    if (d.A2 != nullptr) {
      if (d.K2 <= 0 || d.K2 % BK != 0) return "gemm: K2 must be a multiple of the 128-byte k-block";
      kp.kb_split = K / BK;
      if (d.a2_stride == 1) {
        if ((static_cast<long long>(d.lda2) * es) % 16 != 0) return "gemm: lda2 must keep rows 16-byte aligned";
        cuuint64_t dims2[2] = {static_cast<cuuint64_t>(d.K2), static_cast<cuuint64_t>(d.M)};
        cuuint64_t str2[1] = {static_cast<cuuint64_t>(d.lda2) * es};
        err = encode_map(&tmA2, dt, 2, d.A2, dims2, str2, box);
      } else if (d.a2_stride == 2) {
        const int Ho = (d.a2_H - 1) / 2 + 1, Wo = (d.a2_W - 1) / 2 + 1;
        if (static_cast<long long>(d.a2_NB) * Ho * Wo != d.M) return "gemm: second source does not sample M pixels";
        kp.a2_im2col = 1;
        kp.HW = Ho * Wo; kp.H = Ho; kp.W = Wo; kp.cstride = 2;
        err = encode_map_im2col(&tmA2, dt, d.A2, d.K2, d.a2_W, d.a2_H, d.a2_NB, 1, 1, 0, 2, BK, BM);
      } else {
        return "gemm: second-source stride must be 1 or 2";
      }
      if (!err.empty()) return "gemm (second source): " + err;
      K += d.K2;
    }
  } else if (d.mode == 2) {
    const int Cp = 16 / es;                              // padded channels per pixel (16 bytes)
    if (d.C != Cp) return "stem: input must be the padded NHWC-Cp image";
    if (d.H % 2 || d.W % 2) return "stem: input extent must be even";
    const int Ho = d.H / 2, Wo = d.W / 2, Hp = d.H + 6, Wp = d.W + 6;
    K = 7 * BK;
    const int bw = Wo < BM ? Wo : BM;
    kp.tiles_per_row = (Wo + BM - 1) / BM;
    kp.tiles_per_img = Ho * kp.tiles_per_row;
    kp.num_m_tiles = kp.tiles_per_img * d.NB;
    kp.M = d.NB * Ho * Wo;
    kp.HW = Ho * Wo;
    kp.H = Ho;
    kp.W = Wo;
    kp.a_bytes = static_cast<uint32_t>(bw * 128);
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(BK), static_cast<cuuint64_t>(Wo), static_cast<cuuint64_t>(Hp),
                          static_cast<cuuint64_t>(d.NB)};
    cuuint64_t str[3] = {static_cast<cuuint64_t>(2 * Cp) * es, static_cast<cuuint64_t>(Wp) * Cp * es,
                         static_cast<cuuint64_t>(Hp) * Wp * Cp * es};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(bw), 1, 1};
    // output rows shorter than the tile (112 of 128): 128 consecutive output pixels per tile through an im2col view
    static const int stem_im2col_on = getenv("SPE_STEM_IM2COL") ? atoi(getenv("SPE_STEM_IM2COL")) : 1;
    bool stem_i2c = stem_im2col_on && Wo % BM != 0;
    if (stem_i2c) {
      err = encode_map_im2col_stem(&tmA, dt, d.A, BK, Wo, Hp, Wp, Cp, d.NB, BM);
      if (err.empty()) {
        kp.stem_im2col = 1;
        kp.num_m_tiles = (kp.M + BM - 1) / BM;
        kp.a_bytes = BM * 128;
      } else {
        stem_i2c = false;
      }
    }
    if (!stem_i2c) {
      err = encode_map(&tmA, dt, 4, d.A, dims, str, box);
      if (!err.empty()) return "stem: " + err;
    }
  } else {
    if (d.C % BK != 0) return "conv: C must be a multiple of the 128-byte k-block";
    if (d.c_ld > 0 && (d.c_ld < d.C || (static_cast<long long>(d.c_ld) * es) % 16 != 0)) return "conv: bad input pixel stride";
    const int cs = d.conv_stride > 0 ? d.conv_stride : 1;
    if (cs != 1 && cs != 2) return "conv: stride must be 1 or 2";
    const int Ho = (d.H + 2 * d.pad - d.R) / cs + 1, Wo = (d.W + 2 * d.pad - d.S) / cs + 1;
    if (Wo > BM) return "conv: output width must be <= 128";
    K = d.R * d.S * d.C;
    int hrows = BM / Wo;
    if (hrows > Ho) hrows = Ho;
    if (hrows * cs > 256 || Wo * cs > 256) return "conv: TMA box too large";
    // convolutions whose row blocks leave accumulator rows empty (28-wide maps: 4 x 28 = 112 of 128; 14-wide: 9 + 5 rows)
    // read their A tiles through an im2col tensor map instead: 128 consecutive output pixels per tile -- 12.5 % fewer
    // tiles at 28 x 28 (and, at B = 64, three waves of pair tiles instead of three and a bit), 23 % fewer at 14 x 14
    static const int im2col_on = getenv("SPE_CONV_IM2COL") ? atoi(getenv("SPE_CONV_IM2COL")) : 1;
    const bool im2col = im2col_on && (hrows * Wo) % BM != 0 && d.R <= 128 && d.S <= 128;
    if (im2col) {
      kp.mode = 3;
      kp.M = d.NB * Ho * Wo;
      kp.num_m_tiles = (kp.M + BM - 1) / BM;
      kp.HW = Ho * Wo; kp.H = Ho; kp.W = Wo; kp.S = d.S; kp.pad = d.pad; kp.cstride = cs;
      kp.kb_per_tap = d.C / BK;
      kp.a_bytes = BM * 128;
      err = encode_map_im2col(&tmA, dt, d.A, d.C, d.W, d.H, d.NB, d.R, d.S, d.pad, cs, BK, BM, d.c_ld);
      if (!err.empty()) return err;
    } else {
    kp.hrows = hrows;
    kp.tiles_per_img = (Ho + hrows - 1) / hrows;
    kp.num_m_tiles = kp.tiles_per_img * d.NB;
    kp.M = d.NB * Ho * Wo;
    kp.HW = Ho * Wo;
    kp.H = Ho;
    kp.W = Wo;
    kp.S = d.S;
    kp.pad = d.pad;
    kp.cstride = cs;
    kp.kb_per_tap = d.C / BK;
    kp.a_bytes = static_cast<uint32_t>(Wo * hrows * 128);
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d.C), static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H),
                          static_cast<cuuint64_t>(d.NB)};
    const cuuint64_t cld = static_cast<cuuint64_t>(d.c_ld > 0 ? d.c_ld : d.C);
    cuuint64_t str[3] = {cld * es, static_cast<cuuint64_t>(d.W) * cld * es, static_cast<cuuint64_t>(d.H) * d.W * cld * es};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(Wo * cs),
                         static_cast<cuuint32_t>(hrows * cs), 1};
    err = encode_map(&tmA, dt, 4, d.A, dims, str, box, cs);
    if (!err.empty()) return err;
    }
  }
  kp.num_kb = K / BK;
  kp.K = K;
  if (d.A2 == nullptr) { tmA2 = tmA; kp.kb_split = kp.num_kb; }
  {
    const int kw = d.x3 ? 2 * K : K;   // X3 weights: [N, 2K] = [W_hi | W_lo]
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(kw), static_cast<cuuint64_t>(d.N)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(kw) * es};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(cg2 ? BN / 2 : BN)};
    err = encode_map(&tmB, dt, 2, d.Wt, dims, str, box);
    if (!err.empty()) return err;
  }

  if (cg2) {
    const bool e8 = kp.num_kb <= 16;   // short K: the epilogue is the long pole, give it 8 warps
    if (dt == kTF32) return e8 ? launch_cg2<float, true>(kp, tmA, tmB, num_sms, stream)
                               : launch_cg2<float, false>(kp, tmA, tmB, num_sms, stream);
    return e8 ? launch_cg2<__nv_bfloat16, true>(kp, tmA, tmB, num_sms, stream)
              : launch_cg2<__nv_bfloat16, false>(kp, tmA, tmB, num_sms, stream);
  }
  // short K loops cannot hide a 4-warp epilogue: give those GEMMs the 8-warp epilogue (and fewer smem stages)
  static const int epi_force = getenv("SPE_GEMM_EPI8") ? atoi(getenv("SPE_GEMM_EPI8")) : -1;
  const bool epi8 = !d.x3 && (epi_force >= 0 ? epi_force != 0 : kp.num_kb <= 16);
#define SPE_LAUNCH(TT, BNV, X3V, E8V) return launch_t<TT, BNV, X3V, E8V>(d, kp, tmA, tmB, tmA2, num_sms, stream)
  if (dt == kTF32) {
    if (d.x3) {
      if (BN == 64) SPE_LAUNCH(float, 64, true, false);
      SPE_LAUNCH(float, 128, true, false);
    }
    if (epi8) {
      if (BN == 64) SPE_LAUNCH(float, 64, false, true);
      if (BN == 128) SPE_LAUNCH(float, 128, false, true);
      SPE_LAUNCH(float, 256, false, true);
    }
    if (BN == 64) SPE_LAUNCH(float, 64, false, false);
    if (BN == 128) SPE_LAUNCH(float, 128, false, false);
    SPE_LAUNCH(float, 256, false, false);
  } else {
    if (epi8) {
      if (BN == 64) SPE_LAUNCH(__nv_bfloat16, 64, false, true);
      if (BN == 128) SPE_LAUNCH(__nv_bfloat16, 128, false, true);
      SPE_LAUNCH(__nv_bfloat16, 256, false, true);
    }
    if (BN == 64) SPE_LAUNCH(__nv_bfloat16, 64, false, false);
    if (BN == 128) SPE_LAUNCH(__nv_bfloat16, 128, false, false);
    SPE_LAUNCH(__nv_bfloat16, 256, false, false);
  }
#undef SPE_LAUNCH
}

}  // namespace spe
