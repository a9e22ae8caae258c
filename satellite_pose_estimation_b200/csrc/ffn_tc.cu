// Fused transformer feed-forward block of the encoder (fp32 storage / TF32 tensor cores):
//
//     Y = LayerNorm( X + relu(X W1^T + b1) W2^T + b2 )          X, Y: [M, 256]   W1: [F, 256]   W2: [256, F]
//
// (reference: TransformerEncoderLayer.forward_post, RV/models/transformer.py:150-166 -- linear1, activation, linear2,
// residual add, norm2; dropout is identity in eval).  As two GEMMs the F = 2048-wide hidden activation costs a
// 411 MB store and a 411 MB load per encoder layer at B = 64 -- a quarter of the whole forward's HBM traffic -- and
// the K = 256 first GEMM is bound by exactly that store.  Here the hidden tile never leaves
// the SM: it is produced in tensor memory, rectified in place and consumed from tensor memory.
//
// One persistent CTA per SM, one 128-row tile of X at a time, hidden units in chunks of 128:
//   warp 0   TMA producer: the X tile (8 K-slabs of 128 x 32, resident in shared memory for the whole tile: A operand
//            of the first GEMM *and* the residual), then W1 / W2 k-blocks through a ring of three 32 KB entries
//   warp 1   one lane issues  S_j = X W1_j^T        (SS MMA, M128 N128, K = 256)   -> TMEM score buffer j & 1
//                             O  += H_j W2_j^T      (TS MMA, M128 N256, K = 128)   A = H_j read from TMEM
//            S_{j+1} is issued before H_j is waited for, so the tensor pipe works while the activation runs
//   warps 2-5  thread t owns row t (tcgen05.ld layout):  H_j = rna_tf32(relu(S_j + b1_j)) written back over S_j; after
//            the last chunk: v = O + b2 + X (X from the resident smem tile), two-pass LayerNorm statistics over the
//            thread's own row (no cross-thread reduction), normalise, round / split, store.
// A cta_group::2 variant (each CTA keeps half of every weight k-block) was built and measured 5 % slower (258 vs
// 246 us at M = 50176): the kernel is bound by the tensor pipe's sustained TF32 rate, not by its weight stream.
// TMEM: S0 [0,128)  S1 [128,256)  O [256,512).  Tensor-pipe instructions retire in issue order, so S_{j+2} overwrites
// H_j only after O += H_j W2_j^T has read it.
#include "spe_internal.h"
#include "spe_ptx.cuh"
#include "profile.h"

#include <cuda.h>

namespace spe {

namespace {

constexpr int kRows = 128;          // rows of X per tile
constexpr int kD = 256;             // model width
constexpr int kChunk = 128;         // hidden units per chunk
constexpr int kSlab = 16384;        // 128 rows x 128 bytes
constexpr int kEntry = 32768;       // ring entry: two W1 k-blocks (128 x 32 each) or one W2 k-block (256 x 32)
constexpr int kEntries = 3;
constexpr int kThreads = 192;
constexpr int kSmemBytes = 8 * kSlab + kEntries * kEntry + 256 + 1024;

struct FfnParams {
  long long M;
  int num_tiles, num_chunks;
  const float *b1, *b2, *gamma, *beta;
  float* out;
  int out_mode;     // 0: [M,256] rounded to TF32, 1: [M,256] exact fp32, 2: [M,768] = [hi | lo | hi] (3xTF32 operand)
};

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(kThreads, 1)
ffn_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sX = smem;                                   // [8 slabs][128 rows][128 B], SWIZZLE_128B
  uint8_t* sRing = smem + 8 * kSlab;                    // [3][32 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + kEntries * kEntry);
  uint64_t* x_full = bars;                              // X tile landed
  uint64_t* x_empty = bars + 1;                         // last S MMA done (commit) + 4 epilogue warps done with X
  uint64_t* full = bars + 2;                            // [3]
  uint64_t* empty = bars + 5;                           // [3]
  uint64_t* s_full = bars + 8;                          // [2] scores of a chunk complete
  uint64_t* h_ready = bars + 10;                        // [2] activation written back (4 warps)
  uint64_t* o_full = bars + 12;                         // all chunks accumulated
  uint64_t* o_empty = bars + 13;                        // epilogue drained O (4 warps)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    mbar_init(x_full, 1);
    mbar_init(x_empty, 5);
    for (int i = 0; i < kEntries; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&h_ready[i], 4); }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int NC = p.num_chunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int ent = 0;
      uint32_t ph = 0;
      auto w1_chunk = [&](int j) {          // four entries, each two 128 x 32 k-blocks of W1 rows [j*128, +128)
        for (int e = 0; e < 4; ++e) {
          mbar_wait(&empty[ent], ph ^ 1u, 31);
          mbar_expect_tx(&full[ent], kEntry);
          uint8_t* dst = sRing + ent * kEntry;
          tma_load_2d(dst, &tmW1, &full[ent], (2 * e) * 32, j * kChunk);
          tma_load_2d(dst + kSlab, &tmW1, &full[ent], (2 * e + 1) * 32, j * kChunk);
          if (++ent == kEntries) { ent = 0; ph ^= 1u; }
        }
      };
      auto w2_chunk = [&](int j) {          // four entries, each the 256 x 32 k-block of W2 columns [j*128 + e*32, +32)
        for (int e = 0; e < 4; ++e) {
          mbar_wait(&empty[ent], ph ^ 1u, 32);
          mbar_expect_tx(&full[ent], kEntry);
          tma_load_2d(sRing + ent * kEntry, &tmW2, &full[ent], j * kChunk + e * 32, 0);
          if (++ent == kEntries) { ent = 0; ph ^= 1u; }
        }
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(x_empty, (static_cast<uint32_t>(it) & 1u) ^ 1u, 33);
        mbar_expect_tx(x_full, 8 * kSlab);
        for (int s = 0; s < 8; ++s) tma_load_2d(sX + s * kSlab, &tmX, x_full, s * 32, tile * kRows);
        w1_chunk(0);
        for (int j = 0; j < NC; ++j) {
          if (j + 1 < NC) w1_chunk(j + 1);
          w2_chunk(j);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc(2, kRows, kChunk);
      constexpr uint32_t idesc_o = umma_idesc(2, kRows, kD);
      const uint32_t sx = smem_u32(sX);
      const uint32_t sring = smem_u32(sRing);
      const uint32_t tO = tmem_base + 256u;
      int ent = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(x_full, static_cast<uint32_t>(it) & 1u, 34);
        tc_fence_after();
        auto issue_s = [&](int j) {
          const uint32_t sbuf = tmem_base + static_cast<uint32_t>((j & 1) * kChunk);
          for (int e = 0; e < 4; ++e) {
            mbar_wait(&full[ent], ph, 35);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t adesc = umma_desc_sw128(sx + (2 * e + kk) * kSlab);
              const uint64_t bdesc = umma_desc_sw128(sring + ent * kEntry + kk * kSlab);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_ss<true>(sbuf, adesc + 2u * k, bdesc + 2u * k, idesc_s, (e | kk | k) != 0 ? 1u : 0u);
            }
            tc_commit(&empty[ent]);
            if (++ent == kEntries) { ent = 0; ph ^= 1u; }
          }
          tc_commit(&s_full[j & 1]);
          if (j == NC - 1) tc_commit(x_empty);     // the tensor pipe is done with the X tile
        };
        issue_s(0);
        for (int j = 0; j < NC; ++j) {
          if (j + 1 < NC) issue_s(j + 1);                           // next scores while the activation of j runs
          const uint32_t g = static_cast<uint32_t>(it) * static_cast<uint32_t>(NC) + static_cast<uint32_t>(j);
          mbar_wait(&h_ready[j & 1], (g >> 1) & 1u, 36);
          if (j == 0) mbar_wait(o_empty, (static_cast<uint32_t>(it) & 1u) ^ 1u, 37);   // previous tile's O drained
          tc_fence_after();
          const uint32_t hbuf = tmem_base + static_cast<uint32_t>((j & 1) * kChunk);
          for (int e = 0; e < 4; ++e) {
            mbar_wait(&full[ent], ph, 38);
            tc_fence_after();
            const uint64_t bdesc = umma_desc_sw128(sring + ent * kEntry);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ts_tf32(tO, hbuf + static_cast<uint32_t>(e * 32 + k * 8), bdesc + 2u * k, idesc_o,
                           (j | e | k) != 0 ? 1u : 0u);
            tc_commit(&empty[ent]);
            if (++ent == kEntries) { ent = 0; ph ^= 1u; }
          }
        }
        tc_commit(o_full);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ activation + LayerNorm epilogue
    const int q = warp & 3;                         // TMEM lane quarter
    const int row = q * 32 + lane;                  // row of the tile this thread owns
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t xrow = smem_u32(sX) + row * 128;
    const int sw = row & 7;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      for (int j = 0; j < NC; ++j) {
        const uint32_t g = static_cast<uint32_t>(it) * static_cast<uint32_t>(NC) + static_cast<uint32_t>(j);
        mbar_wait(&s_full[j & 1], (g >> 1) & 1u, 39);
        tc_fence_after();
        const uint32_t tb = trow + static_cast<uint32_t>((j & 1) * kChunk);
        const float* b1 = p.b1 + j * kChunk;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t v[32];
          tmem_ld_32x32(tb + static_cast<uint32_t>(cc * 32), v);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(b1 + cc * 32 + 4 * k));   // warp-uniform
            v[4 * k] = __float_as_uint(rna_tf32(fmaxf(__uint_as_float(v[4 * k]) + b4.x, 0.f)));
            v[4 * k + 1] = __float_as_uint(rna_tf32(fmaxf(__uint_as_float(v[4 * k + 1]) + b4.y, 0.f)));
            v[4 * k + 2] = __float_as_uint(rna_tf32(fmaxf(__uint_as_float(v[4 * k + 2]) + b4.z, 0.f)));
            v[4 * k + 3] = __float_as_uint(rna_tf32(fmaxf(__uint_as_float(v[4 * k + 3]) + b4.w, 0.f)));
          }
          tmem_st_32x32(tb + static_cast<uint32_t>(cc * 32), v);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_ready[j & 1]);
      }

      // ---- epilogue of the tile: v = O + b2 + X, LayerNorm over the thread's own row
      mbar_wait(o_full, static_cast<uint32_t>(it) & 1u, 40);
      tc_fence_after();
      const uint32_t to = trow + 256u;
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(to + static_cast<uint32_t>(c * 32), v);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float4 x4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(x4.x), "=f"(x4.y), "=f"(x4.z), "=f"(x4.w)
                       : "r"(xrow + c * kSlab + ((k ^ sw) * 16)));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + c * 32 + 4 * k));
          const float a0 = __uint_as_float(v[4 * k]) + b4.x + x4.x;
          const float a1 = __uint_as_float(v[4 * k + 1]) + b4.y + x4.y;
          const float a2 = __uint_as_float(v[4 * k + 2]) + b4.z + x4.z;
          const float a3 = __uint_as_float(v[4 * k + 3]) + b4.w + x4.w;
          sum += (a0 + a1) + (a2 + a3);
          v[4 * k] = __float_as_uint(a0); v[4 * k + 1] = __float_as_uint(a1);
          v[4 * k + 2] = __float_as_uint(a2); v[4 * k + 3] = __float_as_uint(a3);
        }
        tmem_st_32x32(to + static_cast<uint32_t>(c * 32), v);   // keep v: the X tile can go
      }
      tmem_wait_st();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_empty);          // residual read: the producer may load the next X tile
      const float mean = sum * (1.f / 256.f);
      float qs = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(to + static_cast<uint32_t>(c * 32), v);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float d = __uint_as_float(v[k]) - mean;
          qs = fmaf(d, d, qs);
        }
      }
      const float rstd = rsqrtf(qs * (1.f / 256.f) + 1e-5f);
      const long long grow = static_cast<long long>(tile) * kRows + row;
      const bool row_ok = grow < p.M;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(to + static_cast<uint32_t>(c * 32), v);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + c * 32 + 4 * k));
          const float4 e4 = __ldg(reinterpret_cast<const float4*>(p.beta + c * 32 + 4 * k));
          float4 y;
          y.x = (__uint_as_float(v[4 * k]) - mean) * rstd * g4.x + e4.x;
          y.y = (__uint_as_float(v[4 * k + 1]) - mean) * rstd * g4.y + e4.y;
          y.z = (__uint_as_float(v[4 * k + 2]) - mean) * rstd * g4.z + e4.z;
          y.w = (__uint_as_float(v[4 * k + 3]) - mean) * rstd * g4.w + e4.w;
          if (!row_ok) continue;
          const int col = c * 32 + 4 * k;
          if (p.out_mode == 2) {
            const float4 hi = make_float4(rna_tf32(y.x), rna_tf32(y.y), rna_tf32(y.z), rna_tf32(y.w));
            const float4 lo = make_float4(rna_tf32(y.x - hi.x), rna_tf32(y.y - hi.y), rna_tf32(y.z - hi.z),
                                          rna_tf32(y.w - hi.w));
            float* o = p.out + grow * 768 + col;
            *reinterpret_cast<float4*>(o) = hi;
            *reinterpret_cast<float4*>(o + 256) = lo;
            *reinterpret_cast<float4*>(o + 512) = hi;
          } else {
            if (p.out_mode == 0) y = make_float4(rna_tf32(y.x), rna_tf32(y.y), rna_tf32(y.z), rna_tf32(y.w));
            *reinterpret_cast<float4*>(p.out + grow * 256 + col) = y;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool ffn_fused_supported(Dtype dt, int d_model, int hidden) {
  static const int off = getenv("SPE_FFN_FUSED") ? (atoi(getenv("SPE_FFN_FUSED")) == 0) : 0;
  return !off && dt == kTF32 && d_model == kD && hidden > 0 && hidden % (2 * kChunk) == 0;
}

std::string launch_ffn_fused(const FfnDesc& d, int num_sms, cudaStream_t stream) {
  if (d.M <= 0) return "";
  if (d.hidden <= 0 || d.hidden % (2 * kChunk) != 0) return "ffn: hidden width must be a multiple of 256";
  CUtensorMap tmX, tmW1, tmW2;
  std::string e = encode_tmap_2d(&tmX, kTF32, d.X, kD, d.M, static_cast<long long>(kD) * 4, 32, kRows);
  if (!e.empty()) return "ffn X map: " + e;
  e = encode_tmap_2d(&tmW1, kTF32, d.W1, kD, d.hidden, static_cast<long long>(kD) * 4, 32, kChunk);
  if (!e.empty()) return "ffn W1 map: " + e;
  e = encode_tmap_2d(&tmW2, kTF32, d.W2, d.hidden, kD, static_cast<long long>(d.hidden) * 4, 32, kD);
  if (!e.empty()) return "ffn W2 map: " + e;
  FfnParams p{};
  p.M = d.M;
  p.num_tiles = static_cast<int>((d.M + kRows - 1) / kRows);
  p.num_chunks = d.hidden / kChunk;
  p.b1 = d.b1; p.b2 = d.b2; p.gamma = d.gamma; p.beta = d.beta;
  p.out = reinterpret_cast<float*>(d.out);
  p.out_mode = d.out_mode;
  static bool attr_set = false;
  if (!attr_set) {
    SPE_CUDA_TRY(cudaFuncSetAttribute(ffn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  {
    ProfScope ps(kFamGemm, stream);
    ffn_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tmX, tmW1, tmW2, p);
  }
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

}  // namespace spe
