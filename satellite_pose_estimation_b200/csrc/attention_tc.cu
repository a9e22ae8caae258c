// Encoder self-attention on the 5th-generation tensor cores (tcgen05 + TMEM), head_dim = 32, fp32 storage / TF32.
//
//   O = softmax(Q K^T / sqrt(32)) V      per (image, head); the 784 x 784 weight matrix never leaves the SM
//   (reference: nn.MultiheadAttention inside TransformerEncoderLayer.forward_post, RV/models/transformer.py:154-158)
//
// One CTA = 128 query rows of one head (two CTAs co-reside per SM and overlap each other's phases), 192 threads:
//   warp 0      TMA producer: Q tile once, then K / V tiles of NK keys through a 2-stage mbarrier ring
//   warp 1      owns TMEM, one lane issues   S = Q K^T   (tcgen05.mma kind::tf32, A and B from shared memory)
//                                  and       O += P V     (A = P read straight from TMEM, B = V tile, MN-major,
//                                                          which for TF32 means the 32-byte-granule swizzle)
//   warps 2..9  softmax: tcgen05.ld gives thread t of a warp score row t of its TMEM lane quarter; two warps share a
//               quarter and split the chunk's columns (twice the warps to hide MUFU / TMEM latency), exchanging only
//               their partial row maxima through shared memory.  P is written back over S in TMEM (tcgen05.st); the
//               32-column O accumulator (16 columns per warp) is rescaled when the running maximum moves
//               (online softmax, fp32 statistics)
// Softmax reference: the running row maximum is taken from the FIRST chunk only (two passes over it: maximum, then
// exponentials) and kept for the whole row; later chunks run a single pass p = 2^((s - m_0) c).  Probabilities above 1
// are as exact in floating point as those below, so O / l is the same softmax -- without the per-chunk maximum pass,
// the partial-maximum exchange between the two warps of a row and the rescaling of O (a fifth of the softmax warps'
// instructions, two barrier round trips per chunk).  A row whose later scores exceed m_0 by more than ~350 logits would
// overflow: its row sum then fails the `l < 1e30` check at the end and the whole CTA repeats its tile with the
// classical online softmax (every chunk two passes, O rescaled when the maximum moves) on a second set of barriers.
// TMEM columns: two score / probability buffers [0, NK) and [NK, 2 NK) (QK^T of chunk j+1 is issued while the softmax
// of chunk j runs), output accumulator at [2 NK, 2 NK + 32).
#include "spe_internal.h"
#include "profile.h"
#include "spe_ptx.cuh"
#include <cuda_bf16.h>

#include <stdlib.h>

namespace spe {

namespace {

constexpr int kQRows = 128;
constexpr int kTmemCols = 256;
constexpr int kThreads = 320;        // TMA warp, MMA warp, 8 softmax warps

struct AttnTcParams {
  float* out;
  int ldo;
  int Lq, Lk;
  float scale_log2e;
  int exact_out;
  int debug;
  int out_bf16;     // bf16-storage models: Q / K / V arrive as fp32 (TF32 values), the output is written as bf16
  int q_rows_per_batch;   // Lq, or 0 when every image shares one [Lq, ldq] query block (AttnDesc::bsq == 0)
  int safe_softmax; // 1: classical online softmax from the start (SPE_ATTN_SAFE=1; the fast path falls back to it by itself)
};

template <int NK> struct AttnSmem {
  static constexpr int Q_BYTES = kQRows * 128;
  static constexpr int KV_BYTES = NK * 128;
  static constexpr int STAGE_BYTES = 2 * KV_BYTES;
  static constexpr int STAGES = 3;
  static constexpr int XCHG_BYTES = 2 * 2 * kQRows * 4 + 2 * kQRows * 4;   // partial maxima [2][2][128] + sums [2][128]
  static constexpr int BYTES = Q_BYTES + STAGES * STAGE_BYTES + 16 * 8 + 16 + XCHG_BYTES + 16 * 8 + 1024;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t rna_bits(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

template <int NK>
__global__ void __launch_bounds__(kThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  using SM = AttnSmem<NK>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + SM::Q_BYTES;
  uint64_t* bars0 = reinterpret_cast<uint64_t*>(sKV + SM::STAGES * SM::STAGE_BYTES);
  uint64_t* q_full = bars0 + 0;
  constexpr uint32_t kOCol = 2 * NK;
  static_assert(2 * NK + 32 <= kTmemCols, "two score buffers and the output accumulator must fit the TMEM allocation");
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars0 + 16);
  float* sm_max = reinterpret_cast<float*>(bars0 + 18);       // [2 chunk parities][2 halves][128 rows]
  float* sm_sum = sm_max + 2 * 2 * kQRows;                    // [2 halves][128 rows]
  uint64_t* bars1 = reinterpret_cast<uint64_t*>(sm_sum + 2 * kQRows);   // barrier set of the repeat pass
  volatile int* vote = reinterpret_cast<volatile int*>(bars1 + 14);     // set when a row sum overflowed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * kQRows;
  const int nchunks = (p.Lk + NK - 1) / NK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    vote[0] = 0;
    for (int a = 0; a < 2; ++a) {
      uint64_t* bb = a == 0 ? bars0 : bars1;
      for (int i = 0; i < 3; ++i) { mbar_init(&bb[1 + i], 1); mbar_init(&bb[4 + i], 1); }
      mbar_init(&bb[7], 1);
      mbar_init(&bb[8], 1);
      mbar_init(&bb[9], 8);
      mbar_init(&bb[10], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();     // everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_launch();

  // Producer and MMA warps run warp-uniform loops; one elected lane issues (spe_ptx.cuh: elect_one_sync) -- under
  // `if (lane == 0)` every one of the 18 MMAs of a chunk was wrapped in a ~90-cycle waterfall loop, 3.6 x their 450
  // tensor-pipe cycles.
  for (int attempt = 0; attempt < 2; ++attempt) {
  const bool fast = attempt == 0 && !p.safe_softmax;     // fixed softmax reference (see the header); attempt 1 = classical
  uint64_t* bars = attempt == 0 ? bars0 : bars1;
  uint64_t* kv_full = bars + 1;    // [3]
  uint64_t* kv_empty = bars + 4;   // [3]
  uint64_t* s_full = bars + 7;     // [2]
  uint64_t* p_full = bars + 9;
  uint64_t* pv_done = bars + 10;
  int bad = 0;
  if (warp == 0) {
    {
      if (attempt == 0 && elect_one_sync()) {
        mbar_expect_tx(q_full, SM::Q_BYTES);
        tma_load_2d(sQ, &tmQ, q_full, h * 32, b * p.q_rows_per_batch + q0);
      }
      __syncwarp();
      for (int j = 0; j < nchunks; ++j) {
        const int st = j % 3;
        const uint32_t u = static_cast<uint32_t>(j / 3);
        mbar_wait(&kv_empty[st], (u & 1u) ^ 1u, 11);
        if (elect_one_sync()) {
          mbar_expect_tx(&kv_full[st], SM::STAGE_BYTES);
          uint8_t* sk = sKV + st * SM::STAGE_BYTES;
          tma_load_2d(sk, &tmK, &kv_full[st], h * 32, b * p.Lk + j * NK);
          tma_load_2d(sk + SM::KV_BYTES, &tmV, &kv_full[st], h * 32, b * p.Lk + j * NK);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc_s = umma_idesc(2, kQRows, NK);
      constexpr uint32_t idesc_pv = umma_idesc(2, kQRows, 32) | (1u << 16);   // B (= V tile) is MN-major
      mbar_wait(q_full, 0, 12);
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      // S_i = Q K_i^T into score buffer i & 1.  K dimension = head_dim 32 = four K=8 steps inside one 128-byte
      // swizzle atom.  Tensor-pipe instructions retire in issue order, so this overwrites P_{i-2} only after
      // P V_{i-2} (issued earlier) has read it.
      auto issue_s = [&](int i) {
        const int st = i % 3;
        mbar_wait(&kv_full[st], static_cast<uint32_t>(i / 3) & 1u, 13);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t kdesc = umma_desc_sw128(smem_u32(sKV + st * SM::STAGE_BYTES));
          const uint32_t sbuf = tmem_base + static_cast<uint32_t>((i & 1) * NK);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss<true>(sbuf, qdesc + 2u * k, kdesc + 2u * k, idesc_s, k != 0 ? 1u : 0u);
          tc_commit(&s_full[i & 1]);
        }
        __syncwarp();
      };
      issue_s(0);
      for (int j = 0; j < nchunks; ++j) {
        if (j + 1 < nchunks) issue_s(j + 1);                    // scores of the next chunk while softmax_j runs
        mbar_wait(p_full, static_cast<uint32_t>(j) & 1u, 14);   // softmax wrote P_j (and rescaled O)
        tc_fence_after();
        const int st = j % 3;
        if (elect_one_sync()) {
          const uint64_t vdesc = umma_desc_mn_tf32(smem_u32(sKV + st * SM::STAGE_BYTES) + SM::KV_BYTES);
          const uint32_t pbuf = tmem_base + static_cast<uint32_t>((j & 1) * NK);
#pragma unroll
          for (int kk = 0; kk < NK / 8; ++kk)                     // 8 keys per MMA = two 4-row K atoms (1024 B) of V
            umma_ts_tf32(tmem_base + kOCol, pbuf + static_cast<uint32_t>(kk * 8), vdesc + 64u * kk, idesc_pv,
                         (j | kk) != 0 ? 1u : 0u);
          tc_commit(&kv_empty[st]);
          tc_commit(pv_done);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;                  // TMEM lane quarter
    const int half = (warp - 2) >> 2;        // which part of the chunk's columns / of O's columns this warp owns
    const int row = q * 32 + lane;
    // column split of a chunk: NK = 112 -> [0,64) | [64,112) ; NK = 64 -> [0,32) | [32,64)
    constexpr int C0 = (NK == 112) ? 64 : NK / 2;
    const int cbeg = half == 0 ? 0 : C0;
    const int n32 = half == 0 ? C0 / 32 : (NK - C0) / 32;          // full 32-column pieces
    const bool tail16 = (half == 1) && ((NK - C0) % 32 != 0);      // plus one 16-column piece (NK = 112, upper half)
    const uint32_t trow0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t ocol = 2 * NK + half * 16;                      // this warp's 16 columns of the O accumulator
    const float c = p.scale_log2e;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nchunks; ++j) {
      mbar_wait(&s_full[j & 1], static_cast<uint32_t>(j >> 1) & 1u, 15);
      tc_fence_after();
      const uint32_t trow = trow0 + static_cast<uint32_t>((j & 1) * NK + cbeg);   // this warp's score columns
      const int key0 = j * NK + cbeg;
      const bool full = j * NK + NK <= p.Lk;               // no ragged tail inside this chunk (the common case)
      float alpha = 1.0f;
      if (!fast || j == 0) {
      // ---- pass 1: partial row maximum (four independent chains; the ragged-tail predicate is hoisted out of the
      //      per-element loops -- ncu showed 11 issued instructions per score element with it inside)
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      for (int cc = 0; cc < n32; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32(trow + static_cast<uint32_t>(cc * 32), v);
        tmem_wait_ld();
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (key0 + cc * 32 + i < p.Lk) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
        }
      }
      if (tail16) {
        uint32_t v[16];
        tmem_ld_32x16(trow + static_cast<uint32_t>(n32 * 32), v);
        tmem_wait_ld();
        if (full) {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (key0 + n32 * 32 + i < p.Lk) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
        }
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // exchange with the warp that owns the other columns of the same rows
      float* xm = sm_max + (j & 1) * 2 * kQRows;
      xm[half * kQRows + row] = mx;
      named_bar_sync(1 + q, 64);
      mx = fmaxf(mx, xm[(half ^ 1) * kQRows + row]);
      const float m_new = fmaxf(m, mx);                    // finite: every chunk holds at least one valid key
      alpha = ex2((m - m_new) * c);                        // 0 on the first chunk (m = -inf)
      m = m_new;
      }
      const float mc = m * c;
      bool done = false;
      if constexpr (NK == 112) {
        if (fast && j > 0) {
          // ---- single pass, balanced split: each warp of a row owns 56 columns = pieces of 32 + 16 + 8, all three
          //      loads in flight before the first exponential (one TMEM round trip per chunk instead of three)
          const uint32_t t56 = trow0 + static_cast<uint32_t>((j & 1) * NK + half * 56);
          uint32_t va[32], vb[16], vc[8];
          tmem_ld_32x32(t56, va);
          tmem_ld_32x16(t56 + 32u, vb);
          tmem_ld_32x8(t56 + 48u, vc);
          tmem_wait_ld();
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            va[i] = __float_as_uint(ex2(fmaf(__uint_as_float(va[i]), c, -mc))) & 0xffffe000u;
            s4[i & 3] += __uint_as_float(va[i]);
          }
          tmem_st_32x32(t56, va);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            vb[i] = __float_as_uint(ex2(fmaf(__uint_as_float(vb[i]), c, -mc))) & 0xffffe000u;
            s4[i & 3] += __uint_as_float(vb[i]);
          }
          tmem_st_32x16(t56 + 32u, vb);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            vc[i] = __float_as_uint(ex2(fmaf(__uint_as_float(vc[i]), c, -mc))) & 0xffffe000u;
            s4[i & 3] += __uint_as_float(vc[i]);
          }
          tmem_st_32x8(t56 + 48u, vc);
          l += (s4[0] + s4[1]) + (s4[2] + s4[3]);
          done = true;
        }
      }
      if (!done) {
      // ---- pass 2: probabilities, written back over the scores.  P is cut to TF32 with one LOP3 (the conversion
      //      instruction shares the MUFU pipe with ex2) and the row sum is taken over the SAME cut values the tensor
      //      core will multiply, so the normalisation stays consistent (no truncation bias in O / l).
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int cc = 0; cc < n32; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32(trow + static_cast<uint32_t>(cc * 32), v);
        tmem_wait_ld();
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = __float_as_uint(ex2(fmaf(__uint_as_float(v[i]), c, -mc))) & 0xffffe000u;
            sum4[i & 3] += __uint_as_float(v[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float pr = (key0 + cc * 32 + i < p.Lk) ? ex2(fmaf(__uint_as_float(v[i]), c, -mc)) : 0.f;
            v[i] = __float_as_uint(pr) & 0xffffe000u;
            sum4[i & 3] += __uint_as_float(v[i]);
          }
        }
        tmem_st_32x32(trow + static_cast<uint32_t>(cc * 32), v);
      }
      if (tail16) {
        uint32_t v[16];
        tmem_ld_32x16(trow + static_cast<uint32_t>(n32 * 32), v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float pr = (full || key0 + n32 * 32 + i < p.Lk) ? ex2(fmaf(__uint_as_float(v[i]), c, -mc)) : 0.f;
          v[i] = __float_as_uint(pr) & 0xffffe000u;
          sum4[i & 3] += __uint_as_float(v[i]);
        }
        tmem_st_32x16(trow + static_cast<uint32_t>(n32 * 32), v);
      }
      l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));   // partial sum over this warp's columns
      }
      // ---- P V_{j-1} must have landed before O may be touched; waiting for it on EVERY chunk also keeps this warp
      //      exactly one phase behind the pv_done barrier (a parity wait two phases late would alias and fall through)
      if (j > 0) {
        mbar_wait(pv_done, static_cast<uint32_t>(j - 1) & 1u, 16);
        tc_fence_after();
        if (!fast && __any_sync(0xffffffffu, alpha != 1.0f)) {   // rescale the running output when a row maximum moved
          uint32_t o[16];
          tmem_ld_32x16(trow0 + ocol, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x16(trow0 + ocol, o);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> global (each warp its 16 columns; l = sum of the two partial sums)
    sm_sum[half * kQRows + row] = l;
    named_bar_sync(1 + q, 64);
    l += sm_sum[(half ^ 1) * kQRows + row];
    mbar_wait(pv_done, static_cast<uint32_t>(nchunks - 1) & 1u, 17);
    tc_fence_after();
    bad = (fast && q0 + row < p.Lq && !(l < 1e30f)) ? 1 : 0;      // overflowed row (or NaN): repeat the tile classically
    if (bad) vote[0] = 1;
    named_bar_sync(9, 256);                                        // the eight softmax warps agree before anything is stored
    const bool redo = fast && vote[0] != 0;
    uint32_t o[16];
    tmem_ld_32x16(trow0 + ocol, o);
    tmem_wait_ld();
    if (redo) {
      // nothing is stored; every warp meets at the barrier below and runs the tile again
    } else if (q0 + row < p.Lq && p.out_bf16) {
      const float inv = 1.0f / l;
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) +
                          (static_cast<long long>(b) * p.Lq + q0 + row) * p.ldo + h * 32 + half * 16;
#pragma unroll
      for (int i = 0; i < 16; i += 8) {
        uint4 o8;
        __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&o8);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          hh[u] = __floats2bfloat162_rn(__uint_as_float(o[i + 2 * u]) * inv, __uint_as_float(o[i + 2 * u + 1]) * inv);
        *reinterpret_cast<uint4*>(op + i) = o8;
      }
    } else if (q0 + row < p.Lq) {
      const float inv = 1.0f / l;
      float* op = p.out + (static_cast<long long>(b) * p.Lq + q0 + row) * p.ldo + h * 32 + half * 16;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 r4 = make_float4(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv,
                                __uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
        if (!p.exact_out)
          r4 = make_float4(__uint_as_float(rna_bits(r4.x)), __uint_as_float(rna_bits(r4.y)),
                           __uint_as_float(rna_bits(r4.z)), __uint_as_float(rna_bits(r4.w)));
        *reinterpret_cast<float4*>(op + i) = r4;
      }
    }
  }
  (void)bad;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (!fast || vote[0] == 0) break;       // done, unless a row overflowed the fixed-reference softmax: run the tile again
  }

  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int NK>
std::string launch_tc(const AttnDesc& d, cudaStream_t s) {
  using SM = AttnSmem<NK>;
  static bool attr_set = false;
  auto kfn = attention_tc_kernel<NK>;
  if (!attr_set) {
    SPE_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES));
    attr_set = true;
  }
  CUtensorMap tmQ, tmK, tmV;
  const int hd = d.heads * 32;
  std::string e;
  e = encode_tmap_2d(&tmQ, kTF32, d.q, hd, (d.bsq == 0 ? 1ll : static_cast<long long>(d.B)) * d.Lq, static_cast<long long>(d.ldq) * 4, 32,
                     kQRows);
  if (!e.empty()) return e;
  e = encode_tmap_2d(&tmK, kTF32, d.k, hd, static_cast<long long>(d.B) * d.Lk, static_cast<long long>(d.ldk) * 4, 32, NK);
  if (!e.empty()) return e;
  e = encode_tmap_2d(&tmV, kTF32, d.v, hd, static_cast<long long>(d.B) * d.Lk, static_cast<long long>(d.ldv) * 4, 32, NK,
                     /*swizzle_atom32=*/true);
  if (!e.empty()) return e;
  AttnTcParams p;
  p.out = reinterpret_cast<float*>(d.out);
  p.ldo = d.ldo;
  p.Lq = d.Lq;
  p.Lk = d.Lk;
  p.q_rows_per_batch = d.bsq == 0 ? 0 : d.Lq;
  p.scale_log2e = d.scale * 1.4426950408889634f;
  p.exact_out = d.exact_out;
  p.out_bf16 = d.mixed;
  static const bool dbg = getenv("SPE_ATTN_DEBUG") != nullptr;
  p.debug = dbg ? 1 : 0;
  static const bool safe = getenv("SPE_ATTN_SAFE") != nullptr && atoi(getenv("SPE_ATTN_SAFE")) != 0;
  p.safe_softmax = safe ? 1 : 0;
  dim3 grid((d.Lq + kQRows - 1) / kQRows, d.heads, d.B);
  ProfScope ps(kFamAttention, s);
  SPE_CUDA_TRY(launch_pdl(kfn, grid, dim3(kThreads), SM::BYTES, s, tmQ, tmK, tmV, p));
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

}  // namespace

// fp32 storage only; Q/K/V/O must be batch-contiguous (batch stride = rows * row stride)
bool attention_tc_supported(Dtype dt, const AttnDesc& d) {
  if (dt != kTF32 && !d.mixed) return false;   // bf16 storage: only with fp32 Q / K / V (AttnDesc::mixed)
  // small problems (decoder self-attention, 40 x 40) stay on the register kernel; decoder cross-attention (40 queries
  // x 784 keys) is worth a 128-row tile even at 31 % row occupancy (28 us vs 45 us per layer at B = 64)
  if (d.Lq < 32 || d.Lk < 128) return false;
  if ((d.bsq != 0 && d.bsq != static_cast<long long>(d.Lq) * d.ldq) || d.bsk != static_cast<long long>(d.Lk) * d.ldk ||
      d.bsv != static_cast<long long>(d.Lk) * d.ldv || d.bso != static_cast<long long>(d.Lq) * d.ldo)
    return false;
  if (d.ldq % 4 || d.ldk % 4 || d.ldv % 4 || d.ldo % 4) return false;
  return true;
}

std::string launch_attention_tc(const AttnDesc& d, cudaStream_t s) {
  if (d.Lk % 112 == 0) return launch_tc<112>(d, s);   // 784 = 7 x 112: no ragged chunk at the 224^2 / stride-8 size
  return launch_tc<64>(d, s);
}

}  // namespace spe
