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
// A cluster of two CTAs (cta_group::2) owns 256 rows of X at a time; hidden units go in chunks of 128.  Each CTA keeps
// its own 128-row X tile (8 K-slabs of 128 x 32, resident in shared memory for the whole tile: A operand of the first
// GEMM *and* the residual) and its own TMEM (S0 | S1 | O for its rows), but only HALF of every weight block: 64 of the
// 128 W1 rows of a chunk, 128 of the 256 W2 rows.
//   warp 0   TMA producer (both CTAs): X tile, then W1 / W2 half blocks through a ring of five 16 KB entries (two W1
//            half-slabs of 64 x 32, or one W2 half-slab of 128 x 32: 512 tensor-pipe cycles of work each); all
//            transaction bytes are accounted on the leader CTA's barriers
//   warp 1   (leader CTA) issues  S_j = X W1_j^T    (SS MMA, M256 N128, K = 256)   -> TMEM score buffer j & 1
//                                 O  += H_j W2_j^T  (TS MMA, M256 N256, K = 128)   A = H_j read from TMEM
//            S_{j+1} is issued before H_j is waited for, so the tensor pipe works while the activation runs; commits are
//            multicast to both CTAs
//   warps 2-9  thread t owns row t (tcgen05.ld layout; two warps per lane quarter split a chunk's columns):
//            H_j = rna_tf32(relu(S_j + b1_j)) written back over S_j, reported to the leader's barrier; after the last
//            chunk the same eight warps run the epilogue, 128 of the 256 columns each: v = O + b2 + X (X from the
//            resident smem tile), two-pass LayerNorm statistics (the two warps of a lane quarter exchange their partial
//            sums through shared memory), normalise, round / split, and store through a per-warp 32 x 64-byte
//            XOR-swizzled staging tile so that every store instruction writes eight rows x 64 contiguous bytes.
// TMEM: S0 [0,128)  S1 [128,256)  O [256,512).  Tensor-pipe instructions retire in issue order, so S_{j+2} overwrites
// H_j only after O += H_j W2_j^T has read it.
//
// How it got here (B = 64: M = 50176, F = 2048; 210 GFLOP):
//   one CTA per tile, `if (lane == 0)` issue loops, direct row-per-thread stores   245 us  (ncu r01h: tensor pipe 48 %)
//   + warp-uniform issue loops with elect_one_sync (see spe_ptx.cuh)               224 us
//   + CTA pairs, half the weight stream per SM and twice the ring depth            203 us
//   + coalesced epilogue stores (SPE_FFN_TIMING counters: the 64 row-per-thread STG.128 of a tile -- 32 distinct lines
//     per instruction -- took 32 k cycles, half of the tile's tensor time, during which the activation warps of the
//     next tile's first chunks were not served)                                    170 us  (ncu r01i: tensor pipe 73 %)
//   + epilogue on all eight row warps (half the columns each)                      165 us
#include "spe_internal.h"
#include "spe_ptx.cuh"
#include "profile.h"

#include <cuda.h>
#include <cuda_bf16.h>

namespace spe {

namespace {

constexpr int kRows = 128;          // rows of X per CTA tile
constexpr int kD = 256;             // model width
constexpr int kChunk = 128;         // hidden units per chunk
constexpr int kSlab = 16384;        // 128 rows x 128 bytes
constexpr int kEntry = 16384;       // ring entry: two W1 half k-blocks (64 x 32 each) or one W2 half k-block (128 x 32)
constexpr int kEntries = 5;
constexpr int kStage = 2048;        // per row warp: 32 rows x 64 bytes (16 output columns at a time)
constexpr int kThreads = 320;       // TMA warp, MMA warp, 8 row warps (two per TMEM lane quarter)
constexpr int kXchg = 2 * kRows * 4;   // partial LayerNorm sums exchanged between the two warps of a lane quarter
constexpr int kSmemBytes = 8 * kSlab + kEntries * kEntry + 8 * kStage + kXchg + 256 + 1024;
static_assert(kSmemBytes <= 232448, "shared memory budget");

struct FfnParams {
  long long M;
  int num_tiles, num_chunks;
  const float *b1, *b2, *gamma, *beta;
  float* out;
  long long* dbg;   // SPE_FFN_DBG: per-CTA wait-cycle counters (bring-up only), else nullptr
  int out_mode;     // 0: [M,256] rounded to TF32, 1: [M,256] exact fp32, 2: [M,768] = [hi | lo | hi] (3xTF32 operand)
};

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// mbar_wait that adds the cycles spent waiting to `acc` (bring-up instrumentation: build with -DSPE_FFN_TIMING and run
// with SPE_FFN_DBG=1)
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, int tag, long long& acc) {
#ifdef SPE_FFN_TIMING
  const long long t0 = clock64();
  mbar_wait(bar, parity, tag);
  acc += clock64() - t0;
#else
  mbar_wait(bar, parity, tag);
#endif
}
#ifdef SPE_FFN_TIMING
#define FFN_CLOCK() clock64()
#else
#define FFN_CLOCK() 0ll
#endif

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// One 32-row x 16-column block of the output through the warp's 2 KB staging tile: thread `lane` holds row `lane` (four
// float4 = 64 bytes); afterwards lane l stores chunk (l & 3) of rows 8 i + (l >> 2): eight rows x 64 contiguous bytes per
// instruction (whole 32-byte sectors).  The 16-byte chunks are XOR-swizzled with bits 1..2 of the row so that both the
// row-wise writes and the transposed reads are bank-conflict free.  `dst` points at (first row of the warp, first column
// of the block), `dst2` optionally at a second copy; `ld` is the row pitch in floats.
__device__ __forceinline__ void store_half_block(uint32_t stage, int lane, const float4* y, float* dst, float* dst2,
                                                 long long ld, int rows_ok) {
#pragma unroll
  for (int k = 0; k < 4; ++k) st_shared_v4(stage + lane * 64 + ((k ^ ((lane >> 1) & 3)) * 16), y[k]);
  __syncwarp();
  const int kk = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2);
    const float4 v = ld_shared_v4(stage + r * 64 + ((kk ^ ((r >> 1) & 3)) * 16));
    if (r < rows_ok) {
      *reinterpret_cast<float4*>(dst + r * ld + kk * 4) = v;
      if (dst2 != nullptr) *reinterpret_cast<float4*>(dst2 + r * ld + kk * 4) = v;
    }
  }
  __syncwarp();
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
ffn_tc2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);   // same offset in both CTAs of the pair
  uint8_t* sX = smem;                                   // [8 slabs][128 rows][128 B], SWIZZLE_128B: this CTA's rows
  uint8_t* sRing = smem + 8 * kSlab;                    // [5][16 KB]
  uint8_t* sStage = sRing + kEntries * kEntry;          // [8][2 KB] epilogue staging, one per row warp
  float* sXchg = reinterpret_cast<float*>(sStage + 8 * kStage);   // [2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sXchg) + kXchg);
  uint64_t* x_full = bars;                              // leader only: both CTAs' X tiles landed
  uint64_t* x_empty = bars + 1;                         // per CTA: last S MMA retired (multicast commit) + own 8 row warps
  uint64_t* full = bars + 2;                            // [5] leader only: both halves of the entry landed
  uint64_t* empty = bars + 8;                           // [5] per CTA (multicast commit)
  uint64_t* s_full = bars + 14;                         // [2] per CTA (multicast commit)
  uint64_t* h_ready = bars + 16;                        // [2] leader only: 8 activation warps of each CTA
  uint64_t* o_full = bars + 18;                         // per CTA (multicast commit)
  uint64_t* o_empty = bars + 19;                        // leader only: 8 row warps of each CTA
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_pt = (p.num_tiles + 1) >> 1;            // 256-row pair tiles (an odd tail tile is zero-filled, not stored)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    mbar_init(x_full, 1);
    mbar_init(x_empty, 9);
    for (int i = 0; i < kEntries; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&h_ready[i], 16); }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 16);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_ptr_smem, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits and the TMEM allocation are visible in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();     // everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_launch();
  const int NC = p.num_chunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs, each its own halves)
    // warp-uniform loops; one elected lane issues (spe_ptx.cuh: elect_one_sync)
    int ent = 0;
    uint32_t ph = 0;
    const int r = static_cast<int>(rank);
    long long w_empty = 0, w_xempty = 0;
    const long long t_begin = FFN_CLOCK();
    auto w1_chunk = [&](int j) {          // four entries, each two 64 x 32 half k-blocks of W1 rows [j*128 + r*64, +64)
      for (int e = 0; e < 4; ++e) {
        mbar_wait_t(&empty[ent], ph ^ 1u, 41, w_empty);
        if (elect_one_sync()) {
          if (rank == 0) mbar_expect_tx(&full[ent], 2u * kEntry);     // both CTAs' bytes land on the leader's barrier
          const uint32_t lead = mapa_u32(smem_u32(&full[ent]), 0);
          uint8_t* dst = sRing + ent * kEntry;
          tma_load_2d_cg2(dst, &tmW1, lead, (2 * e) * 32, j * kChunk + r * (kChunk / 2));
          tma_load_2d_cg2(dst + kEntry / 2, &tmW1, lead, (2 * e + 1) * 32, j * kChunk + r * (kChunk / 2));
        }
        __syncwarp();
        if (++ent == kEntries) { ent = 0; ph ^= 1u; }
      }
    };
    auto w2_chunk = [&](int j) {          // four entries: rows [r*128, +128) of the k-block of W2 columns [j*128 + e*32, +32)
      for (int e = 0; e < 4; ++e) {
        mbar_wait_t(&empty[ent], ph ^ 1u, 42, w_empty);
        if (elect_one_sync()) {
          if (rank == 0) mbar_expect_tx(&full[ent], 2u * kEntry);
          const uint32_t lead = mapa_u32(smem_u32(&full[ent]), 0);
          tma_load_2d_cg2(sRing + ent * kEntry, &tmW2, lead, j * kChunk + e * 32, r * (kD / 2));
        }
        __syncwarp();
        if (++ent == kEntries) { ent = 0; ph ^= 1u; }
      }
    };
    int it = 0;
    for (int t = pair; t < num_pt; t += npairs, ++it) {
      const int tile = 2 * t + r;
      w1_chunk(0);   // four of the five ring entries fill while the previous tile's epilogue still reads its X tile
      mbar_wait_t(x_empty, (static_cast<uint32_t>(it) & 1u) ^ 1u, 43, w_xempty);
      if (elect_one_sync()) {
        if (rank == 0) mbar_expect_tx(x_full, 2u * 8u * kSlab);
        const uint32_t lead = mapa_u32(smem_u32(x_full), 0);
        for (int s = 0; s < 8; ++s) tma_load_2d_cg2(sX + s * kSlab, &tmX, lead, s * 32, tile * kRows);
      }
      __syncwarp();
      for (int j = 0; j < NC; ++j) {
        if (j + 1 < NC) w1_chunk(j + 1);
        w2_chunk(j);
      }
    }
    if (p.dbg && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[0] = FFN_CLOCK() - t_begin; d[1] = w_empty; d[2] = w_xempty;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: leader CTA, one elected lane
    if (rank == 0) {
      constexpr uint32_t idesc_s = umma_idesc(2, 2 * kRows, kChunk);
      constexpr uint32_t idesc_o = umma_idesc(2, 2 * kRows, kD);
      const uint32_t sx = smem_u32(sX);
      const uint32_t sring = smem_u32(sRing);
      const uint32_t tO = tmem_base + 256u;
      int ent = 0;
      uint32_t ph = 0;
      int it = 0;
      long long w_full1 = 0, w_full2 = 0, w_h = 0, w_o = 0, w_x = 0;
      const long long t_begin = FFN_CLOCK();
      for (int t = pair; t < num_pt; t += npairs, ++it) {
        mbar_wait_t(x_full, static_cast<uint32_t>(it) & 1u, 44, w_x);
        tc_fence_after();
        auto issue_s = [&](int j) {
          const uint32_t sbuf = tmem_base + static_cast<uint32_t>((j & 1) * kChunk);
          for (int e = 0; e < 4; ++e) {
            mbar_wait_t(&full[ent], ph, 45, w_full1);
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint64_t adesc = umma_desc_sw128(sx + (2 * e + kk) * kSlab);
                const uint64_t bdesc = umma_desc_sw128(sring + ent * kEntry + kk * (kEntry / 2));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_ss_cg2<true>(sbuf, adesc + 2u * k, bdesc + 2u * k, idesc_s, (e | kk | k) != 0 ? 1u : 0u);
              }
              tc_commit_cg2(&empty[ent], 3);
              if (e == 3) {
                tc_commit_cg2(&s_full[j & 1], 3);
                if (j == NC - 1) tc_commit_cg2(x_empty, 3);     // the tensor pipe is done with both X tiles
              }
            }
            __syncwarp();
            if (++ent == kEntries) { ent = 0; ph ^= 1u; }
          }
        };
        issue_s(0);
        for (int j = 0; j < NC; ++j) {
          if (j + 1 < NC) issue_s(j + 1);                           // next scores while the activation of j runs
          const uint32_t g = static_cast<uint32_t>(it) * static_cast<uint32_t>(NC) + static_cast<uint32_t>(j);
          mbar_wait_t(&h_ready[j & 1], (g >> 1) & 1u, 46, w_h);
          if (j == 0) mbar_wait_t(o_empty, (static_cast<uint32_t>(it) & 1u) ^ 1u, 47, w_o);   // previous tile's O drained
          tc_fence_after();
          const uint32_t hbuf = tmem_base + static_cast<uint32_t>((j & 1) * kChunk);
          for (int e = 0; e < 4; ++e) {
            mbar_wait_t(&full[ent], ph, 48, w_full2);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint64_t bdesc = umma_desc_sw128(sring + ent * kEntry);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_ts_cg2_tf32(tO, hbuf + static_cast<uint32_t>(e * 32 + k * 8), bdesc + 2u * k, idesc_o,
                                 (j | e | k) != 0 ? 1u : 0u);
              tc_commit_cg2(&empty[ent], 3);
              if (j == NC - 1 && e == 3) tc_commit_cg2(o_full, 3);
            }
            __syncwarp();
            if (++ent == kEntries) { ent = 0; ph ^= 1u; }
          }
        }
      }
      if (p.dbg && lane == 0) {
        long long* d = p.dbg + blockIdx.x * 16;
        d[3] = FFN_CLOCK() - t_begin; d[4] = w_full1; d[5] = w_full2; d[6] = w_h; d[7] = w_o; d[8] = w_x;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ row warps: activation, then the tile epilogue
    const int q = warp & 3;                         // TMEM lane quarter
    const int half = (warp - 2) >> 2;               // which 64 columns of a chunk this warp rectifies
    const int row = q * 32 + lane;                  // row of the tile this thread owns
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t xrow = smem_u32(sX) + row * 128;
    const uint32_t stage = smem_u32(sStage) + (warp - 2) * kStage;
    const int sw = row & 7;
    long long w_s = 0, w_of = 0, t_p1 = 0, t_p2 = 0, t_p3 = 0, t_epi = 0;
    const long long t_begin = FFN_CLOCK();
    int it = 0;
    for (int t = pair; t < num_pt; t += npairs, ++it) {
      const int tile = 2 * t + static_cast<int>(rank);   // 128-row tile of X this CTA owns
      for (int j = 0; j < NC; ++j) {
        const uint32_t g = static_cast<uint32_t>(it) * static_cast<uint32_t>(NC) + static_cast<uint32_t>(j);
        // the chunk's biases are fetched before the wait: their latency is off the S_j -> H_j critical path
        const float* b1 = p.b1 + j * kChunk + half * 64;
        float4 bb[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) bb[k] = __ldg(reinterpret_cast<const float4*>(b1 + 4 * k));   // warp-uniform
        mbar_wait_t(&s_full[j & 1], (g >> 1) & 1u, 39, w_s);
        const long long a0 = FFN_CLOCK();
        tc_fence_after();
        const uint32_t tb = trow + static_cast<uint32_t>((j & 1) * kChunk + half * 64);
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(tb, v0);
        tmem_ld_32x32(tb + 32u, v1);
        tmem_wait_ld();
        const long long a1 = FFN_CLOCK();
        // relu() has already mapped NaN to 0 and the values are >= 0, so round-to-nearest (ties away) to TF32 is an
        // integer add + mask on the bit pattern
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 b4 = bb[k], c4 = bb[8 + k];
          v0[4 * k] = (__float_as_uint(fmaxf(__uint_as_float(v0[4 * k]) + b4.x, 0.f)) + 0x1000u) & 0xffffe000u;
          v0[4 * k + 1] = (__float_as_uint(fmaxf(__uint_as_float(v0[4 * k + 1]) + b4.y, 0.f)) + 0x1000u) & 0xffffe000u;
          v0[4 * k + 2] = (__float_as_uint(fmaxf(__uint_as_float(v0[4 * k + 2]) + b4.z, 0.f)) + 0x1000u) & 0xffffe000u;
          v0[4 * k + 3] = (__float_as_uint(fmaxf(__uint_as_float(v0[4 * k + 3]) + b4.w, 0.f)) + 0x1000u) & 0xffffe000u;
          v1[4 * k] = (__float_as_uint(fmaxf(__uint_as_float(v1[4 * k]) + c4.x, 0.f)) + 0x1000u) & 0xffffe000u;
          v1[4 * k + 1] = (__float_as_uint(fmaxf(__uint_as_float(v1[4 * k + 1]) + c4.y, 0.f)) + 0x1000u) & 0xffffe000u;
          v1[4 * k + 2] = (__float_as_uint(fmaxf(__uint_as_float(v1[4 * k + 2]) + c4.z, 0.f)) + 0x1000u) & 0xffffe000u;
          v1[4 * k + 3] = (__float_as_uint(fmaxf(__uint_as_float(v1[4 * k + 3]) + c4.w, 0.f)) + 0x1000u) & 0xffffe000u;
        }
        const long long a2 = FFN_CLOCK();
        tmem_st_32x32(tb, v0);
        tmem_st_32x32(tb + 32u, v1);
        tmem_wait_st();
        const long long a3 = FFN_CLOCK();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&h_ready[j & 1]), 0));   // leader's barrier
#ifdef SPE_FFN_TIMING_ACT
        const long long a4 = FFN_CLOCK();
        t_p1 += a1 - a0; t_p2 += a2 - a1; t_p3 += a3 - a2; t_epi += a4 - a3;
#else
        (void)a0; (void)a1; (void)a2; (void)a3;
#endif
      }
      // ---- epilogue of the tile: v = O + b2 + X, LayerNorm over the row.  All eight row warps take part: the two warps
      //      of a lane quarter split the 256 columns (128 each) and exchange their partial sums through shared memory,
      //      which halves the time the tensor pipe waits for the O accumulator and the X tile to be released.
      mbar_wait_t(o_full, static_cast<uint32_t>(it) & 1u, 40, w_of);
      const long long e0 = FFN_CLOCK();
      tc_fence_after();
      const uint32_t to = trow + 256u + static_cast<uint32_t>(half * 128);
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int cb = half * 4 + c;                 // 32-column block of the row
        uint32_t v[32];
        tmem_ld_32x32(to + static_cast<uint32_t>(c * 32), v);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 x4 = ld_shared_v4(xrow + cb * kSlab + ((k ^ sw) * 16));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + cb * 32 + 4 * k));
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
      sXchg[half * kRows + row] = sum;
      named_bar_sync(1 + q, 64);                    // the two warps of this lane quarter
      sum += sXchg[(half ^ 1) * kRows + row];
      named_bar_sync(1 + q, 64);                    // both have read before the array is reused
      const long long e1 = FFN_CLOCK();
      const float mean = sum * (1.f / 256.f);
      float qs = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(to + static_cast<uint32_t>(c * 32), v);
        tmem_wait_ld();
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d0 = __uint_as_float(v[4 * k]) - mean, d1 = __uint_as_float(v[4 * k + 1]) - mean;
          const float d2 = __uint_as_float(v[4 * k + 2]) - mean, d3 = __uint_as_float(v[4 * k + 3]) - mean;
          q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
        }
        qs += (q0 + q1) + (q2 + q3);
      }
      sXchg[half * kRows + row] = qs;
      named_bar_sync(1 + q, 64);
      qs += sXchg[(half ^ 1) * kRows + row];
      const float rstd = rsqrtf(qs * (1.f / 256.f) + 1e-5f);
      const long long e2 = FFN_CLOCK();
      const long long row0 = static_cast<long long>(tile) * kRows + q * 32;      // first row of this warp
      const long long left = p.M - row0;
      const int rows_ok = left >= 32 ? 32 : (left > 0 ? static_cast<int>(left) : 0);
      const long long ld = p.out_mode == 2 ? 768 : 256;
      float* obase = p.out + row0 * ld;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col0 = (half * 4 + c) * 32;
        uint32_t v[32];
        tmem_ld_32x32(to + static_cast<uint32_t>(c * 32), v);
        tmem_wait_ld();
        float4 y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + col0 + 4 * k));
          const float4 e4 = __ldg(reinterpret_cast<const float4*>(p.beta + col0 + 4 * k));
          y[k].x = (__uint_as_float(v[4 * k]) - mean) * rstd * g4.x + e4.x;
          y[k].y = (__uint_as_float(v[4 * k + 1]) - mean) * rstd * g4.y + e4.y;
          y[k].z = (__uint_as_float(v[4 * k + 2]) - mean) * rstd * g4.z + e4.z;
          y[k].w = (__uint_as_float(v[4 * k + 3]) - mean) * rstd * g4.w + e4.w;
        }
        if (p.out_mode == 3) {
          // bf16 storage: the row's 32 columns as 16 packed words (one staged 64-byte segment per row)
          float4 pk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 b0 = __floats2bfloat162_rn(y[2 * k].x, y[2 * k].y), b1 = __floats2bfloat162_rn(y[2 * k].z, y[2 * k].w);
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(y[2 * k + 1].x, y[2 * k + 1].y), b3 = __floats2bfloat162_rn(y[2 * k + 1].z, y[2 * k + 1].w);
            pk[k] = make_float4(__uint_as_float(*reinterpret_cast<const uint32_t*>(&b0)), __uint_as_float(*reinterpret_cast<const uint32_t*>(&b1)),
                                __uint_as_float(*reinterpret_cast<const uint32_t*>(&b2)), __uint_as_float(*reinterpret_cast<const uint32_t*>(&b3)));
          }
          store_half_block(stage, lane, pk, p.out + row0 * 128 + col0 / 2, nullptr, 128, rows_ok);
        } else if (p.out_mode == 2) {
          float4 hi[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            hi[k] = make_float4(rna_tf32(y[k].x), rna_tf32(y[k].y), rna_tf32(y[k].z), rna_tf32(y[k].w));
            y[k] = make_float4(rna_tf32(y[k].x - hi[k].x), rna_tf32(y[k].y - hi[k].y), rna_tf32(y[k].z - hi[k].z),
                               rna_tf32(y[k].w - hi[k].w));
          }
          // [hi | lo | hi]: hi goes to columns [0,256) and [512,768), lo to [256,512)
          store_half_block(stage, lane, hi, obase + col0, obase + 512 + col0, ld, rows_ok);
          store_half_block(stage, lane, hi + 4, obase + col0 + 16, obase + 512 + col0 + 16, ld, rows_ok);
          store_half_block(stage, lane, y, obase + 256 + col0, nullptr, ld, rows_ok);
          store_half_block(stage, lane, y + 4, obase + 256 + col0 + 16, nullptr, ld, rows_ok);
        } else {
          if (p.out_mode == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              y[k] = make_float4(rna_tf32(y[k].x), rna_tf32(y[k].y), rna_tf32(y[k].z), rna_tf32(y[k].w));
          }
          store_half_block(stage, lane, y, obase + col0, nullptr, ld, rows_ok);
          store_half_block(stage, lane, y + 4, obase + col0 + 16, nullptr, ld, rows_ok);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(o_empty), 0));
#if defined(SPE_FFN_TIMING) && !defined(SPE_FFN_TIMING_ACT)
      const long long e3 = FFN_CLOCK();
      t_p1 += e1 - e0; t_p2 += e2 - e1; t_p3 += e3 - e2; t_epi += e3 - e0;
#else
      (void)e0; (void)e1; (void)e2; (void)t_p1; (void)t_p2; (void)t_p3; (void)t_epi;
#endif
    }
    if (p.dbg && warp == 2 && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[9] = FFN_CLOCK() - t_begin; d[10] = w_s; d[11] = w_of; d[12] = t_p1; d[13] = t_p2; d[14] = t_p3; d[15] = t_epi;
    }
  }

  tc_fence_before();
  cluster_sync_all();   // no CTA of the pair may exit (or free TMEM) while its partner still uses it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
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
  if (num_sms < 2) return "ffn: needs a CTA pair";
  FfnParams p{};
  p.M = d.M;
  p.num_tiles = static_cast<int>((d.M + kRows - 1) / kRows);
  p.num_chunks = d.hidden / kChunk;
  p.b1 = d.b1; p.b2 = d.b2; p.gamma = d.gamma; p.beta = d.beta;
  p.out = reinterpret_cast<float*>(d.out);
  p.out_mode = d.out_mode;
  static const bool dbg_on = getenv("SPE_FFN_DBG") != nullptr;
  static long long* dbg_dev = nullptr;
  if (dbg_on && dbg_dev == nullptr) SPE_CUDA_TRY(cudaMalloc(&dbg_dev, 256 * 16 * sizeof(long long)));
  if (dbg_on) SPE_CUDA_TRY(cudaMemsetAsync(dbg_dev, 0, 256 * 16 * sizeof(long long), stream));
  p.dbg = dbg_on ? dbg_dev : nullptr;
  CUtensorMap tmX, tmW1, tmW2;
  std::string e = encode_tmap_2d(&tmX, kTF32, d.X, kD, d.M, static_cast<long long>(kD) * 4, 32, kRows);
  if (!e.empty()) return "ffn X map: " + e;
  e = encode_tmap_2d(&tmW1, kTF32, d.W1, kD, d.hidden, static_cast<long long>(kD) * 4, 32, kChunk / 2);
  if (!e.empty()) return "ffn W1 map: " + e;
  e = encode_tmap_2d(&tmW2, kTF32, d.W2, d.hidden, kD, static_cast<long long>(d.hidden) * 4, 32, kD / 2);
  if (!e.empty()) return "ffn W2 map: " + e;
  static bool attr_set = false;
  if (!attr_set) {
    SPE_CUDA_TRY(cudaFuncSetAttribute(ffn_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int num_pt = (p.num_tiles + 1) / 2;
  const int pairs = num_pt < num_sms / 2 ? num_pt : num_sms / 2;
  {
    ProfScope ps(kFamGemm, stream);
    SPE_CUDA_TRY(launch_pdl(ffn_tc2_kernel, dim3(2 * pairs), dim3(kThreads), kSmemBytes, stream, tmX, tmW1, tmW2, p));
  }
  SPE_CUDA_TRY(cudaGetLastError());
  if (dbg_on) {   // only meaningful in a -DSPE_FFN_TIMING build
    static long long host[256 * 16];
    SPE_CUDA_TRY(cudaStreamSynchronize(stream));
    SPE_CUDA_TRY(cudaMemcpy(host, dbg_dev, sizeof(host), cudaMemcpyDeviceToHost));
    static const char* names[16] = {"prod.total", "prod.wait_empty", "prod.wait_x_empty", "mma.total", "mma.wait_full_w1",
                                    "mma.wait_full_w2", "mma.wait_h_ready", "mma.wait_o_empty", "mma.wait_x_full",
                                    "row.total", "row.wait_s_full", "row.wait_o_full", "epi.pass1", "epi.pass2",
                                    "epi.pass3", "epi.total"};
    for (int cta : {0, 1, 146}) {
      fprintf(stderr, "[ffn dbg] cta %3d:", cta);
      for (int k = 0; k < 16; ++k) fprintf(stderr, " %s=%lld", names[k], host[cta * 16 + k]);
      fprintf(stderr, "\n");
    }
  }
  return "";
}

}  // namespace spe
