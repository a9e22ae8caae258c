// TMA ring micro-benchmark: how long does one 16 KB ring entry take to come back (commit -> producer -> TMA -> L2 -> smem
// -> consumer) as a function of the ring depth and of how many SMs stream at once?  One CTA per SM, warp 0 = producer,
// warp 1 = consumer that holds every entry for `hold` cycles (the tensor-pipe time of an entry: 512) and releases it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I satellite_pose_estimation_b200/csrc \
//        tools/tma_ring_probe.cu -o tools/tma_ring_probe && tools/tma_ring_probe
#include "spe_ptx.cuh"
#include <cstdio>
#include <vector>
using namespace spe;

constexpr int kEntry = 16384;   // 128 rows x 128 bytes
constexpr int kMaxDepth = 12;

__global__ void __launch_bounds__(64, 1)
ring_kernel(const __grid_constant__ CUtensorMap tm, int depth, int n, int hold, int rows_total, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kMaxDepth * kEntry);
  uint64_t* empty = full + kMaxDepth;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm);
    for (int i = 0; i < depth; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    int e = 0; uint32_t ph = 0;
    int row = (blockIdx.x * 7919 * 128) % rows_total;
    for (int i = 0; i < n; ++i) {
      mbar_wait(&empty[e], ph ^ 1u, 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&full[e], kEntry);
        tma_load_2d(smem + e * kEntry, &tm, &full[e], 0, row);
      }
      __syncwarp();
      row += 128 * 149; if (row >= rows_total) row -= rows_total;
      row -= row % 128;
      if (++e == depth) { e = 0; ph ^= 1u; }
    }
  } else {
    int e = 0; uint32_t ph = 0;
    for (int i = 0; i < n; ++i) {
      mbar_wait(&full[e], ph, 2);
      const long long c0 = clock64();
      while (clock64() - c0 < hold) {}
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[e]);
      if (++e == depth) { e = 0; ph ^= 1u; }
    }
    if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const int smem_bytes = kMaxDepth * kEntry + 2 * kMaxDepth * 8 + 1024;
  cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  long long* cyc;
  cudaMalloc(&cyc, sizeof(long long) * 1024);
  for (int big = 0; big < 2; ++big) {
    // L2-resident source (8 MB) or HBM-sized source (2 GB)
    const long long rows = big ? (1ll << 24) : (1ll << 16);
    float* src;
    cudaMalloc(&src, rows * 128);
    cudaMemset(src, 0, rows * 128);
    CUtensorMap tm;
    cuuint64_t dims[2] = {32, static_cast<cuuint64_t>(rows)};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int grid : {1, sms}) {
      for (int hold : {0, 512}) {
        printf("source %s, %3d CTAs, hold %3d cycles: cycles per 16 KB entry at depth", big ? "2 GB (HBM)" : "8 MB (L2) ", grid, hold);
        for (int depth : {1, 2, 3, 4, 6, 8, 12}) {
          const int n = 2000;
          ring_kernel<<<grid, 64, smem_bytes>>>(tm, depth, n, hold, static_cast<int>(rows), cyc);   // warm-up
          ring_kernel<<<grid, 64, smem_bytes>>>(tm, depth, n, hold, static_cast<int>(rows), cyc);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf(" launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          std::vector<long long> h(grid);
          cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (long long v : h) mx = v > mx ? v : mx;
          printf("  %d: %lld", depth, mx / n);
        }
        printf("\n");
      }
    }
    cudaFree(src);
  }
  return 0;
}
