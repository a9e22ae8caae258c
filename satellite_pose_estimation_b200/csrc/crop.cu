// Fused crop -> bicubic resize -> /255 -> ImageNet normalise, straight from the 1920x1200 grayscale frame.
//
// Replaces SpeedSubmission.__getitem__ (reference: RV/datasets/speed.py:113-160): the zero S x S canvas, the
// slice copy of image-intersect-box, cv2.resize(INTER_CUBIC) to R x R (uint8 result), to_tensor (/255) and
// Normalize (RV/datasets/speed.py:25-41).  Nothing but the final fp32 NCHW tensor is written to HBM.
//
// Resampling model (matches cv2 4.x INTER_CUBIC on uint8, see DESIGN.md "crop"): Keys cubic with a = -0.75,
// half-pixel centres f = (d + 0.5) * S / R - 0.5, taps floor(f)-1 .. floor(f)+2 clamped to the canvas (replicate),
// w3 = 1 - w0 - w1 - w2, separable (horizontal then vertical), no antialiasing, result rounded half-to-even and
// saturated to uint8.  Arithmetic is fp64: the kernel is bound by its 3 * R * R * 4-byte store, not by math, and
// fp64 keeps the pre-rounding value far from accidental .5 flips (fp32 flips ~0.03 % of pixels at S > 1000).
//
// The gray frame is read once per tap through the read-only path; the three output channels are the same
// uint8 value pushed through three per-channel affine maps, exactly like the reference's replicated RGB image.
#include "spe_internal.h"

#include <cuda_bf16.h>
#include "profile.h"
#include "spe_ptx.cuh"

#include <stdlib.h>

namespace spe {

namespace {

__device__ __forceinline__ void cubic_w(double t, double (&w)[4]) {
  const double A = -0.75;
  w[0] = ((A * (t + 1.0) - 5.0 * A) * (t + 1.0) + 8.0 * A) * (t + 1.0) - 4.0 * A;
  w[1] = ((A + 2.0) * t - (A + 3.0)) * t * t + 1.0;
  w[2] = ((A + 2.0) * (1.0 - t) - (A + 3.0)) * (1.0 - t) * (1.0 - t) + 1.0;
  w[3] = 1.0 - w[0] - w[1] - w[2];
}

// One thread per output COLUMN, kCropRows output rows per block.
//  * the horizontal taps (four clamped frame columns + four fp64 cubic weights) depend on the column only and are
//    computed once per thread; the vertical taps (four clamped frame rows + weights) depend on the row only and are
//    computed once per BLOCK by its first kCropRows threads and shared through shared memory
//  * ncu on the first version (profiles/r02_ncu_crop.md) showed the kernel bound by the XU pipe (46 % busy: 16 uint8 ->
//    fp64 conversions, floor, rint and fp64 -> int per pixel), not by memory.  Conversions now stay off that pipe: a byte
//    becomes a double by OR-ing it into the mantissa of 2^52 and subtracting 2^52 (exact), and the final rounding is
//    the classic add-and-subtract of 1.5 * 2^52 (round-half-even in the default rounding mode, exactly rint() for
//    |x| < 2^51), the integer read straight from the low word of the sum
//  * threadIdx.x walks ox, so the three channel stores of every row are fully coalesced, and the 4 x 4 source bytes of
//    neighbouring columns share sectors (at the median crop side of 430 px a warp's 32 columns span 61 source bytes)
constexpr int kCropRows = 8;

__device__ __forceinline__ double u8_to_f64(unsigned v) {          // exact, no conversion instruction
  return __hiloint2double(0x43300000, static_cast<int>(v)) - 4503599627370496.0;   // (2^52 + v) - 2^52
}

__global__ void __launch_bounds__(256)
crop_resize_norm_kernel(const uint8_t* __restrict__ frames, int H, int W, long long pitch, long long frame_stride,
                        const int32_t* __restrict__ boxes, int R, float* __restrict__ out) {
  __shared__ double s_wy[kCropRows][4];
  __shared__ int s_gy[kCropRows][4];          // frame row of each vertical tap, -1 = outside the frame (zero canvas)
  const int b = blockIdx.z;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy0 = blockIdx.y * kCropRows;

  const int x1 = boxes[b * 4 + 0];
  const int y1 = boxes[b * 4 + 1];
  // canvas extent: square (clip_size) on the submission path, any rectangle on the eval path (PIL crop squashed to R x R)
  const int S = boxes[b * 4 + 2] - x1;
  const int Sy = boxes[b * 4 + 3] - y1;
  const bool empty = !(S > 0 && Sy > 0);

  if (threadIdx.x < kCropRows && !empty) {
    const int oy = oy0 + threadIdx.x;
    const double fy = (oy + 0.5) * (static_cast<double>(Sy) / static_cast<double>(R)) - 0.5;
    const double fly = floor(fy);
    const int sy = static_cast<int>(fly);
    double wy[4];
    cubic_w(fy - fly, wy);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = sy - 1 + j;
      c = c < 0 ? 0 : (c > Sy - 1 ? Sy - 1 : c);   // replicate the canvas border
      const int gy = y1 + c;                       // canvas -> frame row
      s_wy[threadIdx.x][j] = wy[j];
      s_gy[threadIdx.x][j] = (gy >= 0 && gy < H) ? gy : -1;
    }
  }
  __syncthreads();
  if (ox >= R) return;

  const uint8_t* frame = frames + static_cast<long long>(b) * frame_stride;
  const long long plane = static_cast<long long>(R) * R;
  double wx[4];
  int cx[4];
  if (!empty) {
    const double fx = (ox + 0.5) * (static_cast<double>(S) / static_cast<double>(R)) - 0.5;
    const double flx = floor(fx);
    const int sx = static_cast<int>(flx);
    cubic_w(fx - flx, wx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = sx - 1 + i;
      c = c < 0 ? 0 : (c > S - 1 ? S - 1 : c);  // replicate the canvas border
      const int gx = x1 + c;                    // canvas -> frame column
      const bool in = (gx >= 0) && (gx < W);
      cx[i] = in ? gx : 0;
      if (!in) wx[i] = 0.0;                     // outside the frame: zero canvas (0 * w contributes nothing)
    }
  }
  float* o = out + static_cast<long long>(b) * 3 * plane + static_cast<long long>(oy0) * R + ox;
#pragma unroll 2
  for (int r = 0; r < kCropRows; ++r, o += R) {
    if (oy0 + r >= R) break;
    int iv = 0;
    if (!empty) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gy = s_gy[r][j];
        if (gy >= 0) {
          const uint8_t* rp = frame + static_cast<long long>(gy) * pitch;
          double row = u8_to_f64(__ldg(rp + cx[0])) * wx[0];
          row += u8_to_f64(__ldg(rp + cx[1])) * wx[1];
          row += u8_to_f64(__ldg(rp + cx[2])) * wx[2];
          row += u8_to_f64(__ldg(rp + cx[3])) * wx[3];
          acc += row * s_wy[r][j];
        }
      }
      // round half to even like cvRound / IPP, saturate to uint8
      const double magic = acc + 6755399441055744.0;                 // 1.5 * 2^52: the integer sits in the low word
      iv = __double2loint(magic);
      iv = iv < 0 ? 0 : (iv > 255 ? 255 : iv);
    }
    // to_tensor: uint8 -> float32 / 255 ; Normalize: (x - mean) / std, all IEEE fp32 like the torch CPU ops
    const float x = __fdiv_rn(static_cast<float>(iv), 255.0f);
    o[0] = __fdiv_rn(__fsub_rn(x, 0.485f), 0.229f);
    o[plane] = __fdiv_rn(__fsub_rn(x, 0.456f), 0.224f);
    o[2 * plane] = __fdiv_rn(__fsub_rn(x, 0.406f), 0.225f);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Staged version (the one the library launches when the frame layout allows 16-byte bulk copies).
// ncu on the gather version (profiles/r02_ncu_crop_attention.md): 24.6 MB of DRAM reads for 19.6 MB of algorithmic
// source bytes -- nothing wasted -- but only 6.7 % of the DRAM bandwidth in use: 16 dependent byte gathers per output
// pixel, each waiting out a DRAM (L1-miss) latency.  Here one CTA (kCropRows output rows of one image) first pulls the
// kCropRows x 4 frame-row segments its outputs read into shared memory with cp.async.bulk (one elected thread per
// segment, one mbarrier for all of them: the loads of a CTA are all in flight at once and arrive as whole 16-byte
// aligned, fully coalesced lines), then computes from shared memory with the same arithmetic, in the same order, as
// above -- the results are bit-identical.
// Segment of tap row (r, j): frame columns [a0, a1) = the crop's columns inside the frame, widened to 16-byte
// boundaries; rows outside the frame are not loaded (their weight path is skipped, zero canvas).
__global__ void __launch_bounds__(256, 3)
crop_resize_norm_staged_kernel(const uint8_t* __restrict__ frames, int H, int W, long long pitch, long long frame_stride,
                               const int32_t* __restrict__ boxes, int R, float* __restrict__ out, int row_stride,
                               void* __restrict__ stem_out, int stem_bf16) {
  extern __shared__ __align__(16) uint8_t s_rows[];      // [kCropRows * 4][row_stride]
  __shared__ double s_wy[kCropRows][4];
  __shared__ int s_gy[kCropRows][4];          // frame row of each vertical tap, -1 = outside the frame (zero canvas)
  __shared__ __align__(8) uint64_t s_bar;
  // to_tensor + Normalize of each of the 256 possible resampled values, per channel: the four IEEE divisions per output
  // pixel (a fifth of the kernel's instructions: ncu r02d) become three shared-memory reads of the same bits
  __shared__ float s_norm[3][256];
  const int b = blockIdx.z;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy0 = blockIdx.y * kCropRows;

  const int x1 = boxes[b * 4 + 0];
  const int y1 = boxes[b * 4 + 1];
  const int S = boxes[b * 4 + 2] - x1;
  const int Sy = boxes[b * 4 + 3] - y1;
  const bool empty = !(S > 0 && Sy > 0);
  // the crop's columns inside the frame, widened to 16-byte boundaries (W and pitch are multiples of 16 here)
  const int xa = x1 > 0 ? x1 : 0;
  const int xb = (x1 + S) < W ? (x1 + S) : W;
  const int a0 = xa & ~15;
  const int a1 = (xb + 15) & ~15;
  const bool any_col = !empty && xb > xa;
  const uint8_t* frame = frames + static_cast<long long>(b) * frame_stride;

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  for (int v = threadIdx.x; v < 256; v += blockDim.x) {
    const float x = __fdiv_rn(static_cast<float>(v), 255.0f);
    s_norm[0][v] = __fdiv_rn(__fsub_rn(x, 0.485f), 0.229f);
    s_norm[1][v] = __fdiv_rn(__fsub_rn(x, 0.456f), 0.224f);
    s_norm[2][v] = __fdiv_rn(__fsub_rn(x, 0.406f), 0.225f);
  }
  if (threadIdx.x < kCropRows && !empty) {
    const int oy = oy0 + threadIdx.x;
    const double fy = (oy + 0.5) * (static_cast<double>(Sy) / static_cast<double>(R)) - 0.5;
    const double fly = floor(fy);
    const int sy = static_cast<int>(fly);
    double wy[4];
    cubic_w(fy - fly, wy);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = sy - 1 + j;
      c = c < 0 ? 0 : (c > Sy - 1 ? Sy - 1 : c);   // replicate the canvas border
      const int gy = y1 + c;                       // canvas -> frame row
      const bool live = gy >= 0 && gy < H && oy < R;
      s_wy[threadIdx.x][j] = live ? wy[j] : 0.0;       // rows outside the frame: zero canvas (weight 0 on a stale row)
      s_gy[threadIdx.x][j] = live ? gy : -1;
    }
  }
  __syncthreads();
  if (threadIdx.x < 32 && any_col) {
    // lane t owns segment (row t / 4, tap t % 4); the expected byte count is posted before any copy is issued
    const int gy = threadIdx.x < kCropRows * 4 ? s_gy[threadIdx.x >> 2][threadIdx.x & 3] : -1;
    const uint32_t seg = static_cast<uint32_t>(a1 - a0);
    const unsigned live = __ballot_sync(0xffffffffu, gy >= 0);
    if (threadIdx.x == 0) mbar_expect_tx(&s_bar, seg * static_cast<uint32_t>(__popc(live)));
    __syncwarp();
    if (gy >= 0) bulk_load_1d(s_rows + threadIdx.x * row_stride, frame + static_cast<long long>(gy) * pitch + a0, seg, &s_bar);
  } else if (threadIdx.x == 0) {
    mbar_arrive(&s_bar);                          // nothing to load: complete the phase by hand
  }

  // horizontal taps while the copies fly
  double wx[4];
  int cx[4];
  if (!empty && ox < R) {
    const double fx = (ox + 0.5) * (static_cast<double>(S) / static_cast<double>(R)) - 0.5;
    const double flx = floor(fx);
    const int sx = static_cast<int>(flx);
    cubic_w(fx - flx, wx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = sx - 1 + i;
      c = c < 0 ? 0 : (c > S - 1 ? S - 1 : c);  // replicate the canvas border
      const int gx = x1 + c;                    // canvas -> frame column
      const bool in = (gx >= 0) && (gx < W);
      cx[i] = in ? gx - a0 : 0;                 // offset inside the staged segment
      if (!in) wx[i] = 0.0;                     // outside the frame: zero canvas (0 * w contributes nothing)
    }
  }
  mbar_wait(&s_bar, 0, 31);
  if (ox >= R) return;

  const long long plane = static_cast<long long>(R) * R;
  float* o = out + static_cast<long long>(b) * 3 * plane + static_cast<long long>(oy0) * R + ox;
#pragma unroll 2
  for (int r = 0; r < kCropRows; ++r, o += R) {
    if (oy0 + r >= R) break;
    int iv = 0;
    if (any_col) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint8_t* rp = s_rows + (r * 4 + j) * row_stride;
        double row = u8_to_f64(rp[cx[0]]) * wx[0];
        row += u8_to_f64(rp[cx[1]]) * wx[1];
        row += u8_to_f64(rp[cx[2]]) * wx[2];
        row += u8_to_f64(rp[cx[3]]) * wx[3];
        acc += row * s_wy[r][j];
      }
      const double magic = acc + 6755399441055744.0;                 // 1.5 * 2^52: the integer sits in the low word
      iv = __double2loint(magic);
      iv = iv < 0 ? 0 : (iv > 255 ? 255 : iv);
    }
    if (stem_out != nullptr) {
      // the predictor's stem input directly: zero-bordered NHWC with the three channels padded to 16 bytes (what
      // stem_pad_kernel would make of the NCHW tensor, same rounding) -- the border was zeroed when the buffer was allocated
      const long long Rp = R + 6;
      const long long pix = (static_cast<long long>(b) * Rp + (oy0 + r + 3)) * Rp + (ox + 3);
      if (!stem_bf16) {
        uint32_t a0, a1, a2;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a0) : "f"(s_norm[0][iv]));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a1) : "f"(s_norm[1][iv]));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a2) : "f"(s_norm[2][iv]));
        reinterpret_cast<uint4*>(stem_out)[pix] = make_uint4(a0, a1, a2, 0u);
      } else {
        const __nv_bfloat162 p01 = __floats2bfloat162_rn(s_norm[0][iv], s_norm[1][iv]);
        const __nv_bfloat162 p2z = __floats2bfloat162_rn(s_norm[2][iv], 0.f);
        reinterpret_cast<uint4*>(stem_out)[pix] = make_uint4(*reinterpret_cast<const uint32_t*>(&p01),
                                                              *reinterpret_cast<const uint32_t*>(&p2z), 0u, 0u);
      }
      continue;
    }
    o[0] = s_norm[0][iv];
    o[plane] = s_norm[1][iv];
    o[2 * plane] = s_norm[2][iv];
  }
}

}  // namespace

std::string launch_crop_resize_norm(const uint8_t* frames, int H, int W, long long pitch, long long frame_stride,
                                    const int32_t* boxes, int B, int R, float* out_nchw, cudaStream_t s, void* stem_out,
                                    int stem_bf16, bool* stem_written) {
  if (stem_written) *stem_written = false;
  if (B <= 0) return "";
  if (R <= 0 || R > 4096) return "crop: bad output size";
  const int tx = R >= 256 ? 256 : ((R + 31) / 32) * 32;
  dim3 block(tx);
  dim3 grid((R + tx - 1) / tx, (R + kCropRows - 1) / kCropRows, B);
  ProfScope ps(kFamCrop, s);
  // staged version: every frame row must be addressable in 16-byte units, a CTA must cover whole output rows (the
  // staged segment is the crop's full width) and the 32 segments must fit shared memory
  static const bool no_stage = getenv("SPE_CROP_GATHER") != nullptr && atoi(getenv("SPE_CROP_GATHER")) != 0;
  const int row_stride = ((W + 15) / 16) * 16;
  const size_t smem = static_cast<size_t>(kCropRows) * 4 * row_stride;
  const bool aligned = (reinterpret_cast<uintptr_t>(frames) % 16 == 0) && pitch % 16 == 0 && frame_stride % 16 == 0 &&
                       W % 16 == 0 && pitch >= row_stride;
  if (!no_stage && aligned && grid.x == 1 && smem <= 96 * 1024) {
    static bool attr = false;
    if (!attr) {
      SPE_CUDA_TRY(cudaFuncSetAttribute(crop_resize_norm_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr = true;
    }
    crop_resize_norm_staged_kernel<<<grid, block, smem, s>>>(frames, H, W, pitch, frame_stride, boxes, R, out_nchw, row_stride,
                                                             stem_out, stem_bf16);
    if (stem_written) *stem_written = stem_out != nullptr;
  } else {
    crop_resize_norm_kernel<<<grid, block, 0, s>>>(frames, H, W, pitch, frame_stride, boxes, R, out_nchw);
  }
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

}  // namespace spe
