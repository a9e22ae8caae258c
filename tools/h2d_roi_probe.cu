// H2D rate of the per-image ROI uploads of spe_submit_batch_host (pinned host frames, 1200 x 1920 u8):
//   mode 0: cudaMemcpy2DAsync of the ROI rectangle (w x h bytes, pitch 1920)            -- what the library did
//   mode 1: one contiguous copy of the ROI's row band (h rows x 1920 bytes)
//   mode 2: 2-D copy with the x-range widened to multiples of `align` bytes
// Build: nvcc -O2 -o tools/h2d_roi_probe tools/h2d_roi_probe.cu ; run: tools/h2d_roi_probe [roi_side] [align]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char** argv) {
  const int H = 1200, W = 1920, B = 64;
  const int side = argc > 1 ? atoi(argv[1]) : 664;   // 28.2 MB / 64 images = 441 KB = 664^2
  const int align = argc > 2 ? atoi(argv[2]) : 256;
  unsigned char *h, *d;
  cudaMallocHost(&h, (size_t)B * H * W);
  cudaMalloc(&d, (size_t)B * H * W);
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      long long bytes = 0;
      cudaEventRecord(e0, s);
      for (int i = 0; i < B; ++i) {
        const int x0 = (37 * i) % (W - side), y0 = (53 * i) % (H - side);
        const size_t off = (size_t)i * H * W + (size_t)y0 * W;
        if (mode == 0) {
          cudaMemcpy2DAsync(d + off + x0, W, h + off + x0, W, side, side, cudaMemcpyHostToDevice, s);
          bytes += (long long)side * side;
        } else if (mode == 1) {
          cudaMemcpyAsync(d + off, h + off, (size_t)side * W, cudaMemcpyHostToDevice, s);
          bytes += (long long)side * W;
        } else {
          const int xa = x0 / align * align;
          int xb = (x0 + side + align - 1) / align * align;
          if (xb > W) xb = W;
          cudaMemcpy2DAsync(d + off + xa, W, h + off + xa, W, xb - xa, side, cudaMemcpyHostToDevice, s);
          bytes += (long long)(xb - xa) * side;
        }
      }
      cudaEventRecord(e1, s);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2)
        printf("mode %d roi %d: %.2f MB in %.3f ms = %.1f GB/s (%.3f ms per batch of 64)\n", mode, side, bytes / 1e6, ms,
               bytes / ms / 1e6, ms);
    }
  }
  return 0;
}
