// First kernels of the SA (RT-DETR) variant's decoder, SURVEY.md section 8f rank 2:
//
//   ms_deform_attn_kernel   multi-scale deformable attention core
//                           reference: deformable_attention_core_func, SA/src/zoo/rtdetr/utils.py:15-64
//                           (per level F.grid_sample(bilinear, padding_mode='zeros', align_corners=False), weighted sum
//                           over levels x points), optionally fused with what MSDeformableAttention.forward does between
//                           its linear layers and the core (SA/src/zoo/rtdetr/rtdetr_decoder.py:117-163): softmax of the
//                           attention logits over levels x points and sampling_location = reference_point +
//                           offset / (W_l, H_l)
//   topk_queries_kernel     encoder top-k query selection
//                           reference: torch.topk(enc_outputs_class.max(-1).values, num_queries, dim=1),
//                           SA/src/zoo/rtdetr/rtdetr_decoder.py:646-648
//   gather_rows_kernel      the .gather(dim=1, index=topk_ind...) calls that follow it (:651-680)
//
// Layouts are the reference's: value [B, Lv, heads, 32] with the levels concatenated along Lv (level l holds H_l x W_l
// rows, row-major), so one bilinear tap of one head is 32 consecutive floats = one 128-byte line read by one warp.
// HBM / L2-bound gather: 4 taps x levels x points lines per (query, head); no tensor-core work here.
#include "spe_internal.h"
#include "profile.h"

namespace spe {

namespace {

constexpr int kMaxLevels = 8;
constexpr int kMaxLP = 64;        // levels x points per head

struct DeformParams {
  const float* value;             // [B, Lv, heads, 32]
  const float* loc;               // fused: sampling offsets (raw linear output) [B, Lq, heads, L, P, 2]; else locations in [0,1]
  const float* attn;              // fused: attention logits [B, Lq, heads, L*P]; else softmaxed weights
  const float* ref;               // fused only: reference points [B, Lq, ref_levels, 2] (ref_levels = 1 or L)
  float* out;                     // [B, Lq, heads*32]
  int B, Lq, Lv, heads, L, P, ref_levels, fused;
  // row strides in floats: value rows (default heads*32), offset / attention-logit rows per (image, query) (defaults
  // heads*L*P*2 and heads*L*P) -- the SA forward reads all three straight out of wider GEMM outputs
  long long value_ld, loc_ld, attn_ld;
  int h[kMaxLevels], w[kMaxLevels], start[kMaxLevels];
};

// one warp per (batch, query, head); lane = channel of the head (head_dim 32)
__global__ void __launch_bounds__(256)
ms_deform_attn_kernel(const DeformParams p) {
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.B * p.Lq * p.heads;
  if (warp_global >= total) return;
  const int hd = warp_global % p.heads;
  const int q = (warp_global / p.heads) % p.Lq;
  const int b = warp_global / (p.heads * p.Lq);
  const int LP = p.L * p.P;
  const long long qh = (static_cast<long long>(b) * p.Lq + q) * p.heads + hd;
  const long long bq = static_cast<long long>(b) * p.Lq + q;
  const float* loc = p.loc + bq * p.loc_ld + static_cast<long long>(hd) * LP * 2;
  const float* att = p.attn + bq * p.attn_ld + static_cast<long long>(hd) * LP;

  // attention weights of this (query, head): lanes hold one (level, point) each (LP <= 64 -> two per lane)
  float w0 = lane < LP ? att[lane] : -INFINITY, w1 = lane + 32 < LP ? att[lane + 32] : -INFINITY;
  if (p.fused) {
    // F.softmax(attention_weights, dim=-1) over levels x points, rtdetr_decoder.py:126-128
    float mx = fmaxf(w0, w1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    w0 = lane < LP ? expf(w0 - mx) : 0.f;
    w1 = lane + 32 < LP ? expf(w1 - mx) : 0.f;
    float sum = w0 + w1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    w0 /= sum; w1 /= sum;
  }
  const long long vrow = p.value_ld;
  const float* vb = p.value + static_cast<long long>(b) * p.Lv * vrow + hd * 32 + lane;
  float acc = 0.f;
  for (int l = 0; l < p.L; ++l) {
    const int H = p.h[l], W = p.w[l];
    const float* vl = vb + static_cast<long long>(p.start[l]) * vrow;
    float rx = 0.f, ry = 0.f;
    if (p.fused) {
      const float* r = p.ref + ((static_cast<long long>(b) * p.Lq + q) * p.ref_levels + (p.ref_levels == 1 ? 0 : l)) * 2;
      rx = r[0]; ry = r[1];
    }
    for (int pt = 0; pt < p.P; ++pt) {
      const int i = l * p.P + pt;
      float lx = loc[i * 2 + 0], ly = loc[i * 2 + 1];
      if (p.fused) {            // reference_points + sampling_offsets / (W_l, H_l), rtdetr_decoder.py:141-152
        lx = rx + lx / static_cast<float>(W);
        ly = ry + ly / static_cast<float>(H);
      }
      const float wgt = __shfl_sync(0xffffffffu, i < 32 ? w0 : w1, i & 31);
      // grid = 2 loc - 1; align_corners=False: pixel = ((grid + 1) * size - 1) / 2
      const float gx = 2.f * lx - 1.f, gy = 2.f * ly - 1.f;
      const float ix = ((gx + 1.f) * static_cast<float>(W) - 1.f) * 0.5f;
      const float iy = ((gy + 1.f) * static_cast<float>(H) - 1.f) * 0.5f;
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
      const float tx = ix - fx, ty = iy - fy;
      // bilinear taps, zeros outside the map (padding_mode='zeros')
      const bool xin0 = x0 >= 0 && x0 < W, xin1 = x0 + 1 >= 0 && x0 + 1 < W;
      const bool yin0 = y0 >= 0 && y0 < H, yin1 = y0 + 1 >= 0 && y0 + 1 < H;
      float v = 0.f;
      if (yin0) {
        const float* row = vl + static_cast<long long>(y0) * W * vrow;
        if (xin0) v += (1.f - tx) * (1.f - ty) * row[static_cast<long long>(x0) * vrow];
        if (xin1) v += tx * (1.f - ty) * row[static_cast<long long>(x0 + 1) * vrow];
      }
      if (yin1) {
        const float* row = vl + static_cast<long long>(y0 + 1) * W * vrow;
        if (xin0) v += (1.f - tx) * ty * row[static_cast<long long>(x0) * vrow];
        if (xin1) v += tx * ty * row[static_cast<long long>(x0 + 1) * vrow];
      }
      acc = fmaf(wgt, v, acc);
    }
  }
  p.out[qh * 32 + lane] = acc;
}

// one CTA per image: score_i = max_c cls[i, c]; k rounds of block-wide arg-max (descending; ties -> lower index)
__global__ void __launch_bounds__(256)
topk_queries_kernel(const float* __restrict__ cls, int Lv, int C, int k, int32_t* __restrict__ idx_out,
                    float* __restrict__ val_out) {
  extern __shared__ float sc[];                 // [Lv] scores, then [8] warp winners
  __shared__ float wv[8];
  __shared__ int wi[8];
  const int b = blockIdx.x;
  const float* x = cls + static_cast<long long>(b) * Lv * C;
  for (int i = threadIdx.x; i < Lv; i += blockDim.x) {
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, x[static_cast<long long>(i) * C + c]);
    sc[i] = m;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = 0; r < k; ++r) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < Lv; i += blockDim.x) {
      const float v = sc[i];
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float v2 = __shfl_xor_sync(0xffffffffu, bv, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      if (v2 > bv || (v2 == bv && i2 < bi)) { bv = v2; bi = i2; }
    }
    if (lane == 0) { wv[warp] = bv; wi[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = wv[0];
      int i = wi[0];
      for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w)
        if (wv[w] > v || (wv[w] == v && wi[w] < i)) { v = wv[w]; i = wi[w]; }
      idx_out[static_cast<long long>(b) * k + r] = i;
      if (val_out) val_out[static_cast<long long>(b) * k + r] = v;
      if (i >= 0 && i < Lv) sc[i] = -INFINITY;   // taken
    }
    __syncthreads();
  }
}

// out[b, r, :] = src[b, idx[b, r], :]
__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int Lv, int k, int D,
                                   float* __restrict__ out) {
  const int b = blockIdx.y, r = blockIdx.x;
  const int i = idx[static_cast<long long>(b) * k + r];
  const float* s = src + (static_cast<long long>(b) * Lv + i) * D;
  float* o = out + (static_cast<long long>(b) * k + r) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) o[d] = s[d];
}

}  // namespace

std::string launch_ms_deform_attn(const float* value, const int* shapes_hw, int L, const float* loc, const float* attn,
                                  const float* ref, int ref_levels, int B, int Lq, int heads, int P, int fused,
                                  float* out, cudaStream_t s, long long value_ld, long long loc_ld, long long attn_ld) {
  if (B <= 0 || Lq <= 0) return "";
  if (L <= 0 || L > kMaxLevels || P <= 0 || L * P > kMaxLP) return "ms_deform_attn: levels x points outside [1, 64]";
  if (fused && (!ref || (ref_levels != 1 && ref_levels != L))) return "ms_deform_attn: fused mode needs reference points";
  DeformParams p{};
  p.value = value; p.loc = loc; p.attn = attn; p.ref = ref; p.out = out;
  p.B = B; p.Lq = Lq; p.heads = heads; p.L = L; p.P = P; p.ref_levels = ref_levels; p.fused = fused;
  p.value_ld = value_ld > 0 ? value_ld : static_cast<long long>(heads) * 32;
  p.loc_ld = loc_ld > 0 ? loc_ld : static_cast<long long>(heads) * L * P * 2;
  p.attn_ld = attn_ld > 0 ? attn_ld : static_cast<long long>(heads) * L * P;
  int start = 0;
  for (int l = 0; l < L; ++l) {
    p.h[l] = shapes_hw[2 * l]; p.w[l] = shapes_hw[2 * l + 1]; p.start[l] = start;
    if (p.h[l] <= 0 || p.w[l] <= 0) return "ms_deform_attn: bad level shape";
    start += p.h[l] * p.w[l];
  }
  p.Lv = start;
  const long long warps = static_cast<long long>(B) * Lq * heads;
  const unsigned blocks = static_cast<unsigned>((warps * 32 + 255) / 256);
  ProfScope ps(kFamAttention, s);
  ms_deform_attn_kernel<<<blocks, 256, 0, s>>>(p);
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_topk_queries(const float* cls, int B, int Lv, int C, int k, int32_t* idx, float* vals, cudaStream_t s) {
  if (B <= 0) return "";
  if (k <= 0 || k > Lv) return "topk_queries: k outside [1, Lv]";
  if (Lv > 11000) return "topk_queries: more than 11000 encoder tokens";
  ProfScope ps(kFamHeads, s);
  topk_queries_kernel<<<B, 256, static_cast<size_t>(Lv) * sizeof(float), s>>>(cls, Lv, C, k, idx, vals);
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_gather_rows(const float* src, const int32_t* idx, int B, int Lv, int k, int D, float* out, cudaStream_t s) {
  if (B <= 0 || k <= 0) return "";
  ProfScope ps(kFamElementwise, s);
  gather_rows_kernel<<<dim3(k, B), 128, 0, s>>>(src, idx, Lv, k, D, out);
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

}  // namespace spe
