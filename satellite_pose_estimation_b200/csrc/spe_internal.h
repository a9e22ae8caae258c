// Internal (non-ABI) declarations shared by the translation units of libspe.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace spe {

enum Dtype : int { kTF32 = 0, kBF16 = 1 };  // storage: fp32 (tf32 MMA) or bf16 (bf16 MMA); fp32 accumulate

inline size_t dtype_size(Dtype d) { return d == kTF32 ? 4 : 2; }

// One tensor-core GEMM launch:  out[m, n] = act( scale[n] * sum_k A[m,k] * Wt[n,k] + bias[n] + residual[m', n] )
// A is either a row-major matrix (mode 0) or an NHWC activation read through R*S shifted taps (mode 1: stride 1
// or 2 convolution, zero padding supplied by TMA out-of-bounds fill, stride by the TMA traversal stride).
struct GemmDesc {
  int mode = 0;                   // 0 matrix, 1 NHWC convolution, 2 7x7/stride-2 stem over the padded NHWC-Cp image
  // mode 0: A[M, K] row-major with leading dimension lda (elements)
  const void* A = nullptr;
  long long M = 0;
  int K = 0;
  int lda = 0;
  // mode 1: A = activation [NB, H, W, C]; K = R*S*C; M = NB*H*W
  int NB = 0, H = 0, W = 0, C = 0, R = 1, S = 1, pad = 0;   // H, W = input extent
  int conv_stride = 1;            // 1 or 2 (TMA traversal stride); output extent = (H + 2 pad - R) / stride + 1
  int c_ld = 0;                   // mode 1: elements between consecutive input pixels (0 = C): channels [0, C) of a wider tensor
  // weights Wt[N, K] row-major (K contiguous)
  const void* Wt = nullptr;
  int N = 0;
  // epilogue
  const float* scale = nullptr;   // [N] or null (=1)
  const float* bias = nullptr;    // [N] or null (=0)
  const void* residual = nullptr; // [*, res_ld] or null
  int res_ld = 0;
  int res_mod = 0;                // 0: residual row = output row, >0: row % res_mod (batch-broadcast addend)
  int res_f32 = 0;                // residual stored as fp32 even when the storage dtype is bf16
  int relu = 0;                   // 0 none, 1 ReLU; 3xTF32 kernels also: 2 SiLU, 3 GELU (erf form)
  void* out = nullptr;            // [M, out_ld] storage dtype
  int out_ld = 0;
  int round_out = 1;              // fp32 storage: round the result to TF32 (its consumer is a kind::tf32 MMA)
  int x3 = 0;                     // error-compensated 3xTF32: Wt is [N, 2K] = [W_hi | W_lo]; fp32-grade accuracy
  int out_f32 = 0;                // bf16 storage only: `out` is fp32 [M, out_ld] (values rounded to TF32)
  // mode 0 only: a SECOND operand source concatenated along K -- out = [A | A2] . Wt^T with Wt [N, K + K2].  A2 is
  // either a plain [M, K2] matrix (a2_stride = 1) or the 1x1 / stride-2 sampling of an NHWC activation
  // [a2_NB, a2_H, a2_W, K2] (a2_stride = 2: output pixel (h, w) reads input pixel (2h, 2w)).  This is how a residual
  // block's downsample branch rides in its conv3 GEMM (model.cu: Bottleneck::c3d).
  const void* A2 = nullptr;
  int K2 = 0, lda2 = 0;
  int a2_stride = 1, a2_NB = 0, a2_H = 0, a2_W = 0;
};

// returns empty string on success, else an error message
std::string launch_gemm(Dtype dt, const GemmDesc& d, int num_sms, cudaStream_t stream);

// ---- memory-bound helper kernels (elementwise.cu) ----
std::string launch_stem_im2col(Dtype dt, const float* nchw, int NB, int Hin, int Win, void* out /*[NB*Ho*Wo,192]*/,
                               cudaStream_t s);
std::string launch_stem_pad(Dtype dt, const float* nchw, int NB, int Hin, int Win, void* out /*[NB,H+6,W+6,Cp]*/,
                            cudaStream_t s);
std::string launch_im2col_nhwc(Dtype dt, const void* in, int NB, int H, int W, int C, int R, int S, int stride,
                               int pad, void* out, cudaStream_t s);
std::string launch_maxpool3x3s2(Dtype dt, const void* in, int NB, int H, int W, int C, void* out, cudaStream_t s);
std::string launch_upsample2x(Dtype dt, const void* in, int NB, int H, int W, int C, void* out, cudaStream_t s);
// out[b, p, 0:Cout] (row stride out_ld, storage dtype) = sum over the 3x3 taps inside the 2H x 2W map of the bilinear
// (align_corners=True) x2 upsampling of Y[b, :, :, tap * Cout + o]; Y fp32 [NB, H, W, 9 * Cout]
std::string launch_upsample_tapsum(Dtype dt, const float* Y, int NB, int H, int W, int Cout, void* out, int out_ld,
                                   cudaStream_t s);
// exact = 0: fp32 storage is TF32-rounded; 1: full fp32 result; 2: 3xTF32 operand layout [hi | lo | hi], row stride 768
std::string launch_layernorm(Dtype dt, const void* in, const float* gamma, const float* beta, long long rows,
                             int dim, void* out, cudaStream_t s, int exact = 0, const float* gamma2 = nullptr,
                             const float* beta2 = nullptr, void* out2 = nullptr,    // 2: a second LayerNorm of the result -> out2
                             float* out_f32 = nullptr);   // also the rows as fp32 [rows, 256] holding TF32 values

// ---- attention (attention.cu) ----
struct AttnDesc {
  const void* q; const void* k; const void* v;  // storage dtype; row (token) major, heads along columns
  int ldq, ldk, ldv;                            // row strides in elements
  long long bsq, bsk, bsv;                      // batch strides in elements
  void* out; int ldo; long long bso;
  int B, heads, Lq, Lk;                         // head_dim fixed at 32
  float scale;                                  // 1/sqrt(head_dim)
  int exact_out = 0;                            // fp32 storage: do not round the output to TF32
  int mixed = 0;                                // bf16 storage, tcgen05 path: q / k / v are fp32 (TF32 values), out is bf16
  int x3 = 0;                                   // fp32 storage: error-compensated products (register kernel only)
};
std::string launch_attention(Dtype dt, const AttnDesc& d, cudaStream_t s);
// tcgen05 / TMEM path (attention_tc.cu); launch_attention dispatches to it when supported
bool attention_tc_supported(Dtype dt, const AttnDesc& d);
std::string launch_attention_tc(const AttnDesc& d, cudaStream_t s);
// 2-D row-major tensor map (dim0 contiguous); SWIZZLE_128B, or SWIZZLE_128B_ATOM_32B (32-byte swizzle granules: the
// only layout the tensor core accepts for MN-major TF32 operands); `map` points at a CUtensorMap
std::string encode_tmap_2d(void* map, Dtype dt, const void* base, long long dim0, long long dim1,
                           long long stride1_bytes, int box0, int box1, bool swizzle_atom32 = false);

// ---- fused encoder feed-forward block (ffn_tc.cu): out = LayerNorm(X + relu(X W1^T + b1) W2^T + b2) ----
struct FfnDesc {
  const void* X = nullptr;      // [M, 256] fp32 (TF32 values)
  long long M = 0;
  const void* W1 = nullptr;     // [hidden, 256]
  const float* b1 = nullptr;
  const void* W2 = nullptr;     // [256, hidden]
  const float* b2 = nullptr;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  void* out = nullptr;          // may alias X (each tile is read completely before it is written)
  int out_mode = 0;             // 0: rounded to TF32, 1: exact fp32, 2: [M, 768] = [hi | lo | hi], 3: bf16 [M, 256]
  int hidden = 0;
};
bool ffn_fused_supported(Dtype dt, int d_model, int hidden);
std::string launch_ffn_fused(const FfnDesc& d, int num_sms, cudaStream_t stream);

// ---- heads (heads.cu) ----
std::string launch_head_final(Dtype dt, const void* hs, const void* h2, const void* s2, long long rows,
                              const float* Wc, const float* bc, const float* W3, const float* b3,
                              const float* Ws3, const float* bs3, float* logits, float* points, float* logsig,
                              cudaStream_t s);

// ---- SA (RT-DETR) decoder pieces (deform_attn.cu) ----
std::string launch_ms_deform_attn(const float* value, const int* shapes_hw, int L, const float* loc, const float* attn,
                                  const float* ref, int ref_levels, int B, int Lq, int heads, int P, int fused,
                                  float* out, cudaStream_t s, long long value_ld = 0, long long loc_ld = 0,
                                  long long attn_ld = 0);   // row strides in floats, 0 = dense reference layout
std::string launch_topk_queries(const float* cls, int B, int Lv, int C, int k, int32_t* idx, float* vals, cudaStream_t s);
std::string launch_gather_rows(const float* src, const int32_t* idx, int B, int Lv, int k, int D, float* out, cudaStream_t s);

// ---- SA (RT-DETR) predictor: everything that is not a GEMM / attention / LayerNorm (sa_kernels.cu) ----
// round: 1 = results rounded to TF32 (consumer is a plain kind::tf32 GEMM), 0 = full fp32 (consumer is a 3xTF32 GEMM)
std::string launch_sa_stem_im2col(const float* nchw, int NB, int H, int W, float* out /*[NB*H/2*W/2, 32]*/, int round,
                                  cudaStream_t s);
std::string launch_avgpool2x2(const float* in, int NB, int H, int W, int C, float* out, int round, cudaStream_t s);
// kind: 0 identity, 1 SiLU, 2 GELU (erf), 3 sigmoid; out[r, 0:C] = act(in[r, 0:C]) (+ add[r, 0:C]); round: to TF32
std::string launch_act_rows(const float* in, int in_ld, const float* add, int add_ld, float* out, int out_ld,
                            long long rows, int C, int kind, int round, cudaStream_t s);
std::string launch_upsample_nearest2x(const float* in, int in_ld, int NB, int H, int W, int C, float* out, int out_ld,
                                      cudaStream_t s);
std::string launch_bicubic_half(const float* in, int NB, int H, int W, int C, float* out, int out_ld, int round,
                                cudaStream_t s);
std::string launch_small_linear(const float* x, int ldx, long long rows, int K, const float* Wt, const float* b, int N,
                                float* out, int ldo, const float* addend, int add_mod, cudaStream_t s);
std::string launch_query_pos_hidden(const float* ref, const float* W0, const float* b0, int Hd, long long rows, float* out,
                                    cudaStream_t s);
std::string launch_sa_head(const float* tgt, const float* h2, const float* g2, const float* ref_in, long long rows,
                           const float* Wc, const float* bc, const float* Wb, const float* bb, const float* Ws,
                           const float* bs, float* logits, float* pts, float* logsig, float* ref_out, cudaStream_t s);
std::string launch_add(const float* a, const float* b, float* out, long long n, cudaStream_t s);

// ---- crop (crop.cu) ----
// stem_out (optional): write the predictor's stem input instead of the NCHW tensor -- zero-bordered NHWC [B, R+6, R+6]
// pixels of 16 bytes (3 channels + padding; fp32 rounded to TF32, or bf16), border already zero.  Only the staged kernel
// does that: *stem_written tells whether it ran (otherwise out_nchw was written as usual).
std::string launch_crop_resize_norm(const uint8_t* frames, int H, int W, long long pitch, long long frame_stride,
                                    const int32_t* boxes, int B, int R, float* out_nchw, cudaStream_t s,
                                    void* stem_out = nullptr, int stem_bf16 = 0, bool* stem_written = nullptr);

// ---- pnp (pnp.cu) ----
struct PnpDesc {
  const float* logits;   // [B,Q,12] raw class logits
  const float* points;   // [B,Q,2] normalised crop coordinates
  const float* logsig;   // [B,Q,2] or null
  const int32_t* boxes;  // [B,4] crop boxes x1,y1,x2,y2
  const float* boxes_f;  // or [B,4] fp32 x1,y1,width,height (eval path: unrounded box); takes precedence when set
  int B, Q;
  float reproj_thresh;
  const float* reproj_dev;   // optional [B]: per-image threshold (area-adaptive RANSAC threshold of the SA solver)
  int weighted;
  int reject;            // apply the self-assessment reject filter
  float reject_rms_px;   // filter thresholds
  float reject_sigma;
  double* quat;          // [B,4] wxyz
  double* tvec;          // [B,3]
  int32_t* assign;       // [B,11]
  int32_t* status;       // [B]
  float* probs;          // [B,Q,12] or null: softmax probabilities (PostProcess 'logits')
  float* points_px;      // [B,Q,2] or null: keypoints in original-image pixels
  float* sigmas;         // [B,Q,2] or null: exp(logsig)
  int32_t* inlier_mask;  // [B] bit i = label i used in the final refinement, or null
  int debug_timing;      // print per-phase cycle counts of the first images (SPE_PNP_TIMING)
  // ensemble form (Multi_Mean_PoseSolver): num_models > 0 -> logits / points are [num_models,B,Q,*], every foreground
  // query of every model is pooled per label (mean -> 3-sigma filter -> mean); `assign` then receives the number of
  // predictions each label's mean was taken over (0 = label absent)
  int num_models;
  // inputs are PostProcess OUTPUTS: `logits` holds class probabilities (no softmax is applied: scores are compared bit
  // for bit as the reference's solver sees them) and `points` original-image pixels (boxes must be (0,0,1,1));
  // sigma_px_scale = crop side for the reject filter's sigma criterion (0: criterion skipped)
  int post_processed;
  float sigma_px_scale;
  float* pooled_px;      // [B,11,2] or null: the pooled keypoints in original-image pixels (0 where absent)
};
std::string launch_assign_pnp(const PnpDesc& d, cudaStream_t s);
std::string launch_speed_score(const double* q_pr, const double* t_pr, const double* q_gt, const double* t_gt, int B,
                               double* s_t, double* s_q, cudaStream_t s);

#define SPE_CUDA_TRY(expr)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return std::string(#expr) + ": " + cudaGetErrorString(_e);                               \
  } while (0)

}  // namespace spe
