// Context, weight repacking and the forward schedule of the keypoint-set predictor.
//
// The reference's DETR.forward (RV/models/detr_speed.py:59-92) runs ~400 eager PyTorch ops per batch.  Here the
// whole forward is a fixed schedule of our own kernels over NHWC activations held in a reusable workspace:
//   stem      : im2col(7x7/s2) -> GEMM(+FrozenBN+ReLU) -> maxpool                 (torchvision resnet50 stem)
//   layer1..3 : per bottleneck 3 GEMMs (1x1, implicit 3x3, 1x1+residual+ReLU), FrozenBN folded into the epilogue
//   neck      : s8_latern 1x1, bilinear x2, s16_latern 3x3, output_conv 3x3       (RV/models/backbone.py:139-142)
//   encoder   : QKV GEMM (pos-embedding folded into a batch-broadcast addend), flash attention, out-proj+residual,
//               LayerNorm, FFN GEMM pair, LayerNorm                                (RV/models/transformer.py:154-167)
//   decoder   : cross-attention K/V of all layers in one GEMM, then per layer self-attn / cross-attn / FFN
//                                                                                  (RV/models/transformer.py:218-239)
//   heads     : class logits, keypoint MLP + sigmoid, optional log-sigma MLP      (RV/models/detr_speed.py:83-84)
//
// Identity used for the positional terms: (x + pos) W^T + b = x W^T + (pos W^T + b); the bracket is constant per
// token position, computed once at weight-load time in fp32 and added by the GEMM epilogue (row % tokens).
#include "spe_internal.h"
#include "profile.h"
#include "../../include/spe.h"

#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

namespace spe {

// ------------------------------------------------------------------------------------------------------------------
// small device helpers used at weight-load time
// ------------------------------------------------------------------------------------------------------------------
namespace {

__global__ void f32_to_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(in[i]));
    out[i] = __uint_as_float(r);
  }
}
// [N,K] fp32 -> [N,2K] = [rna(W) | rna(W - rna(W))] for the error-compensated 3xTF32 GEMM
__global__ void f32_split_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int K) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(N) * K) return;
  const int n = static_cast<int>(i / K), k = static_cast<int>(i % K);
  const float w = in[i];
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(w));
  const float hi = __uint_as_float(r);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(w - hi));
  out[static_cast<long long>(n) * 2 * K + k] = hi;
  out[static_cast<long long>(n) * 2 * K + K + k] = __uint_as_float(r);
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}
// out[t, ncol0 + n] = sum_k X[t, k] * W[n, k] + b[n]   (fp32 FMA; setup only)
__global__ void addend_kernel(const float* __restrict__ X, int T, int K, const float* __restrict__ W,
                              const float* __restrict__ b, int N, float* __restrict__ out, int out_ld, int ncol0) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (n >= N || t >= T) return;
  float acc = 0.f;
  if (X != nullptr) {
    const float* x = X + static_cast<long long>(t) * K;
    const float* w = W + static_cast<long long>(n) * K;
    for (int k = 0; k < K; ++k) acc = fmaf(x[k], w[k], acc);
  }
  out[static_cast<long long>(t) * out_ld + ncol0 + n] = acc + (b ? b[n] : 0.f);
}

// ---- calibration (spe_calibrate): per-column sums of a GEMM's A operand, and the bias correction they imply ----
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// what the tensor core makes of a stored operand: kind::tf32 ignores the low 13 mantissa bits (truncation); bf16 is exact
__device__ __forceinline__ float as_mma_operand(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ float as_mma_operand(__nv_bfloat16 v) { return __bfloat162float(v); }

// part[blk][c] = sum over the block's rows of A[r, c], part_mma[blk][c] = the same over the values the tensor core
// reads (A [M, C], row stride lda elements).  No atomics: the partial sums are added up in a fixed order by
// colsum_reduce_kernel, so a calibration is bit-reproducible.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ A, long long M, int C, long long lda, float* __restrict__ part,
                              float* __restrict__ part_mma) {
  const long long rows_per_block = (M + gridDim.x - 1) / gridDim.x;
  const long long r0 = rows_per_block * blockIdx.x;
  const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f, acc_t = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const T v = A[r * lda + c];
      acc += to_f32(v);
      acc_t += as_mma_operand(v);
    }
    part[static_cast<long long>(blockIdx.x) * C + c] = acc;
    part_mma[static_cast<long long>(blockIdx.x) * C + c] = acc_t;
  }
}
__global__ void colsum_reduce_kernel(const float* __restrict__ part, const float* __restrict__ part_mma, int nblk, int C,
                                     float* __restrict__ sum, float* __restrict__ sum_mma) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int i = 0; i < nblk; ++i) { a += part[static_cast<long long>(i) * C + c]; b += part_mma[static_cast<long long>(i) * C + c]; }
  sum[c] = a;
  sum_mma[c] = b;
}
// channel sums of an NCHW fp32 image batch: sum[c] += sum over (n, h, w); C channels, HW pixels per plane
__global__ void nchw_chansum_kernel(const float* __restrict__ x, int NB, int C, long long HW, float* __restrict__ sum) {
  // one block per channel, fixed summation order (bit-reproducible): thread t strides over the pixels of every image
  const int c = blockIdx.x;
  __shared__ float red[256];
  float acc = 0.f;
  for (int n = 0; n < NB; ++n) {
    const float* p = x + (static_cast<long long>(n) * C + c) * HW;
    for (long long i = threadIdx.x; i < HW; i += blockDim.x) acc += p[i];
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) sum[c] = red[0];
}
// One warp per output channel n: delta = scale[n] * sum_k (W32[n,k] * mean[k % C] - W[n,k] * mean_mma[k % C]) -- what
// the product of the stored (rounded) weights and the operand as the tensor core reads it (fp32 activations that were
// left unrounded are truncated to TF32) loses on the mean activation -- folded into the layer's bias (or, for the
// projections whose bias lives in a batch-broadcast addend, into every row of that addend).  `applied` remembers the
// correction so that a second calibration replaces the first instead of stacking on it.
template <typename T>
__global__ void bias_correction_kernel(const float* __restrict__ w32, const T* __restrict__ w, const float* __restrict__ colsum,
                                       const float* __restrict__ colsum_mma,
                                       float inv_rows, int N, int K, int C, const float* __restrict__ scale,
                                       float* __restrict__ bias, float* __restrict__ applied, float* __restrict__ addend,
                                       int addend_rows, int addend_ld) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* a = w32 + static_cast<long long>(n) * K;
  const T* b = w + static_cast<long long>(n) * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    const int c = k % C;
    acc += (a[k] * colsum[c] - to_f32(b[k]) * colsum_mma[c]) * inv_rows;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const float delta = (scale ? scale[n] : 1.0f) * acc;
  const float d = delta - applied[n];
  __syncwarp();
  if (bias != nullptr) {
    if (lane == 0) bias[n] += d;
  } else if (addend != nullptr) {
    for (int t = lane; t < addend_rows; t += 32) addend[static_cast<long long>(t) * addend_ld + n] += d;
  }
  __syncwarp();
  if (lane == 0) applied[n] = delta;
}

// a[i] *= f, b[i] *= f (column sums of a second operand source taken over a different number of rows)
__global__ void scale_pair_kernel(float* __restrict__ a, float* __restrict__ b, int n, float f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { a[i] *= f; b[i] *= f; }
}

// round-to-nearest (ties away) fp32 -> TF32, the host twin of cvt.rna.tf32.f32
inline float host_rna_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

struct HostTensor {
  const float* data;
  std::vector<long long> shape;
  long long numel() const {
    long long n = 1;
    for (long long s : shape) n *= s;
    return n;
  }
};

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------------------------
struct GemmW {          // one GEMM's parameters on the device
  void* w = nullptr;    // [N, K] storage dtype
  float* scale = nullptr;
  float* bias = nullptr;
  int N = 0, K = 0;
  int x3 = 0;           // weights stored as [N, 2K] = [W_hi | W_lo] (3xTF32)
  // calibration (spe_calibrate): the unrounded fp32 weights in the stored layout, the correction currently folded
  // into `bias`, and -- for projections without a bias vector -- the batch-broadcast addend that takes it instead
  float* w32 = nullptr;
  float* applied = nullptr;
  float* addend = nullptr;
  int addend_rows = 0, addend_ld = 0;
};

struct Bottleneck {
  GemmW c1, c2, c3, down;
  // First block of a layer: out = relu(bn3(conv3(t)) + bn_d(down(x))) is ONE GEMM over the concatenated operand
  // [t | x (sampled at the block's stride)] with weights [s3 W3 | s_d W_d] (FrozenBN scales folded into the rows) and
  // bias b3 + b_d: the downsample output (205 MB at layer1, B = 64) is never written or read back, one launch less.
  GemmW c3d;
  int inplanes = 0, planes = 0, stride = 1;
  bool has_down = false;
};

struct EncLayer {
  GemmW qkv, out, ff1, ff2;
  // bf16 storage: TF32 copies of the feed-forward weights (own, uncalibrated bias vectors) for the fused feed-forward
  // kernel, which is built for fp32 / TF32 operands (spe_ctx::mixed_ffn)
  GemmW ff1_32, ff2_32;
  float* addend = nullptr;  // [tokens, 768]
  float *n1g, *n1b, *n2g, *n2b;
};

struct DecLayer {
  GemmW sa_qkv, sa_out, ca_q, ca_out, ff1, ff2;
  float* sa_addend = nullptr;  // [Q, 768]
  float* ca_q_addend = nullptr;  // [Q, 256]
  float *n1g, *n1b, *n2g, *n2b, *n3g, *n3b;
};

}  // namespace spe

using namespace spe;

struct GraphKey {
  int B;
  const void *images, *logits, *points, *logsig, *aux_l, *aux_p;
  int parts = 3, kv_slot = 0;   // 1 = trunk (+ decoder K/V), 2 = decoder + heads, 3 = both
  bool operator==(const GraphKey& o) const {
    return parts == o.parts && kv_slot == o.kv_slot && B == o.B && images == o.images && logits == o.logits && points == o.points && logsig == o.logsig &&
           aux_l == o.aux_l && aux_p == o.aux_p;
  }
};
struct GraphEntry {
  GraphKey key{};
  cudaGraphExec_t exec = nullptr;
  bool failed = false;          // capture / instantiation failed for THIS key: it keeps running eagerly
  long long launches[kNumFamilies] = {0, 0, 0, 0, 0, 0};
};

namespace spe { struct SaModel; }   // SA (RT-DETR) predictor: weights + workspace, sa_model.inl

struct spe_ctx {
  spe_config cfg{};
  spe::SaModel* sa = nullptr;   // cfg.backbone == 2
  int device = 0;
  int num_sms = 148;
  Dtype dt = kTF32;
  std::string err;
  bool weights_loaded = false;

  int featH = 0, tokens = 0, featC = 0;

  std::vector<void*> allocs;  // everything cudaMalloc'ed, freed in destroy

  // weights
  GemmW stem;                  // im2col form [64, 192]
  GemmW stem2;                 // TMA-window form [64, 7 rows x 8 taps x Cp]
  bool stem_windowed = true;   // cleared if the overlapping-window tensor map is refused by the driver
  // pipeline slots: the crop kernel writes the padded stem input (SP) itself and the schedule skips stem_pad_kernel
  // (one launch and a 38 MB write + read per batch of 64 less; SPE_CROP_STEM=0 restores the NCHW hand-over)
  bool crop_writes_stem = getenv("SPE_CROP_STEM") ? atoi(getenv("SPE_CROP_STEM")) != 0 : true;
  bool stem_prefilled = false; // set for the duration of one forward_schedule call (parts bit 2)
  void* SP = nullptr;          // padded NHWC-Cp stem input
  std::vector<Bottleneck> blocks;
  GemmW s8_lat, s16_lat, out_conv, input_proj;
  // output_conv (3x3, 512 -> 512, + bias) is followed by input_proj (1x1, 512 -> 256, + bias) with nothing in between
  // (RV/models/backbone.py:142, RV/models/detr_speed.py:81): one 3x3 convolution 512 -> 256 with W' = W_ip . W_oc and
  // b' = W_ip . b_oc + b_ip computes the same thing with half the multiply-adds of output_conv and no input_proj at all
  // (-2.06 of 26.57 GFLOP per image, one rounding of the 512-channel feature map less).  SPE_FOLD_NECK=0 keeps the two.
  GemmW out_conv_ip;
  // s16_latern applied BEFORE the bilinear x2 upsampling (launch_upsample_tapsum, elementwise.cu): its nine tap matrices
  // as one [9 * 256, 1024] GEMM on the 14 x 14 map, a quarter of the convolution's multiply-adds (-2.77 GFLOP per image),
  // no 1024-channel upsampled tensor (205 MB at B = 64)
  GemmW s16_lr;
  float* YLR = nullptr;        // [B, 14, 14, 9 * 256] fp32
  // stem + max-pool + layer1 run in chunks of `head_chunk` images (0 = whole batch at once): at 112 x 112 x 64 and
  // 56 x 56 x 256 fp32 these layers are HBM-bound (205 MB per tensor at B = 64, against 126 MB of L2); a chunk of 16
  // images keeps producer -> consumer hand-overs inside L2, and because every chunk reuses the SAME scratch addresses
  // the dirty lines of the stem output are overwritten in L2 instead of ever being written back.  SPE_HEAD_CHUNK.
  // The first decoder layer starts from tgt = 0 (RV/models/transformer.py:59, :111): its self-attention block sees
  // only the learned query embeddings, so norm1's output and the cross-attention query projection of layer 0 are the
  // same [Q, 256] matrices for every image.  They are computed once (after a weight load and after every calibration,
  // with the very kernels the per-batch path would run on one image) and broadcast: five launches per batch less.
  void* dec0_tgt = nullptr;    // [Q, 256] storage dtype: tgt after norm1 of decoder layer 0
  void* dec0_q = nullptr;      // [Q, 256] storage dtype: cross-attention query projection of decoder layer 0
  bool dec0_valid = false;
  bool dec0_fold = getenv("SPE_DEC0_FOLD") ? atoi(getenv("SPE_DEC0_FOLD")) != 0 : true;
  bool fuse_down = getenv("SPE_FUSE_DOWN") ? atoi(getenv("SPE_FUSE_DOWN")) != 0 : true;   // Bottleneck::c3d
  void* L1OUT = nullptr;       // [B, 56, 56, 256] layer1 output of all chunks
  int head_chunk = getenv("SPE_HEAD_CHUNK") ? atoi(getenv("SPE_HEAD_CHUNK")) : 0;
  bool fold_neck = getenv("SPE_FOLD_NECK") ? atoi(getenv("SPE_FOLD_NECK")) != 0 : true;
  std::vector<EncLayer> enc;
  std::vector<DecLayer> dec;
  GemmW ca_kv_all;             // [L*512, 256] plain storage-dtype weights
  GemmW ca_kv_x3;              // fp32 storage: [L*512, 768] = [W_hi | W_hi | W_lo], the hoisted 3xTF32 form
  bool kv_split3 = false;      // ca_kv_x3 is loaded
  void* XS = nullptr;          // last encoder output as [hi | lo | hi], [B*tokens, 768]
  float* ca_kv_addend = nullptr;  // [tokens, L*512]
  float *dn_g = nullptr, *dn_b = nullptr;
  GemmW pt0, pt1, sg0, sg1;
  float *cls_w = nullptr, *cls_b = nullptr, *pt2_w = nullptr, *pt2_b = nullptr, *sg2_w = nullptr, *sg2_b = nullptr;

  // workspace (storage dtype unless noted)
  void *S0 = nullptr, *S1 = nullptr, *P0 = nullptr, *P1 = nullptr, *T1 = nullptr, *T2 = nullptr, *DS = nullptr,
       *COL = nullptr, *L2OUT = nullptr, *L3OUT = nullptr, *UP = nullptr, *CAT = nullptr, *FEAT = nullptr,
       *X = nullptr, *X2 = nullptr, *QKV = nullptr, *ATT = nullptr, *HID = nullptr, *KV = nullptr;
  bool mixed_attention = getenv("SPE_MIXED_ATTN") ? atoi(getenv("SPE_MIXED_ATTN")) != 0 : true;
  // bf16 storage: norm1 also writes its rows as fp32 (TF32 values) and the encoder's feed-forward block + norm2 run on
  // the fused tcgen05 kernel (hidden activation in tensor memory) with a bf16 output, instead of two bf16 GEMMs with the
  // 2048-wide hidden tensor in HBM and a LayerNorm pass (SPE_MIXED_FFN=0: the bf16 GEMM pair)
  bool mixed_ffn = getenv("SPE_MIXED_FFN") ? atoi(getenv("SPE_MIXED_FFN")) != 0 : true;
  float* XF = nullptr;         // [B * tokens, 256] fp32: norm1 output for the fused feed-forward kernel (bf16 storage)
  void *TGT = nullptr, *TGT2 = nullptr, *DQKV = nullptr, *DQ = nullptr, *DATT = nullptr, *DHID = nullptr,
       *HS = nullptr, *H1 = nullptr, *H2 = nullptr, *G1 = nullptr, *G2 = nullptr;
  float *logits_all = nullptr, *points_all = nullptr;
  // Activation sets: every pipeline slot owns a full copy of the workspace above, so whole batches can be in flight
  // next to each other.  The named pointers always hold the set selected by use_workspace() (enqueue is host-serial;
  // launched kernels keep the pointers they were given).
  struct WsField { void** field; long long bytes; };
  std::vector<WsField> ws_fields;
  std::vector<std::vector<void*>> ws_sets;
  int ws_current = 0;

  // pipeline buffers for spe_run_batch_host
  uint8_t* frames_dev = nullptr;
  long long frames_cap = 0;
  int32_t* boxes_dev = nullptr;
  float* images_dev = nullptr;
  float *p_logits = nullptr, *p_points = nullptr, *p_logsig = nullptr;
  double *p_quat = nullptr, *p_tvec = nullptr;
  int32_t *p_assign = nullptr, *p_status = nullptr;
  const float *ov_logits = nullptr, *ov_points = nullptr;   // bench hook, see spe_debug_set_pnp_override
  const int32_t* ov_boxes = nullptr;
  long long last_h2d = 0;            // bytes uploaded by the last spe_run_batch_host

  // calibration state (spe_calibrate)
  bool calibrating = false, calibrated = false;
  float* colsum = nullptr;           // [2][4096] column sums of a layer's input as stored / as the MMA reads it, then
                                     // [2][256][4096] per-block partial sums (calibration scratch)
  // fp32 storage: tensors that are both a GEMM operand and a residual (block outputs of the backbone, the encoder
  // stream) stay UNROUNDED in HBM.  The tensor core truncates them on the fly -- same noise as rounding them, plus a
  // small bias that spe_calibrate measures and folds into the biases -- while the skip path adds the exact values, so
  // rounding noise no longer accumulates along the residual stream.  MEASURED AND LEFT OFF (SPE_EXACT_STREAM=1 enables
  // it): the per-element error of the encoder memory drops (5.6e-4 -> 4.2e-4 relative) but the keypoints get WORSE
  // (0.16 -> 0.21 px rms at S = 1748): truncation shrinks every operand by ~3.5e-4, a coherent gain error on every
  // GEMM branch that a bias cannot absorb.
  // error-compensated 3xTF32 for the decoder + head GEMMs / for the cross-attention K/V projection (fp32 storage only)
  bool dec_x3 = getenv("SPE_DEC_X3") ? atoi(getenv("SPE_DEC_X3")) != 0 : true;
  // ... and separately for the decoder's feed-forward pair (linear1 / linear2: 0.28 of the decoder's GEMM time).  With
  // plain TF32 there, norm2 writes a second, TF32-ROUNDED copy of its output as linear1's operand (the exact copy
  // stays the residual), so the tensor core's truncation adds no bias.  SPE_DEC_FFN_X3.
  bool dec_ffn_x3 = getenv("SPE_DEC_FFN_X3") ? atoi(getenv("SPE_DEC_FFN_X3")) != 0 : true;
  // K/V projection: -1 (default) = 3xTF32 until the context is calibrated, plain TF32 afterwards -- with the rounding
  // bias folded into the addend the plain product is as accurate as the compensated one (measured: 0.12 vs 0.14 px at
  // S = 1748, whole-chain keypoints 0.27 vs 0.49 px) and a third of the tensor work (213 -> ~75 us at B = 64);
  // uncalibrated it is not (0.61 vs 0.47 px).  1 = always 3xTF32, 0 = never.
  int kv_x3 = getenv("SPE_KV_X3") ? atoi(getenv("SPE_KV_X3")) : -1;
  bool exact_stream = getenv("SPE_EXACT_STREAM") ? atoi(getenv("SPE_EXACT_STREAM")) != 0 : false;

  // Stream capture is illegal on the legacy default stream (and on cudaStreamPerThread a capture would swallow
  // unrelated work): calls that arrive on one of those run on this private non-blocking stream instead, ordered against
  // the caller's stream by a pair of events, so they too replay their captured graph.
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;

  // forward schedule
  bool use_graphs = true;
  int sub_batch = 0;                 // 0 = automatic (L2-sized chunks)
  std::vector<GraphEntry> graphs;

  // debug taps
  bool taps_enabled = false;
  struct Tap { void* buf; long long bytes; };
  std::map<std::string, Tap> taps;
};

namespace spe {

void pipeline_release(spe_ctx* ctx);   // api.cu
void jpeg_release(spe_ctx* ctx);       // jpeg.cu
bool pipeline_busy(spe_ctx* ctx);      // api.cu

static std::string g_last_error;

static int fail(spe_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  g_last_error = msg;
  return code;
}

template <typename T>
static std::string dmalloc(spe_ctx* ctx, T** p, long long count) {
  void* q = nullptr;
  SPE_CUDA_TRY(cudaMalloc(&q, static_cast<size_t>(count > 0 ? count : 1) * sizeof(T)));
  ctx->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return "";
}
static std::string dmalloc_bytes(spe_ctx* ctx, void** p, long long bytes) {
  void* q = nullptr;
  SPE_CUDA_TRY(cudaMalloc(&q, static_cast<size_t>(bytes > 0 ? bytes : 16)));
  ctx->allocs.push_back(q);
  *p = q;
  return "";
}

#define TRY_S(expr)                         \
  do {                                      \
    std::string _s = (expr);                \
    if (!_s.empty()) return _s;             \
  } while (0)

// upload host fp32 -> device fp32
static std::string upload_f32(spe_ctx* ctx, const float* host, long long n, float** out) {
  TRY_S(dmalloc(ctx, out, n));
  SPE_CUDA_TRY(cudaMemcpy(*out, host, static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice));
  return "";
}

// upload a host fp32 [N,K] matrix as GEMM weights in the storage dtype (tf32-rounded fp32 or bf16)
static std::string upload_gemm_w(spe_ctx* ctx, const std::vector<float>& host, int N, int K, GemmW* g,
                                 bool x3 = false, bool keep32 = true) {
  float* tmp = nullptr;
  const long long n = static_cast<long long>(N) * K;
  x3 = x3 && ctx->dt == kTF32;   // bf16 storage has its own (looser) accuracy contract
  SPE_CUDA_TRY(cudaMalloc(&tmp, static_cast<size_t>(n) * sizeof(float)));
  cudaError_t e = cudaMemcpy(tmp, host.data(), static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(tmp); return std::string("upload weights: ") + cudaGetErrorString(e); }
  std::string s = dmalloc_bytes(ctx, &g->w, (x3 ? 2 : 1) * n * static_cast<long long>(dtype_size(ctx->dt)));
  if (!s.empty()) { cudaFree(tmp); return s; }
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  g->x3 = x3 ? 1 : 0;
  if (x3) f32_split_tf32_kernel<<<blocks, 256>>>(tmp, reinterpret_cast<float*>(g->w), N, K);
  else if (ctx->dt == kTF32) f32_to_tf32_kernel<<<blocks, 256>>>(tmp, reinterpret_cast<float*>(g->w), n);
  else f32_to_bf16_kernel<<<blocks, 256>>>(tmp, reinterpret_cast<__nv_bfloat16*>(g->w), n);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(tmp); return std::string("convert weights: ") + cudaGetErrorString(e); }
  g->N = N;
  g->K = K;
  if (keep32 && !x3) {
    // the unrounded weights stay on the device: spe_calibrate folds (W - round(W)) . mean(activation) into the bias
    ctx->allocs.push_back(tmp);
    g->w32 = tmp;
    TRY_S(dmalloc(ctx, &g->applied, N));
    SPE_CUDA_TRY(cudaMemset(g->applied, 0, sizeof(float) * N));
  } else {
    cudaFree(tmp);
  }
  return "";
}

// host fp32 [N, K] -> device fp32 rounded to TF32, whatever the storage dtype of the ctx
static std::string upload_tf32_copy(spe_ctx* ctx, const float* host, int N, int K, GemmW* g) {
  const long long n = static_cast<long long>(N) * K;
  float* tmp = nullptr;
  SPE_CUDA_TRY(cudaMalloc(&tmp, static_cast<size_t>(n) * sizeof(float)));
  cudaError_t e = cudaMemcpy(tmp, host, static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(tmp); return std::string("upload weights: ") + cudaGetErrorString(e); }
  std::string s = dmalloc_bytes(ctx, &g->w, n * 4);
  if (!s.empty()) { cudaFree(tmp); return s; }
  f32_to_tf32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(tmp, reinterpret_cast<float*>(g->w), n);
  e = cudaDeviceSynchronize();
  cudaFree(tmp);
  if (e != cudaSuccess) return std::string("convert weights: ") + cudaGetErrorString(e);
  g->N = N; g->K = K; g->x3 = 0;
  return "";
}

struct WeightSource {
  std::map<std::string, HostTensor> t;
  std::string missing;
  const HostTensor* get(const std::string& name, std::initializer_list<long long> shape) {
    auto it = t.find(name);
    if (it == t.end()) { if (missing.empty()) missing = "missing tensor '" + name + "'"; return nullptr; }
    std::vector<long long> want(shape);
    if (it->second.shape != want) {
      if (missing.empty()) {
        missing = "tensor '" + name + "' has shape [";
        for (size_t i = 0; i < it->second.shape.size(); ++i) missing += (i ? "," : "") + std::to_string(it->second.shape[i]);
        missing += "], expected [";
        for (size_t i = 0; i < want.size(); ++i) missing += (i ? "," : "") + std::to_string(want[i]);
        missing += "]";
      }
      return nullptr;
    }
    return &it->second;
  }
};

// conv weight (Cout, Cin, R, S) -> [Cout][(r*S + s)*Cin + c], K padded with zeros to Kpad
static std::vector<float> repack_conv(const HostTensor& w, int Kpad = 0) {
  const int Cout = static_cast<int>(w.shape[0]), Cin = static_cast<int>(w.shape[1]);
  const int R = static_cast<int>(w.shape[2]), S = static_cast<int>(w.shape[3]);
  const int K = R * S * Cin;
  const int Kp = Kpad > 0 ? Kpad : K;
  std::vector<float> out(static_cast<size_t>(Cout) * Kp, 0.f);
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s)
          out[static_cast<size_t>(o) * Kp + (r * S + s) * Cin + c] =
              w.data[((static_cast<size_t>(o) * Cin + c) * R + r) * S + s];
  return out;
}

// FrozenBatchNorm2d fold, same fp32 op order as RV/models/backbone.py:44-54
static std::string load_bn(spe_ctx* ctx, WeightSource& ws, const std::string& prefix, int C, GemmW* g) {
  const HostTensor* w = ws.get(prefix + ".weight", {C});
  const HostTensor* b = ws.get(prefix + ".bias", {C});
  const HostTensor* rm = ws.get(prefix + ".running_mean", {C});
  const HostTensor* rv = ws.get(prefix + ".running_var", {C});
  if (!w || !b || !rm || !rv) return ws.missing;
  std::vector<float> scale(C), bias(C);
  for (int i = 0; i < C; ++i) {
    const float sc = w->data[i] * (1.0f / sqrtf(rv->data[i] + 1e-5f));
    scale[i] = sc;
    bias[i] = b->data[i] - rm->data[i] * sc;
  }
  TRY_S(upload_f32(ctx, scale.data(), C, &g->scale));
  TRY_S(upload_f32(ctx, bias.data(), C, &g->bias));
  return "";
}

// FrozenBatchNorm2d scale / bias on the host (same fp32 op order as load_bn)
static bool bn_fold_host(WeightSource& ws, const std::string& prefix, int C, std::vector<float>* scale, std::vector<float>* bias) {
  const HostTensor* w = ws.get(prefix + ".weight", {C});
  const HostTensor* b = ws.get(prefix + ".bias", {C});
  const HostTensor* rm = ws.get(prefix + ".running_mean", {C});
  const HostTensor* rv = ws.get(prefix + ".running_var", {C});
  if (!w || !b || !rm || !rv) return false;
  scale->resize(C); bias->resize(C);
  for (int i = 0; i < C; ++i) {
    const float sc = w->data[i] * (1.0f / sqrtf(rv->data[i] + 1e-5f));
    (*scale)[i] = sc;
    (*bias)[i] = b->data[i] - rm->data[i] * sc;
  }
  return true;
}

static std::string upload_gemm_w(spe_ctx* ctx, const std::vector<float>& host, int N, int K, GemmW* g, bool x3, bool keep32);
// conv3 + bn3 and downsample.0 + downsample.1 of residual block `p` as one [N, planes + inplanes] weight matrix
// (Bottleneck::c3d): rows scaled by the two FrozenBN scales, biases added
static std::string load_fused_down(spe_ctx* ctx, WeightSource& ws, const std::string& p, int planes, int inplanes, GemmW* g) {
  const int N = planes * 4, K = planes + inplanes;
  const HostTensor* w3 = ws.get(p + ".conv3.weight", {N, planes, 1, 1});
  const HostTensor* wd = ws.get(p + ".downsample.0.weight", {N, inplanes, 1, 1});
  std::vector<float> s3, b3, sd, bd;
  if (!w3 || !wd || !bn_fold_host(ws, p + ".bn3", N, &s3, &b3) || !bn_fold_host(ws, p + ".downsample.1", N, &sd, &bd))
    return ws.missing;
  std::vector<float> w(static_cast<size_t>(N) * K), bias(N);
  for (int n = 0; n < N; ++n) {
    float* row = w.data() + static_cast<size_t>(n) * K;
    for (int k = 0; k < planes; ++k) row[k] = s3[n] * w3->data[static_cast<size_t>(n) * planes + k];
    for (int k = 0; k < inplanes; ++k) row[planes + k] = sd[n] * wd->data[static_cast<size_t>(n) * inplanes + k];
    bias[n] = b3[n] + bd[n];
  }
  TRY_S(upload_gemm_w(ctx, w, N, K, g, false, true));
  TRY_S(upload_f32(ctx, bias.data(), N, &g->bias));
  return "";
}

static std::string load_conv_bn(spe_ctx* ctx, WeightSource& ws, const std::string& conv, const std::string& bn,
                                int Cout, int Cin, int R, GemmW* g, int Kpad = 0, bool x3 = false) {
  const HostTensor* w = ws.get(conv + ".weight", {Cout, Cin, R, R});
  if (!w) return ws.missing;
  TRY_S(upload_gemm_w(ctx, repack_conv(*w, Kpad), Cout, Kpad > 0 ? Kpad : R * R * Cin, g, x3));
  if (!bn.empty()) TRY_S(load_bn(ctx, ws, bn, Cout, g));
  return "";
}

static std::string load_linear(spe_ctx* ctx, WeightSource& ws, const std::string& p, int out_f, int in_f, GemmW* g,
                               bool x3 = false) {
  const HostTensor* w = ws.get(p + ".weight", {out_f, in_f});
  const HostTensor* b = ws.get(p + ".bias", {out_f});
  if (!w || !b) return ws.missing;
  TRY_S(upload_gemm_w(ctx, std::vector<float>(w->data, w->data + w->numel()), out_f, in_f, g, x3));
  TRY_S(upload_f32(ctx, b->data, out_f, &g->bias));
  return "";
}

static std::string load_vec(spe_ctx* ctx, WeightSource& ws, const std::string& name, int n, float** out) {
  const HostTensor* t = ws.get(name, {n});
  if (!t) return ws.missing;
  return upload_f32(ctx, t->data, n, out);
}

// PositionEmbeddingSine with an all-False mask, normalize=True (RV/models/position_encoding.py:30-53)
static std::vector<float> make_pos(int H, int W, int E) {
  const int npf = E / 2;
  std::vector<float> pos(static_cast<size_t>(H) * W * E);
  const float two_pi = 2.0f * static_cast<float>(M_PI);
  for (int i = 0; i < H; ++i)
    for (int j = 0; j < W; ++j) {
      const float ye = static_cast<float>(i + 1) / (static_cast<float>(H) + 1e-6f) * two_pi;
      const float xe = static_cast<float>(j + 1) / (static_cast<float>(W) + 1e-6f) * two_pi;
      float* p = pos.data() + (static_cast<size_t>(i) * W + j) * E;
      for (int k = 0; k < npf; ++k) {
        const float dim_t = powf(10000.0f, 2.0f * static_cast<float>(k / 2) / static_cast<float>(npf));
        const float vy = ye / dim_t, vx = xe / dim_t;
        p[k] = (k % 2 == 0) ? sinf(vy) : cosf(vy);
        p[npf + k] = (k % 2 == 0) ? sinf(vx) : cosf(vx);
      }
    }
  return pos;
}

// addend[:, col0 : col0+N] = X W^T + b  (X may be null -> bias only); W/b are host slices
static std::string make_addend(spe_ctx* ctx, const float* X_dev, int T, int K, const float* W_host,
                               const float* b_host, int N, float* out_dev, int out_ld, int col0) {
  float *Wd = nullptr, *bd = nullptr;
  SPE_CUDA_TRY(cudaMalloc(&Wd, static_cast<size_t>(N) * K * sizeof(float)));
  SPE_CUDA_TRY(cudaMalloc(&bd, static_cast<size_t>(N) * sizeof(float)));
  cudaMemcpy(Wd, W_host, static_cast<size_t>(N) * K * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(bd, b_host, static_cast<size_t>(N) * sizeof(float), cudaMemcpyHostToDevice);
  dim3 grid((N + 127) / 128, T);
  addend_kernel<<<grid, 128>>>(X_dev, T, K, Wd, bd, N, out_dev, out_ld, col0);
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(Wd);
  cudaFree(bd);
  if (e != cudaSuccess) return std::string("addend: ") + cudaGetErrorString(e);
  return "";
}

static std::string load_mha_self(spe_ctx* ctx, WeightSource& ws, const std::string& p, const float* posX_dev, int T,
                                 GemmW* qkv, GemmW* out, float** addend, bool x3 = false) {
  const int E = 256;
  const HostTensor* w = ws.get(p + ".in_proj_weight", {3 * E, E});
  const HostTensor* b = ws.get(p + ".in_proj_bias", {3 * E});
  if (!w || !b) return ws.missing;
  TRY_S(upload_gemm_w(ctx, std::vector<float>(w->data, w->data + w->numel()), 3 * E, E, qkv, x3));
  TRY_S(dmalloc(ctx, addend, static_cast<long long>(T) * 3 * E));
  // q and k see (x + pos); v sees x only
  TRY_S(make_addend(ctx, posX_dev, T, E, w->data, b->data, 2 * E, *addend, 3 * E, 0));
  TRY_S(make_addend(ctx, nullptr, T, E, w->data + 2 * E * E, b->data + 2 * E, E, *addend, 3 * E, 2 * E));
  qkv->addend = *addend; qkv->addend_rows = T; qkv->addend_ld = 3 * E;
  TRY_S(load_linear(ctx, ws, p + ".out_proj", E, E, out, x3));
  return "";
}

std::string sa_load_weights(spe_ctx* ctx, WeightSource& ws);   // sa_model.inl
std::string sa_alloc_workspace(spe_ctx* ctx);
void sa_release(spe_ctx* ctx);

std::string load_weights_impl(spe_ctx* ctx, WeightSource& ws) {
  const spe_config& c = ctx->cfg;
  if (c.backbone == 2) return sa_load_weights(ctx, ws);
  const int E = 256, FF = c.dim_feedforward, Q = c.num_queries, T = ctx->tokens;
  const std::string b = "backbone.0.body";
  // ---- stem + layers
  TRY_S(load_conv_bn(ctx, ws, b + ".conv1", b + ".bn1", 64, 3, 7, &ctx->stem, 192));
  {
    // same filter laid out for the windowed stem GEMM: K = (filter row r, tap s in 0..7, channel c in 0..Cp-1)
    const HostTensor* w = ws.get(b + ".conv1.weight", {64, 3, 7, 7});
    if (!w) return ws.missing;
    const int Cp = 16 / static_cast<int>(dtype_size(ctx->dt)), Kw = 7 * 8 * Cp;
    std::vector<float> w2(static_cast<size_t>(64) * Kw, 0.f);
    for (int o = 0; o < 64; ++o)
      for (int ch = 0; ch < 3; ++ch)
        for (int r = 0; r < 7; ++r)
          for (int t = 0; t < 7; ++t)
            w2[static_cast<size_t>(o) * Kw + (r * 8 + t) * Cp + ch] = w->data[((o * 3 + ch) * 7 + r) * 7 + t];
    TRY_S(upload_gemm_w(ctx, w2, 64, Kw, &ctx->stem2, false, false));
    ctx->stem2.scale = ctx->stem.scale;
    ctx->stem2.bias = ctx->stem.bias;
  }
  ctx->blocks.clear();
  int inplanes = 64;
  const int nblk[3] = {3, 4, 6};
  const int planes_[3] = {64, 128, 256};
  for (int li = 0; li < 3; ++li)
    for (int bi = 0; bi < nblk[li]; ++bi) {
      Bottleneck bk;
      bk.inplanes = inplanes;
      bk.planes = planes_[li];
      bk.stride = (bi == 0 && li > 0) ? 2 : 1;
      bk.has_down = (bi == 0);
      const std::string p = b + ".layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      TRY_S(load_conv_bn(ctx, ws, p + ".conv1", p + ".bn1", bk.planes, inplanes, 1, &bk.c1));
      TRY_S(load_conv_bn(ctx, ws, p + ".conv2", p + ".bn2", bk.planes, bk.planes, 3, &bk.c2));
      TRY_S(load_conv_bn(ctx, ws, p + ".conv3", p + ".bn3", bk.planes * 4, bk.planes, 1, &bk.c3));
      if (bk.has_down) {
        TRY_S(load_conv_bn(ctx, ws, p + ".downsample.0", p + ".downsample.1", bk.planes * 4, inplanes, 1, &bk.down));
        if (ctx->fuse_down) TRY_S(load_fused_down(ctx, ws, p, bk.planes, inplanes, &bk.c3d));
      }
      inplanes = bk.planes * 4;
      ctx->blocks.push_back(bk);
    }
  // ---- neck
  if (c.backbone == 0) {
    TRY_S(load_conv_bn(ctx, ws, "backbone.0.s8_latern", "", 256, 512, 1, &ctx->s8_lat));
    TRY_S(load_conv_bn(ctx, ws, "backbone.0.s16_latern", "", 256, 1024, 3, &ctx->s16_lat));
    {
      // both lateral convolutions are bias-free in the reference; a zero bias gives spe_calibrate a place for its correction
      const std::vector<float> zeros(256, 0.f);
      TRY_S(upload_f32(ctx, zeros.data(), 256, &ctx->s8_lat.bias));
      TRY_S(upload_f32(ctx, zeros.data(), 256, &ctx->s16_lat.bias));
    }
    if (ctx->fold_neck) {
      const HostTensor* w16 = ws.get("backbone.0.s16_latern.weight", {256, 1024, 3, 3});
      if (!w16) return ws.missing;
      std::vector<float> wlr(static_cast<size_t>(9) * 256 * 1024);           // row = tap * 256 + o, column = input channel
      for (int o = 0; o < 256; ++o)
        for (int ci = 0; ci < 1024; ++ci)
          for (int t = 0; t < 9; ++t)
            wlr[(static_cast<size_t>(t) * 256 + o) * 1024 + ci] = w16->data[(static_cast<size_t>(o) * 1024 + ci) * 9 + t];
      TRY_S(upload_gemm_w(ctx, wlr, 9 * 256, 1024, &ctx->s16_lr));
      const std::vector<float> z9(9 * 256, 0.f);
      TRY_S(upload_f32(ctx, z9.data(), 9 * 256, &ctx->s16_lr.bias));
    }
    TRY_S(load_conv_bn(ctx, ws, "backbone.0.output_conv", "", 512, 512, 3, &ctx->out_conv));
    TRY_S(load_vec(ctx, ws, "backbone.0.output_conv.bias", 512, &ctx->out_conv.bias));
  }
  TRY_S(load_conv_bn(ctx, ws, "input_proj", "", E, ctx->featC, 1, &ctx->input_proj));
  TRY_S(load_vec(ctx, ws, "input_proj.bias", E, &ctx->input_proj.bias));
  if (c.backbone == 0 && ctx->fold_neck) {
    const HostTensor* woc = ws.get("backbone.0.output_conv.weight", {512, 512, 3, 3});
    const HostTensor* boc = ws.get("backbone.0.output_conv.bias", {512});
    const HostTensor* wip = ws.get("input_proj.weight", {E, 512, 1, 1});
    const HostTensor* bip = ws.get("input_proj.bias", {E});
    if (!woc || !boc || !wip || !bip) return ws.missing;
    const int Kc = 9 * 512;
    const std::vector<float> oc = repack_conv(*woc);                 // [512][(r*3+s)*512 + c]
    std::vector<float> ocT(static_cast<size_t>(Kc) * 512);           // [Kc][512]: addend_kernel computes X . W^T
    for (int m = 0; m < 512; ++m)
      for (int j = 0; j < Kc; ++j) ocT[static_cast<size_t>(j) * 512 + m] = oc[static_cast<size_t>(m) * Kc + j];
    float *wip_dev = nullptr, *prod_dev = nullptr;
    SPE_CUDA_TRY(cudaMalloc(&wip_dev, sizeof(float) * E * 512));
    SPE_CUDA_TRY(cudaMalloc(&prod_dev, sizeof(float) * E * Kc));
    cudaMemcpy(wip_dev, wip->data, sizeof(float) * E * 512, cudaMemcpyHostToDevice);
    const std::vector<float> zeros(Kc, 0.f);
    std::string e2 = make_addend(ctx, wip_dev, E, 512, ocT.data(), zeros.data(), Kc, prod_dev, Kc, 0);   // fp32 FMA
    std::vector<float> prod(static_cast<size_t>(E) * Kc);
    if (e2.empty()) cudaMemcpy(prod.data(), prod_dev, sizeof(float) * E * Kc, cudaMemcpyDeviceToHost);
    cudaFree(wip_dev);
    cudaFree(prod_dev);
    if (!e2.empty()) return e2;
    TRY_S(upload_gemm_w(ctx, prod, E, Kc, &ctx->out_conv_ip));
    std::vector<float> bias(E);
    for (int o = 0; o < E; ++o) {
      double acc = bip->data[o];
      for (int m = 0; m < 512; ++m) acc += static_cast<double>(wip->data[static_cast<size_t>(o) * 512 + m]) * boc->data[m];
      bias[o] = static_cast<float>(acc);
    }
    TRY_S(upload_f32(ctx, bias.data(), E, &ctx->out_conv_ip.bias));
  }

  // ---- positional embedding and query embedding on the device (fp32) for the addends
  std::vector<float> pos = make_pos(ctx->featH, ctx->featH, E);
  {
    // --position_embedding learned ('v3'): PositionEmbeddingLearned (RV/models/position_encoding.py:55-81) -- the two
    // nn.Embedding(50, 128) tables travel with the checkpoint; pos[y, x] = cat(col_embed[x], row_embed[y]), again a
    // constant of the model, so it folds into the same addends as the sine embedding
    auto ir = ws.t.find("backbone.1.row_embed.weight"), ic = ws.t.find("backbone.1.col_embed.weight");
    if (ir != ws.t.end() && ic != ws.t.end()) {
      const HostTensor* row = ws.get("backbone.1.row_embed.weight", {50, E / 2});
      const HostTensor* col = ws.get("backbone.1.col_embed.weight", {50, E / 2});
      if (!row || !col) return ws.missing;
      const int H = ctx->featH;
      if (H > 50) return "learned position embedding: feature map larger than the 50-entry tables";
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < H; ++x) {
          float* p = pos.data() + (static_cast<size_t>(y) * H + x) * E;
          memcpy(p, col->data + static_cast<size_t>(x) * (E / 2), sizeof(float) * (E / 2));
          memcpy(p + E / 2, row->data + static_cast<size_t>(y) * (E / 2), sizeof(float) * (E / 2));
        }
    }
  }
  float *pos_dev = nullptr, *qe_dev = nullptr;
  SPE_CUDA_TRY(cudaMalloc(&pos_dev, pos.size() * sizeof(float)));
  cudaMemcpy(pos_dev, pos.data(), pos.size() * sizeof(float), cudaMemcpyHostToDevice);
  const HostTensor* qe = ws.get("query_embed.weight", {Q, E});
  if (!qe) { cudaFree(pos_dev); return ws.missing; }
  SPE_CUDA_TRY(cudaMalloc(&qe_dev, static_cast<size_t>(Q) * E * sizeof(float)));
  cudaMemcpy(qe_dev, qe->data, static_cast<size_t>(Q) * E * sizeof(float), cudaMemcpyHostToDevice);

  std::string s;
  auto body = [&]() -> std::string {
    // ---- encoder
    ctx->enc.assign(c.enc_layers, EncLayer{});
    for (int i = 0; i < c.enc_layers; ++i) {
      EncLayer& L = ctx->enc[i];
      const std::string p = "transformer.encoder.layers." + std::to_string(i);
      TRY_S(load_mha_self(ctx, ws, p + ".self_attn", pos_dev, T, &L.qkv, &L.out, &L.addend));
      TRY_S(load_linear(ctx, ws, p + ".linear1", FF, E, &L.ff1));
      TRY_S(load_linear(ctx, ws, p + ".linear2", E, FF, &L.ff2));
      if (ctx->dt == kBF16 && ctx->mixed_ffn && ffn_fused_supported(kTF32, E, FF)) {
        const HostTensor* w1 = ws.get(p + ".linear1.weight", {FF, E});
        const HostTensor* w2 = ws.get(p + ".linear2.weight", {E, FF});
        if (!w1 || !w2) return ws.missing;
        TRY_S(upload_tf32_copy(ctx, w1->data, FF, E, &L.ff1_32));
        TRY_S(upload_tf32_copy(ctx, w2->data, E, FF, &L.ff2_32));
        TRY_S(load_vec(ctx, ws, p + ".linear1.bias", FF, &L.ff1_32.bias));
        TRY_S(load_vec(ctx, ws, p + ".linear2.bias", E, &L.ff2_32.bias));
      }
      TRY_S(load_vec(ctx, ws, p + ".norm1.weight", E, &L.n1g));
      TRY_S(load_vec(ctx, ws, p + ".norm1.bias", E, &L.n1b));
      TRY_S(load_vec(ctx, ws, p + ".norm2.weight", E, &L.n2g));
      TRY_S(load_vec(ctx, ws, p + ".norm2.bias", E, &L.n2b));
    }
    // ---- decoder
    ctx->dec.assign(c.dec_layers, DecLayer{});
    const int LD = c.dec_layers;
    std::vector<float> kv_w(static_cast<size_t>(LD) * 2 * E * E);
    TRY_S(dmalloc(ctx, &ctx->ca_kv_addend, static_cast<long long>(T) * LD * 2 * E));
    for (int i = 0; i < LD; ++i) {
      DecLayer& L = ctx->dec[i];
      const std::string p = "transformer.decoder.layers." + std::to_string(i);
      // the decoder and the heads run as 3xTF32: their rounding error dominates the keypoint error budget (0.5 px at
      // crop sides up to 1748 px), while they are < 6 % of the FLOPs
      TRY_S(load_mha_self(ctx, ws, p + ".self_attn", qe_dev, Q, &L.sa_qkv, &L.sa_out, &L.sa_addend, ctx->dec_x3));
      const HostTensor* w = ws.get(p + ".multihead_attn.in_proj_weight", {3 * E, E});
      const HostTensor* bb = ws.get(p + ".multihead_attn.in_proj_bias", {3 * E});
      if (!w || !bb) return ws.missing;
      // query projection: (tgt + query_pos) Wq^T + bq
      TRY_S(upload_gemm_w(ctx, std::vector<float>(w->data, w->data + E * E), E, E, &L.ca_q, ctx->dec_x3));
      TRY_S(dmalloc(ctx, &L.ca_q_addend, static_cast<long long>(Q) * E));
      TRY_S(make_addend(ctx, qe_dev, Q, E, w->data, bb->data, E, L.ca_q_addend, E, 0));
      L.ca_q.addend = L.ca_q_addend; L.ca_q.addend_rows = Q; L.ca_q.addend_ld = E;
      // key/value projections of every layer share the encoder memory: stack them into one GEMM
      memcpy(kv_w.data() + static_cast<size_t>(i) * 2 * E * E, w->data + E * E, sizeof(float) * 2 * E * E);
      TRY_S(make_addend(ctx, pos_dev, T, E, w->data + E * E, bb->data + E, E, ctx->ca_kv_addend, LD * 2 * E,
                        i * 2 * E));
      TRY_S(make_addend(ctx, nullptr, T, E, w->data + 2 * E * E, bb->data + 2 * E, E, ctx->ca_kv_addend,
                        LD * 2 * E, i * 2 * E + E));
      TRY_S(load_linear(ctx, ws, p + ".multihead_attn.out_proj", E, E, &L.ca_out, ctx->dec_x3));
      TRY_S(load_linear(ctx, ws, p + ".linear1", FF, E, &L.ff1, ctx->dec_x3 && ctx->dec_ffn_x3));
      TRY_S(load_linear(ctx, ws, p + ".linear2", E, FF, &L.ff2, ctx->dec_x3 && ctx->dec_ffn_x3));
      TRY_S(load_vec(ctx, ws, p + ".norm1.weight", E, &L.n1g));
      TRY_S(load_vec(ctx, ws, p + ".norm1.bias", E, &L.n1b));
      TRY_S(load_vec(ctx, ws, p + ".norm2.weight", E, &L.n2g));
      TRY_S(load_vec(ctx, ws, p + ".norm2.bias", E, &L.n2b));
      TRY_S(load_vec(ctx, ws, p + ".norm3.weight", E, &L.n3g));
      TRY_S(load_vec(ctx, ws, p + ".norm3.bias", E, &L.n3b));
    }
    TRY_S(upload_gemm_w(ctx, kv_w, LD * 2 * E, E, &ctx->ca_kv_all));
    ctx->ca_kv_all.addend = ctx->ca_kv_addend; ctx->ca_kv_all.addend_rows = T; ctx->ca_kv_all.addend_ld = LD * 2 * E;
    ctx->kv_split3 = false;
    if (ctx->dt == kTF32 && ctx->kv_x3 != 0) {
      // 3xTF32 with the operand split hoisted out of the GEMM: the last encoder LayerNorm emits [x_hi | x_lo | x_hi]
      // and the weights are stored as [W_hi | W_hi | W_lo], so a plain K = 3E GEMM yields the compensated product
      std::vector<float> w3(static_cast<size_t>(LD) * 2 * E * 3 * E);
      for (int n = 0; n < LD * 2 * E; ++n)
        for (int k = 0; k < E; ++k) {
          const float w = kv_w[static_cast<size_t>(n) * E + k];
          const float hi = host_rna_tf32(w), lo = host_rna_tf32(w - hi);
          float* row = w3.data() + static_cast<size_t>(n) * 3 * E;
          row[k] = hi; row[E + k] = hi; row[2 * E + k] = lo;
        }
      TRY_S(upload_gemm_w(ctx, w3, LD * 2 * E, 3 * E, &ctx->ca_kv_x3, false, false));   // already error-compensated
      ctx->kv_split3 = true;
    }
    TRY_S(load_vec(ctx, ws, "transformer.decoder.norm.weight", E, &ctx->dn_g));
    TRY_S(load_vec(ctx, ws, "transformer.decoder.norm.bias", E, &ctx->dn_b));
    // ---- heads
    const HostTensor* cw = ws.get("cls_embed.weight", {12, E});
    if (!cw) return ws.missing;
    TRY_S(upload_f32(ctx, cw->data, 12 * E, &ctx->cls_w));
    TRY_S(load_vec(ctx, ws, "cls_embed.bias", 12, &ctx->cls_b));
    TRY_S(load_linear(ctx, ws, "point_embed.layers.0", E, E, &ctx->pt0, ctx->dec_x3));
    TRY_S(load_linear(ctx, ws, "point_embed.layers.1", E, E, &ctx->pt1, ctx->dec_x3));
    const HostTensor* p2 = ws.get("point_embed.layers.2.weight", {2, E});
    if (!p2) return ws.missing;
    TRY_S(upload_f32(ctx, p2->data, 2 * E, &ctx->pt2_w));
    TRY_S(load_vec(ctx, ws, "point_embed.layers.2.bias", 2, &ctx->pt2_b));
    if (c.has_sigma) {
      TRY_S(load_linear(ctx, ws, "sigma_embed.layers.0", E, E, &ctx->sg0, ctx->dec_x3));
      TRY_S(load_linear(ctx, ws, "sigma_embed.layers.1", E, E, &ctx->sg1, ctx->dec_x3));
      const HostTensor* s2 = ws.get("sigma_embed.layers.2.weight", {1, E});
      if (!s2) return ws.missing;
      TRY_S(upload_f32(ctx, s2->data, E, &ctx->sg2_w));
      TRY_S(load_vec(ctx, ws, "sigma_embed.layers.2.bias", 1, &ctx->sg2_b));
    }
    return "";
  };
  s = body();
  cudaFree(pos_dev);
  cudaFree(qe_dev);
  return s;
}

// ------------------------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------------------------
std::string alloc_workspace(spe_ctx* ctx) {
  const spe_config& c = ctx->cfg;
  const long long B = c.max_batch;
  const long long es = static_cast<long long>(dtype_size(ctx->dt));
  const long long R = c.input_size;
  const long long h2 = R / 2, h4 = R / 4, h8 = R / 8, h16 = R / 16;
  const long long T = ctx->tokens, Q = c.num_queries, LD = c.dec_layers, FF = c.dim_feedforward;
  auto AB = [&](void** p, long long bytes) {
    std::string e = dmalloc_bytes(ctx, p, bytes);
    if (e.empty()) ctx->ws_fields.push_back({p, bytes});
    return e;
  };
  auto A = [&](void** p, long long elems) { return AB(p, elems * es); };
  if (c.backbone == 2) {
    TRY_S(sa_alloc_workspace(ctx));
  } else {
  TRY_S(A(&ctx->S0, B * h2 * h2 * 192));
  TRY_S(AB(&ctx->SP, B * (R + 6) * (R + 6) * 16));
  TRY_S(A(&ctx->S1, B * h2 * h2 * 64));
  TRY_S(A(&ctx->P0, B * h4 * h4 * 256));
  TRY_S(A(&ctx->P1, B * h4 * h4 * 256));
  TRY_S(A(&ctx->L1OUT, ctx->head_chunk > 0 ? B * h4 * h4 * 256 : 4));
  TRY_S(A(&ctx->T1, B * h4 * h4 * 128));
  TRY_S(A(&ctx->T2, B * h4 * h4 * 128));
  TRY_S(A(&ctx->DS, B * h4 * h4 * 256));
  TRY_S(A(&ctx->COL, B * h8 * h8 * 1152));
  TRY_S(A(&ctx->L2OUT, B * h8 * h8 * 512));
  TRY_S(A(&ctx->L3OUT, B * h16 * h16 * 1024));
  if (c.backbone == 0) {
    TRY_S(AB(reinterpret_cast<void**>(&ctx->YLR), B * h16 * h16 * 9 * 256 * 4));
    TRY_S(A(&ctx->UP, B * h8 * h8 * 1024));
    TRY_S(A(&ctx->CAT, B * h8 * h8 * 512));
    TRY_S(A(&ctx->FEAT, B * h8 * h8 * 512));
  }
  TRY_S(A(&ctx->X, B * T * 256));
  TRY_S(A(&ctx->X2, B * T * 256));
  if (ctx->dt == kTF32) TRY_S(A(&ctx->XS, B * T * 768));
  if (ctx->dt == kBF16) TRY_S(AB(reinterpret_cast<void**>(&ctx->XF), B * T * 256 * 4));
  // bf16 storage: the encoder's Q|K|V are written as fp32 (TF32 values) for the tcgen05 attention kernel -> 4 B/element
  TRY_S(A(&ctx->QKV, B * T * 768 * (ctx->dt == kTF32 ? 1 : 2)));
  TRY_S(A(&ctx->ATT, B * T * 256));
  TRY_S(A(&ctx->HID, B * T * FF));
  TRY_S(A(&ctx->KV, B * T * LD * 512));
  TRY_S(A(&ctx->TGT, B * Q * 256));
  TRY_S(A(&ctx->TGT2, B * Q * 256));
  TRY_S(A(&ctx->DQKV, B * Q * 768));
  TRY_S(A(&ctx->DQ, B * Q * 256));
  TRY_S(A(&ctx->DATT, B * Q * 256));
  TRY_S(A(&ctx->DHID, B * Q * FF));
  TRY_S(A(&ctx->HS, LD * B * Q * 256));
  TRY_S(A(&ctx->H1, LD * B * Q * 256));
  TRY_S(A(&ctx->H2, LD * B * Q * 256));
  TRY_S(A(&ctx->G1, B * Q * 256));
  TRY_S(A(&ctx->G2, B * Q * 256));
  }
  if (ctx->SP != nullptr) SPE_CUDA_TRY(cudaMemset(ctx->SP, 0, static_cast<size_t>(B * (R + 6) * (R + 6) * 16)));   // zero border
  {
    std::vector<void*> set0;
    for (const auto& f : ctx->ws_fields) set0.push_back(*f.field);
    ctx->ws_sets.assign(1, set0);
    ctx->ws_current = 0;
  }
  TRY_S(dmalloc(ctx, &ctx->colsum, 8192 + 2 * 256 * 4096));   // sums + per-block partials
  TRY_S(dmalloc_bytes(ctx, &ctx->dec0_tgt, Q * 256 * es));
  TRY_S(dmalloc_bytes(ctx, &ctx->dec0_q, Q * 256 * es));
  // pipeline buffers
  TRY_S(dmalloc(ctx, &ctx->boxes_dev, B * 4));
  TRY_S(dmalloc(ctx, &ctx->images_dev, B * 3 * R * R));
  TRY_S(dmalloc(ctx, &ctx->p_logits, B * Q * 12));
  TRY_S(dmalloc(ctx, &ctx->p_points, B * Q * 2));
  TRY_S(dmalloc(ctx, &ctx->p_logsig, B * Q * 2));
  TRY_S(dmalloc(ctx, &ctx->p_quat, B * 4));
  TRY_S(dmalloc(ctx, &ctx->p_tvec, B * 3));
  TRY_S(dmalloc(ctx, &ctx->p_assign, B * 11));
  TRY_S(dmalloc(ctx, &ctx->p_status, B));
  return "";
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
namespace {

struct Fwd {
  spe_ctx* ctx;
  cudaStream_t st;
  int B;
  Dtype dt;
  long long es;

  void* col(void* base, long long elem_off) const { return static_cast<uint8_t*>(base) + elem_off * es; }

  std::string tap(const char* name, const void* buf, long long elems) {
    if (!ctx->taps_enabled) return "";
    const long long bytes = elems * es;
    auto it = ctx->taps.find(name);
    if (it == ctx->taps.end() || it->second.bytes < bytes) {
      void* p = nullptr;
      SPE_CUDA_TRY(cudaMalloc(&p, static_cast<size_t>(bytes)));
      ctx->allocs.push_back(p);
      ctx->taps[name] = spe_ctx::Tap{p, bytes};
      it = ctx->taps.find(name);
    }
    it->second.bytes = bytes;
    SPE_CUDA_TRY(cudaMemcpyAsync(it->second.buf, buf, static_cast<size_t>(bytes), cudaMemcpyDeviceToDevice, st));
    return "";
  }

  // spe_calibrate: measure the column means of this layer's input ([rows, C], row stride ld) and fold what the rounded
  // weights lose on them into the layer's bias (see bias_correction_kernel)
  std::string calibrate_layer(const void* A, long long rows, int C, long long ld, const GemmW& w) {
    if (!ctx->calibrating || w.w32 == nullptr || w.x3 || rows <= 0) return "";
    if (C > 4096) return "calibration: more than 4096 input channels";
    if (w.bias == nullptr && w.addend == nullptr) return "";
    float* cs = ctx->colsum;
    float* cs_mma = ctx->colsum + 4096;
    constexpr int kColsumBlocks = 256;
    const unsigned blocks = static_cast<unsigned>(rows < kColsumBlocks ? rows : kColsumBlocks);
    float* part = ctx->colsum + 8192;
    float* part_mma = part + static_cast<long long>(kColsumBlocks) * 4096;
    const unsigned cgrid = static_cast<unsigned>((w.N + 7) / 8);
    const float inv = 1.0f / static_cast<float>(rows);
    if (dt == kTF32) {
      colsum_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(A), rows, C, ld, part, part_mma);
      colsum_reduce_kernel<<<(C + 255) / 256, 256, 0, st>>>(part, part_mma, static_cast<int>(blocks), C, cs, cs_mma);
      bias_correction_kernel<float><<<cgrid, 256, 0, st>>>(w.w32, static_cast<const float*>(w.w), cs, cs_mma, inv, w.N, w.K, C,
                                                           w.scale, w.bias, w.applied, w.addend, w.addend_rows, w.addend_ld);
    } else {
      colsum_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(A), rows, C, ld, part, part_mma);
      colsum_reduce_kernel<<<(C + 255) / 256, 256, 0, st>>>(part, part_mma, static_cast<int>(blocks), C, cs, cs_mma);
      bias_correction_kernel<__nv_bfloat16><<<cgrid, 256, 0, st>>>(w.w32, static_cast<const __nv_bfloat16*>(w.w), cs, cs_mma, inv,
                                                                   w.N, w.K, C, w.scale, w.bias, w.applied, w.addend,
                                                                   w.addend_rows, w.addend_ld);
    }
    SPE_CUDA_TRY(cudaGetLastError());
    return "";
  }

  // calibrate_layer for a GEMM over two concatenated operand sources (GemmDesc::A2): columns [0, C1) are measured on
  // A1 [rows1, C1], columns [C1, C1 + C2) on A2 [rows2, C2] (for a stride-2 source the mean is taken over every input
  // pixel, not only the sampled ones: second order, like the zero padding of the 3x3 convolutions)
  std::string calibrate_two(const void* A1, long long rows1, int C1, const void* A2, long long rows2, int C2, const GemmW& w) {
    if (!ctx->calibrating || w.w32 == nullptr || rows1 <= 0 || rows2 <= 0) return "";
    if (C1 + C2 > 4096) return "calibration: more than 4096 input channels";
    float* cs = ctx->colsum;
    float* cs_mma = ctx->colsum + 4096;
    constexpr int kColsumBlocks = 256;
    float* part = ctx->colsum + 8192;
    float* part_mma = part + static_cast<long long>(kColsumBlocks) * 4096;
    const void* src[2] = {A1, A2};
    const long long rows[2] = {rows1, rows2};
    const int C[2] = {C1, C2};
    for (int i = 0; i < 2; ++i) {
      const unsigned blocks = static_cast<unsigned>(rows[i] < kColsumBlocks ? rows[i] : kColsumBlocks);
      float* o = cs + (i ? C1 : 0);
      float* om = cs_mma + (i ? C1 : 0);
      if (dt == kTF32) colsum_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(src[i]), rows[i], C[i], C[i], part, part_mma);
      else colsum_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src[i]), rows[i], C[i], C[i], part, part_mma);
      colsum_reduce_kernel<<<(C[i] + 255) / 256, 256, 0, st>>>(part, part_mma, static_cast<int>(blocks), C[i], o, om);
      if (i == 1 && rows2 != rows1)
        scale_pair_kernel<<<(C2 + 255) / 256, 256, 0, st>>>(o, om, C2, static_cast<float>(rows1) / static_cast<float>(rows2));
    }
    const unsigned cgrid = static_cast<unsigned>((w.N + 7) / 8);
    const float inv = 1.0f / static_cast<float>(rows1);
    if (dt == kTF32)
      bias_correction_kernel<float><<<cgrid, 256, 0, st>>>(w.w32, static_cast<const float*>(w.w), cs, cs_mma, inv, w.N, w.K, C1 + C2,
                                                           w.scale, w.bias, w.applied, w.addend, w.addend_rows, w.addend_ld);
    else
      bias_correction_kernel<__nv_bfloat16><<<cgrid, 256, 0, st>>>(w.w32, static_cast<const __nv_bfloat16*>(w.w), cs, cs_mma, inv,
                                                                   w.N, w.K, C1 + C2, w.scale, w.bias, w.applied, w.addend,
                                                                   w.addend_rows, w.addend_ld);
    SPE_CUDA_TRY(cudaGetLastError());
    return "";
  }

  // out[M, N] = act(scale * A W^T + bias (+ residual))
  std::string gemm(const void* A, long long M, const GemmW& w, void* out, int out_ld, bool relu,
                   const void* residual = nullptr, int res_ld = 0, int res_mod = 0, int res_f32 = 0,
                   bool use_scale_bias = true, int out_f32 = 0, bool exact_out = false) {
    TRY_S(calibrate_layer(A, M, w.K, w.K, w));
    GemmDesc d;
    d.mode = 0;
    d.A = A; d.M = M; d.K = w.K; d.lda = w.K;
    d.Wt = w.w; d.N = w.N;
    d.scale = use_scale_bias ? w.scale : nullptr;
    d.bias = use_scale_bias ? w.bias : nullptr;
    d.residual = residual; d.res_ld = res_ld; d.res_mod = res_mod; d.res_f32 = res_f32;
    d.relu = relu ? 1 : 0;
    d.out = out; d.out_ld = out_ld;
    d.x3 = w.x3;
    // 3xTF32 chains keep full fp32 activations; so do outputs that no tensor-core GEMM reads (`exact_out`: values that
    // only feed a residual add or a LayerNorm -- rounding them would add noise to the residual stream for nothing)
    d.round_out = (w.x3 || exact_out) ? 0 : 1;
    d.out_f32 = out_f32;
    return launch_gemm(dt, d, ctx->num_sms, st);
  }
  // R x R convolution (pad = R/2, stride 1 or 2) as implicit GEMM; H = input extent
  std::string conv(const void* x, int H, int C, int R, int stride, const GemmW& w, void* out, int out_ld,
                   bool relu, bool exact_out = false, int c_ld = 0, int act = 0) {
    // the mean is taken over every input position; the zero padding at the border is ignored (second-order)
    TRY_S(calibrate_layer(x, static_cast<long long>(B) * H * H, C, c_ld > 0 ? c_ld : C, w));
    GemmDesc d;
    d.mode = 1;
    d.A = x; d.NB = B; d.H = H; d.W = H; d.C = C; d.R = R; d.S = R; d.pad = R / 2; d.conv_stride = stride;
    d.c_ld = c_ld;
    d.Wt = w.w; d.N = w.N;
    d.scale = w.scale; d.bias = w.bias;
    d.relu = act > 1 ? act : (relu ? 1 : 0);      // act 2 / 3: SiLU / GELU (3xTF32 kernels only)
    d.out = out; d.out_ld = out_ld;
    d.x3 = w.x3;
    d.round_out = (exact_out || w.x3) ? 0 : 1;
    return launch_gemm(dt, d, ctx->num_sms, st);
  }
  std::string conv3x3(const void* x, int H, int C, const GemmW& w, void* out, int out_ld, bool relu) {
    return conv(x, H, C, 3, 1, w, out, out_ld, relu);
  }
  std::string attn(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int Lq,
                   int Lk, int exact_out = 0, int mixed = 0, int x3 = 0, bool q_broadcast = false) {
    AttnDesc a;
    a.exact_out = exact_out;
    a.x3 = x3;
    a.mixed = mixed;
    a.q = q; a.k = k; a.v = v; a.out = out;
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = 256;
    a.bsq = q_broadcast ? 0 : static_cast<long long>(Lq) * ldq;   // 0: one [Lq, ldq] query block shared by every image
    a.bsk = static_cast<long long>(Lk) * ldk;
    a.bsv = static_cast<long long>(Lk) * ldv; a.bso = static_cast<long long>(Lq) * 256;
    a.B = B; a.heads = 8; a.Lq = Lq; a.Lk = Lk;
    a.scale = 1.0f / sqrtf(32.0f);
    return launch_attention(dt, a, st);
  }
  std::string ln(const void* in, const float* g, const float* b, long long rows, void* out, int exact = 0) {
    return launch_layernorm(dt, in, g, b, rows, 256, out, st, exact);
  }
};

}  // namespace

// Unrounded residual stream (see spe_ctx::exact_stream): only together with the calibration that removes the bias of the
// tensor core's operand truncation -- an uncalibrated context keeps rounding every store.
// the K/V projection runs in its hoisted 3xTF32 form (see spe_ctx::kv_x3)
static bool kv_split_on(const spe_ctx* ctx) {
  if (!ctx->kv_split3) return false;
  return ctx->kv_x3 > 0 || !(ctx->calibrated || ctx->calibrating);
}
// profiling only (results are garbage): SPE_DBG_SKIP bit mask leaves stages out of the schedule so that their cost
// under batches-in-flight can be read off the step time: 1 decoder + heads, 2 encoder attention, 4 encoder FFN,
// 8 decoder K/V projection, 16 layer1, 32 layer2, 64 layer3, 128 neck 3x3 convolutions, 256 encoder Q|K|V + out-proj
static int dbg_skip() {
  static const int m = getenv("SPE_DBG_SKIP") ? atoi(getenv("SPE_DBG_SKIP")) : 0;
  return m;
}
static bool exact_stream_on(const spe_ctx* ctx) {
  return ctx->exact_stream && ctx->dt == kTF32 && (ctx->calibrated || ctx->calibrating);
}

// One residual block of layer1..layer3 on `B` images: conv1 (1x1) -> conv2 (3x3, stride on it) -> conv3 (1x1) + identity / downsample
static std::string run_bottleneck(spe_ctx* ctx, Fwd& f, const Bottleneck& bk, const void* cur, int H, void* nxt) {
  const long long Bl = f.B;
  const int Ho = H / bk.stride;
  const long long Min = Bl * H * H, Mout = Bl * Ho * Ho;
  TRY_S(f.gemm(cur, Min, bk.c1, ctx->T1, bk.planes, true));
  if (bk.stride == 1) {
    TRY_S(f.conv3x3(ctx->T1, H, bk.planes, bk.c2, ctx->T2, bk.planes, true));
  } else {
    TRY_S(f.conv(ctx->T1, H, bk.planes, 3, 2, bk.c2, ctx->T2, bk.planes, true));   // 3x3 / stride 2
  }
  if (bk.has_down && bk.c3d.w != nullptr && !ctx->taps_enabled) {
    // conv3 and the downsample branch as one GEMM over [T2 | cur sampled at the stride] (see Bottleneck::c3d)
    TRY_S(f.calibrate_two(ctx->T2, Mout, bk.planes, cur, Min, bk.inplanes, bk.c3d));
    GemmDesc d;
    d.mode = 0;
    d.A = ctx->T2; d.M = Mout; d.K = bk.planes; d.lda = bk.planes;
    d.A2 = cur; d.K2 = bk.inplanes; d.lda2 = bk.inplanes;
    d.a2_stride = bk.stride; d.a2_NB = f.B; d.a2_H = H; d.a2_W = H;
    d.Wt = bk.c3d.w; d.N = bk.c3d.N;
    d.bias = bk.c3d.bias;
    d.relu = 1;
    d.out = nxt; d.out_ld = bk.planes * 4;
    d.round_out = exact_stream_on(ctx) ? 0 : 1;
    return launch_gemm(f.dt, d, ctx->num_sms, f.st);
  }
  const void* identity = cur;
  if (bk.has_down) {
    if (bk.stride == 1) {
      TRY_S(f.gemm(cur, Min, bk.down, ctx->DS, bk.planes * 4, false, nullptr, 0, 0, 0, true, 0, true));
    } else {
      TRY_S(f.conv(cur, H, bk.inplanes, 1, 2, bk.down, ctx->DS, bk.planes * 4, false, true));  // 1x1 / stride 2
    }
    identity = ctx->DS;
  }
  return f.gemm(ctx->T2, Mout, bk.c3, nxt, bk.planes * 4, true, identity, bk.planes * 4, 0, 0, true, 0, exact_stream_on(ctx));
}

// stem + max-pool + layer1 for `B` images starting at `images`; the layer1 output goes to `l1dst` when given, else to
// the ping-pong buffer the block sequence ends on.  *out = where it is.
static std::string trunk_head(spe_ctx* ctx, const float* images, int B, void* l1dst, const void** out, cudaStream_t st) {
  const spe_config& c = ctx->cfg;
  Fwd f{ctx, st, B, ctx->dt, static_cast<long long>(dtype_size(ctx->dt))};
  const int R = c.input_size;
  const int h2 = R / 2, h4 = R / 4;
  const long long Bl = B;

  // ---- stem
  if (ctx->calibrating && ctx->stem.w32 != nullptr) {
    // channel means of the normalised image; the filter in im2col layout [64, 192] has k % 3 = channel
    nchw_chansum_kernel<<<3, 256, 0, st>>>(images, B, 3, static_cast<long long>(R) * R, ctx->colsum);
    const GemmW& w = ctx->stem;
    const float inv = 1.0f / (static_cast<float>(B) * R * R);
    if (f.dt == kTF32)
      bias_correction_kernel<float><<<(w.N + 7) / 8, 256, 0, st>>>(w.w32, static_cast<const float*>(w.w), ctx->colsum, ctx->colsum,
                                                                  inv, w.N, w.K, 3, w.scale, w.bias, w.applied, nullptr, 0, 0);
    else
      bias_correction_kernel<__nv_bfloat16><<<(w.N + 7) / 8, 256, 0, st>>>(w.w32, static_cast<const __nv_bfloat16*>(w.w),
                                                                          ctx->colsum, ctx->colsum, inv, w.N, w.K, 3, w.scale,
                                                                          w.bias, w.applied, nullptr, 0, 0);
    SPE_CUDA_TRY(cudaGetLastError());
  }
  bool stem_done = false;
  if (ctx->stem_windowed) {
    if (!(ctx->stem_prefilled && !ctx->calibrating && !ctx->taps_enabled)) TRY_S(launch_stem_pad(f.dt, images, B, R, R, ctx->SP, st));
    GemmDesc d;
    d.mode = 2;
    d.A = ctx->SP; d.NB = B; d.H = R; d.W = R; d.C = 16 / static_cast<int>(f.es);
    d.Wt = ctx->stem2.w; d.N = 64; d.scale = ctx->stem2.scale; d.bias = ctx->stem2.bias; d.relu = 1;
    d.out = ctx->S1; d.out_ld = 64;
    const std::string e = launch_gemm(f.dt, d, ctx->num_sms, st);
    if (e.empty()) stem_done = true;
    else if (e.rfind("stem: cuTensorMapEncodeTiled", 0) == 0) ctx->stem_windowed = false;   // use the im2col form
    else return e;
  }
  if (!stem_done) {
    if (ctx->stem_prefilled)
      return "stem: the windowed tensor map was refused although the crop kernel already wrote the padded layout (SPE_CROP_STEM=0 avoids this)";
    TRY_S(launch_stem_im2col(f.dt, images, B, R, R, ctx->S0, st));
    {
      const bool cal = ctx->calibrating;
      ctx->calibrating = false;                 // the stem was calibrated from the image channel means above
      const std::string e = f.gemm(ctx->S0, Bl * h2 * h2, ctx->stem, ctx->S1, 64, true);
      ctx->calibrating = cal;
      if (!e.empty()) return e;
    }
  }
  TRY_S(f.tap("stem", ctx->S1, Bl * h2 * h2 * 64));
  TRY_S(launch_maxpool3x3s2(f.dt, ctx->S1, B, h2, h2, 64, ctx->P0, st));

  // ---- layer1
  const void* cur = ctx->P0;
  if (!(dbg_skip() & 16)) {
    for (int bi = 0; bi < 3; ++bi) {
      void* nxt = (bi == 2 && l1dst != nullptr) ? l1dst : ((cur == ctx->P0) ? ctx->P1 : ctx->P0);
      TRY_S(run_bottleneck(ctx, f, ctx->blocks[bi], cur, h4, nxt));
      cur = nxt;
    }
  }
  TRY_S(f.tap("layer1", cur, Bl * h4 * h4 * 256));
  *out = cur;
  return "";
}

// backbone + neck + input_proj + encoder for `B` images starting at `images`; the encoder output (memory) of those
// images is left in Xc ([B * tokens, 256]).  Every other buffer is the shared scratch region, so consecutive chunks
// reuse the same cache lines (the whole working set of a chunk is sized to stay L2-resident).
static std::string forward_trunk(spe_ctx* ctx, const float* images, int B, void* Xc, cudaStream_t st) {
  const spe_config& c = ctx->cfg;
  Fwd f{ctx, st, B, ctx->dt, static_cast<long long>(dtype_size(ctx->dt))};
  const int R = c.input_size;
  const int h4 = R / 4;
  const long long Bl = B;

  // ---- stem, max-pool, layer1: whole batch, or chunk by chunk (see spe_ctx::head_chunk)
  const void* cur = nullptr;
  const int hc = ctx->head_chunk;
  if (hc > 0 && hc < B && !ctx->taps_enabled && !ctx->calibrating) {
    const long long img_elems = 3ll * R * R, l1_elems = static_cast<long long>(h4) * h4 * 256;
    for (int c0 = 0; c0 < B; c0 += hc) {
      const int nb = (B - c0) < hc ? (B - c0) : hc;
      const void* dummy = nullptr;
      TRY_S(trunk_head(ctx, images + c0 * img_elems, nb, f.col(ctx->L1OUT, c0 * l1_elems), &dummy, st));
    }
    cur = ctx->L1OUT;
  } else {
    TRY_S(trunk_head(ctx, images, B, nullptr, &cur, st));
  }

  // ---- layer2, layer3
  int H = h4;
  int bidx = 3;
  const int nblk[3] = {3, 4, 6};
  for (int li = 1; li < 3; ++li) {
    for (int bi = 0; bi < nblk[li]; ++bi, ++bidx) {
      const Bottleneck& bk = ctx->blocks[bidx];
      const int Ho = H / bk.stride;
      if (dbg_skip() & (16 << li)) { H = Ho; continue; }
      void* nxt;
      if (bi == nblk[li] - 1 && li == 1) nxt = ctx->L2OUT;
      else if (bi == nblk[li] - 1 && li == 2) nxt = ctx->L3OUT;
      else nxt = (cur == ctx->P0) ? ctx->P1 : ctx->P0;
      TRY_S(run_bottleneck(ctx, f, bk, cur, H, nxt));
      cur = nxt;
      H = Ho;
    }
    const char* names[3] = {"layer1", "layer2", "layer3"};
    TRY_S(f.tap(names[li], cur, Bl * H * H * ctx->blocks[bidx - 1].planes * 4));
  }

  // ---- neck
  const int FH = ctx->featH;
  const long long T = ctx->tokens;
  const void* feat = c.backbone == 0 ? ctx->FEAT : ctx->L3OUT;
  if (c.backbone == 0) {
    const int h16 = R / 16;
    TRY_S(f.gemm(ctx->L2OUT, Bl * T, ctx->s8_lat, ctx->CAT, 512, false));
    if (!(dbg_skip() & 128)) {
    if (ctx->s16_lr.w != nullptr && !ctx->taps_enabled) {
      // the nine taps of s16_latern on the 14 x 14 map, then upsample + shift + add (see spe_ctx::s16_lr)
      TRY_S(f.gemm(ctx->L3OUT, Bl * h16 * h16, ctx->s16_lr, ctx->YLR, 9 * 256, false, nullptr, 0, 0, 0, true, 1, true));
      TRY_S(launch_upsample_tapsum(f.dt, ctx->YLR, B, h16, h16, 256, f.col(ctx->CAT, 256), 512, st));
    } else {
      TRY_S(launch_upsample2x(f.dt, ctx->L3OUT, B, h16, h16, 1024, ctx->UP, st));
      TRY_S(f.conv3x3(ctx->UP, FH, 1024, ctx->s16_lat, f.col(ctx->CAT, 256), 512, false));
    }
    if (ctx->out_conv_ip.w != nullptr && !ctx->taps_enabled) {
      // output_conv and input_proj as one 3x3 convolution (see spe_ctx::out_conv_ip); the 512-channel neck output only
      // exists when the activation taps ask for it
      TRY_S(f.conv(ctx->CAT, FH, 512, 3, 1, ctx->out_conv_ip, Xc, 256, false, exact_stream_on(ctx)));
      feat = nullptr;
    } else {
      TRY_S(f.conv3x3(ctx->CAT, FH, 512, ctx->out_conv, ctx->FEAT, 512, false));
      TRY_S(f.tap("neck", ctx->FEAT, Bl * T * 512));
      feat = ctx->FEAT;
    }
    }
  } else {
    feat = ctx->L3OUT;
  }
  if (feat != nullptr)
    TRY_S(f.gemm(feat, Bl * T, ctx->input_proj, Xc, 256, false, nullptr, 0, 0, 0, true, 0, exact_stream_on(ctx)));
  TRY_S(f.tap("input_proj", Xc, Bl * T * 256));

  // ---- encoder
  const int Ti = static_cast<int>(T);
  for (int i = 0; i < c.enc_layers; ++i) {
    const EncLayer& L = ctx->enc[i];
    if (f.dt == kTF32 || !ctx->mixed_attention) {
      if (!(dbg_skip() & 256)) TRY_S(f.gemm(Xc, Bl * T, L.qkv, ctx->QKV, 768, false, L.addend, 768, Ti, 1));
      if (!(dbg_skip() & 2)) TRY_S(f.attn(ctx->QKV, 768, f.col(ctx->QKV, 256), 768, f.col(ctx->QKV, 512), 768, ctx->ATT, Ti, Ti));
    } else {
      // bf16 storage: Q|K|V leave the GEMM as fp32 (TF32 values) so that the encoder attention can run on the tcgen05 /
      // TMEM kernel (184 us per layer at B = 64) instead of the mma.sync register kernel (445 us); its output is bf16
      float* q32 = static_cast<float*>(ctx->QKV);
      TRY_S(f.gemm(Xc, Bl * T, L.qkv, q32, 768, false, L.addend, 768, Ti, 1, true, 1));
      TRY_S(f.attn(q32, 768, q32 + 256, 768, q32 + 512, 768, ctx->ATT, Ti, Ti, 0, 1));
    }
    if (!(dbg_skip() & 256)) TRY_S(f.gemm(ctx->ATT, Bl * T, L.out, ctx->X2, 256, false, Xc, 256, 0, 0, true, 0, true));   // feeds LayerNorm only
    const bool ffn_mixed = f.dt == kBF16 && ctx->mixed_ffn && L.ff1_32.w != nullptr && !ctx->taps_enabled && !ctx->calibrating &&
                           (Bl * T + 127) / 128 >= ctx->num_sms / 2 && !(dbg_skip() & 4);
    if (ffn_mixed) {
      float* XFc = ctx->XF + (static_cast<uint8_t*>(Xc) - static_cast<uint8_t*>(ctx->X)) / 2;   // same row offset, fp32
      TRY_S(launch_layernorm(f.dt, ctx->X2, L.n1g, L.n1b, Bl * T, 256, Xc, st, 0, nullptr, nullptr, nullptr, XFc));
      FfnDesc d;
      d.X = XFc; d.M = Bl * T;
      d.W1 = L.ff1_32.w; d.b1 = L.ff1_32.bias; d.W2 = L.ff2_32.w; d.b2 = L.ff2_32.bias;
      d.gamma = L.n2g; d.beta = L.n2b;
      d.hidden = c.dim_feedforward;
      d.out = Xc;
      d.out_mode = 3;
      TRY_S(launch_ffn_fused(d, ctx->num_sms, st));
      const std::string nm = "enc" + std::to_string(i);
      TRY_S(f.tap(nm.c_str(), Xc, Bl * T * 256));
      continue;
    }
    TRY_S(f.ln(ctx->X2, L.n1g, L.n1b, Bl * T, Xc, exact_stream_on(ctx) ? 1 : 0));
    // the last encoder output feeds only the (3xTF32) cross-attention K/V projection: emit it pre-split
    const bool last = i == c.enc_layers - 1;
    const bool split = last && kv_split_on(ctx);
    void* XSc = split ? static_cast<uint8_t*>(ctx->XS) + (static_cast<uint8_t*>(Xc) - static_cast<uint8_t*>(ctx->X)) * 3
                      : nullptr;
    // feed-forward block + norm2 in one kernel (the 2048-wide hidden activation stays in tensor memory); the tap of
    // the last layer needs both output forms, so bring-up runs take the unfused path there
    // (small batches keep the two-GEMM path: one 128-row tile per CTA cannot fill the machine below ~74 tiles)
    if (dbg_skip() & 4) {
    } else if (ffn_fused_supported(f.dt, 256, c.dim_feedforward) && !(split && ctx->taps_enabled) && !ctx->calibrating &&
        (Bl * T + 127) / 128 >= ctx->num_sms / 2) {
      FfnDesc d;
      d.X = Xc; d.M = Bl * T;
      d.W1 = L.ff1.w; d.b1 = L.ff1.bias; d.W2 = L.ff2.w; d.b2 = L.ff2.bias;
      d.gamma = L.n2g; d.beta = L.n2b;
      d.hidden = c.dim_feedforward;
      d.out = split ? XSc : Xc;
      d.out_mode = split ? 2 : ((last || exact_stream_on(ctx)) ? 1 : 0);
      TRY_S(launch_ffn_fused(d, ctx->num_sms, st));
    } else {
      TRY_S(f.gemm(Xc, Bl * T, L.ff1, ctx->HID, c.dim_feedforward, true));
      TRY_S(f.gemm(ctx->HID, Bl * T, L.ff2, ctx->X2, 256, false, Xc, 256, 0, 0, true, 0, true));   // feeds LayerNorm only
      if (split) {
        TRY_S(f.ln(ctx->X2, L.n2g, L.n2b, Bl * T, XSc, 2));
        if (ctx->taps_enabled) TRY_S(f.ln(ctx->X2, L.n2g, L.n2b, Bl * T, Xc, 1));
      } else {
        TRY_S(f.ln(ctx->X2, L.n2g, L.n2b, Bl * T, Xc, (last || exact_stream_on(ctx)) ? 1 : 0));
      }
    }
    const std::string nm = "enc" + std::to_string(i);
    TRY_S(f.tap(nm.c_str(), Xc, Bl * T * 256));
  }

  return "";
}

// cross-attention K/V of all decoder layers from the encoder memory (last step of the trunk)
static std::string forward_kv(spe_ctx* ctx, int B, void* kv, cudaStream_t st) {
  Fwd f{ctx, st, B, ctx->dt, static_cast<long long>(dtype_size(ctx->dt))};
  const long long T = ctx->tokens;
  const int kvld = ctx->cfg.dec_layers * 512;
  const bool split = kv_split_on(ctx);
  const GemmW& w = split ? ctx->ca_kv_x3 : ctx->ca_kv_all;
  if (!split) TRY_S(f.calibrate_layer(ctx->X, static_cast<long long>(B) * T, 256, 256, w));
  GemmDesc d;
  d.mode = 0;
  d.A = split ? ctx->XS : ctx->X;
  d.M = static_cast<long long>(B) * T; d.K = w.K; d.lda = w.K;
  d.Wt = w.w; d.N = w.N;
  d.residual = ctx->ca_kv_addend; d.res_ld = kvld; d.res_mod = static_cast<int>(T); d.res_f32 = 1;
  d.out = kv; d.out_ld = kvld;
  d.round_out = split ? 0 : 1;   // the decoder attention rounds its own operands
  return launch_gemm(f.dt, d, ctx->num_sms, st);
}

// decoder and heads for the whole batch
static std::string forward_tail(spe_ctx* ctx, int B, void* kv, float* logits, float* points, float* logsig,
                                float* aux_logits, float* aux_points, cudaStream_t st, bool dec0_only = false) {
  const spe_config& c = ctx->cfg;
  Fwd f{ctx, st, B, ctx->dt, static_cast<long long>(dtype_size(ctx->dt))};
  const long long Bl = B;
  const long long T = ctx->tokens;
  const int Ti = static_cast<int>(T);

  // ---- decoder
  const int Q = c.num_queries, LD = c.dec_layers;
  const long long MQ = Bl * Q;
  const int kvld = LD * 512;
  // layer 0's self-attention block and query projection are constants of the model (see spe_ctx::dec0_tgt)
  const bool fold0 = ctx->dec0_fold && ctx->dec0_valid && !ctx->taps_enabled && !ctx->calibrating && !dec0_only;
  if (!fold0) SPE_CUDA_TRY(cudaMemsetAsync(ctx->TGT, 0, static_cast<size_t>(MQ * 256 * f.es), st));
  // fp32 storage: with 3xTF32 GEMMs the decoder state stays unrounded fp32 end to end; with plain TF32 GEMMs
  // (SPE_DEC_X3=0) whatever feeds a GEMM is rounded by its producer.  The attention products are error-compensated in
  // both cases for the SELF-attention (the learned query embeddings can drive its logits into the hundreds).
  const int ex = (f.dt == kTF32 && ctx->dec_x3) ? 1 : 0;
  const int ax3 = f.dt == kTF32 ? 1 : 0;
  for (int i = 0; i < LD; ++i) {
    const DecLayer& L = ctx->dec[i];
    const bool folded = i == 0 && fold0;
    if (!folded) {
      TRY_S(f.gemm(ctx->TGT, MQ, L.sa_qkv, ctx->DQKV, 768, false, L.sa_addend, 768, Q, 1));
      TRY_S(f.attn(ctx->DQKV, 768, f.col(ctx->DQKV, 256), 768, f.col(ctx->DQKV, 512), 768, ctx->DATT, Q, Q, ex, 0, ax3));
      TRY_S(f.gemm(ctx->DATT, MQ, L.sa_out, ctx->TGT2, 256, false, ctx->TGT, 256, 0, 0, true, 0, true));   // feeds LayerNorm only
      TRY_S(f.ln(ctx->TGT2, L.n1g, L.n1b, MQ, ctx->TGT, ex));
      TRY_S(f.gemm(ctx->TGT, MQ, L.ca_q, ctx->DQ, 256, false, L.ca_q_addend, 256, Q, 1));
      if (dec0_only) return "";
    }
    // cross-attention stays on the tcgen05 kernel (plain TF32 operands): its logits are bounded by the LayerNorm-ed
    // memory (tens at most), where TF32 is accurate to ~1e-2 of a logit -- measured harmless; 28 vs ~80 us per layer
    TRY_S(f.attn(folded ? ctx->dec0_q : ctx->DQ, 256, f.col(kv, i * 512), kvld, f.col(kv, i * 512 + 256), kvld, ctx->DATT, Q,
                 Ti, ex, 0, 0, folded));
    TRY_S(f.gemm(ctx->DATT, MQ, L.ca_out, ctx->TGT2, 256, false, folded ? ctx->dec0_tgt : ctx->TGT, 256, folded ? Q : 0, 0,
                 true, 0, true));
    // plain-TF32 feed-forward inside a 3xTF32 decoder: linear1 reads a rounded copy of norm2's output
    const bool ffn_plain = ex && !L.ff1.x3;
    if (ffn_plain) TRY_S(launch_layernorm(f.dt, ctx->TGT2, L.n2g, L.n2b, MQ, 256, ctx->TGT, st, ex, nullptr, nullptr, ctx->DQ));
    else TRY_S(f.ln(ctx->TGT2, L.n2g, L.n2b, MQ, ctx->TGT, ex));
    TRY_S(f.gemm(ffn_plain ? ctx->DQ : ctx->TGT, MQ, L.ff1, ctx->DHID, c.dim_feedforward, true));
    TRY_S(f.gemm(ctx->DHID, MQ, L.ff2, ctx->TGT2, 256, false, ctx->TGT, 256, 0, 0, true, 0, true));
    // norm3, and the decoder's shared output norm of it (return_intermediate), in one launch
    TRY_S(launch_layernorm(f.dt, ctx->TGT2, L.n3g, L.n3b, MQ, 256, ctx->TGT, st, ex, ctx->dn_g, ctx->dn_b,
                           f.col(ctx->HS, static_cast<long long>(i) * MQ * 256)));
  }
  TRY_S(f.tap("hs", ctx->HS, static_cast<long long>(LD) * MQ * 256));

  // ---- heads (all decoder layers only when the caller wants aux outputs)
  const bool want_aux = (aux_logits != nullptr && aux_points != nullptr && LD > 1);
  const int l_first = want_aux ? 0 : LD - 1;
  const long long rows = static_cast<long long>(LD - l_first) * MQ;
  const void* hs = f.col(ctx->HS, static_cast<long long>(l_first) * MQ * 256);
  TRY_S(f.gemm(hs, rows, ctx->pt0, ctx->H1, 256, true));
  TRY_S(f.gemm(ctx->H1, rows, ctx->pt1, ctx->H2, 256, true));
  const void* hs_last = f.col(ctx->HS, static_cast<long long>(LD - 1) * MQ * 256);
  const void* h2_last = f.col(ctx->H2, static_cast<long long>(LD - 1 - l_first) * MQ * 256);
  const bool sig = c.has_sigma && logsig != nullptr;
  if (sig) {
    TRY_S(f.gemm(hs_last, MQ, ctx->sg0, ctx->G1, 256, true));
    TRY_S(f.gemm(ctx->G1, MQ, ctx->sg1, ctx->G2, 256, true));
  }
  TRY_S(launch_head_final(f.dt, hs_last, h2_last, sig ? ctx->G2 : nullptr, MQ, ctx->cls_w, ctx->cls_b, ctx->pt2_w,
                          ctx->pt2_b, ctx->sg2_w, ctx->sg2_b, logits, points, logsig, st));
  if (want_aux) {
    TRY_S(launch_head_final(f.dt, ctx->HS, ctx->H2, nullptr, static_cast<long long>(LD - 1) * MQ, ctx->cls_w,
                            ctx->cls_b, ctx->pt2_w, ctx->pt2_b, nullptr, nullptr, aux_logits, aux_points, nullptr,
                            st));
  }
  return "";
}

#include "sa_model.inl"
void sa_release(spe_ctx* ctx) {
  delete ctx->sa;
  ctx->sa = nullptr;
}

// (re)compute the input-independent part of decoder layer 0 on one image's worth of rows (see spe_ctx::dec0_tgt)
static std::string use_workspace(spe_ctx* ctx, int set);
static std::string fold_dec0(spe_ctx* ctx) {
  ctx->dec0_valid = false;
  if (!ctx->dec0_fold || ctx->cfg.backbone == 2) return "";
  TRY_S(use_workspace(ctx, 0));
  const long long bytes = static_cast<long long>(ctx->cfg.num_queries) * 256 * static_cast<long long>(dtype_size(ctx->dt));
  cudaStream_t st = nullptr;
  TRY_S(forward_tail(ctx, 1, ctx->KV, nullptr, nullptr, nullptr, nullptr, nullptr, st, /*dec0_only=*/true));
  SPE_CUDA_TRY(cudaMemcpyAsync(ctx->dec0_tgt, ctx->TGT, static_cast<size_t>(bytes), cudaMemcpyDeviceToDevice, st));
  SPE_CUDA_TRY(cudaMemcpyAsync(ctx->dec0_q, ctx->DQ, static_cast<size_t>(bytes), cudaMemcpyDeviceToDevice, st));
  SPE_CUDA_TRY(cudaStreamSynchronize(st));
  ctx->dec0_valid = true;
  return "";
}

// Images per trunk pass.  Measured on B200 (B = 64, TF32): one pass over the whole batch is fastest (10.7 ms/step vs
// 11.5 ms at 32 and 17.4 ms at 8 images per pass) -- the per-launch fixed costs of ~90 small GEMMs outweigh the L2
// residency gained by shrinking the working set -- so chunking is opt-in (SPE_SUBBATCH) only.
static int chunk_images(const spe_ctx* ctx, int B) {
  if (ctx->taps_enabled || ctx->sub_batch <= 0) return B;
  return ctx->sub_batch < B ? ctx->sub_batch : B;
}

// select activation set `set` (allocated on first use -- never during a graph capture: the first call with any
// graph key runs eagerly)
static std::string use_workspace(spe_ctx* ctx, int set) {
  if (set < 0 || set >= 8) return "workspace set out of range";
  while (static_cast<int>(ctx->ws_sets.size()) <= set) {
    std::vector<void*> ns;
    for (const auto& f : ctx->ws_fields) {
      void* p = nullptr;
      if (cudaMalloc(&p, static_cast<size_t>(f.bytes > 0 ? f.bytes : 16)) != cudaSuccess) {
        cudaGetLastError();
        for (void* q : ns) cudaFree(q);
        return "out of device memory (activation set " + std::to_string(ctx->ws_sets.size()) + ")";
      }
      if (f.field == &ctx->SP) cudaMemset(p, 0, static_cast<size_t>(f.bytes));   // zero border of the stem input (crop writes the interior)
      ns.push_back(p);
    }
    for (void* q : ns) ctx->allocs.push_back(q);
    ctx->ws_sets.push_back(ns);
  }
  if (set != ctx->ws_current) {
    for (size_t i = 0; i < ctx->ws_fields.size(); ++i) *ctx->ws_fields[i].field = ctx->ws_sets[set][i];
    ctx->ws_current = set;
  }
  return "";
}

// parts: bit 0 = trunk (backbone, neck, encoder, decoder K/V), bit 1 = decoder + heads, on activation set `kv_slot`.
static std::string forward_schedule(spe_ctx* ctx, int parts, int kv_slot, const float* images, int B, float* logits,
                                    float* points, float* logsig, float* aux_logits, float* aux_points,
                                    cudaStream_t st) {
  TRY_S(use_workspace(ctx, kv_slot));
  struct PrefillScope {            // parts bit 2: the crop kernel already wrote this slot's stem input
    spe_ctx* c;
    ~PrefillScope() { c->stem_prefilled = false; }
  } prefill_scope{ctx};
  ctx->stem_prefilled = (parts & 4) != 0;
  parts &= 3;
  if (ctx->cfg.backbone == 2) {
    if (aux_logits != nullptr || aux_points != nullptr) return "the SA predictor returns its aux outputs through spe_forward_sa";
    if (parts != 3) return "the SA predictor runs as one schedule (parts = 3)";
    return sa_forward(ctx, images, B, logits, points, logsig, st);
  }
  void* kv = ctx->KV;
  if (parts & 1) {
    const long long es = static_cast<long long>(dtype_size(ctx->dt));
    const long long img_elems = 3ll * ctx->cfg.input_size * ctx->cfg.input_size;
    const int SB = chunk_images(ctx, B);
    for (int c0 = 0; c0 < B; c0 += SB) {
      const int nb = (B - c0) < SB ? (B - c0) : SB;
      void* Xc = static_cast<uint8_t*>(ctx->X) + static_cast<long long>(c0) * ctx->tokens * 256 * es;
      TRY_S(forward_trunk(ctx, images + c0 * img_elems, nb, Xc, st));
    }
    if (!(dbg_skip() & 8)) TRY_S(forward_kv(ctx, B, kv, st));
  }
  if ((parts & 2) && !(dbg_skip() & 1)) TRY_S(forward_tail(ctx, B, kv, logits, points, logsig, aux_logits, aux_points, st));
  return "";
}

// The schedule is a fixed sequence of ~150-600 launches: after one eager run per (batch, buffer set) it is captured
// into a CUDA graph and replayed, which removes the per-launch CPU cost (tensor-map encodes, launch calls).
static std::string forward_parts_on(spe_ctx* ctx, int parts, int kv_slot, const float* images, int B, float* logits,
                                    float* points, float* logsig, float* aux_logits, float* aux_points, cudaStream_t st);

std::string forward_parts(spe_ctx* ctx, int parts, int kv_slot, const float* images, int B, float* logits,
                          float* points, float* logsig, float* aux_logits, float* aux_points, cudaStream_t st) {
  const bool graphs_ok = ctx->use_graphs && !ctx->taps_enabled && !profile_timing_enabled();
  if (!graphs_ok)
    return forward_schedule(ctx, parts, kv_slot, images, B, logits, points, logsig, aux_logits, aux_points, st);
  const bool default_stream = st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread;
  if (!default_stream)
    return forward_parts_on(ctx, parts, kv_slot, images, B, logits, points, logsig, aux_logits, aux_points, st);
  if (ctx->own_stream == nullptr) {
    SPE_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    SPE_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming));
    SPE_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_out, cudaEventDisableTiming));
  }
  SPE_CUDA_TRY(cudaEventRecord(ctx->ev_in, st));                       // everything the caller enqueued so far ...
  SPE_CUDA_TRY(cudaStreamWaitEvent(ctx->own_stream, ctx->ev_in, 0));   // ... precedes the forward
  std::string s = forward_parts_on(ctx, parts, kv_slot, images, B, logits, points, logsig, aux_logits, aux_points,
                                   ctx->own_stream);
  SPE_CUDA_TRY(cudaEventRecord(ctx->ev_out, ctx->own_stream));
  SPE_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_out, 0));               // and whatever the caller enqueues next follows it
  return s;
}

static std::string forward_parts_on(spe_ctx* ctx, int parts, int kv_slot, const float* images, int B, float* logits,
                                    float* points, float* logsig, float* aux_logits, float* aux_points, cudaStream_t st) {
  GraphKey key{B, images, logits, points, logsig, aux_logits, aux_points, parts, kv_slot};
  if (!(parts & 1)) key.images = nullptr;
  if (!(parts & 2)) key.logits = key.points = key.logsig = key.aux_l = key.aux_p = nullptr;
  for (auto& g : ctx->graphs) {
    if (!(g.key == key)) continue;
    if (g.failed)       // this key could not be captured: eager, without giving up on the other keys
      return forward_schedule(ctx, parts, kv_slot, images, B, logits, points, logsig, aux_logits, aux_points, st);
    if (g.exec == nullptr) {
      // second call with this key: capture
      long long before[kNumFamilies];
      profile_peek_launches(before);
      cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
      if (e != cudaSuccess) {
        cudaGetLastError();
        g.failed = true;
        return forward_schedule(ctx, parts, kv_slot, images, B, logits, points, logsig, aux_logits, aux_points, st);
      }
      std::string s = forward_schedule(ctx, parts, kv_slot, images, B, logits, points, logsig, aux_logits, aux_points, st);
      cudaGraph_t graph = nullptr;
      e = cudaStreamEndCapture(st, &graph);
      if (!s.empty() || e != cudaSuccess || graph == nullptr) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        g.failed = true;
        if (!s.empty()) return s;
        // nothing ran during the broken capture: run this call eagerly
        return forward_schedule(ctx, parts, kv_slot, images, B, logits, points, logsig, aux_logits, aux_points, st);
      }
      e = cudaGraphInstantiate(&g.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) {
        g.exec = nullptr; g.failed = true; cudaGetLastError();
        return forward_schedule(ctx, parts, kv_slot, images, B, logits, points, logsig, aux_logits, aux_points, st);
      }
      long long after[kNumFamilies];
      profile_peek_launches(after);
      for (int i = 0; i < kNumFamilies; ++i) g.launches[i] = after[i] - before[i];
      profile_add_launches(g.launches, -1);   // the capture pass itself launched nothing
    }
    SPE_CUDA_TRY(cudaGraphLaunch(g.exec, st));
    profile_add_launches(g.launches, 1);
    return "";
  }
  if (ctx->use_graphs) {
    if (ctx->graphs.size() >= 12) {              // bounded cache: drop the oldest
      if (ctx->graphs.front().exec) cudaGraphExecDestroy(ctx->graphs.front().exec);
      ctx->graphs.erase(ctx->graphs.begin());
    }
    GraphEntry ge;
    ge.key = key;
    ctx->graphs.push_back(ge);
  }
  // first call: eager
  return forward_schedule(ctx, parts, kv_slot, images, B, logits, points, logsig, aux_logits, aux_points, st);
}

// One eager pass of the whole forward with the per-layer measurement switched on (see Fwd::calibrate_layer): layers are
// corrected in schedule order, so every layer is measured on the already-corrected output of its predecessors.
std::string calibrate_impl(spe_ctx* ctx, const float* images, int B, cudaStream_t st) {
  const bool taps = ctx->taps_enabled;
  ctx->taps_enabled = false;
  ctx->calibrating = true;
  std::string s = forward_schedule(ctx, 3, 0, images, B, ctx->p_logits, ctx->p_points,
                                   ctx->cfg.has_sigma ? ctx->p_logsig : nullptr, nullptr, nullptr, st);
  ctx->calibrating = false;
  ctx->taps_enabled = taps;
  if (!s.empty()) return s;
  SPE_CUDA_TRY(cudaStreamSynchronize(st));
  if (!ctx->calibrated) {
    // the schedule changes with the first calibration (unrounded residual stream): drop the graphs captured before it
    for (auto& g : ctx->graphs)
      if (g.exec) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
  }
  ctx->calibrated = true;
  return fold_dec0(ctx);
}

std::string forward_impl(spe_ctx* ctx, const float* images, int B, float* logits, float* points, float* logsig,
                         float* aux_logits, float* aux_points, cudaStream_t st) {
  return forward_parts(ctx, 3, 0, images, B, logits, points, logsig, aux_logits, aux_points, st);
}

}  // namespace spe

// ------------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------------
extern "C" {

const char* spe_global_last_error(void) { return g_last_error.c_str(); }

const char* spe_last_error(const spe_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int spe_create(const spe_config* cfg, int device, spe_ctx** out) {
  if (!cfg || !out) return fail(nullptr, SPE_ERR_INVALID, "spe_create: null argument");
  *out = nullptr;
  if (cfg->hidden_dim != 256 || cfg->nheads != 8)
    return fail(nullptr, SPE_ERR_INVALID, "spe_create: only hidden_dim=256, nheads=8 (head_dim 32) is built");
  if (cfg->input_size <= 0 || cfg->input_size % 32 != 0)
    return fail(nullptr, SPE_ERR_INVALID, "spe_create: input_size must be a positive multiple of 32");
  if (cfg->backbone < 0 || cfg->backbone > 2) return fail(nullptr, SPE_ERR_INVALID, "spe_create: backbone must be 0, 1 or 2");
  if (cfg->backbone == 2) {
    if (cfg->precision != 0 || cfg->has_sigma != 1 || cfg->enc_layers != 1)
      return fail(nullptr, SPE_ERR_INVALID, "spe_create: the SA predictor needs precision 0 (fp32 / TF32), has_sigma 1, enc_layers 1");
    if (cfg->input_size > 256)   // the implicit-GEMM convolution takes output rows of up to 128 pixels: conv1_2 / conv1_3 run at R / 2
      return fail(nullptr, SPE_ERR_INVALID, "spe_create: the SA predictor is built for inputs up to 256 x 256 (the recipe's eval_spatial_size)");
    const int h8 = cfg->input_size / 8;
    if (cfg->num_queries > h8 * h8 + (h8 / 2) * (h8 / 2) + (h8 / 4) * (h8 / 4) || cfg->num_queries % 2)
      return fail(nullptr, SPE_ERR_INVALID, "spe_create: the SA predictor needs an even num_queries <= the number of anchors");
  }
  if (cfg->precision != 0 && cfg->precision != 1) return fail(nullptr, SPE_ERR_INVALID, "spe_create: precision must be 0 or 1");
  if (cfg->num_queries <= 0 || cfg->enc_layers <= 0 || cfg->dec_layers <= 0 || cfg->max_batch <= 0 ||
      cfg->dim_feedforward <= 0 || cfg->dim_feedforward % 64 != 0)
    return fail(nullptr, SPE_ERR_INVALID, "spe_create: bad layer/query/batch configuration");
  const int featH = cfg->input_size / (cfg->backbone == 0 ? 8 : 16);
  if (featH > 128) return fail(nullptr, SPE_ERR_INVALID, "spe_create: feature map wider than 128 is not built");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return fail(nullptr, SPE_ERR_DEVICE, std::string("spe_create: no CUDA device (") + cudaGetErrorString(e) +
                                             "); this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(nullptr, SPE_ERR_INVALID, "spe_create: bad device index");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, SPE_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, SPE_ERR_DEVICE, std::string("spe_create: device '") + prop.name +
                                             "' is not sm_100; kernels are built for sm_100a only");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, SPE_ERR_CUDA, cudaGetErrorString(e));
  spe_ctx* ctx = new spe_ctx();
  ctx->cfg = *cfg;
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->dt = cfg->precision == 0 ? kTF32 : kBF16;
  ctx->featH = featH;
  ctx->tokens = featH * featH;
  ctx->featC = cfg->backbone == 0 ? 512 : 1024;
  if (const char* e = getenv("SPE_NO_GRAPH")) ctx->use_graphs = !(e[0] == '1');
  if (const char* e = getenv("SPE_SUBBATCH")) ctx->sub_batch = atoi(e);
  std::string s = alloc_workspace(ctx);
  if (!s.empty()) {
    fail(nullptr, SPE_ERR_CUDA, "spe_create: " + s);
    spe_destroy(ctx);
    return SPE_ERR_CUDA;
  }
  *out = ctx;
  return SPE_OK;
}

void spe_destroy(spe_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  pipeline_release(ctx);
  jpeg_release(ctx);
  for (auto& g : ctx->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (ctx->own_stream) { cudaStreamSynchronize(ctx->own_stream); cudaStreamDestroy(ctx->own_stream); }
  if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
  if (ctx->ev_out) cudaEventDestroy(ctx->ev_out);
  for (void* p : ctx->allocs) cudaFree(p);
  if (ctx->frames_dev) cudaFree(ctx->frames_dev);
  sa_release(ctx);
  delete ctx;
}

int spe_load_weights(spe_ctx* ctx, const spe_tensor_desc* tensors, int n) {
  if (!ctx || !tensors || n <= 0) return fail(ctx, SPE_ERR_INVALID, "spe_load_weights: null argument");
  cudaSetDevice(ctx->device);
  WeightSource ws;
  for (int i = 0; i < n; ++i) {
    if (!tensors[i].name || !tensors[i].data || tensors[i].ndim < 0 || tensors[i].ndim > 4)
      return fail(ctx, SPE_ERR_INVALID, "spe_load_weights: malformed tensor descriptor");
    HostTensor t;
    t.data = tensors[i].data;
    for (int k = 0; k < tensors[i].ndim; ++k) t.shape.push_back(tensors[i].shape[k]);
    ws.t[tensors[i].name] = t;
  }
  // NB: re-loading leaks the previous device weights until spe_destroy (weights are loaded once per ctx in practice)
  for (auto& g : ctx->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  ctx->graphs.clear();               // captured launches point at the previous weights
  ctx->calibrated = false;
  std::string s = load_weights_impl(ctx, ws);
  if (!s.empty()) return fail(ctx, SPE_ERR_WEIGHTS, "spe_load_weights: " + s);
  ctx->weights_loaded = true;
  s = fold_dec0(ctx);
  if (!s.empty()) return fail(ctx, SPE_ERR_CUDA, "spe_load_weights: " + s);
  return SPE_OK;
}

int spe_sync(spe_ctx* ctx, void* stream) {
  if (!ctx) return fail(nullptr, SPE_ERR_INVALID, "spe_sync: null ctx");
  cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(ctx, SPE_ERR_CUDA, std::string("spe_sync: ") + cudaGetErrorString(e));
  return SPE_OK;
}

int spe_forward(spe_ctx* ctx, const float* images_dev, int B, float* logits_dev, float* points_dev,
                float* log_sigma_dev, float* aux_logits_dev, float* aux_points_dev, void* stream) {
  if (!ctx) return fail(nullptr, SPE_ERR_INVALID, "spe_forward: null ctx");
  if (!ctx->weights_loaded) return fail(ctx, SPE_ERR_STATE, "spe_forward: call spe_load_weights first");
  if (!images_dev || !logits_dev || !points_dev) return fail(ctx, SPE_ERR_INVALID, "spe_forward: null buffer");
  if (B <= 0 || B > ctx->cfg.max_batch)
    return fail(ctx, SPE_ERR_INVALID, "spe_forward: batch " + std::to_string(B) + " outside [1, max_batch=" +
                                          std::to_string(ctx->cfg.max_batch) + "]");
  if (log_sigma_dev && !ctx->cfg.has_sigma)
    return fail(ctx, SPE_ERR_INVALID, "spe_forward: log_sigma requested but the model has no sigma head");
  cudaSetDevice(ctx->device);
  std::string s = forward_impl(ctx, images_dev, B, logits_dev, points_dev, log_sigma_dev, aux_logits_dev,
                               aux_points_dev, static_cast<cudaStream_t>(stream));
  if (!s.empty()) return fail(ctx, SPE_ERR_CUDA, "spe_forward: " + s);
  return SPE_OK;
}

int spe_forward_sa(spe_ctx* ctx, const float* images_dev, int B, float* logits_dev, float* points_dev, float* log_sigma_dev,
                   float* aux_logits_dev, float* aux_points_dev, float* aux_log_sigma_dev, int32_t* topk_idx_dev,
                   const int32_t* topk_override_dev, void* stream) {
  if (!ctx) return fail(nullptr, SPE_ERR_INVALID, "spe_forward_sa: null ctx");
  if (ctx->cfg.backbone != 2 || ctx->sa == nullptr) return fail(ctx, SPE_ERR_INVALID, "spe_forward_sa: the ctx was not created with backbone 2");
  if (!ctx->weights_loaded) return fail(ctx, SPE_ERR_STATE, "spe_forward_sa: call spe_load_weights first");
  if (!images_dev || !logits_dev || !points_dev) return fail(ctx, SPE_ERR_INVALID, "spe_forward_sa: null buffer");
  if (B <= 0 || B > ctx->cfg.max_batch) return fail(ctx, SPE_ERR_INVALID, "spe_forward_sa: batch outside [1, max_batch]");
  cudaSetDevice(ctx->device);
  std::string s = use_workspace(ctx, 0);
  if (s.empty()) {
    SaModel& m = *ctx->sa;
    m.aux_logits = aux_logits_dev; m.aux_points = aux_points_dev; m.aux_logsig = aux_log_sigma_dev;
    m.topk_out = topk_idx_dev; m.topk_in = topk_override_dev;
    s = sa_forward(ctx, images_dev, B, logits_dev, points_dev, log_sigma_dev, static_cast<cudaStream_t>(stream));
    m.aux_logits = m.aux_points = m.aux_logsig = nullptr;
    m.topk_out = nullptr; m.topk_in = nullptr;
  }
  if (!s.empty()) return fail(ctx, SPE_ERR_CUDA, "spe_forward_sa: " + s);
  return SPE_OK;
}

int spe_calibrate(spe_ctx* ctx, const float* images_dev, int B, void* stream) {
  if (!ctx) return fail(nullptr, SPE_ERR_INVALID, "spe_calibrate: null ctx");
  if (!ctx->weights_loaded) return fail(ctx, SPE_ERR_STATE, "spe_calibrate: call spe_load_weights first");
  if (!images_dev) return fail(ctx, SPE_ERR_INVALID, "spe_calibrate: null buffer");
  if (B <= 0 || B > ctx->cfg.max_batch) return fail(ctx, SPE_ERR_INVALID, "spe_calibrate: batch outside [1, max_batch]");
  if (pipeline_busy(ctx)) return fail(ctx, SPE_ERR_STATE, "spe_calibrate: collect every pipeline slot first (the biases it rewrites are shared)");
  cudaSetDevice(ctx->device);
  std::string s = calibrate_impl(ctx, images_dev, B, static_cast<cudaStream_t>(stream));
  if (!s.empty()) return fail(ctx, SPE_ERR_CUDA, "spe_calibrate: " + s);
  return SPE_OK;
}

int spe_is_calibrated(const spe_ctx* ctx) { return ctx && ctx->calibrated ? 1 : 0; }

int spe_debug_graph_stats(const spe_ctx* ctx, int* captured, int* failed) {
  if (!ctx || !captured || !failed) return SPE_ERR_INVALID;
  *captured = *failed = 0;
  for (const auto& g : ctx->graphs) { if (g.exec) ++*captured; if (g.failed) ++*failed; }
  return SPE_OK;
}

int spe_debug_enable_taps(spe_ctx* ctx, int enable) {
  if (!ctx) return SPE_ERR_INVALID;
  ctx->taps_enabled = enable != 0;
  return SPE_OK;
}

long long spe_debug_read_tap(spe_ctx* ctx, const char* name, void* host_out, long long max_bytes) {
  if (!ctx || !name) return SPE_ERR_INVALID;
  auto it = ctx->taps.find(name);
  if (it == ctx->taps.end()) return fail(ctx, SPE_ERR_INVALID, std::string("no tap named ") + name);
  if (host_out == nullptr) return it->second.bytes;
  const long long n = it->second.bytes < max_bytes ? it->second.bytes : max_bytes;
  cudaError_t e = cudaMemcpy(host_out, it->second.buf, static_cast<size_t>(n), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail(ctx, SPE_ERR_CUDA, cudaGetErrorString(e));
  return n;
}

}  // extern "C"

// accessors used by api.cu (spe_ctx is private to this translation unit)
namespace spe {
struct PipelineBuffers {
  uint8_t** frames_dev; long long* frames_cap; int32_t* boxes_dev; float* images_dev; float* logits; float* points;
  float* logsig; double* quat; double* tvec; int32_t* assign; int32_t* status; int device; int max_batch; int R; int Q;
  int has_sigma; const float* ov_logits; const float* ov_points; long long* last_h2d_bytes;
  const int32_t* ov_boxes;
};
PipelineBuffers pipeline_buffers(spe_ctx* ctx) {
  return PipelineBuffers{&ctx->frames_dev, &ctx->frames_cap, ctx->boxes_dev, ctx->images_dev, ctx->p_logits,
                         ctx->p_points,    ctx->p_logsig,    ctx->p_quat,    ctx->p_tvec,     ctx->p_assign,
                         ctx->p_status,    ctx->device,      ctx->cfg.max_batch, ctx->cfg.input_size,
                         ctx->cfg.num_queries, ctx->cfg.has_sigma, ctx->ov_logits, ctx->ov_points,
                         &ctx->last_h2d, ctx->ov_boxes};
}
long long last_h2d_bytes(spe_ctx* ctx) { return ctx->last_h2d; }
void set_pnp_override(spe_ctx* ctx, const float* logits, const float* points, const int32_t* boxes) {
  ctx->ov_logits = logits;
  ctx->ov_points = points;
  ctx->ov_boxes = boxes;
}
int set_error(spe_ctx* ctx, int code, const std::string& msg) { return fail(ctx, code, msg); }
// the stem input buffer of activation set `slot` when the crop kernel may write it directly (else null)
void* stem_input_buffer(spe_ctx* ctx, int slot, int* is_bf16) {
  if (!ctx->crop_writes_stem || ctx->cfg.backbone == 2 || !ctx->stem_windowed || ctx->head_chunk > 0 || ctx->taps_enabled ||
      ctx->calibrating || ctx->sub_batch > 0)
    return nullptr;
  if (!use_workspace(ctx, slot).empty()) return nullptr;
  *is_bf16 = ctx->dt == kBF16 ? 1 : 0;
  return ctx->SP;
}
// the forward (parts: 1 = trunk + decoder K/V, 2 = decoder + heads, 3 = both) on activation set `kv_slot`
int forward_half(spe_ctx* ctx, int parts, int kv_slot, const float* images, int B, float* logits, float* points,
                 float* logsig, cudaStream_t st) {
  if (!ctx->weights_loaded) return fail(ctx, SPE_ERR_STATE, "pipeline: call spe_load_weights first");
  std::string s = forward_parts(ctx, parts, kv_slot, images, B, logits, points, logsig, nullptr, nullptr, st);
  if (!s.empty()) return fail(ctx, SPE_ERR_CUDA, "pipeline forward: " + s);
  return SPE_OK;
}
}  // namespace spe
