#include <cmath>
// C-ABI entry points for the crop and pose stages, the host-buffer pipeline and the bring-up hooks.
// (ctx lifetime, weights and spe_forward live in model.cu.)
#include "spe_internal.h"
#include "profile.h"
#include "../../include/spe.h"

#include <string.h>

#include <map>
#include <string>
#include <vector>

// spe_ctx is defined in model.cu; the pipeline needs a few of its fields, exposed through these accessors.
namespace spe {
struct PipelineBuffers {
  uint8_t** frames_dev; long long* frames_cap; int32_t* boxes_dev; float* images_dev; float* logits; float* points;
  float* logsig; double* quat; double* tvec; int32_t* assign; int32_t* status; int device; int max_batch; int R; int Q;
  int has_sigma; const float* ov_logits; const float* ov_points; long long* last_h2d_bytes;
  const int32_t* ov_boxes;
};
PipelineBuffers pipeline_buffers(spe_ctx* ctx);
long long last_h2d_bytes(spe_ctx* ctx);
void set_pnp_override(spe_ctx* ctx, const float* logits, const float* points, const int32_t* boxes);
int set_error(spe_ctx* ctx, int code, const std::string& msg);
int forward_half(spe_ctx* ctx, int parts, int kv_slot, const float* images, int B, float* logits, float* points,
                 float* logsig, cudaStream_t st);
void* stem_input_buffer(spe_ctx* ctx, int slot, int* is_bf16);
}  // namespace spe

using namespace spe;

namespace {
constexpr int kPipeSlots = SPE_PIPELINE_SLOTS;
struct PipeSlot {
  cudaStream_t stream = nullptr;          // upload stream of this slot
  cudaStream_t compute = nullptr;         // crop -> forward -> assignment + PnP -> result download of this slot
  cudaEvent_t upload_done = nullptr, done = nullptr;
  float* images_dev = nullptr;
  uint8_t* frames_dev = nullptr;
  long long frames_cap = 0;
  int32_t *boxes_dev = nullptr, *status_dev = nullptr, *assign_dev = nullptr;
  double *quat_dev = nullptr, *tvec_dev = nullptr;
  float *logits_dev = nullptr, *points_dev = nullptr, *logsig_dev = nullptr;   // network outputs of this slot's batch
  int32_t *boxes_h = nullptr, *status_h = nullptr;   // pinned
  double *quat_h = nullptr, *tvec_h = nullptr;       // pinned
  bool busy = false;
  bool ready = false;                                // every buffer below was allocated
  bool host_boxes = false;                           // boxes_h holds this batch's crop boxes (host submissions only)
  int B = 0;
};
// Every slot is a complete, independent instance of the path: its own streams, frame / result buffers and activation
// set (model.cu: use_workspace).  Batches submitted on different slots share nothing but the weights, so the GPU runs
// them next to each other: the ~100 small latency-bound launches of one batch's decoder / pose stage and the drained
// last wave of every persistent GEMM are filled with another batch's tiles.
struct Pipe {
  PipeSlot slot[kPipeSlots];
  cudaEvent_t caller_ready = nullptr;
};
std::map<spe_ctx*, Pipe*> g_pipes;
Pipe& pipe_of(spe_ctx* ctx) {
  auto it = g_pipes.find(ctx);
  if (it == g_pipes.end()) it = g_pipes.emplace(ctx, new Pipe()).first;
  return *it->second;
}
}  // namespace

static void slot_free(PipeSlot& S) {
  if (S.stream) cudaStreamSynchronize(S.stream);
  if (S.compute) cudaStreamSynchronize(S.compute);
  if (S.frames_dev) cudaFree(S.frames_dev);
  if (S.boxes_dev) cudaFree(S.boxes_dev);
  if (S.status_dev) cudaFree(S.status_dev);
  if (S.assign_dev) cudaFree(S.assign_dev);
  if (S.quat_dev) cudaFree(S.quat_dev);
  if (S.tvec_dev) cudaFree(S.tvec_dev);
  if (S.logits_dev) cudaFree(S.logits_dev);
  if (S.points_dev) cudaFree(S.points_dev);
  if (S.logsig_dev) cudaFree(S.logsig_dev);
  if (S.boxes_h) cudaFreeHost(S.boxes_h);
  if (S.status_h) cudaFreeHost(S.status_h);
  if (S.quat_h) cudaFreeHost(S.quat_h);
  if (S.tvec_h) cudaFreeHost(S.tvec_h);
  if (S.images_dev) cudaFree(S.images_dev);
  if (S.upload_done) cudaEventDestroy(S.upload_done);
  if (S.done) cudaEventDestroy(S.done);
  if (S.stream) cudaStreamDestroy(S.stream);
  if (S.compute) cudaStreamDestroy(S.compute);
  S = PipeSlot();
}

namespace spe {
bool pipeline_busy(spe_ctx* ctx) {
  auto it = g_pipes.find(ctx);
  if (it == g_pipes.end()) return false;
  for (const PipeSlot& S : it->second->slot)
    if (S.busy) return true;
  return false;
}
void pipeline_release(spe_ctx* ctx) {   // called by spe_destroy
  auto it = g_pipes.find(ctx);
  if (it == g_pipes.end()) return;
  for (PipeSlot& S : it->second->slot) slot_free(S);
  Pipe* P = it->second;
  if (P->caller_ready) cudaEventDestroy(P->caller_ready);
  delete it->second;
  g_pipes.erase(it);
}
}  // namespace spe

extern "C" {

int spe_clip_boxes(const double* det, int B, int32_t* boxes) {
  if (!det || !boxes || B < 0) return set_error(nullptr, SPE_ERR_INVALID, "spe_clip_boxes: null argument");
  for (int i = 0; i < 4 * B; ++i)     // int() of NaN / inf / 1e300 is undefined behaviour: refuse instead
    if (!std::isfinite(det[i]) || std::fabs(det[i]) > 1e8)
      return set_error(nullptr, SPE_ERR_INVALID, "spe_clip_boxes: non-finite or absurd detector box");
  for (int i = 0; i < B; ++i) {
    // SpeedSubmission.generate_clip_bbox, RV/datasets/speed.py:92-108 (float64; int() truncates toward zero)
    const double x1 = det[4 * i + 0], y1 = det[4 * i + 1], x2 = det[4 * i + 2], y2 = det[4 * i + 3];
    const double bw = x2 - x1, bh = y2 - y1;
    const double scale = (bw > bh ? bw : bh) * 1.2;
    const double xc = (x1 + x2) / 2, yc = (y1 + y2) / 2;
    const double half = scale / 2;
    const int32_t cx1 = static_cast<int32_t>(xc - half), cy1 = static_cast<int32_t>(yc - half);
    const int32_t s = static_cast<int32_t>(scale);
    boxes[4 * i + 0] = cx1; boxes[4 * i + 1] = cy1; boxes[4 * i + 2] = cx1 + s; boxes[4 * i + 3] = cy1 + s;
  }
  return SPE_OK;
}

int spe_clip_boxes_val(const double* det, int B, int W, int H, double* fbox, int32_t* ibox) {
  if (!det || !fbox || !ibox || B < 0) return set_error(nullptr, SPE_ERR_INVALID, "spe_clip_boxes_val: null argument");
  for (int i = 0; i < 4 * B; ++i)
    if (!std::isfinite(det[i]) || std::fabs(det[i]) > 1e8)
      return set_error(nullptr, SPE_ERR_INVALID, "spe_clip_boxes_val: non-finite or absurd detector box");
  for (int i = 0; i < B; ++i) {
    // SpeedTrain.generate_clip_bbox_val, RV/datasets/speed.py:246-260 (float64; x clipped to [0,W], y to [0,H])
    const double x1 = det[4 * i + 0], y1 = det[4 * i + 1], x2 = det[4 * i + 2], y2 = det[4 * i + 3];
    const double bw = x2 - x1, bh = y2 - y1;
    const double scale = (bw > bh ? bw : bh) * 1.2;
    const double xc = (x1 + x2) / 2, yc = (y1 + y2) / 2;
    const double half = scale / 2;
    double b[4] = {xc - half, yc - half, xc + half, yc + half};
    for (int k = 0; k < 4; ++k) {
      const double hi = (k & 1) ? static_cast<double>(H) : static_cast<double>(W);
      b[k] = b[k] < 0.0 ? 0.0 : (b[k] > hi ? hi : b[k]);
      fbox[4 * i + k] = b[k];
      // PIL.Image.crop: int(round(v)) -- Python's round() is round-half-to-even, like nearbyint in the default mode
      ibox[4 * i + k] = static_cast<int32_t>(nearbyint(b[k]));
    }
  }
  return SPE_OK;
}

int spe_crop_resize_norm(spe_ctx* ctx, const uint8_t* frames_dev, int H, int W, long long pitch,
                         long long frame_stride, const int32_t* boxes_dev, int B, int R, float* out_nchw_dev,
                         void* stream) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_crop_resize_norm: null ctx");
  if (!frames_dev || !boxes_dev || !out_nchw_dev || H <= 0 || W <= 0 || pitch < W || B < 0)
    return set_error(ctx, SPE_ERR_INVALID, "spe_crop_resize_norm: bad argument");
  std::string s = launch_crop_resize_norm(frames_dev, H, W, pitch, frame_stride, boxes_dev, B, R, out_nchw_dev,
                                          static_cast<cudaStream_t>(stream));
  if (!s.empty()) return set_error(ctx, SPE_ERR_CUDA, "spe_crop_resize_norm: " + s);
  return SPE_OK;
}

int spe_assign_pnp(spe_ctx* ctx, const float* logits_dev, const float* points_dev, const float* log_sigma_dev,
                   const int32_t* boxes_dev, int B, int Q, const spe_pnp_params* params, double* quat_dev,
                   double* tvec_dev, int32_t* assign_dev, int32_t* status_dev, float* probs_dev,
                   float* points_px_dev, float* sigmas_dev, int32_t* inlier_mask_dev, void* stream) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_assign_pnp: null ctx");
  if (!logits_dev || !points_dev || !boxes_dev || !quat_dev || !tvec_dev || !assign_dev || !status_dev || !params)
    return set_error(ctx, SPE_ERR_INVALID, "spe_assign_pnp: null buffer");
  if (params->weighted && !log_sigma_dev)
    return set_error(ctx, SPE_ERR_INVALID, "spe_assign_pnp: weighted solve needs log_sigma");
  PnpDesc d{};
  d.logits = logits_dev; d.points = points_dev; d.logsig = log_sigma_dev; d.boxes = boxes_dev;
  d.boxes_f = params->float_boxes_dev;
  d.B = B; d.Q = Q;
  d.reproj_thresh = params->reproj_thresh;
  d.reproj_dev = params->reproj_thresh_dev;
  d.weighted = params->weighted;
  d.reject = params->reject;
  d.reject_rms_px = params->reject_rms_px > 0 ? params->reject_rms_px : 5.0f;
  d.reject_sigma = params->reject_sigma_px > 0 ? params->reject_sigma_px : 12.0f;
  d.post_processed = params->inputs_post_processed; d.sigma_px_scale = params->sigma_px_scale;
  d.quat = quat_dev; d.tvec = tvec_dev; d.assign = assign_dev; d.status = status_dev;
  d.probs = probs_dev; d.points_px = points_px_dev; d.sigmas = sigmas_dev; d.inlier_mask = inlier_mask_dev;
  std::string s = launch_assign_pnp(d, static_cast<cudaStream_t>(stream));
  if (!s.empty()) return set_error(ctx, SPE_ERR_CUDA, "spe_assign_pnp: " + s);
  return SPE_OK;
}

int spe_speed_score(spe_ctx* ctx, const double* quat_pr_dev, const double* tvec_pr_dev, const double* quat_gt_dev,
                    const double* tvec_gt_dev, int B, double* score_t_dev, double* score_q_dev, void* stream) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_speed_score: null ctx");
  if (!quat_pr_dev || !tvec_pr_dev || !quat_gt_dev || !tvec_gt_dev || !score_t_dev || !score_q_dev)
    return set_error(ctx, SPE_ERR_INVALID, "spe_speed_score: null buffer");
  std::string s = launch_speed_score(quat_pr_dev, tvec_pr_dev, quat_gt_dev, tvec_gt_dev, B, score_t_dev, score_q_dev,
                                     static_cast<cudaStream_t>(stream));
  if (!s.empty()) return set_error(ctx, SPE_ERR_CUDA, "spe_speed_score: " + s);
  return SPE_OK;
}

int spe_ensemble_pnp(spe_ctx* ctx, const float* logits_dev, const float* points_dev, const int32_t* boxes_dev,
                     int num_models, int B, int Q, const spe_pnp_params* params, double* quat_dev, double* tvec_dev,
                     int32_t* count_dev, int32_t* status_dev, float* pooled_px_dev, int32_t* inlier_mask_dev,
                     void* stream) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_ensemble_pnp: null ctx");
  if (!logits_dev || !points_dev || !boxes_dev || !quat_dev || !tvec_dev || !count_dev || !status_dev || !params)
    return set_error(ctx, SPE_ERR_INVALID, "spe_ensemble_pnp: null buffer");
  if (num_models < 1) return set_error(ctx, SPE_ERR_INVALID, "spe_ensemble_pnp: needs at least one model");
  if (params->weighted) return set_error(ctx, SPE_ERR_INVALID, "spe_ensemble_pnp: the ensemble solver has no sigma-weighted form");
  PnpDesc d{};
  d.logits = logits_dev; d.points = points_dev; d.logsig = nullptr; d.boxes = boxes_dev;
  d.boxes_f = params->float_boxes_dev;
  d.B = B; d.Q = Q; d.num_models = num_models;
  d.reproj_thresh = params->reproj_thresh;
  d.reproj_dev = params->reproj_thresh_dev;
  d.reject = params->reject;
  d.reject_rms_px = params->reject_rms_px > 0 ? params->reject_rms_px : 5.0f;
  d.reject_sigma = params->reject_sigma_px > 0 ? params->reject_sigma_px : 12.0f;
  d.post_processed = params->inputs_post_processed; d.sigma_px_scale = params->sigma_px_scale;
  d.quat = quat_dev; d.tvec = tvec_dev; d.assign = count_dev; d.status = status_dev;
  d.pooled_px = pooled_px_dev; d.inlier_mask = inlier_mask_dev;
  std::string s = launch_assign_pnp(d, static_cast<cudaStream_t>(stream));
  if (!s.empty()) return set_error(ctx, SPE_ERR_CUDA, "spe_ensemble_pnp: " + s);
  return SPE_OK;
}

int spe_ms_deform_attn(spe_ctx* ctx, const float* value_dev, const int32_t* shapes_hw_host, int L, const float* loc_dev,
                       const float* attn_dev, const float* ref_dev, int ref_levels, int B, int Lq, int heads, int P,
                       int fused, float* out_dev, void* stream) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_ms_deform_attn: null ctx");
  if (!value_dev || !shapes_hw_host || !loc_dev || !attn_dev || !out_dev || heads <= 0)
    return set_error(ctx, SPE_ERR_INVALID, "spe_ms_deform_attn: null buffer");
  std::string s = launch_ms_deform_attn(value_dev, shapes_hw_host, L, loc_dev, attn_dev, ref_dev, ref_levels, B, Lq, heads,
                                        P, fused, out_dev, static_cast<cudaStream_t>(stream));
  if (!s.empty()) return set_error(ctx, SPE_ERR_INVALID, "spe_ms_deform_attn: " + s);
  return SPE_OK;
}

int spe_topk_queries(spe_ctx* ctx, const float* cls_dev, int B, int Lv, int C, int k, int32_t* idx_dev, float* vals_dev,
                     void* stream) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_topk_queries: null ctx");
  if (!cls_dev || !idx_dev || Lv <= 0 || C <= 0) return set_error(ctx, SPE_ERR_INVALID, "spe_topk_queries: bad argument");
  std::string s = launch_topk_queries(cls_dev, B, Lv, C, k, idx_dev, vals_dev, static_cast<cudaStream_t>(stream));
  if (!s.empty()) return set_error(ctx, SPE_ERR_INVALID, "spe_topk_queries: " + s);
  return SPE_OK;
}

int spe_gather_rows(spe_ctx* ctx, const float* src_dev, const int32_t* idx_dev, int B, int Lv, int k, int D,
                    float* out_dev, void* stream) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_gather_rows: null ctx");
  if (!src_dev || !idx_dev || !out_dev || D <= 0) return set_error(ctx, SPE_ERR_INVALID, "spe_gather_rows: bad argument");
  std::string s = launch_gather_rows(src_dev, idx_dev, B, Lv, k, D, out_dev, static_cast<cudaStream_t>(stream));
  if (!s.empty()) return set_error(ctx, SPE_ERR_INVALID, "spe_gather_rows: " + s);
  return SPE_OK;
}

int spe_run_batch_host(spe_ctx* ctx, const uint8_t* frames_host, int H, int W, const double* det_boxes_host, int B,
                       const spe_pnp_params* params, double* quat_host, double* tvec_host, int32_t* status_host,
                       int32_t* boxes_host, void* stream) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_run_batch_host: null ctx");
  if (!frames_host || !det_boxes_host || !params || !quat_host || !tvec_host || !status_host)
    return set_error(ctx, SPE_ERR_INVALID, "spe_run_batch_host: null buffer");
  PipelineBuffers pb = pipeline_buffers(ctx);
  if (B <= 0 || B > pb.max_batch) return set_error(ctx, SPE_ERR_INVALID, "spe_run_batch_host: batch outside [1, max_batch]");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaSetDevice(pb.device);
  const long long need = static_cast<long long>(B) * H * W;
  if (*pb.frames_cap < need) {
    if (*pb.frames_dev) cudaFree(*pb.frames_dev);
    *pb.frames_dev = nullptr;
    *pb.frames_cap = 0;
    const long long cap = static_cast<long long>(pb.max_batch) * H * W;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(pb.frames_dev), static_cast<size_t>(cap));
    if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string("frame buffer: ") + cudaGetErrorString(e));
    *pb.frames_cap = cap;
  }
  std::vector<int32_t> boxes(static_cast<size_t>(B) * 4);
  spe_clip_boxes(det_boxes_host, B, boxes.data());
  if (boxes_host) memcpy(boxes_host, boxes.data(), boxes.size() * sizeof(int32_t));
  // Upload only what the crop kernel can touch: the intersection of each crop box with its frame (the bicubic taps
  // are clamped to the crop canvas, so nothing outside the box is ever read).  The rectangle lands at its own
  // position inside the device frame, so the kernel needs no extra indirection.
  cudaError_t e = cudaSuccess;
  long long h2d = 0;
  for (int i = 0; i < B && e == cudaSuccess; ++i) {
    const int x0 = boxes[4 * i + 0] < 0 ? 0 : boxes[4 * i + 0], y0 = boxes[4 * i + 1] < 0 ? 0 : boxes[4 * i + 1];
    const int x1 = boxes[4 * i + 2] > W ? W : boxes[4 * i + 2], y1 = boxes[4 * i + 3] > H ? H : boxes[4 * i + 3];
    if (x1 <= x0 || y1 <= y0) continue;   // box entirely outside the frame: the crop is all padding
    const long long off = static_cast<long long>(i) * H * W + static_cast<long long>(y0) * W + x0;
    e = cudaMemcpy2DAsync(*pb.frames_dev + off, W, frames_host + off, W, static_cast<size_t>(x1 - x0),
                          static_cast<size_t>(y1 - y0), cudaMemcpyHostToDevice, st);
    h2d += static_cast<long long>(x1 - x0) * (y1 - y0);
  }
  *pb.last_h2d_bytes = h2d + static_cast<long long>(boxes.size() * sizeof(int32_t));
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(pb.boxes_dev, boxes.data(), boxes.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string("H2D: ") + cudaGetErrorString(e));
  int rc = spe_crop_resize_norm(ctx, *pb.frames_dev, H, W, W, static_cast<long long>(H) * W, pb.boxes_dev, B, pb.R,
                                pb.images_dev, stream);
  if (rc != SPE_OK) return rc;
  const bool sig = pb.has_sigma != 0;
  rc = spe_forward(ctx, pb.images_dev, B, pb.logits, pb.points, sig ? pb.logsig : nullptr, nullptr, nullptr, stream);
  if (rc != SPE_OK) return rc;
  spe_pnp_params pp = *params;
  if (!sig) pp.weighted = 0;
  // bench hook: with random-init weights every query collapses to one label and the solve would exit early, so the
  // benchmark may substitute resident synthetic keypoint sets for the pose stage (spe_debug_set_pnp_override)
  const float* pl = pb.ov_logits ? pb.ov_logits : pb.logits;
  const float* pp_pts = pb.ov_points ? pb.ov_points : pb.points;
  rc = spe_assign_pnp(ctx, pl, pp_pts, sig ? pb.logsig : nullptr, pb.ov_boxes ? pb.ov_boxes : pb.boxes_dev, B, pb.Q, &pp, pb.quat,
                      pb.tvec, pb.assign, pb.status, nullptr, nullptr, nullptr, nullptr, stream);
  if (rc != SPE_OK) return rc;
  e = cudaMemcpyAsync(quat_host, pb.quat, sizeof(double) * 4 * B, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tvec_host, pb.tvec, sizeof(double) * 3 * B, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(status_host, pb.status, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string("spe_run_batch_host: ") + cudaGetErrorString(e));
  return SPE_OK;
}

long long spe_last_h2d_bytes(spe_ctx* ctx) { return ctx ? last_h2d_bytes(ctx) : -1; }

// ---- multi-slot pipeline: whole batches in flight next to each other ------------------------------------------
static int pipe_prepare(spe_ctx* ctx, const char* who, int slot, int B, PipelineBuffers& pb, Pipe** Pp, PipeSlot** Sp) {
  if (slot < 0 || slot >= kPipeSlots)
    return set_error(ctx, SPE_ERR_INVALID, std::string(who) + ": slot outside [0, SPE_PIPELINE_SLOTS)");
  pb = pipeline_buffers(ctx);
  if (B <= 0 || B > pb.max_batch) return set_error(ctx, SPE_ERR_INVALID, std::string(who) + ": batch outside [1, max_batch]");
  cudaSetDevice(pb.device);
  Pipe& P = pipe_of(ctx);
  PipeSlot& S = P.slot[slot];
  if (S.busy) return set_error(ctx, SPE_ERR_STATE, std::string(who) + ": slot still in flight (collect it first)");
  cudaError_t e = cudaSuccess;
  if (!P.caller_ready) {
    e = cudaEventCreateWithFlags(&P.caller_ready, cudaEventDisableTiming);
    if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string("pipeline: ") + cudaGetErrorString(e));
  }
  if (!S.ready) {
    e = cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&S.compute, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.upload_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&S.images_dev, sizeof(float) * 3 * pb.R * pb.R * pb.max_batch);
    if (e == cudaSuccess) e = cudaMalloc(&S.boxes_dev, sizeof(int32_t) * 4 * pb.max_batch);
    if (e == cudaSuccess) e = cudaMalloc(&S.quat_dev, sizeof(double) * 4 * pb.max_batch);
    if (e == cudaSuccess) e = cudaMalloc(&S.tvec_dev, sizeof(double) * 3 * pb.max_batch);
    if (e == cudaSuccess) e = cudaMalloc(&S.status_dev, sizeof(int32_t) * pb.max_batch);
    if (e == cudaSuccess) e = cudaMalloc(&S.assign_dev, sizeof(int32_t) * 11 * pb.max_batch);
    if (e == cudaSuccess) e = cudaMalloc(&S.logits_dev, sizeof(float) * 12 * pb.Q * pb.max_batch);
    if (e == cudaSuccess) e = cudaMalloc(&S.points_dev, sizeof(float) * 2 * pb.Q * pb.max_batch);
    if (e == cudaSuccess) e = cudaMalloc(&S.logsig_dev, sizeof(float) * 2 * pb.Q * pb.max_batch);
    if (e == cudaSuccess) e = cudaMallocHost(&S.boxes_h, sizeof(int32_t) * 4 * pb.max_batch);
    if (e == cudaSuccess) e = cudaMallocHost(&S.quat_h, sizeof(double) * 4 * pb.max_batch);
    if (e == cudaSuccess) e = cudaMallocHost(&S.tvec_h, sizeof(double) * 3 * pb.max_batch);
    if (e == cudaSuccess) e = cudaMallocHost(&S.status_h, sizeof(int32_t) * pb.max_batch);
    if (e != cudaSuccess) {
      // a half-built slot must not look initialised to the next submit: release what was allocated
      cudaGetLastError();
      slot_free(S);
      return set_error(ctx, SPE_ERR_CUDA, std::string("pipeline slot: ") + cudaGetErrorString(e));
    }
    S.ready = true;
  }
  *Pp = &P;
  *Sp = &S;
  return SPE_OK;
}

// the whole path of one batch on the slot's compute stream (after `ready`), on the slot's activation set
static int pipe_enqueue(spe_ctx* ctx, const char* who, int slot, const PipelineBuffers& pb, Pipe& P, PipeSlot& S,
                        cudaEvent_t ready, const uint8_t* frames_dev, int H, int W, long long pitch,
                        long long frame_stride, const int32_t* boxes_dev, int B, const spe_pnp_params* params) {
  (void)P;
  cudaStream_t st = S.compute;
  cudaError_t e = cudaStreamWaitEvent(st, ready, 0);
  if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string(who) + ": " + cudaGetErrorString(e));
  // the crop writes the predictor's padded stem input of this slot directly when it can (staged kernel, windowed stem)
  int stem_bf16 = 0;
  void* stem_buf = stem_input_buffer(ctx, slot, &stem_bf16);
  bool stem_written = false;
  {
    const std::string s = launch_crop_resize_norm(frames_dev, H, W, pitch, frame_stride, boxes_dev, B, pb.R, S.images_dev, st,
                                                  stem_buf, stem_bf16, &stem_written);
    if (!s.empty()) return set_error(ctx, SPE_ERR_CUDA, std::string(who) + ": crop: " + s);
  }
  int rc;
  const bool sig = pb.has_sigma != 0;
  rc = forward_half(ctx, stem_written ? 7 : 3, slot, S.images_dev, B, S.logits_dev, S.points_dev, sig ? S.logsig_dev : nullptr, st);
  if (rc != SPE_OK) return rc;
  spe_pnp_params pp = *params;
  if (!sig) pp.weighted = 0;
  const float* pl = pb.ov_logits ? pb.ov_logits : S.logits_dev;
  const float* pp_pts = pb.ov_points ? pb.ov_points : S.points_dev;
  rc = spe_assign_pnp(ctx, pl, pp_pts, sig ? S.logsig_dev : nullptr, pb.ov_boxes ? pb.ov_boxes : boxes_dev, B, pb.Q,
                      &pp, S.quat_dev, S.tvec_dev, S.assign_dev, S.status_dev, nullptr, nullptr, nullptr, nullptr, st);
  if (rc != SPE_OK) return rc;
  e = cudaMemcpyAsync(S.quat_h, S.quat_dev, sizeof(double) * 4 * B, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(S.tvec_h, S.tvec_dev, sizeof(double) * 3 * B, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(S.status_h, S.status_dev, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaEventRecord(S.done, st);
  if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string(who) + ": " + cudaGetErrorString(e));
  S.busy = true;
  S.B = B;
  S.host_boxes = (boxes_dev == S.boxes_dev);   // spe_submit_batch_host computed S.boxes_h for this batch
  return SPE_OK;
}

int spe_submit_batch_host(spe_ctx* ctx, int slot, const uint8_t* frames_host, int H, int W,
                          const double* det_boxes_host, int B, const spe_pnp_params* params) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_submit_batch_host: null ctx");
  if (!frames_host || !det_boxes_host || !params) return set_error(ctx, SPE_ERR_INVALID, "spe_submit_batch_host: null buffer");
  PipelineBuffers pb;
  Pipe* P = nullptr;
  PipeSlot* Sp = nullptr;
  int rc = pipe_prepare(ctx, "spe_submit_batch_host", slot, B, pb, &P, &Sp);
  if (rc != SPE_OK) return rc;
  PipeSlot& S = *Sp;
  cudaError_t e = cudaSuccess;
  const long long need = static_cast<long long>(B) * H * W;
  if (S.frames_cap < need) {
    if (S.frames_dev) cudaFree(S.frames_dev);
    S.frames_dev = nullptr;
    S.frames_cap = 0;
    const long long cap = static_cast<long long>(pb.max_batch) * H * W;
    e = cudaMalloc(reinterpret_cast<void**>(&S.frames_dev), static_cast<size_t>(cap));
    if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string("frame buffer: ") + cudaGetErrorString(e));
    S.frames_cap = cap;
  }
  spe_clip_boxes(det_boxes_host, B, S.boxes_h);
  // upload only the intersection of each crop box with its frame (see spe_run_batch_host)
  long long h2d = 0;
  for (int i = 0; i < B && e == cudaSuccess; ++i) {
    const int32_t* bx = S.boxes_h + 4 * i;
    const int x0 = bx[0] < 0 ? 0 : bx[0], y0 = bx[1] < 0 ? 0 : bx[1];
    const int x1 = bx[2] > W ? W : bx[2], y1 = bx[3] > H ? H : bx[3];
    if (x1 <= x0 || y1 <= y0) continue;
    const long long off = static_cast<long long>(i) * H * W + static_cast<long long>(y0) * W + x0;
    e = cudaMemcpy2DAsync(S.frames_dev + off, W, frames_host + off, W, static_cast<size_t>(x1 - x0),
                          static_cast<size_t>(y1 - y0), cudaMemcpyHostToDevice, S.stream);
    h2d += static_cast<long long>(x1 - x0) * (y1 - y0);
  }
  *pb.last_h2d_bytes = h2d + static_cast<long long>(sizeof(int32_t)) * 4 * B;
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(S.boxes_dev, S.boxes_h, sizeof(int32_t) * 4 * B, cudaMemcpyHostToDevice, S.stream);
  if (e == cudaSuccess) e = cudaEventRecord(S.upload_done, S.stream);
  if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string("spe_submit_batch_host: ") + cudaGetErrorString(e));
  return pipe_enqueue(ctx, "spe_submit_batch_host", slot, pb, *P, S, S.upload_done, S.frames_dev, H, W, W,
                      static_cast<long long>(H) * W, S.boxes_dev, B, params);
}

int spe_submit_batch_dev(spe_ctx* ctx, int slot, const uint8_t* frames_dev, int H, int W, long long pitch,
                         long long frame_stride, const int32_t* boxes_dev, int B, const spe_pnp_params* params,
                         void* stream) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_submit_batch_dev: null ctx");
  if (!frames_dev || !boxes_dev || !params || H <= 0 || W <= 0 || pitch < W)
    return set_error(ctx, SPE_ERR_INVALID, "spe_submit_batch_dev: bad argument");
  PipelineBuffers pb;
  Pipe* P = nullptr;
  PipeSlot* Sp = nullptr;
  int rc = pipe_prepare(ctx, "spe_submit_batch_dev", slot, B, pb, &P, &Sp);
  if (rc != SPE_OK) return rc;
  // the frames and boxes are whatever `stream` has produced up to this call
  cudaError_t e = cudaEventRecord(P->caller_ready, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string("spe_submit_batch_dev: ") + cudaGetErrorString(e));
  *pb.last_h2d_bytes = 0;
  return pipe_enqueue(ctx, "spe_submit_batch_dev", slot, pb, *P, *Sp, P->caller_ready, frames_dev, H, W, pitch,
                      frame_stride, boxes_dev, B, params);
}

int spe_collect_batch_host(spe_ctx* ctx, int slot, double* quat_host, double* tvec_host, int32_t* status_host,
                           int32_t* boxes_host) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_collect_batch_host: null ctx");
  if (slot < 0 || slot >= kPipeSlots)
    return set_error(ctx, SPE_ERR_INVALID, "spe_collect_batch_host: slot outside [0, SPE_PIPELINE_SLOTS)");
  if (!quat_host || !tvec_host || !status_host) return set_error(ctx, SPE_ERR_INVALID, "spe_collect_batch_host: null buffer");
  PipeSlot& S = pipe_of(ctx).slot[slot];
  if (!S.busy) return set_error(ctx, SPE_ERR_STATE, "spe_collect_batch_host: nothing submitted on this slot");
  cudaError_t e = cudaEventSynchronize(S.done);
  S.busy = false;
  if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string("spe_collect_batch_host: ") + cudaGetErrorString(e));
  memcpy(quat_host, S.quat_h, sizeof(double) * 4 * S.B);
  memcpy(tvec_host, S.tvec_h, sizeof(double) * 3 * S.B);
  memcpy(status_host, S.status_h, sizeof(int32_t) * S.B);
  if (boxes_host) {   // the crop boxes libspe computed: host submissions only (device submissions brought their own)
    if (S.host_boxes) memcpy(boxes_host, S.boxes_h, sizeof(int32_t) * 4 * S.B);
    else memset(boxes_host, 0, sizeof(int32_t) * 4 * S.B);
  }
  return SPE_OK;
}

int spe_debug_read_slot_outputs(spe_ctx* ctx, int slot, float* logits_host, float* points_host) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_debug_read_slot_outputs: null ctx");
  if (slot < 0 || slot >= kPipeSlots || !logits_host || !points_host)
    return set_error(ctx, SPE_ERR_INVALID, "spe_debug_read_slot_outputs: bad argument");
  PipeSlot& S = pipe_of(ctx).slot[slot];
  if (S.busy || S.B <= 0 || !S.logits_dev)
    return set_error(ctx, SPE_ERR_STATE, "spe_debug_read_slot_outputs: collect the slot first");
  const PipelineBuffers pb = pipeline_buffers(ctx);
  cudaError_t e = cudaMemcpy(logits_host, S.logits_dev, sizeof(float) * 12 * pb.Q * S.B, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(points_host, S.points_dev, sizeof(float) * 2 * pb.Q * S.B, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string("spe_debug_read_slot_outputs: ") + cudaGetErrorString(e));
  return SPE_OK;
}

int spe_profile_enable(int on) {
  profile_enable(on != 0);
  return SPE_OK;
}

int spe_profile_collect(double* ms_by_family, long long* launches_by_family) {
  if (!ms_by_family || !launches_by_family) return set_error(nullptr, SPE_ERR_INVALID, "spe_profile_collect: null");
  profile_collect(ms_by_family, launches_by_family);
  return SPE_OK;
}

int spe_debug_set_pnp_override(spe_ctx* ctx, const float* logits_dev, const float* points_dev,
                               const int32_t* boxes_dev) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_debug_set_pnp_override: null ctx");
  set_pnp_override(ctx, logits_dev, points_dev, boxes_dev);
  return SPE_OK;
}

// ---- bring-up hooks ---------------------------------------------------------------------------------------------
static int sm_count() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

int spe_debug_gemm(int dtype, const void* A, const void* Wt, long long M, int N, int K, const float* scale,
                   const float* bias, const void* residual, int res_mod, int relu, void* out, void* stream) {
  GemmDesc d;
  d.mode = 0; d.A = A; d.M = M; d.K = K; d.lda = K; d.Wt = Wt; d.N = N;
  d.scale = scale; d.bias = bias; d.residual = residual; d.res_ld = N; d.res_mod = res_mod; d.relu = relu;
  d.out = out; d.out_ld = N;
  if (dtype == 2) { d.x3 = 1; d.round_out = 0; }   // 3xTF32: Wt is the pre-split [N, 2K] matrix
  std::string s = launch_gemm(dtype == 1 ? kBF16 : kTF32, d, sm_count(), static_cast<cudaStream_t>(stream));
  if (!s.empty()) return set_error(nullptr, SPE_ERR_CUDA, "spe_debug_gemm: " + s);
  return SPE_OK;
}

int spe_debug_gemm2(int dtype, const void* A, int K, const void* A2, int K2, int a2_stride, int NB, int H, int W,
                    const void* Wt, long long M, int N, const float* bias, int relu, void* out, void* stream) {
  GemmDesc d;
  d.mode = 0; d.A = A; d.M = M; d.K = K; d.lda = K; d.Wt = Wt; d.N = N;
  d.A2 = A2; d.K2 = K2; d.lda2 = K2; d.a2_stride = a2_stride; d.a2_NB = NB; d.a2_H = H; d.a2_W = W;
  d.bias = bias; d.relu = relu; d.out = out; d.out_ld = N;
  std::string s = launch_gemm(dtype == 1 ? kBF16 : kTF32, d, sm_count(), static_cast<cudaStream_t>(stream));
  if (!s.empty()) return set_error(nullptr, SPE_ERR_CUDA, "spe_debug_gemm2: " + s);
  return SPE_OK;
}

int spe_debug_conv(int dtype, const void* x, const void* w, int NB, int H, int W, int C, int Cout, int R, int S,
                   int pad, int stride, const float* scale, const float* bias, int relu, void* out, void* stream) {
  GemmDesc d;
  d.mode = 1; d.A = x; d.NB = NB; d.H = H; d.W = W; d.C = C; d.R = R; d.S = S; d.pad = pad; d.Wt = w; d.N = Cout;
  d.conv_stride = stride;
  d.scale = scale; d.bias = bias; d.relu = relu; d.out = out; d.out_ld = Cout;
  if (dtype == 2) { d.x3 = 1; d.round_out = 0; }   // 3xTF32: w is the pre-split [Cout, 2 R S C] matrix
  std::string s = launch_gemm(dtype == 1 ? kBF16 : kTF32, d, sm_count(), static_cast<cudaStream_t>(stream));
  if (!s.empty()) return set_error(nullptr, SPE_ERR_CUDA, "spe_debug_conv: " + s);
  return SPE_OK;
}

int spe_debug_ffn(const float* X, long long M, const float* W1, const float* b1, const float* W2, const float* b2,
                  const float* gamma, const float* beta, int hidden, int out_mode, float* out, void* stream) {
  FfnDesc d;
  d.X = X; d.M = M; d.W1 = W1; d.b1 = b1; d.W2 = W2; d.b2 = b2; d.gamma = gamma; d.beta = beta;
  d.hidden = hidden; d.out_mode = out_mode; d.out = out;
  if (!ffn_fused_supported(kTF32, 256, hidden)) return set_error(nullptr, SPE_ERR_INVALID, "spe_debug_ffn: unsupported shape");
  std::string s = launch_ffn_fused(d, sm_count(), static_cast<cudaStream_t>(stream));
  if (!s.empty()) return set_error(nullptr, SPE_ERR_CUDA, "spe_debug_ffn: " + s);
  return SPE_OK;
}

int spe_debug_attention(int dtype, const void* q, const void* k, const void* v, void* out, int B, int heads, int Lq,
                        int Lk, int ldq, int ldk, int ldv, int ldo, void* stream) {
  AttnDesc a;
  a.q = q; a.k = k; a.v = v; a.out = out; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
  a.bsq = static_cast<long long>(Lq) * ldq; a.bsk = static_cast<long long>(Lk) * ldk;
  a.bsv = static_cast<long long>(Lk) * ldv; a.bso = static_cast<long long>(Lq) * ldo;
  a.B = B; a.heads = heads; a.Lq = Lq; a.Lk = Lk; a.scale = 0.17677669529663687f;
  std::string s = launch_attention(dtype == 0 ? kTF32 : kBF16, a, static_cast<cudaStream_t>(stream));
  if (!s.empty()) return set_error(nullptr, SPE_ERR_CUDA, "spe_debug_attention: " + s);
  return SPE_OK;
}

}  // extern "C"
