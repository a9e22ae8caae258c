/* libspe.so -- C ABI of the B200-native crop -> keypoint-set predictor -> PnP path.
 *
 * The reference (wwhitecyan/satellite-pose-estimation) has no FFI for this path: its boundary is four Python call
 * signatures.  Each entry point below names the reference interface it sits under; INTEGRATION.md shows the ctypes
 * stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, a negative spe_status otherwise; spe_last_error() gives the message
 *   - the caller owns every input/output buffer; "dev" pointers are CUDA device pointers on the ctx's device,
 *     "host" pointers are ordinary (preferably pinned) host memory
 *   - work is enqueued on the caller's CUDA stream (pass cudaStream_t / torch.cuda.current_stream().cuda_stream as
 *     void*; NULL = default stream); no entry point synchronises except spe_run_batch_host and spe_sync
 *   - one spe_ctx per process per GPU; a ctx is not thread-safe
 *   - sm_100a only: spe_create fails on any other device; there is no CPU fallback anywhere in this library
 */
#ifndef SPE_H_
#define SPE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; only this ABI is exported */
#endif

typedef struct spe_ctx spe_ctx;

typedef enum {
  SPE_OK = 0,
  SPE_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  SPE_ERR_CUDA = -2,      /* a CUDA call failed */
  SPE_ERR_DEVICE = -3,    /* not an sm_100 device */
  SPE_ERR_WEIGHTS = -4,   /* missing / mis-shaped tensor in spe_load_weights */
  SPE_ERR_STATE = -5      /* call order (e.g. forward before weights) */
} spe_status;

/* pose status written by spe_assign_pnp (reference behaviour: RV/gen_submission_single.py:169-175 maps solver
 * failures to the zero pose) */
enum { SPE_POSE_OK = 0, SPE_POSE_TOO_FEW = 1, SPE_POSE_DIVERGED = 2, SPE_POSE_REJECTED = 3 };

/* Mirrors the argparse fields the reference's build_model(args) reads (RV/main.py:90-187, RV/models/detr_speed.py:
 * 296-336, RV/models/backbone.py:184-198, RV/models/transformer.py:284-294). */
typedef struct {
  int input_size;       /* R: network input is [B,3,R,R]                               (--input_size)      */
  int num_queries;      /* Q                                                             (--num_queries)     */
  int enc_layers;       /*                                                               (--enc_layers)      */
  int dec_layers;       /*                                                               (--dec_layers)      */
  int hidden_dim;       /* must be 256                                                   (--hidden_dim)      */
  int nheads;           /* must be 8 (head_dim 32)                                       (--nheads)          */
  int dim_feedforward;  /*                                                               (--dim_feedforward) */
  int backbone;         /* 0: ResNet-50 stride-8 fusion neck (Backbone8s), 1: ResNet-50 layer3 stride 16,
                         * 2: the SA drop's RT-DETR predictor (PResNet-50-vd + HybridEncoder + RTDETRTransformer,
                         *    SA/configs/rtdetr_speed/rtdetr_r50vd_6x_speed_kl_*.yml): input_size = eval_spatial_size
                         *    (256), num_queries 30, enc_layers 1 (AIFI), dec_layers 3, dim_feedforward 1024,
                         *    precision 0, has_sigma 1; weights in the SA state_dict layout                        */
  int precision;        /* 0: fp32 storage + TF32 tensor cores, 1: bf16 storage + BF16 tensor cores          */
  int has_sigma;        /* 1: self-assessment variant, sigma_embed.* head -> pred_sigmas                     */
  int max_batch;        /* workspace is sized for this many images per call                                  */
} spe_config;

/* One tensor of the reference state_dict, fp32, HOST memory, C-contiguous. */
typedef struct {
  const char* name;
  const float* data;
  int ndim;
  long long shape[4];
} spe_tensor_desc;

/* ---- lifetime --------------------------------------------------------------------------------------------- */
/* replaces: build_model(args) + model.to(device)            (RV/models/__init__.py:5-6, RV/main.py:210-211) */
int spe_create(const spe_config* cfg, int device, spe_ctx** out);
void spe_destroy(spe_ctx* ctx);
/* message of the last failure on this ctx (ctx == NULL: last failure of spe_create) */
const char* spe_last_error(const spe_ctx* ctx);
/* replaces: model.load_state_dict(checkpoint['model'])      (RV/gen_submission_single.py:216-218)
 * folds FrozenBatchNorm2d (RV/models/backbone.py:44-54) and repacks to kernel layouts once */
int spe_load_weights(spe_ctx* ctx, const spe_tensor_desc* tensors, int n);
int spe_sync(spe_ctx* ctx, void* stream);

/* ---- stage 0: image file -> frame --------------------------------------------------------- */
/* replaces: Image.open(img_path).convert('RGB')   (RV/datasets/speed.py:116, :212; SA/src/data/speed/speed_dataset.py)
 * for what SPEED ships: baseline (SOF0 / SOF1) Huffman-coded 8-bit single-component JPEG files.  PIL decodes them
 * through libjpeg's JDCT_ISLOW integer inverse DCT; the frames written here are bit-identical to np.asarray(Image.open(f))
 * (and .convert('RGB') of a grayscale image replicates that plane, which the crop kernel does on the fly).
 * files_host[i] / sizes[i]: the bytes of file i in host memory.  The call parses the headers, copies the compressed
 * scans into a pinned staging buffer owned by the ctx, and enqueues ONE upload + the decode kernel (one warp per image)
 * on `stream`; frames_dev (image b at frames_dev + b*frame_stride, rows `pitch` bytes apart, every file exactly W x H)
 * is valid when the stream reaches that point.  Progressive, arithmetic-coded, 12-bit and multi-component files are
 * refused with SPE_ERR_INVALID (message names the file index); nothing is enqueued in that case. */
int spe_jpeg_decode_batch(spe_ctx* ctx, const uint8_t* const* files_host, const long long* sizes, int B,
                          uint8_t* frames_dev, int H, int W, long long pitch, long long frame_stride, void* stream);
/* header walk only (host; no device needed): width / height of a file spe_jpeg_decode_batch would accept */
int spe_jpeg_info(const uint8_t* file_host, long long size, int* width, int* height);

/* ---- stage 1: crop -------------------------------------------------------------------------------------------- */
/* replaces: SpeedSubmission.generate_clip_bbox               (RV/datasets/speed.py:92-108)
 * host-side, float64, int() truncation toward zero; det_boxes [B,4] = x1,y1,x2,y2 ; boxes [B,4] int32 */
int spe_clip_boxes(const double* det_boxes_host, int B, int32_t* boxes_host);
/* replaces: SpeedTrain.generate_clip_bbox_val + the box rounding of PIL.Image.crop   (RV/datasets/speed.py:246-260,
 * :223; the main.py --eval path).  float_boxes [B,4] = centre +- 0.6 max(w,h) clipped to the W x H frame (float64, the
 * box PostProcess later de-normalises with); crop_boxes [B,4] int32 = each coordinate rounded half-to-even like
 * Python's round(): the (generally non-square) pixel rectangle that spe_crop_resize_norm squashes to R x R */
int spe_clip_boxes_val(const double* det_boxes_host, int B, int W, int H, double* float_boxes_host,
                       int32_t* crop_boxes_host);
/* replaces: canvas copy + cv2.resize(INTER_CUBIC) + to_tensor + Normalize   (RV/datasets/speed.py:113-160), and
 * PIL crop + A.Resize + to_tensor + Normalize of the eval path (:219-236): boxes may be any pixel rectangle
 * frames: uint8 grayscale, image b at frames + b*frame_stride, rows `pitch` bytes apart; out: fp32 [B,3,R,R] */
int spe_crop_resize_norm(spe_ctx* ctx, const uint8_t* frames_dev, int H, int W, long long pitch,
                         long long frame_stride, const int32_t* boxes_dev, int B, int R, float* out_nchw_dev,
                         void* stream);

/* ---- stage 2: keypoint-set predictor ------------------------------------------------------------------------- */
/* replaces: DETR.forward                                     (RV/models/detr_speed.py:59-92)
 * images [B,3,R,R] fp32 NCHW.  logits [B,Q,12], points [B,Q,2] (sigmoid, normalised), log_sigma [B,Q,2] or NULL,
 * aux_logits [(L-1),B,Q,12] / aux_points [(L-1),B,Q,2] or NULL (the 'aux_outputs' list). */
int spe_forward(spe_ctx* ctx, const float* images_dev, int B, float* logits_dev, float* points_dev,
                float* log_sigma_dev, float* aux_logits_dev, float* aux_points_dev, void* stream);

/* replaces: RTDETR.forward in eval mode                      (SA/src/zoo/rtdetr/rtdetr.py:36-52, rtdetr_decoder.py:686-751)
 * for a ctx created with backbone = 2 (spe_forward serves the same ctx when only the last layer is wanted).
 * logits [B,Q,12] 'pred_logits', points [B,Q,2] 'pred_pts', log_sigma [B,Q,2] 'pred_sigmas' (or NULL);
 * the 'aux_outputs' list, each or all NULL: aux_logits [L,B,Q,12] / aux_points [L,B,Q,2] = decoder layers 0..L-2 followed
 * by the encoder's top-k proposals (enc_topk_logits / enc_topk_bboxes), aux_log_sigma [(L-1),B,Q,2];
 * topk_idx [B,Q] int32 or NULL receives the selected anchor of every query (torch.topk order, :646-648);
 * topk_override [B,Q] int32 or NULL replaces the selection (parity tests: the selection is a discontinuous function of
 * scores that differ by rounding between any two implementations).  Runs eagerly (no graph replay). */
int spe_forward_sa(spe_ctx* ctx, const float* images_dev, int B, float* logits_dev, float* points_dev, float* log_sigma_dev,
                   float* aux_logits_dev, float* aux_points_dev, float* aux_log_sigma_dev, int32_t* topk_idx_dev,
                   const int32_t* topk_override_dev, void* stream);

/* Optional accuracy step with no counterpart in the reference (whose fp32 weights need none).  The tensor cores read
 * TF32 / BF16 weights; what the rounded weights lose, (W - round(W)) . x, is to first order the same for every image:
 * (W - round(W)) . mean(x).  spe_calibrate runs one forward over `images` ([B,3,R,R] fp32, e.g. the first batch of
 * crops), measures each layer's per-channel input mean and folds that term into the layer's bias (FrozenBN bias,
 * linear bias or positional addend), once, at no inference cost.  Measured on the benchmarked frames: TF32 keypoint
 * error 0.18 -> 0.05 px rms at the largest crop.  Results of later forwards depend on the calibration
 * batch only through this correction, which is itself below the TF32 rounding noise it removes.  Must not be called
 * while pipeline slots are in flight; synchronises `stream`.  Re-calibration replaces the previous correction. */
int spe_calibrate(spe_ctx* ctx, const float* images_dev, int B, void* stream);
int spe_is_calibrated(const spe_ctx* ctx);

/* ---- stage 3: set post-processing + pose -------------------------------------------------------------------- */
/* replaces: PostProcess.forward (RV/models/detr_speed.py:264-293), SimplePoseSolver.__call__
 * (RV/utils/speed_eval.py:164-242), SimplePoseSolverSigma (SA/utils/speed_eval.py:322-420).
 * Optional outputs may be NULL.  quat is (w,x,y,z) with w >= 0; failed images get the zero pose + status != 0. */
typedef struct {
  float reproj_thresh;   /* RANSAC reprojectionError in pixels (args.repro: 20 / 25)                      */
  int weighted;          /* 1: sigma-weighted Huber LM in normalised image coordinates (needs log_sigma)  */
  int reject;            /* 1: apply the self-assessment reject filter (status SPE_POSE_REJECTED)          */
  float reject_rms_px;   /* reject when inlier RMS reprojection error exceeds this (default 5)            */
  float reject_sigma_px; /* reject when mean predicted sigma (pixels) exceeds this (default 12)           */
  const float* float_boxes_dev; /* optional [B,4] fp32 (x1, y1, width, height): de-normalise the keypoints with
                                   these instead of the int32 boxes -- the main.py --eval path hands PostProcess the
                                   UNROUNDED crop box (RV/datasets/speed.py:246-260, spe_clip_boxes_val)        */
  const float* reproj_thresh_dev; /* optional [B] fp32: per-image RANSAC threshold instead of reproj_thresh -- the SA solver
                                     derives it from the detection area, int(area / input_size * 10) clamped to [1.5, 20]
                                     (SA/utils/speed_eval_ceres.py:53-58)                                          */
  int inputs_post_processed; /* 1: `logits` are PostProcess OUTPUTS -- class probabilities, used as scores without another
                                softmax, so the assignment sees exactly the numbers the reference's solver sees -- and
                                `points` original-image pixels (pass boxes (0,0,1,1)): the reference's per-image
                                solver(points, logits) signature (RV/utils/speed_eval.py:164-200)                     */
  float sigma_px_scale;      /* with inputs_post_processed: crop side in pixels for the reject filter's sigma criterion
                                (sigmas are in normalised crop units); 0 = criterion skipped                        */
} spe_pnp_params;

int spe_assign_pnp(spe_ctx* ctx, const float* logits_dev, const float* points_dev, const float* log_sigma_dev,
                   const int32_t* boxes_dev, int B, int Q, const spe_pnp_params* params,
                   double* quat_dev /*[B,4]*/, double* tvec_dev /*[B,3]*/, int32_t* assign_dev /*[B,11]*/,
                   int32_t* status_dev /*[B]*/, float* probs_dev /*[B,Q,12] or NULL*/,
                   float* points_px_dev /*[B,Q,2] or NULL*/, float* sigmas_dev /*[B,Q,2] or NULL*/,
                   int32_t* inlier_mask_dev /*[B] or NULL*/, void* stream);

/* replaces: Multi_Mean_PoseSolver.__call__ (RV/utils/speed_eval.py:42-140) over a batch, fed with the per-model network
 * outputs that gen_prediction collects (RV/gen_submission_multi.py:122-141): logits [num_models,B,Q,12] and points
 * [num_models,B,Q,2] (normalised; the PostProcess de-normalisation with `boxes` happens inside).  Every foreground
 * query of every model is pooled per keypoint label (mean -> drop predictions farther than 3 std of the distances
 * -> mean), then the same RANSAC-P3P consensus + LM refinement as spe_assign_pnp.  count [B,11] receives the number of
 * predictions each label's mean was taken over (0 = label absent), pooled_px [B,11,2] the pooled keypoints. */
int spe_ensemble_pnp(spe_ctx* ctx, const float* logits_dev, const float* points_dev, const int32_t* boxes_dev,
                     int num_models, int B, int Q, const spe_pnp_params* params, double* quat_dev /*[B,4]*/,
                     double* tvec_dev /*[B,3]*/, int32_t* count_dev /*[B,11]*/, int32_t* status_dev /*[B]*/,
                     float* pooled_px_dev /*[B,11,2] or NULL*/, int32_t* inlier_mask_dev /*[B] or NULL*/, void* stream);

/* replaces: speed_score (RV/utils/speed_eval.py:245-262) for a batch that stays on the device: score_t = |t^ - t| / |t|,
 * score_q = 2 acos(min(|q^ . q|, 1)); all arrays float64, quaternions (w,x,y,z) */
int spe_speed_score(spe_ctx* ctx, const double* quat_pr_dev, const double* tvec_pr_dev, const double* quat_gt_dev,
                    const double* tvec_gt_dev, int B, double* score_t_dev, double* score_q_dev, void* stream);

/* ---- SA (RT-DETR) variant: first decoder pieces ----------------------------------------- */
/* replaces: deformable_attention_core_func (SA/src/zoo/rtdetr/utils.py:15-64) -- per level F.grid_sample(bilinear,
 * zeros padding, align_corners=False) and the weighted sum over levels x points -- and, with fused = 1, also what
 * MSDeformableAttention.forward computes between its linear layers and the core (SA/src/zoo/rtdetr/rtdetr_decoder.py:
 * 117-163): softmax of the attention logits over levels x points, sampling location = reference point + offset / (W_l, H_l).
 *   value [B, Lv, heads, 32] fp32, levels concatenated along Lv; shapes_hw_host [L, 2] = (H_l, W_l) in HOST memory
 *   fused = 0: loc [B, Lq, heads, L, P, 2] sampling locations in [0, 1], attn [B, Lq, heads, L, P] softmaxed weights
 *   fused = 1: loc = raw sampling offsets (same shape), attn = raw logits, ref [B, Lq, ref_levels, 2] (ref_levels 1 or L)
 *   out [B, Lq, heads * 32] fp32.  head_dim is 32 (embed 256 / 8 heads), L * P <= 64. */
int spe_ms_deform_attn(spe_ctx* ctx, const float* value_dev, const int32_t* shapes_hw_host, int L, const float* loc_dev,
                       const float* attn_dev, const float* ref_dev, int ref_levels, int B, int Lq, int heads, int P,
                       int fused, float* out_dev, void* stream);
/* replaces: torch.topk(enc_outputs_class.max(-1).values, num_queries, dim=1) (SA/src/zoo/rtdetr/rtdetr_decoder.py:
 * 646-648): cls [B, Lv, C] fp32 -> idx [B, k] int32 in descending score order (ties: lower index), vals [B, k] or NULL */
int spe_topk_queries(spe_ctx* ctx, const float* cls_dev, int B, int Lv, int C, int k, int32_t* idx_dev, float* vals_dev,
                     void* stream);
/* replaces: tensor.gather(dim=1, index=topk_ind.unsqueeze(-1).repeat(1, 1, D)) (rtdetr_decoder.py:651-680):
 * out[b, r, :] = src[b, idx[b, r], :] for src [B, Lv, D] fp32 */
int spe_gather_rows(spe_ctx* ctx, const float* src_dev, const int32_t* idx_dev, int B, int Lv, int k, int D,
                    float* out_dev, void* stream);

/* ---- whole path, host buffers in, host buffers out ---------------------------------------------------------- */
/* replaces the hot loop of gen_submission (RV/gen_submission_single.py:136-181): frames + detector boxes in
 * host memory -> poses in host memory.  Uploads, runs crop -> forward -> assign/PnP on `stream`, downloads and
 * synchronises.  frames_host: uint8 [B,H,W]; det_boxes_host: double [B,4]. */
int spe_run_batch_host(spe_ctx* ctx, const uint8_t* frames_host, int H, int W, const double* det_boxes_host, int B,
                       const spe_pnp_params* params, double* quat_host, double* tvec_host, int32_t* status_host,
                       int32_t* boxes_host /*[B,4] or NULL*/, void* stream);

/* Pipelined form of the same loop.  A slot (0 .. SPE_PIPELINE_SLOTS-1) is a complete, independent instance of the
 * path: own streams, frame / result buffers and activation set (allocated on first use, ~3.5 GB at B=64).  submit
 * enqueues upload + compute + download of one batch on the slot and returns at once; collect waits for that slot and
 * hands out the poses.  Keeping several slots in flight lets the GPU run whole batches next to each other: the upload
 * of one hides behind the kernels of the others, and the small latency-bound launches of a batch's decoder / pose
 * stage and the drained last wave of every persistent GEMM are filled with another batch's work (B=64 on B200:
 * 6.7 ms/batch one at a time, 5.9 ms with three in flight when the pipeline was built; four is the measured optimum
 * with host frames).  Results are bit-identical to the one-stream calls.
 * frames_host must be pinned and stay valid until the slot is collected.  Do not mix with spe_forward /
 * spe_run_batch_host while slots are in flight (slot 0 shares their activation set). */
#define SPE_PIPELINE_SLOTS 8
int spe_submit_batch_host(spe_ctx* ctx, int slot, const uint8_t* frames_host, int H, int W,
                          const double* det_boxes_host, int B, const spe_pnp_params* params);
int spe_collect_batch_host(spe_ctx* ctx, int slot, double* quat_host, double* tvec_host, int32_t* status_host,
                           int32_t* boxes_host /*[B,4] or NULL*/);

/* Same pipeline for frames that are already resident in device memory (a decoder / camera stack that delivers into
 * HBM): frames_dev uint8 [B] frames of H x W with row pitch `pitch` and frame stride `frame_stride` bytes, boxes_dev
 * int32 [B,4] clip boxes (spe_clip_boxes), both complete on `stream` at the time of the call and untouched until
 * the slot is collected.  Collect with spe_collect_batch_host (boxes_host is not filled). */
int spe_submit_batch_dev(spe_ctx* ctx, int slot, const uint8_t* frames_dev, int H, int W, long long pitch,
                         long long frame_stride, const int32_t* boxes_dev, int B, const spe_pnp_params* params,
                         void* stream);
/* bytes spe_run_batch_host uploaded on its last call (only the crop-box/frame intersections travel) */
long long spe_last_h2d_bytes(spe_ctx* ctx);

/* ---- test / bring-up hooks (not part of the drop-in surface) ------------------------------------------------- */
/* out = act(scale * A.W^T + bias + residual) with A [M,K], W [N,K], storage dtype 0=fp32/TF32, 1=bf16,
 * 2=fp32 with error-compensated 3xTF32 (W is then [N,2K] = [rna(W) | rna(W - rna(W))], output not rounded) */
int spe_debug_gemm(int dtype, const void* A_dev, const void* W_dev, long long M, int N, int K, const float* scale_dev,
                   const float* bias_dev, const void* residual_dev, int res_mod, int relu, void* out_dev,
                   void* stream);
/* out[M, N] = act([A | A2] . Wt^T + bias): A [M, K]; A2 either [M, K2] (a2_stride 1) or the (2h, 2w) sampling of an NHWC
 * activation [NB, H, W, K2] with M = NB * ceil(H/2) * ceil(W/2) (a2_stride 2); Wt [N, K + K2] */
int spe_debug_gemm2(int dtype, const void* A_dev, int K, const void* A2_dev, int K2, int a2_stride, int NB, int H, int W,
                    const void* Wt_dev, long long M, int N, const float* bias_dev, int relu, void* out_dev, void* stream);
/* convolution (stride 1 or 2) as implicit GEMM: x [NB,H,W,C] NHWC, w [Cout, R*S*C] (tap-major, channel-minor),
 * out [NB,Ho,Wo,Cout] with Ho = (H + 2 pad - R) / stride + 1; dtype as in spe_debug_gemm (2: w is [Cout, 2*R*S*C]) */
int spe_debug_conv(int dtype, const void* x_dev, const void* w_dev, int NB, int H, int W, int C, int Cout, int R,
                   int S, int pad, int stride, const float* scale_dev, const float* bias_dev, int relu, void* out_dev,
                   void* stream);
/* fused encoder feed-forward block, fp32/TF32: out = LayerNorm(X + relu(X W1^T + b1) W2^T + b2); X [M,256], W1
 * [hidden,256], W2 [256,hidden]; out_mode 0 rounded / 1 exact [M,256], 2 = [M,768] 3xTF32 operand form, 3 = bf16 [M,256] */
int spe_debug_ffn(const float* X, long long M, const float* W1, const float* b1, const float* W2, const float* b2,
                  const float* gamma, const float* beta, int hidden, int out_mode, float* out, void* stream);
int spe_debug_attention(int dtype, const void* q_dev, const void* k_dev, const void* v_dev, void* out_dev, int B,
                        int heads, int Lq, int Lk, int ldq, int ldk, int ldv, int ldo, void* stream);
/* keep copies of intermediate activations during spe_forward (names: stem, layer1, layer2, layer3, neck,
 * input_proj, enc<i>, hs) and read them back as raw storage-dtype bytes in kernel layout (NHWC / [rows, C]) */
int spe_debug_enable_taps(spe_ctx* ctx, int enable);
long long spe_debug_read_tap(spe_ctx* ctx, const char* name, void* host_out, long long max_bytes);
const char* spe_global_last_error(void);
/* launch accounting for bench.py: per kernel family {gemm, attention, elementwise, heads, crop, pnp} the number of
 * launches since the last collect and, while enabled, their CUDA-event-timed device milliseconds */
int spe_profile_enable(int on);
int spe_profile_collect(double* ms_by_family /*[6]*/, long long* launches_by_family /*[6]*/);
/* bench hook: the batch pipelines (run / submit) feed these resident [B,Q,12] / [B,Q,2] tensors -- and, if given, the
 * int32 [B,4] crop boxes they were generated for -- to the pose stage instead of the network outputs (random-init
 * weights collapse to one label, which would make the solve exit early); NULL resets */
int spe_debug_set_pnp_override(spe_ctx* ctx, const float* logits_dev, const float* points_dev,
                               const int32_t* boxes_dev);
/* test hook: how many (batch size, buffer set) keys of the forward schedule replay a captured CUDA graph, and for how
 * many the capture failed (those keep running kernel by kernel) */
int spe_debug_graph_stats(const spe_ctx* ctx, int* captured, int* failed);
/* test hook: network outputs (logits [B,Q,12], points [B,Q,2]) of the batch last collected from pipeline `slot` */
int spe_debug_read_slot_outputs(spe_ctx* ctx, int slot, float* logits_host, float* points_host);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* SPE_H_ */
