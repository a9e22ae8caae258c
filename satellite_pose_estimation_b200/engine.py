"""Thin Python owner of one ``spe_ctx``: PyTorch supplies device memory and streams, libspe.so does the work."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SpeConfig, SpePnpParams, SpeTensorDesc, check

BACKBONE_S8, BACKBONE_S16, BACKBONE_SA_RTDETR = 0, 1, 2
SA_BACKBONES = ("rtdetr_r50vd", "sa_rtdetr")
PRECISION = {"tf32": 0, "fp32": 0, "bf16": 1}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    """One context per process per GPU (include/spe.h).  All tensors passed in must live on ``device``."""

    def __init__(self, *, input_size=224, num_queries=40, enc_layers=4, dec_layers=4, hidden_dim=256, nheads=8,
                 dim_feedforward=2048, backbone="resnet50s8", precision="tf32", has_sigma=False, max_batch=64,
                 device=0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("satellite_pose_estimation_b200 needs a CUDA (sm_100a) device; there is no CPU path")
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.cfg = SpeConfig(
            input_size=input_size, num_queries=num_queries, enc_layers=enc_layers, dec_layers=dec_layers,
            hidden_dim=hidden_dim, nheads=nheads, dim_feedforward=dim_feedforward,
            backbone=(BACKBONE_SA_RTDETR if backbone in SA_BACKBONES else
                      BACKBONE_S16 if backbone in ("resnet18", "resnet34", "resnet50") else BACKBONE_S8),
            precision=PRECISION[precision], has_sigma=int(bool(has_sigma)), max_batch=max_batch)
        self.precision = precision
        self.has_sigma = bool(has_sigma)
        self.is_sa = backbone in SA_BACKBONES
        self.max_batch = max_batch
        self.Q, self.L, self.R = num_queries, dec_layers, input_size
        self._ctx = C.c_void_p()
        torch.cuda.init()
        with torch.cuda.device(self.device):
            check(self.lib.spe_create(C.byref(self.cfg), self.device.index, C.byref(self._ctx)))
        self.weights_loaded = False
        self._fwd_bufs = {}
        self._inflight = {}
        self._stable_inputs = set()

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.spe_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights ---------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        """Reference ``state_dict`` -> device weights (BN folded, kernel layouts)."""
        keep, descs = [], []
        for name, t in state_dict.items():
            if name.endswith("num_batches_tracked"):
                continue  # dropped by FrozenBatchNorm2d._load_from_state_dict, RV/models/backbone.py:34-42
            a = t.detach().to("cpu", torch.float32).contiguous()
            if a.dim() > 4:
                continue
            keep.append(a)
            d = SpeTensorDesc()
            d.name = name.encode()
            d.data = a.data_ptr()
            d.ndim = a.dim()
            for i, s in enumerate(a.shape):
                d.shape[i] = s
            descs.append(d)
        arr = (SpeTensorDesc * len(descs))(*descs)
        with torch.cuda.device(self.device):
            check(self.lib.spe_load_weights(self._ctx, arr, len(descs)), self._ctx)
        self.weights_loaded = True

    # ---- stage 1 ---------------------------------------------------------------------------------------------
    def clip_boxes(self, det_boxes):
        det = np.ascontiguousarray(np.asarray(det_boxes, dtype=np.float64).reshape(-1, 4))
        out = np.empty((det.shape[0], 4), dtype=np.int32)
        check(self.lib.spe_clip_boxes(det.ctypes.data_as(C.c_void_p), det.shape[0], out.ctypes.data_as(C.c_void_p)))
        return out

    def clip_boxes_val(self, det_boxes, W=1920, H=1200):
        """``main.py --eval`` crop boxes (RV/datasets/speed.py:246-260 + PIL's rounding): returns (float64 [B,4] unrounded
        boxes -- what PostProcess de-normalises with --, int32 [B,4] pixel rectangles for ``crop_resize_norm``)."""
        det = np.ascontiguousarray(np.asarray(det_boxes, dtype=np.float64).reshape(-1, 4))
        fbox = np.empty((det.shape[0], 4), dtype=np.float64)
        ibox = np.empty((det.shape[0], 4), dtype=np.int32)
        check(self.lib.spe_clip_boxes_val(det.ctypes.data_as(C.c_void_p), det.shape[0], int(W), int(H),
                                          fbox.ctypes.data_as(C.c_void_p), ibox.ctypes.data_as(C.c_void_p)))
        return fbox, ibox

    def speed_score(self, quat_pr, tvec_pr, quat_gt, tvec_gt):
        """Batched ``speed_score`` (RV/utils/speed_eval.py:245-262) on device: float64 cuda [B,4] / [B,3] -> (s_t, s_q)."""
        q, t, qg, tg = (x.to(torch.float64).contiguous() for x in (quat_pr, tvec_pr, quat_gt, tvec_gt))
        B = q.shape[0]
        s_t = torch.empty((B,), dtype=torch.float64, device=q.device)
        s_q = torch.empty((B,), dtype=torch.float64, device=q.device)
        check(self.lib.spe_speed_score(self._ctx, _ptr(q), _ptr(t), _ptr(qg), _ptr(tg), B, _ptr(s_t), _ptr(s_q),
                                       _stream(q.device)), self._ctx)
        return s_t, s_q

    # ---- stage 0 ---------------------------------------------------------------------------------------------
    def decode_jpeg(self, files, out=None):
        """files: list of ``bytes`` (baseline grayscale JPEG files of one size, as SPEED ships them) -> uint8 cuda
        [B,H,W], bit-identical to ``np.asarray(PIL.Image.open(f))`` (RV/datasets/speed.py:116 decodes with PIL and
        replicates the plane to RGB; the crop stage replicates on the fly).  Huffman decoding and the inverse DCT run
        on the GPU (one warp per image); only the compressed bytes cross PCIe.  Refuses progressive / colour files."""
        B = len(files)
        if B == 0:
            raise ValueError("decode_jpeg: empty batch")
        w, h = C.c_int(0), C.c_int(0)
        files = [bytes(f) if not isinstance(f, bytes) else f for f in files]
        bufs = [C.c_char_p(f) for f in files]            # pointers into the bytes objects themselves: no copy
        check(self.lib.spe_jpeg_info(C.cast(bufs[0], C.c_void_p), len(files[0]), C.byref(w), C.byref(h)), None)
        H, W = h.value, w.value
        dev = self.device
        if out is None:
            out = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        assert out.dtype == torch.uint8 and out.is_cuda and tuple(out.shape) == (B, H, W) and out.stride(2) == 1
        ptrs = (C.c_void_p * B)(*[C.cast(b, C.c_void_p).value for b in bufs])
        sizes = (C.c_longlong * B)(*[len(f) for f in files])
        check(self.lib.spe_jpeg_decode_batch(self._ctx, ptrs, sizes, B, _ptr(out), H, W, out.stride(1), out.stride(0),
                                             _stream(dev)), self._ctx)
        return out

    def crop_resize_norm(self, frames, boxes, out=None, R=None):
        """frames: uint8 cuda [B,H,W]; boxes: int32 cuda [B,4] -> float32 cuda [B,3,R,R]."""
        R = R or self.R
        assert frames.dtype == torch.uint8 and frames.is_cuda and frames.dim() == 3 and frames.stride(2) == 1
        assert boxes.dtype == torch.int32 and boxes.is_cuda and boxes.is_contiguous()
        B, H, W = frames.shape
        if out is None:
            out = torch.empty((B, 3, R, R), dtype=torch.float32, device=frames.device)
        check(self.lib.spe_crop_resize_norm(self._ctx, _ptr(frames), H, W, frames.stride(1), frames.stride(0),
                                            _ptr(boxes), B, R, _ptr(out), _stream(frames.device)), self._ctx)
        return out

    # ---- stage 2 ---------------------------------------------------------------------------------------------
    def forward(self, images, want_aux=False):
        """images float32 cuda [B,3,R,R] -> dict of float32 cuda tensors (reference output layout).

        Inputs are staged into, and outputs returned from, per-batch-size persistent buffers (stable addresses let
        libspe replay the captured CUDA graph): the returned tensors are overwritten by the next ``forward`` call
        with the same batch size -- clone them if they must outlive it."""
        assert images.is_cuda and images.dtype == torch.float32
        B = images.shape[0]
        if images.shape[1:] != (3, self.R, self.R):
            raise ValueError(f"expected images [B,3,{self.R},{self.R}], got {tuple(images.shape)}")
        dev = images.device
        key = (B, bool(want_aux))
        bufs = self._fwd_bufs.get(key)
        if bufs is None:
            bufs = {"in": torch.empty((B, 3, self.R, self.R), dtype=torch.float32, device=dev),
                    "logits": torch.empty((B, self.Q, 12), dtype=torch.float32, device=dev),
                    "points": torch.empty((B, self.Q, 2), dtype=torch.float32, device=dev),
                    "logsig": torch.empty((B, self.Q, 2), dtype=torch.float32, device=dev) if self.has_sigma else None,
                    "aux_l": None, "aux_p": None}
            if want_aux and self.L > 1:
                bufs["aux_l"] = torch.empty((self.L - 1, B, self.Q, 12), dtype=torch.float32, device=dev)
                bufs["aux_p"] = torch.empty((self.L - 1, B, self.Q, 2), dtype=torch.float32, device=dev)
            self._fwd_bufs[key] = bufs
        if images.data_ptr() != bufs["in"].data_ptr():
            if images.is_contiguous() and images.data_ptr() in self._stable_inputs:
                src = images                      # caller-owned persistent buffer (e.g. the crop output)
            else:
                bufs["in"].copy_(images)
                src = bufs["in"]
        else:
            src = images
        logits, points, logsig, aux_l, aux_p = (bufs[k] for k in ("logits", "points", "logsig", "aux_l", "aux_p"))
        check(self.lib.spe_forward(self._ctx, _ptr(src), B, _ptr(logits), _ptr(points), _ptr(logsig), _ptr(aux_l),
                                   _ptr(aux_p), _stream(dev)), self._ctx)
        out = {"pred_logits": logits, "pred_points": points}
        if logsig is not None:
            out["pred_sigmas"] = logsig
        if aux_l is not None:
            out["aux_outputs"] = [{"pred_logits": aux_l[i], "pred_points": aux_p[i]} for i in range(self.L - 1)]
        return out

    def forward_sa(self, images, want_aux=True, topk_override=None):
        """The SA drop's RT-DETR predictor (``Engine(backbone="rtdetr_r50vd", ...)``): images float32 cuda [B,3,R,R] ->
        the dict ``RTDETR.forward`` returns in eval mode (SA/src/zoo/rtdetr/rtdetr_decoder.py:732-751): ``pred_logits``
        [B,Q,12], ``pred_pts`` [B,Q,2], ``pred_sigmas`` [B,Q,2] (log sigma) and ``aux_outputs`` (decoder layers 0..L-2 with
        sigmas, then the encoder's top-k proposals without), plus ``topk_ind`` [B,Q] int32, the anchors the queries were
        taken from.  ``topk_override`` (int32 cuda [B,Q]) replaces the top-k selection."""
        if not self.is_sa:
            raise ValueError("forward_sa needs an Engine built with backbone='rtdetr_r50vd'")
        assert images.is_cuda and images.dtype == torch.float32
        B = images.shape[0]
        if images.shape[1:] != (3, self.R, self.R):
            raise ValueError(f"expected images [B,3,{self.R},{self.R}], got {tuple(images.shape)}")
        dev, Q, L = images.device, self.Q, self.L
        x = images.contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        logits, points, logsig = torch.empty((B, Q, 12), **f32), torch.empty((B, Q, 2), **f32), torch.empty((B, Q, 2), **f32)
        aux_l = torch.empty((L, B, Q, 12), **f32) if want_aux else None
        aux_p = torch.empty((L, B, Q, 2), **f32) if want_aux else None
        aux_s = torch.empty((max(L - 1, 1), B, Q, 2), **f32) if want_aux else None
        topk = torch.empty((B, Q), dtype=torch.int32, device=dev)
        if topk_override is not None:
            topk_override = topk_override.to(device=dev, dtype=torch.int32).contiguous()
            assert tuple(topk_override.shape) == (B, Q)
        check(self.lib.spe_forward_sa(self._ctx, _ptr(x), B, _ptr(logits), _ptr(points), _ptr(logsig), _ptr(aux_l),
                                      _ptr(aux_p), _ptr(aux_s), _ptr(topk), _ptr(topk_override), _stream(dev)), self._ctx)
        out = {"pred_logits": logits, "pred_pts": points, "pred_sigmas": logsig, "topk_ind": topk}
        if want_aux:
            out["aux_outputs"] = [{"pred_logits": aux_l[i], "pred_pts": aux_p[i], "pred_sigmas": aux_s[i]} for i in range(L - 1)]
            out["aux_outputs"].append({"pred_logits": aux_l[L - 1], "pred_pts": aux_p[L - 1]})
        return out

    def calibrate(self, images, max_images=16):
        """Fold the mean effect of the TF32 / BF16 weight rounding into the layer biases, measured on ``images``
        (float32 cuda [B,3,R,R], typically the first batch of crops; at most ``max_images`` are used).  See
        ``spe_calibrate`` in include/spe.h.  Not while pipeline slots are in flight."""
        assert images.is_cuda and images.dtype == torch.float32 and images.shape[1:] == (3, self.R, self.R)
        x = images[:min(images.shape[0], max_images, self.max_batch)].contiguous()
        check(self.lib.spe_calibrate(self._ctx, _ptr(x), x.shape[0], _stream(x.device)), self._ctx)

    @property
    def calibrated(self):
        return bool(self.lib.spe_is_calibrated(self._ctx))

    def register_stable_input(self, tensor):
        """Declare a caller-owned, persistent, contiguous input buffer: ``forward`` reads it in place (no staging
        copy), keeping the CUDA-graph key stable."""
        self._stable_inputs.add(tensor.data_ptr())

    @staticmethod
    def sa_detection_area(det_boxes):
        """The "area" the SA dataset hands its solver (SA/src/data/speed/speed_dataset.py:370-373), reproduced with its
        parenthesisation: ``sqrt((x2 - x1) * y2 - y1)`` of the detector box -- a length in pixels, not an area."""
        b = np.asarray(det_boxes, dtype=np.float64).reshape(-1, 4)
        return np.sqrt((b[:, 2] - b[:, 0]) * b[:, 3] - b[:, 1])

    @staticmethod
    def area_repro_threshold(areas, input_size):
        """``EPnPCeresSolver.get_repro_th`` (SA/utils/speed_eval_ceres.py:53-58): RANSAC threshold from the detection
        area, ``int(area / input_size * 10)`` clamped to [1.5, 20]; float32 numpy [B] for ``assign_pnp(reproj=...)``."""
        a = np.asarray(areas, dtype=np.float64).reshape(-1)
        return np.clip(np.trunc(a / float(input_size) * 10), 1.5, 20).astype(np.float32)

    # ---- stage 3 ---------------------------------------------------------------------------------------------
    def assign_pnp(self, logits, points, boxes, log_sigma=None, reproj=20.0, weighted=False, reject=False,
                   reject_rms_px=5.0, reject_sigma_px=12.0, want_post=False, post_processed=False, sigma_px_scale=0.0):
        """Batched PostProcess + assignment + PnP.  All inputs cuda; returns dict of cuda tensors.  Integer ``boxes``
        are the submission path's crop boxes; floating-point ``boxes`` [B,4] (x1,y1,x2,y2) are the eval path's
        unrounded boxes (``clip_boxes_val``), applied like PostProcess applies a float64 ``clip_bbox``.
        ``post_processed``: ``logits`` are PostProcess outputs (class probabilities, used as scores as they are) and
        ``points`` original-image pixels (``boxes`` must then be (0,0,1,1)); ``sigma_px_scale`` = crop side for the
        reject filter's sigma criterion in that mode (0 skips the criterion)."""
        logits = logits.contiguous().float(); points = points.contiguous().float()
        fbox = None
        if boxes.is_floating_point():
            b64 = boxes.to(torch.float64)
            fbox = torch.stack([b64[:, 0], b64[:, 1], b64[:, 2] - b64[:, 0], b64[:, 3] - b64[:, 1]], 1).float().contiguous()
        boxes = boxes.to(torch.int32).contiguous()
        B, Q = logits.shape[0], logits.shape[1]
        dev = logits.device
        quat = torch.empty((B, 4), dtype=torch.float64, device=dev)
        tvec = torch.empty((B, 3), dtype=torch.float64, device=dev)
        assign = torch.empty((B, 11), dtype=torch.int32, device=dev)
        status = torch.empty((B,), dtype=torch.int32, device=dev)
        inl = torch.empty((B,), dtype=torch.int32, device=dev)
        probs = pts_px = sig = None
        if want_post:
            probs = torch.empty((B, Q, 12), dtype=torch.float32, device=dev)
            pts_px = torch.empty((B, Q, 2), dtype=torch.float32, device=dev)
            if log_sigma is not None:
                sig = torch.empty((B, Q, 2), dtype=torch.float32, device=dev)
        if log_sigma is not None:
            log_sigma = log_sigma.contiguous().float()
        thr = None
        if torch.is_tensor(reproj):          # per-image thresholds (e.g. area_repro_threshold), cuda float [B]
            thr = reproj.to(device=dev, dtype=torch.float32).contiguous()
            assert thr.shape == (B,)
        p = SpePnpParams(reproj_thresh=0.0 if thr is not None else float(reproj), weighted=int(weighted),
                         reject=int(reject), reject_rms_px=float(reject_rms_px), reject_sigma_px=float(reject_sigma_px),
                         float_boxes_dev=fbox.data_ptr() if fbox is not None else None,
                         reproj_thresh_dev=thr.data_ptr() if thr is not None else None,
                         inputs_post_processed=int(bool(post_processed)), sigma_px_scale=float(sigma_px_scale))
        check(self.lib.spe_assign_pnp(self._ctx, _ptr(logits), _ptr(points), _ptr(log_sigma), _ptr(boxes), B, Q,
                                      C.byref(p), _ptr(quat), _ptr(tvec), _ptr(assign), _ptr(status), _ptr(probs),
                                      _ptr(pts_px), _ptr(sig), _ptr(inl), _stream(dev)), self._ctx)
        out = {"quat": quat, "tvec": tvec, "assign": assign, "status": status, "inlier_mask": inl}
        if want_post:
            out["probs"], out["points_px"] = probs, pts_px
            if sig is not None:
                out["sigmas"] = sig
        return out

    def ensemble_pnp(self, logits, points, boxes, reproj=25.0, reject=False, want_pooled=False, post_processed=False):
        """Batched ``Multi_Mean_PoseSolver`` (RV/utils/speed_eval.py:42-140): ``logits`` [Nm,B,Q,12] and ``points``
        [Nm,B,Q,2] are the raw outputs of the Nm ensemble members for the same B crops (cuda), ``boxes`` int [B,4]
        the crop boxes.  Returns dict of cuda tensors: quat, tvec, status, count [B,11] (predictions pooled per
        keypoint after the 3-sigma filter), inlier_mask, and with ``want_pooled`` the pooled keypoints [B,11,2] px."""
        logits = logits.contiguous().float(); points = points.contiguous().float()
        boxes = boxes.to(torch.int32).contiguous()
        Nm, B, Q = logits.shape[0], logits.shape[1], logits.shape[2]
        assert points.shape[:3] == (Nm, B, Q) and boxes.shape == (B, 4)
        dev = logits.device
        quat = torch.empty((B, 4), dtype=torch.float64, device=dev)
        tvec = torch.empty((B, 3), dtype=torch.float64, device=dev)
        count = torch.empty((B, 11), dtype=torch.int32, device=dev)
        status = torch.empty((B,), dtype=torch.int32, device=dev)
        inl = torch.empty((B,), dtype=torch.int32, device=dev)
        pooled = torch.empty((B, 11, 2), dtype=torch.float32, device=dev) if want_pooled else None
        p = SpePnpParams(reproj_thresh=float(reproj), weighted=0, reject=int(reject), reject_rms_px=5.0,
                         reject_sigma_px=12.0, inputs_post_processed=int(bool(post_processed)))
        check(self.lib.spe_ensemble_pnp(self._ctx, _ptr(logits), _ptr(points), _ptr(boxes), Nm, B, Q, C.byref(p),
                                        _ptr(quat), _ptr(tvec), _ptr(count), _ptr(status), _ptr(pooled), _ptr(inl),
                                        _stream(dev)), self._ctx)
        out = {"quat": quat, "tvec": tvec, "count": count, "status": status, "inlier_mask": inl}
        if want_pooled:
            out["pooled_px"] = pooled
        return out

    # ---- SA (RT-DETR) decoder pieces ----------------------------------------------------------------------------
    def ms_deform_attn(self, value, spatial_shapes, loc, attn, ref=None):
        """Multi-scale deformable attention core (``deformable_attention_core_func``, SA/src/zoo/rtdetr/utils.py:15-64).
        ``value`` cuda float32 [B, Lv, heads, 32]; ``spatial_shapes`` [(H_l, W_l)]; without ``ref``: ``loc``
        [B, Lq, heads, L, P, 2] sampling locations in [0, 1] and ``attn`` [B, Lq, heads, L, P] softmaxed weights;
        with ``ref`` [B, Lq, 1 or L, 2]: ``loc`` / ``attn`` are the raw outputs of ``MSDeformableAttention``'s
        ``sampling_offsets`` / ``attention_weights`` linear layers (rtdetr_decoder.py:117-163).  -> [B, Lq, heads*32]."""
        value, loc, attn = value.contiguous().float(), loc.contiguous().float(), attn.contiguous().float()
        B, Lv, heads, hd = value.shape
        assert hd == 32 and value.is_cuda
        Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
        shp = np.ascontiguousarray(np.asarray(spatial_shapes, dtype=np.int32).reshape(L, 2))
        assert int((shp[:, 0] * shp[:, 1]).sum()) == Lv
        if ref is not None:
            ref = ref.contiguous().float()
        out = torch.empty((B, Lq, heads * 32), dtype=torch.float32, device=value.device)
        check(self.lib.spe_ms_deform_attn(self._ctx, _ptr(value), shp.ctypes.data_as(C.c_void_p), L, _ptr(loc), _ptr(attn),
                                          _ptr(ref), ref.shape[2] if ref is not None else 0, B, Lq, heads, P,
                                          int(ref is not None), _ptr(out), _stream(value.device)), self._ctx)
        return out

    def topk_queries(self, enc_class, k):
        """``torch.topk(enc_outputs_class.max(-1).values, k, dim=1)`` (rtdetr_decoder.py:646-648): cuda float32
        [B, Lv, C] -> (values [B, k], indices int32 [B, k]) in descending score order."""
        x = enc_class.contiguous().float()
        B, Lv, Cc = x.shape
        idx = torch.empty((B, k), dtype=torch.int32, device=x.device)
        val = torch.empty((B, k), dtype=torch.float32, device=x.device)
        check(self.lib.spe_topk_queries(self._ctx, _ptr(x), B, Lv, Cc, k, _ptr(idx), _ptr(val), _stream(x.device)), self._ctx)
        return val, idx

    def gather_rows(self, src, idx):
        """``src.gather(1, idx[..., None].repeat(1, 1, D))`` (rtdetr_decoder.py:651-680): [B, Lv, D], int32 [B, k] -> [B, k, D]."""
        src = src.contiguous().float(); idx = idx.contiguous().to(torch.int32)
        B, Lv, D = src.shape
        k = idx.shape[1]
        out = torch.empty((B, k, D), dtype=torch.float32, device=src.device)
        check(self.lib.spe_gather_rows(self._ctx, _ptr(src), _ptr(idx), B, Lv, k, D, _ptr(out), _stream(src.device)), self._ctx)
        return out

    # ---- whole path, host in / host out -----------------------------------------------------------------------
    def run_batch_host(self, frames_host, det_boxes, reproj=20.0, weighted=False, reject=False):
        """frames_host: uint8 (pinned) torch/numpy [B,H,W]; det_boxes: float64 [B,4] -> numpy quat/tvec/status/boxes."""
        if isinstance(frames_host, torch.Tensor):
            assert not frames_host.is_cuda and frames_host.dtype == torch.uint8 and frames_host.is_contiguous()
            fptr, (B, H, W) = C.c_void_p(frames_host.data_ptr()), frames_host.shape
        else:
            frames_host = np.ascontiguousarray(frames_host, dtype=np.uint8)
            fptr, (B, H, W) = frames_host.ctypes.data_as(C.c_void_p), frames_host.shape
        det = np.ascontiguousarray(np.asarray(det_boxes, dtype=np.float64).reshape(-1, 4))
        quat = np.empty((B, 4), dtype=np.float64); tvec = np.empty((B, 3), dtype=np.float64)
        status = np.empty((B,), dtype=np.int32); boxes = np.empty((B, 4), dtype=np.int32)
        p = SpePnpParams(reproj_thresh=float(reproj), weighted=int(weighted), reject=int(reject),
                         reject_rms_px=5.0, reject_sigma_px=12.0)
        check(self.lib.spe_run_batch_host(self._ctx, fptr, H, W, det.ctypes.data_as(C.c_void_p), B, C.byref(p),
                                          quat.ctypes.data_as(C.c_void_p), tvec.ctypes.data_as(C.c_void_p),
                                          status.ctypes.data_as(C.c_void_p), boxes.ctypes.data_as(C.c_void_p),
                                          _stream(self.device)), self._ctx)
        return {"quat": quat, "tvec": tvec, "status": status, "boxes": boxes,
                "h2d_bytes": int(self.lib.spe_last_h2d_bytes(self._ctx))}

    def submit_batch_host(self, slot, frames_host, det_boxes, reproj=20.0, weighted=False, reject=False):
        """Asynchronous half of ``run_batch_host`` (slot 0 .. SPE_PIPELINE_SLOTS - 1 = 7): returns as soon as the work is enqueued.
        ``frames_host`` must be a pinned uint8 torch tensor [B,H,W] that stays alive until ``collect_batch_host``."""
        assert isinstance(frames_host, torch.Tensor) and frames_host.is_pinned() and frames_host.dtype == torch.uint8
        B, H, W = frames_host.shape
        det = np.ascontiguousarray(np.asarray(det_boxes, dtype=np.float64).reshape(-1, 4))
        p = SpePnpParams(reproj_thresh=float(reproj), weighted=int(weighted), reject=int(reject),
                         reject_rms_px=5.0, reject_sigma_px=12.0)
        check(self.lib.spe_submit_batch_host(self._ctx, slot, C.c_void_p(frames_host.data_ptr()), H, W,
                                             det.ctypes.data_as(C.c_void_p), B, C.byref(p)), self._ctx)
        self._inflight[slot] = (frames_host, B, int(self.lib.spe_last_h2d_bytes(self._ctx)))

    def submit_batch_dev(self, slot, frames_dev, boxes_dev, reproj=20.0, weighted=False, reject=False):
        """``submit_batch_host`` for frames already resident in HBM: uint8 CUDA tensor [B,H,W] (contiguous) and the
        int32 clip boxes [B,4] (``clip_boxes``), both produced on the current torch stream.  Collect with
        ``collect_batch_host``; the tensors must stay untouched until then."""
        assert frames_dev.is_cuda and frames_dev.dtype == torch.uint8 and frames_dev.is_contiguous()
        assert boxes_dev.is_cuda and boxes_dev.dtype == torch.int32 and boxes_dev.is_contiguous()
        B, H, W = frames_dev.shape
        p = SpePnpParams(reproj_thresh=float(reproj), weighted=int(weighted), reject=int(reject),
                         reject_rms_px=5.0, reject_sigma_px=12.0)
        check(self.lib.spe_submit_batch_dev(self._ctx, slot, C.c_void_p(frames_dev.data_ptr()), H, W, W, H * W,
                                            C.c_void_p(boxes_dev.data_ptr()), B, C.byref(p), _stream(frames_dev.device)), self._ctx)
        self._inflight[slot] = ((frames_dev, boxes_dev), B, 0)

    def collect_batch_host(self, slot):
        frames_host, B, h2d = self._inflight.pop(slot)
        quat = np.empty((B, 4), dtype=np.float64); tvec = np.empty((B, 3), dtype=np.float64)
        status = np.empty((B,), dtype=np.int32); boxes = np.empty((B, 4), dtype=np.int32)
        check(self.lib.spe_collect_batch_host(self._ctx, slot, quat.ctypes.data_as(C.c_void_p),
                                              tvec.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p),
                                              boxes.ctypes.data_as(C.c_void_p)), self._ctx)
        out = {"quat": quat, "tvec": tvec, "status": status, "h2d_bytes": h2d}
        if not isinstance(frames_host, tuple):      # host submissions: the crop boxes libspe computed for this batch
            out["boxes"] = boxes
        return out

    def graph_stats(self):
        """Test hook: (keys of the forward schedule that replay a captured CUDA graph, keys whose capture failed)."""
        a, b = C.c_int(0), C.c_int(0)
        check(self.lib.spe_debug_graph_stats(self._ctx, C.byref(a), C.byref(b)), self._ctx)
        return a.value, b.value

    def read_slot_outputs(self, slot, B):
        """Test hook: (logits [B,Q,12], points [B,Q,2]) the network produced for the batch last collected from `slot`."""
        Q = self.cfg.num_queries
        logits = np.empty((B, Q, 12), dtype=np.float32); points = np.empty((B, Q, 2), dtype=np.float32)
        check(self.lib.spe_debug_read_slot_outputs(self._ctx, slot, logits.ctypes.data_as(C.c_void_p),
                                                   points.ctypes.data_as(C.c_void_p)), self._ctx)
        return logits, points

    # ---- measurement hooks (bench.py) ------------------------------------------------------------------------------
    FAMILIES = ("gemm", "attention", "elementwise", "heads", "crop", "pnp")

    def profile_enable(self, on=True):
        check(self.lib.spe_profile_enable(int(on)))

    def profile_collect(self):
        """-> ({family: event-timed ms}, {family: launches}) since the previous collect."""
        ms = (C.c_double * 6)()
        n = (C.c_longlong * 6)()
        check(self.lib.spe_profile_collect(ms, n))
        return dict(zip(self.FAMILIES, list(ms))), dict(zip(self.FAMILIES, list(n)))

    def set_pnp_override(self, logits=None, points=None, boxes=None):
        """Bench hook: the batch pipelines feed these resident tensors to the pose stage (None resets)."""
        self._override = (logits, points, boxes)   # keep alive
        check(self.lib.spe_debug_set_pnp_override(self._ctx, _ptr(logits), _ptr(points), _ptr(boxes)), self._ctx)

    # ---- bring-up -----------------------------------------------------------------------------------------------
    def enable_taps(self, on=True):
        check(self.lib.spe_debug_enable_taps(self._ctx, int(on)), self._ctx)

    def read_tap(self, name, shape):
        """Intermediate activation as float32 torch CPU tensor of ``shape`` (kernel layout: NHWC / [rows, C])."""
        nbytes = self.lib.spe_debug_read_tap(self._ctx, name.encode(), None, 0)
        if nbytes < 0:
            check(int(nbytes), self._ctx)
        buf = torch.empty(nbytes, dtype=torch.uint8)
        got = self.lib.spe_debug_read_tap(self._ctx, name.encode(), C.c_void_p(buf.data_ptr()), nbytes)
        assert got == nbytes
        dt = torch.float32 if self.cfg.precision == 0 else torch.bfloat16
        return buf.view(dt).float().reshape(shape)
