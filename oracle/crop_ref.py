"""CPU restatement of the reference's test-time crop (TEST INFRASTRUCTURE; see oracle/__init__.py).

Follows ``SpeedSubmission`` in RV/datasets/speed.py:
  * ``generate_clip_bbox``                      :92-108
  * zero canvas + slice copy                    :127-144
  * ``A.Resize(R, R, cv2.INTER_CUBIC)``         :49-53, :146-149   (albumentations 0.5.1 ``Resize`` is exactly
                                                 ``cv2.resize(img, (R, R), interpolation=INTER_CUBIC)``)
  * ``F.to_tensor`` + ``Normalize``             :152-153, :25-41

The third-party arithmetic (bicubic resampling) lives in OpenCV (reference pins opencv-python==4.4.0.44,
RV/requirements.txt:6; this image has 4.13.0): the oracle calls the same ``cv2.resize`` the reference calls.
``bicubic_f64`` is a plain-numpy model of that resampler used to explain the CUDA kernel's arithmetic; it is
checked against cv2 in tests/test_oracle_crop.py.
"""
import numpy as np
import cv2
import torch

from .constants import MEAN, STD


def generate_clip_bbox(bbox):
    """RV/datasets/speed.py:92-108.  float64 arithmetic, ``int()`` truncates toward zero. Returns int64[4]."""
    x1, y1, x2, y2 = [float(v) for v in bbox]
    bbox_width, bbox_height = x2 - x1, y2 - y1
    scale = max(bbox_width, bbox_height) * 1.2
    x_center, y_center = (x1 + x2) / 2, (y1 + y2) / 2
    half_scale = scale / 2
    x1, y1 = int(x_center - half_scale), int(y_center - half_scale)
    scale = int(scale)
    return np.asarray([x1, y1, x1 + scale, y1 + scale], dtype=np.int64)


def make_canvas(gray, bbox_clip):
    """RV/datasets/speed.py:116-144.  ``gray`` is uint8 [H, W]; the reference converts to RGB (three equal
    channels), so the canvas is S x S x 3 with the frame-intersect-box region copied and zeros elsewhere."""
    height, width = gray.shape
    img = np.repeat(gray[:, :, None], 3, axis=2)
    clip_size = int(bbox_clip[2] - bbox_clip[0])
    canvas = np.zeros((clip_size, clip_size, 3), dtype=img.dtype)
    x1 = max(0, bbox_clip[0]); crop_x1 = int(x1 - bbox_clip[0])
    y1 = max(0, bbox_clip[1]); crop_y1 = int(y1 - bbox_clip[1])
    x2 = min(width, bbox_clip[2]); y2 = min(height, bbox_clip[3])
    x1, x2, y1, y2 = [int(v) for v in (x1, x2, y1, y2)]
    if y2 > y1 and x2 > x1:
        canvas[crop_y1:crop_y1 + y2 - y1, crop_x1:crop_x1 + x2 - x1] = img[y1:y2, x1:x2]
    return canvas


def crop_resize_u8(gray, bbox_clip, R):
    """canvas -> cv2 bicubic -> uint8 [R, R, 3]  (RV/datasets/speed.py:146-149)."""
    canvas = make_canvas(gray, bbox_clip)
    return cv2.resize(canvas, (R, R), interpolation=cv2.INTER_CUBIC)


def normalize_u8(img_u8):
    """``F.to_tensor`` then ``F.normalize``  (RV/datasets/speed.py:152-153, :25-41) -> float32 [3, R, R]."""
    t = torch.from_numpy(np.ascontiguousarray(img_u8)).permute(2, 0, 1).contiguous()
    t = t.to(torch.float32).div(255)
    mean = torch.as_tensor(MEAN, dtype=torch.float32)[:, None, None]
    std = torch.as_tensor(STD, dtype=torch.float32)[:, None, None]
    return t.sub(mean).div(std)


def crop_resize_normalize(gray, det_bbox, R):
    """Full ``SpeedSubmission.__getitem__``: returns (float32 tensor [3, R, R], int64 clip box [4])."""
    bbox_clip = generate_clip_bbox(det_bbox)
    return normalize_u8(crop_resize_u8(gray, bbox_clip, R)), bbox_clip


# ----------------------------------------------------------------------------------------------------------
# numpy model of cv2.resize(INTER_CUBIC) on uint8 (what the CUDA kernel computes, in the same fp64 arithmetic)
# ----------------------------------------------------------------------------------------------------------
def _cubic_weights(t):
    A = -0.75
    w0 = ((A * (t + 1) - 5 * A) * (t + 1) + 8 * A) * (t + 1) - 4 * A
    w1 = ((A + 2) * t - (A + 3)) * t * t + 1
    w2 = ((A + 2) * (1 - t) - (A + 3)) * (1 - t) * (1 - t) + 1
    w3 = 1.0 - w0 - w1 - w2
    return np.stack([w0, w1, w2, w3], -1)


def bicubic_f64(canvas_gray, R):
    """Keys cubic a=-0.75, half-pixel centres, replicate border, separable, round-half-even + saturate."""
    S = canvas_gray.shape[0]
    d = np.arange(R, dtype=np.float64)
    f = (d + 0.5) * (S / R) - 0.5
    s = np.floor(f)
    w = _cubic_weights(f - s)
    idx = np.clip(s.astype(np.int64)[:, None] + np.arange(-1, 3)[None], 0, S - 1)
    im = canvas_gray.astype(np.float64)
    hz = (im[:, idx] * w[None]).sum(-1)
    out = (hz[idx] * w[:, :, None]).sum(1)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------------------------
# main.py --eval path (SpeedTrain with train=False), RV/datasets/speed.py:209-260
# ----------------------------------------------------------------------------------------------------------
def generate_clip_bbox_val(bbox, image_size):
    """RV/datasets/speed.py:246-260: centre +- 0.6 max(w, h), clipped to the frame; float64[4], generally not square."""
    x1, y1, x2, y2 = [float(v) for v in bbox]
    bbox_width, bbox_height = x2 - x1, y2 - y1
    scale = max(bbox_width, bbox_height) * 1.2
    x_center, y_center = (x1 + x2) / 2, (y1 + y2) / 2
    half_scale = scale / 2
    bbox_clip = np.asarray([x_center - half_scale, y_center - half_scale, x_center + half_scale, y_center + half_scale])
    bbox_clip[0::2] = bbox_clip[0::2].clip(min=0, max=image_size[0])
    bbox_clip[1::2] = bbox_clip[1::2].clip(min=0, max=image_size[1])
    return bbox_clip


def eval_crop_resize_u8(gray, bbox_clip, R):
    """:219-230: ``np.array(img.crop(bbox_clip))`` (PIL rounds every coordinate with ``round``) then
    ``A.Resize(R, R, cv2.INTER_CUBIC)``.  The random ``img_trunc`` (:230, applied even in eval) is left out: it
    depends on numpy's global RNG state."""
    from PIL import Image
    img = Image.fromarray(gray).convert("RGB")
    crop = np.array(img.crop(bbox_clip))
    return cv2.resize(crop, (R, R), interpolation=cv2.INTER_CUBIC)


def eval_crop_resize_normalize(gray, det_bbox, R):
    bbox_clip = generate_clip_bbox_val(det_bbox, (gray.shape[1], gray.shape[0]))
    return normalize_u8(eval_crop_resize_u8(gray, bbox_clip, R)), bbox_clip
