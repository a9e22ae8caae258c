"""Deterministic synthetic inputs for the hot path (TEST INFRASTRUCTURE; spec: SURVEY.md section 8d).

Everything is seeded ``numpy.random.default_rng`` (PCG64: identical streams on every machine), so the build
container and the GPU box regenerate bit-identical weights / frames / correspondences instead of shipping
100 MB fixtures.  ``weights_checksum`` lets the golden files assert that.
"""
import hashlib
import os
from argparse import Namespace

import numpy as np
import torch

from .constants import CAMERA_K, TANGO_POINTS, IMG_H, IMG_W
from .model_ref import ModelCfg, RESNET50_BLOCKS

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def reference_args(cfg: ModelCfg, device="cpu"):
    """argparse namespace the reference's ``build_model(args)`` reads (RV/main.py:90-187 defaults)."""
    return Namespace(
        backbone=cfg.backbone, dilation=False, position_embedding=getattr(cfg, "position_embedding", "sine"), bn="frozen_bn",
        enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers, dim_feedforward=cfg.dim_feedforward,
        hidden_dim=cfg.hidden_dim, dropout=0.1, nheads=cfg.nheads, num_queries=cfg.num_queries, pre_norm=False,
        aux_loss=cfg.aux_loss, lr_backbone=1e-5, device=device, set_cost_class=1, set_cost_pts=5,
        pts_loss_coef=5.0, eos_coef=0.1, repro=20)


# -------------------------------------------------------------------------------------------------------------
# weights: reference state_dict key layout, He/Xavier-scaled random values, and
# *perturbed* FrozenBN buffers so the BN fold is exercised (identity BN would hide a wrong fold).
# -------------------------------------------------------------------------------------------------------------
def _conv(rng, cout, cin, k, gain=1.0):
    fan_in = cin * k * k
    return (rng.standard_normal((cout, cin, k, k)) * (gain * np.sqrt(2.0 / fan_in))).astype(np.float32)


def _bn(rng, sd, prefix, c, gamma_scale=1.0):
    sd[prefix + ".weight"] = (gamma_scale * (1.0 + 0.1 * rng.standard_normal(c))).astype(np.float32)
    sd[prefix + ".bias"] = (0.05 * rng.standard_normal(c)).astype(np.float32)
    sd[prefix + ".running_mean"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
    sd[prefix + ".running_var"] = (1.0 + 0.2 * rng.random(c)).astype(np.float32)


def _xavier(rng, out_f, in_f):
    a = np.sqrt(6.0 / (in_f + out_f))
    return rng.uniform(-a, a, (out_f, in_f)).astype(np.float32)


def _linear(rng, sd, prefix, out_f, in_f):
    sd[prefix + ".weight"] = _xavier(rng, out_f, in_f)
    sd[prefix + ".bias"] = (0.02 * rng.standard_normal(out_f)).astype(np.float32)


def _mha(rng, sd, prefix, e):
    sd[prefix + ".in_proj_weight"] = _xavier(rng, 3 * e, e)
    sd[prefix + ".in_proj_bias"] = (0.02 * rng.standard_normal(3 * e)).astype(np.float32)
    _linear(rng, sd, prefix + ".out_proj", e, e)


def _ln(rng, sd, prefix, e):
    sd[prefix + ".weight"] = (1.0 + 0.05 * rng.standard_normal(e)).astype(np.float32)
    sd[prefix + ".bias"] = (0.02 * rng.standard_normal(e)).astype(np.float32)


def make_state_dict(cfg: ModelCfg, seed=0, spread_labels=False):
    """Random weights in the reference ``state_dict`` layout.  ``spread_labels`` swaps in the calibrated heads of
    ``tests/golden/chain_heads.npz`` (oracle/make_chain_fixture.py): query embeddings x16, ``cls_embed`` and
    ``point_embed`` fitted so that queries 0..10 emit the 11 keypoint labels at a PnP-consistent layout and the other
    queries background -- random init collapses every query to one label, which would leave the
    crop -> predictor -> PnP chain untestable on the network's own output."""
    rng = np.random.default_rng(seed)
    sd = {}
    b = "backbone.0.body"
    sd[b + ".conv1.weight"] = _conv(rng, 64, 3, 7)
    _bn(rng, sd, b + ".bn1", 64)
    inplanes = 64
    for li, (name, planes) in enumerate((("layer1", 64), ("layer2", 128), ("layer3", 256))):
        for bi in range(RESNET50_BLOCKS[name]):
            p = f"{b}.{name}.{bi}"
            sd[p + ".conv1.weight"] = _conv(rng, planes, inplanes, 1)
            _bn(rng, sd, p + ".bn1", planes)
            sd[p + ".conv2.weight"] = _conv(rng, planes, planes, 3)
            _bn(rng, sd, p + ".bn2", planes)
            sd[p + ".conv3.weight"] = _conv(rng, planes * 4, planes, 1)
            _bn(rng, sd, p + ".bn3", planes * 4, gamma_scale=0.5)
            if bi == 0:
                sd[p + ".downsample.0.weight"] = _conv(rng, planes * 4, inplanes, 1)
                _bn(rng, sd, p + ".downsample.1", planes * 4, gamma_scale=0.5)
            inplanes = planes * 4
    if cfg.stride8:
        sd["backbone.0.s8_latern.weight"] = _conv(rng, 256, 512, 1, gain=0.7)
        sd["backbone.0.s16_latern.weight"] = _conv(rng, 256, 1024, 3, gain=0.7)
        sd["backbone.0.output_conv.weight"] = _conv(rng, 512, 512, 3, gain=0.7)
        sd["backbone.0.output_conv.bias"] = (0.02 * rng.standard_normal(512)).astype(np.float32)
        nch = 512
    else:
        nch = 1024
    e = cfg.hidden_dim
    if getattr(cfg, "position_embedding", "sine") in ("learned", "v3"):
        # nn.init.uniform_ tables of PositionEmbeddingLearned (RV/models/position_encoding.py:59-67); drawn from their own
        # generator so that the other tensors of a seed do not depend on the embedding type
        rpe = np.random.default_rng(seed + 7919)
        sd["backbone.1.row_embed.weight"] = rpe.random((50, e // 2)).astype(np.float32)
        sd["backbone.1.col_embed.weight"] = rpe.random((50, e // 2)).astype(np.float32)
    sd["input_proj.weight"] = _conv(rng, e, nch, 1, gain=0.7)
    sd["input_proj.bias"] = (0.02 * rng.standard_normal(e)).astype(np.float32)
    sd["query_embed.weight"] = rng.standard_normal((cfg.num_queries, e)).astype(np.float32)
    for i in range(cfg.enc_layers):
        p = f"transformer.encoder.layers.{i}"
        _mha(rng, sd, p + ".self_attn", e)
        _linear(rng, sd, p + ".linear1", cfg.dim_feedforward, e)
        _linear(rng, sd, p + ".linear2", e, cfg.dim_feedforward)
        _ln(rng, sd, p + ".norm1", e)
        _ln(rng, sd, p + ".norm2", e)
    for i in range(cfg.dec_layers):
        p = f"transformer.decoder.layers.{i}"
        _mha(rng, sd, p + ".self_attn", e)
        _mha(rng, sd, p + ".multihead_attn", e)
        _linear(rng, sd, p + ".linear1", cfg.dim_feedforward, e)
        _linear(rng, sd, p + ".linear2", e, cfg.dim_feedforward)
        for n in (1, 2, 3):
            _ln(rng, sd, f"{p}.norm{n}", e)
    _ln(rng, sd, "transformer.decoder.norm", e)
    _linear(rng, sd, "cls_embed", 12, e)
    _linear(rng, sd, "point_embed.layers.0", e, e)
    _linear(rng, sd, "point_embed.layers.1", e, e)
    _linear(rng, sd, "point_embed.layers.2", 2, e)
    if cfg.sigma_head:
        _linear(rng, sd, "sigma_embed.layers.0", e, e)
        _linear(rng, sd, "sigma_embed.layers.1", e, e)
        _linear(rng, sd, "sigma_embed.layers.2", 1, e)
    sd = {k: torch.from_numpy(v) for k, v in sd.items()}
    if spread_labels:
        apply_chain_heads(sd, cfg, seed)
    return sd


CHAIN_QUERY_SCALE = 16.0


def apply_chain_heads(sd, cfg, seed):
    """Overwrite the heads of the C* seed-0 weights with the calibrated ones (see ``make_state_dict``)."""
    assert seed == 0 and cfg.stride8 and (cfg.num_queries, cfg.enc_layers, cfg.dec_layers) == (40, 4, 4), \
        "the calibrated heads exist for the canonical config C* with seed 0 only"
    fx = np.load(os.path.join(GOLDEN_DIR, "chain_heads.npz"))
    sd["query_embed.weight"] = sd["query_embed.weight"] * float(fx["query_embed_scale"])
    for k in fx.files:
        if k.startswith(("cls_embed.", "point_embed.")):
            sd[k] = torch.from_numpy(fx[k].astype(np.float32))
    if cfg.sigma_head:
        # log-sigma around log(2 px / 430 px): random init would predict sigma ~ 1 crop side and reject every pose
        sd["sigma_embed.layers.2.weight"] = sd["sigma_embed.layers.2.weight"] * 0.1
        sd["sigma_embed.layers.2.bias"] = torch.full((1,), float(np.log(2.0 / 430.0)), dtype=torch.float32)
    return sd


def canonical_layout():
    """Normalised (crop-relative) positions of the 11 Tango keypoints for one fixed attitude, target centred and
    filling 1/1.2 of the crop like the detector boxes do: the layout the calibrated point head is fitted to."""
    q = np.array([0.8, 0.3, -0.4, 0.33])
    q /= np.linalg.norm(q)
    pc = TANGO_POINTS @ quat_to_rot(q).T + np.array([0.0, 0.0, 8.0])
    uv = pc[:, :2] / pc[:, 2:3]
    lo, hi = uv.min(0), uv.max(0)
    return (uv - (lo + hi) / 2) / ((hi - lo).max() * 1.2) + 0.5


def weights_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


# -------------------------------------------------------------------------------------------------------------
# frames and detector boxes
# -------------------------------------------------------------------------------------------------------------
def load_detector_boxes():
    """Real detector boxes shipped with the reference (RV/annos/wz_synt_test.json, 2998 x [x1,y1,x2,y2]),
    committed as a small fixture by oracle/make_golden.py."""
    return np.load(os.path.join(GOLDEN_DIR, "wz_synt_test_boxes.npy"))


def _box_blur_int(a, k):
    """Exact integer box filter (k x k, edge-replicated) via cumulative sums -> int64."""
    pad = k // 2
    a = np.pad(a.astype(np.int64), ((pad, k - 1 - pad), (pad, k - 1 - pad)), mode="edge")
    c = np.pad(a.cumsum(0).cumsum(1), ((1, 0), (1, 0)))
    s = c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]
    return s // (k * k)


def make_frames(n, boxes, seed=0, H=IMG_H, W=IMG_W):
    """uint8 [n, H, W]: smooth low-frequency background + a bright textured blob inside each detector box.
    Integer-only numpy arithmetic, so every machine regenerates bit-identical frames."""
    rng = np.random.default_rng(seed)
    frames = np.empty((n, H, W), dtype=np.uint8)
    for i in range(n):
        low = rng.integers(0, 90, (H // 32 + 1, W // 32 + 1))
        bg = np.kron(low, np.ones((32, 32), dtype=np.int64))[:H, :W]
        bg = _box_blur_int(bg, 33)
        x1, y1, x2, y2 = [int(round(float(v))) for v in boxes[i % len(boxes)][:4]]
        x1, y1 = min(max(x1, 0), W - 2), min(max(y1, 0), H - 2)
        x2, y2 = min(max(x2, x1 + 2), W), min(max(y2, y1 + 2), H)
        bh, bw = y2 - y1, x2 - x1
        tex = rng.integers(60, 256, (bh // 6 + 1, bw // 6 + 1))
        blob = _box_blur_int(np.kron(tex, np.ones((6, 6), dtype=np.int64))[:bh, :bw], 5)
        bg[y1:y2, x1:x2] = (35 * bg[y1:y2, x1:x2] + 65 * blob) // 100
        noise = rng.integers(-6, 7, (H, W))
        frames[i] = np.clip(bg + noise, 0, 255).astype(np.uint8)
    return frames


_BENCH_BASE = {}


def bench_set(s, batch=64):
    """Frame set ``s`` of bench.py (and of the B = 64 / 256 parity goldens): uint8 frames [batch, H, W] built from 8
    seeded base frames shifted horizontally, with detector boxes ``s*batch .. (s+1)*batch`` of the real distribution."""
    det_all = load_detector_boxes()
    if "base" not in _BENCH_BASE:
        _BENCH_BASE["base"] = make_frames(8, det_all, seed=100)
    base = _BENCH_BASE["base"]
    det = det_all[s * batch:(s + 1) * batch]
    frames = np.concatenate([np.roll(base, 37 * (s * 8 + k), axis=2) for k in range(batch // 8)])
    return frames, det


# -------------------------------------------------------------------------------------------------------------
# synthetic keypoint-set predictions for the PnP stage
# -------------------------------------------------------------------------------------------------------------
def _random_quat(rng):
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    return q if q[0] >= 0 else -q


def quat_to_rot(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def make_predictions(n, Q=40, seed=1, noise_px=1.0, outlier_frac=0.10, few_frac=0.05, with_sigma=False):
    """Model-output-shaped PnP inputs with known poses (last table row).

    Returns dict: logits [n,Q,12] f32 (raw class logits), points [n,Q,2] f32 normalised to the crop box,
    boxes [n,4] int64 crop boxes, (logsig [n,Q,2] f32), q_gt [n,4], t_gt [n,3], n_outliers [n], n_visible [n].
    """
    rng = np.random.default_rng(seed)
    logits = np.full((n, Q, 12), -4.0, dtype=np.float32)
    points = rng.uniform(0.05, 0.95, (n, Q, 2)).astype(np.float32)
    boxes = np.zeros((n, 4), dtype=np.int64)
    logsig = np.zeros((n, Q, 2), dtype=np.float32)
    q_gt = np.zeros((n, 4)); t_gt = np.zeros((n, 3))
    n_out = np.zeros(n, dtype=np.int64); n_vis = np.zeros(n, dtype=np.int64)
    for i in range(n):
        q = _random_quat(rng)
        t = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.3, 0.3), rng.uniform(3.0, 20.0)])
        R = quat_to_rot(q)
        pc = TANGO_POINTS @ R.T + t
        uv = pc[:, :2] / pc[:, 2:3] * np.array([CAMERA_K[0, 0], CAMERA_K[1, 1]]) + CAMERA_K[:2, 2]
        uv = uv + rng.normal(0, noise_px, uv.shape)
        q_gt[i], t_gt[i] = q, t
        # square crop box around the projected target, like generate_clip_bbox output
        lo, hi = uv.min(0), uv.max(0)
        side = int(max(hi - lo) * 1.2) + 8
        cx, cy = (lo + hi) / 2
        x1, y1 = int(cx - side / 2), int(cy - side / 2)
        boxes[i] = (x1, y1, x1 + side, y1 + side)
        labels = np.arange(11)
        u = rng.random()
        if u < few_frac:
            labels = rng.permutation(11)[: rng.integers(0, 4)]           # < 4 visible: failure path
        elif u < few_frac + 0.10:
            labels = np.sort(rng.permutation(11)[: rng.integers(4, 11)])  # partially visible
        n_vis[i] = len(labels)
        uv_i = uv.copy()
        if len(labels) >= 6 and rng.random() < outlier_frac:
            k = int(rng.integers(1, 3))
            bad = rng.choice(labels, k, replace=False)
            ang = rng.uniform(0, 2 * np.pi, k)
            mag = rng.uniform(50, 200, k)
            uv_i[bad] += np.stack([np.cos(ang), np.sin(ang)], 1) * mag[:, None]
            n_out[i] = k
        slots = rng.permutation(Q)
        used = 0
        for l_ in labels:
            s = slots[used]; used += 1
            top = rng.uniform(0.5, 0.99)
            # logits whose softmax puts `top` on the label: others share the rest equally
            logits[i, s, :] = np.log((1 - top) / 11)
            logits[i, s, l_] = np.log(top)
            points[i, s] = (uv_i[l_] - boxes[i, :2]) / side
        n_decoy = int(rng.integers(0, 6)) if len(labels) else 0
        for _ in range(n_decoy):
            s = slots[used]; used += 1
            l_ = int(rng.choice(labels))
            top = rng.uniform(0.2, 0.45)                                   # below every true detection's score
            logits[i, s, :] = np.log((1 - top) / 11)
            logits[i, s, l_] = np.log(top)
            points[i, s] = rng.uniform(0.05, 0.95, 2)
        for s in slots[used:]:
            top = rng.uniform(0.6, 0.99)
            logits[i, s, :] = np.log((1 - top) / 11)
            logits[i, s, 11] = np.log(top)                                 # background
        logits[i] += rng.normal(0, 0.01, (Q, 12)).astype(np.float32)
        if with_sigma:
            sig_px = np.exp(rng.uniform(np.log(0.5), np.log(8.0), (Q, 1)))
            logsig[i] = np.log(sig_px / side).repeat(2, 1)
    out = dict(logits=logits, points=points, boxes=boxes, q_gt=q_gt, t_gt=t_gt, n_outliers=n_out, n_visible=n_vis)
    if with_sigma:
        out["logsig"] = logsig
    return out


def make_multi_predictions(n, num_models=5, Q=40, seed=7, noise_px=1.0, outlier_frac=0.08, miss_frac=0.05,
                           few_frac=0.05):
    """Ensemble-shaped PnP inputs (RV/gen_submission_multi.py): ``num_models`` members predict the same ``n`` crops.

    Every member finds each visible keypoint with independent N(0, noise_px) noise (one query per label in a random
    slot, occasionally a second lower-score query nearby: the ensemble solver pools EVERY foreground query), misses a
    label with probability ``miss_frac`` and puts it grossly wrong (50..200 px) with probability ``outlier_frac`` --
    the 3-sigma filter's job.  Returns dict: logits [Nm,n,Q,12] f32 raw logits, points [Nm,n,Q,2] f32 normalised,
    boxes [n,4] int64, q_gt [n,4], t_gt [n,3]."""
    rng = np.random.default_rng(seed)
    logits = np.zeros((num_models, n, Q, 12), dtype=np.float32)
    points = rng.uniform(0.05, 0.95, (num_models, n, Q, 2)).astype(np.float32)
    boxes = np.zeros((n, 4), dtype=np.int64)
    q_gt = np.zeros((n, 4)); t_gt = np.zeros((n, 3))
    for i in range(n):
        q = _random_quat(rng)
        t = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.3, 0.3), rng.uniform(3.0, 20.0)])
        pc = TANGO_POINTS @ quat_to_rot(q).T + t
        uv = pc[:, :2] / pc[:, 2:3] * np.array([CAMERA_K[0, 0], CAMERA_K[1, 1]]) + CAMERA_K[:2, 2]
        q_gt[i], t_gt[i] = q, t
        lo, hi = uv.min(0), uv.max(0)
        side = int(max(hi - lo) * 1.2) + 8
        cx, cy = (lo + hi) / 2
        x1, y1 = int(cx - side / 2), int(cy - side / 2)
        boxes[i] = (x1, y1, x1 + side, y1 + side)
        visible = np.arange(11)
        if rng.random() < few_frac:
            visible = rng.permutation(11)[: rng.integers(0, 4)]            # < 4 keypoints: failure path
        for m in range(num_models):
            slots = rng.permutation(Q)
            used = 0
            for l_ in rng.permutation(visible):                            # label order differs between members
                if rng.random() < miss_frac:
                    continue
                p = uv[l_] + rng.normal(0, noise_px, 2)
                if len(visible) >= 6 and rng.random() < outlier_frac:
                    ang, mag = rng.uniform(0, 2 * np.pi), rng.uniform(50, 200)
                    p = p + mag * np.array([np.cos(ang), np.sin(ang)])
                reps = 2 if rng.random() < 0.15 else 1                     # a second, weaker query on the same keypoint
                for r in range(reps):
                    s = slots[used]; used += 1
                    top = rng.uniform(0.5, 0.99) if r == 0 else rng.uniform(0.3, 0.45)
                    logits[m, i, s, :] = np.log((1 - top) / 11)
                    logits[m, i, s, l_] = np.log(top)
                    points[m, i, s] = (p + (rng.normal(0, 2.0, 2) if r else 0.0) - boxes[i, :2]) / side
            for s in slots[used:]:
                top = rng.uniform(0.6, 0.99)
                logits[m, i, s, :] = np.log((1 - top) / 11)
                logits[m, i, s, 11] = np.log(top)
    return {"logits": logits, "points": points, "boxes": boxes, "q_gt": q_gt, "t_gt": t_gt}


# -------------------------------------------------------------------------------------------------------------
# SA drop: RT-DETR keypoint predictor (oracle/sa_model_ref.py) -- state_dict of
# configs/rtdetr_speed/rtdetr_r50vd_6x_speed_kl_1.yml (636 tensors, 31.3 M values), seeded
# -------------------------------------------------------------------------------------------------------------
def _bn2d(rng, sd, prefix, c, gamma_scale=1.0):
    _bn(rng, sd, prefix, c, gamma_scale)
    sd[prefix + ".num_batches_tracked"] = np.zeros((), dtype=np.int64)


def _conv_norm(rng, sd, prefix, cout, cin, k, gain=1.0, gamma_scale=1.0):
    sd[prefix + ".conv.weight"] = _conv(rng, cout, cin, k, gain)
    _bn2d(rng, sd, prefix + ".norm", cout, gamma_scale)


def _mlp(rng, sd, prefix, dims, last_gain=1.0):
    for i, (i_f, o_f) in enumerate(zip(dims[:-1], dims[1:])):
        _linear(rng, sd, f"{prefix}.layers.{i}", o_f, i_f)
        if i == len(dims) - 2:
            sd[f"{prefix}.layers.{i}.weight"] *= np.float32(last_gain)


def make_sa_state_dict(cfg=None, seed=0):
    """Random weights in the SA reference's ``state_dict`` layout.  Every tensor is non-degenerate on purpose: the
    reference's own init zeroes the sampling-offset / attention-weight projections and the last layer of every keypoint
    head (SA/src/zoo/rtdetr/rtdetr_decoder.py:70-92, :487-497), which would hide a wrong kernel behind zeros."""
    from .sa_model_ref import SaCfg, PRESNET_BLOCKS, STAGE_PLANES
    cfg = cfg or SaCfg()
    rng = np.random.default_rng(seed)
    sd = {"temper_param": rng.standard_normal(1).astype(np.float32)}
    b = "backbone"
    _conv_norm(rng, sd, b + ".conv1.conv1_1", 32, 3, 3)
    _conv_norm(rng, sd, b + ".conv1.conv1_2", 32, 32, 3)
    _conv_norm(rng, sd, b + ".conv1.conv1_3", 64, 32, 3)
    cin = 64
    bottleneck = cfg.depth >= 50
    exp = 4 if bottleneck else 1
    for si, (nb, planes) in enumerate(zip(PRESNET_BLOCKS[cfg.depth], STAGE_PLANES)):
        for bi in range(nb):
            p = f"{b}.res_layers.{si}.blocks.{bi}"
            if bottleneck:
                _conv_norm(rng, sd, p + ".branch2a", planes, cin, 1)
                _conv_norm(rng, sd, p + ".branch2b", planes, planes, 3)
                _conv_norm(rng, sd, p + ".branch2c", planes * 4, planes, 1, gamma_scale=0.5)
            else:
                _conv_norm(rng, sd, p + ".branch2a", planes, cin, 3)
                _conv_norm(rng, sd, p + ".branch2b", planes, planes, 3, gamma_scale=0.5)
            if bi == 0:
                _conv_norm(rng, sd, p + (".short" if si == 0 else ".short.conv"), planes * exp, cin, 1, gamma_scale=0.5)
                cin = planes * exp
    e, E = "encoder", cfg.hidden_dim
    for i, c in enumerate((128 * exp, 256 * exp, 512 * exp)):
        sd[f"{e}.input_proj.{i}.0.weight"] = _conv(rng, E, c, 1, gain=0.7)
        _bn2d(rng, sd, f"{e}.input_proj.{i}.1", E)
    sd[e + ".encoder_fusion_input.weight"] = _conv(rng, 256, 3 * E, 1)        # defined, never used by forward
    p = e + ".encoder.0.layers.0"
    _mha(rng, sd, p + ".self_attn", E)
    _linear(rng, sd, p + ".linear1", cfg.enc_ff, E)
    _linear(rng, sd, p + ".linear2", E, cfg.enc_ff)
    _ln(rng, sd, p + ".norm1", E)
    _ln(rng, sd, p + ".norm2", E)
    for i in range(2):
        _conv_norm(rng, sd, f"{e}.lateral_convs.{i}", E, E, 1)
    for grp in ("fpn_blocks", "pan_blocks"):
        for i in range(2):
            p = f"{e}.{grp}.{i}"
            h = cfg.csp_hidden
            _conv_norm(rng, sd, p + ".conv1", h, 2 * E, 1)
            _conv_norm(rng, sd, p + ".conv2", h, 2 * E, 1)
            _conv_norm(rng, sd, p + ".bottlenecks.0.conv1", h, h, 3, gain=0.8)
            _conv_norm(rng, sd, p + ".bottlenecks.0.conv2", h, h, 1, gain=0.6)
            _conv_norm(rng, sd, p + ".conv3", E, h, 1)
    d = "decoder"
    for i in range(cfg.num_levels):
        sd[f"{d}.input_proj.{i}.conv.weight"] = _conv(rng, E, E, 1, gain=0.7)
        _bn2d(rng, sd, f"{d}.input_proj.{i}.norm", E)
    nl_np = cfg.nheads * cfg.num_levels * cfg.num_points
    for i in range(cfg.dec_layers):
        p = f"{d}.decoder.layers.{i}"
        _mha(rng, sd, p + ".self_attn", E)
        _ln(rng, sd, p + ".norm1", E)
        _linear(rng, sd, p + ".cross_attn.sampling_offsets", nl_np * 2, E)
        # offsets of a few cells around the reference point, like the reference's grid_init bias (:72-84)
        sd[p + ".cross_attn.sampling_offsets.bias"] = rng.uniform(-3.0, 3.0, nl_np * 2).astype(np.float32)
        _linear(rng, sd, p + ".cross_attn.attention_weights", nl_np, E)
        _linear(rng, sd, p + ".cross_attn.value_proj", E, E)
        _linear(rng, sd, p + ".cross_attn.output_proj", E, E)
        _ln(rng, sd, p + ".norm2", E)
        _linear(rng, sd, p + ".linear1", cfg.dec_ff, E)
        _linear(rng, sd, p + ".linear2", E, cfg.dec_ff)
        _ln(rng, sd, p + ".norm3", E)
    for i in range(cfg.dec_layers):
        _mlp(rng, sd, f"{d}.decoder.sigma_embed.{i}", (E, E, E, 1))
    _mlp(rng, sd, d + ".query_pos_head", (2, 2 * E, E))
    _linear(rng, sd, d + ".enc_output.0", E, E)
    _ln(rng, sd, d + ".enc_output.1", E)
    _linear(rng, sd, d + ".enc_score_head", cfg.num_classes + 1, E)
    sd[d + ".enc_score_head.weight"] *= np.float32(4.0)          # spread the anchor scores: robust top-k selection
    _mlp(rng, sd, d + ".enc_bbox_head", (E, E, E, 2), last_gain=0.5)
    for i in range(cfg.dec_layers):
        _linear(rng, sd, f"{d}.dec_score_head.{i}", cfg.num_classes + 1, E)
    for i in range(cfg.dec_layers):
        _mlp(rng, sd, f"{d}.dec_bbox_head.{i}", (E, E, E, 2), last_gain=0.5)
    return {k: torch.from_numpy(np.ascontiguousarray(v)).reshape(v.shape) for k, v in sd.items()}
