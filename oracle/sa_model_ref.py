"""CPU restatement of the SA drop's RT-DETR keypoint predictor forward (TEST INFRASTRUCTURE; see oracle/__init__.py).

SA = "Monocular Satellite Pose Estimation Based on Uncertainty Estimation and Self-Assessment" (second code drop under
/root/reference).  Model of ``configs/rtdetr_speed/rtdetr_r50vd_6x_speed_kl_1.yml``: PResNet-50-vd -> HybridEncoder
(AIFI on the stride-32 level, CSPRep top-down / bottom-up fusion, bicubic x0.5 instead of the stride-2 convolutions) ->
RTDETRTransformer (top-k query selection, 3 decoder layers with multi-scale deformable cross-attention, iterative
keypoint refinement, per-layer log-sigma head), eval mode.

Plain functional PyTorch fp32 driven by the reference ``state_dict`` (636 tensors at this config).  Every function
cites the reference lines it follows.  Pinned against the LIVE reference model (oracle/ref_import.py:import_sa_rtdetr)
by oracle/make_golden.py (``sa_model_golden.npz``) and, when /root/reference is mounted, by
tests/test_oracle.py::test_sa_model_ref_matches_live_reference.
"""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class SaCfg:
    input_size: int = 256            # eval_spatial_size of the speed configs
    num_queries: int = 30
    dec_layers: int = 3
    hidden_dim: int = 256
    nheads: int = 8
    enc_ff: int = 1024               # HybridEncoder.dim_feedforward (AIFI layer, GELU)
    dec_ff: int = 1024               # RTDETRTransformer.dim_feedforward (ReLU)
    csp_hidden: int = 128            # CSPRepLayer hidden channels = hidden_dim * expansion (0.5)
    num_levels: int = 3
    num_points: int = 4
    num_classes: int = 11            # + 1 background logit
    anchor_eps: float = 1e-2
    depth: int = 50                  # PResNet depth: 50 (BottleNeck, rtdetr_r50vd_*.yml) or 18 / 34 (BasicBlock, rtdetr_r18vd_*.yml)


PRESNET50_BLOCKS = (3, 4, 6, 3)
PRESNET_BLOCKS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3)}      # ResNet_cfg, SA/nn/backbone/presnet.py:18-24
STAGE_PLANES = (64, 128, 256, 512)


def bn(x, sd, p, eps=1e-5):
    """nn.BatchNorm2d in eval mode (the kl configs set freeze_norm False; FrozenBatchNorm2d of SA/nn/backbone/
    common.py:27-80 is the same arithmetic)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, eps)


def act(x, name):
    if name is None:
        return x
    return {"relu": F.relu, "silu": F.silu, "gelu": F.gelu}[name](x)


def conv_norm(x, sd, p, stride=1, a=None):
    """ConvNormLayer.forward (SA/nn/backbone/common.py:8-24, SA/src/zoo/rtdetr/hybrid_encoder.py:17-35):
    bias-free conv with padding (k-1)//2 -> BatchNorm2d -> activation."""
    w = sd[p + ".conv.weight"]
    return act(bn(F.conv2d(x, w, stride=stride, padding=(w.shape[-1] - 1) // 2), sd, p + ".norm"), a)


def bottleneck_vd(x, sd, p, stride, shortcut):
    """BottleNeck.forward, variant 'd' (SA/nn/backbone/presnet.py:73-123): stride on the 3x3 convolution; the first
    block's shortcut is AvgPool2d(2, 2) + 1x1 conv when stride == 2, a plain 1x1 conv otherwise."""
    out = conv_norm(x, sd, p + ".branch2a", 1, "relu")
    out = conv_norm(out, sd, p + ".branch2b", stride, "relu")
    out = conv_norm(out, sd, p + ".branch2c", 1, None)
    if shortcut:
        short = x
    elif stride == 2:
        short = conv_norm(F.avg_pool2d(x, 2, 2, 0, ceil_mode=True), sd, p + ".short.conv", 1, None)
    else:
        short = conv_norm(x, sd, p + ".short", 1, None)
    return F.relu(out + short)


def basic_block_vd(x, sd, p, stride, shortcut):
    """BasicBlock.forward, variant 'd' (SA/nn/backbone/presnet.py:35-70): 3x3 (stride) + BN + ReLU, 3x3 + BN, shortcut as
    in the bottleneck (AvgPool2d(2, 2) + 1x1 when stride == 2, plain 1x1 in the first stage), ReLU of the sum."""
    out = conv_norm(x, sd, p + ".branch2a", stride, "relu")
    out = conv_norm(out, sd, p + ".branch2b", 1, None)
    if shortcut:
        short = x
    elif stride == 2:
        short = conv_norm(F.avg_pool2d(x, 2, 2, 0, ceil_mode=True), sd, p + ".short.conv", 1, None)
    else:
        short = conv_norm(x, sd, p + ".short", 1, None)
    return F.relu(out + short)


def presnet(x, sd, taps=None, depth=50):
    """PResNet.forward (SA/nn/backbone/presnet.py:245-265), variant d, return_idx [1, 2, 3]; depth 50 = BottleNeck,
    18 / 34 = BasicBlock."""
    b = "backbone"
    x = conv_norm(x, sd, b + ".conv1.conv1_1", 2, "relu")
    x = conv_norm(x, sd, b + ".conv1.conv1_2", 1, "relu")
    x = conv_norm(x, sd, b + ".conv1.conv1_3", 1, "relu")
    if taps is not None:
        taps["stem"] = x
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = []
    block = bottleneck_vd if depth >= 50 else basic_block_vd
    for si, nb in enumerate(PRESNET_BLOCKS[depth]):
        for bi in range(nb):
            stride = 2 if (bi == 0 and si != 0) else 1          # Blocks.__init__: stage_num != 2 (:137)
            x = block(x, sd, f"{b}.res_layers.{si}.blocks.{bi}", stride, shortcut=bi != 0)
        if taps is not None:
            taps[f"stage{si}"] = x
        if si >= 1:
            outs.append(x)
    return outs


def sincos_pos_embed(w, h, embed_dim=256, temperature=10000.0):
    """HybridEncoder.build_2d_sincos_position_embedding (hybrid_encoder.py:306-326)."""
    grid_w = torch.arange(int(w), dtype=torch.float32)
    grid_h = torch.arange(int(h), dtype=torch.float32)
    grid_w, grid_h = torch.meshgrid(grid_w, grid_h, indexing="ij")
    pos_dim = embed_dim // 4
    omega = torch.arange(pos_dim, dtype=torch.float32) / pos_dim
    omega = 1.0 / (temperature ** omega)
    out_w = grid_w.flatten()[..., None] @ omega[None]
    out_h = grid_h.flatten()[..., None] @ omega[None]
    return torch.concat([out_w.sin(), out_w.cos(), out_h.sin(), out_h.cos()], dim=1)[None, :, :]


def mha(query, key, value, sd, p, nheads):
    """nn.MultiheadAttention(batch_first=True) forward in eval mode: packed in_proj, scaled dot-product, out_proj."""
    E = query.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(query, w[:E], b[:E])
    k = F.linear(key, w[E:2 * E], b[E:2 * E])
    v = F.linear(value, w[2 * E:], b[2 * E:])
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    hd = E // nheads
    q = q.reshape(B, Lq, nheads, hd).transpose(1, 2)
    k = k.reshape(B, Lk, nheads, hd).transpose(1, 2)
    v = v.reshape(B, Lk, nheads, hd).transpose(1, 2)
    a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, Lq, E)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def layer_norm(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def aifi_layer(src, pos, sd, p, nheads):
    """TransformerEncoderLayer.forward, post-norm, GELU (hybrid_encoder.py:153-173)."""
    q = src + pos
    src = layer_norm(src + mha(q, q, src, sd, p + ".self_attn", nheads), sd, p + ".norm1")
    ff = F.linear(F.gelu(F.linear(src, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                  sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return layer_norm(src + ff, sd, p + ".norm2")


def rep_vgg(x, sd, p):
    """RepVggBlock.forward, training-form branches (hybrid_encoder.py:47-53): silu(conv3x3+bn + conv1x1+bn)."""
    return F.silu(conv_norm(x, sd, p + ".conv1") + conv_norm(x, sd, p + ".conv2"))


def csp_rep(x, sd, p):
    """CSPRepLayer.forward with one bottleneck and hidden != out (hybrid_encoder.py:119-123)."""
    x1 = rep_vgg(conv_norm(x, sd, p + ".conv1", 1, "silu"), sd, p + ".bottlenecks.0")
    x2 = conv_norm(x, sd, p + ".conv2", 1, "silu")
    return conv_norm(x1 + x2, sd, p + ".conv3", 1, "silu")


def hybrid_encoder(feats, sd, cfg: SaCfg, taps=None):
    """HybridEncoder.forward (hybrid_encoder.py:328-401), use_encoder_idx [2], one AIFI layer."""
    e = "encoder"
    proj = []
    for i, f in enumerate(feats):
        y = F.conv2d(f, sd[f"{e}.input_proj.{i}.0.weight"])
        proj.append(bn(y, sd, f"{e}.input_proj.{i}.1"))
    B, C, h, w = proj[2].shape
    src = proj[2].flatten(2).permute(0, 2, 1)
    pos = sincos_pos_embed(w, h, cfg.hidden_dim)
    mem = aifi_layer(src, pos, sd, f"{e}.encoder.0.layers.0", cfg.nheads)
    proj[2] = mem.permute(0, 2, 1).reshape(B, C, h, w).contiguous()
    if taps is not None:
        taps["aifi"] = proj[2]
    inner = [proj[2]]
    for idx in (2, 1):                                            # top-down
        high = conv_norm(inner[0], sd, f"{e}.lateral_convs.{2 - idx}", 1, "silu")
        inner[0] = high
        up = F.interpolate(high, scale_factor=2.0, mode="nearest")
        inner.insert(0, csp_rep(torch.concat([up, proj[idx - 1]], dim=1), sd, f"{e}.fpn_blocks.{2 - idx}"))
    outs = [inner[0]]
    for idx in (0, 1):                                            # bottom-up, bicubic x0.5 (:394)
        down = F.interpolate(outs[-1], scale_factor=0.5, mode="bicubic")
        outs.append(csp_rep(torch.concat([down, inner[idx + 1]], dim=1), sd, f"{e}.pan_blocks.{idx}"))
    if taps is not None:
        for i, o in enumerate(outs):
            taps[f"enc_out{i}"] = o
    return outs


def mlp(x, sd, p, n):
    """MLP.forward (rtdetr_decoder.py:24-37), ReLU between layers."""
    for i in range(n):
        x = F.linear(x, sd[f"{p}.layers.{i}.weight"], sd[f"{p}.layers.{i}.bias"])
        if i < n - 1:
            x = F.relu(x)
    return x


def inverse_sigmoid(x, eps=1e-5):
    """SA/src/zoo/rtdetr/utils.py:10-12."""
    x = x.clip(min=0.0, max=1.0)
    return torch.log(x.clip(min=eps) / (1 - x).clip(min=eps))


def make_anchors(cfg: SaCfg):
    """RTDETRTransformer._generate_anchors (rtdetr_decoder.py:570-600): logit of the cell centres, two coordinates."""
    anchors = []
    for s in (8, 16, 32):
        h = w = cfg.input_size // s
        gy, gx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        g = (torch.stack([gx, gy], -1).unsqueeze(0) + 0.5) / torch.tensor([w, h], dtype=torch.float32)
        anchors.append(g.reshape(-1, h * w, 2))
    a = torch.concat(anchors, 1)
    valid = ((a > cfg.anchor_eps) * (a < 1 - cfg.anchor_eps)).all(-1, keepdim=True)
    a = torch.log(a / (1 - a))
    return torch.where(valid, a, torch.inf)


def deform_core(value, shapes, loc, aw):
    """deformable_attention_core_func (SA/src/zoo/rtdetr/utils.py:15-64)."""
    bs, _, nh, c = value.shape
    _, Lq, _, nl, npnt, _ = loc.shape
    vl = value.split([h * w for h, w in shapes], dim=1)
    grids = 2 * loc - 1
    sampled = []
    for lvl, (h, w) in enumerate(shapes):
        v = vl[lvl].flatten(2).permute(0, 2, 1).reshape(bs * nh, c, h, w)
        g = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).flatten(0, 1)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    aw = aw.permute(0, 2, 1, 3, 4).reshape(bs * nh, 1, Lq, nl * npnt)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1).reshape(bs, nh * c, Lq)
    return out.permute(0, 2, 1)


def ms_deform_attn(query, ref, memory, shapes, sd, p, cfg: SaCfg):
    """MSDeformableAttention.forward with 2-coordinate reference points (rtdetr_decoder.py:100-191): the offsets are
    divided by the level's (W, H) and added to the SAME reference point on every level."""
    bs, Lq, _ = query.shape
    nh, nl, npnt = cfg.nheads, cfg.num_levels, cfg.num_points
    value = F.linear(memory, sd[p + ".value_proj.weight"], sd[p + ".value_proj.bias"]).reshape(bs, -1, nh, cfg.hidden_dim // nh)
    off = F.linear(query, sd[p + ".sampling_offsets.weight"], sd[p + ".sampling_offsets.bias"]).reshape(bs, Lq, nh, nl, npnt, 2)
    aw = F.linear(query, sd[p + ".attention_weights.weight"], sd[p + ".attention_weights.bias"]).reshape(bs, Lq, nh, nl * npnt)
    aw = F.softmax(aw, dim=-1).reshape(bs, Lq, nh, nl, npnt)
    norm = torch.tensor(shapes, dtype=torch.float32).flip([1]).reshape(1, 1, 1, nl, 1, 2)
    loc = ref[:, :, None, None, None, :] + off / norm
    out = deform_core(value, shapes, loc, aw)
    return F.linear(out, sd[p + ".output_proj.weight"], sd[p + ".output_proj.bias"])


def decoder_layer(tgt, ref, memory, shapes, qpos, sd, p, cfg: SaCfg):
    """TransformerDecoderLayer.forward (rtdetr_decoder.py:245-290), post-norm, ReLU feed-forward."""
    q = tgt + qpos
    tgt = layer_norm(tgt + mha(q, q, tgt, sd, p + ".self_attn", cfg.nheads), sd, p + ".norm1")
    tgt = layer_norm(tgt + ms_deform_attn(tgt + qpos, ref, memory, shapes, sd, p + ".cross_attn", cfg), sd, p + ".norm2")
    ff = F.linear(F.relu(F.linear(tgt, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                  sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return layer_norm(tgt + ff, sd, p + ".norm3")


def rtdetr_decoder(feats, sd, cfg: SaCfg, taps=None, topk_override=None):
    """RTDETRTransformer.forward in eval mode without denoising (rtdetr_decoder.py:686-751) and
    TransformerDecoder.forward (:300-372)."""
    d = "decoder"
    proj, shapes = [], []
    for i, f in enumerate(feats):
        y = bn(F.conv2d(f, sd[f"{d}.input_proj.{i}.conv.weight"]), sd, f"{d}.input_proj.{i}.norm")
        shapes.append([y.shape[2], y.shape[3]])
        proj.append(y.flatten(2).permute(0, 2, 1))
    memory = torch.concat(proj, 1)                                       # [B, 1344, 256]
    om = F.linear(memory, sd[d + ".enc_output.0.weight"], sd[d + ".enc_output.0.bias"])
    om = F.layer_norm(om, (om.shape[-1],), sd[d + ".enc_output.1.weight"], sd[d + ".enc_output.1.bias"], 1e-5)
    enc_cls = F.linear(om, sd[d + ".enc_score_head.weight"], sd[d + ".enc_score_head.bias"])
    enc_xy = mlp(om, sd, d + ".enc_bbox_head", 3) + make_anchors(cfg)
    scores = enc_cls.max(-1).values
    _, topk = torch.topk(scores, cfg.num_queries, dim=1)
    if topk_override is not None:
        topk = topk_override
    ref_unact = enc_xy.gather(1, topk.unsqueeze(-1).repeat(1, 1, 2))
    enc_topk_pts = torch.sigmoid(ref_unact)
    enc_topk_logits = enc_cls.gather(1, topk.unsqueeze(-1).repeat(1, 1, enc_cls.shape[-1]))
    tgt = om.gather(1, topk.unsqueeze(-1).repeat(1, 1, om.shape[-1]))
    if taps is not None:
        taps.update(memory=memory, enc_scores=scores, topk=topk, enc_out_memory=om, ref0=enc_topk_pts)
    ref = torch.sigmoid(ref_unact)
    pts, logits, sigmas = [], [], []
    for i in range(cfg.dec_layers):
        qpos = mlp(ref, sd, d + ".query_pos_head", 2)
        tgt = decoder_layer(tgt, ref, memory, shapes, qpos, sd, f"{d}.decoder.layers.{i}", cfg)
        ref = torch.sigmoid(mlp(tgt, sd, f"{d}.dec_bbox_head.{i}", 3) + inverse_sigmoid(ref))
        logits.append(F.linear(tgt, sd[f"{d}.dec_score_head.{i}.weight"], sd[f"{d}.dec_score_head.{i}.bias"]))
        pts.append(ref)
        sigmas.append(mlp(tgt, sd, f"{d}.decoder.sigma_embed.{i}", 3).repeat(1, 1, 2))
        if taps is not None:
            taps[f"dec{i}"] = tgt
    out = {"pred_logits": logits[-1], "pred_pts": pts[-1], "pred_sigmas": sigmas[-1]}
    out["aux_outputs"] = [{"pred_logits": a, "pred_pts": b, "pred_sigmas": c}
                          for a, b, c in zip(logits[:-1], pts[:-1], sigmas[:-1])]
    out["aux_outputs"].append({"pred_logits": enc_topk_logits, "pred_pts": enc_topk_pts})
    return out


def forward(sd, cfg: SaCfg, images, taps=None, topk_override=None):
    """RTDETR.forward (SA/src/zoo/rtdetr/rtdetr.py:36-52), eval mode."""
    with torch.no_grad():
        feats = presnet(images, sd, taps, cfg.depth)
        feats = hybrid_encoder(feats, sd, cfg, taps)
        return rtdetr_decoder(feats, sd, cfg, taps, topk_override)
