"""CPU restatement of the reference keypoint-set predictor forward (TEST INFRASTRUCTURE; see oracle/__init__.py).

Plain functional PyTorch fp32 driven by a ``state_dict`` in the reference key layout.
Each function cites the reference lines it follows.  Pinned against the real reference modules by
``tests/golden/model_*.npz`` (written by oracle/make_golden.py) and, when /root/reference is mounted, directly
by tests/test_oracle_model.py::test_matches_live_reference.
"""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class ModelCfg:
    backbone: str = "resnet50s8"     # 'resnet50s8' (Backbone8s) or 'resnet50' (Backbone, stride 16)
    num_queries: int = 40
    enc_layers: int = 4
    dec_layers: int = 4
    hidden_dim: int = 256
    nheads: int = 8
    dim_feedforward: int = 2048
    aux_loss: bool = True
    sigma_head: bool = False         # self-assessment variant: per-keypoint log-sigma head
    position_embedding: str = "sine"  # 'sine' ('v2') or 'learned' ('v3'), RV/models/position_encoding.py:84-95

    @property
    def stride8(self):
        return self.backbone not in ("resnet18", "resnet34", "resnet50")


RESNET50_BLOCKS = {"layer1": 3, "layer2": 4, "layer3": 6}


def frozen_bn(x, sd, prefix):
    """FrozenBatchNorm2d.forward, RV/models/backbone.py:44-54 (eps 1e-5, same op order)."""
    w = sd[prefix + ".weight"].reshape(1, -1, 1, 1)
    b = sd[prefix + ".bias"].reshape(1, -1, 1, 1)
    rv = sd[prefix + ".running_var"].reshape(1, -1, 1, 1)
    rm = sd[prefix + ".running_mean"].reshape(1, -1, 1, 1)
    scale = w * (rv + 1e-5).rsqrt()
    bias = b - rm * scale
    return x * scale + bias


def bottleneck(x, sd, p, stride, has_down):
    """torchvision.models.resnet.Bottleneck.forward (v1.5: stride on conv2), as instantiated by
    RV/models/backbone.py:95-99 / :113-117 with norm_layer=FrozenBatchNorm2d."""
    out = F.relu(frozen_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
    out = F.relu(frozen_bn(F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1), sd, p + ".bn2"))
    out = frozen_bn(F.conv2d(out, sd[p + ".conv3.weight"]), sd, p + ".bn3")
    if has_down:
        idt = frozen_bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
    else:
        idt = x
    return F.relu(out + idt)


def resnet_body(x, sd, taps=None):
    """conv1/bn1/relu/maxpool + layer1..layer3 of torchvision resnet50 (IntermediateLayerGetter,
    RV/models/backbone.py:120-123).  Returns (layer2_out, layer3_out)."""
    b = "backbone.0.body"
    x = F.conv2d(x, sd[b + ".conv1.weight"], stride=2, padding=3)
    x = F.relu(frozen_bn(x, sd, b + ".bn1"))
    if taps is not None:
        taps["stem"] = x
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = {}
    for li, name in enumerate(("layer1", "layer2", "layer3")):
        for bi in range(RESNET50_BLOCKS[name]):
            stride = 2 if (bi == 0 and li > 0) else 1
            x = bottleneck(x, sd, f"{b}.{name}.{bi}", stride, bi == 0)
        outs[name] = x
        if taps is not None:
            taps[name] = x
    return outs["layer2"], outs["layer3"]


def backbone8s(x, sd, taps=None):
    """Backbone8s.forward, RV/models/backbone.py:133-149."""
    xs8, xs16 = resnet_body(x, sd, taps)
    b = "backbone.0"
    xs8 = F.conv2d(xs8, sd[b + ".s8_latern.weight"])
    up = F.interpolate(xs16, scale_factor=2, mode="bilinear", align_corners=True)  # nn.UpsamplingBilinear2d
    xs16 = F.conv2d(up, sd[b + ".s16_latern.weight"], padding=1)
    out = F.conv2d(torch.cat([xs8, xs16], 1), sd[b + ".output_conv.weight"], sd[b + ".output_conv.bias"], padding=1)
    if taps is not None:
        taps["neck"] = out
    return out


def backbone16(x, sd, taps=None):
    """Backbone.forward (stride 16, layer3 output), RV/models/backbone.py:76-102."""
    _, xs16 = resnet_body(x, sd, taps)
    return xs16


def position_embedding_sine(B, H, W, hidden_dim=256):
    """PositionEmbeddingSine.forward with an all-False mask (equal-size batch), normalize=True,
    RV/models/position_encoding.py:30-53, built by :84-89 with N_steps = hidden_dim // 2."""
    num_pos_feats = hidden_dim // 2
    not_mask = torch.ones(B, H, W, dtype=torch.bool)
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps, scale = 1e-6, 2 * math.pi
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = 10000 ** (2 * (dim_t // 2) / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def mha(query, key, value, sd, p, nheads):
    """nn.MultiheadAttention forward (seq-first, no masks, eval) as used at RV/models/transformer.py:157, :225,
    :229: packed in_proj (rows 0:E q, E:2E k, 2E:3E v), q scaled by 1/sqrt(head_dim), softmax, out_proj."""
    L, B, E = query.shape
    S = key.shape[0]
    hd = E // nheads
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(query, w[:E], b[:E])
    k = F.linear(key, w[E:2 * E], b[E:2 * E])
    v = F.linear(value, w[2 * E:], b[2 * E:])
    q = q.reshape(L, B * nheads, hd).transpose(0, 1) * (hd ** -0.5)
    k = k.reshape(S, B * nheads, hd).transpose(0, 1)
    v = v.reshape(S, B * nheads, hd).transpose(0, 1)
    attn = torch.softmax(torch.bmm(q, k.transpose(1, 2)), dim=-1)
    out = torch.bmm(attn, v).transpose(0, 1).reshape(L, B, E)
    return F.linear(out, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def layer_norm(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def encoder_layer(src, pos, sd, p, nheads):
    """TransformerEncoderLayer.forward_post, RV/models/transformer.py:154-167 (dropout inactive in eval)."""
    q = k = src + pos
    src2 = mha(q, k, src, sd, p + ".self_attn", nheads)
    src = layer_norm(src + src2, sd, p + ".norm1")
    src2 = F.linear(F.relu(F.linear(src, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                    sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return layer_norm(src + src2, sd, p + ".norm2")


def decoder_layer(tgt, memory, pos, query_pos, sd, p, nheads):
    """TransformerDecoderLayer.forward_post, RV/models/transformer.py:218-239."""
    q = k = tgt + query_pos
    tgt2 = mha(q, k, tgt, sd, p + ".self_attn", nheads)
    tgt = layer_norm(tgt + tgt2, sd, p + ".norm1")
    tgt2 = mha(tgt + query_pos, memory + pos, memory, sd, p + ".multihead_attn", nheads)
    tgt = layer_norm(tgt + tgt2, sd, p + ".norm2")
    tgt2 = F.linear(F.relu(F.linear(tgt, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                    sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return layer_norm(tgt + tgt2, sd, p + ".norm3")


def mlp3(x, sd, p):
    """MLP.forward with num_layers=3, RV/models/detr_speed.py:16-29."""
    x = F.relu(F.linear(x, sd[p + ".layers.0.weight"], sd[p + ".layers.0.bias"]))
    x = F.relu(F.linear(x, sd[p + ".layers.1.weight"], sd[p + ".layers.1.bias"]))
    return F.linear(x, sd[p + ".layers.2.weight"], sd[p + ".layers.2.bias"])


@torch.no_grad()
def forward(sd, cfg: ModelCfg, images, taps=None):
    """DETR.forward, RV/models/detr_speed.py:59-92, for an equal-size batch ``images`` [B,3,R,R] float32.

    Returns {'pred_logits' [B,Q,12], 'pred_points' [B,Q,2], ('pred_sigmas' [B,Q,2]), 'aux_outputs': [...]}.
    """
    sd = {k: v.float() for k, v in sd.items()}
    x = images.float()
    B = x.shape[0]
    feat = backbone8s(x, sd, taps) if cfg.stride8 else backbone16(x, sd, taps)
    _, _, H, W = feat.shape
    if cfg.position_embedding in ("learned", "v3"):
        # PositionEmbeddingLearned.forward, RV/models/position_encoding.py:69-81: cat(col_embed[x], row_embed[y])
        x_emb = sd["backbone.1.col_embed.weight"][:W]
        y_emb = sd["backbone.1.row_embed.weight"][:H]
        pos = torch.cat([x_emb.unsqueeze(0).repeat(H, 1, 1), y_emb.unsqueeze(1).repeat(1, W, 1)], dim=-1)
        pos = pos.permute(2, 0, 1).unsqueeze(0).repeat(B, 1, 1, 1)
    else:
        pos = position_embedding_sine(B, H, W, cfg.hidden_dim)
    src = F.conv2d(feat, sd["input_proj.weight"], sd["input_proj.bias"])          # detr_speed.py:54-55, :81
    # Transformer.forward, RV/models/transformer.py:51-63
    src = src.flatten(2).permute(2, 0, 1)
    pos = pos.flatten(2).permute(2, 0, 1)
    query_embed = sd["query_embed.weight"].unsqueeze(1).repeat(1, B, 1)
    memory = src
    for i in range(cfg.enc_layers):
        memory = encoder_layer(memory, pos, sd, f"transformer.encoder.layers.{i}", cfg.nheads)
        if taps is not None:
            taps[f"enc{i}"] = memory
    if taps is not None:
        taps["input_proj"] = src
    tgt = torch.zeros_like(query_embed)
    inter = []
    out = tgt
    for i in range(cfg.dec_layers):                                                # transformer.py:111-118
        out = decoder_layer(out, memory, pos, query_embed, sd, f"transformer.decoder.layers.{i}", cfg.nheads)
        inter.append(layer_norm(out, sd, "transformer.decoder.norm"))
    hs = torch.stack(inter).transpose(1, 2)                                        # [L, B, Q, C]
    if taps is not None:
        taps["hs"] = hs
    logits = F.linear(hs, sd["cls_embed.weight"], sd["cls_embed.bias"])            # detr_speed.py:83
    points = mlp3(hs, sd, "point_embed").sigmoid()                                 # detr_speed.py:84
    res = {"pred_logits": logits[-1], "pred_points": points[-1]}
    if cfg.sigma_head:
        # SA/src/zoo/rtdetr/rtdetr_decoder.py:295-297, :367: one log-sigma per query, repeated to (x, y)
        s = mlp3(hs[-1], sd, "sigma_embed")
        res["pred_sigmas"] = s.repeat(1, 1, 2)
    if cfg.aux_loss:
        res["aux_outputs"] = [{"pred_logits": a, "pred_points": b} for a, b in zip(logits[:-1], points[:-1])]
    return res
