#!/usr/bin/env python
"""Which stage's TF32 rounding costs how much keypoint error?  (CPU emulation; build-time analysis tool, not product)

Runs the oracle forward with cvt.rna.tf32 rounding applied to the operands (activations AND weights) of the GEMMs /
convolutions of ONE stage group at a time (fp32 accumulation, everything else exact) and reports the resulting keypoint
error in pixels at the largest crop side (1748 px).  The decoder and the heads run 3xTF32 on the GPU (error ~1e-6) and
are therefore treated as exact.  Usage: python tests/probes/tf32_budget.py [n_images] [--spread]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import crop_ref, model_ref, synth  # noqa: E402


def rna_tf32(x):
    u = x.contiguous().view(torch.int32)
    finite = (u & 0x7f800000) != 0x7f800000
    r = torch.where(finite, (u + 0x1000) & ~0x1fff, u)
    return r.view(torch.float32)


class Ctx:
    stage = "none"
    active = set()
    log = []


def _round_if(x):
    return rna_tf32(x) if Ctx.stage in Ctx.active or "all" in Ctx.active else x


class FProxy:
    def __init__(self, real):
        self._r = real

    def __getattr__(self, k):
        return getattr(self._r, k)

    def conv2d(self, x, w, b=None, **kw):
        return self._r.conv2d(_round_if(x), _round_if(w), b, **kw)

    def linear(self, x, w, b=None):
        return self._r.linear(_round_if(x), _round_if(w), b)


class TorchProxy:
    def __init__(self, real):
        self._r = real

    def __getattr__(self, k):
        return getattr(self._r, k)

    def bmm(self, a, b):
        if Ctx.stage.startswith("enc"):
            st, Ctx.stage = Ctx.stage, Ctx.stage.split(".")[0] + ".attn"
            out = self._r.bmm(_round_if(a), _round_if(b))
            Ctx.stage = st
            return out
        return self._r.bmm(a, b)


def install():
    """stage tracking by wrapping the oracle's layer functions"""
    model_ref.F = FProxy(torch.nn.functional)
    model_ref.torch = TorchProxy(torch)
    orig_bottleneck, orig_enc, orig_dec, orig_mlp3 = (model_ref.bottleneck, model_ref.encoder_layer,
                                                      model_ref.decoder_layer, model_ref.mlp3)
    orig_mha, orig_body, orig_b8 = model_ref.mha, model_ref.resnet_body, model_ref.backbone8s

    def bottleneck(x, sd, p, stride, has_down):
        Ctx.stage = p.split(".")[3]                 # layer1 / layer2 / layer3
        out = orig_bottleneck(x, sd, p, stride, has_down)
        Ctx.stage = "neck"
        return out

    def resnet_body(x, sd, taps=None):
        Ctx.stage = "stem"
        return orig_body(x, sd, taps)

    def encoder_layer(src, pos, sd, p, nheads):
        i = p.split(".")[-1]
        Ctx.stage = f"enc{i}.ffn"                   # mha() switches to .proj while it runs
        return orig_enc(src, pos, sd, p, nheads)

    def mha(query, key, value, sd, p, nheads):
        st = Ctx.stage
        if st.startswith("enc"):
            Ctx.stage = st.split(".")[0] + ".proj"
        out = orig_mha(query, key, value, sd, p, nheads)
        Ctx.stage = st
        return out

    def decoder_layer(*a, **k):
        Ctx.stage = "dec"
        return orig_dec(*a, **k)

    def mlp3(x, sd, p):
        Ctx.stage = "heads"
        return orig_mlp3(x, sd, p)

    model_ref.bottleneck, model_ref.resnet_body, model_ref.encoder_layer = bottleneck, resnet_body, encoder_layer
    model_ref.mha, model_ref.decoder_layer, model_ref.mlp3 = mha, decoder_layer, mlp3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
    spread = "--spread" in sys.argv
    torch.set_num_threads(os.cpu_count())
    cfg = model_ref.ModelCfg(aux_loss=False)
    sd = synth.make_state_dict(cfg, seed=0, spread_labels=spread)
    det = synth.load_detector_boxes()
    frames = synth.make_frames(8, det, seed=100)
    x = torch.stack([crop_ref.crop_resize_normalize(np.roll(frames[i % 8], 37 * (i // 8), axis=1), det[i], 224)[0]
                     for i in range(n)])
    install()
    Ctx.active = set()
    ref = model_ref.forward(sd, cfg, x)
    groups = ["stem", "layer1", "layer2", "layer3", "neck", "enc_input_proj"]
    for i in range(cfg.enc_layers):
        groups += [f"enc{i}.proj", f"enc{i}.attn", f"enc{i}.ffn"]
    rows = []
    # input_proj runs under stage "neck" in the oracle (called right after the backbone); split it out by name
    orig_conv = model_ref.F.conv2d

    for g in groups + ["trunk (everything the GPU runs in plain TF32)"]:
        if g.startswith("trunk"):
            Ctx.active = set(groups) | {"neck"}
        elif g == "enc_input_proj":
            continue
        else:
            Ctx.active = {g}
        out = model_ref.forward(sd, cfg, x)
        d = (out["pred_points"] - ref["pred_points"]).abs()
        dl = (out["pred_logits"] - ref["pred_logits"]).abs()
        rows.append((g, d.max().item() * 1748, d.pow(2).mean().sqrt().item() * 1748, dl.max().item()))
        print(f"{g:55s} max {rows[-1][1]:.3f} px  rms {rows[-1][2]:.4f} px  logits max {rows[-1][3]:.2e}", flush=True)
    return rows


if __name__ == "__main__":
    main()
