"""Host-side mirror of the reference's model interface (RV/models/__init__.py, RV/models/detr_speed.py).

``build_model(args) -> (model, criterion, postprocessors)`` keeps the reference signature; ``model`` is an
``nn.Module`` with the reference's exact ``state_dict`` layout (so ``load_state_dict(checkpoint['model'],
strict=True)`` works unchanged) whose ``forward`` hands the batch to libspe.so and returns the reference's output
dict.  PyTorch is only the owner of parameters, device memory and streams here; no PyTorch op runs on the hot path
and there is no CPU / eager fallback: a model that is not on a CUDA (sm_100a) device raises.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch
from torch import nn

from .engine import Engine

RESNET50_BLOCKS = (("layer1", 64, 3), ("layer2", 128, 4), ("layer3", 256, 6))


def is_stride8(backbone):
    """RV/models/backbone.py:187-195: anything but resnet18/34/50 builds Backbone8s (ResNet-50, stride-8 neck)."""
    return backbone not in ("resnet18", "resnet34", "resnet50")


def param_specs(backbone, num_queries, enc_layers, dec_layers, hidden_dim, dim_feedforward, sigma_head,
                position_embedding="sine"):
    """(name, shape, is_buffer) for every tensor of the reference state_dict."""
    if backbone in ("resnet18", "resnet34"):
        raise ValueError("only the ResNet-50 backbones of the reference recipes are built (resnet50 / resnet50s8)")
    specs = []
    b = "backbone.0.body"

    def bn(prefix, c):
        for n in ("weight", "bias", "running_mean", "running_var"):   # FrozenBatchNorm2d buffers, backbone.py:29-32
            specs.append((f"{prefix}.{n}", (c,), True))

    specs.append((b + ".conv1.weight", (64, 3, 7, 7), False))
    bn(b + ".bn1", 64)
    inplanes = 64
    for name, planes, nblk in RESNET50_BLOCKS:
        for bi in range(nblk):
            p = f"{b}.{name}.{bi}"
            specs.append((p + ".conv1.weight", (planes, inplanes, 1, 1), False)); bn(p + ".bn1", planes)
            specs.append((p + ".conv2.weight", (planes, planes, 3, 3), False)); bn(p + ".bn2", planes)
            specs.append((p + ".conv3.weight", (planes * 4, planes, 1, 1), False)); bn(p + ".bn3", planes * 4)
            if bi == 0:
                specs.append((p + ".downsample.0.weight", (planes * 4, inplanes, 1, 1), False))
                bn(p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    if is_stride8(backbone):
        specs += [("backbone.0.s8_latern.weight", (256, 512, 1, 1), False),
                  ("backbone.0.s16_latern.weight", (256, 1024, 3, 3), False),
                  ("backbone.0.output_conv.weight", (512, 512, 3, 3), False),
                  ("backbone.0.output_conv.bias", (512,), False)]
        nch = 512
    else:
        nch = 1024
    e, ff = hidden_dim, dim_feedforward

    def linear(p, o, i):
        specs.append((p + ".weight", (o, i), False)); specs.append((p + ".bias", (o,), False))

    def mha(p):
        specs.append((p + ".in_proj_weight", (3 * e, e), False)); specs.append((p + ".in_proj_bias", (3 * e,), False))
        linear(p + ".out_proj", e, e)

    def ln(p):
        specs.append((p + ".weight", (e,), False)); specs.append((p + ".bias", (e,), False))

    for i in range(enc_layers):
        p = f"transformer.encoder.layers.{i}"
        mha(p + ".self_attn"); linear(p + ".linear1", ff, e); linear(p + ".linear2", e, ff)
        ln(p + ".norm1"); ln(p + ".norm2")
    for i in range(dec_layers):
        p = f"transformer.decoder.layers.{i}"
        mha(p + ".self_attn"); mha(p + ".multihead_attn"); linear(p + ".linear1", ff, e); linear(p + ".linear2", e, ff)
        ln(p + ".norm1"); ln(p + ".norm2"); ln(p + ".norm3")
    ln("transformer.decoder.norm")
    linear("cls_embed", 12, e)
    linear("point_embed.layers.0", e, e); linear("point_embed.layers.1", e, e); linear("point_embed.layers.2", 2, e)
    if sigma_head:
        linear("sigma_embed.layers.0", e, e); linear("sigma_embed.layers.1", e, e); linear("sigma_embed.layers.2", 1, e)
    if position_embedding in ("learned", "v3"):     # PositionEmbeddingLearned, RV/models/position_encoding.py:59-63
        specs.append(("backbone.1.row_embed.weight", (50, e // 2), False))
        specs.append(("backbone.1.col_embed.weight", (50, e // 2), False))
    elif position_embedding not in ("sine", "v2"):
        raise ValueError(f"not supported {position_embedding}")
    specs.append(("query_embed.weight", (num_queries, e), False))
    specs.append(("input_proj.weight", (e, nch, 1, 1), False)); specs.append(("input_proj.bias", (e,), False))
    return specs


class _Holder(nn.Module):
    """Parameter container; gives the tree of sub-module names the reference's state_dict keys imply."""

    def forward(self, *a, **k):
        raise NotImplementedError(
            "sub-modules of B200DETR only hold parameters; the forward pass runs as one fused schedule in libspe.so")


def position_embedding_sine(B, H, W, hidden_dim=256, device="cpu"):
    """``PositionEmbeddingSine.forward`` for an all-False mask (RV/models/position_encoding.py:30-53): normalised
    cumulative coordinates x 2 pi / 10000^(2 floor(k/2)/128), sin on even / cos on odd channels, cat(y, x)."""
    npf = hidden_dim // 2
    ones = torch.ones((B, H, W), dtype=torch.float32, device=device)
    y_embed, x_embed = ones.cumsum(1), ones.cumsum(2)
    eps, scale = 1e-6, 2 * math.pi
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(npf, dtype=torch.float32, device=device)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / npf)
    pos_x, pos_y = x_embed[:, :, :, None] / dim_t, y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


class _BackboneView(_Holder):
    """``model.backbone(samples)`` as the reference's ``Joiner`` answers it (RV/models/backbone.py:156-165; called
    directly by RV/get_backbone_time.py:110): ``([NestedTensor-like features], [position embedding])``.  The features are
    read back from the fused schedule's activation taps -- a compatibility / inspection accessor, not the hot path."""

    def forward(self, samples):
        root = self.__dict__["_root"]()
        x = root._as_batch(samples)
        dev = root.query_embed.weight.device
        x = x.to(device=dev, dtype=torch.float32)
        B, R = x.shape[0], x.shape[-1]
        eng = root._get_engine(dev, R, B)
        if B > eng.max_batch:
            raise ValueError(f"backbone view: batch {B} exceeds max_batch {eng.max_batch}")
        eng.enable_taps(True)
        try:
            eng.forward(x)
            torch.cuda.synchronize(dev)
            if is_stride8(root.cfg.backbone):
                h = R // 8
                feat = eng.read_tap("neck", (B, h, h, 512))
            else:
                h = R // 16
                feat = eng.read_tap("layer3", (B, h, h, 1024))
        finally:
            eng.enable_taps(False)
        feat = feat.permute(0, 3, 1, 2).contiguous().to(dev)
        mask = torch.zeros((B, h, h), dtype=torch.bool, device=dev)
        if root.cfg.position_embedding in ("learned", "v3"):       # RV/models/position_encoding.py:69-81
            tables = self._modules["1"]
            x_emb, y_emb = tables.col_embed.weight[:h], tables.row_embed.weight[:h]
            pos = torch.cat([x_emb.unsqueeze(0).repeat(h, 1, 1), y_emb.unsqueeze(1).repeat(1, h, 1)], dim=-1)
            pos = pos.permute(2, 0, 1).unsqueeze(0).repeat(B, 1, 1, 1)
        else:
            pos = position_embedding_sine(B, h, h, root.cfg.hidden_dim, device=dev)
        return [SimpleNamespace(tensors=feat, mask=mask, decompose=lambda: (feat, mask))], [pos]


def _register(root, name, tensor, is_buffer):
    parts = name.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Holder())
        mod = mod._modules[p]
    if is_buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class B200DETR(nn.Module):
    """Drop-in for ``DETR`` (RV/models/detr_speed.py:32-100): same constructor-level configuration, same
    ``state_dict`` keys, same ``forward(samples)`` contract and output dict; inference only."""

    def __init__(self, *, backbone="resnet50s8", num_queries=40, enc_layers=4, dec_layers=4, hidden_dim=256, nheads=8,
                 dim_feedforward=2048, aux_loss=True, input_size=None, precision="tf32", sigma_head=False,
                 max_batch=64, calibrate=True, position_embedding="sine"):
        super().__init__()
        if hidden_dim != 256 or nheads != 8:
            raise ValueError("libspe.so is built for hidden_dim=256, nheads=8 (every recipe of the reference)")
        self.num_queries, self.aux_loss = num_queries, aux_loss
        self.cfg = SimpleNamespace(backbone=backbone, num_queries=num_queries, enc_layers=enc_layers,
                                   dec_layers=dec_layers, hidden_dim=hidden_dim, nheads=nheads,
                                   dim_feedforward=dim_feedforward, sigma_head=sigma_head,
                                   position_embedding=position_embedding)
        self.input_size, self.precision, self.max_batch = input_size, precision, max_batch
        # fold the mean effect of the TF32 / BF16 weight rounding into the biases, measured on the first batch this
        # model sees after (re)loading weights (Engine.calibrate / spe_calibrate)
        self.calibrate = bool(calibrate)
        gen = torch.Generator().manual_seed(0)
        for name, shape, is_buffer in param_specs(backbone, num_queries, enc_layers, dec_layers, hidden_dim,
                                                  dim_feedforward, sigma_head, position_embedding):
            _register(self, name, self._init_tensor(name, shape, gen), is_buffer)
        self._engine = None
        self._engine_key = None
        self._weights_dirty = True
        import weakref
        self._modules["backbone"].__class__ = _BackboneView          # same parameters, plus the Joiner-style call
        self._modules["backbone"].__dict__["_root"] = weakref.ref(self)

    @staticmethod
    def _init_tensor(name, shape, gen):
        """Random init in the spirit of the reference (torchvision kaiming for convs, identity FrozenBN,
        xavier_uniform for the transformer, RV/models/transformer.py:46-49); real use loads a checkpoint."""
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "running_var" or (leaf == "weight" and len(shape) == 1):
            return torch.ones(shape)
        if len(shape) == 1:
            return torch.zeros(shape)
        if len(shape) == 4:
            fan_out = shape[0] * shape[2] * shape[3]
            return torch.randn(shape, generator=gen) * math.sqrt(2.0 / fan_out)
        if name == "query_embed.weight":
            return torch.randn(shape, generator=gen)
        if name.endswith("_embed.weight"):                       # PositionEmbeddingLearned: nn.init.uniform_
            return torch.rand(shape, generator=gen)
        a = math.sqrt(6.0 / (shape[0] + shape[1]))
        return (torch.rand(shape, generator=gen) * 2 - 1) * a

    # ---- weight synchronisation with the device-side repacked copy ----------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        state_dict = {k: v for k, v in state_dict.items() if not k.endswith("num_batches_tracked")}
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._weights_dirty = True
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._weights_dirty = True
        return out

    def refresh_weights(self):
        """Call after modifying parameters in place: the kernels read a repacked device copy."""
        self._weights_dirty = True

    def train(self, mode=True):
        if mode:
            raise RuntimeError("B200DETR is the inference path of the reference; training stays on the reference model")
        return super().train(False)

    def _get_engine(self, device, R, B):
        if device.type != "cuda":
            raise RuntimeError("B200DETR must live on a CUDA (sm_100a) device: call model.to('cuda'). "
                               "There is no CPU fallback.")
        c = self.cfg
        key = (device.index or 0, R, self.precision)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(input_size=R, num_queries=c.num_queries, enc_layers=c.enc_layers,
                                  dec_layers=c.dec_layers, hidden_dim=c.hidden_dim, nheads=c.nheads,
                                  dim_feedforward=c.dim_feedforward, backbone=c.backbone, precision=self.precision,
                                  has_sigma=c.sigma_head, max_batch=self.max_batch, device=device.index or 0)
            self._engine_key = key
            self._weights_dirty = True
        if self._weights_dirty:
            self._engine.load_state_dict(self.state_dict())
            self._weights_dirty = False
        return self._engine

    @property
    def engine(self):
        return self._engine

    # ---- forward -------------------------------------------------------------------------------------------------
    @staticmethod
    def _as_batch(samples):
        """NestedTensor | list[Tensor] | Tensor -> Tensor [B,3,R,R]   (RV/models/detr_speed.py:76-77,
        RV/utils/misc.py:310-333).  Ragged batches (non-trivial padding mask) are not built: the test-time crop
        always produces equal R x R inputs."""
        if hasattr(samples, "decompose"):
            tensors, mask = samples.decompose()
            if mask is not None and bool(mask.any()):
                raise ValueError("padded (ragged) batches are not supported by the B200 path")
            return tensors
        if isinstance(samples, (list, tuple)):
            shapes = {tuple(t.shape) for t in samples}
            if len(shapes) != 1:
                raise ValueError("all images of a batch must have the same size")
            return torch.stack(list(samples))
        return samples

    @torch.no_grad()
    def forward(self, samples):
        x = self._as_batch(samples)
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise ValueError(f"expected [B,3,R,R] input, got {tuple(x.shape)}")
        dev = self.query_embed.weight.device
        x = x.to(device=dev, dtype=torch.float32)
        R = x.shape[-1]
        if self.input_size is not None and R != self.input_size:
            raise ValueError(f"model was built for input_size={self.input_size}, got {R}")
        eng = self._get_engine(dev, R, x.shape[0])
        if self.calibrate and not eng.calibrated:
            eng.calibrate(x)

        def run(chunk):   # engine outputs live in persistent buffers: detach them from the next call
            o = eng.forward(chunk, want_aux=self.aux_loss)
            res = {k: v.clone() for k, v in o.items() if k != "aux_outputs"}
            if "aux_outputs" in o:
                res["aux_outputs"] = [{k: v.clone() for k, v in a.items()} for a in o["aux_outputs"]]
            return res

        outs = [run(x[i:i + eng.max_batch]) for i in range(0, x.shape[0], eng.max_batch)]
        if len(outs) == 1:
            out = outs[0]
        else:
            out = {k: torch.cat([o[k] for o in outs]) for k in outs[0] if k != "aux_outputs"}
            if "aux_outputs" in outs[0]:
                out["aux_outputs"] = [{k: torch.cat([o["aux_outputs"][i][k] for o in outs]) for k in a}
                                      for i, a in enumerate(outs[0]["aux_outputs"])]
        return out


class PostProcess(nn.Module):
    """Drop-in for ``PostProcess`` (RV/models/detr_speed.py:264-293; with sigmas:
    SA/src/zoo/rtdetr/rtdetr_postprocessor.py:43-78).  One kernel launch does softmax, the pixel de-normalisation,
    the query->keypoint assignment and the PnP solve for the whole batch; the per-image poses are remembered so that
    the reference's per-image ``solver(points, logits)`` calls that follow are answered without extra GPU work."""

    def __init__(self, engine_getter=None, reproj=20.0, weighted=False, reject=False):
        super().__init__()
        self._engine_getter = engine_getter
        self.reproj, self.weighted, self.reject = reproj, weighted, reject
        self.pose_cache = {}

    @torch.no_grad()
    def forward(self, outputs, clip_bbox):
        logits, points = outputs["pred_logits"], outputs["pred_points"]
        assert len(logits) == len(clip_bbox)
        eng = self._engine_getter() if self._engine_getter else None
        if eng is None or not logits.is_cuda:
            raise RuntimeError("PostProcess needs the CUDA outputs of a B200DETR forward (no CPU fallback)")
        boxes = torch.stack([torch.as_tensor(b) for b in clip_bbox]).to(logits.device)
        sig = outputs.get("pred_sigmas")
        r = eng.assign_pnp(logits, points, boxes, log_sigma=sig, reproj=self.reproj,
                           weighted=self.weighted and sig is not None, reject=self.reject, want_post=True)
        probs, pts = r["probs"].cpu().numpy(), r["points_px"].cpu().numpy()
        quat, tvec, status = r["quat"].cpu().numpy(), r["tvec"].cpu().numpy(), r["status"].cpu().numpy()
        sigmas = r["sigmas"].cpu().numpy() if "sigmas" in r else None
        results = []
        self.pose_cache.clear()
        for i in range(len(probs)):
            d = {"logits": probs[i], "points": pts[i]}
            if sigmas is not None:
                d["sigmas"] = sigmas[i]
            results.append(d)
            self.pose_cache[id(d["points"])] = (d["points"], quat[i], tvec[i], int(status[i]))
        return results


class NoCriterion(nn.Module):
    """Stand-in for ``SetCriterion`` (RV/models/detr_speed.py:103-261) on the inference path.  ``evaluate`` calls
    ``criterion(outputs, targets)`` and reads ``criterion.weight_dict`` purely to log validation losses
    (RV/engine.py:99-113); the training loss (Hungarian matching + set loss) is outside this path, so this returns
    no weighted losses and ``class_error`` = NaN ("not computed" -- ``evaluate`` reads that key unconditionally), which
    lets ``main.py --eval`` run unedited and log nothing misleading."""

    def __init__(self):
        super().__init__()
        self.weight_dict = {}

    def forward(self, outputs, targets):
        dev = outputs["pred_logits"].device if isinstance(outputs, dict) and "pred_logits" in outputs else "cpu"
        return {"class_error": torch.full((), float("nan"), device=dev)}


def build_model(args):
    """Same contract as the reference's ``build_model(args)`` (RV/models/__init__.py:5-6 ->
    RV/models/detr_speed.py:296-336): returns ``(model, criterion, postprocessors)``.  ``criterion`` (training loss)
    is outside this path: a ``NoCriterion`` that yields no losses keeps ``engine.evaluate`` running unedited.  Additive optional attributes on ``args``:
    ``precision`` ('tf32' | 'bf16'), ``sigma_head`` (bool), ``max_batch`` (int), ``input_size``, ``calibrate`` (bool,
    default True: rounding-bias calibration on the first batch, see ``Engine.calibrate``)."""
    if getattr(args, "pre_norm", False):
        # RV/models/transformer.py:284-294 builds normalize_before layers + an encoder norm for --pre_norm; no recipe of the
        # reference uses it and the fused schedule is the post-norm one: refuse instead of computing something else
        raise NotImplementedError("--pre_norm is not built: every recipe of the reference trains the post-norm transformer")
    model = B200DETR(
        backbone=args.backbone, num_queries=args.num_queries, enc_layers=args.enc_layers, dec_layers=args.dec_layers,
        hidden_dim=args.hidden_dim, nheads=args.nheads, dim_feedforward=args.dim_feedforward,
        aux_loss=getattr(args, "aux_loss", True), input_size=getattr(args, "input_size", None),
        precision=getattr(args, "precision", "tf32"), sigma_head=getattr(args, "sigma_head", False),
        calibrate=getattr(args, "calibrate", True), position_embedding=getattr(args, "position_embedding", "sine"),
        max_batch=getattr(args, "max_batch", None) or max(int(getattr(args, "batch_size", 64) or 64), 1))
    if is_stride8(args.backbone):
        args.backbone = "resnet50"   # the reference's build_backbone rewrites it too (RV/models/backbone.py:193)
    post = PostProcess(engine_getter=lambda: model.engine, reproj=float(getattr(args, "repro", 20)),
                       weighted=bool(getattr(args, "sigma_head", False)),
                       reject=bool(getattr(args, "self_assessment", False)))
    model.eval()   # like the reference, the caller moves it: model.to(device)
    return model, NoCriterion(), {"points": post}
