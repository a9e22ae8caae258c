"""Host-side mirror of the SA drop's model interface ("Monocular Satellite Pose Estimation Based on Uncertainty
Estimation and Self-Assessment": SA/src/zoo/rtdetr/rtdetr.py, rtdetr_postprocessor.py).

The reference builds ``cfg.model`` / ``cfg.postprocessor`` from a YAML file (SA/src/core/yaml_config.py) and calls
``outputs = model(samples)`` and ``results = postprocessor(outputs, clip_bbox)`` (SA/src/solver/speed_engine.py:153-182).
``build_sa_model(**yaml_values)`` returns objects with those call signatures: ``B200RTDETR`` is an ``nn.Module`` with
the reference's exact ``state_dict`` layout (636 tensors for rtdetr_r50vd_6x_speed_kl_*.yml, ``load_state_dict(strict=
True)`` of a reference checkpoint works unchanged) whose forward runs as one schedule of sm_100a kernels in libspe.so
(``spe_forward_sa``); ``RTDETRPostProcessor`` does softmax, exp(sigma), de-normalisation and -- in the same launch --
the batched pose solve.  Inference only; no CPU / eager fallback.
"""
import math
from types import SimpleNamespace

import torch
from torch import nn

from .engine import Engine
from .models import _Holder, _register

PRESNET_BLOCKS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3)}   # SA/nn/backbone/presnet.py:18-24
STAGE_PLANES = (64, 128, 256, 512)


def sa_param_specs(num_queries=30, dec_layers=3, hidden_dim=256, nheads=8, enc_ff=1024, dec_ff=1024, csp_hidden=128,
                   num_levels=3, num_points=4, num_classes=11, depth=50):
    """(name, shape, kind) for every tensor of the SA reference's state_dict; kind: 'p' parameter, 'b' float buffer,
    'n' the int64 ``num_batches_tracked`` counters of nn.BatchNorm2d (the kl configs set ``freeze_norm: False``)."""
    specs, E = [("temper_param", (1,), "p")], hidden_dim      # RTDETR.temper_param, rtdetr.py:34 (unused in forward)

    def bn(p, c):
        specs.extend([(p + ".weight", (c,), "p"), (p + ".bias", (c,), "p"), (p + ".running_mean", (c,), "b"),
                      (p + ".running_var", (c,), "b"), (p + ".num_batches_tracked", (), "n")])

    def conv_norm(p, cout, cin, k):           # ConvNormLayer, SA/nn/backbone/common.py:8-24
        specs.append((p + ".conv.weight", (cout, cin, k, k), "p"))
        bn(p + ".norm", cout)

    def linear(p, o, i):
        specs.extend([(p + ".weight", (o, i), "p"), (p + ".bias", (o,), "p")])

    def mha(p):
        specs.extend([(p + ".in_proj_weight", (3 * E, E), "p"), (p + ".in_proj_bias", (3 * E,), "p")])
        linear(p + ".out_proj", E, E)

    def ln(p):
        specs.extend([(p + ".weight", (E,), "p"), (p + ".bias", (E,), "p")])

    def mlp(p, dims):
        for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
            linear(f"{p}.layers.{i}", b, a)

    conv_norm("backbone.conv1.conv1_1", 32, 3, 3)
    conv_norm("backbone.conv1.conv1_2", 32, 32, 3)
    conv_norm("backbone.conv1.conv1_3", 64, 32, 3)
    cin = 64
    exp = 4 if depth >= 50 else 1            # BottleNeck (depth 50) / BasicBlock (depth 18, 34), presnet.py:35-123
    for si, (nb, planes) in enumerate(zip(PRESNET_BLOCKS[depth], STAGE_PLANES)):
        for bi in range(nb):
            p = f"backbone.res_layers.{si}.blocks.{bi}"
            if exp == 4:
                conv_norm(p + ".branch2a", planes, cin, 1)
                conv_norm(p + ".branch2b", planes, planes, 3)
                conv_norm(p + ".branch2c", planes * 4, planes, 1)
            else:
                conv_norm(p + ".branch2a", planes, cin, 3)
                conv_norm(p + ".branch2b", planes, planes, 3)
            if bi == 0:
                conv_norm(p + (".short" if si == 0 else ".short.conv"), planes * exp, cin, 1)
                cin = planes * exp
    for i in range(num_levels):
        specs.append((f"decoder.input_proj.{i}.conv.weight", (E, E, 1, 1), "p"))
        bn(f"decoder.input_proj.{i}.norm", E)
    nlp = nheads * num_levels * num_points
    for i in range(dec_layers):
        p = f"decoder.decoder.layers.{i}"
        mha(p + ".self_attn"); ln(p + ".norm1")
        linear(p + ".cross_attn.sampling_offsets", nlp * 2, E); linear(p + ".cross_attn.attention_weights", nlp, E)
        linear(p + ".cross_attn.value_proj", E, E); linear(p + ".cross_attn.output_proj", E, E)
        ln(p + ".norm2"); linear(p + ".linear1", dec_ff, E); linear(p + ".linear2", E, dec_ff); ln(p + ".norm3")
    for i in range(dec_layers):
        mlp(f"decoder.decoder.sigma_embed.{i}", (E, E, E, 1))
    mlp("decoder.query_pos_head", (2, 2 * E, E))
    linear("decoder.enc_output.0", E, E); ln("decoder.enc_output.1")
    linear("decoder.enc_score_head", num_classes + 1, E)
    mlp("decoder.enc_bbox_head", (E, E, E, 2))
    for i in range(dec_layers):
        linear(f"decoder.dec_score_head.{i}", num_classes + 1, E)
    for i in range(dec_layers):
        mlp(f"decoder.dec_bbox_head.{i}", (E, E, E, 2))
    for i, c in enumerate((128 * exp, 256 * exp, 512 * exp)):
        specs.append((f"encoder.input_proj.{i}.0.weight", (E, c, 1, 1), "p"))
        bn(f"encoder.input_proj.{i}.1", E)
    specs.append(("encoder.encoder_fusion_input.weight", (256, 3 * E, 1, 1), "p"))   # defined, unused by forward
    p = "encoder.encoder.0.layers.0"
    mha(p + ".self_attn"); linear(p + ".linear1", enc_ff, E); linear(p + ".linear2", E, enc_ff); ln(p + ".norm1"); ln(p + ".norm2")
    for i in range(num_levels - 1):
        conv_norm(f"encoder.lateral_convs.{i}", E, E, 1)
    for grp in ("fpn_blocks", "pan_blocks"):
        for i in range(num_levels - 1):
            q = f"encoder.{grp}.{i}"
            conv_norm(q + ".conv1", csp_hidden, 2 * E, 1); conv_norm(q + ".conv2", csp_hidden, 2 * E, 1)
            conv_norm(q + ".bottlenecks.0.conv1", csp_hidden, csp_hidden, 3)
            conv_norm(q + ".bottlenecks.0.conv2", csp_hidden, csp_hidden, 1)
            conv_norm(q + ".conv3", E, csp_hidden, 1)
    return specs


class B200RTDETR(nn.Module):
    """Drop-in for the SA drop's ``RTDETR`` in eval mode (SA/src/zoo/rtdetr/rtdetr.py:19-52 with PResNet-50-vd,
    HybridEncoder and RTDETRTransformer as configured by rtdetr_r50vd_6x_speed_kl_*.yml): same ``state_dict`` keys,
    ``forward(x, targets=None)`` -> ``{'pred_logits', 'pred_pts', 'pred_sigmas', 'aux_outputs'}``."""

    def __init__(self, *, input_size=256, num_queries=30, dec_layers=3, hidden_dim=256, nheads=8, enc_ff=1024,
                 dec_ff=1024, expansion=0.5, num_classes=11, max_batch=64, calibrate=False, depth=50):
        super().__init__()
        if depth not in PRESNET_BLOCKS:
            raise ValueError("PResNet depth must be 18, 34 or 50 (rtdetr_r18vd_* / rtdetr_r50vd_* recipes)")
        if hidden_dim != 256 or nheads != 8 or enc_ff != dec_ff or int(hidden_dim * expansion) != 128 or num_classes != 11:
            raise ValueError("libspe.so builds the rtdetr_r50vd speed recipe: hidden_dim 256, 8 heads, equal encoder / "
                             "decoder feed-forward widths, expansion 0.5, 11 keypoint classes")
        self.cfg = SimpleNamespace(input_size=input_size, num_queries=num_queries, dec_layers=dec_layers,
                                   hidden_dim=hidden_dim, nheads=nheads, dim_feedforward=dec_ff)
        self.max_batch, self.calibrate = max_batch, bool(calibrate)
        gen = torch.Generator().manual_seed(0)
        for name, shape, kind in sa_param_specs(num_queries, dec_layers, hidden_dim, nheads, enc_ff, dec_ff,
                                                int(hidden_dim * expansion), depth=depth):
            if kind == "n":
                _register(self, name, torch.zeros((), dtype=torch.int64), True)
            else:
                _register(self, name, self._init_tensor(name, shape, gen), kind == "b")
        self._engine = None
        self._engine_key = None
        self._weights_dirty = True

    @staticmethod
    def _init_tensor(name, shape, gen):
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "running_var" or (leaf == "weight" and len(shape) == 1):
            return torch.ones(shape)
        if len(shape) == 1:
            return torch.zeros(shape)
        if len(shape) == 4:
            return torch.randn(shape, generator=gen) * math.sqrt(2.0 / (shape[1] * shape[2] * shape[3]))
        a = math.sqrt(6.0 / (shape[0] + shape[1]))
        return (torch.rand(shape, generator=gen) * 2 - 1) * a

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._weights_dirty = True
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._weights_dirty = True
        return out

    def refresh_weights(self):
        self._weights_dirty = True

    def train(self, mode=True):
        if mode:
            raise RuntimeError("B200RTDETR is the inference path of the reference; training stays on the reference model")
        return super().train(False)

    def deploy(self):
        """``RTDETR.deploy`` (rtdetr.py:54-61) re-parameterises the RepVgg blocks; the library does that at weight load."""
        return self.eval()

    def _get_engine(self, device):
        if device.type != "cuda":
            raise RuntimeError("B200RTDETR must live on a CUDA (sm_100a) device: call model.to('cuda'). There is no CPU fallback.")
        c = self.cfg
        key = (device.index or 0, c.input_size)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(input_size=c.input_size, num_queries=c.num_queries, enc_layers=1, dec_layers=c.dec_layers,
                                  hidden_dim=c.hidden_dim, nheads=c.nheads, dim_feedforward=c.dim_feedforward,
                                  backbone="rtdetr_r50vd", precision="tf32", has_sigma=True, max_batch=self.max_batch,
                                  device=device.index or 0)
            self._engine_key = key
            self._weights_dirty = True
        if self._weights_dirty:
            self._engine.load_state_dict(self.state_dict())
            self._weights_dirty = False
        return self._engine

    @property
    def engine(self):
        return self._engine

    @torch.no_grad()
    def forward(self, x, targets=None):
        if isinstance(x, (list, tuple)):
            x = torch.stack(list(x))
        R = self.cfg.input_size
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, R, R):
            raise ValueError(f"expected [B,3,{R},{R}] input (eval_spatial_size of the recipe), got {tuple(x.shape)}")
        dev = self.temper_param.device
        x = x.to(device=dev, dtype=torch.float32)
        eng = self._get_engine(dev)
        if self.calibrate and not eng.calibrated:
            eng.calibrate(x)
        outs = [eng.forward_sa(x[i:i + eng.max_batch]) for i in range(0, x.shape[0], eng.max_batch)]
        for o in outs:
            o.pop("topk_ind")
        if len(outs) == 1:
            return outs[0]
        out = {k: torch.cat([o[k] for o in outs]) for k in outs[0] if k != "aux_outputs"}
        out["aux_outputs"] = [{k: torch.cat([o["aux_outputs"][i][k] for o in outs]) for k in a}
                              for i, a in enumerate(outs[0]["aux_outputs"])]
        return out


class RTDETRPostProcessor(nn.Module):
    """Drop-in for the SA drop's ``RTDETRPostProcessor.forward(outputs, clip_bbox)`` (SA/src/zoo/rtdetr/
    rtdetr_postprocessor.py:43-78): per image ``{'logits': softmax probabilities, 'points': pixels, 'sigmas':
    exp(pred_sigmas)}`` as numpy.  The same launch also solves the poses (EPnP-RANSAC-equivalent consensus + sigma-weighted
    refinement, SA/utils/speed_eval.py:332-420); they are kept in ``pose_cache`` for the per-image solver calls."""

    def __init__(self, engine_getter, reproj=25.0, weighted=True, reject=False):
        super().__init__()
        self._engine_getter = engine_getter
        self.reproj, self.weighted, self.reject = reproj, weighted, reject
        self.pose_cache = {}

    @torch.no_grad()
    def forward(self, outputs, clip_bbox):
        logits, points, sig = outputs["pred_logits"], outputs["pred_pts"], outputs["pred_sigmas"]
        assert len(logits) == len(clip_bbox)
        eng = self._engine_getter()
        if eng is None or not logits.is_cuda:
            raise RuntimeError("RTDETRPostProcessor needs the CUDA outputs of a B200RTDETR forward (no CPU fallback)")
        boxes = torch.stack([torch.as_tensor(b) for b in clip_bbox]).to(logits.device)
        r = eng.assign_pnp(logits, points, boxes, log_sigma=sig, reproj=self.reproj, weighted=self.weighted,
                           reject=self.reject, want_post=True)
        probs, pts, sigmas = r["probs"].cpu().numpy(), r["points_px"].cpu().numpy(), r["sigmas"].cpu().numpy()
        quat, tvec, status = r["quat"].cpu().numpy(), r["tvec"].cpu().numpy(), r["status"].cpu().numpy()
        results = []
        self.pose_cache.clear()
        for i in range(len(probs)):
            d = {"logits": probs[i], "points": pts[i], "sigmas": sigmas[i]}
            results.append(d)
            self.pose_cache[id(d["points"])] = (d["points"], quat[i], tvec[i], int(status[i]))
        return results


def build_sa_model(*, input_size=256, num_queries=30, num_decoder_layers=3, hidden_dim=256, nhead=8, dim_feedforward=1024,
                   expansion=0.5, max_batch=64, reproj=25.0, self_assessment=False, calibrate=False, depth=50):
    """What ``cfg.model`` / ``cfg.postprocessor`` give the SA drop's engine, from the values its YAML recipe sets
    (``HybridEncoder`` / ``RTDETRTransformer`` blocks of configs/rtdetr_speed/rtdetr_r50vd_6x_speed_kl_*.yml; ``eval_
    spatial_size`` -> ``input_size``; ``PResNet.depth`` -> ``depth``: 50 for the rtdetr_r50vd recipes, 18 / 34 for rtdetr_r18vd).  Returns ``(model, postprocessor)``; the caller moves the model: model.to('cuda')."""
    model = B200RTDETR(input_size=input_size, num_queries=num_queries, dec_layers=num_decoder_layers, hidden_dim=hidden_dim,
                       nheads=nhead, enc_ff=dim_feedforward, dec_ff=dim_feedforward, expansion=expansion,
                       max_batch=max_batch, calibrate=calibrate, depth=depth)
    model.eval()
    post = RTDETRPostProcessor(lambda: model.engine, reproj=reproj, weighted=True, reject=self_assessment)
    return model, post


def build_sigma_solver(model=None, postprocessor=None, reproj=25.0, self_assessment=False):
    """``build_sigma_solver()`` of the SA drop (SA/utils/speed_eval.py:27-28 -> ``SimplePoseSolverSigma``, :322-420):
    ``solver(points, logits, sigmas) -> (quat, tvec)``, ``IndexError`` when fewer than four keypoints are found or the
    solve fails.  With the ``model`` / ``postprocessor`` of ``build_sa_model`` it answers from the batched solve the
    post-processor already launched."""
    from .solver import BatchedPoseSolver
    return BatchedPoseSolver(reproj=reproj, weighted=True, reject=self_assessment, post=postprocessor, model=model)
