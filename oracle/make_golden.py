"""Writes tests/golden/* by running the REAL reference (build container only; TEST INFRASTRUCTURE).

    python -m oracle.make_golden

The reference ships no golden vectors for this path, so they are produced here by executing the
reference's own classes from /root/reference -- the DETR model, ``SpeedSubmission.__getitem__``, ``PostProcess`` and
``SimplePoseSolver`` -- on the seeded synthetic inputs of oracle/synth.py.  Missing third-party modules are replaced
by minimal in-process stand-ins (no reference file is edited):
  albumentations  ``Compose([Resize(h, w, interp)])(image=img)`` -> ``cv2.resize(img, (w, h), interpolation=interp)``,
                  which is what albumentations 0.5.1 ``Resize`` does (pinned in RV/requirements.txt:5)
  mathutils       ``Matrix(R).to_quaternion()`` -> rotation matrix to (w, x, y, z), w >= 0
  matplotlib      empty module (only imported for plotting helpers)
Fixtures written:
  wz_synt_test_boxes.npy / wz_real_test_boxes.npy   detector boxes shipped with the reference (RV/annos/*.json)
  crop_golden.npz     reference crop boxes + uint8 resized crops for a spread of boxes
  model_golden.npz    reference model outputs for seeded weights/inputs (weights are regenerated, never stored)
  pnp_golden.npz      reference SimplePoseSolver poses for seeded synthetic predictions
  deform_attn_golden.npz  the SA drop's deformable_attention_core_func and live MSDeformableAttention module on seeded inputs
  sa_model_golden.npz  the SA drop's live RT-DETR model (PResNet-50-vd + HybridEncoder + RTDETRTransformer) on seeded weights
  model_b256_golden.npz  reference model (calibrated heads, oracle/make_chain_fixture.py) + PostProcess + solver on the
                      256 crops of the benchmarked frame sets; sigma head through the SA drop's own MLP class
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

from . import crop_ref, model_ref, pnp_ref, ref_import, synth
from .model_ref import ModelCfg

GOLDEN = synth.GOLDEN_DIR

CROP_INDICES = [0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144]   # + largest / smallest / most out-of-frame boxes
MODEL_CASES = {
    # name: (cfg kwargs, batch, input size, seed)
    "s8_q40": (dict(backbone="resnet50s8", num_queries=40, enc_layers=4, dec_layers=4), 2, 224, 0),
    "s16_q100": (dict(backbone="resnet50", num_queries=100, enc_layers=6, dec_layers=6), 1, 256, 1),
}


def model_inputs(B, R, seed):
    rng = np.random.default_rng(1000 + seed)
    return torch.from_numpy(rng.standard_normal((B, 3, R, R)).astype(np.float32))


def _install_stubs():
    import cv2
    if "albumentations" not in sys.modules:
        A = types.ModuleType("albumentations")

        class Resize:
            def __init__(self, height, width, interpolation=cv2.INTER_LINEAR, **kw):
                self.h, self.w, self.interp = height, width, interpolation

            def __call__(self, image):
                return cv2.resize(image, (self.w, self.h), interpolation=self.interp)

        class Compose:
            def __init__(self, transforms, **kw):
                self.transforms = transforms

            def __call__(self, image, **kw):
                for t in self.transforms:
                    image = t(image)
                return {"image": image}

        def _unsupported(*a, **k):
            return None

        def _getattr(name):          # training-only augmentations are never called on this path
            if name.startswith("__"):
                raise AttributeError(name)
            return _unsupported
        A.Resize, A.Compose = Resize, Compose
        A.__getattr__ = _getattr
        A.__file__ = "<oracle stub>"
        sys.modules["albumentations"] = A
    if "mathutils" not in sys.modules:
        M = types.ModuleType("mathutils")

        class Matrix:
            def __init__(self, m):
                self.m = np.asarray(m, dtype=np.float64)

            def to_quaternion(self):
                return pnp_ref.rot_to_quat(self.m)
        M.Matrix = Matrix
        M.Quaternion = object
        M.__file__ = "<oracle stub>"
        sys.modules["mathutils"] = M
    if "matplotlib" not in sys.modules:
        mp = types.ModuleType("matplotlib")
        mp.pyplot = types.ModuleType("matplotlib.pyplot")
        mp.__file__ = mp.pyplot.__file__ = "<oracle stub>"
        sys.modules["matplotlib"] = mp
        sys.modules["matplotlib.pyplot"] = mp.pyplot


def import_rv_dataset_and_solver():
    _install_stubs()
    ref_import.import_rv_models()
    with ref_import._rv_on_path():
        import datasets.speed as rv_speed
        import utils.speed_eval as rv_eval
    rv_eval.world_pt_path = os.path.join(ref_import.RV_ROOT, "all_result.json")
    return rv_speed, rv_eval


def write_boxes():
    for name in ("wz_synt_test", "wz_real_test"):
        with open(os.path.join(ref_import.RV_ROOT, "annos", name + ".json")) as f:
            anns = json.load(f)
        boxes = np.asarray([v[0][:4] for v in anns.values()], dtype=np.float64)
        np.save(os.path.join(GOLDEN, name + "_boxes.npy"), boxes)
        print(name, boxes.shape)


def crop_case_boxes():
    boxes = synth.load_detector_boxes()
    sides = np.maximum(boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1])
    clip = np.stack([crop_ref.generate_clip_bbox(b) for b in boxes])
    outside = np.maximum(0, -clip[:, 0]) + np.maximum(0, -clip[:, 1]) + np.maximum(0, clip[:, 2] - 1920) + \
        np.maximum(0, clip[:, 3] - 1200)
    idx = CROP_INDICES + [int(np.argmax(sides)), int(np.argmin(sides))] + [int(i) for i in np.argsort(-outside)[:4]]
    return np.asarray(idx), boxes[idx]


def write_crop(rv_speed):
    from PIL import Image
    idx, det = crop_case_boxes()
    frames = synth.make_frames(len(idx), det, seed=0)
    R = 224
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "annos")); os.makedirs(os.path.join(tmp, "images"))
        anns = {}
        for i in range(len(idx)):
            fn = f"img{i:06d}.png"
            Image.fromarray(frames[i]).save(os.path.join(tmp, "images", fn))
            anns[fn] = [list(det[i]) + [1.0]]
        with open(os.path.join(tmp, "annos", "a.json"), "w") as f:
            json.dump(anns, f)
        rv_speed.DATA_ROOT = tmp
        ds = rv_speed.SpeedSubmission("a.json", "images", R)       # the reference's own dataset class
        u8s, clips = [], []
        for i in range(len(ds)):
            img, target = ds[i]
            clip = target["clip_bbox"].numpy()
            u8 = crop_ref.crop_resize_u8(frames[i], clip, R)
            # the reference tensor must equal normalise(oracle uint8 crop) bit for bit
            assert torch.equal(img, crop_ref.normalize_u8(u8)), f"oracle crop != reference for case {i}"
            assert (clip == crop_ref.generate_clip_bbox(det[i])).all()
            u8s.append(u8[:, :, 0]); clips.append(clip)
    np.savez_compressed(os.path.join(GOLDEN, "crop_golden.npz"), box_index=idx, det_boxes=det,
                        clip_boxes=np.stack(clips), crops_u8=np.stack(u8s), input_size=R, frame_seed=0)
    print("crop cases", len(idx), "sides", [int(c[2] - c[0]) for c in clips])


def write_crop_eval(rv_speed):
    """main.py --eval crop: the reference's own SpeedTrain(train=False).__getitem__ (RV/datasets/speed.py:209-244) with
    its random ``img_trunc`` switched off, against the oracle restatement; uint8 crops + boxes are committed."""
    from PIL import Image
    idx, det = crop_case_boxes()
    idx, det = idx[:10], det[:10]
    frames = synth.make_frames(len(idx), det, seed=0)
    R = 224
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "annos")); os.makedirs(os.path.join(tmp, "images"))
        anns = []
        for i in range(len(idx)):
            fn = f"img{i:06d}.png"
            Image.fromarray(frames[i]).save(os.path.join(tmp, "images", fn))
            anns.append({"filename": fn, "landmarks": np.zeros((11, 3)).tolist(), "bbox_xxyy": list(det[i])})
        with open(os.path.join(tmp, "annos", "a.json"), "w") as f:
            json.dump(anns, f)
        np.savetxt(os.path.join(tmp, "annos", "idx.txt"), np.arange(len(idx)), fmt="%d")
        rv_speed.DATA_ROOT = tmp
        rv_speed.img_trunc = lambda img, **kw: img            # np.random-driven blackout, applied even in eval (:230)

        class T:                                              # make_transforms(False): A.Resize only (keypoints pass through)
            def __call__(self, image, keypoints):
                import cv2
                return {"image": cv2.resize(image, (R, R), interpolation=cv2.INTER_CUBIC), "keypoints": keypoints}
        ds = rv_speed.SpeedTrain("a.json", "idx.txt", "images", resize=R, train=False, transforms=T())
        u8s, fboxes = [], []
        for i in range(len(ds)):
            img, target = ds[i]
            fbox = target["clip_bbox"].numpy()
            want = crop_ref.generate_clip_bbox_val(det[i], (1920, 1200))
            assert np.array_equal(fbox, want), (fbox, want)
            u8 = crop_ref.eval_crop_resize_u8(frames[i], fbox, R)
            assert torch.equal(img, crop_ref.normalize_u8(u8)), f"oracle eval crop != reference for case {i}"
            u8s.append(u8[:, :, 0]); fboxes.append(fbox)
    np.savez_compressed(os.path.join(GOLDEN, "crop_eval_golden.npz"), box_index=idx, det_boxes=det,
                        float_boxes=np.stack(fboxes), crops_u8=np.stack(u8s), input_size=R, frame_seed=0)
    print("eval crop cases", len(idx), "sizes", [(int(round(b[2]) - round(b[0])), int(round(b[3]) - round(b[1]))) for b in fboxes])


def write_model():
    out = {}
    for name, (kw, B, R, seed) in MODEL_CASES.items():
        cfg = ModelCfg(**kw)
        sd = synth.make_state_dict(cfg, seed=seed)
        model, _, _ = ref_import.build_reference_model(cfg, sd)
        x = model_inputs(B, R, seed)
        with torch.no_grad():
            ref = model(x)
        port = model_ref.forward(sd, cfg, x)
        err = max((ref[k] - port[k]).abs().max().item() for k in ("pred_logits", "pred_points"))
        print(name, "restatement vs live reference max|d|", err)
        assert err < 1e-4
        out[name + "/checksum"] = np.frombuffer(synth.weights_checksum(sd).encode(), dtype=np.uint8)
        out[name + "/pred_logits"] = ref["pred_logits"].numpy()
        out[name + "/pred_points"] = ref["pred_points"].numpy()
        out[name + "/aux_logits"] = torch.stack([a["pred_logits"] for a in ref["aux_outputs"]]).numpy()
        out[name + "/aux_points"] = torch.stack([a["pred_points"] for a in ref["aux_outputs"]]).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "model_golden.npz"), **out)


B256_SETS = 4                     # synth.bench_set(0..3): the frames bench.py times (sets 0, 1) + two more


def b256_inputs():
    """The 256 crops of the batch-64 / batch-256 parity goldens: oracle crops (cv2) of synth.bench_set(0..3)."""
    crops, clips = [], []
    for s in range(B256_SETS):
        frames, det = synth.bench_set(s)
        for i in range(len(frames)):
            t, c = crop_ref.crop_resize_normalize(frames[i], det[i], 224)
            crops.append(t); clips.append(c)
    return torch.stack(crops), np.stack(clips)


def write_model_b256(rv_eval):
    """The benchmarked configurations end to end: the LIVE reference model (calibrated heads, so its queries emit 11
    distinct labels) on 256 crops, the reference's own PostProcess + SimplePoseSolver behind it, and the sigma head
    evaluated by the SA drop's own ``MLP`` class (SA/src/zoo/rtdetr/rtdetr_decoder.py:24-37, :295-297, :367)."""
    import cv2
    cfg = ModelCfg(sigma_head=True)
    sd = synth.make_state_dict(cfg, seed=0, spread_labels=True)
    rv_sd = {k: v for k, v in sd.items() if not k.startswith("sigma_embed.")}
    model, _, _ = ref_import.build_reference_model(ModelCfg(), rv_sd)
    hs_store = {}
    model.transformer.register_forward_hook(lambda m, i, o: hs_store.__setitem__("hs", o[0]))
    x, clips = b256_inputs()
    logits, points, hs_last = [], [], []
    with torch.no_grad():
        for i in range(0, len(x), 32):
            ref = model(x[i:i + 32])
            logits.append(ref["pred_logits"]); points.append(ref["pred_points"]); hs_last.append(hs_store["hs"][-1])
    logits, points, hs_last = torch.cat(logits), torch.cat(points), torch.cat(hs_last)
    port = model_ref.forward(sd, cfg, x[:16])
    err = max((port["pred_logits"] - logits[:16]).abs().max().item(), (port["pred_points"] - points[:16]).abs().max().item())
    print("b256: restatement vs live reference max|d|", err)
    assert err < 1e-4
    # sigma head through the SA drop's own MLP module
    sa = ref_import.import_sa_rtdetr()
    from src.zoo.rtdetr.rtdetr_decoder import MLP as SA_MLP
    mlp = SA_MLP(256, 256, 1, num_layers=3)
    mlp.load_state_dict({k[len("sigma_embed."):]: v for k, v in sd.items() if k.startswith("sigma_embed.")}, strict=True)
    with torch.no_grad():
        sig = mlp(hs_last).repeat(1, 1, 2)                                   # rtdetr_decoder.py:367
    assert (sig[:16] - port["pred_sigmas"]).abs().max().item() < 1e-5
    # the reference's own post-processing + solver on the reference's own network outputs
    post = sys.modules["models"].PostProcess()
    res = post({"pred_logits": logits, "pred_points": points.clone()}, [torch.from_numpy(c) for c in clips])
    solver = rv_eval.SimplePoseSolver(synth.reference_args(ModelCfg()))
    n = len(res)
    quat = np.zeros((n, 4)); tvec = np.zeros((n, 3)); ok = np.zeros(n, dtype=np.int32)
    for i, r in enumerate(res):
        try:
            q, t = solver(r["points"], r["logits"])
            quat[i], tvec[i], ok[i] = q, t, 1
        except (IndexError, cv2.error):
            pass
    assign = np.stack([pnp_ref.assign_table(r["points"], r["logits"]) for r in res])
    np.savez_compressed(os.path.join(GOLDEN, "model_b256_golden.npz"),
                        checksum=np.frombuffer(synth.weights_checksum(sd).encode(), dtype=np.uint8),
                        pred_logits=logits.numpy(), pred_points=points.numpy(), pred_sigmas=sig.numpy(),
                        clip_boxes=clips, quat=quat, tvec=tvec, ok=ok, assign=assign)
    print("b256: reference chain solved", int(ok.sum()), "of", n, "; distinct labels per image",
          np.unique((assign >= 0).sum(1)))


def write_model_b64_random():
    """BASELINE configs[1] as specified (random-init weights, batch 64): the live reference on the 64 crops of frame
    set 0.  Every query collapses to one label with these weights, so only the network outputs are pinned here."""
    cfg = ModelCfg()
    sd = synth.make_state_dict(cfg, seed=0)
    model, _, _ = ref_import.build_reference_model(cfg, sd)
    x, _ = b256_inputs()
    x = x[:64]
    with torch.no_grad():
        ref = [model(x[i:i + 32]) for i in range(0, 64, 32)]
    np.savez_compressed(os.path.join(GOLDEN, "model_b64_random_golden.npz"),
                        checksum=np.frombuffer(synth.weights_checksum(sd).encode(), dtype=np.uint8),
                        pred_logits=torch.cat([r["pred_logits"] for r in ref]).numpy(),
                        pred_points=torch.cat([r["pred_points"] for r in ref]).numpy())
    print("b64 random-init golden written")


DEFORM_CASE = dict(bs=3, Lq=30, heads=8, levels=((32, 32), (16, 16), (8, 8)), points=4, seed=41)


def deform_inputs(case=DEFORM_CASE):
    """Seeded inputs + module weights of the deformable-attention goldens (numpy PCG64: identical on every machine)."""
    rng = np.random.default_rng(case["seed"])
    bs, Lq, H, P = case["bs"], case["Lq"], case["heads"], case["points"]
    L = len(case["levels"])
    Lv = sum(h * w for h, w in case["levels"])
    f = lambda *shape, scale=1.0: torch.from_numpy((rng.standard_normal(shape) * scale).astype(np.float32))
    d = {"query": f(bs, Lq, 256), "value_in": f(bs, Lv, 256),
         # reference points as the decoder passes them: one (x, y) in (0, 1) per query, shared by the levels
         # (rtdetr_decoder.py:320 `ref_points_detach.unsqueeze(2)`); a few outside [0, 1] exercise the zero padding
         "ref": torch.from_numpy(rng.uniform(-0.1, 1.1, (bs, Lq, 1, 2)).astype(np.float32)),
         "weights": {
             "sampling_offsets.weight": f(H * L * P * 2, 256, scale=0.05), "sampling_offsets.bias": f(H * L * P * 2, scale=2.0),
             "attention_weights.weight": f(H * L * P, 256, scale=0.1), "attention_weights.bias": f(H * L * P, scale=0.5),
             "value_proj.weight": f(256, 256, scale=0.06), "value_proj.bias": f(256, scale=0.02),
             "output_proj.weight": f(256, 256, scale=0.06), "output_proj.bias": f(256, scale=0.02)},
         # direct inputs of the core function: sampling locations spilling over the map border, softmaxed weights
         "core_value": f(bs, Lv, H, 32), "core_loc": torch.from_numpy(rng.uniform(-0.15, 1.15, (bs, Lq, H, L, P, 2)).astype(np.float32)),
         "core_attn": torch.softmax(f(bs, Lq, H, L * P), -1).reshape(bs, Lq, H, L, P)}
    return d


def write_deform_attn():
    """SA/src/zoo/rtdetr/utils.py:15-64 (``deformable_attention_core_func``) and the LIVE ``MSDeformableAttention`` module
    (rtdetr_decoder.py:40-191) on seeded inputs: the core output the module hands to ``output_proj`` is captured with a
    forward-pre hook, so the softmax / sampling-location arithmetic in between is the reference's own."""
    ref_import.import_sa_rtdetr()
    from src.zoo.rtdetr.rtdetr_decoder import MSDeformableAttention
    from src.zoo.rtdetr.utils import deformable_attention_core_func
    case = DEFORM_CASE
    d = deform_inputs(case)
    shapes = [list(s_) for s_ in case["levels"]]
    with torch.no_grad():
        core = deformable_attention_core_func(d["core_value"], shapes, d["core_loc"], d["core_attn"])
        m = MSDeformableAttention(256, case["heads"], len(shapes), case["points"])
        m.load_state_dict(d["weights"], strict=True)
        m.eval()
        got = {}
        m.output_proj.register_forward_pre_hook(lambda mod, inp: got.__setitem__("core", inp[0].clone()))
        out = m(d["query"], d["ref"], d["value_in"], shapes)
    np.savez_compressed(os.path.join(GOLDEN, "deform_attn_golden.npz"), core_func_out=core.numpy(),
                        module_core_out=got["core"].numpy(), module_out=out.numpy())
    print("deformable attention goldens:", core.shape, got["core"].shape, out.shape)


SA_MODEL_CASE = dict(batch=4, seed=7, weights_seed=0)
SA_R18_CASE = dict(batch=3, seed=8, weights_seed=1, depth=18)


def write_sa_model(case=None, name="sa_model_golden.npz", yml=None):
    """The LIVE SA ``RTDETR`` model (ref_import.build_sa_reference_model) on seeded weights / inputs -> sa_model_golden.npz:
    the eval-mode output dict (SA/src/zoo/rtdetr/rtdetr_decoder.py:732-751), the anchor scores and the top-k selection
    (recomputed with the reference's own ``torch.topk`` call, :646-648, on the live ``enc_score_head`` output)."""
    from . import sa_model_ref
    case = case or SA_MODEL_CASE
    cfg = sa_model_ref.SaCfg(depth=case.get("depth", 50))
    sd = synth.make_sa_state_dict(cfg, seed=case["weights_seed"])
    model = ref_import.build_sa_reference_model(sd, yml)
    x = model_inputs(case["batch"], cfg.input_size, case["seed"])
    got = {}
    model.decoder.enc_score_head.register_forward_hook(lambda mod, inp, out: got.__setitem__("cls", out.detach().clone()))
    with torch.no_grad():
        out = model(x)
    scores = got["cls"].max(-1).values
    topk = torch.topk(scores, cfg.num_queries, dim=1)[1]
    taps = {}
    mine = sa_model_ref.forward(sd, cfg, x, taps)
    assert torch.equal(taps["topk"], topk), "restatement selects other anchors than the live model"
    for k in ("pred_logits", "pred_pts", "pred_sigmas"):
        d = (mine[k] - out[k]).abs().max().item()
        assert d < 5e-5, (k, d)
    aux = out["aux_outputs"]
    assert len(aux) == cfg.dec_layers
    np.savez_compressed(
        os.path.join(GOLDEN, name), weights_sha256=synth.weights_checksum(sd),
        pred_logits=out["pred_logits"].numpy(), pred_pts=out["pred_pts"].numpy(), pred_sigmas=out["pred_sigmas"].numpy(),
        aux_logits=torch.stack([a["pred_logits"] for a in aux]).numpy(), aux_pts=torch.stack([a["pred_pts"] for a in aux]).numpy(),
        aux_sigmas=torch.stack([a["pred_sigmas"] for a in aux[:-1]]).numpy(), enc_scores=scores.numpy(),
        topk=topk.numpy().astype(np.int32))
    print("SA model golden written:", tuple(out["pred_logits"].shape), "restatement within 5e-5 of the live model")


def write_pnp(rv_eval, n=400):
    import cv2
    d = synth.make_predictions(n, seed=1)
    solver = rv_eval.SimplePoseSolver(synth.reference_args(ModelCfg()))     # the reference's own solver class
    models = sys.modules["models"]
    post = models.PostProcess()                                            # the reference's own PostProcess
    res = post({"pred_logits": torch.from_numpy(d["logits"]), "pred_points": torch.from_numpy(d["points"]).clone()},
               [torch.from_numpy(b) for b in d["boxes"]])
    quat = np.zeros((n, 4)); tvec = np.zeros((n, 3)); ok = np.zeros(n, dtype=np.int32)
    probs = np.stack([r["logits"] for r in res]); pts = np.stack([r["points"] for r in res])
    assign = np.stack([pnp_ref.assign_table(r["points"], r["logits"]) for r in res])
    for i, r in enumerate(res):
        try:                                                               # RV/gen_submission_single.py:169-175
            q, t = solver(r["points"], r["logits"])
            quat[i], tvec[i], ok[i] = q, t, 1
        except (IndexError, cv2.error):
            pass
    np.savez_compressed(os.path.join(GOLDEN, "pnp_golden.npz"), quat=quat, tvec=tvec, ok=ok, probs=probs,
                        points_px=pts, assign=assign, n=n, seed=1)
    print("pnp cases", n, "solved", int(ok.sum()))


def write_pnp_multi(rv_eval, n=200, num_models=5):
    """Ensemble solver: the reference's own PostProcess per member + its own Multi_Mean_PoseSolver."""
    import cv2
    import io
    import contextlib
    d = synth.make_multi_predictions(n, num_models=num_models, seed=7)
    args = synth.reference_args(ModelCfg())
    args.repro = 25                                                         # RV/gen_submission_multi.py:106
    solver = rv_eval.Multi_Mean_PoseSolver(args)                            # the reference's own ensemble solver
    post = sys.modules["models"].PostProcess()
    boxes = [torch.from_numpy(b) for b in d["boxes"]]
    per_model = [post({"pred_logits": torch.from_numpy(d["logits"][m]),
                       "pred_points": torch.from_numpy(d["points"][m]).clone()}, boxes) for m in range(num_models)]
    quat = np.zeros((n, 4)); tvec = np.zeros((n, 3)); ok = np.zeros(n, dtype=np.int32)
    pooled = np.zeros((n, 11, 2), dtype=np.float32); count = np.zeros((n, 11), dtype=np.int32)
    oracle = pnp_ref.MultiMeanPoseSolver(25, return_details=True)
    nan_cases = 0
    for i in range(n):
        mp = [per_model[m][i]["points"] for m in range(num_models)]
        ml = [per_model[m][i]["logits"] for m in range(num_models)]
        # the restatement's pooling must equal the reference's own mean_and_filter bit for bit
        mean_o, cnt_o = oracle.pool(mp, ml)
        from collections import defaultdict
        orig = defaultdict(list)
        for points, logits in zip(mp, ml):
            labels, scores = solver.find_index(logits)
            fg = labels != logits.shape[1] - 1
            for pt_, l_ in zip(points[fg], labels[fg]):
                orig[l_].append(pt_)
        import warnings
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            mean_r = solver.mean_and_filter(orig)
        assert list(mean_r.keys()) == list(mean_o.keys())
        for l_ in mean_r:
            assert np.array_equal(mean_r[l_], mean_o[l_], equal_nan=True), (i, l_)
        if any(np.isnan(v).any() for v in mean_r.values()):
            nan_cases += 1
        pooled[i], count[i] = pnp_ref.pooled_table(mean_o, cnt_o)
        try:                                                                # RV/gen_submission_multi.py:166-171
            with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
                warnings.simplefilter("ignore")
                q, t = solver(mp, ml)
            if np.all(np.isfinite(q)) and np.all(np.isfinite(t)):
                quat[i], tvec[i], ok[i] = q, t, 1
        except (IndexError, cv2.error):
            pass
    np.savez_compressed(os.path.join(GOLDEN, "pnp_multi_golden.npz"), quat=quat, tvec=tvec, ok=ok, pooled_px=pooled,
                        count=count, n=n, num_models=num_models, seed=7)
    print("ensemble pnp cases", n, "solved", int(ok.sum()), "with an emptied label", nan_cases)


def main():
    if not ref_import.available():
        raise SystemExit("/root/reference is not mounted: golden vectors can only be regenerated in the build container")
    os.makedirs(GOLDEN, exist_ok=True)
    write_boxes()
    rv_speed, rv_eval = import_rv_dataset_and_solver()
    write_crop(rv_speed)
    write_crop_eval(rv_speed)
    write_model()
    write_model_b256(rv_eval)
    write_model_b64_random()
    write_deform_attn()
    write_sa_model()
    write_sa_model(SA_R18_CASE, "sa_r18_model_golden.npz", ref_import.SA_MODEL_YML_R18)   # rtdetr_r18vd_6x_speed_kl_1.yml
    write_pnp(rv_eval)
    write_pnp_multi(rv_eval)


if __name__ == "__main__":
    main()
