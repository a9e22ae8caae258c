"""Keypoint-set predictor forward through the C ABI against the oracle (restated reference forward, fp32 CPU) and
the golden outputs of the real reference."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import model_ref, pnp_ref, synth
from oracle.make_golden import MODEL_CASES, model_inputs
from satellite_pose_estimation_b200 import Engine, _lib, build_model, build_solver

pytestmark = pytest.mark.gpu

# north_star: keypoints within 0.5 px in fp32/TF32 mode.  pixel = sigmoid * S + x1, so the bar on the normalised
# output is 0.5 / S; we hold it at the LARGEST crop of the real box distribution (S = 1748) -> 2.86e-4.
S_MAX = 1748
PTS_TOL_TF32 = 0.5 / S_MAX
LOGIT_TOL_TF32 = 1e-2
PTS_TOL_BF16 = 8e-3          # bf16 cannot meet 0.5 px at large crops: reported, looser bar
LOGIT_TOL_BF16 = 0.15


def _engine(cfg, R, B, precision):
    e = Engine(input_size=R, num_queries=cfg.num_queries, enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers,
               backbone=cfg.backbone, precision=precision, has_sigma=cfg.sigma_head, max_batch=B)
    return e


@pytest.mark.parametrize("case", list(MODEL_CASES))
def test_forward_matches_reference_golden_tf32(lib, cuda_dev, case):
    g = np.load(os.path.join(synth.GOLDEN_DIR, "model_golden.npz"))
    kw, B, R, seed = MODEL_CASES[case]
    cfg = model_ref.ModelCfg(**kw)
    sd = synth.make_state_dict(cfg, seed=seed)
    assert synth.weights_checksum(sd).encode() == g[case + "/checksum"].tobytes()
    eng = _engine(cfg, R, B, "tf32")
    eng.load_state_dict(sd)
    out = eng.forward(model_inputs(B, R, seed).cuda(), want_aux=True)
    torch.cuda.synchronize()
    dp = np.abs(out["pred_points"].cpu().numpy() - g[case + "/pred_points"]).max()
    dl = np.abs(out["pred_logits"].cpu().numpy() - g[case + "/pred_logits"]).max()
    assert dp <= PTS_TOL_TF32, f"keypoints off by {dp * S_MAX:.3f} px at S={S_MAX}"
    assert dl <= LOGIT_TOL_TF32
    aux_p = torch.stack([a["pred_points"] for a in out["aux_outputs"]]).cpu().numpy()
    aux_l = torch.stack([a["pred_logits"] for a in out["aux_outputs"]]).cpu().numpy()
    assert np.abs(aux_p - g[case + "/aux_points"]).max() <= PTS_TOL_TF32
    assert np.abs(aux_l - g[case + "/aux_logits"]).max() <= LOGIT_TOL_TF32
    assert np.array_equal(out["pred_logits"].argmax(-1).cpu().numpy(), g[case + "/pred_logits"].argmax(-1))
    eng.close()


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_forward_layerwise_vs_oracle(lib, cuda_dev, precision):
    cfg = model_ref.ModelCfg(sigma_head=True)
    sd = synth.make_state_dict(cfg, seed=3)
    B = 3
    x = model_inputs(B, 224, 11)
    eng = _engine(cfg, 224, 4, precision)
    eng.load_state_dict(sd)
    eng.enable_taps(True)
    out = eng.forward(x.cuda(), want_aux=True)
    torch.cuda.synchronize()
    taps = {}
    ref = model_ref.forward(sd, cfg, x, taps)
    rel = 6e-3 if precision == "tf32" else 4e-2

    def check(name, got, want):
        err = (got - want).abs().max().item() / want.abs().max().item()
        assert err < rel, f"{name}: rel err {err:.2e}"

    nhwc = lambda t: t.permute(0, 2, 3, 1)
    check("stem", eng.read_tap("stem", (B, 112, 112, 64)), nhwc(taps["stem"]))
    check("layer1", eng.read_tap("layer1", (B, 56, 56, 256)), nhwc(taps["layer1"]))
    check("layer2", eng.read_tap("layer2", (B, 28, 28, 512)), nhwc(taps["layer2"]))
    check("layer3", eng.read_tap("layer3", (B, 14, 14, 1024)), nhwc(taps["layer3"]))
    check("neck", eng.read_tap("neck", (B, 28, 28, 512)), nhwc(taps["neck"]))
    for i in range(cfg.enc_layers):
        check(f"enc{i}", eng.read_tap(f"enc{i}", (B, 784, 256)), taps[f"enc{i}"].permute(1, 0, 2))
    check("hs", eng.read_tap("hs", (cfg.dec_layers, B, 40, 256)), taps["hs"])
    ptol, ltol = (PTS_TOL_TF32, LOGIT_TOL_TF32) if precision == "tf32" else (PTS_TOL_BF16, LOGIT_TOL_BF16)
    assert (out["pred_points"].cpu() - ref["pred_points"]).abs().max() <= ptol
    assert (out["pred_logits"].cpu() - ref["pred_logits"]).abs().max() <= ltol
    assert (out["pred_sigmas"].cpu() - ref["pred_sigmas"]).abs().max() <= ltol
    assert torch.equal(out["pred_sigmas"][..., 0], out["pred_sigmas"][..., 1])
    eng.close()


def test_batch_invariance_and_chunking(lib, cuda_dev):
    """An image's outputs do not depend on its batch neighbours; batches above max_batch are chunked."""
    cfg = model_ref.ModelCfg()
    sd = synth.make_state_dict(cfg, seed=0)
    args = SimpleNamespace(backbone="resnet50s8", num_queries=40, enc_layers=4, dec_layers=4, hidden_dim=256, nheads=8,
                           dim_feedforward=2048, aux_loss=False, device="cuda", repro=20, max_batch=4)
    model, _, post = build_model(args)
    model.to("cuda")
    model.load_state_dict(sd, strict=True)
    x = model_inputs(6, 224, 5).cuda()
    full = model(x)                       # 6 > max_batch=4 -> two chunks
    one = model(x[4:5])
    assert "aux_outputs" not in full
    assert torch.allclose(full["pred_points"][4], one["pred_points"][0], atol=1e-6)
    assert torch.allclose(full["pred_logits"][4], one["pred_logits"][0], atol=1e-5)
    lst = model([x[0], x[1]])            # list[Tensor] input like the reference
    assert torch.allclose(lst["pred_points"], full["pred_points"][:2], atol=1e-6)


def test_large_batch_kernels_agree_with_small_batch_path(lib, cuda_dev):
    """At B = 16 the schedule switches to the machine-filling kernels (fused feed-forward block + norm2, wide / paired
    GEMM tiles); an image's keypoints must stay within a fraction of the 0.5 px budget of its batch-1 result
    (which the golden tests pin against the reference).  1748 px = the largest crop side of the box distribution."""
    cfg = model_ref.ModelCfg()
    eng = _engine(cfg, 224, 16, "tf32")
    eng.load_state_dict(synth.make_state_dict(cfg, seed=0))
    x = model_inputs(16, 224, 11).cuda()
    big = eng.forward(x)
    pts_big = big["pred_points"].cpu().numpy().copy()
    log_big = big["pred_logits"].cpu().numpy().copy()
    for i in (0, 7, 15):
        one = eng.forward(x[i:i + 1])
        d_px = np.abs(one["pred_points"].cpu().numpy()[0] - pts_big[i]).max() * 1748
        assert d_px < 0.1, (i, d_px)
        assert np.abs(one["pred_logits"].cpu().numpy()[0] - log_big[i]).max() < 5e-3
    eng.close()


def test_drop_in_loop_like_gen_submission(lib, cuda_dev):
    """The reference's hot loop (RV/gen_submission_single.py:136-181) with the drop-in objects, on seeded inputs;
    outputs are compared with the oracle chain (restated forward + PostProcess + cv2 solver)."""
    cfg = model_ref.ModelCfg()
    sd = synth.make_state_dict(cfg, seed=0)
    args = SimpleNamespace(backbone="resnet50s8", num_queries=40, enc_layers=4, dec_layers=4, hidden_dim=256, nheads=8,
                           dim_feedforward=2048, aux_loss=True, device="cuda", repro=20)
    model, criterion, postprocessors = build_model(args)
    model.to("cuda")
    model.load_state_dict(sd, strict=True)
    model.eval()
    solver = build_solver(args, model, postprocessors)
    x = model_inputs(2, 224, 9)
    clip = [torch.tensor([300, 200, 816, 716]), torch.tensor([-20, 40, 1128, 1188])]
    outputs = model(x.to("cuda"))
    results = postprocessors["points"](outputs, clip)
    ref_out = model_ref.forward(sd, cfg, x)
    ref_res = pnp_ref.post_process(ref_out["pred_logits"], ref_out["pred_points"], clip)
    for r, rr, box in zip(results, ref_res, clip):
        side = float(box[2] - box[0])
        assert np.abs(r["points"] - rr["points"]).max() <= 0.5 * side / S_MAX + 1e-3   # <= 0.5 px at S_MAX scale
        assert np.abs(r["logits"] - rr["logits"]).max() < 5e-3
        try:
            q, t = solver(r["points"], r["logits"])
            ok = True
        except IndexError:
            q, t, ok = np.zeros(4), np.zeros(3), False
        q_ref, t_ref, ok_ref = pnp_ref.solve_or_zero(pnp_ref.SimplePoseSolver(20), rr["points"], rr["logits"])
        assert ok == ok_ref                              # random-init weights collapse to one label -> both fail


def test_host_pipeline_matches_stagewise(lib, cuda_dev):
    """spe_run_batch_host (host in, host out) == the three stage calls on device buffers."""
    cfg = model_ref.ModelCfg()
    eng = _engine(cfg, 224, 8, "tf32")
    eng.load_state_dict(synth.make_state_dict(cfg, seed=0))
    det = synth.load_detector_boxes()[:8]
    frames = synth.make_frames(8, det, seed=2)
    r = eng.run_batch_host(torch.from_numpy(frames).pin_memory(), det)
    clip = eng.clip_boxes(det)
    assert np.array_equal(r["boxes"], clip)
    img = eng.crop_resize_norm(torch.from_numpy(frames).cuda(), torch.from_numpy(clip).cuda())
    out = eng.forward(img)
    p = eng.assign_pnp(out["pred_logits"], out["pred_points"], torch.from_numpy(clip).cuda())
    assert np.array_equal(r["status"], p["status"].cpu().numpy())
    assert np.allclose(r["quat"], p["quat"].cpu().numpy()) and np.allclose(r["tvec"], p["tvec"].cpu().numpy())
    # double-buffered form: two batches in flight give the same answers as the synchronous call
    fh = [torch.from_numpy(frames).pin_memory(), torch.from_numpy(frames[::-1].copy()).pin_memory()]
    dets = [det, det[::-1].copy()]
    eng.submit_batch_host(0, fh[0], dets[0])
    eng.submit_batch_host(1, fh[1], dets[1])
    a, b = eng.collect_batch_host(0), eng.collect_batch_host(1)
    assert np.array_equal(a["status"], r["status"]) and np.allclose(a["quat"], r["quat"])
    assert np.array_equal(b["boxes"], clip[::-1]) and np.array_equal(b["status"], r["status"][::-1])
    assert 0 < a["h2d_bytes"] < frames.size          # only the crop-box / frame intersections are uploaded
    with pytest.raises(Exception):
        eng.collect_batch_host(0)                    # nothing in flight any more
    eng.close()


@pytest.mark.parametrize("slots", [2, 4])
def test_multi_slot_pipeline_matches_serial(lib, cuda_dev, slots):
    """Whole batches in flight next to each other (own streams + activation set per slot) change nothing: network
    outputs and poses of every batch are bit-identical to the one-stream stage calls."""
    cfg = model_ref.ModelCfg()
    B = 8
    eng = _engine(cfg, 224, B, "tf32")
    eng.load_state_dict(synth.make_state_dict(cfg, seed=0))
    det_all = synth.load_detector_boxes()
    preds = synth.make_predictions(B, Q=cfg.num_queries, seed=5)
    syn = [torch.from_numpy(preds[k]).cuda() for k in ("logits", "points")]
    syn_boxes = torch.from_numpy(preds["boxes"]).to(torch.int32).cuda()
    sets = []
    for k in range(7):
        det = det_all[k * B:(k + 1) * B]
        frames = torch.from_numpy(synth.make_frames(B, det, seed=10 + k)).cuda()
        boxes = torch.from_numpy(eng.clip_boxes(det)).cuda()
        out = eng.forward(eng.crop_resize_norm(frames, boxes))
        ref = (out["pred_logits"].cpu().numpy().copy(), out["pred_points"].cpu().numpy().copy())
        sets.append((frames, boxes, ref))
    pose_ref = eng.assign_pnp(syn[0], syn[1], syn_boxes)
    assert int((pose_ref["status"] == 0).sum()) >= B // 2
    eng.set_pnp_override(syn[0], syn[1], syn_boxes)
    n = len(sets)
    for rep in range(2):                                 # second round replays the captured graphs
        for i in range(min(slots, n)):
            eng.submit_batch_dev(i % slots, sets[i][0], sets[i][1])
        for i in range(n):
            r = eng.collect_batch_host(i % slots)
            logits, points = eng.read_slot_outputs(i % slots, B)
            assert np.array_equal(logits, sets[i][2][0]) and np.array_equal(points, sets[i][2][1]), (rep, i)
            assert np.array_equal(r["status"], pose_ref["status"].cpu().numpy())
            assert np.array_equal(r["quat"], pose_ref["quat"].cpu().numpy())
            assert np.array_equal(r["tvec"], pose_ref["tvec"].cpu().numpy())
            if i + slots < n:
                eng.submit_batch_dev(i % slots, sets[i + slots][0], sets[i + slots][1])
    # the one-stream call still works afterwards (activation set 0 is selected again)
    out = eng.forward(eng.crop_resize_norm(sets[3][0], sets[3][1]))
    assert np.array_equal(out["pred_points"].cpu().numpy(), sets[3][2][1])
    with pytest.raises(Exception):
        eng.submit_batch_dev(_lib.PIPELINE_SLOTS, sets[0][0], sets[0][1])   # no such slot
    eng.set_pnp_override(None, None, None)
    eng.close()


def test_errors_are_loud(lib, cuda_dev):
    cfg = model_ref.ModelCfg()
    eng = _engine(cfg, 224, 2, "tf32")
    from satellite_pose_estimation_b200._lib import SpeError
    with pytest.raises(SpeError, match="spe_load_weights first"):
        eng.forward(torch.zeros(1, 3, 224, 224, device="cuda"))
    sd = synth.make_state_dict(cfg, seed=0)
    bad = dict(sd); bad.pop("input_proj.bias")
    with pytest.raises(SpeError, match="input_proj.bias"):
        eng.load_state_dict(bad)
    bad = dict(sd); bad["cls_embed.weight"] = torch.zeros(13, 256)
    with pytest.raises(SpeError, match="cls_embed.weight"):
        eng.load_state_dict(bad)
    eng.load_state_dict(sd)
    with pytest.raises(SpeError, match="max_batch"):
        eng.forward(torch.zeros(3, 3, 224, 224, device="cuda"))
    eng.close()


def test_image_set_runner_shards_and_matches_batch_calls(lib, cuda_dev):
    """BASELINE configs[4] in miniature: a 150-image set through run_image_set (three batches in flight, ragged last
    batch, two emulated ranks) gives exactly what one spe_run_batch_host call per batch gives, keyed by filename."""
    from satellite_pose_estimation_b200.submission import run_image_set
    cfg = model_ref.ModelCfg()
    eng = Engine(max_batch=64)
    eng.load_state_dict(synth.make_state_dict(cfg, seed=0))
    n = 150
    det = synth.load_detector_boxes()[:n]
    base = synth.make_frames(8, det, seed=3)
    frames = np.stack([np.roll(base[i % 8], 11 * i, axis=1) for i in range(n)])
    names = [f"img{(7 * i) % n:06d}.jpg" for i in range(n)]                          # not in filename order
    preds = synth.make_predictions(64, seed=5)
    eng.set_pnp_override(torch.from_numpy(preds["logits"]).cuda(), torch.from_numpy(preds["points"]).cuda(),
                         torch.from_numpy(preds["boxes"]).to(torch.int32).cuda())     # so that poses are non-trivial
    get = lambda a, b: frames[a:b]
    merged = {}
    for rank in range(2):
        part = run_image_set(eng, get, det, names, batch_size=64, rank=rank, world_size=2, slots=3, gather=False)
        assert not (set(part) & set(merged))
        merged.update(part)
    assert sorted(merged) == sorted(names)
    from satellite_pose_estimation_b200.sharding import shard_range, batches
    from satellite_pose_estimation_b200.submission import log_entry
    for rank in range(2):
        a, b = shard_range(n, rank, 2)
        for i0, i1 in batches(a, b, 64):
            r = eng.run_batch_host(torch.from_numpy(frames[i0:i1]).pin_memory(), det[i0:i1])
            for j, i in enumerate(range(i0, i1)):
                ok = r["status"][j] == 0
                want = log_entry(r["quat"][j] if ok else np.zeros(4), r["tvec"][j] if ok else np.zeros(3))
                got = merged[names[i]]
                assert got["status"] == int(r["status"][j]) and got["quat_pr"] == want["quat_pr"] and \
                    got["tvec_pr"] == want["tvec_pr"], (i, got, want)
    assert sum(1 for v in merged.values() if v["status"] == 0) > 100
    eng.set_pnp_override(None, None, None)
    eng.close()


def test_ensemble_submission_loop(lib, cuda_dev):
    """gen_submission of the ensemble path (RV/gen_submission_multi.py:145-186): the batched kernel path equals the
    per-file solver calls, failures become the zero pose, values are rounded to 6 decimals."""
    from collections import defaultdict
    from satellite_pose_estimation_b200 import MultiMeanPoseSolver, gen_submission
    nm, n = 3, 40
    d = synth.make_multi_predictions(n, num_models=nm, seed=11)
    prediction = defaultdict(list)
    for m in range(nm):
        for i, r in enumerate(pnp_ref.post_process(d["logits"][m], d["points"][m], d["boxes"])):
            prediction[f"img{i:06d}.jpg"].append(r)
    eng = Engine(max_batch=1)
    solver = MultiMeanPoseSolver(reproj=25, engine=eng)
    log = gen_submission(prediction, solver)
    assert list(log) == list(prediction)

    class PerFile:                       # any non-batched solver object goes through the reference's per-file loop
        def __call__(self, mp, ml):
            return solver(mp, ml)
    log2 = gen_submission(prediction, PerFile())
    assert log == log2
    zero = [f for f, v in log.items() if not any(v["quat_pr"])]
    assert 0 < len(zero) < n // 2 and all(log[f]["tvec_pr"] == [0.0, 0.0, 0.0] for f in zero)
    eng.close()


@pytest.mark.parametrize("precision,sigma", [("bf16", False), ("tf32", True)])
def test_host_pipeline_at_batch_256(lib, cuda_dev, precision, sigma):
    """BASELINE.json configs[2] / configs[3] at their stated batch size through the whole host pipeline (crop ->
    predictor -> PnP on 256 pinned frames); their numerics are pinned against the live reference in
    tests/test_gpu_bench_configs.py::test_batch256_configs_vs_live_reference."""
    cfg = model_ref.ModelCfg(sigma_head=sigma)
    B = 256
    eng = _engine(cfg, 224, B, precision)
    eng.load_state_dict(synth.make_state_dict(cfg, seed=0, spread_labels=True))
    det = synth.load_detector_boxes()[:B]
    base = synth.make_frames(8, det, seed=2)
    frames = torch.from_numpy(np.concatenate([base] * (B // 8))).pin_memory()
    out = eng.run_batch_host(frames, det, reproj=25.0 if sigma else 20.0, weighted=sigma, reject=sigma)
    assert out["status"].shape == (B,) and np.isfinite(out["quat"]).all() and np.isfinite(out["tvec"]).all()
    # the sigma variant runs the reject filter: poses whose inlier RMS reprojection error exceeds 5 px are flagged (3)
    assert (out["status"] == 0).mean() > (0.6 if sigma else 0.9) and set(np.unique(out["status"])) <= {0, 1, 3}
    eng.close()


def test_backbone_view_like_get_backbone_time(lib, cuda_dev):
    """``model.backbone(sample)`` (RV/get_backbone_time.py:110): features + sine position embedding as the reference's
    Joiner returns them, against the oracle's backbone and position embedding."""
    cfg = model_ref.ModelCfg()
    sd = synth.make_state_dict(cfg, seed=0)
    args = SimpleNamespace(backbone="resnet50s8", num_queries=40, enc_layers=4, dec_layers=4, hidden_dim=256, nheads=8,
                           dim_feedforward=2048, aux_loss=False, device="cuda", repro=20, max_batch=4)
    model, _, _ = build_model(args)
    model.to("cuda")
    model.load_state_dict(sd, strict=True)
    x = model_inputs(2, 224, 3)
    feats, pos = model.backbone(x.cuda())
    assert len(feats) == 1 and len(pos) == 1
    taps = {}
    model_ref.forward(sd, cfg, x, taps)
    f = feats[0].tensors.cpu()
    assert f.shape == (2, 512, 28, 28) and not feats[0].mask.any()
    assert (f - taps["neck"]).abs().max().item() / taps["neck"].abs().max().item() < 6e-3
    assert torch.allclose(pos[0].cpu(), model_ref.position_embedding_sine(2, 28, 28), atol=1e-6)
    out = model(x.cuda())                     # the fused forward still works after the tap round trip
    assert out["pred_points"].shape == (2, 40, 2)


def test_learned_position_embedding_vs_oracle(lib, cuda_dev):
    """--position_embedding learned: the embedding tables of the checkpoint replace the sine table in the folded
    positional addends (encoder Q|K, decoder cross-attention K); forward against the oracle."""
    cfg = model_ref.ModelCfg(num_queries=20, enc_layers=2, dec_layers=2, position_embedding="learned")
    sd = synth.make_state_dict(cfg, seed=6)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(4))
    ref = model_ref.forward(sd, cfg, x)
    sine = model_ref.forward({k: v for k, v in sd.items() if not k.startswith("backbone.1.")},
                             model_ref.ModelCfg(num_queries=20, enc_layers=2, dec_layers=2), x)
    e = Engine(num_queries=20, enc_layers=2, dec_layers=2, max_batch=2)
    try:
        e.load_state_dict(sd)
        out = e.forward(x.cuda())
        torch.cuda.synchronize()
        d = (out["pred_points"].cpu() - ref["pred_points"]).abs().max().item()
        assert d <= 0.5 / 1748, d * 1748
        # the two embeddings give different predictions: the tables were really used
        assert (ref["pred_points"] - sine["pred_points"]).abs().max().item() > 20 * d
    finally:
        e.close()


def test_main_py_default_configuration_vs_oracle(lib, cuda_dev):
    """The reference's main.py defaults (secondary configuration): --backbone resnet50 (stride 16), 512 x 512
    crops, 100 queries, 6 + 6 layers -- 32 x 32 = 1024 tokens, a 256-wide stem output row (two window tiles per row),
    128 x 128 layer1 maps.  Keypoints within 0.5 px at S = 1748 after the rounding-bias calibration, same arg-max labels."""
    cfg = model_ref.ModelCfg(backbone="resnet50", num_queries=100, enc_layers=6, dec_layers=6)
    sd = synth.make_state_dict(cfg, seed=2)
    x = torch.randn(2, 3, 512, 512, generator=torch.Generator().manual_seed(12))
    ref = model_ref.forward(sd, cfg, x)
    e = Engine(input_size=512, num_queries=100, enc_layers=6, dec_layers=6, backbone="resnet50", max_batch=2)
    try:
        e.load_state_dict(sd)
        xc = x.cuda()
        e.calibrate(xc)
        out = e.forward(xc)
        torch.cuda.synchronize()
        d = (out["pred_points"].cpu() - ref["pred_points"]).abs().max().item()
        print(f"main.py defaults (s16, 512^2, Q=100, 6+6): keypoints {d * 1748:.3f} px at S=1748")
        assert d <= 0.5 / 1748
        assert torch.equal(out["pred_logits"].argmax(-1).cpu(), ref["pred_logits"].argmax(-1))
    finally:
        e.close()
