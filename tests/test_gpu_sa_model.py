"""The SA drop's RT-DETR keypoint predictor on the GPU (spe_config.backbone = 2) against
  * oracle/sa_model_ref.py, stage by stage (activation taps), and
  * tests/golden/sa_model_golden.npz: outputs of the SA drop's LIVE model (oracle/make_golden.py:write_sa_model).

The top-k query selection (SA/src/zoo/rtdetr/rtdetr_decoder.py:646-648) is a discontinuous function of anchor scores
that no two floating-point implementations reproduce bit for bit (the 30th and 31st score of an image are 4e-4 apart
here, the TF32 backbone is accurate to ~2e-2 on them), so parity is stated in two parts: (1) with the reference's
selection handed in (``topk_override``) every output matches the reference within the tolerance below; (2) the
library's own selection is an exact top-k of its own scores, and every anchor it picks is within rounding distance of
the reference's cut.  Tolerances: the north-star bar, keypoints within 0.5 px at the largest crop side (1748 px) =
2.9e-4 of the unit square; measured 0.06 px with the default 3xTF32 schedule (fp32 storage, error-compensated tensor-core
products everywhere) -- the tests assert 0.15 px and 2e-3 on logits / log-sigma.  The plain-TF32 trunk (SPE_SA_X3=0, 1.45x
the throughput) measures 0.1 px rms but up to 0.7 px on single keypoints, i.e. it does NOT meet the bar at the largest crops;
it is an opt-in and its test states what it delivers.
"""
import os

import numpy as np
import pytest
import torch

from oracle import pnp_ref, sa_model_ref, synth
from oracle.make_golden import SA_MODEL_CASE, model_inputs
from satellite_pose_estimation_b200 import Engine
from satellite_pose_estimation_b200.sa_models import build_sa_model, build_sigma_solver

pytestmark = pytest.mark.gpu

PTS_TOL = 0.15 / 1748
LOGIT_TOL = 2e-3


@pytest.fixture(scope="module")
def case():
    cfg = sa_model_ref.SaCfg()
    sd = synth.make_sa_state_dict(cfg, seed=SA_MODEL_CASE["weights_seed"])
    x = model_inputs(SA_MODEL_CASE["batch"], cfg.input_size, SA_MODEL_CASE["seed"])
    g = np.load(os.path.join(synth.GOLDEN_DIR, "sa_model_golden.npz"))
    assert synth.weights_checksum(sd) == str(g["weights_sha256"])
    return cfg, sd, x, g


@pytest.fixture(scope="module")
def eng(lib, cuda_dev, case):
    cfg, sd, x, g = case
    e = Engine(input_size=cfg.input_size, num_queries=cfg.num_queries, enc_layers=1, dec_layers=cfg.dec_layers,
               dim_feedforward=cfg.dec_ff, backbone="rtdetr_r50vd", precision="tf32", has_sigma=True, max_batch=8)
    e.load_state_dict(sd)
    yield e
    e.close()


def _rel_rms(a, b):
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()


def test_sa_forward_layerwise_vs_oracle(eng, case):
    """Every stage of the schedule against the oracle's activations: stem, the four residual stages, the AIFI layer, the
    three fused encoder outputs, the decoder memory, the anchor scores and the decoder layers (reference selection)."""
    cfg, sd, x, g = case
    B, R, Q = x.shape[0], cfg.input_size, cfg.num_queries
    taps = {}
    sa_model_ref.forward(sd, cfg, x, taps)
    eng.enable_taps(True)
    try:
        eng.forward_sa(x.cuda(), topk_override=taps["topk"].to(torch.int32).cuda())
        torch.cuda.synchronize()
        nhwc = lambda name, H, C: eng.read_tap(name, (B, H, H, C)).permute(0, 3, 1, 2)
        errs = {"stem": _rel_rms(nhwc("sa_stem", R // 2, 64), taps["stem"])}
        for i, (s, c) in enumerate(((4, 256), (8, 512), (16, 1024), (32, 2048))):
            errs[f"stage{i}"] = _rel_rms(nhwc(f"sa_stage{i}", R // s, c), taps[f"stage{i}"])
        errs["aifi"] = _rel_rms(nhwc("sa_aifi", R // 32, 256), taps["aifi"])
        for i, s in enumerate((8, 16, 32)):
            errs[f"enc_out{i}"] = _rel_rms(nhwc(f"sa_enc{i}", R // s, 256), taps[f"enc_out{i}"])
        Lv = taps["memory"].shape[1]
        errs["memory"] = _rel_rms(eng.read_tap("sa_memory", (B, Lv, 256)), taps["memory"])
        errs["scores"] = _rel_rms(eng.read_tap("sa_enc_scores", (B, Lv, 12)).max(-1).values, taps["enc_scores"])
        for i in range(cfg.dec_layers):
            errs[f"dec{i}"] = _rel_rms(eng.read_tap(f"sa_dec{i}", (B, Q, 256)), taps[f"dec{i}"])
    finally:
        eng.enable_taps(False)
    print("SA layer-wise relative rms error:", {k: f"{v:.1e}" for k, v in errs.items()})
    assert max(errs.values()) < 4e-4, errs


def test_sa_outputs_vs_live_reference_golden(eng, case):
    """Part (1): the live model's selection handed in -> the whole output dict of the live model."""
    cfg, sd, x, g = case
    out = eng.forward_sa(x.cuda(), topk_override=torch.from_numpy(g["topk"]).cuda())
    torch.cuda.synchronize()
    d_l = np.abs(out["pred_logits"].cpu().numpy() - g["pred_logits"]).max()
    d_p = np.abs(out["pred_pts"].cpu().numpy() - g["pred_pts"]).max()
    d_s = np.abs(out["pred_sigmas"].cpu().numpy() - g["pred_sigmas"]).max()
    aux = out["aux_outputs"]
    assert len(aux) == cfg.dec_layers and "pred_sigmas" not in aux[-1]
    a_l = np.abs(torch.stack([a["pred_logits"] for a in aux]).cpu().numpy() - g["aux_logits"]).max()
    a_p = np.abs(torch.stack([a["pred_pts"] for a in aux]).cpu().numpy() - g["aux_pts"]).max()
    a_s = np.abs(torch.stack([a["pred_sigmas"] for a in aux[:-1]]).cpu().numpy() - g["aux_sigmas"]).max()
    print(f"SA forward vs live reference: keypoints {d_p * 1748:.3f} px at S=1748 (aux {a_p * 1748:.3f}), logits {d_l:.1e} "
          f"(aux {a_l:.1e}), log-sigma {d_s:.1e} (aux {a_s:.1e})")
    assert d_p <= PTS_TOL and a_p <= PTS_TOL
    assert d_l <= LOGIT_TOL and a_l <= 2 * LOGIT_TOL and d_s <= LOGIT_TOL and a_s <= LOGIT_TOL
    assert np.array_equal(out["pred_logits"].argmax(-1).cpu().numpy(), g["pred_logits"].argmax(-1))
    s = out["pred_sigmas"]
    assert torch.equal(s[..., 0], s[..., 1])


def test_sa_own_topk_selection(eng, case):
    """Part (2): the library's own selection.  It is the exact top-k (torch.topk order: descending, ties to the lower
    index) of the library's own anchor scores; against the reference's scores every selected anchor lies above the
    reference's cut minus the score tolerance; and the outputs equal the oracle's when the oracle is handed this
    selection."""
    cfg, sd, x, g = case
    B, Q = x.shape[0], cfg.num_queries
    eng.enable_taps(True)
    try:
        out = eng.forward_sa(x.cuda())
        torch.cuda.synchronize()
        scores = eng.read_tap("sa_enc_scores", (B, g["enc_scores"].shape[1], 12)).max(-1).values
    finally:
        eng.enable_taps(False)
    tk = out["topk_ind"].cpu().long()
    assert torch.equal(tk, torch.topk(scores, Q, dim=1)[1])
    ref_scores = torch.from_numpy(g["enc_scores"])
    tol = (scores - ref_scores).abs().max().item()
    assert tol < 5e-3
    cut = ref_scores.sort(dim=1, descending=True).values[:, Q - 1]
    assert (ref_scores.gather(1, tk) >= cut[:, None] - 2 * tol).all()
    common = [len(set(tk[b].tolist()) & set(g["topk"][b].tolist())) for b in range(B)]
    ref = sa_model_ref.forward(sd, cfg, x, None, topk_override=tk)
    d_p = (out["pred_pts"].cpu() - ref["pred_pts"]).abs().max().item()
    d_l = (out["pred_logits"].cpu() - ref["pred_logits"]).abs().max().item()
    d_s = (out["pred_sigmas"].cpu() - ref["pred_sigmas"]).abs().max().item()
    print(f"SA own top-k: {common} of {Q} anchors shared with the live model per image, score tolerance {tol:.1e}; "
          f"vs oracle on the same selection: keypoints {d_p * 1748:.3f} px, logits {d_l:.1e}, log-sigma {d_s:.1e}")
    assert min(common) >= Q - 2
    assert d_p <= PTS_TOL and d_l <= LOGIT_TOL and d_s <= LOGIT_TOL


def test_sa_plain_forward_replays_graph_and_batches_agree(eng, case):
    """spe_forward serves the same ctx (last layer only) and replays a captured graph from the third call on; a batch of
    one and a larger batch give the same per-image results (the reference itself cannot run batch 1:
    rtdetr_decoder.py:168 squeezes the batch dimension away)."""
    cfg, sd, x, g = case
    xc = x.cuda()
    ref = eng.forward_sa(xc)
    outs = [{k: v.clone() for k, v in eng.forward(xc).items()} for _ in range(3)]
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o["pred_logits"], ref["pred_logits"]) and torch.equal(o["pred_points"], ref["pred_pts"])
        assert torch.equal(o["pred_sigmas"], ref["pred_sigmas"])
    one = eng.forward_sa(xc[1:2])
    assert (one["pred_pts"][0] - ref["pred_pts"][1]).abs().max().item() <= PTS_TOL
    assert torch.equal(one["topk_ind"][0], ref["topk_ind"][1]) or len(
        set(one["topk_ind"][0].tolist()) & set(ref["topk_ind"][1].tolist())) >= cfg.num_queries - 2


def test_sa_model_mirror_and_postprocessor_chain(lib, cuda_dev, case):
    """The host mirror: ``build_sa_model`` -> model(x) -> postprocessor(outputs, clip_bbox), against the oracle forward +
    the SA post-processing restated from rtdetr_postprocessor.py:43-78 (softmax, exp(sigma), de-normalisation)."""
    cfg, sd, x, g = case
    model, post = build_sa_model(max_batch=4)
    model.load_state_dict(sd, strict=True)
    model.to("cuda")
    out = model(x[:3].cuda())
    assert set(out) == {"pred_logits", "pred_pts", "pred_sigmas", "aux_outputs"} and len(out["aux_outputs"]) == cfg.dec_layers
    boxes = [torch.tensor([100, 50, 612, 562]), torch.tensor([-40, 300, 700, 1040]), torch.tensor([900, 200, 1500, 800])]
    res = post(out, boxes)
    assert len(res) == 3 and set(res[0]) == {"logits", "points", "sigmas"}
    logits, pts, sig = out["pred_logits"].cpu(), out["pred_pts"].cpu(), out["pred_sigmas"].cpu()
    want = pnp_ref.post_process(logits, pts, boxes)
    for i in range(3):
        assert np.abs(res[i]["logits"] - want[i]["logits"]).max() < 1e-6
        assert np.array_equal(res[i]["points"], want[i]["points"])
        assert np.abs(res[i]["sigmas"] - np.exp(sig[i].numpy())).max() < 1e-5
    # per-image solver calls of the SA engine (SA/src/data/speed/speed_dataset.py:399): answered from the batched solve;
    # seeded random heads emit one label, i.e. fewer than four keypoints -> the reference's IndexError contract
    solver = build_sigma_solver(model, post)
    for i in range(3):
        hit = post.pose_cache[id(res[i]["points"])]
        if hit[3] == 0:
            q, t = solver(res[i]["points"], res[i]["logits"], res[i]["sigmas"])
            assert np.array_equal(q, hit[1]) and np.array_equal(t, hit[2])
        else:
            with pytest.raises(IndexError):
                solver(res[i]["points"], res[i]["logits"], res[i]["sigmas"])
    # ... and without the cache (a fresh array): the per-image kernel call on post-processed inputs agrees
    d = synth.make_predictions(4, Q=cfg.num_queries, seed=9, with_sigma=True)
    r = model.engine.assign_pnp(torch.from_numpy(d["logits"]).cuda(), torch.from_numpy(d["points"]).cuda(),
                                torch.from_numpy(d["boxes"]).cuda(), log_sigma=torch.from_numpy(d["logsig"]).cuda(),
                                reproj=25.0, weighted=True, want_post=True)
    for i in range(4):
        if int(r["status"][i]) != 0:
            continue
        q, t = solver(r["points_px"][i].cpu().numpy().copy(), r["probs"][i].cpu().numpy(), r["sigmas"][i].cpu().numpy())
        assert np.abs(q - r["quat"][i].cpu().numpy()).max() < 1e-6 and np.abs(t - r["tvec"][i].cpu().numpy()).max() < 1e-5
    model.engine.close()


def test_sa_plain_tf32_trunk_opt_in(lib, cuda_dev, case, monkeypatch):
    """SPE_SA_X3=0: backbone + encoder on plain TF32 products (the decoder side stays 3xTF32).  States what that mode
    delivers against the live reference: ~1e-3 relative noise on the encoder memory, keypoints ~0.1 px rms / below 1 px
    worst case at S = 1748 -- outside the 0.5 px bar at the largest crops, hence not the default."""
    cfg, sd, x, g = case
    monkeypatch.setenv("SPE_SA_X3", "0")
    e = Engine(input_size=cfg.input_size, num_queries=cfg.num_queries, enc_layers=1, dec_layers=cfg.dec_layers,
               dim_feedforward=cfg.dec_ff, backbone="rtdetr_r50vd", precision="tf32", has_sigma=True, max_batch=4)
    try:
        e.load_state_dict(sd)
        out = e.forward_sa(x.cuda(), topk_override=torch.from_numpy(g["topk"]).cuda())
        torch.cuda.synchronize()
        d = (out["pred_pts"].cpu().numpy() - g["pred_pts"]) * 1748
        print(f"SA plain-TF32 trunk: keypoints max {np.abs(d).max():.3f} px, rms {np.sqrt((d ** 2).mean()):.3f} px at S=1748")
        assert np.abs(d).max() < 1.0 and np.sqrt((d ** 2).mean()) < 0.2
        assert np.abs(out["pred_logits"].cpu().numpy() - g["pred_logits"]).max() < 2e-2
    finally:
        e.close()


def test_sa_whole_path_through_pipeline_slots(eng, case):
    """crop (256 x 256) -> SA predictor -> sigma-weighted assignment + PnP through spe_submit_batch_dev / spe_collect_batch_host on
    two slots equals the same three stages called one by one (the pipeline is the same kernels on a slot's own buffers)."""
    cfg, sd, x, g = case
    B = 6
    det = synth.load_detector_boxes()[:B]
    frames = torch.from_numpy(synth.make_frames(B, det, seed=3)).cuda()
    boxes = torch.from_numpy(eng.clip_boxes(det)).cuda()
    imgs = eng.crop_resize_norm(frames, boxes)
    assert tuple(imgs.shape) == (B, 3, cfg.input_size, cfg.input_size)
    out = eng.forward_sa(imgs, want_aux=False)
    r = eng.assign_pnp(out["pred_logits"], out["pred_pts"], boxes, log_sigma=out["pred_sigmas"], reproj=25.0, weighted=True)
    torch.cuda.synchronize()
    for slot in (0, 1):
        eng.submit_batch_dev(slot, frames, boxes, reproj=25.0, weighted=True)
    for slot in (0, 1):
        got = eng.collect_batch_host(slot)
        assert np.array_equal(got["status"], r["status"].cpu().numpy())
        assert np.allclose(got["quat"], r["quat"].cpu().numpy(), atol=1e-9) and np.allclose(got["tvec"], r["tvec"].cpu().numpy(), atol=1e-9)


@pytest.mark.parametrize("R,B", [(224, 3), (160, 1)])
def test_sa_other_input_sizes(lib, cuda_dev, R, B):
    """Input sizes other than the recipe's 256: odd feature maps (224 -> 28 / 14 / 7, 160 -> 20 / 10 / 5) through every
    convolution path, the bicubic x0.5 and nearest x2 kernels, and a batch of one -- against the oracle."""
    cfg = sa_model_ref.SaCfg(input_size=R)
    sd = synth.make_sa_state_dict(cfg, seed=0)
    x = torch.randn(B, 3, R, R, generator=torch.Generator().manual_seed(R))
    taps = {}
    ref = sa_model_ref.forward(sd, cfg, x, taps)
    e = Engine(input_size=R, num_queries=cfg.num_queries, enc_layers=1, dec_layers=cfg.dec_layers, dim_feedforward=cfg.dec_ff,
               backbone="rtdetr_r50vd", precision="tf32", has_sigma=True, max_batch=B)
    try:
        e.load_state_dict(sd)
        o = e.forward_sa(x.cuda(), topk_override=taps["topk"].to(torch.int32).cuda())
        torch.cuda.synchronize()
        assert (o["pred_pts"].cpu() - ref["pred_pts"]).abs().max().item() <= PTS_TOL
        assert (o["pred_logits"].cpu() - ref["pred_logits"]).abs().max().item() <= LOGIT_TOL
        assert (o["pred_sigmas"].cpu() - ref["pred_sigmas"]).abs().max().item() <= LOGIT_TOL
    finally:
        e.close()
    with pytest.raises(Exception, match="256"):
        Engine(input_size=512, num_queries=30, enc_layers=1, dec_layers=3, dim_feedforward=1024, backbone="rtdetr_r50vd",
               precision="tf32", has_sigma=True, max_batch=1)


def test_sa_r18vd_recipe_vs_live_reference_golden(lib, cuda_dev):
    """rtdetr_r18vd_6x_speed_kl_*.yml: PResNet depth 18 (BasicBlocks: 3x3 + 3x3 with the residual in the second
    convolution's epilogue, 128 / 256 / 512-channel pyramid) behind the same encoder / decoder -- the library reads the
    depth off the checkpoint's tensor names.  Against the live model's outputs (its own top-k selection handed in)."""
    from oracle.make_golden import SA_R18_CASE
    g = np.load(os.path.join(synth.GOLDEN_DIR, "sa_r18_model_golden.npz"))
    cfg = sa_model_ref.SaCfg(depth=18)
    sd = synth.make_sa_state_dict(cfg, seed=SA_R18_CASE["weights_seed"])
    assert synth.weights_checksum(sd) == str(g["weights_sha256"])
    x = model_inputs(SA_R18_CASE["batch"], cfg.input_size, SA_R18_CASE["seed"])
    model, post = build_sa_model(max_batch=4, depth=18)
    model.load_state_dict(sd, strict=True)
    model.to("cuda")
    model(x[:2].cuda())                       # creates the engine
    out = model.engine.forward_sa(x.cuda(), topk_override=torch.from_numpy(g["topk"]).cuda())
    torch.cuda.synchronize()
    d_p = np.abs(out["pred_pts"].cpu().numpy() - g["pred_pts"]).max()
    d_l = np.abs(out["pred_logits"].cpu().numpy() - g["pred_logits"]).max()
    d_s = np.abs(out["pred_sigmas"].cpu().numpy() - g["pred_sigmas"]).max()
    a_p = np.abs(torch.stack([a["pred_pts"] for a in out["aux_outputs"]]).cpu().numpy() - g["aux_pts"]).max()
    print(f"SA r18vd vs live reference: keypoints {d_p * 1748:.3f} px at S=1748 (aux {a_p * 1748:.3f}), logits {d_l:.1e}, log-sigma {d_s:.1e}")
    assert d_p <= PTS_TOL and a_p <= PTS_TOL and d_l <= LOGIT_TOL and d_s <= LOGIT_TOL
    own = model.engine.forward_sa(x.cuda())
    assert min(len(set(own["topk_ind"][b].tolist()) & set(g["topk"][b].tolist())) for b in range(x.shape[0])) >= cfg.num_queries - 2
    model.engine.close()
