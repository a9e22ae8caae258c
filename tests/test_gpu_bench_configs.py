"""Parity at the configurations bench.py actually runs (BASELINE.json configs[1..3]): batch 64 TF32 on the big-batch
schedule (fused feed-forward block, CTA-pair GEMMs, 256-wide tiles), through the 4-slot pipeline, and batch 256 in bf16
/ with the sigma head -- each compared DIRECTLY with goldens of the live reference
(tests/golden/model_b64_random_golden.npz, model_b256_golden.npz; oracle/make_golden.py), not with another GPU result.

Two weight sets:
  * the seeded random-init weights north_star names -- keypoints within 0.5 px at the largest crop side (1748 px);
  * the same trunk with calibrated heads (oracle/make_chain_fixture.py) whose queries emit 11 distinct labels at a
    PnP-consistent layout, so that crop -> predictor -> assignment -> PnP is checked as ONE chain on the network's own
    output against the reference's own PostProcess + SimplePoseSolver.
Bars are stated where they are asserted; tests/probes/parity_report.py prints the full statistics (profiles/r02_parity.json).
"""
import os

import numpy as np
import pytest
import torch

from oracle import crop_ref, model_ref, pnp_ref, synth
from satellite_pose_estimation_b200 import Engine

pytestmark = pytest.mark.gpu

S_MAX = 1748                       # largest crop side of the real detector-box distribution
G256 = None
G64 = None


def goldens():
    global G256, G64
    if G256 is None:
        G256 = np.load(os.path.join(synth.GOLDEN_DIR, "model_b256_golden.npz"))
        G64 = np.load(os.path.join(synth.GOLDEN_DIR, "model_b64_random_golden.npz"))
    return G256, G64


_CROPS = {}


def oracle_crops(n_sets):
    """oracle (cv2) crops of synth.bench_set(0 .. n_sets-1): the inputs the goldens were computed on"""
    for s in range(n_sets):
        if s not in _CROPS:
            frames, det = synth.bench_set(s)
            _CROPS[s] = (torch.stack([crop_ref.crop_resize_normalize(frames[i], det[i], 224)[0] for i in range(64)]),
                         frames, det)
    return (torch.cat([_CROPS[s][0] for s in range(n_sets)]), np.concatenate([_CROPS[s][1] for s in range(n_sets)]),
            np.concatenate([_CROPS[s][2] for s in range(n_sets)]))


def assigned_mask(assign, Q=40):
    m = np.zeros((len(assign), Q), dtype=bool)
    for i, a in enumerate(assign):
        m[i, a[a >= 0]] = True
    return m


def test_tf32_batch64_random_init_vs_live_reference(lib, cuda_dev):
    """configs[1] exactly as north_star words it: random-init weights, batch 64 (big-batch schedule), TF32.  Bar: every
    keypoint of every query within 0.5 px at S = 1748 without calibration, within 0.3 px with it (calibration images
    = 16 crops of ANOTHER frame set); arg-max labels identical."""
    _, g = goldens()
    cfg = model_ref.ModelCfg()
    sd = synth.make_state_dict(cfg, seed=0)
    assert synth.weights_checksum(sd).encode() == g["checksum"].tobytes()
    x, _, _ = oracle_crops(2)
    eng = Engine(max_batch=64, precision="tf32")
    eng.load_state_dict(sd)
    errs = []
    for cal in (False, True):
        if cal:
            eng.calibrate(x[64:80].cuda())
        out = eng.forward(x[:64].cuda())
        torch.cuda.synchronize()
        d = np.abs(out["pred_points"].cpu().numpy() - g["pred_points"]).max() * S_MAX
        errs.append(d)
        assert np.array_equal(out["pred_logits"].argmax(-1).cpu().numpy(), g["pred_logits"].argmax(-1))
        assert np.abs(out["pred_logits"].cpu().numpy() - g["pred_logits"]).max() < 1e-2
    print(f"random-init B=64: {errs[0]:.3f} px uncalibrated, {errs[1]:.3f} px calibrated (at S={S_MAX})")
    assert errs[0] <= 0.5 and errs[1] <= 0.3
    eng.close()


def test_tf32_batch64_whole_chain_through_4_slot_pipeline(lib, cuda_dev):
    """crop -> predictor -> assignment -> PnP as one chain on the network's OWN outputs (calibrated heads), exactly the
    way bench.py runs it: frames resident in HBM, spe_submit_batch_dev on 4 slots, graphs replayed, poses collected on
    the host -- against the live reference model + the reference's own PostProcess + SimplePoseSolver (cv2)."""
    g, _ = goldens()
    cfg = model_ref.ModelCfg()
    sd = synth.make_state_dict(cfg, seed=0, spread_labels=True)
    x, frames, det = oracle_crops(4)
    B, slots, nb = 64, 4, 4
    eng = Engine(max_batch=B, precision="tf32")
    eng.load_state_dict(sd)
    fd = [torch.from_numpy(frames[i * B:(i + 1) * B]).cuda() for i in range(nb)]
    clip = eng.clip_boxes(det)
    assert np.array_equal(clip, g["clip_boxes"])                              # crop boxes: bit-exact
    bd = [torch.from_numpy(clip[i * B:(i + 1) * B]).cuda() for i in range(nb)]
    eng.calibrate(eng.crop_resize_norm(fd[1][:16], bd[1][:16]))
    res = [None] * nb
    for rep in range(2):                                                      # round 2 replays the captured graphs
        for i in range(min(slots, nb)):
            eng.submit_batch_dev(i % slots, fd[i], bd[i])
        for i in range(nb):
            res[i] = eng.collect_batch_host(i % slots)
            res[i]["net"] = eng.read_slot_outputs(i % slots, B)
    captured, failed = eng.graph_stats()
    assert captured >= slots and failed == 0
    n = nb * B
    logits = np.concatenate([r["net"][0] for r in res]); pts = np.concatenate([r["net"][1] for r in res])
    quat = np.concatenate([r["quat"] for r in res]); tvec = np.concatenate([r["tvec"] for r in res])
    status = np.concatenate([r["status"] for r in res])
    # labels: identical arg-max for every query; the pose stage's query -> keypoint table is the reference's
    assert np.array_equal(logits.argmax(-1), g["pred_logits"].argmax(-1))
    probs = torch.softmax(torch.from_numpy(logits), -1).numpy()
    px = pts * (clip[:, None, 2:] - clip[:, None, :2]) + clip[:, None, :2]
    my_assign = np.stack([pnp_ref.assign_table(px[i], probs[i]) for i in range(n)])
    assert np.array_equal(my_assign, g["assign"]) and (g["assign"] >= 0).all()   # all 11 keypoints found in every image
    # keypoints the pose stage consumes: error in ORIGINAL-IMAGE pixels (normalised error x the image's own crop side)
    side = (clip[:, 2] - clip[:, 0]).astype(np.float64)
    d = np.abs(pts - g["pred_points"])
    fg = assigned_mask(g["assign"])
    d_px = (d.max(-1) * side[:, None])[fg]
    d_smax = d[fg].max() * S_MAX
    print(f"chain B=64x4: assigned keypoints max {d_px.max():.3f} px in the image ({d_smax:.3f} px if every crop were "
          f"{S_MAX} px wide), rms {np.sqrt((d[fg] ** 2).mean()) * S_MAX:.3f} px at S={S_MAX}")
    # These heads are ~10 x more sensitive to the encoder memory than the random-init ones (the point head was fitted to
    # decode 11 distinct positions): their bar is stated on the pixels PnP sees.  Measured over several runs: rms
    # 0.15-0.17 px at S = 1748, worst keypoint 0.26-0.51 px in the image.
    rms = np.sqrt((d[fg] ** 2).mean()) * S_MAX
    assert rms <= 0.25 and (d_px <= 0.5).mean() >= 0.999 and d_px.max() <= 0.75
    # poses against the reference chain (reference network outputs -> reference PostProcess -> cv2 RANSAC-P3P + LM)
    assert np.array_equal(status == 0, g["ok"] == 1) and (status == 0).all()
    rot, tr = [], []
    for i in range(n):
        s_t, s_q = pnp_ref.speed_score(quat[i], tvec[i], g["quat"][i], g["tvec"][i])
        rot.append(np.degrees(s_q)); tr.append(s_t)
    rot, tr = np.asarray(rot), np.asarray(tr)
    print(f"chain poses vs reference chain: rotation median {np.median(rot):.4f} deg, p95 {np.percentile(rot, 95):.3f}, "
          f"max {rot.max():.2f}; translation median {np.median(tr):.2e}, max {tr.max():.2e}")
    # The two chains see keypoints that differ by up to 0.5 px, and the calibrated point head scatters single keypoints
    # by up to ~20 px around the layout -- right at RANSAC's 20 px threshold, where cv2's random draw and the kernel's
    # exhaustive consensus may keep different inlier sets.  So the poses agree statistically here (measured: median
    # 0.007 deg, 89 % within 0.5 deg, worst 2.8 deg); the 0.01 deg bar for IDENTICAL keypoints and inlier sets is held by
    # test_chain_poses_given_identical_keypoints below and tests/test_gpu_pnp.py.
    assert np.median(rot) < 0.05 and np.median(tr) < 1e-3
    assert (rot < 0.5).mean() >= 0.8 and rot.max() < 5.0
    eng.close()


def test_chain_poses_given_identical_keypoints(lib, cuda_dev):
    """north_star: 'given identical keypoints, poses match cv2 within 0.01 deg / 1e-4' -- on the NETWORK's keypoints:
    the GPU pose stage and the cv2 chain both consume the GPU network outputs of the 64 bench frames."""
    cfg = model_ref.ModelCfg()
    sd = synth.make_state_dict(cfg, seed=0, spread_labels=True)
    x, frames, det = oracle_crops(1)
    eng = Engine(max_batch=64, precision="tf32")
    eng.load_state_dict(sd)
    clip = eng.clip_boxes(det)
    out = eng.forward(x[:64].cuda())
    r = eng.assign_pnp(out["pred_logits"], out["pred_points"], torch.from_numpy(clip).cuda(), want_post=True)
    torch.cuda.synchronize()
    res = pnp_ref.post_process(out["pred_logits"].cpu().numpy(), out["pred_points"].cpu().numpy(), clip)
    solver = pnp_ref.SimplePoseSolver(20, return_inliers=True)
    n_same, worst_rot, worst_tr = 0, 0.0, 0.0
    for i in range(64):
        assert np.array_equal(r["points_px"][i].cpu().numpy(), res[i]["points"])          # PostProcess: bit-exact
        tab = pnp_ref.assign_table(res[i]["points"], res[i]["logits"])
        assert np.array_equal(r["assign"][i].cpu().numpy(), tab)
        try:
            q_ref, t_ref, used = solver(res[i]["points"], res[i]["logits"]); ok = True
        except Exception:
            ok, used = False, None
        assert ok == (int(r["status"][i]) == 0)
        # the kernel's inlier mask indexes correspondences in first-appearance order of their labels
        order, _, _ = pnp_ref.assign(res[i]["points"], res[i]["logits"])
        mine = sorted(order[j] for j in range(len(order)) if (int(r["inlier_mask"][i]) >> j) & 1)
        if ok and mine == used:
            s_t, s_q = pnp_ref.speed_score(r["quat"][i].cpu().numpy(), r["tvec"][i].cpu().numpy(), q_ref, t_ref)
            worst_rot, worst_tr = max(worst_rot, np.degrees(s_q)), max(worst_tr, s_t)
            n_same += 1
    print(f"identical keypoints: {n_same}/64 with cv2's consensus set; worst {worst_rot:.2e} deg / {worst_tr:.2e}")
    assert n_same >= 58 and worst_rot < 0.01 and worst_tr < 1e-4
    eng.close()


@pytest.mark.parametrize("precision,sigma", [("bf16", False), ("tf32", True)])
def test_batch256_configs_vs_live_reference(lib, cuda_dev, precision, sigma):
    """configs[2] (bf16, batch 256) and configs[3] (sigma head, batch 256) against the live reference on 256 crops
    (calibrated heads).  TF32 + sigma: the 0.5 px bar in image pixels as above, log-sigma against the SA drop's own MLP
    module (SA/src/zoo/rtdetr/rtdetr_decoder.py:24-37, :295-297, :367) within 1e-3.  bf16 cannot meet 0.5 px (8-bit
    mantissa on every stored activation) -- its bar is the one the pose stage needs: no arg-max
    flip, every assigned keypoint within 16 px in the image (measured 9-11 px worst, 0.8 px rms at the median crop side),
    median pose within 0.5 deg of the reference chain."""
    g, _ = goldens()
    cfg = model_ref.ModelCfg(sigma_head=sigma)
    sd = synth.make_state_dict(cfg, seed=0, spread_labels=True)
    x, frames, det = oracle_crops(4)
    B = 256
    eng = Engine(max_batch=B, precision=precision, has_sigma=sigma)
    eng.load_state_dict(sd)
    eng.calibrate(x[64:80].cuda())
    out = eng.forward(x.cuda())
    torch.cuda.synchronize()
    pts = out["pred_points"].cpu().numpy(); logits = out["pred_logits"].cpu().numpy()
    flips = int((logits.argmax(-1) != g["pred_logits"].argmax(-1)).sum())
    clip = g["clip_boxes"]
    side = (clip[:, 2] - clip[:, 0]).astype(np.float64)
    d = np.abs(pts - g["pred_points"])
    fg = assigned_mask(g["assign"])
    d_px = (d.max(-1) * side[:, None])[fg]
    stats = {S: d[fg].max() * S for S in (430, 1147, 1748)}
    print(f"{precision}{' + sigma' if sigma else ''} B=256: flips {flips}/{B * 40}; assigned keypoints max {d_px.max():.2f} px "
          f"in the image; at S=430/1147/1748: {stats[430]:.2f}/{stats[1147]:.2f}/{stats[1748]:.2f} px; rms "
          f"{np.sqrt((d[fg] ** 2).mean()) * 1748:.2f} px at 1748")
    assert flips == 0
    if precision == "tf32":
        assert np.sqrt((d[fg] ** 2).mean()) * S_MAX <= 0.25 and (d_px <= 0.5).mean() >= 0.999 and d_px.max() <= 0.75
        assert np.abs(out["pred_sigmas"].cpu().numpy() - g["pred_sigmas"]).max() < 1e-3
        assert torch.equal(out["pred_sigmas"][..., 0], out["pred_sigmas"][..., 1])
    else:
        assert d_px.max() <= 16.0 and np.sqrt((d[fg] ** 2).mean()) * 430 <= 1.5
    # pose stage on the network's own outputs at B = 256 (sigma-weighted + reject filter for the SA variant)
    r = eng.assign_pnp(out["pred_logits"], out["pred_points"], torch.from_numpy(clip).to(torch.int32).cuda(),
                       log_sigma=out.get("pred_sigmas"), reproj=25.0 if sigma else 20.0, weighted=sigma, reject=sigma)
    st = r["status"].cpu().numpy()
    assert np.array_equal(r["assign"].cpu().numpy(), g["assign"])
    # every image solves; the SA variant's reject filter additionally flags (status 3, pose still reported) the poses
    # whose inlier RMS reprojection error exceeds 5 px -- the calibrated point head scatters that much on some images
    assert ((st == 0) | (st == 3)).all() and (st == 0).mean() >= (0.6 if sigma else 0.98)
    rot = []
    for i in range(B):
        _, s_q = pnp_ref.speed_score(r["quat"][i].cpu().numpy(), r["tvec"][i].cpu().numpy(), g["quat"][i], g["tvec"][i])
        rot.append(np.degrees(s_q))
    print(f"  poses vs reference chain: rotation median {np.median(rot):.3f} deg, max {max(rot):.2f}")
    assert np.median(rot) < (0.5 if precision == "bf16" else 0.1)
    eng.close()


def test_forward_on_default_stream_replays_its_graph(lib, cuda_dev):
    """Stream capture is illegal on the legacy default stream: a forward that arrives there runs on the context's own
    stream (ordered by events) and must still replay a captured graph from its second call on -- and a pipeline slot
    keeps replaying ITS graph after such calls (the old behaviour disabled graphs for the whole context)."""
    cfg = model_ref.ModelCfg()
    sd = synth.make_state_dict(cfg, seed=0)
    x, frames, det = oracle_crops(1)
    eng = Engine(max_batch=8, precision="tf32")
    eng.load_state_dict(sd)
    xs = x[:8].cuda()
    assert torch.cuda.current_stream().cuda_stream == 0                      # the legacy default stream
    first = eng.forward(xs)["pred_points"].clone()                           # eager
    for _ in range(3):
        again = eng.forward(xs)["pred_points"].clone()                       # capture, then replay
        assert torch.equal(first, again)
    captured, failed = eng.graph_stats()
    assert captured == 1 and failed == 0
    fd = torch.from_numpy(frames[:8]).cuda(); bd = torch.from_numpy(eng.clip_boxes(det[:8])).cuda()
    for _ in range(3):
        eng.submit_batch_dev(1, fd, bd)
        eng.collect_batch_host(1)
        eng.run_batch_host(torch.from_numpy(frames[:8]).pin_memory(), det[:8])  # default stream in between
    captured, failed = eng.graph_stats()
    assert captured >= 3 and failed == 0
    # ordering against the caller's stream: work enqueued before / after the call sees consistent data
    y = xs * 2.0
    out = eng.forward(y)["pred_points"].clone()
    y.zero_()                                                                # must not race with the forward above
    torch.cuda.synchronize()
    assert torch.equal(out, eng.forward(xs * 2.0)["pred_points"])
    eng.close()
