"""Fused crop/resize/normalise kernel against the oracle (= the reference's own cv2 call chain)."""
import os

import numpy as np
import pytest
import torch

from oracle import crop_ref, synth
from oracle.make_golden import crop_case_boxes
from satellite_pose_estimation_b200 import Engine

pytestmark = pytest.mark.gpu
LSB = 1.0 / 255 / 0.224   # one uint8 step after normalisation (largest channel scale ~ 1/(255*0.224))


@pytest.fixture(scope="module")
def eng(lib, cuda_dev):
    e = Engine(max_batch=32)
    yield e
    e.close()


def _run(eng, frames, det, R=224):
    clip = eng.clip_boxes(det)
    out = eng.crop_resize_norm(torch.from_numpy(frames).cuda(), torch.from_numpy(clip).cuda(), R=R)
    torch.cuda.synchronize()
    return clip, out.cpu()


def test_crop_matches_golden_and_oracle(eng):
    g = np.load(os.path.join(synth.GOLDEN_DIR, "crop_golden.npz"))
    idx, det = crop_case_boxes()
    frames = synth.make_frames(len(idx), det, seed=0)
    clip, out = _run(eng, frames, det)
    assert np.array_equal(clip, g["clip_boxes"])                                   # bit-exact integer boxes
    bad = tot = 0
    for i in range(len(idx)):
        gold = crop_ref.normalize_u8(np.repeat(g["crops_u8"][i][:, :, None], 3, 2))   # reference output tensor
        d = (out[i] - gold).abs()
        assert d.max() <= LSB * 1.001, f"case {i}: more than one uint8 step off"
        bad += int((d > 1e-6).sum()); tot += d.numel()
    assert bad / tot <= 1e-4, f"{bad}/{tot} values differ from cv2 (bar: 0.01 %)"


def test_crop_random_boxes_incl_out_of_frame(eng):
    rng = np.random.default_rng(5)
    det = []
    for _ in range(24):
        w, h = rng.uniform(40, 1500, 2)
        cx, cy = rng.uniform(-100, 2020), rng.uniform(-100, 1300)
        det.append([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
    det = np.asarray(det)
    frames = synth.make_frames(24, np.clip(det, 0, 1900), seed=9)
    clip, out = _run(eng, frames, det)
    bad = tot = 0
    for i in range(24):
        ref, rclip = crop_ref.crop_resize_normalize(frames[i], det[i], 224)
        assert np.array_equal(clip[i], rclip)
        d = (out[i] - ref).abs()
        assert d.max() <= LSB * 1.001
        bad += int((d > 1e-6).sum()); tot += d.numel()
    assert bad / tot <= 1e-4


def test_crop_properties(eng):
    # constant frame -> constant crop wherever the box is inside the frame; fully outside -> exact -mean/std
    frame = np.full((2, 1200, 1920), 137, np.uint8)
    det = np.array([[500, 300, 900, 600], [-900, -900, -500, -600]], dtype=np.float64)
    _, out = _run(eng, frame, det, R=96)
    inside = crop_ref.normalize_u8(np.full((96, 96, 3), 137, np.uint8))
    black = crop_ref.normalize_u8(np.zeros((96, 96, 3), np.uint8))
    assert torch.equal(out[0], inside) and torch.equal(out[1], black)
    # pitched input: a view with a row stride larger than the width gives the same result
    big = torch.zeros(1, 1200, 2048, dtype=torch.uint8, device="cuda")
    f = synth.make_frames(1, det[:1], seed=3)
    big[0, :, :1920] = torch.from_numpy(f[0]).cuda()
    clip = torch.from_numpy(eng.clip_boxes(det[:1])).cuda()
    a = eng.crop_resize_norm(big[:, :, :1920], clip)
    b = eng.crop_resize_norm(torch.from_numpy(f).cuda(), clip)
    assert torch.equal(a, b)


def test_staged_and_gather_kernels_are_bit_identical(eng):
    """Two kernels sit behind spe_crop_resize_norm: the shared-memory staged one (frame rows addressable in 16-byte units:
    the normal case) and the byte-gather one (any layout).  A frame whose row stride is not a multiple of 16 takes the
    gather kernel; the same pixels in the normal layout take the staged kernel; the results must be the same bits --
    for boxes inside, across every border of, and outside the frame."""
    rng = np.random.default_rng(11)
    det = synth.load_detector_boxes()[200:232].copy()
    det[:8, [0, 2]] -= 700; det[8:12, [1, 3]] -= 500; det[12:16, [0, 2]] += 900; det[16:20, [1, 3]] += 600
    det[20] = [-900, -900, -500, -600]
    frames = rng.integers(0, 256, size=(32, 1200, 1920), dtype=np.uint8)      # white noise: every tap matters
    fd = torch.from_numpy(frames).cuda()
    odd = torch.zeros(32, 1200, 1929, dtype=torch.uint8, device="cuda")
    odd[:, :, :1920] = fd
    clip = torch.from_numpy(eng.clip_boxes(det)).cuda()
    staged = eng.crop_resize_norm(fd, clip)
    gather = eng.crop_resize_norm(odd[:, :, :1920], clip)
    assert torch.equal(staged, gather)
    ref, _ = crop_ref.crop_resize_normalize(frames[3], det[3], 224)
    assert (staged[3].cpu() - ref).abs().max() <= LSB * 1.001


def test_crop_other_input_size(eng):
    det = synth.load_detector_boxes()[100:104]
    frames = synth.make_frames(4, det, seed=4)
    clip, out = _run(eng, frames, det, R=256)
    for i in range(4):
        ref, _ = crop_ref.crop_resize_normalize(frames[i], det[i], 256)
        d = (out[i] - ref).abs()
        assert d.max() <= LSB * 1.001 and (d > 1e-6).float().mean() <= 2e-4


def test_eval_path_crop_matches_reference_golden(eng):
    """main.py --eval crop (non-square PIL rectangle squashed to R x R, RV/datasets/speed.py:219-236) against the
    reference's own SpeedTrain(train=False) output."""
    g = np.load(os.path.join(synth.GOLDEN_DIR, "crop_eval_golden.npz"))
    det = g["det_boxes"]
    frames = synth.make_frames(len(det), det, seed=int(g["frame_seed"]))
    fbox, ibox = eng.clip_boxes_val(det)
    assert np.array_equal(fbox, g["float_boxes"])
    out = eng.crop_resize_norm(torch.from_numpy(frames).cuda(), torch.from_numpy(ibox).cuda(), R=int(g["input_size"]))
    torch.cuda.synchronize()
    out = out.cpu()
    bad = tot = 0
    for i in range(len(det)):
        gold = crop_ref.normalize_u8(np.repeat(g["crops_u8"][i][:, :, None], 3, 2))
        d = (out[i] - gold).abs()
        assert d.max() <= LSB * 1.001, f"case {i}: more than one uint8 step off"
        bad += int((d > 1e-6).sum()); tot += d.numel()
    assert bad / tot <= 1e-4, f"{bad}/{tot} values differ from cv2"
    assert any(ibox[i, 2] - ibox[i, 0] != ibox[i, 3] - ibox[i, 1] for i in range(len(det)))   # non-square cases present
