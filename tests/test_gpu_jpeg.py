"""GPU JPEG decode (spe_jpeg_decode_batch) against PIL -- the reference's own decoder call
(RV/datasets/speed.py:116: Image.open(img_path).convert('RGB')) -- bit for bit."""
import io

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import crop_ref, jpeg_ref, synth
from satellite_pose_estimation_b200 import Engine
from satellite_pose_estimation_b200._lib import SpeError

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(lib, cuda_dev):
    e = Engine(max_batch=16)
    yield e
    e.close()


def _jpeg(a, **kw):
    buf = io.BytesIO()
    Image.fromarray(a, "L").save(buf, "JPEG", **kw)
    return buf.getvalue()


def _pil(b):
    return np.asarray(Image.open(io.BytesIO(b)))


def test_speed_sized_frames_bit_exact_with_pil(eng):
    """1920 x 1200 frames of the benchmark's synthetic set at three qualities, default and optimised Huffman tables,
    with and without restart intervals: every pixel equals PIL's; .convert('RGB') is that plane three times."""
    det = synth.load_detector_boxes()[:6]
    frames = synth.make_frames(6, det, seed=2)
    opts = [dict(quality=75), dict(quality=95), dict(quality=50, optimize=True), dict(quality=90, restart_marker_rows=2),
            dict(quality=75, restart_marker_blocks=100), dict(quality=100)]
    files = [_jpeg(frames[i], **opts[i]) for i in range(6)]
    out = eng.decode_jpeg(files)
    torch.cuda.synchronize()
    assert out.shape == (6, 1200, 1920) and out.dtype == torch.uint8
    for i in range(6):
        ref = _pil(files[i])
        assert np.array_equal(out[i].cpu().numpy(), ref), f"file {i} ({opts[i]})"
        rgb = np.asarray(Image.open(io.BytesIO(files[i])).convert("RGB"))
        assert all(np.array_equal(rgb[..., c], ref) for c in range(3))
    print(f"compressed {sum(map(len, files)) / 1e6:.2f} MB for {6 * 1200 * 1920 / 1e6:.1f} MB of frames")


def test_white_noise_and_flat_frames(eng):
    """white noise at quality 98 drives the coefficients to their extremes (long Huffman codes, the range-limit
    table's saturated part); a flat frame is DC-only blocks (the inverse DCT's shortcut form)"""
    rng = np.random.default_rng(5)
    files = [_jpeg(rng.integers(0, 256, (1200, 1920), dtype=np.uint8), quality=98),
             _jpeg(np.full((1200, 1920), 137, np.uint8), quality=80),
             _jpeg((rng.integers(0, 2, (1200, 1920)) * 255).astype(np.uint8), quality=100)]
    out = eng.decode_jpeg(files)
    for i, f in enumerate(files):
        assert np.array_equal(out[i].cpu().numpy(), _pil(f)), i


@pytest.mark.parametrize("h,w", [(8, 8), (117, 203), (33, 9), (1200, 1928)])
def test_other_sizes_partial_blocks_and_pitched_output(eng, h, w):
    rng = np.random.default_rng(h * 1000 + w)
    y, x = np.mgrid[0:h, 0:w]
    a = np.clip(90 + 60 * np.sin(x / 5.0) * np.cos(y / 9.0) + rng.normal(0, 20, (h, w)), 0, 255).astype(np.uint8)
    files = [_jpeg(a, quality=q) for q in (60, 92)]
    out = eng.decode_jpeg(files)
    for i, f in enumerate(files):
        assert np.array_equal(out[i].cpu().numpy(), _pil(f))
        if h * w < 50000:
            assert np.array_equal(jpeg_ref.decode(f), _pil(f))      # the CPU restatement agrees too
    # a view into a wider buffer (row stride != width, odd alignment: the byte-store path)
    big = torch.zeros(2, h, w + 13, dtype=torch.uint8, device="cuda")
    eng.decode_jpeg(files, out=big[:, :, 5:5 + w])
    assert np.array_equal(big[0, :, 5:5 + w].cpu().numpy(), _pil(files[0]))
    assert int(big[:, :, :5].max()) == 0 and int(big[:, :, 5 + w:].max()) == 0


def test_decode_feeds_the_crop_stage(eng):
    """file -> frame -> crop on the device equals PIL decode -> the oracle's crop (the reference's data path)"""
    det = synth.load_detector_boxes()[40:44]
    frames = synth.make_frames(4, det, seed=9)
    files = [_jpeg(frames[i], quality=90) for i in range(4)]
    fd = eng.decode_jpeg(files)
    clip = eng.clip_boxes(det)
    out = eng.crop_resize_norm(fd, torch.from_numpy(clip).cuda()).cpu()
    for i in range(4):
        ref, rclip = crop_ref.crop_resize_normalize(_pil(files[i]), det[i], 224)
        assert np.array_equal(rclip, clip[i])
        d = (out[i] - ref).abs()
        assert d.max() <= 1.0 / 255 / 0.224 * 1.001 and (d > 1e-6).float().mean() <= 2e-4


def test_unsupported_files_are_refused(eng):
    a = synth.make_frames(1, synth.load_detector_boxes()[:1], seed=1)[0][:64, :64].copy()
    buf = io.BytesIO(); Image.fromarray(a, "L").save(buf, "JPEG", progressive=True)
    with pytest.raises(SpeError, match="progressive"):
        eng.decode_jpeg([buf.getvalue()])
    buf = io.BytesIO(); Image.fromarray(np.stack([a] * 3, -1), "RGB").save(buf, "JPEG")
    with pytest.raises(SpeError, match="3 components"):
        eng.decode_jpeg([buf.getvalue()])
    with pytest.raises(SpeError, match="not a JPEG"):
        eng.decode_jpeg([b"\x89PNG\r\n\x1a\n" + b"\0" * 64])
    good = _jpeg(a, quality=80)
    with pytest.raises(SpeError, match="file 1"):           # mixed sizes in one batch
        eng.decode_jpeg([good, _jpeg(a[:32], quality=80)])
    assert np.array_equal(eng.decode_jpeg([good])[0].cpu().numpy(), _pil(good))   # the ctx is still usable


def test_image_set_from_jpeg_files_equals_pil_decoded_frames(lib, cuda_dev):
    """config 5 with the reference's real input: the runner fed JPEG files (GPU decode of the whole shard, frames stay
    in HBM) returns, per filename, exactly what it returns when fed the frames PIL decodes from the same files."""
    from oracle import model_ref
    from satellite_pose_estimation_b200.submission import run_image_set
    cfg = model_ref.ModelCfg()
    e = Engine(max_batch=32)
    e.load_state_dict(synth.make_state_dict(cfg, seed=0, spread_labels=True))
    n = 70
    det = synth.load_detector_boxes()[:n]
    base = synth.make_frames(6, det, seed=4)
    frames = [np.roll(base[i % 6], 13 * i, axis=1) for i in range(n)]
    files = [_jpeg(f, quality=85) for f in frames]
    decoded = np.stack([_pil(f) for f in files])
    names = [f"img{i:06d}.jpg" for i in range(n)]
    a = run_image_set(e, None, det, names, batch_size=32, slots=3, gather=False, jpeg_files=files)
    b = run_image_set(e, lambda i0, i1: decoded[i0:i1], det, names, batch_size=32, slots=3, gather=False)
    assert a == b and len(a) == n
    assert sum(1 for v in a.values() if v["status"] == 0) >= 0.9 * n
    e.close()
