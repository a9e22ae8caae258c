"""The C-ABI library loads and exports every symbol include/spe.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import torch

from satellite_pose_estimation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "spe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spe_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    names = declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"include/spe.h declares {n} but libspe.so does not export it"
    assert set(names) == set(_lib.SYMBOLS), "ctypes prototypes out of sync with include/spe.h"
    slots = int(re.search(r"#define SPE_PIPELINE_SLOTS (\d+)", open(os.path.join(ROOT, "include", "spe.h")).read()).group(1))
    assert slots == _lib.PIPELINE_SLOTS


def test_only_abi_symbols_are_exported():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert exported and all(s.startswith("spe_") for s in exported), exported


def test_kernels_are_blackwell_native():
    """SASS evidence: tcgen05 MMA (UTC*MMA), TMEM loads (LDTM) and TMA (UTMALDG) are in the shipped binary."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, f"{mnemonic} missing from libspe.so SASS"
    # issue loops are warp-uniform (elect_one_sync): no per-instruction waterfall loops around UTCHMMA / UTMALDG.
    # The one place a waterfall is the intent: the staged crop kernel, where each of 32 lanes issues the bulk copy of
    # its OWN frame-row segment (32 different addresses, once per CTA).
    fn, waterfall = None, set()
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        elif "BRA.U.ANY" in line:
            waterfall.add(fn)
    assert all("crop_resize_norm_staged_kernel" in f for f in waterfall), waterfall
    # programmatic dependent launch: griddepcontrol.wait / .launch_dependents compiled into the step's kernels
    assert sass.count("ACQBULK") >= 20 and sass.count("PREEXIT") >= 20


def test_clip_boxes_bit_exact_on_every_reference_box(lib):
    """Row 1 of SURVEY.md section 8a: int64 crop boxes must be bit-exact (host-side f64 + truncation)."""
    from oracle import crop_ref, synth
    for name in ("wz_synt_test_boxes.npy", "wz_real_test_boxes.npy"):
        det = np.load(os.path.join(synth.GOLDEN_DIR, name))
        out = np.empty((len(det), 4), dtype=np.int32)
        assert lib.spe_clip_boxes(det.ctypes.data_as(C.c_void_p), len(det), out.ctypes.data_as(C.c_void_p)) == 0
        ref = np.stack([crop_ref.generate_clip_bbox(b) for b in det])
        assert (out == ref).all()
    neg = np.array([[-30.7, -12.2, 55.5, 40.1], [10.0, 10.0, 10.0, 10.0]])   # truncation toward zero, empty box
    out = np.empty((2, 4), dtype=np.int32)
    lib.spe_clip_boxes(neg.ctypes.data_as(C.c_void_p), 2, out.ctypes.data_as(C.c_void_p))
    assert (out == np.stack([crop_ref.generate_clip_bbox(b) for b in neg])).all()


def test_eval_path_boxes_bit_exact_on_every_reference_box(lib):
    """main.py --eval crop boxes (RV/datasets/speed.py:246-260): the float64 box and its PIL rounding
    (``int(round(v))``, half to even) for every detector box the reference ships."""
    from oracle import crop_ref, synth
    for name in ("wz_synt_test_boxes.npy", "wz_real_test_boxes.npy"):
        det = np.load(os.path.join(synth.GOLDEN_DIR, name))
        fb = np.empty((len(det), 4), dtype=np.float64); ib = np.empty((len(det), 4), dtype=np.int32)
        assert lib.spe_clip_boxes_val(det.ctypes.data_as(C.c_void_p), len(det), 1920, 1200,
                                      fb.ctypes.data_as(C.c_void_p), ib.ctypes.data_as(C.c_void_p)) == 0
        ref = np.stack([crop_ref.generate_clip_bbox_val(b, (1920, 1200)) for b in det])
        assert np.array_equal(fb, ref)
        assert np.array_equal(ib, np.asarray([[int(round(v)) for v in r] for r in ref]))
    half = np.array([[100.0, 100.0, 105.0, 105.0], [0.5, 0.5, 3.0, 3.0]])      # x.5 coordinates: ties go to even
    fb = np.empty((2, 4)); ib = np.empty((2, 4), dtype=np.int32)
    lib.spe_clip_boxes_val(half.ctypes.data_as(C.c_void_p), 2, 1920, 1200, fb.ctypes.data_as(C.c_void_p),
                           ib.ctypes.data_as(C.c_void_p))
    assert np.array_equal(ib, np.asarray([[int(round(v)) for v in r] for r in fb])) and ib[0].tolist() == [100, 100, 106, 106]


def test_no_cpu_fallback(lib):
    """Without a GPU the product path fails loudly instead of computing on the CPU."""
    if torch.cuda.is_available():
        return
    cfg = _lib.SpeConfig(224, 40, 4, 4, 256, 8, 2048, 0, 0, 0, 4)
    ctx = C.c_void_p()
    rc = lib.spe_create(C.byref(cfg), 0, C.byref(ctx))
    assert rc != 0 and not ctx.value
    assert b"no CPU fallback" in lib.spe_global_last_error()
    import pytest
    from satellite_pose_estimation_b200 import Engine
    with pytest.raises(RuntimeError):
        Engine()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "satellite_pose_estimation_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


def test_jpeg_header_walk_on_the_host():
    """spe_jpeg_info needs no device: sizes of baseline grayscale files, refusal of what the decode kernel does not take"""
    import io
    from PIL import Image
    lib = _lib.load()
    a = (np.arange(48 * 100).reshape(48, 100) % 251).astype(np.uint8)

    def info(**kw):
        buf = io.BytesIO()
        img = Image.fromarray(a, "L") if kw.pop("gray", True) else Image.fromarray(np.stack([a] * 3, -1), "RGB")
        img.save(buf, "JPEG", **kw)
        b = buf.getvalue()
        w, h = C.c_int(0), C.c_int(0)
        rc = lib.spe_jpeg_info(C.cast(C.create_string_buffer(b, len(b)), C.c_void_p), len(b), C.byref(w), C.byref(h))
        return rc, w.value, h.value, (lib.spe_global_last_error() or b"").decode()

    assert info(quality=80)[:3] == (0, 100, 48)
    assert info(quality=95, optimize=True, restart_marker_rows=1)[:3] == (0, 100, 48)
    rc, _, _, msg = info(progressive=True)
    assert rc == -1 and "progressive" in msg
    rc, _, _, msg = info(gray=False)
    assert rc == -1 and "3 components" in msg
    w, h = C.c_int(0), C.c_int(0)
    assert lib.spe_jpeg_info(C.cast(C.create_string_buffer(b"abcdefgh", 8), C.c_void_p), 8, C.byref(w), C.byref(h)) == -1


def test_sa_entry_points_validate_before_touching_the_device(lib):
    """The SA predictor's ABI (spe_config.backbone = 2, spe_forward_sa): argument errors are reported without a device."""
    assert lib.spe_forward_sa(None, None, 1, None, None, None, None, None, None, None, None, None) != 0
    assert b"null ctx" in lib.spe_global_last_error()
    for bad in (dict(precision=1), dict(has_sigma=0), dict(enc_layers=2), dict(input_size=512), dict(num_queries=31)):
        kw = dict(input_size=256, num_queries=30, enc_layers=1, dec_layers=3, hidden_dim=256, nheads=8, dim_feedforward=1024,
                  backbone=2, precision=0, has_sigma=1, max_batch=2)
        kw.update(bad)
        cfg = _lib.SpeConfig(**kw)
        ctx = C.c_void_p()
        assert lib.spe_create(C.byref(cfg), 0, C.byref(ctx)) != 0 and not ctx.value, bad
        assert b"SA predictor" in lib.spe_global_last_error(), bad
