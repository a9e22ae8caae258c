"""ctypes binding of libspe.so (the C ABI declared in include/spe.h).

There is no fallback: if the shared library is missing or cannot be loaded this module raises, and every product
entry point fails loudly rather than silently running something else.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspe.so")


class SpeConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "input_size", "num_queries", "enc_layers", "dec_layers", "hidden_dim", "nheads", "dim_feedforward",
        "backbone", "precision", "has_sigma", "max_batch")]


class SpeTensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int), ("shape", C.c_longlong * 4)]


class SpePnpParams(C.Structure):
    _fields_ = [("reproj_thresh", C.c_float), ("weighted", C.c_int), ("reject", C.c_int),
                ("reject_rms_px", C.c_float), ("reject_sigma_px", C.c_float), ("float_boxes_dev", C.c_void_p),
                ("reproj_thresh_dev", C.c_void_p), ("inputs_post_processed", C.c_int),
                ("sigma_px_scale", C.c_float)]


# every symbol include/spe.h declares: (restype, argtypes)
_vp, _i, _ll = C.c_void_p, C.c_int, C.c_longlong
PIPELINE_SLOTS = 8   # include/spe.h: SPE_PIPELINE_SLOTS

SYMBOLS = {
    "spe_create": (_i, [C.POINTER(SpeConfig), _i, C.POINTER(_vp)]),
    "spe_destroy": (None, [_vp]),
    "spe_last_error": (C.c_char_p, [_vp]),
    "spe_global_last_error": (C.c_char_p, []),
    "spe_load_weights": (_i, [_vp, C.POINTER(SpeTensorDesc), _i]),
    "spe_sync": (_i, [_vp, _vp]),
    "spe_clip_boxes": (_i, [_vp, _i, _vp]),
    "spe_clip_boxes_val": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "spe_speed_score": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "spe_jpeg_decode_batch": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _ll, _ll, _vp]),
    "spe_jpeg_info": (_i, [_vp, _ll, _vp, _vp]),
    "spe_crop_resize_norm": (_i, [_vp, _vp, _i, _i, _ll, _ll, _vp, _i, _i, _vp, _vp]),
    "spe_forward": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spe_forward_sa": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spe_calibrate": (_i, [_vp, _vp, _i, _vp]),
    "spe_is_calibrated": (_i, [_vp]),
    "spe_assign_pnp": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, C.POINTER(SpePnpParams), _vp, _vp, _vp, _vp, _vp, _vp,
                            _vp, _vp, _vp]),
    "spe_ensemble_pnp": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, C.POINTER(SpePnpParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spe_ms_deform_attn": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "spe_topk_queries": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "spe_gather_rows": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "spe_run_batch_host": (_i, [_vp, _vp, _i, _i, _vp, _i, C.POINTER(SpePnpParams), _vp, _vp, _vp, _vp, _vp]),
    "spe_debug_gemm": (_i, [_i, _vp, _vp, _ll, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "spe_debug_gemm2": (_i, [_i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _ll, _i, _vp, _i, _vp, _vp]),
    "spe_debug_conv": (_i, [_i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "spe_debug_attention": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "spe_debug_ffn": (_i, [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "spe_debug_enable_taps": (_i, [_vp, _i]),
    "spe_debug_read_tap": (_ll, [_vp, C.c_char_p, _vp, _ll]),
    "spe_submit_batch_host": (_i, [_vp, _i, _vp, _i, _i, _vp, _i, C.POINTER(SpePnpParams)]),
    "spe_collect_batch_host": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "spe_submit_batch_dev": (_i, [_vp, _i, _vp, _i, _i, _ll, _ll, _vp, _i, C.POINTER(SpePnpParams), _vp]),
    "spe_last_h2d_bytes": (_ll, [_vp]),
    "spe_profile_enable": (_i, [_i]),
    "spe_profile_collect": (_i, [_vp, _vp]),
    "spe_debug_set_pnp_override": (_i, [_vp, _vp, _vp, _vp]),
    "spe_debug_graph_stats": (_i, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "spe_debug_read_slot_outputs": (_i, [_vp, _i, _vp, _vp]),
}

_lib = None


def load():
    """Load libspe.so once and attach prototypes.  Raises RuntimeError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m satellite_pose_estimation_b200.build` "
            "(nvcc, sm_100a).  This package has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SpeError(RuntimeError):
    pass


def check(rc, ctx=None):
    if rc == 0:
        return
    lib = load()
    msg = lib.spe_last_error(ctx) if ctx else lib.spe_global_last_error()
    raise SpeError(f"libspe error {rc}: {(msg or b'').decode(errors='replace')}")
