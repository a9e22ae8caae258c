"""Import the REAL reference from /root/reference (build container only; TEST INFRASTRUCTURE).

The reference is not importable as shipped on this image; two in-process shims (no file edits) are needed
(SURVEY.md section 8c):
  1. ``torchvision.__version__`` is parsed with ``float(v[:3]) < 0.7`` (RV/utils/misc.py:21-23): "0.26.0" reads as
     0.2 and pulls removed private symbols -> present a large version string while importing.
  2. ``Backbone8s`` asks torchvision for pretrained weights on the main process (RV/models/backbone.py:98, :116):
     no network here -> force ``is_main_process() == False`` so ``pretrained=False``.
"""
import os
import sys
import contextlib

REFERENCE_ROOT = "/root/reference"
RV_ROOT = os.path.join(REFERENCE_ROOT, "Revisiting Monocular Satellite Pose Estimation With Transformer")


def available():
    return os.path.isdir(RV_ROOT)


@contextlib.contextmanager
def _rv_on_path():
    sys.path.insert(0, RV_ROOT)
    cwd = os.getcwd()
    os.chdir(RV_ROOT)
    try:
        yield
    finally:
        os.chdir(cwd)
        sys.path.remove(RV_ROOT)


def import_rv_models():
    """Returns the reference's ``models`` package (``models.build_model``, ``models.PostProcess``)."""
    import torchvision
    real_version = torchvision.__version__
    with _rv_on_path():
        torchvision.__version__ = "9.9.0"
        try:
            import utils.misc  # noqa: F401  (reference module)
            import models  # noqa: F401
            import models.backbone as rv_backbone
        finally:
            torchvision.__version__ = real_version
        rv_backbone.is_main_process = lambda: False
    return sys.modules["models"]


def build_reference_model(cfg, state_dict=None):
    """``build_model(args)`` of the reference (RV/models/__init__.py:5-6), eval mode, optional weights."""
    import warnings
    from .synth import reference_args
    models = import_rv_models()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model, criterion, postprocessors = models.build_model(reference_args(cfg))
    model.eval()
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model, criterion, postprocessors


SA_ROOT = os.path.join(REFERENCE_ROOT, "Monocular Satellite Pose Estimation Based on Uncertainty Estimation and Self-Assessment")


def import_sa_rtdetr():
    """The SA drop's ``src.zoo.rtdetr`` package (``rtdetr_decoder.MLP`` / ``TransformerDecoder.sigma_embed``,
    ``utils.deformable_attention_core_func``, ...).  The tree is not importable as shipped (SURVEY.md section 2b:
    ``src/__init__.py`` expects ``nn`` / ``optim`` / ``solver`` under ``src/`` but they sit at the project root, and
    ``nn/backbone/__init__.py`` needs ``timm``): a synthetic ``src`` namespace package spanning both directories is
    registered instead -- no reference file is edited."""
    import types
    if "src.zoo.rtdetr" not in sys.modules:
        src = types.ModuleType("src")
        src.__path__ = [os.path.join(SA_ROOT, "src"), SA_ROOT]
        sys.modules["src"] = src
        nn_ = types.ModuleType("src.nn")
        nn_.__path__ = [os.path.join(SA_ROOT, "nn")]
        sys.modules["src.nn"] = nn_
        bb = types.ModuleType("src.nn.backbone")
        bb.__path__ = [os.path.join(SA_ROOT, "nn", "backbone")]
        sys.modules["src.nn.backbone"] = bb
        import src.core  # noqa: F401
        import src.nn.backbone.presnet  # noqa: F401
        import src.zoo.rtdetr  # noqa: F401
    return sys.modules["src.zoo.rtdetr"]


SA_MODEL_YML = "configs/rtdetr_speed/rtdetr_r50vd_6x_speed_kl_1.yml"


SA_MODEL_YML_R18 = "configs/rtdetr_speed/rtdetr_r18vd_6x_speed_kl_1.yml"


def build_sa_reference_model(state_dict=None, yml=None):
    """The SA drop's live ``RTDETR`` model of ``configs/rtdetr_speed/rtdetr_r50vd_6x_speed_kl_1.yml`` (PResNet-50-vd,
    HybridEncoder, RTDETRTransformer with 30 queries / 3 decoder layers, eval size 256), built by the reference's own
    ``YAMLConfig`` with the pretrained-weight download switched off; eval mode, optional weights (strict)."""
    import warnings
    import torch
    import_sa_rtdetr()
    from src.core import YAMLConfig
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfg = YAMLConfig(os.path.join(SA_ROOT, yml or SA_MODEL_YML))
        cfg.yaml_cfg["PResNet"]["pretrained"] = False
        torch.manual_seed(0)
        model = cfg.model
    model.eval()
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model
