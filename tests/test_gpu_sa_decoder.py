"""First kernels of the SA (RT-DETR) decoder against goldens of the SA drop's own code:
``deformable_attention_core_func`` (SA/src/zoo/rtdetr/utils.py:15-64), the live ``MSDeformableAttention`` module
(SA/src/zoo/rtdetr/rtdetr_decoder.py:40-191) and the encoder top-k query selection (:646-680)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import synth
from oracle.make_golden import DEFORM_CASE, deform_inputs
from satellite_pose_estimation_b200 import Engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(lib, cuda_dev):
    e = Engine(max_batch=1)
    yield e
    e.close()


def test_deformable_attention_core_matches_reference(eng):
    g = np.load(os.path.join(synth.GOLDEN_DIR, "deform_attn_golden.npz"))
    d = deform_inputs()
    out = eng.ms_deform_attn(d["core_value"].cuda(), DEFORM_CASE["levels"], d["core_loc"].cuda(), d["core_attn"].cuda())
    torch.cuda.synchronize()
    err = np.abs(out.cpu().numpy() - g["core_func_out"]).max()
    print(f"deformable attention core: max |d| {err:.2e} (outputs up to {np.abs(g['core_func_out']).max():.2f})")
    assert err < 2e-6


def test_deformable_attention_fused_matches_live_module(eng):
    """softmax over levels x points + reference point + offset / (W, H) + sampling, as the live module computes them
    between its linear layers; the linears themselves are plain GEMMs (run here with torch on the same weights)."""
    g = np.load(os.path.join(synth.GOLDEN_DIR, "deform_attn_golden.npz"))
    d = deform_inputs()
    w = d["weights"]
    bs, Lq, H, P = DEFORM_CASE["bs"], DEFORM_CASE["Lq"], DEFORM_CASE["heads"], DEFORM_CASE["points"]
    L = len(DEFORM_CASE["levels"])
    value = F.linear(d["value_in"], w["value_proj.weight"], w["value_proj.bias"]).reshape(bs, -1, H, 32)
    off = F.linear(d["query"], w["sampling_offsets.weight"], w["sampling_offsets.bias"]).reshape(bs, Lq, H, L, P, 2)
    logit = F.linear(d["query"], w["attention_weights.weight"], w["attention_weights.bias"]).reshape(bs, Lq, H, L, P)
    core = eng.ms_deform_attn(value.cuda(), DEFORM_CASE["levels"], off.cuda(), logit.cuda(), ref=d["ref"].cuda())
    torch.cuda.synchronize()
    err = np.abs(core.cpu().numpy() - g["module_core_out"]).max()
    out = F.linear(core.cpu(), w["output_proj.weight"], w["output_proj.bias"]).numpy()
    err_out = np.abs(out - g["module_out"]).max()
    print(f"fused deformable attention vs live module: core max |d| {err:.2e}, module output max |d| {err_out:.2e}")
    assert err < 2e-5 and err_out < 2e-5


def test_topk_query_selection_matches_torch_topk(eng):
    """rtdetr_decoder.py:646-680: topk over the per-token maximum class score, then three gathers."""
    rng = np.random.default_rng(9)
    B, Lv, C, k = 5, 1344, 12, 30
    cls = torch.from_numpy(rng.standard_normal((B, Lv, C)).astype(np.float32))
    coord = torch.from_numpy(rng.standard_normal((B, Lv, 2)).astype(np.float32))
    mem = torch.from_numpy(rng.standard_normal((B, Lv, 256)).astype(np.float32))
    val_ref, ind_ref = torch.topk(cls.max(-1).values, k, dim=1)                       # the reference's call
    val, idx = eng.topk_queries(cls.cuda(), k)
    assert torch.equal(idx.cpu().long(), ind_ref) and torch.equal(val.cpu(), val_ref)
    for src in (coord, cls, mem):
        want = src.gather(dim=1, index=ind_ref.unsqueeze(-1).repeat(1, 1, src.shape[-1]))
        assert torch.equal(eng.gather_rows(src.cuda(), idx).cpu(), want)
    # ties resolve to the lower token index, deterministically
    flat = torch.zeros(1, 64, 3); flat[0, 10, 1] = flat[0, 20, 2] = 1.0
    _, i2 = eng.topk_queries(flat.cuda(), 4)
    assert i2.cpu().tolist() == [[10, 20, 0, 1]]
