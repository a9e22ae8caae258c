"""GPU bring-up diagnostics (run on a B200 via gpurun; prints rich numbers, never asserts).

    python tests/gpu_bringup.py [gemm] [conv] [attn] [crop] [pnp] [model] [e2e]

The pytest parity suite (tests/test_*_gpu.py) holds the pass/fail bars; this script is the microscope.
"""
import ctypes as C
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satellite_pose_estimation_b200 import _lib  # noqa: E402
from satellite_pose_estimation_b200.engine import Engine  # noqa: E402

DEV = "cuda:0"


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stat(name, got, ref):
    got, ref = got.double().flatten(), ref.double().flatten()
    d = (got - ref).abs()
    rel = d.max().item() / max(ref.abs().max().item(), 1e-30)
    print(f"   {name:34s} max|d| {d.max().item():.3e}  mean|d| {d.mean().item():.3e}  ref max {ref.abs().max().item():.3e} "
          f" rel {rel:.2e}  nan {int(torch.isnan(got).sum())}")
    return rel


def tf32_round(x):
    return (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32) if False else x


def run_gemm():
    lib = _lib.load()
    torch.manual_seed(0)
    for dt, tdt in ((0, torch.float32), (1, torch.bfloat16)):
        for (M, N, K, relu, use_res, res_mod) in [(128, 64, 64, 0, 0, 0), (128, 128, 256, 0, 0, 0), (300, 256, 512, 1, 1, 0),
                                                   (3136 * 2, 64, 256, 1, 0, 0), (784 * 3, 768, 256, 0, 1, 784),
                                                   (40, 256, 2048, 0, 1, 0), (12544, 64, 192, 1, 0, 0),
                                                   (784 * 64, 2048, 256, 1, 0, 0)]:
            A = torch.randn(M, K, device=DEV).to(tdt)
            W = (torch.randn(N, K, device=DEV) / K ** 0.5).to(tdt)
            scale = torch.rand(N, device=DEV) + 0.5
            bias = torch.randn(N, device=DEV)
            rows = res_mod if res_mod else M
            res = torch.randn(rows, N, device=DEV).to(tdt) if use_res else None
            out = torch.full((M, N), float("nan"), device=DEV).to(tdt)
            rc = lib.spe_debug_gemm(dt, _p(A), _p(W), M, N, K, _p(scale), _p(bias), _p(res), res_mod, relu, _p(out), None)
            torch.cuda.synchronize()
            ref = A.double() @ W.double().t() * scale.double() + bias.double()
            if use_res:
                r = res.double()
                ref = ref + (r.repeat(M // rows + 1, 1)[:M] if res_mod else r)
            if relu:
                ref = ref.clamp_min(0)
            print(f"gemm dt={dt} M={M} N={N} K={K} relu={relu} res={use_res}/{res_mod} rc={rc}")
            stat("out", out, ref)
    # timing of a big one
    for dt, tdt in ((0, torch.float32), (1, torch.bfloat16)):
        M, N, K = 784 * 64, 2048, 256
        A = torch.randn(M, K, device=DEV).to(tdt); W = torch.randn(N, K, device=DEV).to(tdt)
        out = torch.empty(M, N, device=DEV).to(tdt)
        for _ in range(3):
            lib.spe_debug_gemm(dt, _p(A), _p(W), M, N, K, None, None, None, 0, 0, _p(out), None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            lib.spe_debug_gemm(dt, _p(A), _p(W), M, N, K, None, None, None, 0, 0, _p(out), None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"gemm timing dt={dt} {M}x{N}x{K}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s  "
              f"out GB/s {M * N * out.element_size() / ms / 1e6:.0f}")
        M, N, K = 784 * 64, 256, 2048
        A = torch.randn(M, K, device=DEV).to(tdt); W = torch.randn(N, K, device=DEV).to(tdt)
        out = torch.empty(M, N, device=DEV).to(tdt)
        for _ in range(3):
            lib.spe_debug_gemm(dt, _p(A), _p(W), M, N, K, None, None, None, 0, 0, _p(out), None)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            lib.spe_debug_gemm(dt, _p(A), _p(W), M, N, K, None, None, None, 0, 0, _p(out), None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"gemm timing dt={dt} {M}x{N}x{K}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")


def run_conv():
    lib = _lib.load()
    torch.manual_seed(1)
    for dt, tdt in ((0, torch.float32), (1, torch.bfloat16)):
        for (NB, H, C, Cout) in [(2, 28, 64, 64), (3, 56, 64, 64), (2, 14, 256, 256), (2, 28, 1024, 256), (1, 7, 64, 128)]:
            x = torch.randn(NB, H, H, C, device=DEV).to(tdt)
            w = (torch.randn(Cout, C, 3, 3, device=DEV) / (9 * C) ** 0.5).to(tdt)
            wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * C).contiguous()
            bias = torch.randn(Cout, device=DEV)
            out = torch.full((NB, H, H, Cout), float("nan"), device=DEV).to(tdt)
            rc = lib.spe_debug_conv(dt, _p(x), _p(wk), NB, H, H, C, Cout, 3, 3, 1, 1, None, _p(bias), 1, _p(out), None)
            torch.cuda.synchronize()
            ref = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double(), bias.double(), padding=1)
            ref = ref.clamp_min(0).permute(0, 2, 3, 1)
            print(f"conv dt={dt} NB={NB} H={H} C={C} Cout={Cout} rc={rc}")
            stat("out", out, ref)


def run_attn():
    lib = _lib.load()
    torch.manual_seed(2)
    for dt, tdt in ((0, torch.float32), (1, torch.bfloat16)):
        for (B, Lq, Lk) in [(2, 784, 784), (3, 40, 40), (2, 40, 784), (1, 100, 1024)]:
            q = torch.randn(B, Lq, 256, device=DEV).to(tdt)
            k = torch.randn(B, Lk, 256, device=DEV).to(tdt)
            v = torch.randn(B, Lk, 256, device=DEV).to(tdt)
            out = torch.full((B, Lq, 256), float("nan"), device=DEV).to(tdt)
            rc = lib.spe_debug_attention(dt, _p(q), _p(k), _p(v), _p(out), B, 8, Lq, Lk, 256, 256, 256, 256, None)
            torch.cuda.synchronize()
            qh = q.double().view(B, Lq, 8, 32).transpose(1, 2)
            kh = k.double().view(B, Lk, 8, 32).transpose(1, 2)
            vh = v.double().view(B, Lk, 8, 32).transpose(1, 2)
            att = torch.softmax(qh @ kh.transpose(-1, -2) / 32 ** 0.5, -1) @ vh
            ref = att.transpose(1, 2).reshape(B, Lq, 256)
            print(f"attn dt={dt} B={B} Lq={Lq} Lk={Lk} rc={rc}")
            stat("out", out, ref)
    B, L = 64, 784
    q = torch.randn(B, L, 768, device=DEV)
    out = torch.empty(B, L, 256, device=DEV)
    for _ in range(2):
        lib.spe_debug_attention(0, _p(q), C.c_void_p(q.data_ptr() + 1024), C.c_void_p(q.data_ptr() + 2048), _p(out), B, 8, L, L, 768, 768, 768, 256, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        lib.spe_debug_attention(0, _p(q), C.c_void_p(q.data_ptr() + 1024), C.c_void_p(q.data_ptr() + 2048), _p(out), B, 8, L, L, 768, 768, 768, 256, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"attn timing tf32 B=64 784x784: {ms:.3f} ms  {4 * B * 8 * L * L * 32 / ms / 1e9:.1f} TFLOP/s")


def run_crop():
    from oracle import crop_ref, synth
    boxes_all = synth.load_detector_boxes()
    idx = [0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987]
    # add extreme sizes
    sides = np.maximum(boxes_all[:, 2] - boxes_all[:, 0], boxes_all[:, 3] - boxes_all[:, 1])
    idx += [int(np.argmax(sides)), int(np.argmin(sides))]
    det = boxes_all[idx]
    frames = synth.make_frames(len(idx), det, seed=0)
    eng = Engine(max_batch=len(idx))
    clip = eng.clip_boxes(det)
    ref_clip = np.stack([crop_ref.generate_clip_bbox(b) for b in det])
    print("crop boxes bit-exact:", bool((clip == ref_clip).all()))
    out = eng.crop_resize_norm(torch.from_numpy(frames).to(DEV), torch.from_numpy(clip).to(DEV))
    torch.cuda.synchronize()
    out = out.cpu()
    tot = bad = 0
    for i in range(len(idx)):
        ref, _ = crop_ref.crop_resize_normalize(frames[i], det[i], 224)
        d = (out[i] - ref).abs()
        nbad = int((d > 1e-6).sum()); tot += d.numel(); bad += nbad
        print(f"   img {i} S={clip[i, 2] - clip[i, 0]:5d} x1={clip[i, 0]:5d} y1={clip[i, 1]:5d} mismatching values {nbad:4d} max|d| {d.max().item():.4f}")
    print(f"crop total mismatch fraction {bad / tot:.2e} (1 LSB = {1 / 255 / 0.225:.4f})")
    B = 64
    det = boxes_all[:B]
    frames = torch.from_numpy(synth.make_frames(4, det, seed=1)).to(DEV).repeat(16, 1, 1)
    eng2 = Engine(max_batch=B)
    clip = torch.from_numpy(eng2.clip_boxes(det)).to(DEV)
    o = eng2.crop_resize_norm(frames, clip)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        eng2.crop_resize_norm(frames, clip, out=o)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"crop timing B=64: {ms * 1e3:.1f} us  -> {B / ms * 1e3:.0f} img/s, store GB/s {o.numel() * 4 / ms / 1e6:.0f}")


def run_pnp(n=2000):
    from oracle import pnp_ref, synth
    d = synth.make_predictions(n, seed=1)
    eng = Engine(max_batch=8)
    t0 = time.time()
    out = eng.assign_pnp(torch.from_numpy(d["logits"]).to(DEV), torch.from_numpy(d["points"]).to(DEV),
                         torch.from_numpy(d["boxes"]).to(DEV), reproj=20.0, want_post=True)
    torch.cuda.synchronize()
    print(f"pnp kernel wall (incl. first launch) {time.time() - t0:.3f}s")
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    lg, pt, bx = torch.from_numpy(d["logits"][:64]).to(DEV), torch.from_numpy(d["points"][:64]).to(DEV), torch.from_numpy(d["boxes"][:64]).to(DEV)
    eng.assign_pnp(lg, pt, bx)
    e0.record()
    for _ in range(10):
        eng.assign_pnp(lg, pt, bx)
    e1.record(); torch.cuda.synchronize()
    print(f"pnp timing B=64: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
    q, t, st, asg = out["quat"].cpu().numpy(), out["tvec"].cpu().numpy(), out["status"].cpu().numpy(), out["assign"].cpu().numpy()
    probs, pts_px, inl = out["probs"].cpu().numpy(), out["points_px"].cpu().numpy(), out["inlier_mask"].cpu().numpy()
    res = pnp_ref.post_process(d["logits"], d["points"], d["boxes"])
    solver = pnp_ref.SimplePoseSolver(20, return_inliers=True)
    n_assign_bad = n_fail_mis = n_inl_mis = n_bad = 0
    max_ang = max_t = 0.0
    max_prob = max_px = 0.0
    for i in range(n):
        max_prob = max(max_prob, float(np.abs(probs[i] - res[i]["logits"]).max()))
        max_px = max(max_px, float(np.abs(pts_px[i] - res[i]["points"]).max()))
        tab = pnp_ref.assign_table(res[i]["points"], res[i]["logits"])
        if not (tab == asg[i]).all():
            n_assign_bad += 1
        try:
            q_ref, t_ref, used = solver(res[i]["points"], res[i]["logits"]); ok_ref = True
        except Exception:
            ok_ref = False
        ok_me = st[i] == 0
        if ok_ref != ok_me:
            n_fail_mis += 1
            print("   fail mismatch", i, ok_ref, st[i], d["n_visible"][i])
            continue
        if not ok_ref:
            continue
        labels = [l for l in range(11) if asg[i][l] >= 0]
        used_me = sorted(labels[j] for j in range(len(labels)) if (inl[i] >> j) & 1)
        if used_me != used:
            n_inl_mis += 1
            continue
        s_t, s_q = pnp_ref.speed_score(q[i], t[i], q_ref, t_ref)
        max_ang = max(max_ang, np.degrees(s_q)); max_t = max(max_t, s_t)
        if np.degrees(s_q) > 0.01 or s_t > 1e-4:
            n_bad += 1
    print(f"pnp n={n}: assign mismatches {n_assign_bad}, fail/ok mismatches {n_fail_mis}, inlier-set mismatches {n_inl_mis}, "
          f"out-of-tol {n_bad}, max rot err {max_ang:.3e} deg, max rel t err {max_t:.3e}, prob max|d| {max_prob:.2e}, px max|d| {max_px:.2e}")


def run_model():
    from oracle import model_ref, synth
    for prec in ("tf32", "bf16"):
        cfg = model_ref.ModelCfg(sigma_head=True)
        sd = synth.make_state_dict(cfg, seed=0)
        B = 3
        torch.manual_seed(3)
        x = torch.randn(B, 3, 224, 224)
        eng = Engine(max_batch=4, precision=prec, has_sigma=True)
        t0 = time.time(); eng.load_state_dict(sd); print(f"[{prec}] load weights {time.time() - t0:.2f}s")
        eng.enable_taps(True)
        out = eng.forward(x.to(DEV), want_aux=True)
        torch.cuda.synchronize()
        taps = {}
        ref = model_ref.forward(sd, cfg, x, taps)
        print(f"model parity [{prec}] B={B}")
        nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
        stat("stem", eng.read_tap("stem", (B, 112, 112, 64)), nhwc(taps["stem"]))
        stat("layer1", eng.read_tap("layer1", (B, 56, 56, 256)), nhwc(taps["layer1"]))
        stat("layer2", eng.read_tap("layer2", (B, 28, 28, 512)), nhwc(taps["layer2"]))
        stat("layer3", eng.read_tap("layer3", (B, 14, 14, 1024)), nhwc(taps["layer3"]))
        stat("neck", eng.read_tap("neck", (B, 28, 28, 512)), nhwc(taps["neck"]))
        stat("input_proj", eng.read_tap("input_proj", (B, 784, 256)), taps["input_proj"].permute(1, 0, 2))
        for i in range(4):
            stat(f"enc{i}", eng.read_tap(f"enc{i}", (B, 784, 256)), taps[f"enc{i}"].permute(1, 0, 2))
        stat("hs", eng.read_tap("hs", (4, B, 40, 256)), taps["hs"])
        stat("pred_logits", out["pred_logits"].cpu(), ref["pred_logits"])
        stat("pred_points", out["pred_points"].cpu(), ref["pred_points"])
        stat("pred_sigmas", out["pred_sigmas"].cpu(), ref["pred_sigmas"])
        for i, (a, b) in enumerate(zip(out["aux_outputs"], ref["aux_outputs"])):
            stat(f"aux{i} logits", a["pred_logits"].cpu(), b["pred_logits"])
            stat(f"aux{i} points", a["pred_points"].cpu(), b["pred_points"])
        print("   label agreement", float((out["pred_logits"].cpu().argmax(-1) == ref["pred_logits"].argmax(-1)).float().mean()))
        eng.enable_taps(False)
        eng.close()
        # timing
        for Bt in (1, 64):
            eng = Engine(max_batch=Bt, precision=prec)
            cfg2 = model_ref.ModelCfg()
            eng.load_state_dict(synth.make_state_dict(cfg2, seed=0))
            xt = torch.randn(Bt, 3, 224, 224, device=DEV)
            for _ in range(3):
                eng.forward(xt)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(5):
                eng.forward(xt)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"forward timing [{prec}] B={Bt}: {ms:.3f} ms -> {Bt / ms * 1e3:.0f} img/s, {Bt * 26.57 / ms:.1f} TFLOP/s")
            eng.close()


def run_e2e():
    from oracle import model_ref, synth
    boxes_all = synth.load_detector_boxes()
    B = 64
    det = boxes_all[:B]
    frames = synth.make_frames(8, det, seed=0)
    frames = np.concatenate([frames] * 8)
    fh = torch.from_numpy(frames).pin_memory()
    eng = Engine(max_batch=B)
    eng.load_state_dict(synth.make_state_dict(model_ref.ModelCfg(), seed=0))
    for _ in range(2):
        r = eng.run_batch_host(fh, det)
    t0 = time.time()
    for _ in range(5):
        r = eng.run_batch_host(fh, det)
    dt = (time.time() - t0) / 5
    print(f"e2e host->host B=64: {dt * 1e3:.2f} ms -> {B / dt:.0f} img/s; status histogram {np.bincount(r['status'], minlength=4)}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "conv", "attn", "crop", "pnp", "model", "e2e"]
    print(torch.cuda.get_device_name(0), torch.__version__)
    for w in which:
        print(f"================ {w}")
        try:
            globals()["run_" + w]()
        except Exception:
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e:
                print("CUDA context dead:", e)
                break
