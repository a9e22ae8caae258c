"""Calibrated heads for the whole-chain tests (TEST INFRASTRUCTURE; writes tests/golden/chain_heads.npz).

    python -m oracle.make_chain_fixture

Why: with random-init weights every query of the keypoint-set predictor emits the same label,
so the pose stage behind it would always take the "< 4 keypoints" exit and crop -> predictor -> PnP could never be
checked, or timed, as ONE chain on the network's own output.  This script makes the seeded C* weights behave like a
trained predictor without training the trunk:
  * ``query_embed`` x16: the learned queries dominate the decoder state (trained DETR queries have large norms);
    the query-dependent part of the decoder output then exceeds its image-dependent part,
  * ``cls_embed``: ridge regression on the oracle's decoder outputs of 256 bench crops so that query q < 11 scores
    keypoint label q and every other query scores background (held-out accuracy is checked below),
  * ``point_embed`` (all three layers): Adam on the same outputs so that query q < 11 lands on keypoint q of a fixed,
    PnP-consistent layout (``synth.canonical_layout``) -- a few px of scatter remain, like a real predictor's noise.
The trunk, encoder and decoder weights stay the seeded random ones.  The result is committed because the fit is not
bit-reproducible across CPUs; ``synth.make_state_dict(cfg, seed=0, spread_labels=True)`` applies it.
"""
import os
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import crop_ref, model_ref, pnp_ref, synth

R = 224
N_FIT_SETS, N_VAL_SETS = 4, 2            # 64-frame sets of synth.bench_set: 256 images fitted, 128 held out


def decoder_outputs(sd, cfg, n_sets):
    hs, clips = [], []
    for s in range(n_sets):
        frames, det = synth.bench_set(s)
        crops = []
        for i in range(len(frames)):
            t, c = crop_ref.crop_resize_normalize(frames[i], det[i], R)
            crops.append(t); clips.append(c)
        for i in range(0, len(crops), 32):
            taps = {}
            model_ref.forward(sd, cfg, torch.stack(crops[i:i + 32]), taps)
            hs.append(taps["hs"][-1])
    return torch.cat(hs), np.stack(clips)


def fit_cls(hs, lam=1.0):
    """ridge regression to +-4 logits: label q for query q < 11, background (11) otherwise"""
    B, Q, E = hs.shape
    X = hs.reshape(-1, E).double()
    y = torch.tensor([q if q < 11 else 11 for q in range(Q)]).repeat(B)
    T = F.one_hot(y, 12).double() * 8 - 4
    Xa = torch.cat([X, torch.ones(len(X), 1, dtype=torch.double)], 1)
    W = torch.linalg.solve(Xa.T @ Xa + lam * torch.eye(E + 1, dtype=torch.double), Xa.T @ T)
    return W[:-1].T.float().contiguous(), W[-1].float().contiguous()


def cls_margin(hs, w, b):
    B, Q, _ = hs.shape
    logits = F.linear(hs, w, b)
    y = torch.tensor([q if q < 11 else 11 for q in range(Q)]).expand(B, Q)
    top2 = logits.topk(2, -1).values
    return (logits.argmax(-1) == y).float().mean().item(), (top2[..., 0] - top2[..., 1]).min().item()


def fit_points(hs, sd, layout, iters=8000):
    torch.manual_seed(0)
    X = hs[:, :11].reshape(-1, hs.shape[-1])
    T = torch.from_numpy(layout).float().repeat(hs.shape[0], 1)
    P = [sd[f"point_embed.layers.{i}.{k}"].clone().requires_grad_() for i in range(3) for k in ("weight", "bias")]

    def mlp(x):
        x = F.relu(F.linear(x, P[0], P[1]))
        x = F.relu(F.linear(x, P[2], P[3]))
        return torch.sigmoid(F.linear(x, P[4], P[5]))
    opt = torch.optim.Adam(P, lr=2e-3)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, iters)
    for it in range(iters):
        opt.zero_grad()
        loss = (mlp(X) - T).pow(2).mean()
        loss.backward()
        opt.step(); sched.step()
    return [p.detach() for p in P]


def main():
    torch.set_num_threads(os.cpu_count())
    cfg = model_ref.ModelCfg(aux_loss=False)
    sd = synth.make_state_dict(cfg, seed=0)
    sd["query_embed.weight"] = sd["query_embed.weight"] * synth.CHAIN_QUERY_SCALE
    t0 = time.time()
    hs, clips = decoder_outputs(sd, cfg, N_FIT_SETS + N_VAL_SETS)
    print(f"decoder outputs of {len(hs)} crops in {time.time() - t0:.0f} s")
    nfit = N_FIT_SETS * 64
    layout = synth.canonical_layout()
    cw, cb = fit_cls(hs[:nfit])
    acc, margin = cls_margin(hs[nfit:], cw, cb)
    print(f"cls_embed: held-out accuracy {acc:.4f}, smallest top-1/top-2 logit margin {margin:.2f}")
    assert acc == 1.0 and margin > 2.0
    P = fit_points(hs[:nfit], sd, layout)
    sd2 = dict(sd)
    sd2["cls_embed.weight"], sd2["cls_embed.bias"] = cw, cb
    for i in range(3):
        sd2[f"point_embed.layers.{i}.weight"], sd2[f"point_embed.layers.{i}.bias"] = P[2 * i], P[2 * i + 1]
    pts = model_ref.mlp3(hs, sd2, "point_embed").sigmoid()
    err = (pts[:, :11] - torch.from_numpy(layout).float()).abs()
    print(f"point_embed: |err| fitted mean {err[:nfit].mean():.5f} max {err[:nfit].max():.4f}; held-out mean "
          f"{err[nfit:].mean():.5f} max {err[nfit:].max():.4f} (normalised; x430 px at the median crop)")
    # the oracle chain on these heads: how many poses solve, with how many inliers
    logits = F.linear(hs, cw, cb).numpy()
    res = pnp_ref.post_process(logits, pts.numpy(), clips)
    solver = pnp_ref.SimplePoseSolver(20)
    ok = sum(pnp_ref.solve_or_zero(solver, r["points"], r["logits"])[2] for r in res)
    print(f"oracle chain (PostProcess + cv2 RANSAC-P3P + LM): {ok} of {len(res)} poses solved")
    out = {"query_embed_scale": np.float32(synth.CHAIN_QUERY_SCALE), "layout": layout,
           "cls_embed.weight": cw.numpy(), "cls_embed.bias": cb.numpy()}
    for i in range(3):
        out[f"point_embed.layers.{i}.weight"] = P[2 * i].numpy()
        out[f"point_embed.layers.{i}.bias"] = P[2 * i + 1].numpy()
    np.savez_compressed(os.path.join(synth.GOLDEN_DIR, "chain_heads.npz"), **out)
    print("wrote chain_heads.npz")


if __name__ == "__main__":
    main()
