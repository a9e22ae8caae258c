import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import crop_ref, model_ref, synth
from satellite_pose_estimation_b200 import Engine
torch.set_num_threads(os.cpu_count())
res = {}
for spread in (False, True):
    cfg = model_ref.ModelCfg(aux_loss=False)
    sd = synth.make_state_dict(cfg, seed=0, spread_labels=spread)
    frames, det = synth.bench_set(0)
    frames1, det1 = synth.bench_set(1)
    xc = torch.stack([crop_ref.crop_resize_normalize(frames[i], det[i], 224)[0] for i in range(0, 64, 4)])
    xt = torch.stack([crop_ref.crop_resize_normalize(frames1[i], det1[i], 224)[0] for i in range(3, 64, 4)])
    taps = {}
    ref = model_ref.forward(sd, cfg, xt, taps)
    eng = Engine(max_batch=16)
    eng.load_state_dict(sd)
    for cal in (0, 1, 2):
        if cal:
            eng.calibrate(xc.cuda())
        eng.enable_taps(True)
        out = eng.forward(xt.cuda())
        torch.cuda.synchronize()
        d = (out["pred_points"].cpu() - ref["pred_points"]) * 1748
        r = {"fg_rms": d[:, :11].pow(2).mean().sqrt().item(), "fg_max": d[:, :11].abs().max().item(), "all_max": d.abs().max().item(),
             "calibrated": eng.calibrated}
        for name, shape, t in (("layer1", (16, 56, 56, 256), taps["layer1"].permute(0, 2, 3, 1)), ("layer3", (16, 14, 14, 1024), taps["layer3"].permute(0, 2, 3, 1)),
                               ("neck", (16, 28, 28, 512), taps["neck"].permute(0, 2, 3, 1)), ("enc3", (16, 784, 256), taps["enc3"].permute(1, 0, 2)),
                               ("hs", (4, 16, 40, 256), taps["hs"])):
            g = eng.read_tap(name, shape)
            e = (g - t)
            r[name] = {"rel_rms": (e.pow(2).mean().sqrt() / t.pow(2).mean().sqrt()).item(), "mean_err_over_rms": (e.mean().abs() / t.pow(2).mean().sqrt()).item(),
                       "chan_mean_err_rms": (e.reshape(-1, e.shape[-1]).mean(0).pow(2).mean().sqrt() / t.pow(2).mean().sqrt()).item()}
        eng.enable_taps(False)
        res[f"spread={spread} cal={cal}"] = r
    eng.close()
print(json.dumps(res, indent=1))
