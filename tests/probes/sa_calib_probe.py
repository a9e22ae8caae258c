"""SA predictor: keypoint error against the oracle with and without spe_calibrate (reference selection handed in)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sa_model_ref, synth
from oracle.make_golden import model_inputs
from satellite_pose_estimation_b200 import Engine

cfg = sa_model_ref.SaCfg()
sd = synth.make_sa_state_dict(cfg, seed=0)
B = 16
x = model_inputs(B, 256, 21)
xcal = model_inputs(16, 256, 22)
taps = {}
ref = sa_model_ref.forward(sd, cfg, x, taps)
eng = Engine(input_size=256, num_queries=30, enc_layers=1, dec_layers=3, dim_feedforward=1024, backbone="rtdetr_r50vd",
             precision="tf32", has_sigma=True, max_batch=B)
eng.load_state_dict(sd)
tk = taps["topk"].to(torch.int32).cuda()
def report(tag):
    eng.enable_taps(True)
    o = eng.forward_sa(x.cuda(), topk_override=tk)
    torch.cuda.synchronize()
    mem = eng.read_tap("sa_memory", (B, 1344, 256))
    eng.enable_taps(False)
    d = (o["pred_pts"].cpu() - ref["pred_pts"]) * 1748
    m = mem - taps["memory"]
    print(f"{tag}: keypoints max {d.abs().max():.3f} px rms {d.pow(2).mean().sqrt():.3f} px | memory rel rms "
          f"{(m.pow(2).mean().sqrt() / taps['memory'].pow(2).mean().sqrt()).item():.2e} mean-bias/rms {(m.mean(dim=(0,1)).abs().mean() / m.pow(2).mean().sqrt()).item():.2f}"
          f" | logits {(o['pred_logits'].cpu() - ref['pred_logits']).abs().max():.1e} sigma {(o['pred_sigmas'].cpu() - ref['pred_sigmas']).abs().max():.1e}")
report("uncalibrated")
eng.calibrate(xcal.cuda())
report("calibrated on 16 other images")
eng.close()
# throughput of the plain forward (graph replay) at batch 64
eng = Engine(input_size=256, num_queries=30, enc_layers=1, dec_layers=3, dim_feedforward=1024, backbone="rtdetr_r50vd",
             precision="tf32", has_sigma=True, max_batch=64)
eng.load_state_dict(sd)
xb = model_inputs(64, 256, 5).cuda()
for _ in range(4):
    eng.forward(xb)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    eng.forward(xb)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"B=64 forward: {ms:.3f} ms per batch = {64 / ms * 1e3:.0f} images/s (SPE_SA_X3={os.environ.get('SPE_SA_X3', '1')})")
