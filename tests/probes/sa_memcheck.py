"""Small SA forward + pipeline batch for compute-sanitizer (run: compute-sanitizer --tool memcheck python tests/probes/sa_memcheck.py)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sa_model_ref, synth
from satellite_pose_estimation_b200 import Engine

cfg = sa_model_ref.SaCfg()
sd = synth.make_sa_state_dict(cfg, seed=0)
B = 3
eng = Engine(input_size=256, num_queries=30, enc_layers=1, dec_layers=3, dim_feedforward=1024, backbone="rtdetr_r50vd",
             precision="tf32", has_sigma=True, max_batch=B)
eng.load_state_dict(sd)
x = torch.randn(B, 3, 256, 256).cuda()
o = eng.forward_sa(x)
torch.cuda.synchronize()
det = synth.load_detector_boxes()[:B]
frames = torch.from_numpy(synth.make_frames(B, det, seed=3)).cuda()
boxes = torch.from_numpy(eng.clip_boxes(det)).cuda()
eng.submit_batch_dev(0, frames, boxes, reproj=25.0, weighted=True)
r = eng.collect_batch_host(0)
print("sa memcheck run done", o["pred_pts"].shape, r["status"])
eng.close()
# RV pipeline slot (crop -> stem fusion, 4-row LayerNorm needs >= 8192 rows: B = 11 -> 8624 token rows)
from oracle import model_ref
B = 11
eng = Engine(max_batch=B)
eng.load_state_dict(synth.make_state_dict(model_ref.ModelCfg(), seed=0))
det = synth.load_detector_boxes()[:B]
frames = torch.from_numpy(synth.make_frames(B, det, seed=4)).cuda()
boxes = torch.from_numpy(eng.clip_boxes(det)).cuda()
for _ in range(2):
    eng.submit_batch_dev(0, frames, boxes)
    r = eng.collect_batch_host(0)
print("rv memcheck run done", r["status"])
eng.close()
