"""Batch-1 forward (crop -> predictor -> PnP) a few times, for latency measurements / ncu launch lists.
   python tests/b1_probe.py [reps]"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model_ref, synth  # noqa: E402
from satellite_pose_estimation_b200 import Engine  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
eng = Engine(max_batch=1)
eng.load_state_dict(synth.make_state_dict(model_ref.ModelCfg(), seed=0))
det = synth.load_detector_boxes()[:1]
frames = torch.from_numpy(synth.make_frames(1, det, seed=0)).cuda()
boxes = torch.from_numpy(eng.clip_boxes(det)).cuda()
p = synth.make_predictions(1, seed=1)
lg, pt, bx = (torch.from_numpy(p[k]).cuda() for k in ("logits", "points", "boxes"))
images = torch.empty((1, 3, 224, 224), device="cuda")
eng.register_stable_input(images)
lat = []
for i in range(reps):
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    eng.crop_resize_norm(frames, boxes, out=images)
    eng.forward(images)
    eng.assign_pnp(lg, pt, bx.to(torch.int32))
    b.record()
    torch.cuda.synchronize()
    lat.append(a.elapsed_time(b))
print(f"batch-1 latency: p50 {statistics.median(lat[reps // 3:]):.3f} ms, min {min(lat):.3f} ms over {reps} reps")
