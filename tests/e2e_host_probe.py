"""Host-side cost of the e2e loop (not a pytest file): where do the milliseconds between `value` and `e2e` go?
Times, per batch, the host thread's submit_batch_host / collect_batch_host calls of the pipelined loop bench.py runs,
and the same loop with frames resident in HBM.  Lives under tests/ because its inputs come from oracle/synth.
    python tests/e2e_host_probe.py [slots] [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model_ref, synth  # noqa: E402
from satellite_pose_estimation_b200 import Engine  # noqa: E402

SLOTS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 40
B = 64
dev = torch.device("cuda", 0)
eng = Engine(max_batch=B)
eng.load_state_dict(synth.make_state_dict(model_ref.ModelCfg(), seed=0))
det_all = synth.load_detector_boxes()
base = synth.make_frames(8, det_all, seed=100)
fh, det, fd, bd = [], [], [], []
for s in range(2):
    d = det_all[s * B:(s + 1) * B]
    f = torch.from_numpy(np.concatenate([np.roll(base, 37 * (s * 8 + k), axis=2) for k in range(B // 8)])).pin_memory()
    fh.append(f); det.append(d); fd.append(f.to(dev)); bd.append(torch.from_numpy(eng.clip_boxes(d)).to(dev))
preds = synth.make_predictions(B, Q=40, seed=1)
eng.set_pnp_override(torch.from_numpy(preds["logits"]).to(dev), torch.from_numpy(preds["points"]).to(dev),
                     torch.from_numpy(preds["boxes"]).to(torch.int32).to(dev))


def loop(submit, n):
    t_sub = t_col = 0.0
    for i in range(min(SLOTS, n)):
        submit(i % SLOTS, i)
    torch.cuda.synchronize()
    # steady state only: everything queued before t0 is done, refill and time from here
    for i in range(min(SLOTS, n)):
        eng.collect_batch_host(i % SLOTS)
    for i in range(min(SLOTS, n)):
        submit(i % SLOTS, i)
    t0 = time.perf_counter()
    for i in range(n):
        a = time.perf_counter()
        eng.collect_batch_host(i % SLOTS)
        b = time.perf_counter()
        if i + SLOTS < n:
            submit(i % SLOTS, i + SLOTS)
        c = time.perf_counter()
        t_col += b - a
        t_sub += c - b
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    return tot / n * 1e3, t_sub / max(n - SLOTS, 1) * 1e3, t_col / n * 1e3


for name, sub in (("dev ", lambda s, i: eng.submit_batch_dev(s, fd[i % 2], bd[i % 2])),
                  ("host", lambda s, i: eng.submit_batch_host(s, fh[i % 2], det[i % 2]))):
    loop(sub, 2 * SLOTS)
    ms, sub_ms, col_ms = loop(sub, STEPS)
    print(f"{name} frames, {SLOTS} slots: {ms:.3f} ms per batch (wall); host thread: submit {sub_ms:.3f} ms, "
          f"collect (mostly waiting) {col_ms:.3f} ms per batch")
