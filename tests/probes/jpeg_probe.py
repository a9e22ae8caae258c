"""Throughput of the GPU JPEG decode: B synthetic SPEED-sized frames, PIL-encoded, decoded as one batch.
    python tests/probes/jpeg_probe.py [B] [quality]
"""
import io
import os
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import synth                                         # noqa: E402  (synthetic frames only)
from satellite_pose_estimation_b200 import Engine               # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 90
frames, det = synth.bench_set(0)
files = []
for i in range(B):
    buf = io.BytesIO()
    Image.fromarray(frames[i % len(frames)], "L").save(buf, "JPEG", quality=Q)
    files.append(buf.getvalue())
t0 = time.perf_counter()
ref = [np.asarray(Image.open(io.BytesIO(f))) for f in files[:16]]
t_pil = (time.perf_counter() - t0) / 16
eng = Engine(max_batch=8)
out = torch.empty((B, 1200, 1920), dtype=torch.uint8, device="cuda")
for r in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    eng.decode_jpeg(files, out=out)
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rep {r}: {B} files ({sum(map(len, files)) / 1e6:.1f} MB, q={Q}) host {t_host * 1e3:.2f} ms, device {ms:.2f} ms "
          f"-> {B / ms * 1e3:.0f} images/s; PIL on one core {t_pil * 1e3:.2f} ms per image")
ok = all(np.array_equal(out[i].cpu().numpy(), ref[i]) for i in range(min(B, 16)))
print("bit-exact with PIL:", ok)
eng.close()
