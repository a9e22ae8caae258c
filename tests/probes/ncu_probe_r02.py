"""Stand-alone launches of the crop kernel (64 frames of bench.py's frame set 0, real box distribution) and of the encoder
self-attention shape (B = 64, 8 heads, 784 x 784, d = 32) for `ncu --set full` captures (profiles/r02_ncu_*.md).

    python tests/probes/ncu_probe_r02.py [reps]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import synth                                         # noqa: E402  (synthetic frames only)
from satellite_pose_estimation_b200 import Engine, _lib          # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
lib = _lib.load()
eng = Engine(max_batch=64)
frames, det = synth.bench_set(0)
fd = torch.from_numpy(frames).cuda()
bd = torch.from_numpy(eng.clip_boxes(det)).cuda()
out = torch.empty((64, 3, 224, 224), device="cuda")
src_bytes = sum(min(int(b[2] - b[0]) ** 2, 16 * 224 * 224) for b in bd.cpu().numpy())
for r in range(reps):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    eng.crop_resize_norm(fd, bd, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3
    alg = src_bytes + out.numel() * 4
    print(f"crop rep {r}: {us:.1f} us; algorithmic bytes {alg / 1e6:.1f} MB -> {alg / us / 1e6:.2f} TB/s")
p = lambda t: C.c_void_p(t.data_ptr())
torch.manual_seed(0)
qkv = torch.randn(64, 784, 768, device="cuda")
att = torch.empty(64, 784, 256, device="cuda")
for r in range(reps):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    rc = lib.spe_debug_attention(0, p(qkv), C.c_void_p(qkv.data_ptr() + 1024), C.c_void_p(qkv.data_ptr() + 2048), p(att),
                                 64, 8, 784, 784, 768, 768, 768, 256, None)
    e1.record()
    torch.cuda.synchronize()
    print(f"attention rep {r}: rc={rc} {e0.elapsed_time(e1) * 1e3:.1f} us")
eng.close()
