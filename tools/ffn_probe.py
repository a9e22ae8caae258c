"""Standalone launches of the fused feed-forward kernel at the encoder's shape (for timing / ncu captures).
   python tools/ffn_probe.py [reps]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satellite_pose_estimation_b200 import _lib  # noqa: E402

lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = "cuda:0"
torch.manual_seed(0)
M, F = 50176, 2048
X = torch.randn(M, 256, device=dev)
W1 = torch.randn(F, 256, device=dev) / 16
W2 = torch.randn(256, F, device=dev) / F ** 0.5
b1, b2 = torch.randn(F, device=dev), torch.randn(256, device=dev)
g, be = torch.rand(256, device=dev) + 0.5, torch.randn(256, device=dev)
out = torch.empty(M, 256, device=dev)
for r in range(reps):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    rc = lib.spe_debug_ffn(p(X), M, p(W1), p(b1), p(W2), p(b2), p(g), p(be), F, 0, p(out), None)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3
    print(f"rep {r} fused ffn rc={rc} {us:8.1f} us  {4 * M * 256 * F / us / 1e6:6.1f} TFLOP/s")
