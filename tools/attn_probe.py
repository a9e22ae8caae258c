"""Standalone launches of the encoder self-attention shape (B=64, 8 heads, 784x784, d=32) for ncu captures."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satellite_pose_estimation_b200 import _lib  # noqa: E402

lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr())
B, L = 64, 784
torch.manual_seed(0)
qkv = torch.randn(B, L, 768, device="cuda:0")
out = torch.empty(B, L, 256, device="cuda:0")
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    rc = lib.spe_debug_attention(0, p(qkv), C.c_void_p(qkv.data_ptr() + 1024), C.c_void_p(qkv.data_ptr() + 2048), p(out),
                                 B, 8, L, L, 768, 768, 768, 256, None)
    e1.record()
    torch.cuda.synchronize()
    print(f"rep {r} rc={rc} {e0.elapsed_time(e1) * 1e3:.1f} us")
