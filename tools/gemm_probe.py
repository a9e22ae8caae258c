"""Standalone launches of the three epilogue-heavy GEMM shapes of the forward (for ncu --set full captures).
   python tools/gemm_probe.py [reps]"""
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
cases = [  # name, M, N, K, relu, residual rows (0 = none, -1 = full), res_mod
    ("l1.conv3+res", 200704, 256, 64, 1, -1, 0),
    ("enc.ff1", 50176, 2048, 256, 1, 0, 0),
    ("enc.qkv+addend", 50176, 768, 256, 0, 784, 784),
    ("enc.out+res", 50176, 256, 256, 0, -1, 0),
    ("ff1 K=32", 50176, 2048, 32, 1, 0, 0),
    ("qkv K=32", 50176, 768, 32, 0, 0, 0),
    ("N256 K=32", 401408, 256, 32, 1, 0, 0),
    ("N64 K=32", 1605632, 64, 32, 1, 0, 0),
]
bufs = []
for name, M, N, K, relu, rrows, rmod in cases:
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    sc = torch.rand(N, device=dev) + 0.5
    bi = torch.randn(N, device=dev)
    R = None if rrows == 0 else torch.randn(M if rrows < 0 else rrows, N, device=dev)
    out = torch.empty(M, N, device=dev)
    bufs.append((name, A, W, sc, bi, R, out, M, N, K, relu, rmod))
for r in range(reps):
    for name, A, W, sc, bi, R, out, M, N, K, relu, rmod in bufs:
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        rc = lib.spe_debug_gemm(0, p(A), p(W), M, N, K, p(sc), p(bi), p(R), rmod, relu, p(out), None)
        e1.record()
        torch.cuda.synchronize()
        by = (A.numel() + out.numel() + (R.numel() if R is not None and rmod == 0 else 0)) * 4
        print(f"rep {r} {name:16s} rc={rc} {e0.elapsed_time(e1) * 1e3:8.1f} us  {2 * M * N * K / e0.elapsed_time(e1) / 1e9:6.1f} TFLOP/s "
              f"{by / e0.elapsed_time(e1) / 1e6:6.0f} GB/s")
