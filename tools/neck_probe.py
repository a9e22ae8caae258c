"""Standalone launches of the two neck 3x3 convolutions at B = 64 (the largest GEMM launches of the step), for ncu
captures.   python tools/neck_probe.py [reps]"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satellite_pose_estimation_b200 import _lib
lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
dev = "cuda:0"
torch.manual_seed(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for name, NB, H, Cin, Cout in [("s16_latern", 64, 28, 1024, 256), ("output_conv", 64, 28, 512, 512)]:
    x = torch.randn(NB, H, H, Cin, device=dev)
    wk = torch.randn(Cout, 9 * Cin, device=dev) / (9 * Cin) ** 0.5
    b = torch.randn(Cout, device=dev)
    out = torch.empty(NB, H, H, Cout, device=dev)
    for r in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        rc = lib.spe_debug_conv(0, p(x), p(wk), NB, H, H, Cin, Cout, 3, 3, 1, 1, None, p(b), 0, p(out), None)
        e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3
    print(f"{name}: conv3x3 {Cin}->{Cout} @{H}x{H} B={NB}: rc={rc} {us:7.1f} us {2*NB*H*H*Cout*9*Cin/us/1e6:6.1f} TFLOP/s")
