"""Timing of the 3x3 layers of layer1 / layer2 at B = 64 (SPE_CONV3_REUSE=0 selects the generic implicit GEMM)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satellite_pose_estimation_b200 import _lib
lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
dev = "cuda:0"
torch.manual_seed(0)
for NB, H, Cin, Cout in [(64, 56, 64, 64), (64, 28, 128, 128)]:
    x = torch.randn(NB, H, H, Cin, device=dev)
    wk = torch.randn(Cout, 9 * Cin, device=dev) / (9 * Cin) ** 0.5
    b = torch.randn(Cout, device=dev)
    out = torch.empty(NB, H, H, Cout, device=dev)
    for r in range(4):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        rc = lib.spe_debug_conv(0, p(x), p(wk), NB, H, H, Cin, Cout, 3, 3, 1, 1, None, p(b), 1, p(out), None)
        e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3
    print(f"conv3x3 {Cin}->{Cout} @{H}x{H} B={NB}: rc={rc} {us:7.1f} us {2*NB*H*H*Cout*9*Cin/us/1e6:6.1f} TFLOP/s")
