"""Stand-alone launches of the encoder Q|K|V projection (M=50176, N=768, K=256, batch-broadcast addend) for ncu."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satellite_pose_estimation_b200 import _lib
lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
dev = "cuda:0"
torch.manual_seed(0)
M, N, K = 50176, 768, 256
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / 16
sc = torch.ones(N, device=dev); bi = torch.zeros(N, device=dev); R = torch.randn(784, N, device=dev)
out = torch.empty(M, N, device=dev)
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    rc = lib.spe_debug_gemm(0, p(A), p(W), M, N, K, p(sc), p(bi), p(R), 784, 0, p(out), None)
    e1.record(); torch.cuda.synchronize()
    print(f"rep {r} rc={rc} {e0.elapsed_time(e1) * 1e3:.1f} us")
