"""3xTF32 on the convolution modes (tiled / im2col tensor maps, stride 1 / 2) and on GEMMs with BN scale + residual."""
import os, sys, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from satellite_pose_estimation_b200 import _lib
lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
dev = "cuda"
def rna(x):
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
def split(w):
    hi = rna(w); lo = rna(w - hi)
    return torch.cat([hi, lo], 1).contiguous()
rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
torch.manual_seed(0)
for (M, N, K) in ((8192, 256, 64), (8192, 64, 64), (8192, 64, 256), (2048, 512, 128), (128, 2048, 512), (60, 320, 256)):
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5
    sc = torch.rand(N, device=dev) + 0.5; bi = torch.randn(N, device=dev); R = torch.randn(M, N, device=dev)
    out = torch.full((M, N), float("nan"), device=dev)
    assert lib.spe_debug_gemm(2, p(A), p(split(W)), M, N, K, p(sc), p(bi), p(R), 0, 1, p(out), None) == 0, lib.spe_global_last_error()
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t() * sc.double() + bi.double() + R.double()).clamp_min(0)
    print("gemm x3", (M, N, K), "rel", f"{rel(out, ref):.2e}")
for (NB, H, Cin, Cout, R, stride) in ((2, 128, 64, 64, 3, 1), (2, 64, 64, 64, 3, 1), (2, 64, 128, 128, 3, 2), (2, 32, 128, 128, 3, 1), (2, 32, 256, 256, 3, 2),
                                      (2, 16, 256, 256, 3, 1), (2, 16, 512, 512, 3, 2), (2, 8, 512, 512, 3, 1), (3, 8, 128, 128, 3, 1), (2, 16, 128, 128, 3, 1)):
    x = torch.randn(NB, H, H, Cin, device=dev)
    w = torch.randn(Cout, Cin, R, R, device=dev) / (R * R * Cin) ** 0.5
    wk = w.permute(0, 2, 3, 1).reshape(Cout, R * R * Cin).contiguous()
    sc = torch.rand(Cout, device=dev) + 0.5; bi = torch.randn(Cout, device=dev)
    Ho = (H + 2 * (R // 2) - R) // stride + 1
    out = torch.full((NB, Ho, Ho, Cout), float("nan"), device=dev)
    rc = lib.spe_debug_conv(2, p(x), p(split(wk)), NB, H, H, Cin, Cout, R, R, R // 2, stride, p(sc), p(bi), 1, p(out), None)
    assert rc == 0, lib.spe_global_last_error()
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double(), None, stride=stride, padding=R // 2)
    ref = (ref * sc.double()[None, :, None, None] + bi.double()[None, :, None, None]).clamp_min(0).permute(0, 2, 3, 1)
    print("conv x3", (NB, H, Cin, Cout, R, stride), "rel", f"{rel(out, ref):.2e}", "nan" if torch.isnan(out).any() else "")
