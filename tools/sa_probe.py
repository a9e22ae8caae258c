"""SA (RT-DETR) predictor forward at batch B (default 64), a few calls: the program ncu captures for the SA launch list.
    SPE_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum ... python tools/sa_probe.py 64"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satellite_pose_estimation_b200 import Engine
from satellite_pose_estimation_b200.sa_models import sa_param_specs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator().manual_seed(0)
sd = {}
for name, shape, kind in sa_param_specs():        # random weights straight from the layout table (no oracle import)
    if kind == "n":
        sd[name] = torch.zeros((), dtype=torch.int64)
    elif name.endswith("running_var"):
        sd[name] = torch.rand(shape, generator=g) * 0.2 + 1.0
    elif len(shape) == 4:
        sd[name] = torch.randn(shape, generator=g) * (2.0 / (shape[1] * shape[2] * shape[3])) ** 0.5
    elif len(shape) == 2:
        sd[name] = torch.randn(shape, generator=g) * (1.0 / shape[1]) ** 0.5
    elif name.endswith("norm.weight") or name.endswith(".1.weight"):
        sd[name] = torch.ones(shape)
    else:
        sd[name] = torch.randn(shape, generator=g) * 0.02
eng = Engine(input_size=256, num_queries=30, enc_layers=1, dec_layers=3, dim_feedforward=1024, backbone="rtdetr_r50vd",
             precision="tf32", has_sigma=True, max_batch=B)
eng.load_state_dict(sd)
x = torch.randn(B, 3, 256, 256, generator=g).cuda()
for _ in range(3):
    eng.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.forward(x)
e1.record()
torch.cuda.synchronize()
print(f"SA forward B={B}: {e0.elapsed_time(e1) / 5:.3f} ms")
eng.close()
