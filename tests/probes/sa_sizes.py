"""SA predictor at other input sizes than the recipe's 256 (odd feature maps: 224 -> 28 / 14 / 7) against the oracle."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sa_model_ref, synth
from satellite_pose_estimation_b200 import Engine

for R, B in ((224, 3), (128, 2), (160, 1)):
    cfg = sa_model_ref.SaCfg(input_size=R)
    sd = synth.make_sa_state_dict(cfg, seed=0)
    x = torch.randn(B, 3, R, R, generator=torch.Generator().manual_seed(R))
    taps = {}
    ref = sa_model_ref.forward(sd, cfg, x, taps)
    eng = Engine(input_size=R, num_queries=30, enc_layers=1, dec_layers=3, dim_feedforward=1024, backbone="rtdetr_r50vd",
                 precision="tf32", has_sigma=True, max_batch=B)
    eng.load_state_dict(sd)
    o = eng.forward_sa(x.cuda(), topk_override=taps["topk"].to(torch.int32).cuda())
    torch.cuda.synchronize()
    print(R, B, "pts", (o["pred_pts"].cpu() - ref["pred_pts"]).abs().max().item() * 1748, "px; logits",
          (o["pred_logits"].cpu() - ref["pred_logits"]).abs().max().item(), "sigma",
          (o["pred_sigmas"].cpu() - ref["pred_sigmas"]).abs().max().item())
    eng.close()
