"""B200-native crop -> keypoint-set predictor -> PnP path behind the reference's Python interface.

    from satellite_pose_estimation_b200 import build_model, build_solver
    model, _, postprocessors = build_model(args); model.to('cuda'); model.load_state_dict(ckpt['model'])
    solver = build_solver(args, model, postprocessors)

All compute runs in ``libspe.so`` (hand-written sm_100a CUDA, C ABI in include/spe.h).  No CPU fallback.
"""
from .models import B200DETR, PostProcess, build_model  # noqa: F401
from .solver import BatchedPoseSolver, MultiMeanPoseSolver, build_solver  # noqa: F401
from .engine import Engine  # noqa: F401
from .submission import (SpeedEval, SubmissionWriter, gen_prediction, gen_submission, run_image_set,  # noqa: F401
                         save_prediction, speed_score)
