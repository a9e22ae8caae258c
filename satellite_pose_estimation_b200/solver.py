"""Host-side mirror of the reference's pose-solver interface (RV/utils/speed_eval.py:21-22, :143-242;
SA/utils/speed_eval.py:27-28, :322-420).

``build_solver(args)`` returns a callable with the reference signature ``solver(points, logits) -> (quat, tvec)``
that raises ``IndexError`` on failure (the reference's callers map ``IndexError`` / ``cv2.error`` to the zero pose,
RV/gen_submission_single.py:169-175).  The batched form ``solver.solve_batch(...)`` is the fast path: one kernel
launch per batch instead of one cv2 call chain per image.  Everything runs in libspe.so; there is no cv2 / CPU
fallback here.
"""
import numpy as np
import torch

from .engine import Engine


class BatchedPoseSolver:
    def __init__(self, engine=None, reproj=20.0, weighted=False, reject=False, post=None, device=0, model=None):
        self._engine = engine
        self._model = model   # B200DETR whose context is shared once its first forward has created it
        self._device = device
        self.reprojectionError = float(reproj)
        self.weighted, self.reject = bool(weighted), bool(reject)
        self._post = post   # PostProcess whose batched launch already solved these images

    @property
    def engine(self):
        if self._model is not None and self._model.engine is not None:
            return self._model.engine
        if self._engine is None:
            # assign/PnP needs no model weights: a minimal context is enough
            self._engine = Engine(max_batch=1, device=self._device)
        return self._engine

    def solve_batch(self, logits, points, boxes, log_sigma=None):
        """cuda tensors [B,Q,12], [B,Q,2] (normalised), int [B,4] -> dict of cuda tensors (quat wxyz f64, tvec f64,
        status int32, assign int32 [B,11], inlier_mask)."""
        return self.engine.assign_pnp(logits, points, boxes, log_sigma=log_sigma, reproj=self.reprojectionError,
                                      weighted=self.weighted and log_sigma is not None, reject=self.reject)

    def __call__(self, points, logits, sigmas=None, crop_side=None):
        """Reference per-image signature: ``points`` [Q,2] in original-image pixels, ``logits`` [Q,12] class
        probabilities (PostProcess output).  Answers from the batch solve PostProcess already did when possible.
        A pose flagged by the self-assessment filter (status 3) is treated like a failed solve everywhere (here:
        ``IndexError``, i.e. the callers' zero pose).  ``crop_side`` (pixels) lets the filter's sigma criterion work on
        this path too: the reference signature carries no crop box, and sigmas are in crop units."""
        if self._post is not None:
            hit = self._post.pose_cache.get(id(points))
            if hit is not None and hit[0] is points:
                _, quat, tvec, status = hit
                if status != 0:
                    raise IndexError(f"pose solve failed (status {status})")
                return quat.copy(), tvec.copy()
        points = np.asarray(points, dtype=np.float32)
        probs = np.asarray(logits, dtype=np.float32)
        assert points.shape[0] == probs.shape[0], "[Solver]: num_queries!"
        dev = self.engine.device
        # post_processed: the probabilities are the scores (no second softmax: the assignment compares exactly the
        # numbers the reference's find_index compares) and the pixel points pass through box (0,0,1,1) unchanged
        lg = torch.from_numpy(probs)[None].to(dev)
        pt = torch.from_numpy(points)[None].to(dev)
        box = torch.tensor([[0, 0, 1, 1]], dtype=torch.int32, device=dev)
        ls = None
        if sigmas is not None:
            ls = torch.from_numpy(np.log(np.asarray(sigmas, dtype=np.float32)))[None].to(dev)
        r = self.engine.assign_pnp(lg, pt, box, log_sigma=ls, reproj=self.reprojectionError,
                                   weighted=self.weighted and ls is not None, reject=self.reject, post_processed=True,
                                   sigma_px_scale=float(crop_side or 0.0))
        status = int(r["status"].item())
        if status != 0:
            raise IndexError(f"pose solve failed (status {status})")
        return r["quat"][0].cpu().numpy(), r["tvec"][0].cpu().numpy()


class MultiMeanPoseSolver:
    """Mirror of ``Multi_Mean_PoseSolver`` (RV/utils/speed_eval.py:42-140), the ensemble solver of
    ``gen_submission_multi.py``: ``solver(multi_points, multi_logits) -> (quat, tvec)`` with one [Q,2] pixel-point
    array and one [Q,12] probability array per ensemble member; raises ``IndexError`` where the reference's callers
    expect a failure (RV/gen_submission_multi.py:166-171).  ``solve_batch`` is the fast path: all images of a batch and
    all members in one kernel launch, straight from the members' raw outputs."""

    def __init__(self, args=None, engine=None, reproj=None, device=0, model=None):
        self._engine = engine
        self._model = model
        self._device = device
        self.reprojectionError = float(reproj if reproj is not None else getattr(args, "repro", 25))

    @property
    def engine(self):
        if self._model is not None and self._model.engine is not None:
            return self._model.engine
        if self._engine is None:
            self._engine = Engine(max_batch=1, device=self._device)
        return self._engine

    def solve_batch(self, multi_logits, multi_points, boxes, want_pooled=False):
        """cuda tensors [Nm,B,Q,12] raw logits, [Nm,B,Q,2] normalised points, int [B,4] crop boxes."""
        return self.engine.ensemble_pnp(multi_logits, multi_points, boxes, reproj=self.reprojectionError,
                                        want_pooled=want_pooled)

    def __call__(self, multi_points, multi_logits):
        assert isinstance(multi_points, list) and isinstance(multi_logits, list)
        assert len(multi_points) == len(multi_logits)
        dev = self.engine.device
        pts = np.stack([np.asarray(p, dtype=np.float32) for p in multi_points])[:, None]       # [Nm,1,Q,2] pixels
        prb = np.stack([np.asarray(l, dtype=np.float32) for l in multi_logits])[:, None]       # [Nm,1,Q,12] probabilities
        # pixel coordinates pass through the kernel's de-normalisation unchanged with box (0,0,1,1); the labels are the
        # arg-max of the probabilities themselves
        lg = torch.from_numpy(prb).to(dev)
        box = torch.tensor([[0, 0, 1, 1]], dtype=torch.int32, device=dev)
        r = self.engine.ensemble_pnp(lg, torch.from_numpy(pts).to(dev), box, reproj=self.reprojectionError,
                                     post_processed=True)
        status = int(r["status"].item())
        if status != 0:
            raise IndexError(f"pose solve failed (status {status})")
        return r["quat"][0].cpu().numpy(), r["tvec"][0].cpu().numpy()


def build_solver(args, model=None, postprocessors=None):
    """Reference contract ``build_solver(args)`` (RV/utils/speed_eval.py:21-22).  Passing the ``model`` /
    ``postprocessors`` returned by ``build_model`` lets the solver share their context and batched results."""
    post = postprocessors["points"] if postprocessors else None
    dev = str(getattr(args, "device", "cuda"))
    idx = (torch.device(dev).index or 0) if dev.startswith("cuda") else 0
    return BatchedPoseSolver(reproj=float(getattr(args, "repro", 20)),
                             weighted=bool(getattr(args, "sigma_head", False)),
                             reject=bool(getattr(args, "self_assessment", False)), post=post, device=idx, model=model)
