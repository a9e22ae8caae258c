"""CPU restatement of set post-processing and the pose solvers (TEST INFRASTRUCTURE; see oracle/__init__.py).

* ``post_process``        PostProcess.forward                 RV/models/detr_speed.py:264-293
                          (+ sigmas = exp(pred_sigmas)         SA/src/zoo/rtdetr/rtdetr_postprocessor.py:53)
* ``assign``              find_index + per-label best query    RV/utils/speed_eval.py:152-162, :184-200
* ``SimplePoseSolver``    cv2 RANSAC-P3P -> ITERATIVE refine   RV/utils/speed_eval.py:143-242
* ``MultiMeanPoseSolver`` ensemble pooling + the same chain    RV/utils/speed_eval.py:42-140
* ``sigma_pnp``           sigma-weighted Huber LM              SA/utils/speed_eval.py:269-319 (ceres_pnp) -- the
                          cost functor is a private PyCeres build (absent): restated with scipy, PARITY UNPINNED
* ``self_assessment``     reject filter                        no reference code exists:
                          builder-defined spec, PARITY UNPINNED
* ``speed_score``         RV/utils/speed_eval.py:245-262

The third-party arithmetic is OpenCV's (reference pins opencv-python==4.4.0.44; this image has 4.13.0): the
oracle calls the very functions the reference calls.  ``mathutils.Matrix(R).to_quaternion()`` (absent) is
replaced by a standard rotation-matrix -> (w, x, y, z) conversion normalised to w >= 0.
"""
import numpy as np
import cv2
import torch
import torch.nn.functional as F

from .constants import CAMERA_K, CAMERA_DIST, TANGO_POINTS


def post_process(pred_logits, pred_points, clip_bbox, pred_sigmas=None):
    """PostProcess.forward, RV/models/detr_speed.py:275-291 (same fp32 torch ops, same order)."""
    out_logits = torch.as_tensor(pred_logits).to("cpu").float()
    out_points = torch.as_tensor(pred_points).to("cpu").float().clone()
    assert len(out_logits) == len(clip_bbox)
    prob = F.softmax(out_logits, -1)
    for pt, bbox in zip(out_points, clip_bbox):
        bbox = torch.as_tensor(bbox)
        width, height = bbox[2] - bbox[0], bbox[3] - bbox[1]
        x1, y1 = bbox[0], bbox[1]
        pt[:, 0] = pt[:, 0] * width + x1
        pt[:, 1] = pt[:, 1] * height + y1
    results = [{"logits": np.asarray(s), "points": np.asarray(p)} for s, p in zip(prob, out_points)]
    if pred_sigmas is not None:
        sig = torch.exp(torch.as_tensor(pred_sigmas).to("cpu").float())   # rtdetr_postprocessor.py:53
        for r, s in zip(results, sig):
            r["sigmas"] = np.asarray(s)
    return results


def assign(points, probs):
    """Query -> keypoint assignment, RV/utils/speed_eval.py:184-200.

    Returns (labels list in first-occurrence order, query index per label, image points [n,2])."""
    points = np.asarray(points); probs = np.asarray(probs)
    labels, scores = probs.argmax(1), probs.max(1)
    fg = np.nonzero(labels != probs.shape[1] - 1)[0]
    best = {}
    for qi in fg:                        # dict insertion order == first occurrence, like the reference's P
        l_ = int(labels[qi])
        if l_ not in best:
            best[l_] = []
        best[l_].append(qi)
    order, qidx = [], []
    for l_, qs in best.items():
        sc = np.asarray([scores[qi] for qi in qs])
        qidx.append(int(qs[int(sc.argmax())]))   # first max, like pt_score[:, -1].argmax()
        order.append(l_)
    pts = np.asarray([points[qi] for qi in qidx], dtype=np.float64).reshape(-1, 2)
    return order, qidx, pts


def assign_table(points, probs):
    """assignment as int32[11]: query index chosen for each keypoint label, -1 if the label is absent."""
    order, qidx, _ = assign(points, probs)
    tab = -np.ones(11, dtype=np.int32)
    for l_, qi in zip(order, qidx):
        tab[l_] = qi
    return tab


def rot_to_quat(R):
    """Rotation matrix -> unit quaternion (w, x, y, z), w >= 0 (stands in for mathutils, speed_eval.py:233)."""
    R = np.asarray(R, dtype=np.float64)
    tr = np.trace(R)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    q /= np.linalg.norm(q)
    return q if q[0] >= 0 else -q


class SimplePoseSolver:
    """RV/utils/speed_eval.py:143-242.  ``solver(points, probs) -> (quat wxyz, tvec)``; raises ``cv2.error`` /
    ``IndexError`` on failure exactly where the reference does (callers map that to the zero pose)."""

    def __init__(self, repro=20, return_inliers=False):
        self.W_Pt = TANGO_POINTS
        self.reprojectionError = repro
        self.return_inliers = return_inliers

    def __call__(self, points, logits):
        order, _, pts = assign(points, logits)
        obj_pts = pts[:, np.newaxis, :].astype(np.float32)                        # :203-204 (IndexError if empty)
        wld_pts = np.asarray([self.W_Pt[l_] for l_ in order])[:, np.newaxis, :].astype(np.float32)
        retval, rvec, tvec, inliers = cv2.solvePnPRansac(                        # :209-214
            wld_pts, obj_pts, CAMERA_K, CAMERA_DIST, useExtrinsicGuess=False, flags=cv2.SOLVEPNP_P3P,
            reprojectionError=self.reprojectionError)
        used = None
        if inliers is not None:                                                   # :219-230
            idx = inliers.flatten()
            retval, rvecs, tvecs, _ = cv2.solvePnPGeneric(
                wld_pts[idx], obj_pts[idx], CAMERA_K, CAMERA_DIST, useExtrinsicGuess=True, rvec=rvec, tvec=tvec,
                flags=cv2.SOLVEPNP_ITERATIVE)
            assert len(rvecs) == 1
            rvec, tvec = rvecs[0], tvecs[0]
            used = sorted(order[i] for i in idx)
        Rmat = cv2.Rodrigues(rvec)[0]                                             # :232
        quat = rot_to_quat(Rmat)
        if self.return_inliers:
            return quat.flatten(), np.asarray(tvec).flatten(), used
        return quat.flatten(), np.asarray(tvec).flatten()


class MultiMeanPoseSolver:
    """Multi_Mean_PoseSolver, RV/utils/speed_eval.py:42-140 (the ensemble solver of gen_submission_multi.py).
    ``solver(multi_points, multi_probs)``: one [Q,2] pixel array and one [Q,12] probability array per member."""

    def __init__(self, repro=25, return_details=False):
        self.W_Pt = TANGO_POINTS
        self.reprojectionError = repro
        self.return_details = return_details

    @staticmethod
    def mean_and_filter(obj_pts):
        """:57-75 -- per label: mean of all pooled predictions; with >= 3 of them, drop those whose distance to the
        mean is not below 3 std of the distances, and average the rest (an empty rest gives NaN, as in the reference)."""
        from scipy.spatial.distance import cdist
        result, counts = {}, {}
        for label, points in obj_pts.items():
            num_points = len(points)
            points = np.vstack(points)
            points_mean = np.mean(points, axis=0, keepdims=True)
            if num_points < 3:
                result[label] = points_mean.flatten()
                counts[label] = num_points
            else:
                distances = cdist(points, points_mean).flatten()
                std_dist = np.std(distances)
                inlier_index = distances < std_dist * 3
                with np.errstate(all="ignore"):
                    import warnings
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        result[label] = np.mean(points[inlier_index], axis=0).flatten()
                counts[label] = int(inlier_index.sum())
        return result, counts

    def pool(self, multi_points, multi_logits):
        """:77-92 -- every foreground query of every member, appended per label in (member, query) order."""
        from collections import defaultdict
        obj_pts_original = defaultdict(list)
        for points, logits in zip(multi_points, multi_logits):
            points = np.asarray(points); logits = np.asarray(logits)
            labels = logits.argmax(1)
            fg = labels != logits.shape[1] - 1
            for pt_, l_ in zip(points[fg], labels[fg]):
                obj_pts_original[int(l_)].append(pt_)
        return self.mean_and_filter(obj_pts_original)

    def __call__(self, multi_points, multi_logits):
        assert len(multi_points) == len(multi_logits)
        obj_pts_mean, counts = self.pool(multi_points, multi_logits)
        order = list(obj_pts_mean.keys())
        obj_pts = np.asarray([obj_pts_mean[l_][:2] for l_ in order])[:, np.newaxis, :].astype(np.float32)   # :94-102
        wld_pts = np.asarray([self.W_Pt[l_] for l_ in order])[:, np.newaxis, :].astype(np.float32)
        retval, rvec, tvec, inliers = cv2.solvePnPRansac(                                                   # :105-110
            wld_pts, obj_pts, CAMERA_K, CAMERA_DIST, useExtrinsicGuess=False, flags=cv2.SOLVEPNP_P3P,
            reprojectionError=self.reprojectionError)
        used = None
        if inliers is not None:                                                                             # :115-126
            idx = inliers.flatten()
            retval, rvecs, tvecs, _ = cv2.solvePnPGeneric(
                wld_pts[idx], obj_pts[idx], CAMERA_K, CAMERA_DIST, useExtrinsicGuess=True, rvec=rvec, tvec=tvec,
                flags=cv2.SOLVEPNP_ITERATIVE)
            assert len(rvecs) == 1
            rvec, tvec = rvecs[0], tvecs[0]
            used = sorted(order[i] for i in idx)
        quat = rot_to_quat(cv2.Rodrigues(rvec)[0])                                                          # :128-131
        if self.return_details:
            return quat.flatten(), np.asarray(tvec).flatten(), obj_pts_mean, counts, used
        return quat.flatten(), np.asarray(tvec).flatten()


def pooled_table(obj_pts_mean, counts):
    """pooled keypoints as float32 [11,2] (0 where absent) and the per-label counts int32 [11]."""
    pts = np.zeros((11, 2), dtype=np.float32); cnt = np.zeros(11, dtype=np.int32)
    for l_, p in obj_pts_mean.items():
        pts[l_] = p[:2]; cnt[l_] = counts[l_]
    return pts, cnt


def solve_or_zero(solver, points, probs):
    """Caller contract, RV/gen_submission_single.py:169-175: failures become the zero pose."""
    try:
        out = solver(points, probs)
        return out[0], out[1], True
    except (IndexError, cv2.error):
        return np.zeros(4), np.zeros(3), False


# -------------------------------------------------------------------------------------------------------------
# self-assessment variant (PARITY UNPINNED: no runnable reference code)
# -------------------------------------------------------------------------------------------------------------
def project(rvec, tvec, X):
    R = cv2.Rodrigues(np.asarray(rvec, dtype=np.float64).reshape(3, 1))[0]
    pc = X @ R.T + np.asarray(tvec, dtype=np.float64).reshape(1, 3)
    return pc[:, :2] / pc[:, 2:3]


def sigma_pnp(world, img_px, sigmas, rvec0, tvec0, huber=0.005, max_iter=20):
    """ceres_pnp, SA/utils/speed_eval.py:269-319: residual_i = [w_u (x/z - u), w_v (y/z - v)] in normalised image
    coordinates, w = 1/(sqrt(sigma)+1e-6) normalised to sum 1 per axis (:285-288), HuberLoss(0.005) on each
    2-vector residual block, LM from (rvec0, tvec0)."""
    from scipy.optimize import least_squares
    world = np.asarray(world, dtype=np.float64).reshape(-1, 3)
    img_px = np.asarray(img_px, dtype=np.float64).reshape(-1, 2)
    un = np.stack([(img_px[:, 0] - CAMERA_K[0, 2]) / CAMERA_K[0, 0],
                   (img_px[:, 1] - CAMERA_K[1, 2]) / CAMERA_K[1, 1]], 1)       # cv2.undistortPoints, dist = 0
    w = 1.0 / (np.sqrt(np.asarray(sigmas, dtype=np.float64).reshape(-1, 2)) + 1e-6)
    w = w / w.sum(0, keepdims=True)

    def cost_and_res(p):
        r = (project(p[:3], p[3:], world) - un) * w
        s = (r ** 2).sum(1)
        rho_scale = np.where(s <= huber ** 2, 1.0, np.sqrt(np.maximum(2 * huber * np.sqrt(s) - huber ** 2, 0)
                                                             / np.maximum(s, 1e-300)))
        return (r * rho_scale[:, None]).ravel()

    p0 = np.concatenate([np.asarray(rvec0, dtype=np.float64).ravel(), np.asarray(tvec0, dtype=np.float64).ravel()])
    sol = least_squares(cost_and_res, p0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=2000)
    return sol.x[:3], sol.x[3:]


class SigmaPoseSolver:
    """SimplePoseSolverSigma, SA/utils/speed_eval.py:322-420: same assignment, solvePnPRansac(EPNP, 25 px),
    then the sigma-weighted LM on the inliers."""

    def __init__(self, repro=25):
        self.reprojectionError = repro

    def __call__(self, points, logits, sigmas):
        order, qidx, pts = assign(points, logits)
        obj_pts = pts[:, np.newaxis, :].astype(np.float32)
        wld_pts = np.asarray([TANGO_POINTS[l_] for l_ in order])[:, np.newaxis, :].astype(np.float32)
        sig = np.asarray([np.asarray(sigmas)[qi] for qi in qidx], dtype=np.float64).reshape(-1, 2)
        retval, rvec, tvec, inliers = cv2.solvePnPRansac(
            wld_pts, obj_pts, CAMERA_K, CAMERA_DIST, useExtrinsicGuess=False, flags=cv2.SOLVEPNP_EPNP,
            reprojectionError=self.reprojectionError)
        used = None
        if inliers is not None:
            idx = inliers.flatten()
            rvec, tvec = sigma_pnp(wld_pts[idx, 0], obj_pts[idx, 0], sig[idx], rvec, tvec)
            used = sorted(order[i] for i in idx)
        Rmat = cv2.Rodrigues(np.asarray(rvec, dtype=np.float64).reshape(3, 1))[0]
        return rot_to_quat(Rmat), np.asarray(tvec).flatten(), used


def reproj_rms_px(quat, tvec, world, img_px):
    from .synth import quat_to_rot
    pc = np.asarray(world) @ quat_to_rot(quat).T + np.asarray(tvec).reshape(1, 3)
    uv = pc[:, :2] / pc[:, 2:3] * np.array([CAMERA_K[0, 0], CAMERA_K[1, 1]]) + CAMERA_K[:2, 2]
    return float(np.sqrt(((uv - np.asarray(img_px)) ** 2).sum(1).mean()))


def self_assessment(n_inliers, rms_px, mean_sigma_px, max_rms_px=5.0, max_sigma_px=12.0, min_inliers=4):
    """Builder-defined reject filter (DESIGN.md "self-assessment"): a pose is rejected when fewer than
    ``min_inliers`` keypoints support it, when the RMS reprojection error of its inliers exceeds ``max_rms_px``,
    or when the network's own mean predicted sigma (pixels) exceeds ``max_sigma_px``."""
    return (n_inliers < min_inliers) or (rms_px > max_rms_px) or (mean_sigma_px > max_sigma_px)


def speed_score(q_pr, t_pr, q_gt, t_gt):
    """RV/utils/speed_eval.py:245-262."""
    q_pr = np.asarray(q_pr, dtype=np.float64).flatten(); t_pr = np.asarray(t_pr, dtype=np.float64).flatten()
    q_gt = np.asarray(q_gt, dtype=np.float64).flatten(); t_gt = np.asarray(t_gt, dtype=np.float64).flatten()
    if q_pr[0] < 0:
        q_pr = -q_pr
    if q_gt[0] < 0:
        q_gt = -q_gt
    s_t = np.linalg.norm(t_pr - t_gt) / np.linalg.norm(t_gt)
    s_q = 2 * np.arccos(min(abs(float(np.dot(q_pr, q_gt))), 1.0))
    return s_t, s_q
