"""Port of the reference's pipeline helpers, making the same OpenCV/SciPy calls.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is the CPU arm that
`bench.py --impl reference` / `cpu_baseline` time ("kind": "port") and the
behavioural oracle the `-m gpu` parity tests compare the CUDA engine against.
It is validated against the reference's own defs (AST-loaded from
/root/reference/sfm.py) by oracle/make_golden.py -> tests/golden/.

Every function cites the reference lines it follows.  Images, SIFT, GUI and file
I/O are not part of the hot path: keypoints/descriptors arrive as arrays.
"""
from __future__ import annotations

import contextlib
import io

import cv2
import numpy as np
from scipy.optimize import least_squares

RATIO = 0.70  # sfm.py:264, isfm.py:75


# --------------------------------------------------------------------------- matching
def knn_match(des0: np.ndarray, des1: np.ndarray):
    """sfm.py:259-260 — `cv2.BFMatcher().knnMatch(des0, des1, k=2)` (NORM_L2, no cross-check)."""
    return cv2.BFMatcher().knnMatch(des0, des1, k=2)


def knn2_arrays(des0: np.ndarray, des1: np.ndarray):
    """Array form of knn_match: (idx (n,2) i32, dist (n,2) f32) via cv2.batchDistance, which is
    what BFMatcher::knnMatchImpl calls; identical indices/distances, no DMatch objects."""
    dist, idx = cv2.batchDistance(des0, des1, cv2.CV_32F, K=2, normType=cv2.NORM_L2)
    return idx, dist


def ratio_filter(matches, ratio: float = RATIO):
    """sfm.py:262-265 — Lowe test in Python doubles, strict '<'."""
    return [m for m, n in matches if m.distance < ratio * n.distance]


def match_keypoints(kp0: np.ndarray, des0: np.ndarray, kp1: np.ndarray, des1: np.ndarray,
                    ratio: float = RATIO):
    """Matching half of find_features (sfm.py:259-268) on arrays: kp (n,2) f32 replace kp[i].pt."""
    good = ratio_filter(knn_match(des0, des1), ratio)
    q = np.fromiter((m.queryIdx for m in good), dtype=np.int64, count=len(good))
    t = np.fromiter((m.trainIdx for m in good), dtype=np.int64, count=len(good))
    return np.float32(kp0[q]), np.float32(kp1[t])


# --------------------------------------------------------------------------- geometry
def Triangulation(P1, P2, pts1, pts2, K=None, repeat=False):
    """sfm.py:45-56 — transpose to 2xN unless `repeat`, cv2.triangulatePoints, divide by w."""
    a = pts1 if repeat else pts1.T
    b = pts2 if repeat else pts2.T
    cloud = cv2.triangulatePoints(P1, P2, a, b)
    return a, b, cloud / cloud[3]


def ReprojectionError(X, pts, Rt, K, homogenity):
    """sfm.py:79-100 — Frobenius norm of (projected - observed) divided by N."""
    rvec, _ = cv2.Rodrigues(Rt[:3, :3])
    if homogenity == 1:
        X = cv2.convertPointsFromHomogeneous(X.T)
    proj, _ = cv2.projectPoints(X, rvec, Rt[:3, 3], K, distCoeffs=None)
    proj = np.float32(proj[:, 0, :])
    obs = np.float32(pts)
    total = cv2.norm(proj, obs.T if homogenity == 1 else obs, cv2.NORM_L2)
    return total / len(proj), X, proj


def PnP(X, p, K, d, p_0, initial):
    """sfm.py:60-76 — NB the 5th positional of solvePnPRansac is `rvec`, so all defaults apply."""
    if initial == 1:
        X = X[:, 0, :]
        p = p.T
        p_0 = p_0.T
    ok, rvec, t, inliers = cv2.solvePnPRansac(X, p, K, d, cv2.SOLVEPNP_ITERATIVE)
    R, _ = cv2.Rodrigues(rvec)
    if inliers is not None:
        sel = inliers[:, 0]
        p, X, p_0 = p[sel], X[sel], p_0[sel]
    return R, t, p, X, p_0


def common_points(pts1, pts2, pts3):
    """sfm.py:215-239 — association by float equality.  `np.where(pts2 == pts1[i])` is
    element-wise: a row of pts2 is a hit when its x *or* its y equals; first hit wins."""
    i1, i2 = [], []
    for i in range(pts1.shape[0]):
        rows = np.where(pts2 == pts1[i, :])[0]
        if rows.size:
            i1.append(i)
            i2.append(rows[0])
    keep = np.ones(pts2.shape[0], dtype=bool)
    keep[i2] = False
    return np.array(i1), np.array(i2), pts2[keep], pts3[keep]


# --------------------------------------------------------------------------- single-camera BA
def OptimReprojectionError(x):
    """sfm.py:104-136 — residual ((p - proj)^2).ravel()/N for x=[Rt 12 | K 9 | p (2,N) | X (N,3)].
    (The reference also prints the sum on every call, sfm.py:132; omitted.)"""
    Rt = x[0:12].reshape(3, 4)
    K = x[12:21].reshape(3, 3)
    rest = int(len(x[21:]) * 0.4)
    p = x[21:21 + rest].reshape(2, rest // 2).T
    X = x[21 + rest:].reshape(-1, 3)
    rvec, _ = cv2.Rodrigues(Rt[:3, :3])
    proj, _ = cv2.projectPoints(X, rvec, Rt[:3, 3], K, distCoeffs=None)
    return ((p - proj[:, 0, :]) ** 2).ravel() / len(p)


def BundleAdjustment(points_3d, temp2, Rtnew, K, r_error):
    """sfm.py:138-157 — least_squares(TRF, 2-point FD Jacobian, gtol=r_error) over everything."""
    x0 = np.hstack((Rtnew.ravel(), K.ravel(), temp2.ravel(), points_3d.ravel()))
    sol = least_squares(fun=OptimReprojectionError, x0=x0, gtol=r_error).x
    rest = int(len(sol[21:]) * 0.4)
    return (sol[21 + rest:].reshape(-1, 3), sol[21:21 + rest].reshape(2, rest // 2).T,
            sol[0:12].reshape(3, 4))


def OptimReprojectionError_tracks(x, cloud_len, poses_len, tracks_len, img_tot, track):
    """test.py:85-113 — the multi-view residual of the track pipeline: x = [K 9 | poses img_tot x 12 | cloud N x 3 |
    tracks N x 2 img_tot]; one value per (view, point), sqrt(dx^2 + dy^2) / (number of values), view-major.  The
    reference reads the module-level `track` instead of the `tracks` it unpacks and takes columns i, i + 1 for view i
    (test.py:99) — both reproduced, `track` passed explicitly."""
    K = x[0:9].reshape(3, 3)
    poses = x[9:9 + poses_len].reshape(img_tot, 12)
    cloud = x[9 + poses_len:9 + poses_len + cloud_len].reshape(cloud_len // 3, 3)
    error = []
    for i in range(img_tot):
        Rt = poses[i].reshape(3, 4)
        r, _ = cv2.Rodrigues(Rt[:3, :3])
        p = track[:, i:i + 2]
        proj, _ = cv2.projectPoints(cloud, r, Rt[:3, 3], K, distCoeffs=None)
        proj = proj[:, 0, :]
        error += [np.sqrt((p[idx][0] - proj[idx][0]) ** 2 + (p[idx][1] - proj[idx][1]) ** 2) for idx in range(len(p))]
    return np.array(error).ravel() / len(error)


# --------------------------------------------------------------------------- per-view loop
def bootstrap_two_views(scene):
    """State the reference holds when its loop starts (sfm.py:304-339), with the E-matrix/
    recoverPose initialisation (covered separately: restated.find_essential_mat / recover_pose and
    tests/test_gpu_essential.py) replaced by the scene's ground-truth second pose so that every arm of the
    loop comparison starts from identical bytes."""
    K = scene["K"]
    v0, v1 = scene["views"][0], scene["views"][1]
    Rt0 = np.hstack([v0["R"], v0["t"].reshape(3, 1)])
    Rt1 = np.hstack([v1["R"], v1["t"].reshape(3, 1)])
    P1, P2 = K @ Rt0, K @ Rt1
    pts0, pts1 = match_keypoints(v0["kp"], v0["des"], v1["kp"], v1["des"])
    pts0, pts1, X = Triangulation(P1, P2, pts0, pts1)
    err, X, _ = ReprojectionError(X, pts1, Rt1, K, homogenity=1)
    _, _, pts1, X, _ = PnP(X, pts1, K, np.zeros((5, 1), np.float32), pts0, initial=1)
    return dict(K=K, P1=P1, P2=P2, pts0=pts0, pts1=pts1, points_3d=X, err0=err)


def register_view(state, view_prev, view_new, first: bool):
    """One iteration of sfm.py:341-409 (imread/SIFT/GUI/colour lookup removed).

    Returns the new state plus what the parity tests compare: Rt of the new view, the two
    reprojection errors and the newly triangulated points."""
    K = state["K"]
    P1, P2 = state["P1"], state["P2"]
    pts0, pts1, points_3d = state["pts0"], state["pts1"], state["points_3d"]
    pts_, pts2 = match_keypoints(view_prev["kp"], view_prev["des"], view_new["kp"], view_new["des"])
    if not first:                                                         # sfm.py:348-352
        pts0, pts1, points_3d = Triangulation(P1, P2, pts0, pts1)
        pts1 = pts1.T
        points_3d = cv2.convertPointsFromHomogeneous(points_3d.T)[:, 0, :]
    with contextlib.redirect_stdout(io.StringIO()):
        i1, i2, temp1, temp2 = common_points(pts1, pts_, pts2)            # sfm.py:356
    com2, com_ = pts2[i2], pts_[i2]
    pnp_X, pnp_p = points_3d[i1], com2                                    # what the reference hands to solvePnPRansac
    Rot, trans, com2, X_in, com_ = PnP(pnp_X, com2, K, np.zeros((5, 1), np.float32),
                                       com_, initial=0)                  # sfm.py:362
    Rt = np.hstack((Rot, trans))
    Pnew = K @ Rt
    err_pnp, _, _ = ReprojectionError(X_in, com2, Rt, K, homogenity=0)    # sfm.py:368
    temp1, temp2, X_new = Triangulation(P2, Pnew, temp1, temp2)           # sfm.py:371
    err_new, X_new, _ = ReprojectionError(X_new, temp2, Rt, K, homogenity=1)
    new_state = dict(K=K, P1=P2.copy(), P2=Pnew.copy(), pts0=pts_.copy(), pts1=pts2.copy(),
                     points_3d=None)
    out = dict(Rt=Rt, err_pnp=err_pnp, err_new=err_new, X_new=X_new[:, 0, :], n_pnp=len(i1),
               n_inl=len(com2), n_match=len(pts_), pnp_X=pnp_X, pnp_p=pnp_p)
    return new_state, out


def register_chain(scene, n_views: int | None = None):
    """Run the reference loop over a synthetic scene; returns per-view outputs."""
    views = scene["views"] if n_views is None else scene["views"][:n_views]
    state = bootstrap_two_views(scene)
    outs = []
    for i in range(len(views) - 2):
        state, out = register_view(state, views[i + 1], views[i + 2], first=(i == 0))
        outs.append(out)
    return outs
