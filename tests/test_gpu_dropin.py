"""GPU: the zero-edit integration route — with cv2's three hot-path names patched, the
oracle port of the reference's helpers (which calls cv2.BFMatcher / cv2.triangulatePoints /
cv2.solvePnPRansac exactly as sfm.py does) runs on the engine and returns what it returns on OpenCV."""
import cv2
import numpy as np
import pytest

import sfm_mvs_b200 as sfm
from oracle import cvpath
from sfm_mvs_b200 import synth

pytestmark = pytest.mark.gpu


def test_patch_cv2_routes_the_reference_helpers(engine):
    scene = synth.orbit_scene(3, 700, seed=8)
    K = scene["K"]
    v0, v1 = scene["views"][0], scene["views"][1]
    P1 = K @ np.hstack([v0["R"], v0["t"]])
    P2 = K @ np.hstack([v1["R"], v1["t"]])
    ref_p0, ref_p1 = cvpath.match_keypoints(v0["kp"], v0["des"], v1["kp"], v1["des"])
    _, _, ref_cloud = cvpath.Triangulation(P1, P2, ref_p0, ref_p1)
    ref_err, ref_X, _ = cvpath.ReprojectionError(ref_cloud, ref_p1.T, np.hstack([v1["R"], v1["t"]]), K, 1)
    ref_R, ref_t, ref_pin, _, _ = cvpath.PnP(ref_X[:, 0, :], ref_p1, K, np.zeros((5, 1), np.float32), ref_p0, 0)

    sfm.set_default_context(engine)
    saved = sfm.patch_cv2(cv2)
    try:
        assert cv2.BFMatcher is sfm.BFMatcher
        p0, p1 = cvpath.match_keypoints(v0["kp"], v0["des"], v1["kp"], v1["des"])       # sfm.py:259-268 on the engine
        _, _, cloud = cvpath.Triangulation(P1, P2, p0, p1)                              # sfm.py:45-56 on the engine
        R, t, pin, _, _ = cvpath.PnP(ref_X[:, 0, :], p1, K, np.zeros((5, 1), np.float32), p0, 0)
    finally:
        sfm.unpatch_cv2(saved, cv2)
        sfm.set_default_context(None)
    assert cv2.BFMatcher is not sfm.BFMatcher
    assert np.array_equal(p0, ref_p0) and np.array_equal(p1, ref_p1)
    assert np.abs(cloud[:3] - ref_cloud[:3]).max() / np.abs(ref_cloud[:3]).max() < 1e-4 and np.all(cloud[3] == 1)
    assert np.array_equal(pin, ref_pin)                   # same inlier rows: the engine's minimal solver is OpenCV's, bit for bit
    assert np.abs(R - ref_R).max() < 1e-6 and np.abs(t - ref_t).max() < 1e-5


def test_config0_two_view_plumbing_on_the_real_pair(engine, golden):
    """BASELINE configs[0] — the reference's two-view initialisation, sfm.py:304-325, on the only real data shipped
    with it: SIFT keypoints / descriptors of image.jpg and of its warped copy (tests/golden/real_pair.npz, produced by
    the reference's own find_features).  Every stage runs on OpenCV and on the engine FROM THE SAME INPUTS (OpenCV's
    output of the stage before): matches, recovered pose and its mask, triangulated cloud, reprojection error and PnP
    inliers must agree — masks and matches exactly, numbers within the north_star bars.
    The essential matrix is the exception, and it is the fixture's doing: the second Gustav frame is not shipped, so
    view 2 is a plane-induced warp of view 1 — a planar scene, for which the five-point problem is degenerate (a
    family of essential matrices fits) and the polynomial's roots are ill-conditioned.  There the engine's root finder
    and OpenCV's return different, equally supported models (this pair: 1461 vs 1460 inliers of 1501), so that stage is
    held to "the same consensus up to 1 %"; on non-degenerate scenes the masks are bit-exact
    (tests/test_gpu_essential.py)."""
    g = golden("real_pair")
    K = synth.K_GUSTAV
    des0, des1 = g["des0"].astype(np.float32), g["des1"].astype(np.float32)
    D0 = np.zeros((5, 1), np.float32)
    ctx = engine
    # sfm.py:259-268
    pts0, pts1 = cvpath.match_keypoints(g["kp0"], des0, g["kp1"], des1)
    e0, e1 = sfm.match_keypoints(g["kp0"], des0, g["kp1"], des1, ctx=ctx)
    assert np.array_equal(pts0, g["pts0"]) and np.array_equal(e0, pts0) and np.array_equal(e1, pts1)
    # sfm.py:307
    E, mask = cv2.findEssentialMat(pts0, pts1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None)
    Ee, maske = sfm.findEssentialMat(pts0, pts1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None, ctx=ctx)
    both = np.logical_and(mask.ravel() == 1, maske.ravel() == 1).sum()
    assert abs(int(maske.sum()) - int(mask.sum())) <= 0.01 * mask.sum() and both >= 0.98 * mask.sum()
    Ue, Se, _ = np.linalg.svd(Ee[:3])
    assert abs(Se[0] - Se[1]) < 1e-9 * Se[0] and Se[2] < 1e-9 * Se[0]                # a valid essential matrix
    pts0, pts1 = pts0[mask.ravel() == 1], pts1[mask.ravel() == 1]
    # sfm.py:311 — cheirality keeps no point of this planar pair; the count (0) and the pose must agree
    n_pose, R, t, mask2 = cv2.recoverPose(E, pts0, pts1, K)
    n_e, Re, te, mask2e = sfm.recoverPose(E, pts0, pts1, K, ctx=ctx)
    assert n_e == n_pose and np.array_equal(mask2e, mask2)
    assert np.abs(Re - R).max() < 1e-9 and np.abs(te - t).max() < 1e-9
    # sfm.py:314-325 on the essential-matrix inliers
    Rt0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    Rt1 = np.hstack([R @ Rt0[:, :3], Rt0[:, 3:] + Rt0[:, :3] @ t])
    a, b, cloud = cvpath.Triangulation(K @ Rt0, K @ Rt1, pts0, pts1)
    _, _, cloude = sfm.Triangulation(K @ Rt0, K @ Rt1, pts0, pts1, K, ctx=ctx)
    # low parallax: depths up to ~1e5 units; compare each point against its own magnitude
    rel = np.abs(cloude[:3] - cloud[:3]).max(0) / np.abs(cloud[:3]).max(0)
    assert np.all(cloude[3] == 1) and np.median(rel) < 1e-6 and rel.max() < 1e-4
    err, X, _ = cvpath.ReprojectionError(cloud, b, Rt1, K, homogenity=1)
    erre, Xe, _ = sfm.ReprojectionError(cloud, b, Rt1, K, 1, ctx=ctx)
    assert abs(erre - err) <= 1e-9 * err and np.array_equal(Xe, X)
    Rp, tp, p_in, X_in, _ = cvpath.PnP(X, b, K, D0, a, initial=1)
    Rpe, tpe, p_ine, _, _ = sfm.PnP(X, b, K, D0, a, 1, ctx=ctx)
    assert np.array_equal(p_ine, p_in) and len(p_in) > 1000
    assert np.abs(Rpe - Rp).max() < 1e-5 and np.abs(tpe - tp).max() < 1e-4
