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
    assert np.abs(R - ref_R).max() < 1e-3 and np.abs(t - ref_t).max() < 1e-2      # engine EPnP: same pose, not bitwise
    assert abs(len(pin) - len(ref_pin)) <= max(2, 0.01 * len(ref_pin))
