"""GPU parity: hot path 2 (DLT triangulation, reprojection error) and common_points."""
import cv2
import numpy as np
import pytest

import sfm_mvs_b200 as sfm
from oracle import cvpath, restated
from sfm_mvs_b200 import synth

pytestmark = pytest.mark.gpu
REL = 1e-4   # north_star tolerance for 3-D points and reprojection errors


def _two_view(n, seed, step=0.05, noise=0.4):
    rng = np.random.default_rng(seed)
    K = synth.K_GUSTAV
    R0, t0 = synth.orbit_pose(0.0)
    R1, t1 = synth.orbit_pose(step)
    X = np.c_[rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n), rng.uniform(5, 11, n)]
    x0, _ = synth.project(K, R0, t0, X)
    x1, _ = synth.project(K, R1, t1, X)
    x0 = (x0 + rng.normal(0, noise, x0.shape)).astype(np.float32)
    x1 = (x1 + rng.normal(0, noise, x1.shape)).astype(np.float32)
    return K, K @ np.hstack([R0, t0]), K @ np.hstack([R1, t1]), np.hstack([R1, t1]), x0, x1


def test_triangulation_golden(engine, golden):
    g = golden("geometry")
    a, b, cloud = sfm.Triangulation(g["P1"], g["P2"], g["x0"], g["x1"], g["K"], repeat=False, ctx=engine)
    ref = g["cloud"]
    assert cloud.shape == ref.shape and cloud.dtype == ref.dtype
    assert np.all(cloud[3] == 1.0)
    rel = np.abs(cloud[:3] - ref[:3]).max() / np.abs(ref[:3]).max()
    assert rel < REL, rel
    assert np.abs(cloud[:3] - ref[:3]).max() / np.abs(ref[:3]).max() < 2e-6   # what we actually expect
    err, Xc, proj = sfm.ReprojectionError(cloud, b, g["Rt1"], g["K"], 1, ctx=engine)
    assert abs(err - float(g["tri_err"])) <= REL * float(g["tri_err"])
    assert Xc.shape == g["tri_X"].shape


@pytest.mark.parametrize("n,seed,step", [(1, 0, 0.05), (7, 1, 0.05), (1000, 2, 0.04), (100000, 3, 0.02), (4096, 4, 0.002)])
def test_triangulatepoints_equals_cv2(engine, n, seed, step):
    K, P1, P2, Rt1, x0, x1 = _two_view(n, seed, step)
    ours = sfm.triangulatePoints(P1, P2, x0.T.copy(), x1.T.copy(), ctx=engine)
    ref = cv2.triangulatePoints(P1, P2, x0.T.copy(), x1.T.copy())
    assert ours.shape == ref.shape == (4, n) and ours.dtype == ref.dtype == np.float32
    # unit-norm homogeneous vectors, sign arbitrary in theory — cv2's Jacobi SVD yields one, ours must be the same
    a, b = ours / ours[3], ref / ref[3]
    rel = np.abs(a[:3] - b[:3]).max() / np.abs(b[:3]).max()
    assert rel < REL, rel
    assert np.allclose(np.linalg.norm(ours.astype(np.float64), axis=0), 1.0, atol=1e-6)
    assert np.array_equal(np.sign(ours[3]), np.sign(ref[3]))
    # the Nx1x2 form the API also accepts
    ours2 = sfm.triangulatePoints(P1, P2, x0.reshape(-1, 1, 2), x1.reshape(-1, 1, 2), ctx=engine)
    assert np.array_equal(ours, ours2)


def test_triangulatepoints_rejects_what_cv2_rejects(engine):
    K, P1, P2, Rt1, x0, x1 = _two_view(10, 0)
    with pytest.raises(sfm.error):
        sfm.triangulatePoints(P1, P2, x0, x1, ctx=engine)                     # (N,2)
    with pytest.raises(sfm.error):
        sfm.triangulatePoints(P1, P2, x0.T[:, :0].copy(), x1.T[:, :0].copy(), ctx=engine)   # N = 0
    with pytest.raises(sfm.error):
        sfm.triangulatePoints(P1[:, :3], P2, x0.T.copy(), x1.T.copy(), ctx=engine)


@pytest.mark.parametrize("n,seed", [(5, 0), (995, 1), (200000, 2)])
def test_reprojection_error_equals_reference_formula(engine, n, seed):
    K, P1, P2, Rt1, x0, x1 = _two_view(n, seed)
    a, b, cloud = cvpath.Triangulation(P1, P2, x0, x1)
    e_ref, X_ref, p_ref = cvpath.ReprojectionError(cloud, b, Rt1, K, 1)
    e, X, p = sfm.ReprojectionError(cloud, b, Rt1, K, 1, ctx=engine)
    assert np.array_equal(X, X_ref)                       # convertPointsFromHomogeneous, bit exact
    assert np.array_equal(p, p_ref)                       # float32 projections, bit exact
    assert abs(e - e_ref) <= 1e-9 * e_ref               # float64 sums in a different order
    e0_ref, _, p0_ref = cvpath.ReprojectionError(X_ref[:, 0, :], x1, Rt1, K, 0)
    e0, _, p0 = sfm.ReprojectionError(X_ref[:, 0, :], x1, Rt1, K, 0, ctx=engine)
    assert np.array_equal(p0, p0_ref) and abs(e0 - e0_ref) <= 1e-9 * e0_ref
    eb, _ = sfm.ba.ReprojectionError(X_ref[:, 0, :], b, Rt1, K, ctx=engine)   # ba.pyc L44-63
    assert abs(eb - e0_ref) <= 1e-9 * e0_ref


def test_common_points_golden_and_quirk(engine, golden):
    g = golden("geometry")
    i1, i2, tA, tB = sfm.common_points(g["cp_A"], g["cp_B"], g["cp_C"], ctx=engine)
    assert np.array_equal(i1, g["cp_i1"]) and np.array_equal(i2, g["cp_i2"])
    assert np.array_equal(tA, g["cp_tA"]) and np.array_equal(tB, g["cp_tB"])


@pytest.mark.parametrize("n1,n2,seed", [(3000, 5000, 0), (1, 1, 1), (50, 0, 2), (0, 40, 3)])
def test_common_points_equals_reference_semantics(engine, n1, n2, seed):
    rng = np.random.default_rng(seed)
    A = rng.uniform(0, 900, (n1, 2)).astype(np.float32)
    B = rng.uniform(0, 900, (n2, 2)).astype(np.float32)
    k = min(n1, n2) // 2
    if k:
        B[rng.permutation(n2)[:k]] = A[rng.permutation(n1)[:k]]
        B[0, 0] = A[min(3, n1 - 1), 0]                  # x-only coincidence
    Cc = rng.uniform(0, 900, (n2, 2)).astype(np.float32)
    i1, i2, tA, tB = sfm.common_points(A, B, Cc, ctx=engine)
    r1, r2, rA, rB = cvpath.common_points(A, B, Cc)
    assert np.array_equal(i1, r1.astype(np.int64).reshape(-1)) and np.array_equal(i2, r2.astype(np.int64).reshape(-1))
    assert np.array_equal(tA, rA) and np.array_equal(tB, rB)


@pytest.mark.parametrize("n,seed,dtype", [(800, 0, np.float32), (2500, 1, np.float32), (300, 2, np.float64), (50, 3, np.float32)])
def test_recover_pose_equals_cv2(engine, n, seed, dtype):
    """cv2.recoverPose (sfm.py:311): same R, t, count and per-point mask.  Synthetic two-view geometry with pixel
    noise, 10 % gross outliers (behind-camera / wrong-depth cases for the cheirality test) and an essential matrix
    from cv2.findEssentialMat, exactly the reference's sequence."""
    rng = np.random.default_rng(seed)
    K = synth.K_GUSTAV
    X = np.column_stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(4, 40, n)])
    rvec = np.array([0.03, -0.25, 0.02]); tvec = np.array([1.0, 0.1, 0.2])
    R_gt = cv2.Rodrigues(rvec)[0]
    def proj(Xc):
        return np.column_stack([K[0, 0] * Xc[:, 0] / Xc[:, 2] + K[0, 2], K[1, 1] * Xc[:, 1] / Xc[:, 2] + K[1, 2]])
    p1 = proj(X) + rng.normal(0, 0.3, (n, 2))
    p2 = proj(X @ R_gt.T + tvec) + rng.normal(0, 0.3, (n, 2))
    bad = rng.choice(n, n // 10, replace=False)
    p2[bad] += rng.uniform(-150, 150, (len(bad), 2))
    p1, p2 = p1.astype(dtype), p2.astype(dtype)
    E, emask = cv2.findEssentialMat(p1, p2, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None)
    E = E[:3]
    for use_mask in (False, True):
        if use_mask:
            rc, Rc, tc, mc = cv2.recoverPose(E, p1, p2, K, mask=emask.copy())
            ro, Ro, to, mo = sfm.recoverPose(E, p1, p2, K, mask=emask.copy(), ctx=engine)
        else:
            rc, Rc, tc, mc = cv2.recoverPose(E, p1, p2, K)
            ro, Ro, to, mo = sfm.recoverPose(E, p1, p2, K, ctx=engine)
        assert np.abs(Ro - Rc).max() < 1e-9 and np.abs(to - tc).max() < 1e-9
        assert ro == rc
        assert np.array_equal(mo.ravel() != 0, mc.ravel() != 0)
        rr, Rr, tr, mr = restated.recover_pose(E, p1, p2, K, mask=emask if use_mask else None)     # the oracle agrees too
        assert rr == ro and np.array_equal(mr, mo.ravel() != 0)
