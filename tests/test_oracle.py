"""CPU tests: pin the oracle (oracle/cvpath.py, oracle/restated.py) against the golden
fixtures produced by the reference's own functions (oracle/make_golden.py) and against
in-process cv2."""
import cv2
import numpy as np
import pytest

from oracle import cv_exact, cvpath, restated
from sfm_mvs_b200 import synth


# ----------------------------------------------------------------------------- matching
def test_real_pair_matches_reference_find_features(golden):
    g = golden("real_pair")
    des0, des1 = g["des0"].astype(np.float32), g["des1"].astype(np.float32)
    p0, p1 = cvpath.match_keypoints(g["kp0"], des0, g["kp1"], des1)
    assert np.array_equal(p0, g["pts0"]) and np.array_equal(p1, g["pts1"])
    # restated exact 2-NN + double-precision ratio test reproduce the same survivors
    idx, dist = restated.knn2_l2(des0, des1)
    good = restated.ratio_mask(dist)
    assert np.array_equal(g["kp0"][good], g["pts0"])
    assert np.array_equal(g["kp1"][idx[good, 0]], g["pts1"])


@pytest.mark.parametrize("nq,nt,seed", [(257, 301, 0), (1000, 900, 1), (64, 2, 2), (5, 1, 3)])
def test_restated_knn2_equals_cv2(nq, nt, seed):
    q, t, _ = synth.matching_pair(nq, nt, seed=seed)
    idx, dist = restated.knn2_l2(q, t)
    cidx, cdist = cvpath.knn2_arrays(q, t)
    assert np.array_equal(idx, cidx[:, :idx.shape[1]])
    assert np.array_equal(dist, cdist[:, :dist.shape[1]])


def test_knn2_ties_go_to_lower_train_index():
    q = synth.sift_like_descriptors(40, 5)
    t = np.vstack([q[::-1], q[::-1]])            # every train row duplicated
    idx, dist = restated.knn2_l2(q, t)
    cidx, cdist = cvpath.knn2_arrays(q, t)
    assert np.array_equal(idx, cidx) and np.array_equal(dist, cdist)
    assert np.all(idx[:, 0] < idx[:, 1]) and np.all(dist == 0)


# ----------------------------------------------------------------------------- geometry
def test_cvpath_geometry_equals_reference(golden):
    g = golden("geometry")
    a, b, cloud = cvpath.Triangulation(g["P1"], g["P2"], g["x0"], g["x1"])
    assert np.array_equal(cloud, g["cloud"])
    err, Xc, proj = cvpath.ReprojectionError(cloud, b, g["Rt1"], g["K"], 1)
    assert err == float(g["tri_err"]) and np.array_equal(proj, g["tri_proj"])
    R, t, p_in, X_in, _ = cvpath.PnP(g["pnp_X"], g["pnp_p"], g["K"], np.zeros((5, 1), np.float32),
                                     g["x0"], 0)
    assert np.array_equal(R, g["pnp_R"]) and np.array_equal(t, g["pnp_t"])
    assert np.array_equal(p_in, g["pnp_p_in"])
    i1, i2, tA, tB = cvpath.common_points(g["cp_A"], g["cp_B"], g["cp_C"])
    assert np.array_equal(i1, g["cp_i1"]) and np.array_equal(i2, g["cp_i2"])
    assert np.array_equal(tA, g["cp_tA"]) and np.array_equal(tB, g["cp_tB"])


def test_restated_triangulation(golden):
    g = golden("geometry")
    X = restated.triangulate_dlt(g["P1"], g["P2"], g["x0"].T, g["x1"].T)
    X = X / X[3]
    ref = g["cloud"]
    rel = np.abs(X[:3] - ref[:3]).max() / np.abs(ref[:3]).max()
    assert rel < 1e-6, rel


def test_restated_rodrigues_and_projection():
    rng = np.random.default_rng(0)
    for _ in range(50):
        r = rng.normal(0, 1.0, 3)
        R = restated.rodrigues_to_matrix(r)
        Rc, _ = cv2.Rodrigues(r)
        assert np.abs(R - Rc).max() < 1e-15
        rv = restated.rodrigues_to_vector(Rc)
        rc, _ = cv2.Rodrigues(Rc)
        assert np.abs(rv - rc.ravel()).max() < 1e-12
    X = rng.uniform(-3, 3, (20000, 3)).astype(np.float32); X[:, 2] += 8
    r = np.array([0.1, -0.3, 0.05]); t = np.array([0.2, -0.1, 0.4])
    pc, _ = cv2.projectPoints(X, r, t, synth.K_GUSTAV, None)
    pr = restated.project_pinhole(X, restated.rodrigues_to_matrix(r), t, synth.K_GUSTAV).astype(np.float32)
    assert np.array_equal(pr, pc[:, 0, :])          # float32-rounded output bit-identical


def test_restated_reproj_error(golden):
    g = golden("geometry")
    X = g["tri_X"][:, 0, :]
    e = restated.reproj_error(X, g["x1"], g["Rt1"], g["K"])
    assert abs(e - float(g["tri_err"])) <= 1e-9 * float(g["tri_err"])


# ----------------------------------------------------------------------------- PnP RANSAC
def _epnp(X5, p5, K=synth.K_GUSTAV):
    ok, r, t = cv2.solvePnP(X5, p5, K, np.zeros((5, 1), np.float32), flags=cv2.SOLVEPNP_EPNP)
    return ok, r.ravel(), t.ravel()


def _epnp_exact(X5, p5, K=synth.K_GUSTAV):
    """The oracle's own minimal solver (oracle/cv_epnp.c) handed on as OpenCV does: (cv2.Rodrigues(R), t)."""
    R, t, _ = cv_exact.epnp(X5, p5, K)
    return True, cv2.Rodrigues(R)[0].ravel(), t


def _pnp_problem(rng, n, K=synth.K_GUSTAV):
    R, t = synth.orbit_pose(rng.uniform(0, 0.5))
    X = np.c_[rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n), rng.uniform(5, 11, n)].astype(np.float32)
    uv, _ = synth.project(K, R, t, X.astype(np.float64))
    return X, (uv + rng.normal(0, rng.uniform(0.1, 2.0), uv.shape)).astype(np.float32)


@pytest.mark.parametrize("n", [2, 3, 4, 6, 12, 24])
def test_c_restatement_of_opencv_small_svd_is_bit_exact(n):
    """cv2 decomposes matrices with fewer than 25 rows with its own one-sided Jacobi (LAPACK only above): the C
    restatement reproduces w, U and Vt bit for bit — full-rank and rank-deficient (the 5-point M^T M case)."""
    rng = np.random.default_rng(n)
    for trial in range(40):
        A = rng.standard_normal((n, n))
        if trial % 2 and n >= 6:                      # rank n - 2, symmetric: the basis of the null space is rounding
            B = rng.standard_normal((n - 2, n))
            A = B.T @ B
        w, U, Vt = cv2.SVDecomp(A)
        w2, U2, Vt2 = cv_exact.svd(A)
        assert np.array_equal(w.ravel(), w2) and np.array_equal(U, U2) and np.array_equal(Vt, Vt2)
    A = rng.standard_normal((6, min(n, 5)))           # tall, as in the beta least squares
    w, U, Vt = cv2.SVDecomp(A)
    w2, U2, Vt2 = cv_exact.svd(A)
    assert np.array_equal(w.ravel(), w2) and np.array_equal(U, U2) and np.array_equal(Vt, Vt2)


def test_c_restatement_of_opencv_primitives_is_bit_exact():
    rng = np.random.default_rng(0)
    for _ in range(100):
        M = rng.standard_normal((10, 12))
        assert np.array_equal(cv_exact.mul_transposed(M), cv2.mulTransposed(M, True))
        A = rng.standard_normal((3, 3))
        assert np.array_equal(cv_exact.invert_svd(A), cv2.invert(A, flags=cv2.DECOMP_SVD)[1])
        for nc in (3, 4, 5):
            A, b = rng.standard_normal((6, nc)), rng.standard_normal(6)
            assert np.array_equal(cv_exact.solve_svd(A, b), cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_SVD)[1].ravel())


@pytest.mark.parametrize("n", [5, 6, 9, 16])
def test_c_restatement_of_opencv_epnp_is_bit_exact(n):
    """cv2.solvePnP(flags=EPNP) — the minimal solver inside cv2.solvePnPRansac (sfm.py:67) — against oracle/cv_epnp.c:
    the same rvec and tvec BIT FOR BIT, for the 5-point minimal case (2-dimensional null space) and for more points."""
    rng = np.random.default_rng(40 + n)
    D0 = np.zeros((5, 1), np.float32)
    for _ in range(300):
        X, p = _pnp_problem(rng, n)
        ok, rvec, tvec = cv2.solvePnP(X, p, synth.K_GUSTAV, D0, flags=cv2.SOLVEPNP_EPNP)
        R, t, info = cv_exact.epnp(X, p, synth.K_GUSTAV)
        assert ok and np.array_equal(t, tvec.ravel()) and np.array_equal(cv2.Rodrigues(R)[0], rvec)
        assert 1 <= info["N"] <= 3 and info["sweeps"] >= 2


def test_cv2_epnp_null_space_basis_is_decided_by_rounding():
    """Why the minimal solver has to be restated operation for operation: perturb ONE coordinate of the five points
    by one float32 ulp and cv2's own hypothesis moves by far more than the perturbation, because the basis of the
    2-dimensional null space of M^T M that the Jacobi sweeps end in is decided by rounding.  (Any solver that is not
    bit-identical therefore returns different hypotheses for most subsets, even though each is a valid EPnP pose.)"""
    rng = np.random.default_rng(7)
    D0 = np.zeros((5, 1), np.float32)
    moved, big = 0, 0
    for _ in range(200):
        X, p = _pnp_problem(rng, 5)
        _, r0, t0 = cv2.solvePnP(X, p, synth.K_GUSTAV, D0, flags=cv2.SOLVEPNP_EPNP)
        X2 = X.copy()
        X2[0, 0] = np.nextafter(X2[0, 0], np.float32(np.inf))
        _, r1, t1 = cv2.solvePnP(X2, p, synth.K_GUSTAV, D0, flags=cv2.SOLVEPNP_EPNP)
        d = np.abs(t1 - t0).max()
        moved += d > 0
        big += d > 1e-4                      # a one-ulp input change is ~1e-7 relative
        # ... while the restatement follows cv2 through the perturbation
        R2, t2, _ = cv_exact.epnp(X2, p, synth.K_GUSTAV)
        assert np.array_equal(t2, t1.ravel())
    assert moved >= 190 and big >= 20, (moved, big)


@pytest.mark.parametrize("seed", range(12))
def test_restated_ransac_loop_reproduces_cv2_mask(seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(30, 1500))
    K = synth.K_GUSTAV
    R, t = synth.orbit_pose(rng.uniform(0, 0.5))
    X = np.c_[rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n), rng.uniform(5, 11, n)].astype(np.float32)
    uv, _ = synth.project(K, R, t, X.astype(np.float64))
    p = (uv + rng.normal(0, rng.uniform(0.1, 2.0), uv.shape)).astype(np.float32)
    bad = rng.random(n) < rng.uniform(0, 0.6)
    p[bad] += rng.uniform(-80, 80, (int(bad.sum()), 2)).astype(np.float32)
    ok, rvec, tvec, inl = cv2.solvePnPRansac(X, p, K, np.zeros((5, 1), np.float32))
    for solver in (_epnp, _epnp_exact):          # cv2's minimal solver, and the oracle's own restatement of it
        res = restated.pnp_ransac(X, p, K, solver)
        assert res["ok"] == ok
        if ok:
            assert np.array_equal(np.nonzero(res["mask"])[0], inl[:, 0])


def test_golden_pnp_inliers(golden):
    g = golden("geometry")
    for solver in (_epnp, _epnp_exact):
        res = restated.pnp_ransac(g["pnp_X"], g["pnp_p"], g["K"], lambda a, b: solver(a, b, g["K"]))
        assert np.array_equal(np.nonzero(res["mask"])[0], g["pnp_inliers"][:, 0])


def test_subset_stream_and_update_rule():
    s = restated.ransac_subsets(50, 100)
    assert s.shape == (100, 5) and all(len(set(r)) == 5 for r in s.tolist())
    assert s.min() >= 0 and s.max() < 50
    assert restated.update_num_iters(0.99, 0.0, 5, 100) == 0
    assert restated.update_num_iters(0.99, 0.5, 5, 100) == 100
    assert restated.update_num_iters(0.99, 0.2, 5, 100) == 12


# ----------------------------------------------------------------------------- BA
def test_cvpath_ba_equals_reference(golden):
    g = golden("ba_small")
    res = cvpath.OptimReprojectionError(g["x"])
    assert np.array_equal(res, g["residual"])
    X, p, Rt = cvpath.BundleAdjustment(g["X0"], g["obs"], g["Rt"], g["K"], 0.5)
    assert np.allclose(X, g["ba_X"], rtol=0, atol=1e-9) and np.allclose(Rt, g["ba_Rt"], atol=1e-9)


def test_cvpath_track_residual_equals_reference(golden):
    """test.py:85-113 (the residual of the track pipeline's BundleAdjustment) against the fixture produced by the
    reference's own def, its global-`track` and column-stride quirks included."""
    g = golden("ba_tracks")
    res = cvpath.OptimReprojectionError_tracks(g["x"], g["cloud"].size, g["poses"].size, g["track"].size, int(g["img_tot"]), g["track"])
    assert np.array_equal(res, g["residual"])


def test_restated_ba_jacobian_equals_cv2_projectpoints():
    rng = np.random.default_rng(3)
    K = synth.K_GUSTAV
    cams = np.array([[0.05, -0.2, 0.1, 0.3, -0.1, 0.5], [1e-14, 0, 0, 0, 0, 0.2]])
    pts = np.c_[rng.uniform(-2, 2, (30, 2)), rng.uniform(5, 11, 30)]
    cam_idx = np.repeat([0, 1], 30).astype(np.int32)
    pt_idx = np.tile(np.arange(30), 2).astype(np.int32)
    obs = np.zeros((60, 2), np.float32)
    r, Jc, Jp = restated.ba_residual_jacobian(cams, pts, cam_idx, pt_idx, obs, K)
    for c in range(2):
        proj, J = cv2.projectPoints(pts, cams[c, :3], cams[c, 3:], K, None)
        sel = cam_idx == c
        assert np.allclose(r[sel], proj[:, 0, :], rtol=1e-12)
        Jcv = J.reshape(30, 2, 15)
        assert np.allclose(Jc[sel], Jcv[:, :, :6], rtol=1e-7, atol=1e-7)
        R = restated.rodrigues_to_matrix(cams[c, :3])
        assert np.allclose(Jp[sel], Jcv[:, :, 3:6] @ R, rtol=1e-9, atol=1e-9)


def test_chain_port_equals_reference_loop(golden):
    g = golden("chain")
    scene = synth.orbit_scene(int(g["n_views"]), int(g["n_pts"]), seed=int(g["seed"]))
    outs = cvpath.register_chain(scene)
    assert np.array_equal(np.array([o["Rt"] for o in outs]), g["Rt"])
    assert np.array_equal(np.array([o["err_new"] for o in outs]), g["err_new"])
    assert np.array_equal(np.array([o["err_pnp"] for o in outs]), g["err_pnp"])
    assert np.array_equal(np.vstack([o["X_new"] for o in outs]), g["X_new"])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_restated_recover_pose_equals_cv2(seed):
    """The numpy restatement of cv2.recoverPose (sfm.py:311) against cv2 itself: same pose, count and mask."""
    rng = np.random.default_rng(seed)
    K = synth.K_GUSTAV
    n = 600
    X = np.column_stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(4, 40, n)])
    R_gt = cv2.Rodrigues(np.array([0.03, -0.25, 0.02]))[0]
    t_gt = np.array([1.0, 0.1, 0.2])
    proj = lambda Y: np.column_stack([K[0, 0] * Y[:, 0] / Y[:, 2] + K[0, 2], K[1, 1] * Y[:, 1] / Y[:, 2] + K[1, 2]])
    p1 = (proj(X) + rng.normal(0, 0.3, (n, 2))).astype(np.float32)
    p2 = proj(X @ R_gt.T + t_gt) + rng.normal(0, 0.3, (n, 2))
    bad = rng.choice(n, n // 10, replace=False)
    p2[bad] += rng.uniform(-150, 150, (len(bad), 2))
    p2 = p2.astype(np.float32)
    E, emask = cv2.findEssentialMat(p1, p2, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None)
    E = E[:3]
    rc, Rc, tc, mc = cv2.recoverPose(E, p1, p2, K)
    ro, Ro, to, mo = restated.recover_pose(E, p1, p2, K)
    assert ro == rc and np.abs(Ro - Rc).max() < 1e-9 and np.abs(to - tc).max() < 1e-9
    assert np.array_equal(mo, mc.ravel() != 0)
    rc, Rc, tc, mc = cv2.recoverPose(E, p1, p2, K, mask=emask.copy())
    ro, Ro, to, mo = restated.recover_pose(E, p1, p2, K, mask=emask)
    assert ro == rc and np.array_equal(mo, mc.ravel() != 0)


def _e_close(a, b):
    """distance between two unit-norm essential matrices, sign-free"""
    return min(np.abs(a - b).max(), np.abs(a + b).max())


@pytest.mark.parametrize("seed", range(12))
def test_restated_five_point_equals_cv2(seed):
    """Nister's five-point solver restated (oracle/restated.py five_point) against OpenCV's: with exactly five
    correspondences cv2.findEssentialMat returns every model of the minimal sample, stacked (sfm.py:307's solver).
    Same number of models and the same matrices up to sign; the order is not comparable (see five_point)."""
    K = synth.K_GUSTAV
    p0, p1, _, _ = synth.two_view_pair(5, seed=100 + seed, noise=0.0, outliers=0.0)
    Ec, mc = cv2.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None)
    Ec = np.zeros((0, 3, 3)) if Ec is None else Ec.reshape(-1, 3, 3)
    q0 = (p0.astype(np.float64) - K[:2, 2]) / [K[0, 0], K[1, 1]]
    q1 = (p1.astype(np.float64) - K[:2, 2]) / [K[0, 0], K[1, 1]]
    Eo = restated.five_point(q0, q1)
    assert len(Eo) == len(Ec) and len(Eo) >= 1
    for e in Eo:
        assert min(_e_close(e, c) for c in Ec) < 1e-6
        # every model satisfies the five epipolar constraints and the cubic constraints
        x0 = np.column_stack([q0, np.ones(5)]); x1 = np.column_stack([q1, np.ones(5)])
        assert np.abs(np.einsum("ni,ij,nj->n", x1, e, x0)).max() < 1e-9
        assert abs(np.linalg.det(e)) < 1e-9
        assert np.abs(2 * e @ e.T @ e - np.trace(e @ e.T) * e).max() < 1e-9


@pytest.mark.parametrize("n,seed", [(60, 0), (400, 1), (1500, 2)])
def test_restated_find_essential_mat_equals_cv2(n, seed):
    """The whole RANSAC loop restated (RNG subsets, Sampson error in float32, accept / RANSACUpdateNumIters) against
    cv2.findEssentialMat with the reference's arguments (sfm.py:307): same inlier mask bit for bit, same E up to
    sign."""
    K = synth.K_GUSTAV
    p0, p1, _, _ = synth.two_view_pair(n, seed=seed)
    Ec, mc = cv2.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None)
    Eo, mo, info = restated.find_essential_mat(p0, p1, K, 0.999, 0.4)
    assert np.array_equal(mo, mc.ravel() != 0) and int(mo.sum()) == info["best_count"] > 5
    assert _e_close(Eo, Ec[:3]) < 1e-7


E_VARIANTS = [(500, 0, 0.99, 1.0, 50, np.float32), (800, 1, 0.9, 3.0, 1000, np.float32), (300, 2, 0.999, 0.4, 5, np.float64),
              (1000, 3, 0.5, 0.2, 1000, np.float32), (400, 4, 0.999, 0.05, 200, np.float64)]


@pytest.mark.parametrize("n,seed,prob,thr,max_iters,dtype", E_VARIANTS)
def test_restated_find_essential_mat_parameter_variants(n, seed, prob, thr, max_iters, dtype):
    """Other confidences, thresholds, iteration caps and float64 points than the reference's call: the stopping rule
    (RANSACUpdateNumIters against maxIters) and the threshold scaling by the mean focal length still give cv2's mask."""
    K = synth.K_GUSTAV
    p0, p1, _, _ = synth.two_view_pair(n, seed=seed, dtype=dtype)
    Ec, mc = cv2.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=prob, threshold=thr, maxIters=max_iters)
    Eo, mo, info = restated.find_essential_mat(p0, p1, K, prob, thr, max_iters)
    assert np.array_equal(mo, mc.ravel() != 0) and info["iters"] <= max_iters
    assert _e_close(Eo, Ec[:3]) < 1e-7


def _geometry_case(name):
    K = synth.K_GUSTAV
    r = np.random.default_rng(0)
    n = 600
    Xg = np.column_stack([r.uniform(-4, 4, n), r.uniform(-3, 3, n), r.uniform(4, 40, n)])
    Xp = np.column_stack([r.uniform(-4, 4, n), r.uniform(-3, 3, n), np.full(n, 10.0)])
    R = cv2.Rodrigues(np.array([0.03, -0.25, 0.02]))[0]
    t = np.array([1.0, 0.1, 0.2])
    X, R, t, noise, seed = {"planar": (Xp, R, t, 0.3, 1), "pure_rotation": (Xg, R, np.zeros(3), 0.3, 2),
                            "tiny_baseline": (Xg, R, t * 1e-3, 0.3, 3), "forward_motion": (Xg, np.eye(3), np.array([0, 0, 1.0]), 0.3, 4),
                            "high_noise": (Xg, R, t, 2.0, 5), "noise_free": (Xg, R, t, 0.0, 6)}[name]
    proj = lambda Y: np.column_stack([K[0, 0] * Y[:, 0] / Y[:, 2] + K[0, 2], K[1, 1] * Y[:, 1] / Y[:, 2] + K[1, 2]])
    g = np.random.default_rng(seed)
    p0 = (proj(X) + g.normal(0, noise, (n, 2))).astype(np.float32)
    p1 = (proj(X @ R.T + t) + g.normal(0, noise, (n, 2))).astype(np.float32)
    return K, p0, p1


@pytest.mark.parametrize("name", ["planar", "tiny_baseline", "forward_motion", "high_noise", "noise_free", "pure_rotation"])
def test_restated_find_essential_mat_geometries(name):
    """Scene geometries the five-point method is known for (planar scenes, forward motion, vanishing baseline) and the
    noise extremes: the restated loop still returns cv2's mask.  Pure rotation (t = 0) makes the essential matrix itself
    degenerate — the polynomial's roots are ill-conditioned, two implementations agree on E to ~1e-5 only and may
    differ on a borderline correspondence; that is the documented limit of parity for this call."""
    K, p0, p1 = _geometry_case(name)
    Ec, mc = cv2.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None)
    Eo, mo, info = restated.find_essential_mat(p0, p1, K, 0.999, 0.4)
    if name == "pure_rotation":
        assert int((mo != (mc.ravel() != 0)).sum()) <= 3 and _e_close(Eo, Ec[:3]) < 1e-3
    else:
        assert np.array_equal(mo, mc.ravel() != 0)
        assert _e_close(Eo, Ec[:3]) < 1e-5


@pytest.mark.parametrize("C", [1, 7, 60])
def test_restated_conjugate_gradients_equal_lapack(C):
    """oracle.restated.reduced_solve_pcg — the algorithm of the engine's LM-step solver (csrc/pcg.cu) — against LAPACK:
    to 1e-6 at the solver's own tolerance (1e-8), and at the tolerance an LM step uses (1e-5) a solution whose residual
    is below it and which needs fewer iterations; an indefinite system is reported, not solved."""
    rng = np.random.default_rng(200 + C)
    n = 6 * C
    B = rng.normal(size=(n, n + 8))
    S = (B @ B.T / n + 0.5 * np.eye(n)).astype(np.float32)
    g = rng.normal(size=n).astype(np.float32)
    S64 = np.tril(S).astype(np.float64) + np.tril(S, -1).T.astype(np.float64)
    ref = np.linalg.solve(S64, -g.astype(np.float64))
    x, solved, its = restated.reduced_solve_pcg(S, g)
    assert solved and 1 <= its <= 400
    assert np.abs(x - ref).max() <= 1e-6 * np.abs(ref).max()
    x5, solved5, its5 = restated.reduced_solve_pcg(S, g, tol=1e-5)
    assert solved5 and its5 <= its
    assert np.linalg.norm(-g - S64 @ x5) <= 1e-5 * np.linalg.norm(g) * (1 + 1e-6)
    assert np.abs(x5 - ref).max() <= 1e-3 * np.abs(ref).max()
    Sbad = S.copy()
    Sbad[n // 2, n // 2] = -1.0
    assert not restated.reduced_solve_pcg(Sbad, g)[1]


def test_restated_conjugate_gradients_on_a_damped_schur_system():
    """The same on what the LM step really solves: the damped reduced camera system of a small BA problem."""
    from sfm_mvs_b200 import synth
    pb = synth.ba_problem(8, 300, 4, seed=4)
    r, Jc, Jp = restated.ba_residual_jacobian(pb["cams0"], pb["pts0"], pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
    S, g, *_ = restated.ba_schur(r, Jc, Jp, pb["cam_idx"], pb["pt_idx"], 8, 300, lam=1e-3)
    ref = np.linalg.solve(S, -g)
    x, solved, its = restated.reduced_solve_pcg(S, g)
    assert solved and np.abs(x - ref).max() <= 1e-6 * np.abs(ref).max(), its
    x5, solved5, its5 = restated.reduced_solve_pcg(S, g, tol=1e-5)
    assert solved5 and its5 < its
