"""GPU parity: hot path 3b (BA residual/Jacobian evaluation, Schur system, LM iterations)."""
import numpy as np
import pytest

import sfm_mvs_b200 as sfm
from oracle import cvpath, restated
from sfm_mvs_b200 import synth

pytestmark = pytest.mark.gpu
REL = 1e-4


def _small(seed=0, n_cam=6, n_pt=300, opp=4):
    return synth.ba_problem(n_cam, n_pt, opp, seed=seed)


def _make(engine, pb):
    prob = sfm.BAProblem(engine, len(pb["cams0"]), len(pb["pts0"]), pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
    prob.set_params(pb["cams0"], pb["pts0"])
    return prob


def test_residuals_and_jacobians_equal_projectpoints_restatement(engine):
    pb = _small()
    prob = _make(engine, pb)
    out = prob.eval(0)
    r, Jc, Jp = restated.ba_residual_jacobian(pb["cams0"], pb["pts0"], pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
    assert np.abs(out["r"] - r).max() <= REL * np.abs(r).max()
    assert np.abs(out["r"] - r).max() < 2e-5                       # float32 storage of a ~1 px residual
    assert np.abs(out["Jc"] - Jc).max() <= 1e-5 * np.abs(Jc).max()
    assert np.abs(out["Jp"] - Jp).max() <= 1e-5 * np.abs(Jp).max()
    assert abs(out["cost"] - 0.5 * (r ** 2).sum()) <= 1e-9 * 0.5 * (r ** 2).sum()


def test_reference_residual_modes(engine, golden):
    """mode 1 = OptimReprojectionError (sfm.py:104-136) on the golden single-camera problem."""
    g = golden("ba_small")
    n = g["X0"].shape[0]
    Rt = g["Rt"]
    cam = np.concatenate([sfm.rodrigues_to_vector(Rt[:, :3]), Rt[:, 3]])
    prob = sfm.BAProblem(engine, 1, n, np.zeros(n, np.int32), np.arange(n, dtype=np.int32), g["obs"].T.copy(), g["K"])
    prob.set_params(cam.reshape(1, 6), g["X0"].reshape(n, 3).astype(np.float64))
    r1 = prob.eval(1, want_J=False)["r"]
    ref = g["residual"].reshape(n, 2)
    assert np.abs(r1 - ref).max() <= REL * np.abs(ref).max()
    r2 = prob.eval(2, want_J=False)["r"]
    r0 = prob.eval(0, want_J=False)["r"].astype(np.float64)
    assert np.allclose(r2, np.sqrt((r0 ** 2).sum(1)) / n, rtol=1e-5)


def test_schur_system_equals_oracle(engine):
    pb = _small(seed=1, n_cam=5, n_pt=120, opp=3)
    prob = _make(engine, pb)
    lam = 1e-3
    S, g, hd = prob.build_system(lam)
    r, Jc, Jp = restated.ba_residual_jacobian(pb["cams0"], pb["pts0"], pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
    S_ref, g_ref, Hcc, bc, Hpp, bp = restated.ba_schur(r, Jc, Jp, pb["cam_idx"], pb["pt_idx"], 5, 120, lam)
    L = np.tril(np.ones_like(S_ref, dtype=bool))
    # block-lower triangle is what the solver reads
    blk = np.kron(np.tril(np.ones((5, 5), bool)), np.ones((6, 6), bool))
    scale = np.abs(S_ref).max()
    assert np.abs(S[blk] - S_ref[blk]).max() <= 2e-5 * scale, np.abs(S[blk] - S_ref[blk]).max() / scale
    assert np.abs(g - g_ref).max() <= 2e-5 * np.abs(g_ref).max()
    assert np.allclose(hd, np.array([np.diag(h) for h in Hcc]).ravel(), rtol=1e-5)


@pytest.mark.parametrize("C", [1, 7, 11, 50, 200, 500])
def test_reduced_solve_equals_dense_factorisation(engine, C):
    """The tile-Cholesky graph + flag-driven back substitution (csrc/solve.cu) on its own: S x = -g against LAPACK on
    the same float32 data, from one tile (n = 6) to BASELINE configs[3]'s 500 cameras (n = 3000, 47 tile columns)."""
    rng = np.random.default_rng(C)
    n = 6 * C
    B = rng.normal(size=(n, n + 8))
    S = (B @ B.T / n + 0.5 * np.eye(n)).astype(np.float32)
    S = np.tril(S) + np.tril(S, -1).T                      # what the solver reads: the lower triangle
    g = rng.normal(size=n).astype(np.float32)
    x, info = engine.reduced_solve(S, g)
    ref = np.linalg.solve(S.astype(np.float64), -g.astype(np.float64))
    assert info == 0
    assert np.abs(x - ref).max() <= 1e-10 * np.abs(ref).max(), np.abs(x - ref).max() / np.abs(ref).max()
    # not positive definite: reported, not silently solved
    Sbad = S.copy()
    k = n // 2
    Sbad[k, k] = -1.0
    _, info_bad = engine.reduced_solve(Sbad, g)
    assert 1 <= info_bad <= k + 1


@pytest.mark.parametrize("C", [1, 7, 50, 200, 500])
def test_reduced_solve_by_conjugate_gradients(engine, C):
    """The LM step's default linear solver (csrc/pcg.cu: block-Jacobi preconditioned conjugate gradients in one persistent
    kernel) against LAPACK on the same float32 data; a system that is not positive definite is reported as unsolved
    (the LM step then runs the factorisation)."""
    rng = np.random.default_rng(100 + C)
    n = 6 * C
    B = rng.normal(size=(n, n + 8))
    S = (B @ B.T / n + 0.5 * np.eye(n)).astype(np.float32)
    S = np.tril(S) + np.tril(S, -1).T
    g = rng.normal(size=n).astype(np.float32)
    x, solved, its = engine.reduced_solve(S, g, method="pcg")
    ref = np.linalg.solve(S.astype(np.float64), -g.astype(np.float64))
    assert solved and 1 <= its <= 400
    assert np.abs(x - ref).max() <= 1e-6 * np.abs(ref).max(), (its, np.abs(x - ref).max() / np.abs(ref).max())
    # the same system twice gives the same bits (fixed summation order, no atomics)
    x2, _, its2 = engine.reduced_solve(S, g, method="pcg")
    assert its2 == its and np.array_equal(x, x2)
    Sbad = S.copy()
    Sbad[n // 2, n // 2] = -1.0
    _, solved_bad, _ = engine.reduced_solve(Sbad, g, method="pcg")
    assert not solved_bad


def test_lm_iterations_reduce_cost_like_dense_gauss_newton(engine):
    pb = _small(seed=2, n_cam=8, n_pt=400, opp=5)
    prob = _make(engine, pb)
    c0 = prob.eval(0, want_r=False, want_J=False)["cost"]
    hist = prob.solve(max_iters=15, lam=1e-3)
    assert all(h["solve_info"] == 0 for h in hist)
    c1 = prob.eval(0, want_r=False, want_J=False)["cost"]
    assert c1 < 0.05 * c0
    # at the optimum the cost is about the noise floor: 0.5 * sigma^2 * (2*O - dof)
    O = len(pb["cam_idx"])
    floor = 0.5 * 0.25 * (2 * O - (6 * 8 + 3 * 400))
    assert c1 < 1.6 * floor, (c1, floor)
    cams, pts = prob.get_params()
    assert np.isfinite(cams).all() and np.isfinite(pts).all()
    # rejected steps leave the parameters untouched
    st = prob.gn_step(1e-15)
    c2 = prob.eval(0, want_r=False, want_J=False)["cost"]
    assert c2 <= c1 * (1 + 1e-9) or not st["accepted"]


def test_first_step_matches_dense_normal_equations(engine):
    """One damped step against the oracle's dense solve of the same normal equations."""
    pb = _small(seed=3, n_cam=4, n_pt=60, opp=3)
    prob = _make(engine, pb)
    lam = 1e-2
    r, Jc, Jp = restated.ba_residual_jacobian(pb["cams0"], pb["pts0"], pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
    C, P, O = 4, 60, len(pb["cam_idx"])
    J = np.zeros((2 * O, 6 * C + 3 * P))
    for o in range(O):
        c, p = pb["cam_idx"][o], pb["pt_idx"][o]
        J[2 * o:2 * o + 2, 6 * c:6 * c + 6] = Jc[o]
        J[2 * o:2 * o + 2, 6 * C + 3 * p:6 * C + 3 * p + 3] = Jp[o]
    H = J.T @ J
    H += lam * np.diag(np.diag(H))
    dx = -np.linalg.solve(H, J.T @ r.ravel())
    st = prob.gn_step(lam)
    assert st["accepted"] and st["solve_info"] == 0
    cams, pts = prob.get_params()
    d_c = (cams - pb["cams0"]).ravel()
    d_p = (pts - pb["pts0"]).ravel()
    assert np.abs(d_c - dx[:6 * C]).max() <= 2e-3 * np.abs(dx[:6 * C]).max()
    assert np.abs(d_p - dx[6 * C:]).max() <= 2e-3 * np.abs(dx[6 * C:]).max()


def test_lm_step_falls_back_to_the_factorisation(engine, monkeypatch):
    """When the conjugate-gradient solver reports failure (here: forced by allowing it one iteration) the tile Cholesky
    behind it produces the step — same answer as the dense solve, solve_info 0.  (An LM step stops the iteration at a
    relative residual of 1e-5; for the comparison with the factorisation it is solved to 1e-8.)"""
    monkeypatch.setenv("SFM_PCG_MAX_ITER", "1")
    monkeypatch.setenv("SFM_PCG_MIN_N", "6")           # (systems this small are factored directly by default)
    monkeypatch.setenv("SFM_BA_CG_TOL", "1e-8")
    pb = _small(seed=5, n_cam=12, n_pt=300, opp=4)
    prob = _make(engine, pb)
    st = prob.gn_step(1e-3)
    monkeypatch.delenv("SFM_PCG_MAX_ITER")
    prob2 = _make(engine, pb)                          # conjugate gradients this time
    st2 = prob2.gn_step(1e-3)
    assert st["solve_info"] == 0 and st2["solve_info"] == 0 and st["accepted"] and st2["accepted"]
    c1, p1 = prob.get_params()
    c2, p2 = prob2.get_params()
    # (the two systems are accumulated separately with float32 atomics: they differ by rounding, and so do the steps)
    assert np.abs(c1 - c2).max() <= 2e-3 * np.abs(c2 - pb["cams0"]).max()
    assert abs(st["cost_after"] - st2["cost_after"]) <= 1e-4 * st2["cost_after"]
    # the step an LM iteration really takes (default tolerance): accepted, and as good as the exact one to a percent
    monkeypatch.delenv("SFM_BA_CG_TOL")
    prob3 = _make(engine, pb)
    st3 = prob3.gn_step(1e-3)
    assert st3["solve_info"] == 0 and st3["accepted"]
    assert abs(st3["cost_after"] - st2["cost_after"]) <= 1e-2 * st2["cost_after"]


def test_reference_formulation_residual_and_fd_jacobian(engine, golden):
    """sfm_ba_reference_fd: OptimReprojectionError (sfm.py:104-136) at the golden x — against the value the reference's
    own def produced — and its finite-difference Jacobian against scipy's approx_derivative over the oracle's port."""
    from scipy.optimize._numdiff import approx_derivative
    g = golden("ba_small")
    n = g["X0"].shape[0]
    f0, J = engine.ba_reference_fd(g["x"], n)
    assert np.abs(f0 - g["residual"]).max() <= 1e-12 * np.abs(g["residual"]).max()
    J_ref = approx_derivative(cvpath.OptimReprojectionError, g["x"], method="2-point", f0=cvpath.OptimReprojectionError(g["x"]))
    assert J.shape == J_ref.shape == (2 * n, 21 + 5 * n)
    assert np.abs(J - J_ref).max() <= 1e-6 * np.abs(J_ref).max()
    assert np.array_equal(J != 0, J_ref != 0) or np.abs(J - J_ref)[(J != 0) != (J_ref != 0)].max() < 1e-9


def test_bundleadjustment_equals_the_reference_result(engine, golden):
    """sfm.py:138-157 through the drop-in: same formulation (pose as 12 numbers, K and observations free), same
    optimiser (scipy TRF, gtol = r_error), residual and Jacobian evaluated on the GPU — against ba_X / ba_p / ba_Rt as
    the reference's own BundleAdjustment returned them (tests/golden/ba_small.npz), and against the oracle's port on
    a second problem."""
    g = golden("ba_small")
    X, p, Rt = sfm.BundleAdjustment(g["X0"], g["obs"], g["Rt"], g["K"], 0.5, ctx=engine)
    n = g["X0"].shape[0]
    assert X.shape == (n, 3) and p.shape == (n, 2) and Rt.shape == (3, 4)
    assert np.abs(g["ba_X"] - g["X0"][:, 0]).max() > 1e-3          # the reference result is not the input ...
    assert np.abs(X - g["ba_X"]).max() <= 1e-4 * np.abs(g["ba_X"]).max()
    assert np.abs(p - g["ba_p"]).max() <= 1e-4 * np.abs(g["ba_p"]).max()
    assert np.abs(Rt - g["ba_Rt"]).max() <= 1e-4
    rng = np.random.default_rng(5)
    K = synth.K_GUSTAV
    R, t = synth.orbit_pose(0.11)
    Xw = np.c_[rng.uniform(-2, 2, 40), rng.uniform(-1.5, 1.5, 40), rng.uniform(5, 11, 40)]
    uv, _ = synth.project(K, R, t, Xw)
    obs = (uv + rng.normal(0, 0.7, uv.shape)).astype(np.float32).T.copy()
    X0 = (Xw + rng.normal(0, 0.03, Xw.shape)).astype(np.float32).reshape(40, 1, 3)
    Rt0 = np.hstack([R, t])
    Xr, pr, Rtr = cvpath.BundleAdjustment(X0, obs, Rt0, K, 0.5)
    Xe, pe, Rte = sfm.BundleAdjustment(X0, obs, Rt0, K, 0.5, ctx=engine)
    assert np.abs(Xe - Xr).max() <= 1e-4 * np.abs(Xr).max() and np.abs(pe - pr).max() <= 1e-4 * np.abs(pr).max()
    assert np.abs(Rte - Rtr).max() <= 1e-4


def test_bundleadjustment_wrappers(engine, golden):
    g = golden("ba_small")
    X, p, Rt = sfm.BundleAdjustmentSE3(g["X0"], g["obs"], g["Rt"], g["K"], ctx=engine)
    n = g["X0"].shape[0]
    assert X.shape == (n, 3) and p.shape == (n, 2) and Rt.shape == (3, 4)
    e_before = cvpath.ReprojectionError(g["X0"].reshape(n, 3), g["obs"].T, g["Rt"], g["K"], 0)[0]
    e_after = cvpath.ReprojectionError(np.float32(X), g["obs"].T, Rt, g["K"], 0)[0]
    assert e_after < e_before
    X2, R2, t2 = sfm.ba.bundle_adjustment(g["X0"].reshape(n, 3), g["obs"].T, None, g["K"], g["Rt"][:, :3], g["Rt"][:, 3:])
    assert np.allclose(X2, X, atol=1e-6) and R2.shape == (3, 3) and t2.shape == (3, 1)


def test_track_pipeline_residual_mode_equals_the_reference(engine, golden):
    """mode 2 of the evaluation kernel = the residual of test.py:85-113, one value per (view, point),
    sqrt(dx^2 + dy^2) / n_obs — against the fixture produced by the reference's own def (which compares view i with
    columns i, i + 1 of its global track table: the observations are built the same way)."""
    g = golden("ba_tracks")
    V, n = int(g["img_tot"]), g["cloud"].shape[0]
    cams = np.array([np.concatenate([sfm.rodrigues_to_vector(P.reshape(3, 4)[:, :3]), P.reshape(3, 4)[:, 3]]) for P in g["poses"]])
    cam_idx = np.tile(np.arange(V, dtype=np.int32), n)                      # point-major: point p, views 0..V-1
    pt_idx = np.repeat(np.arange(n, dtype=np.int32), V)
    obs = np.array([g["track"][pt, v:v + 2] for pt in range(n) for v in range(V)], np.float32)
    prob = sfm.BAProblem(engine, V, n, cam_idx, pt_idx, obs, g["K"])
    prob.set_params(cams, g["cloud"])
    r2 = prob.eval(2, want_J=False)["r"].reshape(n, V)
    ref = g["residual"].reshape(V, n).T                                     # the reference is view-major
    assert np.abs(r2 - ref).max() <= 1e-4 * np.abs(ref).max()
    prob.close()


def test_fixture_problem_from_the_reference_artifacts(engine, golden):
    """The reference's own reconstruction as a BA problem: the 57 cameras of pose.csv and the 19 282
    points of Point_Cloud/sparse.ply (tests/golden/gustav_scene.npz, written by oracle/make_golden.py through the
    product's pose.csv / PLY readers), 1 061 813 observations.  K5 against the numpy restatement on a sample, the
    cost against the float64 sum of the residuals, and LM from the perturbed start back to the noise floor."""
    g = golden("gustav_scene")
    pb = synth.ba_problem_from_scene(g["K"], g["cams"], g["pts"])
    assert len(pb["obs"]) == 1_061_813 and len(pb["pts0"]) == 19_282 and len(pb["cams0"]) == 57
    prob = sfm.BAProblem(engine, 57, 19_282, pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
    prob.set_params(pb["cams0"], pb["pts0"])
    out = prob.eval(0)
    sel = np.random.default_rng(2).choice(len(pb["obs"]), 3000, replace=False)
    rr, Jc, Jp = restated.ba_residual_jacobian(pb["cams0"], pb["pts0"], pb["cam_idx"][sel], pb["pt_idx"][sel], pb["obs"][sel], pb["K"])
    assert np.abs(out["r"][sel] - rr).max() <= 1e-4 * np.abs(rr).max()
    assert np.abs(out["Jc"][sel] - Jc).max() <= 1e-5 * np.abs(Jc).max()
    r = out["r"].astype(np.float64)
    assert abs(out["cost"] - 0.5 * (r * r).sum()) <= 1e-6 * out["cost"]
    hist = prob.solve(max_iters=15)
    noise_floor = 0.5 * 2 * len(pb["obs"]) * 0.5 ** 2                  # 0.5 px noise on both coordinates
    assert hist[-1]["cost_after"] < 1.05 * noise_floor < 0.2 * hist[0]["cost_before"]
    cams, pts = prob.get_params()
    # gauge freedom (a similarity transform) aside, the scene comes back: compare reprojection, not coordinates
    assert np.sqrt(2 * hist[-1]["cost_after"] / (2 * len(pb["obs"]))) < 0.52
    prob.close()


def test_ba_create_rejects_bad_input(engine):
    pb = _small()
    with pytest.raises(sfm.error):
        sfm.BAProblem(engine, 6, 300, pb["cam_idx"], pb["pt_idx"][::-1].copy(), pb["obs"], pb["K"])   # not point-major
    with pytest.raises(sfm.error):
        sfm.BAProblem(engine, 2, 300, pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])                 # cam index out of range


def test_full_size_problem_by_properties(engine):
    """BASELINE configs[3] (500 cameras / 100k points / 1M observations): properties instead of the oracle —
    K5's cost equals the float64 sum of its own residuals, the Jacobian blocks of a sample of observations equal the
    numpy restatement, and LM iterations decrease the cost monotonically on accepted steps."""
    import torch
    pb = synth.ba_problem(500, 100_000, 10, seed=0)
    prob = sfm.BAProblem(engine, 500, 100_000, pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
    prob.set_params(pb["cams0"], pb["pts0"])
    out = prob.eval(0)
    r = out["r"].astype(np.float64)
    assert abs(out["cost"] - 0.5 * (r * r).sum()) <= 1e-6 * out["cost"]
    sel = np.random.default_rng(1).choice(len(pb["obs"]), 2000, replace=False)
    rr, Jc, Jp = restated.ba_residual_jacobian(pb["cams0"], pb["pts0"], pb["cam_idx"][sel], pb["pt_idx"][sel], pb["obs"][sel], pb["K"])
    assert np.abs(out["r"][sel] - rr).max() <= 1e-4 * np.abs(rr).max()
    assert np.abs(out["Jc"][sel] - Jc).max() <= 1e-5 * np.abs(Jc).max()
    assert np.abs(out["Jp"][sel] - Jp).max() <= 1e-5 * np.abs(Jp).max()
    hist = prob.solve(max_iters=4)
    for h in hist:
        if h["accepted"]:
            assert h["cost_after"] <= h["cost_before"]
    assert hist[-1]["cost_after"] < hist[0]["cost_before"]
    prob.close()
