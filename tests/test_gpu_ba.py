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


def test_bundleadjustment_wrappers(engine, golden):
    g = golden("ba_small")
    X, p, Rt = sfm.BundleAdjustment(g["X0"], g["obs"], g["Rt"], g["K"], 0.5, ctx=engine)
    n = g["X0"].shape[0]
    assert X.shape == (n, 3) and p.shape == (n, 2) and Rt.shape == (3, 4)
    e_before = cvpath.ReprojectionError(g["X0"].reshape(n, 3), g["obs"].T, g["Rt"], g["K"], 0)[0]
    e_after = cvpath.ReprojectionError(np.float32(X), g["obs"].T, Rt, g["K"], 0)[0]
    assert e_after < e_before
    X2, R2, t2 = sfm.ba.bundle_adjustment(g["X0"].reshape(n, 3), g["obs"].T, None, g["K"], g["Rt"][:, :3], g["Rt"][:, 3:])
    assert np.allclose(X2, X, atol=1e-6) and R2.shape == (3, 3) and t2.shape == (3, 1)


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
