"""GPU parity: hot path 3a (PnP-RANSAC scoring, stopping rule, inliers, refinement) against cv2."""
import cv2
import numpy as np
import pytest

import sfm_mvs_b200 as sfm
from oracle import restated
from sfm_mvs_b200 import synth

pytestmark = pytest.mark.gpu
K = synth.K_GUSTAV
D0 = np.zeros((5, 1), np.float32)


def _problem(seed, n=None):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(30, 3000)) if n is None else n
    R, t = synth.orbit_pose(rng.uniform(0, 0.5))
    X = np.c_[rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n), rng.uniform(5, 11, n)].astype(np.float32)
    uv, _ = synth.project(K, R, t, X.astype(np.float64))
    p = (uv + rng.normal(0, rng.uniform(0.1, 2.0), uv.shape)).astype(np.float32)
    bad = rng.random(n) < rng.uniform(0, 0.6)
    p[bad] += rng.uniform(-80, 80, (int(bad.sum()), 2)).astype(np.float32)
    return X, p


def _cv_hypotheses(X, p, iters=100):
    """OpenCV's own minimal solutions for the RNG subset stream (what cv2.solvePnPRansac computes inside)."""
    subs = sfm.ransac_subsets(len(X), iters)
    hyp = np.zeros((iters, 6))
    valid = np.zeros(iters, np.uint8)
    for i, s in enumerate(subs):
        ok, r, t = cv2.solvePnP(X[s], p[s], K, D0, flags=cv2.SOLVEPNP_EPNP)
        if ok:
            hyp[i, :3], hyp[i, 3:] = r.ravel(), t.ravel()
            valid[i] = 1
    return hyp, valid


@pytest.mark.parametrize("seed", range(6))
def test_hypothesis_scoring_bit_exact(engine, seed):
    X, p = _problem(seed)
    hyp, valid = _cv_hypotheses(X, p, 40)
    Rt = np.array([np.hstack([cv2.Rodrigues(h[:3])[0], h[3:].reshape(3, 1)]) for h in hyp])
    counts, masks = engine.pnp_score(X, p, K, Rt, 8.0)
    for h in range(len(hyp)):
        ref = restated.score_hypothesis(X, p, cv2.Rodrigues(hyp[h, :3])[0], hyp[h, 3:], K, 8.0)
        assert np.array_equal(masks[h].astype(bool), ref), h
        assert counts[h] == int(ref.sum())


@pytest.mark.parametrize("seed", range(24))
def test_ransac_inlier_mask_bit_exact_given_opencv_minimal_solutions(engine, seed):
    X, p = _problem(seed)
    ok_ref, rvec_ref, tvec_ref, inl_ref = cv2.solvePnPRansac(X, p, K, D0, cv2.SOLVEPNP_ITERATIVE)
    hyp, valid = _cv_hypotheses(X, p)
    ok, rvec, tvec, inl, info = engine.pnp_ransac(X, p, K, hypotheses=hyp, hyp_valid=valid)
    assert ok == ok_ref
    if ok:
        assert np.array_equal(inl, inl_ref[:, 0])                   # bit-exact mask
        assert np.abs(rvec - rvec_ref.ravel()).max() <= 1e-4 * max(1.0, np.abs(rvec_ref).max())
        assert np.abs(tvec - tvec_ref.ravel()).max() <= 1e-4 * max(1.0, np.abs(tvec_ref).max())


def test_golden_pnp_inliers(engine, golden):
    g = golden("geometry")
    X, p = g["pnp_X"], g["pnp_p"]
    hyp, valid = _cv_hypotheses(X, p)
    ok, rvec, tvec, inl, _ = engine.pnp_ransac(X, p, g["K"], hypotheses=hyp, hyp_valid=valid)
    ok_now, _, _, inl_now = cv2.solvePnPRansac(X, p, g["K"], D0)
    assert ok and np.array_equal(inl, inl_now[:, 0])
    # the committed fixture was produced by the reference's PnP() in the build container.  Everything up to the inlier
    # list is OpenCV's own fixed-order arithmetic (no LAPACK below 25 rows), so it holds on every host; the default
    # path (engine minimal solver) must reproduce it too
    assert np.array_equal(inl, g["pnp_inliers"][:, 0])
    ok2, rvec2, tvec2, inl2, _ = engine.pnp_ransac(X, p, g["K"])
    assert ok2 and np.array_equal(inl2, g["pnp_inliers"][:, 0])
    for rv, tv in ((rvec, tvec), (rvec2, tvec2)):
        R = sfm.rodrigues_to_matrix(rv)
        assert np.abs(R - g["pnp_R"]).max() < 1e-6 and np.abs(tv - g["pnp_t"].ravel()).max() < 1e-5


@pytest.mark.parametrize("seed", range(8))
def test_device_minimal_solver_is_bit_identical_to_cv2(engine, seed):
    """pnp_epnp_kernel (csrc/pnp_epnp.cu) on the subsets of OpenCV's RANSAC stream against cv2.solvePnP(EPNP) on
    the same five points: the same rotation and translation BIT FOR BIT (the raw solver output; cv2 hands it on
    as cv2.Rodrigues(R), compared through that very call)."""
    X, p = _problem(seed)
    subs = sfm.ransac_subsets(len(X), 25)
    R, t = engine.epnp_batch(X, p, K, subs)
    for s, Re, te in zip(subs, R, t):
        ok, rvec, tvec = cv2.solvePnP(X[s], p[s], K, D0, flags=cv2.SOLVEPNP_EPNP)
        assert ok and np.array_equal(te, tvec.ravel()), s
        assert np.array_equal(cv2.Rodrigues(Re)[0], rvec), s


@pytest.mark.parametrize("seed", range(24))
def test_ransac_default_path_inlier_mask_bit_exact(engine, seed):
    """The product's default call — minimal solver, scoring, stopping rule and refinement all on the GPU — against
    cv2.solvePnPRansac on the same inputs (0-60 % outliers, 0.1-2 px noise, 30-3000 points): identical inlier
    list, pose within 1e-4."""
    X, p = _problem(seed)
    ok_ref, rvec_ref, tvec_ref, inl_ref = cv2.solvePnPRansac(X, p, K, D0, cv2.SOLVEPNP_ITERATIVE)
    ok, rvec, tvec, inl, info = engine.pnp_ransac(X, p, K)
    assert ok == ok_ref
    if ok:
        assert np.array_equal(inl, inl_ref[:, 0])
        assert np.abs(rvec - rvec_ref.ravel()).max() <= 1e-4 * max(1.0, np.abs(rvec_ref).max())
        assert np.abs(tvec - tvec_ref.ravel()).max() <= 1e-4 * max(1.0, np.abs(tvec_ref).max())
        assert info["hyp_solved"] == 100


def _solvable_problem(seed):
    """Outlier ratio <= 0.4 and noise <= 1 px: the 100 subsets then contain all-inlier ones, so RANSAC's
    outcome does not hinge on which contaminated subset happens to score best (with ~60 % outliers none of
    the 100 five-point subsets is clean and any two minimal solvers legitimately pick different models)."""
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(60, 3000))
    R, t = synth.orbit_pose(rng.uniform(0, 0.5))
    X = np.c_[rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n), rng.uniform(5, 11, n)].astype(np.float32)
    uv, _ = synth.project(K, R, t, X.astype(np.float64))
    p = (uv + rng.normal(0, rng.uniform(0.1, 1.0), uv.shape)).astype(np.float32)
    bad = rng.random(n) < rng.uniform(0, 0.4)
    p[bad] += rng.uniform(-80, 80, (int(bad.sum()), 2)).astype(np.float32)
    return X, p


@pytest.mark.parametrize("seed", range(16))
def test_ransac_with_engine_minimal_solver(engine, seed):
    """Full GPU path on the easier problems (<= 40 % outliers): identical inlier list, and the refined pose within
    1e-3 px of OpenCV's on every consensus point."""
    X, p = _solvable_problem(seed)
    ok_ref, rvec_ref, tvec_ref, inl_ref = cv2.solvePnPRansac(X, p, K, D0)
    ok, rvec, tvec, inl, info = engine.pnp_ransac(X, p, K)
    assert ok == ok_ref
    if ok:
        assert np.array_equal(inl, inl_ref[:, 0])
        proj_ref, _ = cv2.projectPoints(X[inl_ref[:, 0]], rvec_ref, tvec_ref, K, None)
        proj, _ = cv2.projectPoints(X[inl_ref[:, 0]], rvec, tvec, K, None)
        assert np.abs(proj - proj_ref).max() < 1e-3
        assert np.all(np.diff(inl) > 0) and info["hyp_solved"] == 100


def test_solvepnpransac_surface(engine):
    X, p = _problem(3, n=600)
    ok, rvec, tvec, inl = sfm.solvePnPRansac(X, p, K, D0, cv2.SOLVEPNP_ITERATIVE, ctx=engine)   # the reference's call
    assert ok and rvec.shape == (3, 1) and tvec.shape == (3, 1) and inl.dtype == np.int32 and inl.shape[1] == 1
    R, t, p_in, X_in, p0_in = sfm.PnP(X, p, K, D0, p.copy(), 0, ctx=engine)
    assert R.shape == (3, 3) and len(p_in) == len(X_in) == len(inl)
    # failure is a return value, not an exception (pure noise -> no consensus of more than 4)
    rng = np.random.default_rng(0)
    Xn = rng.uniform(-1, 1, (40, 3)).astype(np.float32); Xn[:, 2] += 8
    pn = rng.uniform(0, 900, (40, 2)).astype(np.float32)
    okn, _, _, inln = sfm.solvePnPRansac(Xn, pn, K, D0, ctx=engine)
    okc, _, _, inlc = cv2.solvePnPRansac(Xn, pn, K, D0)
    assert okn == okc and ((inln is None) == (inlc is None))
    with pytest.raises(sfm.error):
        sfm.solvePnPRansac(X[:3], p[:3], K, D0, ctx=engine)
    # five points: OpenCV returns the EPnP pose with all five as inliers
    ok5, r5, t5, inl5 = sfm.solvePnPRansac(X[:5], p[:5], K, D0, ctx=engine)
    assert ok5 and inl5[:, 0].tolist() == [0, 1, 2, 3, 4]


@pytest.mark.parametrize("seed", range(12))
def test_four_points_is_p3p_with_the_fourth_as_referee(engine, seed):
    """npoints == 4: cv2.solvePnPRansac runs no RANSAC — model_points == npoints — but one solvePnP(SOLVEPNP_P3P): the
    three-point problem on rows 0-2, row 3 choosing among its solutions, all four reported as inliers, no refinement.
    The engine solves the same minimal problem on the device (Grunert's quartic, csrc/pnp.cu pnp_p3p4_kernel): same
    return convention, the same solution chosen, pose within the north_star 1e-4 (cv2's own P3P leaves ~1e-5 px on its
    three points, so tighter than ~1e-5 is not there to be had)."""
    rng = np.random.default_rng(100 + seed)
    X = (rng.random((4, 3)) * 2).astype(np.float32)
    X[:, 2] += 4
    R, _ = cv2.Rodrigues(rng.normal(size=3) * 0.3)
    t = rng.normal(size=3) * 0.3
    p = (K @ (R @ X.T + t[:, None])).T
    p = (p[:, :2] / p[:, 2:] + rng.normal(size=(4, 2)) * (seed % 3)).astype(np.float32)     # exact, 1 px and 2 px noise
    okc, rc, tc, inlc = cv2.solvePnPRansac(X, p, K, D0, cv2.SOLVEPNP_ITERATIVE)
    ok, rv, tv, inl = sfm.solvePnPRansac(X, p, K, D0, cv2.SOLVEPNP_ITERATIVE, ctx=engine)
    assert ok == okc
    if okc:
        assert np.array_equal(inl, inlc) and inl.dtype == inlc.dtype and inl[:, 0].tolist() == [0, 1, 2, 3]
        assert np.abs(cv2.Rodrigues(rv)[0] - cv2.Rodrigues(rc)[0]).max() < 1e-4 and np.abs(tv - tc).max() < 1e-4
        pr, _ = cv2.projectPoints(X[:3], rv, tv, K, D0)
        assert np.abs(pr[:, 0] - p[:3]).max() < 1e-3          # a P3P solution: exact on its three points
    else:
        assert inl is None


# ----------------------------------------------------------------------------- the per-view loop
def _cv_hyp_fn(X, p):
    return _cv_hypotheses(np.ascontiguousarray(X), np.ascontiguousarray(p))


def test_registration_chain_equals_reference_loop(engine, golden):
    """sfm.py:341-409 over a 7-view scene: the engine's device-resident loop against the oracle port of
    the reference loop (and the golden fixture produced by the reference's own defs when this host's
    OpenCV reproduces it).  PnP runs on OpenCV's minimal solutions, so poses must agree to rounding."""
    from oracle import cvpath
    from sfm_mvs_b200 import pipeline
    g = golden("chain")
    scene = synth.orbit_scene(int(g["n_views"]), int(g["n_pts"]), seed=int(g["seed"]))
    ref = cvpath.register_chain(scene)
    outs = pipeline.register_chain(scene, ctx=engine, hypothesis_fn=_cv_hyp_fn)
    assert len(outs) == len(ref) == 5
    for o, r in zip(outs, ref):
        assert (o["n_match"], o["n_pnp"], o["n_inl"], o["n_new"]) == (r["n_match"], r["n_pnp"], r["n_inl"], len(r["X_new"]))
        assert np.abs(o["Rt"] - r["Rt"]).max() < 1e-5
        assert abs(o["err_pnp"] - r["err_pnp"]) <= 1e-4 * r["err_pnp"]
        assert abs(o["err_new"] - r["err_new"]) <= 1e-4 * r["err_new"]
        rel = np.abs(o["X_new"] - r["X_new"]).max() / np.abs(r["X_new"]).max()
        assert rel < 1e-4, rel
    if np.array_equal(np.array([r["Rt"] for r in ref]), g["Rt"]):
        assert np.abs(np.array([o["Rt"] for o in outs]) - g["Rt"]).max() < 1e-5
        assert np.abs(np.vstack([o["X_new"] for o in outs]) - g["X_new"]).max() / np.abs(g["X_new"]).max() < 1e-4


def test_registration_chain_gpu_minimal_solver(engine):
    """Throughput configuration (engine EPnP): same views registered, poses close to ground truth."""
    from sfm_mvs_b200 import pipeline
    scene = synth.orbit_scene(8, 1500, seed=4)
    outs = pipeline.register_chain(scene, ctx=engine)
    assert len(outs) == 6
    for i, o in enumerate(outs):
        v = scene["views"][i + 2]
        Rt_gt = np.hstack([v["R"], v["t"].reshape(3, 1)])
        assert np.abs(o["Rt"][:, :3] - Rt_gt[:, :3]).max() < 5e-3
        assert np.abs(o["Rt"][:, 3] - Rt_gt[:, 3]).max() < 5e-2
        assert o["err_new"] < 0.05 and o["n_inl"] > 0.8 * o["n_pnp"]


def test_native_and_streamed_chain_equal_the_python_loop(engine, monkeypatch):
    """sfm_chain_run (the loop in the library) and register_host (chunked upload + sfm_chain_extend) run the same
    kernels in the same order as the Python loop.  With SFM_CHAIN_SYNC=1 (host reads counts and pose per view, as
    the Python loop does) the results are identical bit for bit; the default synchronisation-free loop forms the
    pose matrices on the device (device libm / FMA contraction) and sums the LM normal equations over a cluster of
    CTAs (different summation order), so poses agree to ~1e-8 and points to ~1e-6 — four orders inside the 1e-4 bar."""
    import torch
    from sfm_mvs_b200 import pipeline
    scene = synth.orbit_scene(9, 1200, seed=7)
    K = scene["K"]
    Rt0 = np.hstack([scene["views"][0]["R"], scene["views"][0]["t"]])
    Rt1 = np.hstack([scene["views"][1]["R"], scene["views"][1]["t"]])
    views = [pipeline.DeviceView(engine, v["kp"], v["des"]) for v in scene["views"]]
    chain = pipeline.RegistrationChain(engine, K)
    matches = chain.match_pairs(views, [(i, i + 1) for i in range(len(views) - 1)])
    # Python loop
    chain.bootstrap(views, Rt0, Rt1, matches[0])
    py = [chain.register(matches[i] if i > 0 else None, matches[i + 1], first=(i == 0)) for i in range(len(views) - 2)]
    engine.sync()
    host_args = (engine, K, [v["kp"] for v in scene["views"]], [v["des"] for v in scene["views"]], Rt0, Rt1)
    monkeypatch.setenv("SFM_CHAIN_SYNC", "1")
    native = chain.run(views, Rt0, Rt1, matches=matches)
    streamed = pipeline.register_host(*host_args, chunk=4)
    monkeypatch.delenv("SFM_CHAIN_SYNC")
    native_sf = chain.run(views, Rt0, Rt1, matches=matches)
    streamed_sf = pipeline.register_host(*host_args, chunk=4)
    assert len(py) == len(native) == len(streamed) == len(native_sf) == len(streamed_sf) == 7
    for a, b, c, d, f in zip(py, native, streamed, native_sf, streamed_sf):
        e = a["errs"].cpu().numpy()
        for o in (b, c, d, f):
            assert (o["n_match"], o["n_pnp"], o["n_inl"], o["n_new"]) == (a["n_match"], a["n_pnp"], a["n_inl"], a["n_new"])
        for o in (b, c):
            assert np.array_equal(o["Rt"], a["Rt"])
            assert o["err_pnp"] == e[0] and o["err_new"] == e[1]
            assert torch.equal(o["X_new"][:o["n_new"]], a["X_new"][:a["n_new"]])
        for o in (d, f):
            # (the LM refinement stops when the relative parameter change drops below FLT_EPSILON, as OpenCV's does: two
            # summation orders of its normal equations end within ~1e-7 of each other, not closer)
            assert np.abs(o["Rt"] - a["Rt"]).max() < 1e-6
            # (projections are rounded to float32 before the differences: a 1e-9 pose change moves a few by one ulp)
            assert abs(o["err_pnp"] - e[0]) <= 1e-5 * e[0] and abs(o["err_new"] - e[1]) <= 1e-5 * e[1]
            xa, xo = a["X_new"][:a["n_new"]].cpu().numpy(), o["X_new"][:o["n_new"]].cpu().numpy()
            assert np.abs(xo - xa).max() <= 1e-5 * np.abs(xa).max()


def test_registration_at_baseline_descriptor_count(engine):
    """BASELINE configs[2] shape (5000 descriptors per view; 30 views here to stay in seconds): against the oracle's
    loop (the reference's cv2 calls) — identical match and association counts for every view, poses and new points
    agreeing far inside the drift of incremental SfM — and the host-array entry (chunked upload, batched K1b / K1,
    sfm_chain_extend) against the resident driver."""
    import torch
    from oracle import cvpath
    from sfm_mvs_b200 import pipeline
    scene = synth.orbit_scene(30, 5000, seed=11)
    K = scene["K"]
    Rt0 = np.hstack([scene["views"][0]["R"], scene["views"][0]["t"]])
    Rt1 = np.hstack([scene["views"][1]["R"], scene["views"][1]["t"]])
    outs = pipeline.register_host(engine, K, [v["kp"] for v in scene["views"]], [v["des"] for v in scene["views"]], Rt0, Rt1,
                                  chunk=7)
    ref = cvpath.register_chain(scene)
    assert len(outs) == len(ref) == 28
    for o, r in zip(outs, ref):
        assert (o["n_match"], o["n_pnp"]) == (r["n_match"], r["n_pnp"])
        assert abs(o["n_inl"] - r["n_inl"]) <= 0.01 * r["n_inl"] and o["n_new"] == len(r["X_new"])
        assert np.abs(o["Rt"] - r["Rt"]).max() < 2e-3            # different minimal solver, same LM fixed point up to its tolerance
        assert abs(o["err_new"] - r["err_new"]) < 0.05 * r["err_new"] + 1e-4
    dev = engine.torch_device
    with torch.cuda.stream(engine.torch_stream()):
        kps = [torch.from_numpy(v["kp"]).to(dev) for v in scene["views"]]
        dess = [torch.from_numpy(v["des"]).to(dev) for v in scene["views"]]
    res = pipeline.RegistrationChain(engine, K).run(pipeline.DeviceView.batch(engine, kps, dess), Rt0, Rt1)
    assert len(res) == 28
    for a, b in zip(outs, res):
        assert (a["n_match"], a["n_pnp"], a["n_inl"], a["n_new"]) == (b["n_match"], b["n_pnp"], b["n_inl"], b["n_new"])
        assert np.abs(a["Rt"] - b["Rt"]).max() < 1e-7
    # the pipelined resident driver (matching context + sfm_chain_extend_async / collect, what bench.py times):
    # the same kernels on the same data, so the same records bit for bit
    pip = pipeline.register_device(engine, K, kps, dess, Rt0, Rt1)
    assert len(pip) == 28
    for a, b in zip(pip, res):
        assert (a["n_match"], a["n_pnp"], a["n_inl"], a["n_new"]) == (b["n_match"], b["n_pnp"], b["n_inl"], b["n_new"])
        assert np.array_equal(a["Rt"], b["Rt"]) and a["err_pnp"] == b["err_pnp"] and a["err_new"] == b["err_new"]
        assert torch.equal(a["X_new"][:a["n_new"]], b["X_new"][:b["n_new"]])


def test_chain_extend_async_needs_collect(engine):
    """One asynchronous call in flight per chain: a second sfm_chain_extend_async before sfm_chain_collect is an
    error, collect without a pending call returns nothing."""
    from sfm_mvs_b200 import pipeline
    scene = synth.orbit_scene(5, 900, seed=4)
    K = scene["K"]
    Rt0 = np.hstack([scene["views"][0]["R"], scene["views"][0]["t"]])
    Rt1 = np.hstack([scene["views"][1]["R"], scene["views"][1]["t"]])
    views = [pipeline.DeviceView(engine, v["kp"], v["des"]) for v in scene["views"]]
    chain = pipeline.RegistrationChain(engine, K)
    matches = chain.match_pairs(views, [(i, i + 1) for i in range(4)])
    nc = pipeline.NativeChain(engine, K, Rt0, Rt1, max(pm.n for pm in matches))
    try:
        assert nc.collect() == []
        nc.launch(matches[:3])
        with pytest.raises(sfm.error):
            nc.launch(matches[3:])
        first = nc.collect()
        assert len(first) == 2
        nc.launch(matches[3:])
        assert len(nc.collect()) == 1
    finally:
        nc.close()


@pytest.mark.parametrize("views,desc,seed", [(50, 2000, 3), (200, 5000, 0)])
def test_every_view_of_the_baseline_scenes_per_call_parity(engine, views, desc, seed):
    """BASELINE configs[1] (50 x 2000) and configs[2] (200 x 5000, the benchmarked scene) in full: the CPU arm
    (oracle port of sfm.py:341-409, ~30 s for the large scene) registers every view and records what the reference
    hands to solvePnPRansac (sfm.py:362); the engine's default call on those same inputs must return the IDENTICAL
    inlier list for every view and the refined pose within 1e-4.  The engine's own loop over the same scene must see
    the same matches and associations, the same number of new points, and poses / errors / points within the
    north_star bars of the CPU loop (bit-identity of the two loops is not defined: cv2's LM refinement sums its
    normal equations through OpenBLAS, see bench.parity_check)."""
    from oracle import cvpath
    from sfm_mvs_b200 import pipeline
    scene = synth.orbit_scene(views, desc, seed=seed)
    Ks = scene["K"]
    ref = cvpath.register_chain(scene)
    assert len(ref) == views - 2
    for v, r in enumerate(ref):
        ok_ref, rv_ref, tv_ref, inl_ref = cv2.solvePnPRansac(r["pnp_X"], r["pnp_p"], Ks, D0, cv2.SOLVEPNP_ITERATIVE)
        ok, rv, tv, inl, _ = engine.pnp_ransac(r["pnp_X"], r["pnp_p"], Ks)
        assert ok and ok_ref and len(inl_ref) == r["n_inl"]
        assert np.array_equal(inl, inl_ref[:, 0]), f"view {v + 2}: inlier mask differs from cv2"
        assert np.abs(rv - rv_ref.ravel()).max() <= 1e-4 and np.abs(tv - tv_ref.ravel()).max() <= 1e-4
    outs = pipeline.register_chain(scene, ctx=engine)
    assert len(outs) == len(ref)
    same_inl = 0
    for o, r in zip(outs, ref):
        assert (o["n_match"], o["n_pnp"], o["n_new"]) == (r["n_match"], r["n_pnp"], len(r["X_new"]))
        same_inl += o["n_inl"] == r["n_inl"]
        assert abs(o["n_inl"] - r["n_inl"]) <= max(2, 0.01 * r["n_inl"])
        assert np.abs(o["Rt"] - r["Rt"]).max() < 1e-4
        assert abs(o["err_new"] - r["err_new"]) <= 1e-4 * r["err_new"] + 1e-7
        assert np.abs(o["X_new"] - r["X_new"]).max() <= 1e-4 * np.abs(r["X_new"]).max()
    print(f"{views} x {desc}: engine loop and CPU loop agree on the inlier count of {same_inl} / {len(ref)} views")
    assert same_inl >= 0.8 * len(ref)
