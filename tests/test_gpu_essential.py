"""GPU parity: two-view initialisation, cv2.findEssentialMat (sfm.py:307, isfm.py:80, test.py:247)
against in-process cv2 and the numpy restatement (oracle/restated.py find_essential_mat / five_point)."""
import os

import cv2
import numpy as np
import pytest

import sfm_mvs_b200 as sfm
from oracle import restated
from sfm_mvs_b200 import synth

pytestmark = pytest.mark.gpu
K = synth.K_GUSTAV


def _e_close(a, b):
    return min(np.abs(a - b).max(), np.abs(a + b).max())


@pytest.mark.parametrize("n,seed,dtype,outl", [(60, 0, np.float32, 0.2), (400, 1, np.float32, 0.2),
                                               (1500, 2, np.float32, 0.3), (3000, 3, np.float32, 0.4),
                                               (3000, 4, np.float64, 0.1), (700, 5, np.float64, 0.5),
                                               (6, 6, np.float32, 0.0), (12, 7, np.float32, 0.0)])
def test_find_essential_mat_equals_cv2(engine, n, seed, dtype, outl):
    """The reference's call (method=RANSAC, prob=0.999, threshold=0.4): the inlier mask must equal cv2's bit for
    bit (index work), E the same matrix up to sign (float64, <= 1e-7 on the unit-norm matrix)."""
    p0, p1, _, _ = synth.two_view_pair(n, seed=seed, dtype=dtype, outliers=outl)
    Ec, mc = cv2.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None)
    Eo, mo = sfm.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None, ctx=engine)
    assert Eo.shape == (3, 3) and Eo.dtype == np.float64 and mo.shape == (n, 1) and mo.dtype == np.uint8
    assert set(np.unique(mo)) <= {0, 1}                      # sfm.py:308 selects with mask.ravel() == 1
    assert np.array_equal(mo, mc)
    assert _e_close(Eo, Ec[:3]) < 1e-7
    assert abs(np.linalg.norm(Eo) - 1.0) < 1e-12


@pytest.mark.parametrize("n,seed", [(200, 10), (1000, 11)])
def test_find_essential_mat_equals_oracle(engine, n, seed):
    """Against the restated loop: same winner (iteration and model), same iteration count, same mask."""
    p0, p1, _, _ = synth.two_view_pair(n, seed=seed)
    Er, mr, info = restated.find_essential_mat(p0, p1, K, 0.999, 0.4)
    Eo, mo = sfm.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, ctx=engine)
    got = sfm.findEssentialMat.last_info
    assert np.array_equal(mo.ravel() != 0, mr)
    assert (got["iters"], got["best_iter"], got["best_model"], got["inliers"]) == \
           (info["iters"], info["best_iter"], info["best_model"], info["best_count"])
    assert _e_close(Eo, Er) < 1e-8


@pytest.mark.parametrize("seed", range(16))
def test_five_point_models_equal_cv2(engine, seed):
    """Exactly five correspondences: cv2 returns every model of the minimal sample stacked as (3k,3).  Same
    number of models, the same matrices up to sign and order, each satisfying the epipolar and cubic constraints."""
    p0, p1, _, _ = synth.two_view_pair(5, seed=100 + seed, noise=0.0, outliers=0.0)
    Ec, mc = cv2.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None)
    Eo, mo = sfm.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, ctx=engine)
    assert Eo.shape == Ec.shape and np.array_equal(mo, mc)
    Ec, Eo = Ec.reshape(-1, 3, 3), Eo.reshape(-1, 3, 3)
    q0 = np.column_stack([(p0.astype(np.float64) - K[:2, 2]) / [K[0, 0], K[1, 1]], np.ones(5)])
    q1 = np.column_stack([(p1.astype(np.float64) - K[:2, 2]) / [K[0, 0], K[1, 1]], np.ones(5)])
    keys = [e[0, 0] ** 2 for e in Eo]
    assert keys == sorted(keys)
    for e in Eo:
        assert min(_e_close(e, c) for c in Ec) < 1e-6
        assert np.abs(np.einsum("ni,ij,nj->n", q1, e, q0)).max() < 1e-9
        assert np.abs(2 * e @ e.T @ e - np.trace(e @ e.T) * e).max() < 1e-9


def test_find_essential_mat_edge_cases(engine):
    """n < 5 -> (None, None) like cv2; coincident points (rank-deficient samples) -> no model, no crash;
    unsupported overloads raise."""
    p0, p1, _, _ = synth.two_view_pair(4, seed=0)
    assert sfm.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, ctx=engine) == (None, None)
    assert cv2.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4) == (None, None)
    z = np.full((40, 2), 100.0, np.float32)
    E, m = sfm.findEssentialMat(z, z, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, ctx=engine)
    assert E is None or np.all(np.isfinite(E))
    with pytest.raises(sfm.error):
        sfm.findEssentialMat(p0, p1, K, method=cv2.LMEDS, ctx=engine)
    with pytest.raises(sfm.error):
        sfm.findEssentialMat(p0[:3], p1, K, ctx=engine)
    with pytest.raises(sfm.error):
        sfm.findEssentialMat(p0, p1, K, prob=1.5, ctx=engine)


def test_reference_bootstrap_sequence_patched(engine):
    """sfm.py:307-313 verbatim through patch_cv2: findEssentialMat -> mask select -> recoverPose -> mask select,
    against the same lines on stock cv2."""
    p0, p1, _, _ = synth.two_view_pair(2000, seed=21)

    def bootstrap(cv):
        E, mask = cv.findEssentialMat(p0, p1, K, method=cv.RANSAC, prob=0.999, threshold=0.4, mask=None)
        a, b = p0[mask.ravel() == 1], p1[mask.ravel() == 1]
        _, R, t, mask2 = cv.recoverPose(E, a, b, K)
        return R, t, a[mask2.ravel() > 0], b[mask2.ravel() > 0]
    Rc, tc, ac, bc = bootstrap(cv2)
    saved = sfm.patch_cv2()
    try:
        Ro, to, ao, bo = bootstrap(cv2)
    finally:
        sfm.unpatch_cv2(saved)
    assert np.array_equal(ao, ac) and np.array_equal(bo, bc)
    assert np.abs(Ro - Rc).max() < 1e-7 and np.abs(to - tc).max() < 1e-7


def test_two_view_init_equals_reference_lines(engine):
    """pipeline.two_view_init = sfm.py:307-316 (E, masks, recoverPose, pose composition) against the same lines run
    on stock cv2; the pose then starts a registration chain exactly like a ground-truth one."""
    from sfm_mvs_b200 import pipeline
    p0, p1, R_gt, t_gt = synth.two_view_pair(3000, seed=33, outliers=0.3)
    out = pipeline.two_view_init(p0, p1, K, ctx=engine)
    E, mask = cv2.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None)
    a, b = p0[mask.ravel() == 1], p1[mask.ravel() == 1]
    _, R, t, mask = cv2.recoverPose(E, a, b, K)
    a, b = a[mask.ravel() > 0], b[mask.ravel() > 0]
    assert out["n_essential"] == int((cv2.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4)[1] == 1).sum())
    assert np.array_equal(out["pts0"], a) and np.array_equal(out["pts1"], b)
    assert np.abs(out["Rt1"][:, :3] - R).max() < 1e-7 and np.abs(out["Rt1"][:, 3] - t.ravel()).max() < 1e-7
    # and it is the scene's motion: rotation within 0.5 degree, translation direction within 2 degrees
    ang = np.degrees(np.arccos(np.clip((np.trace(out["Rt1"][:, :3] @ R_gt.T) - 1) / 2, -1, 1)))
    tdir = np.degrees(np.arccos(np.clip(out["Rt1"][:, 3] @ t_gt / np.linalg.norm(t_gt), -1, 1)))
    assert ang < 0.5 and tdir < 2.0


def test_pairwise_init_equals_the_isfm_loop(engine):
    """isfm.py:68-87 — all earlier views against each new view: match, findEssentialMat, recoverPose, survivor count
    (the number isfm.py prints).  Against the same lines on stock cv2 (oracle.cvpath matching + cv2 calls)."""
    from oracle import cvpath
    from sfm_mvs_b200 import pipeline
    scene = synth.orbit_scene(4, 1500, seed=11)
    Ks = scene["K"]
    views = [pipeline.DeviceView(engine, v["kp"], v["des"]) for v in scene["views"]]
    got = pipeline.pairwise_init(engine, views, Ks)
    assert [g["pair"] for g in got] == [(0, 1), (0, 2), (1, 2), (0, 3), (1, 3), (2, 3)]
    for g in got:
        j, i = g["pair"]
        vj, vi = scene["views"][j], scene["views"][i]
        p0, p1 = cvpath.match_keypoints(vj["kp"], vj["des"], vi["kp"], vi["des"])
        assert g["n_match"] == len(p0)
        E, mask = cv2.findEssentialMat(p0, p1, Ks, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None)
        a, b = p0[mask.ravel() == 1], p1[mask.ravel() == 1]
        _, R, t, mask = cv2.recoverPose(E, a, b, Ks)
        a, b = a[mask.ravel() > 0], b[mask.ravel() > 0]
        assert g["n_essential"] == int((cv2.findEssentialMat(p0, p1, Ks, method=cv2.RANSAC, prob=0.999, threshold=0.4)[1] == 1).sum())
        assert g["n_pose"] == len(a) and np.array_equal(g["pts0"], a) and np.array_equal(g["pts1"], b)
        assert np.abs(g["R"] - R).max() < 1e-6 and np.abs(g["t"] - t).max() < 1e-6


@pytest.mark.parametrize("n,seed,prob,thr,max_iters,dtype",
                         [(500, 0, 0.99, 1.0, 50, np.float32), (800, 1, 0.9, 3.0, 1000, np.float32),
                          (300, 2, 0.999, 0.4, 5, np.float64), (1000, 3, 0.5, 0.2, 1000, np.float32),
                          (400, 4, 0.999, 0.05, 200, np.float64)])
def test_find_essential_mat_parameter_variants(engine, n, seed, prob, thr, max_iters, dtype):
    """Other confidences, thresholds, iteration caps and float64 points than the reference's call: same mask as cv2,
    same iteration count and winner as the restated loop."""
    p0, p1, _, _ = synth.two_view_pair(n, seed=seed, dtype=dtype)
    Ec, mc = cv2.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=prob, threshold=thr, maxIters=max_iters)
    Eo, mo = sfm.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=prob, threshold=thr, maxIters=max_iters, ctx=engine)
    got = dict(sfm.findEssentialMat.last_info)
    assert np.array_equal(mo, mc) and _e_close(Eo, Ec[:3]) < 1e-7
    _, _, info = restated.find_essential_mat(p0, p1, K, prob, thr, max_iters)
    assert (got["iters"], got["best_iter"], got["inliers"]) == (info["iters"], info["best_iter"], info["best_count"])


def test_find_essential_mat_batched(engine):
    """All pairs' essential matrices in a few launches: the same records as one call per pair."""
    from sfm_mvs_b200 import pipeline
    scene = synth.orbit_scene(4, 1500, seed=11)
    views = [pipeline.DeviceView(engine, v["kp"], v["des"]) for v in scene["views"]]
    one = pipeline.pairwise_init(engine, views, scene["K"])
    many = pipeline.pairwise_init(engine, views, scene["K"], batched=True)
    for a, b in zip(one, many):
        assert (a["pair"], a["n_match"], a["n_essential"], a["n_pose"]) == (b["pair"], b["n_match"], b["n_essential"], b["n_pose"])
        assert np.array_equal(a["pts0"], b["pts0"]) and np.abs(a["R"] - b["R"]).max() < 1e-9
