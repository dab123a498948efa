"""Host-side logic that needs no GPU: chunking of the pipelined drivers, argument checks of the cv2-compatible
wrappers that run before any device work."""
import numpy as np
import pytest

import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import pipeline, synth


def test_chunk_bounds():
    assert pipeline._chunk_bounds(200, (12, 60)) == [(0, 12), (12, 60), (60, 200)]
    assert pipeline._chunk_bounds(200, [8, 25] + list(range(50, 200, 25))) == \
        [(0, 8), (8, 25), (25, 50), (50, 75), (75, 100), (100, 125), (125, 150), (150, 175), (175, 200)]
    assert pipeline._chunk_bounds(5, (12, 60)) == [(0, 5)]
    assert pipeline._chunk_bounds(12, (12, 60)) == [(0, 12)]
    assert pipeline._chunk_bounds(30, (25, 8, 8, 0)) == [(0, 8), (8, 25), (25, 30)]
    for V in (3, 13, 61, 200):
        b = pipeline._chunk_bounds(V, (12, 60))
        assert b[0][0] == 0 and b[-1][1] == V and all(x[1] == y[0] for x, y in zip(b, b[1:]))


def test_find_essential_mat_argument_checks_need_no_device():
    """cv2's behaviour at the boundary (sfm.py:307): fewer than five correspondences -> (None, None); mismatched
    point arrays, a non-3x3 camera matrix, an unsupported method or a confidence outside (0, 1) raise."""
    K = synth.K_GUSTAV
    p0, p1, _, _ = synth.two_view_pair(4, seed=0)
    assert sfm.findEssentialMat(p0, p1, K, method=8, prob=0.999, threshold=0.4) == (None, None)
    q0, q1, _, _ = synth.two_view_pair(20, seed=0)
    with pytest.raises(sfm.error):
        sfm.findEssentialMat(q0, q1[:10], K)
    with pytest.raises(sfm.error):
        sfm.findEssentialMat(q0, q1, K[:2])
    with pytest.raises(sfm.error):
        sfm.findEssentialMat(q0, q1, K, method=4)
    with pytest.raises(sfm.error):
        sfm.findEssentialMat(q0, q1, K, prob=1.0)
    with pytest.raises(sfm.error):
        sfm.findEssentialMat(q0, q1, focal=1.0)          # the focal / pp overload is not implemented


def test_two_view_pair_is_deterministic_and_has_the_stated_outlier_share():
    a0, a1, R, t = synth.two_view_pair(500, seed=3, outliers=0.2)
    b0, b1, _, _ = synth.two_view_pair(500, seed=3, outliers=0.2)
    assert np.array_equal(a0, b0) and np.array_equal(a1, b1) and a0.dtype == np.float32
    assert abs(np.linalg.det(R) - 1) < 1e-12 and np.abs(R @ R.T - np.eye(3)).max() < 1e-12


def test_five_point_kernel_code_on_the_host_equals_the_oracle():
    """sfm_five_point is the hypothesis kernel's solver (null space, cubic constraints, Gauss-Jordan, det B(z), real
    roots by derivative bracketing) compiled for the host: 300 minimal samples against oracle/restated.py five_point
    (numpy SVD null space + companion-matrix roots).  Same number of models, the same matrices in the same order."""
    from oracle import restated
    K = synth.K_GUSTAV
    n_models, off_count, worst, loose = 0, 0, 0.0, 0
    for seed in range(300):
        p0, p1, _, _ = synth.two_view_pair(5, seed=1000 + seed, noise=0.3, outliers=0.0)
        q0 = (p0.astype(np.float64) - K[:2, 2]) / [K[0, 0], K[1, 1]]
        q1 = (p1.astype(np.float64) - K[:2, 2]) / [K[0, 0], K[1, 1]]
        Eh = sfm.five_point(q0, q1)
        Eo = restated.five_point(q0, q1)
        x0, x1 = np.column_stack([q0, np.ones(5)]), np.column_stack([q1, np.ones(5)])
        for e in Eh:                       # every model solves the minimal problem
            assert np.abs(np.einsum("ni,ij,nj->n", x1, e, x0)).max() < 1e-9
            assert np.abs(2 * e @ e.T @ e - np.trace(e @ e.T) * e).max() < 1e-7
            assert abs(np.linalg.norm(e) - 1) < 1e-12
        if len(Eh) != len(Eo):             # a near-double root found by one root finder only
            off_count += 1
            continue
        for a, b in zip(Eh, Eo):
            d = min(np.abs(a - b).max(), np.abs(a + b).max())
            worst = max(worst, d)
            loose += d > 1e-8
            n_models += 1
    assert off_count <= 3 and n_models > 1000
    assert worst < 1e-5 and loose <= 0.02 * n_models


def test_five_point_kernel_code_on_the_host_equals_cv2_minimal_solver():
    """Against OpenCV's own minimal solver: with exactly five correspondences cv2.findEssentialMat returns every model
    of the sample.  Same number of models on every sample; matrices equal up to sign (a handful of ill-conditioned
    roots differ beyond 1e-6 between ANY two implementations — Durand-Kerner vs bracketing vs companion matrix)."""
    import cv2
    K = synth.K_GUSTAV
    same, models, matched = 0, 0, 0
    N = 500
    for seed in range(N):
        p0, p1, _, _ = synth.two_view_pair(5, seed=5000 + seed, noise=0.3, outliers=0.0)
        Ec, _ = cv2.findEssentialMat(p0, p1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4)
        Ec = np.zeros((0, 3, 3)) if Ec is None else Ec.reshape(-1, 3, 3)
        q0 = (p0.astype(np.float64) - K[:2, 2]) / [K[0, 0], K[1, 1]]
        q1 = (p1.astype(np.float64) - K[:2, 2]) / [K[0, 0], K[1, 1]]
        Eh = sfm.five_point(q0, q1)
        same += len(Eh) == len(Ec)
        for e in Eh:
            models += 1
            matched += bool(len(Ec)) and min(min(np.abs(e - c).max(), np.abs(e + c).max()) for c in Ec) < 1e-6
    assert same >= N - 2
    assert matched >= 0.99 * models


def test_five_point_degenerate_samples_terminate_with_finite_output():
    """Identical views, duplicated correspondences, collinear points, extreme scales, all-zero input: the solver returns
    (possibly zero) finite models and never hangs — the RANSAC loop feeds it whatever the index stream draws."""
    r = np.random.default_rng(0)
    for k in range(1200):
        kind = k % 6
        q0, q1 = r.uniform(-1, 1, (5, 2)), r.uniform(-1, 1, (5, 2))
        if kind == 0:
            q1 = q0.copy()
        elif kind == 1:
            q0[1], q1[1] = q0[0], q1[0]
        elif kind == 2:
            q0[:, 1], q1[:, 1] = 0.3 * q0[:, 0], 0.3 * q1[:, 0]
        elif kind == 3:
            q0, q1 = q0 * 1e-9, q1 * 1e-9
        elif kind == 4:
            q0, q1 = q0 * 1e6, q1 * 1e6
        else:
            q0[:], q1[:] = 0, 0
        E = sfm.five_point(q0, q1)
        assert E.shape[0] <= 10 and np.all(np.isfinite(E))


def test_epnp_host_solver_is_bit_identical_to_cv2():
    """sfm_epnp runs the minimal solver's code (csrc/epnp.h, hostmath.h — what pnp_epnp.cu spreads over a CTA)
    on the host: the same rvec / tvec as cv2.solvePnP(flags=SOLVEPNP_EPNP) bit for bit, 5 points and more."""
    import cv2
    D0 = np.zeros((5, 1), np.float32)
    K = synth.K_GUSTAV
    for n in (5, 6, 11, 40):
        rng = np.random.default_rng(n)
        for _ in range(100):
            R, t = synth.orbit_pose(rng.uniform(0, 0.5))
            X = np.c_[rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n), rng.uniform(5, 11, n)].astype(np.float32)
            uv, _ = synth.project(K, R, t, X.astype(np.float64))
            p = (uv + rng.normal(0, 1.0, uv.shape)).astype(np.float32)
            ok, rvec, tvec = cv2.solvePnP(X, p, K, D0, flags=cv2.SOLVEPNP_EPNP)
            Re, te = sfm.epnp(X, p, K)
            assert np.array_equal(te, tvec.ravel()) and np.array_equal(cv2.Rodrigues(Re)[0], rvec)
