"""numpy restatements of the OpenCV algorithms behind the reference's hot-path calls.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference calls into
OpenCV (not vendored, not pinned — see oracle/__init__.py); these functions restate
the *published* algorithm of each call so that the CUDA engine can be compared
against inspectable intermediates, and are themselves pinned against in-process
cv2 4.13.0 and the golden fixtures by tests/test_oracle.py.

Reference call sites: sfm.py:259-265 (2-NN + ratio), sfm.py:53-54 (triangulate),
sfm.py:84-95 (Rodrigues/projectPoints/norm), sfm.py:67 (solvePnPRansac),
sfm.py:104-136 (BA residual), notebook cell 6 (BA block structure).
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------- 2-NN matching
def knn2_l2(q: np.ndarray, t: np.ndarray):
    """Exact brute-force 2-NN under L2 (cv2.BFMatcher(NORM_L2).knnMatch(k=2), sfm.py:259-260).

    Squared distances are accumulated in float64 (exact for integer-valued SIFT
    descriptors); ties go to the lower train index; distance = float32(sqrt(.)).
    Returns idx (nq,k) int32 and dist (nq,k) float32 with k = min(2, nt).
    """
    q64 = np.asarray(q, dtype=np.float64)
    t64 = np.asarray(t, dtype=np.float64)
    nq, nt = q64.shape[0], t64.shape[0]
    k = min(2, nt)
    idx = np.zeros((nq, k), dtype=np.int32)
    dist = np.zeros((nq, k), dtype=np.float32)
    if nq == 0 or nt == 0:
        return idx, dist
    tn = (t64 * t64).sum(1)
    B = max(1, min(nq, (1 << 24) // max(nt, 1)))
    for s in range(0, nq, B):
        qq = q64[s:s + B]
        d2 = (qq * qq).sum(1)[:, None] + tn[None, :] - 2.0 * (qq @ t64.T)
        d2 = np.maximum(d2, 0.0)
        rows = np.arange(qq.shape[0])
        i1 = np.argmin(d2, axis=1)                       # first occurrence -> lowest index on ties
        idx[s:s + B, 0] = i1
        dist[s:s + B, 0] = np.sqrt(d2[rows, i1].astype(np.float32))
        if k == 2:
            d2[rows, i1] = np.inf
            i2 = np.argmin(d2, axis=1)
            idx[s:s + B, 1] = i2
            dist[s:s + B, 1] = np.sqrt(d2[rows, i2].astype(np.float32))
    return idx, dist


def ratio_mask(dist: np.ndarray, ratio: float = 0.70) -> np.ndarray:
    """Lowe test exactly as the Python loop does it (sfm.py:264): float32 distances widened
    to double, `d1 < ratio*d2` in double, strict."""
    d = dist.astype(np.float64)
    return d[:, 0] < ratio * d[:, 1]


# --------------------------------------------------------------------------- triangulation
def triangulate_dlt(P1: np.ndarray, P2: np.ndarray, x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """cv2.triangulatePoints (sfm.py:53): per correspondence the 4x4 system with rows
    x*P[2]-P[0], y*P[2]-P[1] for both views; X = right singular vector of the smallest
    singular value (float64), unit norm, sign arbitrary.  x1,x2 are (2,N).  Returns (4,N) f64."""
    P1 = np.asarray(P1, np.float64); P2 = np.asarray(P2, np.float64)
    x1 = np.asarray(x1, np.float64); x2 = np.asarray(x2, np.float64)
    n = x1.shape[1]
    A = np.empty((n, 4, 4))
    A[:, 0] = x1[0][:, None] * P1[2] - P1[0]
    A[:, 1] = x1[1][:, None] * P1[2] - P1[1]
    A[:, 2] = x2[0][:, None] * P2[2] - P2[0]
    A[:, 3] = x2[1][:, None] * P2[2] - P2[1]
    _, _, vt = np.linalg.svd(A)
    return vt[:, 3, :].T.copy()


# --------------------------------------------------------------------------- Rodrigues / projection
def rodrigues_to_matrix(rvec) -> np.ndarray:
    """cv2.Rodrigues(vector->matrix): R = c*I + (1-c)*k k^T + s*[k]x ; identity below DBL_EPSILON."""
    r = np.asarray(rvec, np.float64).ravel()
    theta = math.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    c, s = math.cos(theta), math.sin(theta)
    c1 = 1.0 - c
    it = 1.0 / theta
    kx, ky, kz = r[0] * it, r[1] * it, r[2] * it
    return np.array([
        [c + c1 * kx * kx, c1 * kx * ky - s * kz, c1 * kx * kz + s * ky],
        [c1 * kx * ky + s * kz, c + c1 * ky * ky, c1 * ky * kz - s * kx],
        [c1 * kx * kz - s * ky, c1 * ky * kz + s * kx, c + c1 * kz * kz],
    ])


def rodrigues_to_vector(R) -> np.ndarray:
    """cv2.Rodrigues(matrix->vector): orthonormalise with SVD (R <- U V^T), then the log map
    with OpenCV's branch for theta near pi."""
    R = np.asarray(R, np.float64)
    u, _, vt = np.linalg.svd(R)
    R = u @ vt
    rx, ry, rz = R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]
    s = math.sqrt((rx * rx + ry * ry + rz * rz) * 0.25)
    c = min(max((R[0, 0] + R[1, 1] + R[2, 2] - 1.0) * 0.5, -1.0), 1.0)
    theta = math.acos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = (R[0, 0] + 1) * 0.5
        x = math.sqrt(max(t, 0.0))
        t = (R[1, 1] + 1) * 0.5
        y = math.sqrt(max(t, 0.0)) * (-1.0 if R[0, 1] < 0 else 1.0)
        t = (R[2, 2] + 1) * 0.5
        z = math.sqrt(max(t, 0.0)) * (-1.0 if R[0, 2] < 0 else 1.0)
        if abs(x) < abs(y) and abs(x) < abs(z) and (R[1, 2] > 0) != (y * z > 0):
            z = -z
        v = np.array([x, y, z])
        return v * (theta / np.linalg.norm(v))
    vth = 1.0 / (2.0 * s) * theta
    return np.array([rx, ry, rz]) * vth


def project_pinhole(X, R, t, K) -> np.ndarray:
    """cv2.projectPoints with zero distortion (sfm.py:88): Y=R X+t ; s = 1/Y.z (1 if Y.z==0) ;
    u = fx*(Y.x*s)+cx ; v = fy*(Y.y*s)+cy.  Skew K[0,1] is ignored, as in OpenCV.  float64."""
    X = np.asarray(X, np.float64).reshape(-1, 3)
    R = np.asarray(R, np.float64); t = np.asarray(t, np.float64).ravel()
    K = np.asarray(K, np.float64)
    x = R[0, 0] * X[:, 0] + R[0, 1] * X[:, 1] + R[0, 2] * X[:, 2] + t[0]
    y = R[1, 0] * X[:, 0] + R[1, 1] * X[:, 1] + R[1, 2] * X[:, 2] + t[1]
    z = R[2, 0] * X[:, 0] + R[2, 1] * X[:, 1] + R[2, 2] * X[:, 2] + t[2]
    with np.errstate(divide="ignore"):
        s = np.where(z != 0.0, 1.0 / z, 1.0)
    x = x * s
    y = y * s
    return np.stack([x * K[0, 0] + K[0, 2], y * K[1, 1] + K[1, 2]], axis=1)


def reproj_error(X, pts, Rt, K) -> float:
    """ReprojectionError (sfm.py:79-100): rvec=Rodrigues(R) -> projectPoints -> both sides
    rounded to float32 -> cv2.norm(L2) (float32 differences, float64 accumulation) / N.
    X (N,3), pts (N,2)."""
    Rt = np.asarray(Rt, np.float64)
    R = rodrigues_to_matrix(rodrigues_to_vector(Rt[:, :3]))
    p = project_pinhole(np.asarray(X, np.float32), R, Rt[:, 3], K).astype(np.float32)
    d = (p - np.asarray(pts, np.float32)).astype(np.float64)
    return math.sqrt(float((d * d).sum())) / len(p)


# --------------------------------------------------------------------------- solvePnPRansac loop
class CvRNG:
    """cv::RNG — multiply-with-carry, the generator solvePnPRansac seeds with (uint64)-1."""
    COEF = 4164903690

    def __init__(self, state: int = 0xFFFFFFFFFFFFFFFF):
        self.state = state & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.state = ((self.state & 0xFFFFFFFF) * self.COEF + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a: int, b: int) -> int:
        return a if a == b else a + self.next() % (b - a)


def ransac_subsets(n: int, n_iters: int = 100, m: int = 5) -> np.ndarray:
    """The index stream RANSACPointSetRegistrator::getSubset draws: m distinct indices per
    iteration, redrawing on duplicates.  Depends only on n.  Returns (n_iters, m) int32."""
    rng = CvRNG()
    out = np.empty((n_iters, m), dtype=np.int32)
    for it in range(n_iters):
        i = 0
        while i < m:
            while True:
                j = rng.uniform(0, n)
                if all(j != out[it, k] for k in range(i)):
                    break
            out[it, i] = j
            i += 1
    return out


def update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    """cv::RANSACUpdateNumIters."""
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, np.finfo(np.float64).tiny)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < np.finfo(np.float64).tiny:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))      # cvRound: round-half-even


def score_hypothesis(X32, p32, R, t, K, thr: float = 8.0):
    """PnPRansacCallback::computeError + findInliers: projection in float64, stored float32,
    err = dx*dx + dy*dy in float32 (separately rounded products), inlier iff err <= thr^2."""
    proj = project_pinhole(X32, R, t, K).astype(np.float32)
    d = p32 - proj
    err = (d[:, 0] * d[:, 0]).astype(np.float32) + (d[:, 1] * d[:, 1]).astype(np.float32)
    return err <= np.float32(thr * thr)


def replay_ransac(counts: np.ndarray, valid: np.ndarray, n: int, max_iters: int = 100,
                  conf: float = 0.99, m: int = 5):
    """The accept/stop recursion of RANSACPointSetRegistrator::run over a vector of
    per-hypothesis inlier counts.  Returns (best_iter or -1, iterations_executed)."""
    niters = max_iters
    best, best_it = 0, -1
    it = 0
    while it < niters:
        if valid[it] and counts[it] > max(best, m - 1):
            best, best_it = int(counts[it]), it
            niters = update_num_iters(conf, (n - best) / n, m, niters)
        it += 1
    return best_it, it


def pnp_ransac(X, p, K, minimal_solver, max_iters: int = 100, thr: float = 8.0, conf: float = 0.99):
    """Restated cv2.solvePnPRansac loop (defaults as the reference gets them at sfm.py:67), up to
    but excluding the final LM refinement.  `minimal_solver(X5, p5) -> (ok, rvec, tvec)` is the
    5-point EPnP.  Returns dict(ok, best_iter, iters, mask, rvec, tvec, counts)."""
    X32 = np.ascontiguousarray(X, np.float32).reshape(-1, 3)
    p32 = np.ascontiguousarray(p, np.float32).reshape(-1, 2)
    n = len(X32)
    subsets = ransac_subsets(n, max_iters)
    niters = max_iters
    best, best_it, best_mask, best_pose = 0, -1, None, None
    counts = np.full(max_iters, -1, dtype=np.int64)
    it = 0
    while it < niters:
        s = subsets[it]
        ok, rvec, tvec = minimal_solver(X32[s], p32[s])
        if ok:
            mask = score_hypothesis(X32, p32, rodrigues_to_matrix(rvec), tvec, K, thr)
            c = int(mask.sum())
            counts[it] = c
            if c > max(best, 4):
                best, best_it, best_mask, best_pose = c, it, mask, (rvec, tvec)
                niters = update_num_iters(conf, (n - c) / n, 5, niters)
        it += 1
    if best_it < 0:
        return dict(ok=False, best_iter=-1, iters=it, mask=None, rvec=None, tvec=None, counts=counts)
    return dict(ok=True, best_iter=best_it, iters=it, mask=best_mask, rvec=best_pose[0],
                tvec=best_pose[1], counts=counts)


# --------------------------------------------------------------------------- bundle adjustment
def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def drot_drvec(rvec) -> np.ndarray:
    """dR/drvec as (3 params, 3, 3), closed form (Gallego & Yezzi 2015); first order at theta->0."""
    r = np.asarray(rvec, np.float64).ravel()
    th2 = float(r @ r)
    R = rodrigues_to_matrix(r)
    out = np.empty((3, 3, 3))
    if th2 < 1e-24:
        for i in range(3):
            e = np.zeros(3); e[i] = 1.0
            out[i] = _skew(e)
        return out
    S = _skew(r)
    for i in range(3):
        e = np.zeros(3); e[i] = 1.0
        out[i] = (r[i] * S + _skew(np.cross(r, (np.eye(3) - R) @ e))) @ R / th2
    return out


def ba_residual_jacobian(cams, pts, cam_idx, pt_idx, obs, K):
    """Engine BA formulation (SURVEY §8a decision): camera = (rvec, tvec), shared pinhole K,
    residual = project(X) - obs.  Returns r (O,2), Jc (O,2,6), Jp (O,2,3) in float64.
    Same projection as every reference variant (sfm.py:121, test.py:101, ba.pyc L24/L51)."""
    cams = np.asarray(cams, np.float64); pts = np.asarray(pts, np.float64)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    O = len(cam_idx)
    r = np.empty((O, 2)); Jc = np.empty((O, 2, 6)); Jp = np.empty((O, 2, 3))
    Rs = np.array([rodrigues_to_matrix(c[:3]) for c in cams])
    dRs = np.array([drot_drvec(c[:3]) for c in cams])
    for o in range(O):
        c, p = cam_idx[o], pt_idx[o]
        R, t, X = Rs[c], cams[c, 3:], pts[p]
        Y = R @ X + t
        iz = 1.0 / Y[2]
        u = fx * Y[0] * iz + cx
        v = fy * Y[1] * iz + cy
        r[o] = (u - obs[o, 0], v - obs[o, 1])
        dpdY = np.array([[fx * iz, 0.0, -fx * Y[0] * iz * iz], [0.0, fy * iz, -fy * Y[1] * iz * iz]])
        Jc[o, :, 3:] = dpdY
        for k in range(3):
            Jc[o, :, k] = dpdY @ (dRs[c, k] @ X)
        Jp[o] = dpdY @ R
    return r, Jc, Jp


def ba_normal_equations(r, Jc, Jp, cam_idx, pt_idx, n_cam, n_pt):
    """Block normal equations and the point-eliminated reduced camera system.
    Returns Hcc (C,6,6), bc (C,6), Hpp (P,3,3), bp (P,3), and dense S (6C,6C), g (6C,) for lambda=0."""
    Hcc = np.zeros((n_cam, 6, 6)); bc = np.zeros((n_cam, 6))
    Hpp = np.zeros((n_pt, 3, 3)); bp = np.zeros((n_pt, 3))
    for o in range(len(cam_idx)):
        c, p = cam_idx[o], pt_idx[o]
        Hcc[c] += Jc[o].T @ Jc[o]; bc[c] += Jc[o].T @ r[o]
        Hpp[p] += Jp[o].T @ Jp[o]; bp[p] += Jp[o].T @ r[o]
    return Hcc, bc, Hpp, bp


def ba_schur(r, Jc, Jp, cam_idx, pt_idx, n_cam, n_pt, lam: float = 0.0):
    """S = (Hcc + lam*diag) - sum_p W_p (Hpp + lam*diag)^-1 W_p^T ; g = bc - sum_p W_p Hpp^-1 bp."""
    Hcc, bc, Hpp, bp = ba_normal_equations(r, Jc, Jp, cam_idx, pt_idx, n_cam, n_pt)
    S = np.zeros((6 * n_cam, 6 * n_cam)); g = bc.reshape(-1).copy()
    for c in range(n_cam):
        H = Hcc[c] + lam * np.diag(np.diag(Hcc[c]))
        S[6 * c:6 * c + 6, 6 * c:6 * c + 6] = H
    order = np.argsort(pt_idx, kind="stable")
    start = np.searchsorted(pt_idx[order], np.arange(n_pt + 1))
    for p in range(n_pt):
        obs_p = order[start[p]:start[p + 1]]
        if obs_p.size == 0:
            continue
        Hinv = np.linalg.inv(Hpp[p] + lam * np.diag(np.diag(Hpp[p])))
        W = [Jc[o].T @ Jp[o] for o in obs_p]
        for a, oa in enumerate(obs_p):
            ca = cam_idx[oa]
            g[6 * ca:6 * ca + 6] -= W[a] @ Hinv @ bp[p]
            for b, ob in enumerate(obs_p):
                cb = cam_idx[ob]
                S[6 * ca:6 * ca + 6, 6 * cb:6 * cb + 6] -= W[a] @ Hinv @ W[b].T
    return S, g, Hcc, bc, Hpp, bp


def reduced_solve_pcg(S, g, tol: float = 1e-8, max_iter: int = 400):
    """S x = -g by block-Jacobi (6x6) preconditioned conjugate gradients — the algorithm of the engine's LM-step solver
    (csrc/pcg.cu) in float64 numpy: x0 = 0, the iteration ends when the recursively updated residual is below
    tol * |g|, and the TRUE residual b - S x must then be within min(100 tol, 1e-4) of |g|.  Only the lower triangle of
    S is read.  Returns (x, solved, iterations); solved is False on a non-positive diagonal block or p^T S p, on
    max_iter, or when the true residual fails the check (the engine then runs the factorisation)."""
    S = np.asarray(S, np.float64)
    S = np.tril(S) + np.tril(S, -1).T
    b = -np.asarray(g, np.float64).ravel()
    n = b.size
    C = n // 6
    Minv = np.empty((C, 6, 6))
    for a in range(C):
        blk = S[6 * a:6 * a + 6, 6 * a:6 * a + 6]
        try:
            L = np.linalg.cholesky(blk)
        except np.linalg.LinAlgError:
            return np.zeros(n), False, 0
        Li = np.linalg.inv(L)
        Minv[a] = Li.T @ Li
    prec = lambda r: np.einsum("cij,cj->ci", Minv, r.reshape(C, 6)).ravel()
    x = np.zeros(n)
    r = b.copy()
    bb = float(r @ r)
    if bb == 0.0:
        return x, True, 0
    z = prec(r)
    p = z.copy()
    rz = float(r @ z)
    for it in range(1, max_iter + 1):
        w = S @ p
        pw = float(p @ w)
        if not (pw > 0.0) or not np.isfinite(pw):
            return x, False, it - 1
        alpha = rz / pw
        x += alpha * p
        r -= alpha * w
        rr = float(r @ r)
        z = prec(r)
        rz_new = float(r @ z)
        if not np.isfinite(rr) or not np.isfinite(rz_new):
            return x, False, it
        if rr <= tol * tol * bb:
            d = b - S @ x
            return x, bool(d @ d <= min(1e4 * tol * tol, 1e-8) * bb), it
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, False, max_iter


# ------------------------------------------------------------------ recoverPose (sfm.py:311, isfm.py:83, test.py:250)
def recover_pose(E, p1, p2, K, dist: float = 50.0, mask=None):
    """Published algorithm of cv2.recoverPose restated in numpy (OpenCV 4.x, calib3d/five-point.cpp): normalise the
    pixels with K, decompose E = U diag(1,1,0) V^T into R1 = U W V^T, R2 = U W^T V^T, t = u3 (det(U), det(V^T) made
    positive), triangulate every correspondence against [R1|t], [R2|t], [R1|-t], [R2|-t] (DLT, smallest right
    singular vector) and keep it when z*w > 0, z/w < dist and the depth in the second camera is in (0, dist); the
    first candidate whose count is >= all others wins.  -> (count, R, t (3,1), mask (N,) bool)."""
    E = np.asarray(E, np.float64)
    K = np.asarray(K, np.float64)
    a = np.asarray(p1, np.float64).reshape(-1, 2).copy()
    b = np.asarray(p2, np.float64).reshape(-1, 2).copy()
    for q in (a, b):
        q[:, 0] = (q[:, 0] - K[0, 2]) / K[0, 0]
        q[:, 1] = (q[:, 1] - K[1, 2]) / K[1, 1]
    U, _, Vt = np.linalg.svd(E)
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    R1, R2, t = U @ W @ Vt, U @ W.T @ Vt, U[:, 2].copy()
    n = len(a)
    keep_in = np.ones(n, bool) if mask is None else (np.asarray(mask).reshape(-1) != 0)
    P0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    cands = [(R1, t), (R2, t), (R1, -t), (R2, -t)]
    masks = []
    for R, tt in cands:
        P = np.hstack([R, tt.reshape(3, 1)])
        A = np.empty((n, 4, 4))
        A[:, 0] = a[:, 0:1] * P0[2] - P0[0]
        A[:, 1] = a[:, 1:2] * P0[2] - P0[1]
        A[:, 2] = b[:, 0:1] * P[2] - P[0]
        A[:, 3] = b[:, 1:2] * P[2] - P[1]
        Q = np.linalg.svd(A)[2][:, 3, :]                 # (n,4) null vectors
        m = Q[:, 2] * Q[:, 3] > 0
        Xh = Q / Q[:, 3:4]
        m &= Xh[:, 2] < dist
        z2 = Xh @ P[2]
        m &= (z2 > 0) & (z2 < dist)
        masks.append(m & keep_in)
    g = [int(m.sum()) for m in masks]
    if g[0] >= g[1] and g[0] >= g[2] and g[0] >= g[3]:
        w = 0
    elif g[1] >= g[0] and g[1] >= g[2] and g[1] >= g[3]:
        w = 1
    elif g[2] >= g[0] and g[2] >= g[1] and g[2] >= g[3]:
        w = 2
    else:
        w = 3
    return g[w], cands[w][0], cands[w][1].reshape(3, 1), masks[w]


# ------------------------------------------------------------------ findEssentialMat (sfm.py:307, isfm.py:80, test.py:247)
# Monomials of degree <= 3 in (x, y, z), in the elimination order of Nister's five-point method: the first ten are
# eliminated (Gauss-Jordan), the last ten are x*[z^2 z 1], y*[z^2 z 1], [z^3 z^2 z 1].
_E5_MONO = [(3, 0, 0), (0, 3, 0), (2, 1, 0), (1, 2, 0), (2, 0, 1), (2, 0, 0), (0, 2, 1), (0, 2, 0), (1, 1, 1), (1, 1, 0),
            (1, 0, 2), (1, 0, 1), (1, 0, 0), (0, 1, 2), (0, 1, 1), (0, 1, 0), (0, 0, 3), (0, 0, 2), (0, 0, 1), (0, 0, 0)]
_E5_INDEX = {m: i for i, m in enumerate(_E5_MONO)}


def _pmul(a: dict, b: dict) -> dict:
    out: dict = {}
    for (ax, ay, az), av in a.items():
        for (bx, by, bz), bv in b.items():
            k = (ax + bx, ay + by, az + bz)
            out[k] = out.get(k, 0.0) + av * bv
    return out


def _padd(a: dict, b: dict, sb: float = 1.0) -> dict:
    out = dict(a)
    for k, v in b.items():
        out[k] = out.get(k, 0.0) + sb * v
    return out


def five_point_constraints(basis: np.ndarray) -> np.ndarray:
    """The 10 x 20 coefficient matrix of det(E) = 0 and 2 E E^T E - tr(E E^T) E = 0 for E = x X + y Y + z Z + W
    (basis rows X, Y, Z, W as 9-vectors), columns in _E5_MONO order.  Built by polynomial arithmetic."""
    lin = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]
    E = [[{lin[k]: float(basis[k, 3 * i + j]) for k in range(4)} for j in range(3)] for i in range(3)]
    EEt = [[_padd(_padd(_pmul(E[i][0], E[j][0]), _pmul(E[i][1], E[j][1])), _pmul(E[i][2], E[j][2]))
            for j in range(3)] for i in range(3)]
    tr = _padd(_padd(EEt[0][0], EEt[1][1]), EEt[2][2])
    rows = []
    det = _padd(_padd(_pmul(E[0][0], _padd(_pmul(E[1][1], E[2][2]), _pmul(E[1][2], E[2][1]), -1.0)),
                      _pmul(E[0][1], _padd(_pmul(E[1][0], E[2][2]), _pmul(E[1][2], E[2][0]), -1.0)), -1.0),
                _pmul(E[0][2], _padd(_pmul(E[1][0], E[2][1]), _pmul(E[1][1], E[2][0]), -1.0)))
    rows.append(det)
    for i in range(3):
        for j in range(3):
            L = [_padd(EEt[i][k], tr, -0.5) if i == k else EEt[i][k] for k in range(3)]
            rows.append(_padd(_padd(_pmul(L[0], E[0][j]), _pmul(L[1], E[1][j])), _pmul(L[2], E[2][j])))
    M = np.zeros((10, 20))
    for r, p in enumerate(rows):
        for k, v in p.items():
            M[r, _E5_INDEX[k]] = v
    return M


def five_point(q1: np.ndarray, q2: np.ndarray) -> np.ndarray:
    """EMEstimatorCallback::runKernel of OpenCV's five-point.cpp (Nister 2004) restated: null space of the 5 x 9
    epipolar system, the ten cubic constraints reduced against the first ten monomials, the 3 x 3 polynomial
    matrix B(z) whose determinant is the degree-10 polynomial, one E = xX + yY + zZ + W (unit Frobenius norm) per
    real root.  Normalised coordinates q1, q2 (5,2).  Returns (k,3,3), k <= 10.

    Model ORDER: OpenCV visits the roots in the order its Durand-Kerner iteration leaves them, and the roots
    themselves depend on the null-space basis its SVD happens to return — neither is part of the published
    algorithm.  Here (and in the CUDA kernel) an iteration's models are ordered by ascending E[0,0]^2, which
    does not depend on the basis; the order only matters when two models of ONE sample tie on the inlier count."""
    q1 = np.asarray(q1, np.float64).reshape(-1, 2)
    q2 = np.asarray(q2, np.float64).reshape(-1, 2)
    # x2^T E x1 = 0 with E row-major in the 9-vector
    Q = np.stack([q2[:, 0] * q1[:, 0], q2[:, 0] * q1[:, 1], q2[:, 0], q2[:, 1] * q1[:, 0], q2[:, 1] * q1[:, 1],
                  q2[:, 1], q1[:, 0], q1[:, 1], np.ones(len(q1))], axis=1)
    basis = np.linalg.svd(Q, full_matrices=True)[2][5:9]
    M = five_point_constraints(basis)
    try:
        A = np.linalg.solve(M[:, :10], M[:, 10:])
    except np.linalg.LinAlgError:
        return np.zeros((0, 3, 3))
    # rows 4/5, 6/7, 8/9 are <x^2 z>/<x^2>, <y^2 z>/<y^2>, <xyz>/<xy>: (row_a - z row_b) = x p3(z) + y q3(z) + r4(z)
    B = [[None] * 3 for _ in range(3)]
    for i in range(3):
        ra, rb = A[2 * i + 4], A[2 * i + 5]
        for c, (lo, hi) in enumerate(((0, 3), (3, 6), (6, 10))):
            pa = np.concatenate([[0.0], ra[lo:hi]])        # descending powers of z
            pb = np.concatenate([rb[lo:hi], [0.0]])
            B[i][c] = pa - pb
    mul = np.polymul
    det = (np.polysub(mul(mul(B[0][0], B[1][1]), B[2][2]), mul(mul(B[0][0], B[1][2]), B[2][1]))
           - np.polysub(mul(mul(B[0][1], B[1][0]), B[2][2]), mul(mul(B[0][1], B[1][2]), B[2][0]))
           + np.polysub(mul(mul(B[0][2], B[1][0]), B[2][1]), mul(mul(B[0][2], B[1][1]), B[2][0])))
    if not np.all(np.isfinite(det)):
        return np.zeros((0, 3, 3))
    roots = np.roots(det)
    models = []
    for r in roots:
        if abs(r.imag) > 1e-10:
            continue
        z = float(r.real)
        Bz = np.array([[np.polyval(B[i][c], z) for c in range(3)] for i in range(3)])
        xy1 = np.linalg.svd(Bz)[2][2]
        if abs(xy1[2]) < 1e-10:
            continue
        x, y = xy1[0] / xy1[2], xy1[1] / xy1[2]
        Ev = x * basis[0] + y * basis[1] + z * basis[2] + basis[3]
        Ev = Ev / np.linalg.norm(Ev)
        models.append(Ev.reshape(3, 3))
    models.sort(key=lambda e: e[0, 0] * e[0, 0])
    return np.array(models).reshape(-1, 3, 3)


def sampson_error(q1: np.ndarray, q2: np.ndarray, E: np.ndarray) -> np.ndarray:
    """EMEstimatorCallback::computeError: (x2^T E x1)^2 / (|E x1|_xy^2 + |E^T x2|_xy^2), float64, stored float32."""
    x1 = np.column_stack([q1, np.ones(len(q1))])
    x2 = np.column_stack([q2, np.ones(len(q2))])
    Ex1 = x1 @ E.T
    Etx2 = x2 @ E
    num = (x2 * Ex1).sum(1)
    den = Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        return (num * num / den).astype(np.float32)


def find_essential_mat(p1, p2, K, prob: float = 0.999, threshold: float = 1.0, max_iters: int = 1000):
    """cv2.findEssentialMat(p1, p2, K, method=RANSAC, prob, threshold, maxIters) restated (five-point.cpp +
    ptsetreg.cpp): pixels normalised with K in float64, threshold divided by the mean focal length, RNG(2^64-1)
    five-index subsets, every model of every sample scored with the Sampson error (float32, err <= thr^2), a model
    accepted when its count beats max(best, 4), the iteration budget shrunk by RANSACUpdateNumIters.  No refit: the
    returned E is the winning minimal-sample model.  -> (E (3,3) or None, mask (N,) bool, info)."""
    K = np.asarray(K, np.float64)
    q1 = np.asarray(p1, np.float64).reshape(-1, 2).copy()
    q2 = np.asarray(p2, np.float64).reshape(-1, 2).copy()
    n = len(q1)
    for q in (q1, q2):
        q[:, 0] = (q[:, 0] - K[0, 2]) / K[0, 0]
        q[:, 1] = (q[:, 1] - K[1, 2]) / K[1, 1]
    thr = threshold / ((K[0, 0] + K[1, 1]) / 2)
    t2 = np.float32(thr * thr)
    if n < 5:
        return None, np.zeros(n, bool), dict(iters=0, best_iter=-1, best_model=-1)
    if n == 5:
        models = five_point(q1, q2)
        if len(models) == 0:
            return None, np.zeros(n, bool), dict(iters=0, best_iter=-1, best_model=-1)
        return models.reshape(-1, 3), np.ones(n, bool), dict(iters=0, best_iter=0, best_model=0)
    niters = max(max_iters, 1)
    rng = CvRNG()
    best, best_E, best_mask, best_it, best_m = 0, None, np.zeros(n, bool), -1, -1
    it = 0
    while it < niters:
        s = []
        while len(s) < 5:
            j = rng.uniform(0, n)
            if j not in s:
                s.append(j)
        models = five_point(q1[s], q2[s])
        for mi, E in enumerate(models):
            mask = sampson_error(q1, q2, E) <= t2
            c = int(mask.sum())
            if c > max(best, 4):
                best, best_E, best_mask, best_it, best_m = c, E, mask, it, mi
                niters = update_num_iters(prob, (n - c) / n, 5, niters)
        it += 1
    return best_E, best_mask, dict(iters=it, best_iter=best_it, best_model=best_m, best_count=best)
