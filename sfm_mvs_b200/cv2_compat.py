"""The reference's call surface, backed by the CUDA engine.

The reference (FlagArihant2000/sfm-mvs) has no plugin or operator layer: its boundary to the hot
path is the OpenCV Python API as called from sfm.py / isfm.py / test.py, plus its own helper
functions around those calls.  This module re-creates both with the same names, argument order,
shapes, dtypes and error behaviour, so the reference's scripts run on the engine by replacing
`cv2.X` with `sfm_mvs_b200.X` (or calling `patch_cv2()`):

  cv2.BFMatcher().knnMatch(des0, des1, k=2)      sfm.py:259-260, isfm.py:47,71, test.py:41-42,225,352
  cv2.triangulatePoints(P1, P2, x1, x2)          sfm.py:53, test.py:310,367
  cv2.solvePnPRansac(X, p, K, d, ...)            sfm.py:67, test.py:319
  Triangulation / ReprojectionError / PnP / common_points / BundleAdjustment   sfm.py:45-157,215-239
  find_features' matching half on arrays         sfm.py:259-268
"""
from __future__ import annotations

import numpy as np

from ._lib import error
from . import engine as _e

NORM_L2 = 4                   # cv2.NORM_L2
SOLVEPNP_ITERATIVE = 0        # cv2.SOLVEPNP_ITERATIVE
RATIO = 0.70                  # sfm.py:264

_default_ctx = None


def default_context() -> _e.Context:
    """Process-wide context on the current device (cuda:LOCAL_RANK under torchrun)."""
    global _default_ctx
    if _default_ctx is None:
        import os
        _default_ctx = _e.Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


def set_default_context(ctx: _e.Context | None):
    global _default_ctx
    _default_ctx = ctx


# --------------------------------------------------------------------------- matching
class DMatch:
    """cv2.DMatch look-alike: queryIdx, trainIdx, imgIdx, distance (a Python float holding a float32 value)."""

    __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

    def __init__(self, queryIdx=-1, trainIdx=-1, imgIdx=0, distance=float("inf")):
        self.queryIdx, self.trainIdx, self.imgIdx, self.distance = queryIdx, trainIdx, imgIdx, distance

    def __repr__(self):
        return f"DMatch(queryIdx={self.queryIdx}, trainIdx={self.trainIdx}, imgIdx={self.imgIdx}, distance={self.distance})"


class BFMatcher:
    """cv2.BFMatcher(normType=NORM_L2, crossCheck=False) — the only configuration the reference uses."""

    def __init__(self, normType: int = NORM_L2, crossCheck: bool = False, ctx: _e.Context | None = None):
        if normType != NORM_L2 or crossCheck:
            raise error(-4, "BFMatcher: only NORM_L2 without cross-check is provided (what the reference constructs)")
        self._ctx = ctx

    def knnMatch(self, queryDescriptors, trainDescriptors, k: int = 2):
        if k not in (1, 2):
            raise error(-4, "knnMatch: k must be 1 or 2")
        q, t = np.asarray(queryDescriptors), np.asarray(trainDescriptors)
        if q.ndim != 2 or t.ndim != 2:
            raise error(-1, "knnMatch: descriptors must be 2-D")
        if q.shape[0] == 0:
            return ()
        if t.shape[0] == 0:
            return tuple(() for _ in range(q.shape[0]))
        if q.dtype != t.dtype:
            raise error(-1, "knnMatch: query and train descriptor types differ")
        ctx = self._ctx or default_context()
        idx, dist, _, _ = ctx.knn2(q, t, RATIO)
        kk = min(k, t.shape[0])
        idl, dl = idx.tolist(), dist.astype(np.float64).tolist()
        return tuple(tuple(DMatch(i, idl[i][j], 0, dl[i][j]) for j in range(kk)) for i in range(len(idl)))


def knn2(des0, des1, ratio: float = RATIO, ctx: _e.Context | None = None):
    """Array form of knnMatch + the Lowe loop: idx (n,2) i32, dist (n,2) f32, good (n,) bool."""
    ctx = ctx or default_context()
    idx, dist, good, _ = ctx.knn2(des0, des1, ratio)
    return idx, dist, good.astype(bool)


def match_keypoints(kp0, des0, kp1, des1, ratio: float = RATIO, ctx: _e.Context | None = None):
    """Matching half of find_features (sfm.py:259-268) on arrays: kp (n,2) float32 stand for kp[i].pt.
    Returns pts0, pts1 (M,2) float32 in ascending queryIdx order."""
    idx, _, good = knn2(des0, des1, ratio, ctx)
    q = np.nonzero(good)[0]
    return np.float32(np.asarray(kp0)[q]), np.float32(np.asarray(kp1)[idx[q, 0]])


# --------------------------------------------------------------------------- triangulation
def _as_2xn(a, name):
    a = np.asarray(a)
    if a.dtype != np.float32:
        # cv2 also takes float64 points (and then answers in float64); the reference only ever passes
        # float32 and the engine's I/O is float32, so be loud instead of silently changing precision.
        raise error(-1, f"triangulatePoints: {name} must be float32 (got {a.dtype})")
    if a.ndim == 2 and a.shape[0] == 2:
        return np.ascontiguousarray(a)
    if a.ndim == 3 and a.shape[2] == 2 and 1 in a.shape[:2]:
        return np.ascontiguousarray(a.reshape(-1, 2).T)
    raise error(-1, f"triangulatePoints: {name} must be 2xN (or Nx1x2 / 1xNx2), got shape {a.shape}")


def triangulatePoints(projMatr1, projMatr2, projPoints1, projPoints2, ctx: _e.Context | None = None):
    """cv2.triangulatePoints: (4,N) homogeneous points, unit norm, dtype of the input points."""
    x1, x2 = _as_2xn(projPoints1, "projPoints1"), _as_2xn(projPoints2, "projPoints2")
    if x1.shape != x2.shape:
        raise error(-1, "triangulatePoints: point arrays differ in size")
    if x1.shape[1] == 0:
        raise error(-1, "triangulatePoints: no points (cv2 raises as well)")
    P1, P2 = np.asarray(projMatr1, np.float64), np.asarray(projMatr2, np.float64)
    if P1.shape != (3, 4) or P2.shape != (3, 4):
        raise error(-1, "triangulatePoints: projection matrices must be 3x4")
    return (ctx or default_context()).triangulate(P1, P2, x1, x2, 0, 0, False)


RANSAC = 8        # cv2.RANSAC
LMEDS = 4         # cv2.LMEDS


def findEssentialMat(points1, points2, cameraMatrix=None, method: int = RANSAC, prob: float = 0.999,
                     threshold: float = 1.0, maxIters: int = 1000, mask=None, focal=None, pp=None,
                     ctx: _e.Context | None = None):
    """cv2.findEssentialMat(pts0, pts1, K, method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None) as the
    reference calls it (sfm.py:307, isfm.py:80, test.py:247) -> (E (3,3) float64, mask (N,1) uint8 of 1 / 0).
    N == 5 returns every model of the single sample stacked as (3k,3) like cv2; N < 5 or no model returns
    (None, None) / (None, zeros).  Only the RANSAC method with a 3x3 camera matrix is implemented (the
    reference uses nothing else); LMEDS / USAC and the focal+pp overload raise.  `mask` is output-only in cv2."""
    from ._lib import check, lib
    if focal is not None or pp is not None or cameraMatrix is None or np.ndim(cameraMatrix) != 2:
        raise error(-1, "findEssentialMat: the 3x3 cameraMatrix overload is the one implemented")
    if int(method) != RANSAC:
        raise error(-1, "findEssentialMat: only method=cv2.RANSAC is implemented")
    p1, p2 = np.asarray(points1), np.asarray(points2)
    if p1.dtype not in (np.float32, np.float64):
        p1 = p1.astype(np.float64)
    p2 = p2.astype(p1.dtype, copy=False)
    if p1.size % 2 or p1.size != p2.size:
        raise error(-215, "findEssentialMat: npoints >= 0 && points2.checkVector(2) == npoints")
    p1 = np.ascontiguousarray(p1.reshape(-1, 2))
    p2 = np.ascontiguousarray(p2.reshape(-1, 2))
    K = np.ascontiguousarray(cameraMatrix, np.float64)
    if K.shape != (3, 3):
        raise error(-215, "findEssentialMat: cameraMatrix must be 3x3")
    if not (0.0 < float(prob) < 1.0):
        raise error(-215, "findEssentialMat: confidence > 0 && confidence < 1")
    n = p1.shape[0]
    if n < 5:
        return None, None
    ctx = ctx or default_context()
    E = np.zeros((10, 3, 3))
    m_out = np.zeros((n, 1), np.uint8)
    info = np.zeros(6, np.int32)
    check(lib.sfm_find_essential_mat(ctx._h, _e._dptr(p1), _e._dptr(p2), 0 if p1.dtype == np.float32 else 2, n,
                                     _e._dptr(K), float(prob), float(threshold), int(maxIters), _e._dptr(E),
                                     _e._dptr(m_out), _e._dptr(info)))
    k = int(info[0])
    findEssentialMat.last_info = dict(models=k, inliers=int(info[1]), iters=int(info[2]), best_iter=int(info[3]),
                                      best_model=int(info[4]), models_scored=int(info[5]))
    if k == 0:
        return None, m_out
    return E[:k].reshape(3 * k, 3).copy(), m_out


def recoverPose(E, points1, points2, cameraMatrix, R=None, t=None, mask=None, distanceThresh: float = 50.0,
                ctx: _e.Context | None = None):
    """cv2.recoverPose(E, pts0, pts1, K) as the reference calls it (sfm.py:311, isfm.py:83, test.py:250):
    -> (number of points passing the cheirality test, R (3,3), t (3,1), mask (N,1) uint8 with 255 = kept).
    points: (N,2) / (N,1,2) float32 or float64.  A given `mask` restricts the test to its non-zero rows (in/out
    argument of cv2)."""
    import ctypes as C
    from ._lib import check, lib
    ctx = ctx or default_context()
    p1, p2 = np.asarray(points1), np.asarray(points2)
    if p1.dtype not in (np.float32, np.float64):
        p1 = p1.astype(np.float64)
    p2 = p2.astype(p1.dtype, copy=False)
    p1 = np.ascontiguousarray(p1.reshape(-1, 2))
    p2 = np.ascontiguousarray(p2.reshape(-1, 2))
    if p1.shape != p2.shape or p1.shape[0] < 1:
        raise error(-1, "recoverPose: point arrays must both be (N,2) with N >= 1")
    E = np.ascontiguousarray(E, np.float64)
    K = np.ascontiguousarray(cameraMatrix, np.float64)
    if E.shape != (3, 3) or K.shape != (3, 3):
        raise error(-1, "recoverPose: E and cameraMatrix must be 3x3")
    n = p1.shape[0]
    m_in = None if mask is None else np.ascontiguousarray(np.asarray(mask).reshape(-1) != 0, np.uint8)
    if m_in is not None and m_in.shape[0] != n:
        raise error(-1, "recoverPose: mask length differs from the number of points")
    Rm, tv = np.zeros((3, 3)), np.zeros((3, 1))
    m_out = np.zeros((n, 1), np.uint8)
    good = C.c_int32(0)
    check(lib.sfm_recover_pose(ctx._h, _e._dptr(E), _e._dptr(p1), _e._dptr(p2), 0 if p1.dtype == np.float32 else 2, n,
                               _e._dptr(K), float(distanceThresh), None if m_in is None else _e._dptr(m_in), _e._dptr(Rm),
                               _e._dptr(tv), _e._dptr(m_out), C.byref(good)))
    return int(good.value), Rm, tv, m_out


def Triangulation(P1, P2, pts1, pts2, K=None, repeat=False, ctx: _e.Context | None = None):
    """sfm.py:45-56 — returns (pts1 as 2xN, pts2 as 2xN, cloud 4xN with row 3 == 1)."""
    a = pts1 if repeat else pts1.T
    b = pts2 if repeat else pts2.T
    a32, b32 = _as_2xn(a, "pts1"), _as_2xn(b, "pts2")
    cloud = (ctx or default_context()).triangulate(np.asarray(P1, np.float64), np.asarray(P2, np.float64), a32, b32,
                                                   0, 0, True)
    return a, b, cloud


# --------------------------------------------------------------------------- reprojection error
def ReprojectionError(X, pts, Rt, K, homogenity, ctx: _e.Context | None = None):
    """sfm.py:79-100 — (Frobenius norm of projected-observed)/N, X as cv2 returns it, projections (N,2) f32."""
    ctx = ctx or default_context()
    X = np.asarray(X)
    pts = np.asarray(pts)
    if homogenity == 1:
        err, proj, X3 = ctx.reproj_error(np.float32(X), 1, np.float32(pts), 0, Rt, K, True, True)
        Xout = X3.reshape(-1, 1, 3)                      # cv2.convertPointsFromHomogeneous shape
    else:
        X3 = np.float32(X).reshape(-1, 3)
        err, proj, _ = ctx.reproj_error(X3, 0, np.float32(pts), 1, Rt, K, True, False)
        Xout = X
    return err, Xout, proj


# --------------------------------------------------------------------------- PnP
def solvePnPRansac(objectPoints, imagePoints, cameraMatrix, distCoeffs, rvec=None, tvec=None,
                   useExtrinsicGuess=False, iterationsCount=100, reprojectionError=8.0, confidence=0.99,
                   inliers=None, flags=SOLVEPNP_ITERATIVE, ctx: _e.Context | None = None, hypotheses=None):
    """cv2.solvePnPRansac, positional-compatible (the reference's 5th positional lands in `rvec` and is
    ignored because useExtrinsicGuess is False).  Returns (retval, rvec (3,1), tvec (3,1), inliers (n,1) int32 | None)."""
    if useExtrinsicGuess or flags != SOLVEPNP_ITERATIVE:
        raise error(-4, "solvePnPRansac: only flags=SOLVEPNP_ITERATIVE without an extrinsic guess (the reference's call)")
    if distCoeffs is not None and np.any(np.asarray(distCoeffs) != 0):
        raise error(-4, "solvePnPRansac: non-zero distortion is not provided (the reference passes zeros)")
    X = np.asarray(objectPoints, np.float32).reshape(-1, 3)
    p = np.asarray(imagePoints, np.float32).reshape(-1, 2)
    if len(X) != len(p):
        raise error(-1, "solvePnPRansac: object/image point counts differ")
    ok, r, t, inl, _ = (ctx or default_context()).pnp_ransac(X, p, cameraMatrix, iterationsCount, reprojectionError,
                                                            confidence, hypotheses)
    if not ok:
        return False, np.zeros((3, 1)), np.zeros((3, 1)), None
    return True, r.reshape(3, 1), t.reshape(3, 1), inl.reshape(-1, 1).astype(np.int32)


def PnP(X, p, K, d, p_0, initial, ctx: _e.Context | None = None):
    """sfm.py:60-76."""
    if initial == 1:
        X = X[:, 0, :]
        p = p.T
        p_0 = p_0.T
    ok, rvec, t, inliers = solvePnPRansac(X, p, K, d, SOLVEPNP_ITERATIVE, ctx=ctx)
    R = _e.rodrigues_to_matrix(rvec)
    if inliers is not None:
        sel = inliers[:, 0]
        p, X, p_0 = p[sel], X[sel], p_0[sel]
    return R, t, p, X, p_0


def common_points(pts1, pts2, pts3, ctx: _e.Context | None = None):
    """sfm.py:215-239 — returns (indx1, indx2, temp_array1, temp_array2) with the reference's element-wise
    'x or y equal, first hit' semantics."""
    i1, i2, keep, _ = (ctx or default_context()).common_points(np.float32(pts1), np.float32(pts2))
    keep = keep.astype(bool)
    return i1.astype(np.int64), i2.astype(np.int64), np.asarray(pts2)[keep], np.asarray(pts3)[keep]


# --------------------------------------------------------------------------- BundleAdjustment (sfm.py:138-157)
def BundleAdjustment(points_3d, temp2, Rtnew, K, r_error, ctx: _e.Context | None = None):
    """sfm.py:138-157, same signature, same formulation, same optimiser: x0 = [Rt 12 | K 9 | temp2 (2,N) | points
    (N,3)] — the pose as twelve unconstrained numbers, K and the observed pixels free, as the reference has it —
    minimised by scipy.optimize.least_squares(gtol=r_error) (TRF, the reference's own driver, sfm.py:9,146).  What runs
    on the GPU is what the reference spends its "half a minute per frame" on (sfm.py:378): the residual
    OptimReprojectionError and the 22 + 5N residual evaluations of every finite-difference Jacobian, all in one launch
    with scipy's own difference steps (csrc/ba_ref.cu), so the iterates follow the reference's.
    Returns (X (N,3), p (N,2), Rt (3,4)) float64 like the reference."""
    from scipy.optimize import least_squares
    ctx = ctx or default_context()
    x0 = np.hstack((np.asarray(Rtnew, np.float64).ravel(), np.asarray(K, np.float64).ravel(),
                    np.asarray(temp2, np.float64).ravel(), np.asarray(points_3d, np.float64).ravel()))
    n = (len(x0) - 21) // 5
    if 21 + 5 * n != len(x0) or n < 1:
        raise _e.error(-1, f"BundleAdjustment: {len(x0)} parameters are not [Rt 12 | K 9 | p (2,N) | X (N,3)]")
    fun = lambda x: ctx.ba_reference_fd(x, n, want_jac=False)[0]
    jac = lambda x: ctx.ba_reference_fd(x, n)[1]
    sol = least_squares(fun=fun, x0=x0, jac=jac, gtol=r_error).x
    rest = int(len(sol[21:]) * 0.4)                   # sfm.py:148-155
    return sol[21 + rest:].reshape(-1, 3), sol[21:21 + rest].reshape(2, rest // 2).T, sol[0:12].reshape(3, 4)


def BundleAdjustmentSE3(points_3d, temp2, Rtnew, K, ctx: _e.Context | None = None, max_iters: int = 25, ftol: float = 1e-10):
    """The well-posed version of the same refinement, entirely on the GPU: the camera moves on SE(3) (rvec, tvec), K
    and the observations are data, analytic Jacobians, Schur-complement LM (BAProblem).  Same return shapes."""
    ctx = ctx or default_context()
    X = np.asarray(points_3d, np.float64).reshape(-1, 3)
    obs = np.asarray(temp2, np.float32)
    obs = obs.T if obs.shape[0] == 2 else obs
    n = len(X)
    Rt = np.asarray(Rtnew, np.float64)
    cam = np.concatenate([_e.rodrigues_to_vector(Rt[:, :3]), Rt[:, 3]])
    prob = _e.BAProblem(ctx, 1, n, np.zeros(n, np.int32), np.arange(n, dtype=np.int32), obs, K)
    prob.set_params(cam.reshape(1, 6), X)
    prob.solve(max_iters=max_iters, ftol=ftol)
    cams, pts = prob.get_params()
    prob.close()
    Rt_out = np.hstack([_e.rodrigues_to_matrix(cams[0, :3]), cams[0, 3:].reshape(3, 1)])
    return pts, np.float64(obs), Rt_out


def patch_cv2(cv2_module=None):
    """Route the reference's hot-path cv2 calls to the engine: after this, the reference's own
    function definitions run unmodified on the GPU.  Returns a dict of the originals (to undo)."""
    import cv2 as _cv2
    m = cv2_module or _cv2
    saved = dict(BFMatcher=m.BFMatcher, triangulatePoints=m.triangulatePoints, solvePnPRansac=m.solvePnPRansac,
                 recoverPose=m.recoverPose, findEssentialMat=m.findEssentialMat)
    m.BFMatcher = BFMatcher
    m.findEssentialMat = findEssentialMat
    m.triangulatePoints = triangulatePoints
    m.solvePnPRansac = solvePnPRansac
    m.recoverPose = recoverPose
    return saved


def unpatch_cv2(saved, cv2_module=None):
    import cv2 as _cv2
    m = cv2_module or _cv2
    for k, v in saved.items():
        setattr(m, k, v)
