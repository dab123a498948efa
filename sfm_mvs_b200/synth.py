"""Seeded synthetic inputs for the hot path.

Everything here is input *generation* only (numpy on the host); nothing in this
module computes a result of the hot path.  The same generators feed the CUDA
engine, the oracle and the CPU baseline so that all three see identical bytes.

Recipes follow SURVEY.md §8d:
  * SIFT-like descriptors: integer valued float32 in [0, 255], ||d|| ~ 512,
    ~22 % zeros (what `cv2.SIFT` emits for the reference at sfm.py:246-252).
  * orbit scene for the per-view registration loop (sfm.py:341-409).
  * multi-camera bundle-adjustment problem (BASELINE.json configs[3]).
"""
from __future__ import annotations

import numpy as np

# Intrinsics the reference uses after `downscale = 2` (sfm.py:16-23, pose.csv rows 0-8).
K_GUSTAV = np.array(
    [
        [2393.952166119461 / 2.0, -3.410605131648481e-13, 932.3821770809047 / 2.0],
        [0.0, 2398.118540286656 / 2.0, 628.2649953288065 / 2.0],
        [0.0, 0.0, 1.0],
    ],
    dtype=np.float64,
)
IMG_W, IMG_H = 968.0, 648.0  # image.jpg (1296x1936) after img_downscale(.., 2)


# --------------------------------------------------------------------------- descriptors
def _sift_base(n: int, rng: np.random.Generator) -> np.ndarray:
    """Un-quantised unit-norm SIFT-like vectors (n,128) float64."""
    d = rng.exponential(1.0, size=(n, 128)) * (rng.random((n, 128)) < 0.78)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    d = np.minimum(d, 0.2)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    return d


def _quantise(d: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(512.0 * d), 0, 255).astype(np.float32)


def sift_like_descriptors(n: int, seed: int | np.random.Generator = 0) -> np.ndarray:
    """(n,128) float32, integer valued in [0,255] — the dtype/range of cv2 SIFT output."""
    rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
    return _quantise(_sift_base(n, rng))


def matching_pair(nq: int, nt: int, seed: int = 0, shared: float = 0.6, sigma: float = 0.012):
    """Two descriptor sets where `shared*min(nq,nt)` rows are noisy copies (rows permuted).

    Returns (des_q, des_t, gt) with gt[i] = train row matching query row i or -1.
    """
    rng = np.random.default_rng(seed)
    ns = int(shared * min(nq, nt))
    base = _sift_base(ns, rng)
    q = np.vstack([base, _sift_base(nq - ns, rng)])
    t = np.vstack([base + rng.normal(0.0, sigma, base.shape), _sift_base(nt - ns, rng)])
    pq = rng.permutation(nq)
    pt = rng.permutation(nt)
    inv_t = np.empty(nt, dtype=np.int64)
    inv_t[pt] = np.arange(nt)
    gt_unperm = np.full(nq, -1, dtype=np.int64)
    gt_unperm[:ns] = inv_t[:ns]
    return _quantise(q[pq]), _quantise(np.maximum(t, 0.0)[pt]), gt_unperm[pq]


# --------------------------------------------------------------------------- cameras
def orbit_pose(theta: float, radius: float = 8.0):
    """World->camera (R, t) of a camera on a circle about (0,0,radius) looking at it.

    theta = 0 is the identity pose, which is what the reference fixes for its first
    camera (sfm.py:277).
    """
    c = np.array([radius * np.sin(theta), 0.0, radius - radius * np.cos(theta)])
    R = np.array(
        [
            [np.cos(theta), 0.0, np.sin(theta)],
            [0.0, 1.0, 0.0],
            [-np.sin(theta), 0.0, np.cos(theta)],
        ]
    )
    return R, (-R @ c).reshape(3, 1)


def project(K, R, t, X):
    """Pinhole projection, float64. X (n,3) -> (n,2) pixels and depth (n,)."""
    Y = X @ R.T + t.reshape(1, 3)
    z = Y[:, 2]
    uv = np.stack([K[0, 0] * Y[:, 0] / z + K[0, 2], K[1, 1] * Y[:, 1] / z + K[1, 2]], axis=1)
    return uv, z


def orbit_scene(n_views: int, n_pts: int, seed: int = 0, shared: float = 0.6,
                step: float = 0.04, px_noise: float = 0.4, desc_sigma: float = 0.012,
                K: np.ndarray | None = None):
    """Synthetic incremental-SfM input: per view n_pts keypoints + SIFT-like descriptors.

    Returns dict with
      K (3,3) f64; views: list of dict(kp (n,2) f32, des (n,128) f32, pid (n,) i64, R, t)
    About `shared` of every view's keypoints are re-observations of points of the
    previous view (noisy descriptor copies -> they survive the 0.70 ratio test); the
    rest are new points back-projected from uniform pixels at depth U(5,11).
    """
    rng = np.random.default_rng(seed)
    K = K_GUSTAV if K is None else K
    margin = 4.0
    views = []
    pts3d: list[np.ndarray] = []
    n_total = 0
    # only the previous view's points can be re-observed, so only its 3-D points / base descriptors
    # are kept (aligned with prev_ids) — O(V) memory and time
    prev_ids = np.zeros(0, dtype=np.int64)
    prev_X = np.zeros((0, 3))
    prev_bd = np.zeros((0, 128))
    for v in range(n_views):
        R, t = orbit_pose(step * v)
        sel = np.zeros(0, dtype=np.int64)
        keep_uv = np.zeros((0, 2))
        if v > 0 and prev_ids.size:
            uv, z = project(K, R, t, prev_X)
            ok = (z > 0.5) & (uv[:, 0] > margin) & (uv[:, 0] < IMG_W - margin) & \
                 (uv[:, 1] > margin) & (uv[:, 1] < IMG_H - margin)
            cand = np.nonzero(ok)[0]
            take = min(int(shared * n_pts), cand.size)
            sel = rng.choice(cand, size=take, replace=False)
            keep_uv = uv[sel]
        keep_ids = prev_ids[sel]
        n_new = n_pts - keep_ids.size
        uv_new = np.stack([rng.uniform(margin, IMG_W - margin, n_new),
                           rng.uniform(margin, IMG_H - margin, n_new)], axis=1)
        depth = rng.uniform(5.0, 11.0, n_new)
        Yc = np.stack([(uv_new[:, 0] - K[0, 2]) / K[0, 0] * depth,
                       (uv_new[:, 1] - K[1, 2]) / K[1, 1] * depth, depth], axis=1)
        Xw = (Yc - t.reshape(1, 3)) @ R          # R^T (Y - t)
        new_ids = np.arange(n_total, n_total + n_new, dtype=np.int64)
        n_total += n_new
        pts3d.append(Xw)
        new_bd = _sift_base(n_new, rng)
        ids = np.concatenate([keep_ids, new_ids])
        uv_all = np.concatenate([keep_uv, uv_new]) + rng.normal(0.0, px_noise, (n_pts, 2))
        bd = np.concatenate([prev_bd[sel], new_bd])
        des = _quantise(np.maximum(bd + rng.normal(0.0, desc_sigma, bd.shape), 0.0))
        perm = rng.permutation(n_pts)
        views.append(dict(kp=uv_all[perm].astype(np.float32), des=des[perm], pid=ids[perm], R=R, t=t))
        prev_ids = ids
        prev_X = np.concatenate([prev_X[sel], Xw])
        prev_bd = bd
    return dict(K=K.copy(), views=views, X=np.concatenate(pts3d))


# --------------------------------------------------------------------------- bundle adjustment
def _rodrigues_vec(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> rotation vector (generation only; plain log map)."""
    c = np.clip((np.trace(R) - 1.0) * 0.5, -1.0, 1.0)
    th = np.arccos(c)
    if th < 1e-12:
        return np.zeros(3)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2.0 * np.sin(th))
    return w * th


def ba_problem(n_cam: int = 500, n_pt: int = 100_000, obs_per_pt: int = 10, seed: int = 0,
               px_noise: float = 0.5, cam_sigma: float = 1e-2, pt_sigma: float = 5e-2,
               K: np.ndarray | None = None):
    """BASELINE.json configs[3]: cameras on 5 orbit rings, points in a box in front of them,
    every point observed by `obs_per_pt` cameras that see it.  Observations are sorted
    point-major (all observations of a point contiguous) — the layout the engine shards by.

    Returns dict(K, cams_gt (C,6), pts_gt (P,3), cams0, pts0, cam_idx (O,) i32, pt_idx (O,) i32,
                 obs (O,2) f32).
    """
    rng = np.random.default_rng(seed)
    K = K_GUSTAV if K is None else K
    rings = 5
    per = (n_cam + rings - 1) // rings
    cams = []
    Rs, ts = [], []
    for c in range(n_cam):
        ring, k = divmod(c, per)
        theta = 2.0 * np.pi * k / per
        height = (ring - (rings - 1) / 2.0) * 1.2
        R, t = orbit_pose(theta)
        # lift the camera centre along world y and tilt it back toward the centre
        cen = -R.T @ t.ravel() + np.array([0.0, height, 0.0])
        f = np.array([0.0, 0.0, 8.0]) - cen
        f /= np.linalg.norm(f)
        r = np.cross(np.array([0.0, 1.0, 0.0]), f); r /= np.linalg.norm(r)
        d = np.cross(f, r)
        R = np.stack([r, d, f])
        t = -R @ cen
        Rs.append(R); ts.append(t)
        cams.append(np.concatenate([_rodrigues_vec(R), t]))
    cams = np.array(cams)
    Rs = np.array(Rs); ts = np.array(ts)
    pts = np.array([0.0, 0.0, 8.0]) + rng.uniform(-2.0, 2.0, size=(n_pt, 3))
    cam_idx = np.empty((n_pt, obs_per_pt), dtype=np.int32)
    obs = np.empty((n_pt, obs_per_pt, 2), dtype=np.float64)
    B = 8192
    for s in range(0, n_pt, B):
        X = pts[s:s + B]
        Y = np.einsum('cij,pj->pci', Rs, X) + ts[None]          # (p, C, 3)
        z = Y[..., 2]
        u = K[0, 0] * Y[..., 0] / z + K[0, 2]
        v = K[1, 1] * Y[..., 1] / z + K[1, 2]
        vis = (z > 0.5) & (u > 0) & (u < IMG_W) & (v > 0) & (v < IMG_H)
        score = rng.random(vis.shape) + (~vis) * 10.0            # visible cameras sort first
        pick = np.sort(np.argsort(score, axis=1)[:, :obs_per_pt], axis=1)
        cam_idx[s:s + B] = pick
        rows = np.arange(X.shape[0])[:, None]
        obs[s:s + B, :, 0] = u[rows, pick]
        obs[s:s + B, :, 1] = v[rows, pick]
    obs += rng.normal(0.0, px_noise, obs.shape)
    pt_idx = np.repeat(np.arange(n_pt, dtype=np.int32), obs_per_pt)
    return dict(
        K=K.copy(), cams_gt=cams, pts_gt=pts,
        cams0=cams + rng.normal(0.0, cam_sigma, cams.shape),
        pts0=pts + rng.normal(0.0, pt_sigma, pts.shape),
        cam_idx=cam_idx.reshape(-1).copy(), pt_idx=pt_idx,
        obs=obs.reshape(-1, 2).astype(np.float32),
    )


def two_view_pair(n: int, seed: int = 0, noise: float = 0.3, outliers: float = 0.2, dtype=np.float32, K=None):
    """Correspondences of a two-view initialisation (the input of sfm.py:307-311): n points in front of both
    cameras, a small rotation + sideways translation, Gaussian pixel noise and a fraction of gross mismatches.
    -> (pts0 (n,2), pts1 (n,2), R, t)."""
    rng = np.random.default_rng(seed)
    K = K_GUSTAV if K is None else K
    X = np.column_stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(4, 40, n)])
    rv = np.array([0.03, -0.25, 0.02]) + rng.normal(0, 0.02, 3)
    th = np.linalg.norm(rv)
    k = rv / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)
    t = np.array([1.0, 0.1, 0.2]) + rng.normal(0, 0.05, 3)

    def proj(Y):
        return np.column_stack([K[0, 0] * Y[:, 0] / Y[:, 2] + K[0, 2], K[1, 1] * Y[:, 1] / Y[:, 2] + K[1, 2]])
    p0 = proj(X) + rng.normal(0, noise, (n, 2))
    p1 = proj(X @ R.T + t) + rng.normal(0, noise, (n, 2))
    bad = rng.choice(n, int(outliers * n), replace=False)
    p1[bad] += rng.uniform(-150, 150, (len(bad), 2))
    return p0.astype(dtype), p1.astype(dtype), R, t


def ba_problem_from_scene(K, cams, pts, width: float = 648.0, height: float = 968.0, seed: int = 0, px_noise: float = 0.5,
                          cam_sigma: float = 2e-3, pt_sigma: float = 1e-2):
    """A bundle-adjustment problem from a reconstructed scene: every point that projects inside the width x height
    frame of a camera with positive depth is an observation of it (plus pixel noise); the start is the scene
    perturbed.  With tests/golden/gustav_scene.npz (the 57 cameras of the reference's pose.csv and the 19 282
    points of its Point_Cloud/sparse.ply) and the default 648 x 968 frame this is the 1 061 813-observation fixture
    problem of SURVEY section 4 (every camera sees 18.1-19.0 k of the points).
    Same dict as ba_problem; observations point-major."""
    rng = np.random.default_rng(seed)
    K = np.asarray(K, np.float64)
    cams = np.asarray(cams, np.float64)
    pts = np.asarray(pts, np.float64)
    Rs = np.array([_rodrigues_mat(c[:3]) for c in cams])
    ts = cams[:, 3:]
    ci, pi, ob = [], [], []
    B = 4096
    for s in range(0, len(pts), B):
        X = pts[s:s + B]
        Y = np.einsum('cij,pj->pci', Rs, X) + ts[None]
        z = Y[..., 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = K[0, 0] * Y[..., 0] / z + K[0, 2]
            v = K[1, 1] * Y[..., 1] / z + K[1, 2]
        vis = (z > 0) & (u >= 0) & (u < width) & (v >= 0) & (v < height)
        p_loc, c_loc = np.nonzero(vis)                                  # point-major already
        ci.append(c_loc.astype(np.int32))
        pi.append((p_loc + s).astype(np.int32))
        ob.append(np.stack([u[vis], v[vis]], 1))
    cam_idx, pt_idx = np.concatenate(ci), np.concatenate(pi)
    obs = np.concatenate(ob) + rng.normal(0.0, px_noise, (len(cam_idx), 2))
    seen = np.zeros(len(pts), bool)
    seen[pt_idx] = True
    if not seen.all():                                                   # drop points no camera sees, renumber
        remap = np.cumsum(seen) - 1
        pts, pt_idx = pts[seen], remap[pt_idx].astype(np.int32)
    cams0 = cams + rng.normal(0.0, cam_sigma, cams.shape)
    pts0 = pts + rng.normal(0.0, pt_sigma, pts.shape)
    return dict(K=K, cams_gt=cams, pts_gt=pts, cams0=cams0, pts0=pts0, cam_idx=cam_idx, pt_idx=pt_idx,
                obs=obs.astype(np.float32))


def _rodrigues_mat(r):
    r = np.asarray(r, np.float64)
    th = float(np.linalg.norm(r))
    if th < 1e-15:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * Kx
