"""Generate tests/golden/*.npz by running the REFERENCE'S OWN functions.

TEST INFRASTRUCTURE ONLY; runs in the build container only (needs /root/reference):

    python -m oracle.make_golden

The reference ships no tests or golden vectors for the hot path, so the
fixtures that pin the oracle are outputs of the reference's own `def`s (AST-loaded by
oracle/refload.py from /root/reference/sfm.py, executed against cv2 4.13.0) on seeded
inputs.  Only inputs and outputs are stored — no reference source.

Fixtures
  real_pair.npz   image.jpg (the one real frame shipped, README.md:33) -> reference
                  find_features(img, warped img): SIFT keypoints/descriptors of both
                  images (recomputed with the identical cv2 calls) + the pts0/pts1 the
                  reference returned  (sfm.py:242-270)
  geometry.npz    Triangulation / ReprojectionError / PnP / common_points outputs on a
                  seeded two-view problem with 25 % gross outliers (sfm.py:45-100,215-239)
  ba_small.npz    OptimReprojectionError residual and BundleAdjustment result, N=24
                  (sfm.py:104-157)
  chain.npz       the per-view loop sfm.py:341-409 driven over a 7-view synthetic scene
                  with the reference's Triangulation/common_points/PnP/ReprojectionError
"""
from __future__ import annotations

import contextlib
import io
import os

import cv2
import numpy as np

from oracle import refload
from sfm_mvs_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def real_pair(ref):
    img = cv2.imread(os.path.join(refload.REFERENCE_ROOT, "image.jpg"))
    img0 = ref["img_downscale"](img, 2)
    K = synth.K_GUSTAV
    # second view = plane-induced homography of a small camera motion (the real second
    # Gustav frame is not in the container)
    rvec = np.array([0.0, 0.03, 0.005]); t = np.array([[-0.25], [0.01], [0.02]])
    R, _ = cv2.Rodrigues(rvec)
    n = np.array([[0.0, 0.0, 1.0]]); d = 8.0
    H = K @ (R + t @ n / d) @ np.linalg.inv(K)
    img1 = cv2.warpPerspective(img0, H, (img0.shape[1], img0.shape[0]))
    pts0, pts1 = ref["find_features"](img0, img1)
    sift = cv2.SIFT_create()
    kp0, des0 = sift.detectAndCompute(cv2.cvtColor(img0, cv2.COLOR_BGR2GRAY), None)
    kp1, des1 = sift.detectAndCompute(cv2.cvtColor(img1, cv2.COLOR_BGR2GRAY), None)
    assert np.all(des0 == np.rint(des0)) and des0.max() <= 255
    np.savez_compressed(
        os.path.join(OUT, "real_pair.npz"),
        kp0=np.float32([k.pt for k in kp0]), kp1=np.float32([k.pt for k in kp1]),
        des0=des0.astype(np.uint8), des1=des1.astype(np.uint8), pts0=pts0, pts1=pts1)
    print("real_pair:", des0.shape, des1.shape, "good", pts0.shape)


def geometry(ref):
    rng = np.random.default_rng(7)
    K = synth.K_GUSTAV
    R0, t0 = synth.orbit_pose(0.0)
    R1, t1 = synth.orbit_pose(0.05)
    Rt0, Rt1 = np.hstack([R0, t0]), np.hstack([R1, t1])
    P1, P2 = K @ Rt0, K @ Rt1
    n = 400
    X = np.c_[rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n), rng.uniform(5, 11, n)]
    x0, _ = synth.project(K, R0, t0, X)
    x1, _ = synth.project(K, R1, t1, X)
    x0 = (x0 + rng.normal(0, 0.4, x0.shape)).astype(np.float32)
    x1 = (x1 + rng.normal(0, 0.4, x1.shape)).astype(np.float32)
    a, b, cloud = ref["Triangulation"](P1, P2, x0, x1, K, repeat=False)
    err, Xc, proj = ref["ReprojectionError"](cloud, b, Rt1, K, homogenity=1)
    # PnP with gross outliers (sfm.py:362 call shape: X (N,3) f32, p (N,2) f32)
    X3 = Xc[:, 0, :].copy()
    p = x1.copy()
    bad = rng.choice(n, n // 4, replace=False)
    p[bad] += rng.uniform(-80, 80, (len(bad), 2)).astype(np.float32)
    Rp, tp, p_in, X_in, p0_in = ref["PnP"](X3, p, K, np.zeros((5, 1), np.float32), x0.copy(), initial=0)
    ok, rvec, tvec, inl = cv2.solvePnPRansac(X3, p, K, np.zeros((5, 1), np.float32), cv2.SOLVEPNP_ITERATIVE)
    err0, _, proj0 = ref["ReprojectionError"](X_in, p_in, np.hstack([Rp, tp]), K, homogenity=0)
    # common_points: float-equality association incl. the x-or-y quirk
    ptsA = x1[rng.permutation(n)[:300]]
    ptsB = np.vstack([ptsA[rng.permutation(300)[:180]], rng.uniform(0, 900, (120, 2)).astype(np.float32)])
    ptsB[185, 0] = ptsA[3, 0]          # x-only coincidence -> still a "match" in the reference
    ptsB = ptsB[rng.permutation(300)]
    ptsC = rng.uniform(0, 900, (300, 2)).astype(np.float32)
    i1, i2, tA, tB = _quiet(ref["common_points"], ptsA, ptsB, ptsC)
    np.savez_compressed(
        os.path.join(OUT, "geometry.npz"), K=K, P1=P1, P2=P2, Rt1=Rt1, x0=x0, x1=x1,
        cloud=cloud, tri_err=err, tri_X=Xc, tri_proj=proj,
        pnp_X=X3, pnp_p=p, pnp_R=Rp, pnp_t=tp, pnp_inliers=inl, pnp_p_in=p_in, pnp_X_in=X_in,
        pnp_err=err0, pnp_proj=proj0,
        cp_A=ptsA, cp_B=ptsB, cp_C=ptsC, cp_i1=i1, cp_i2=i2, cp_tA=tA, cp_tB=tB)
    print("geometry: tri_err", err, "pnp inliers", len(inl), "common", len(i1))


def ba_small(ref):
    rng = np.random.default_rng(11)
    K = synth.K_GUSTAV
    R, t = synth.orbit_pose(0.08)
    Rt = np.hstack([R, t])
    n = 24
    X = np.c_[rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(5, 11, n)]
    uv, _ = synth.project(K, R, t, X)
    obs = (uv + rng.normal(0, 0.5, uv.shape)).astype(np.float32).T.copy()       # (2,N) like temp2
    X0 = (X + rng.normal(0, 0.02, X.shape)).astype(np.float32).reshape(n, 1, 3)  # like points_3d
    x = np.hstack((Rt.ravel(), K.ravel(), obs.ravel(), X0.ravel()))
    res = _quiet(ref["OptimReprojectionError"], x)
    Xo, po, Rto = _quiet(ref["BundleAdjustment"], X0, obs, Rt, K, 0.5)
    np.savez_compressed(os.path.join(OUT, "ba_small.npz"), K=K, Rt=Rt, obs=obs, X0=X0, x=x,
                        residual=res, ba_X=Xo, ba_p=po, ba_Rt=Rto)
    print("ba_small: cost0", float(res.sum()), "-> X shift", float(np.abs(Xo - X0[:, 0]).max()))


def ba_tracks():
    """test.py:85-113 — the multi-view residual of the track pipeline, executed from the reference's own def.  The def
    reads the module-level `track` (not its `tracks` argument) and slices it `track[:, i:i + 2]` for view i — both
    kept: the fixture records what the reference computes."""
    defs = refload.load_reference_defs("test.py")
    rng = np.random.default_rng(21)
    K = synth.K_GUSTAV
    img_tot, n = 3, 20
    X = np.c_[rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(5, 11, n)]
    poses = []
    for i in range(img_tot):
        R, t = synth.orbit_pose(0.05 * i)
        poses.append(np.hstack([R, t]).ravel())
    poses = np.array(poses)
    track = rng.uniform(100, 900, (n, 2 * img_tot))                # what view i is compared with: columns i, i + 1
    defs["OptimReprojectionError"].__globals__["track"] = track
    x = np.hstack((K.ravel(), poses.ravel(), X.ravel(), track.ravel()))
    res = _quiet(defs["OptimReprojectionError"], x, X.size, poses.size, track.size, img_tot)
    np.savez_compressed(os.path.join(OUT, "ba_tracks.npz"), K=K, poses=poses, cloud=X, track=track, x=x, residual=res,
                        img_tot=img_tot)
    print("ba_tracks: residual sum", float(res.sum()), res.shape)


def gustav_scene():
    """The reference's shipped artifacts as a fixture: pose.csv (K + 57 projection matrices, sfm.py:423) and
    Point_Cloud/sparse.ply (19 282 points, sfm.py:169-201) -> K, cameras as (rvec | tvec) = K^-1 P, points in world
    units (file / 200, sfm.py:170).  Read with the product's own readers (sfm_mvs_b200.io)."""
    from sfm_mvs_b200 import io as sio
    K, Ps = sio.load_poses(os.path.join(refload.REFERENCE_ROOT, "pose.csv"))
    pts, _ = sio.load_ply(os.path.join(refload.REFERENCE_ROOT, "Point_Cloud", "sparse.ply"))
    cams = []
    for P in Ps:
        Rt = np.linalg.inv(K) @ P
        U, _, Vt = np.linalg.svd(Rt[:, :3])
        cams.append(np.concatenate([cv2.Rodrigues(U @ Vt)[0].ravel(), Rt[:, 3]]))
    np.savez_compressed(os.path.join(OUT, "gustav_scene.npz"), K=K, cams=np.array(cams), pts=pts.astype(np.float32))
    print("gustav_scene:", len(cams), "cameras,", len(pts), "points")


def chain(ref):
    from oracle import cvpath  # matching half on arrays (find_features needs images)
    scene = synth.orbit_scene(7, 600, seed=3)
    K = scene["K"]
    st = cvpath.bootstrap_two_views(scene)
    # re-do the bootstrap with the reference's defs to be sure the state is the reference's
    v0, v1 = scene["views"][0], scene["views"][1]
    P1, P2 = st["P1"], st["P2"]
    Rt1 = np.hstack([v1["R"], v1["t"]])
    pts0, pts1 = cvpath.match_keypoints(v0["kp"], v0["des"], v1["kp"], v1["des"])
    pts0, pts1, X = ref["Triangulation"](P1, P2, pts0, pts1, K, repeat=False)
    _, X, _ = ref["ReprojectionError"](X, pts1, Rt1, K, homogenity=1)
    _, _, pts1, X, _ = ref["PnP"](X, pts1, K, np.zeros((5, 1), np.float32), pts0, initial=1)
    points_3d = X
    Rts, e1, e2, nnew, ninl = [], [], [], [], []
    Xnew_all = []
    for i in range(len(scene["views"]) - 2):
        va, vb = scene["views"][i + 1], scene["views"][i + 2]
        pts_, pts2 = cvpath.match_keypoints(va["kp"], va["des"], vb["kp"], vb["des"])
        if i != 0:
            pts0, pts1, points_3d = ref["Triangulation"](P1, P2, pts0, pts1, K, repeat=False)
            pts1 = pts1.T
            points_3d = cv2.convertPointsFromHomogeneous(points_3d.T)[:, 0, :]
        i1, i2, temp1, temp2 = _quiet(ref["common_points"], pts1, pts_, pts2)
        com2, com_ = pts2[i2], pts_[i2]
        Rot, trans, com2, points_3d, com_ = ref["PnP"](points_3d[i1], com2, K, np.zeros((5, 1), np.float32), com_, initial=0)
        Rtnew = np.hstack((Rot, trans)); Pnew = K @ Rtnew
        ea, points_3d, _ = ref["ReprojectionError"](points_3d, com2, Rtnew, K, homogenity=0)
        temp1, temp2, points_3d = ref["Triangulation"](P2, Pnew, temp1, temp2, K, repeat=False)
        eb, points_3d, _ = ref["ReprojectionError"](points_3d, temp2, Rtnew, K, homogenity=1)
        Rts.append(Rtnew); e1.append(ea); e2.append(eb); nnew.append(points_3d.shape[0]); ninl.append(len(com2))
        Xnew_all.append(points_3d[:, 0, :])
        P1, P2 = P2.copy(), Pnew.copy()
        pts0, pts1 = pts_.copy(), pts2.copy()
    np.savez_compressed(os.path.join(OUT, "chain.npz"), seed=3, n_views=7, n_pts=600,
                        Rt=np.array(Rts), err_pnp=np.array(e1), err_new=np.array(e2),
                        n_new=np.array(nnew), n_inl=np.array(ninl), X_new=np.vstack(Xnew_all))
    print("chain: err_new", e2)


def main():
    assert refload.available(), "needs /root/reference (build container only)"
    os.makedirs(OUT, exist_ok=True)
    ref = refload.load_reference_defs()
    real_pair(ref); geometry(ref); ba_small(ref); ba_tracks(); gustav_scene(); chain(ref)


if __name__ == "__main__":
    main()
