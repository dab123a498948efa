"""numpy restatement of EPnP (Lepetit, Moreno-Noguer, Fua, IJCV 2009) as OpenCV runs it for
`cv2.solvePnP(flags=SOLVEPNP_EPNP)` — the 5-point minimal solver inside
`cv2.solvePnPRansac` (reference call site sfm.py:67).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Structure follows the published
algorithm and OpenCV's calling convention: image points are first normalised with K and
ROUNDED TO FLOAT32 (cv2.undistortPoints keeps the input dtype) and then mapped back to
pixels inside the solver (us = x_n*fu + uc), which runs in float64 with the real K; four control points from the PCA of the object points;
M^T M null-space; the three beta approximations (N=4 linearisation, N=2, N=3), 5
Gauss-Newton iterations each; the pose with the smallest reprojection error wins.
"""
from __future__ import annotations

import numpy as np


def normalise_points(p32: np.ndarray, K: np.ndarray) -> np.ndarray:
    ifx, ify = 1.0 / K[0, 0], 1.0 / K[1, 1]
    x = (p32[:, 0].astype(np.float64) - K[0, 2]) * ifx
    y = (p32[:, 1].astype(np.float64) - K[1, 2]) * ify
    return np.stack([x, y], 1).astype(np.float32).astype(np.float64)


def _control_points(pw):
    c0 = pw.mean(0)
    d = pw - c0
    w, v = np.linalg.eigh(d.T @ d)              # ascending
    order = np.argsort(-w)
    cws = [c0]
    for i in range(3):
        k = np.sqrt(max(w[order[i]], 0.0) / len(pw))
        cws.append(c0 + k * v[:, order[i]])
    return np.array(cws)


def _alphas(pw, cws):
    CC = (cws[1:] - cws[0]).T
    a123 = (np.linalg.inv(CC) @ (pw - cws[0]).T).T
    return np.c_[1.0 - a123.sum(1), a123]


def _L6x10(v):
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    dv = np.array([[vi[3 * a:3 * a + 3] - vi[3 * b:3 * b + 3] for a, b in pairs] for vi in v])
    L = np.empty((6, 10))
    for i in range(6):
        d0, d1, d2, d3 = dv[0, i], dv[1, i], dv[2, i], dv[3, i]
        L[i] = [d0 @ d0, 2 * d0 @ d1, d1 @ d1, 2 * d0 @ d2, 2 * d1 @ d2, d2 @ d2,
                2 * d0 @ d3, 2 * d1 @ d3, 2 * d2 @ d3, d3 @ d3]
    return L


def _rho(cws):
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    return np.array([((cws[a] - cws[b]) ** 2).sum() for a, b in pairs])


def _betas_approx_1(L, rho):
    b4 = np.linalg.lstsq(L[:, [0, 1, 3, 6]], rho, rcond=None)[0]
    if b4[0] < 0:
        b0 = np.sqrt(-b4[0]); return np.array([b0, -b4[1] / b0, -b4[2] / b0, -b4[3] / b0])
    b0 = np.sqrt(b4[0]); return np.array([b0, b4[1] / b0, b4[2] / b0, b4[3] / b0])


def _betas_approx_2(L, rho):
    b3 = np.linalg.lstsq(L[:, [0, 1, 2]], rho, rcond=None)[0]
    if b3[0] < 0:
        b0 = np.sqrt(-b3[0]); b1 = np.sqrt(-b3[2]) if b3[2] < 0 else 0.0
    else:
        b0 = np.sqrt(b3[0]); b1 = np.sqrt(b3[2]) if b3[2] > 0 else 0.0
    if b3[1] < 0:
        b0 = -b0
    return np.array([b0, b1, 0.0, 0.0])


def _betas_approx_3(L, rho):
    b5 = np.linalg.lstsq(L[:, [0, 1, 2, 3, 4]], rho, rcond=None)[0]
    if b5[0] < 0:
        b0 = np.sqrt(-b5[0]); b1 = np.sqrt(-b5[2]) if b5[2] < 0 else 0.0
    else:
        b0 = np.sqrt(b5[0]); b1 = np.sqrt(b5[2]) if b5[2] > 0 else 0.0
    if b5[1] < 0:
        b0 = -b0
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.array([b0, b1, b5[3] / b0, 0.0])


def _gauss_newton(L, rho, b):
    b = b.copy()
    for _ in range(5):
        A = np.empty((6, 4)); r = np.empty(6)
        for i in range(6):
            l = L[i]
            A[i] = [2 * l[0] * b[0] + l[1] * b[1] + l[3] * b[2] + l[6] * b[3],
                    l[1] * b[0] + 2 * l[2] * b[1] + l[4] * b[2] + l[7] * b[3],
                    l[3] * b[0] + l[4] * b[1] + 2 * l[5] * b[2] + l[8] * b[3],
                    l[6] * b[0] + l[7] * b[1] + l[8] * b[2] + 2 * l[9] * b[3]]
            r[i] = rho[i] - (l[0] * b[0] * b[0] + l[1] * b[0] * b[1] + l[2] * b[1] * b[1] +
                             l[3] * b[0] * b[2] + l[4] * b[1] * b[2] + l[5] * b[2] * b[2] +
                             l[6] * b[0] * b[3] + l[7] * b[1] * b[3] + l[8] * b[2] * b[3] +
                             l[9] * b[3] * b[3])
        with np.errstate(all="ignore"):
            try:
                b = b + np.linalg.lstsq(A, r, rcond=None)[0]
            except np.linalg.LinAlgError:
                break
    return b


def _pose_from_betas(v, betas, alphas, pw, us, fu, fv, uc, vc):
    ccs = sum(betas[j] * v[j] for j in range(4)).reshape(4, 3)
    pcs = alphas @ ccs
    if pcs[0, 2] < 0:
        ccs, pcs = -ccs, -pcs
    pc0, pw0 = pcs.mean(0), pw.mean(0)
    ABt = (pcs - pc0).T @ (pw - pw0)
    U, _, Vt = np.linalg.svd(ABt)
    R = U @ Vt
    if np.linalg.det(R) < 0:
        R[2] = -R[2]
    t = pc0 - R @ pw0
    Y = pw @ R.T + t
    iz = 1.0 / Y[:, 2]
    e = np.sqrt((us[:, 0] - (uc + fu * Y[:, 0] * iz)) ** 2 + (us[:, 1] - (vc + fv * Y[:, 1] * iz)) ** 2).sum() / len(pw)
    return R, t, e


def epnp(X32: np.ndarray, p32: np.ndarray, K: np.ndarray):
    """Returns (R (3,3), t (3,)) float64."""
    pw = np.asarray(X32, np.float32).astype(np.float64).reshape(-1, 3)
    K = np.asarray(K, np.float64)
    fu, fv, uc, vc = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    # OpenCV hands the float32 *normalised* points to the solver together with the real camera
    # matrix, and the solver maps them back to pixels: us = x_n*fu + uc
    us = normalise_points(np.asarray(p32, np.float32).reshape(-1, 2), K)
    us = np.stack([us[:, 0] * fu + uc, us[:, 1] * fv + vc], 1)
    n = len(pw)
    cws = _control_points(pw)
    al = _alphas(pw, cws)
    M = np.zeros((2 * n, 12))
    for j in range(4):
        M[0::2, 3 * j] = al[:, j] * fu
        M[0::2, 3 * j + 2] = al[:, j] * (uc - us[:, 0])
        M[1::2, 3 * j + 1] = al[:, j] * fv
        M[1::2, 3 * j + 2] = al[:, j] * (vc - us[:, 1])
    w, vec = np.linalg.eigh(M.T @ M)            # ascending eigenvalues
    v = [vec[:, i] for i in range(4)]           # v[0] = smallest
    L, rho = _L6x10(v), _rho(cws)
    best = None
    with np.errstate(all="ignore"):
        for approx in (_betas_approx_1, _betas_approx_2, _betas_approx_3):
            b = _gauss_newton(L, rho, approx(L, rho))
            R, t, e = _pose_from_betas(v, b, al, pw, us, fu, fv, uc, vc)
            if best is None or e < best[2]:
                best = (R, t, e)
    return best[0], best[1]
