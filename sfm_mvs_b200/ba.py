"""The reference's orphaned `ba` module surface (`__pycache__/ba.cpython-37.pyc`, source ba.py L17-72,
reconstructed in SURVEY.md Appendix A), backed by the CUDA engine.

  ba.ReprojectionError(X, pts, Rt, K) -> (||proj - pts^T||_F / N, X)        pyc L44-63 (the live def)
  ba.bundle_adjustment(X, p, img, K, R, t)                                  pyc L66-72

As compiled, the reference's bundle_adjustment cannot run (its 5-argument call at L68 and its
3-argument least_squares callback at L72 both hit the 4-argument ReprojectionError) and returns
None; here it returns the refined (X, R, t) — the one documented deviation.
"""
from __future__ import annotations

import numpy as np

from . import engine as _e
from .cv2_compat import default_context


def ReprojectionError(X, pts, Rt, K, ctx=None):
    """pts is (2,N) as in the reference (it is compared against pts.T); X is (N,3) or (N,1,3)."""
    ctx = ctx or default_context()
    X3 = np.float32(np.asarray(X)).reshape(-1, 3)
    err, _, _ = ctx.reproj_error(X3, 0, np.float32(np.asarray(pts)), 0, Rt, K)
    return err, X


def bundle_adjustment(X, p, img, K, R, t, ctx=None, max_iters: int = 25):
    """Single-camera refinement of [pose ; X] (pyc L66-72: opt = [P.ravel(), X.ravel('F')], least_squares).
    `img` is unused, as in the reference.  p: (N,2) observed pixels.  Returns (X (N,3), R (3,3), t (3,1))."""
    ctx = ctx or default_context()
    X3 = np.asarray(X, np.float64).reshape(-1, 3)
    obs = np.asarray(p, np.float32).reshape(-1, 2)
    n = len(X3)
    cam = np.concatenate([_e.rodrigues_to_vector(np.asarray(R, np.float64)), np.asarray(t, np.float64).ravel()])
    prob = _e.BAProblem(ctx, 1, n, np.zeros(n, np.int32), np.arange(n, dtype=np.int32), obs, K)
    prob.set_params(cam.reshape(1, 6), X3)
    prob.solve(max_iters=max_iters, ftol=1e-10)
    cams, pts = prob.get_params()
    prob.close()
    return pts, _e.rodrigues_to_matrix(cams[0, :3]), cams[0, 3:].reshape(3, 1)
