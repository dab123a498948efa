"""ctypes loader for oracle/_build/libcvoracle.so (oracle/cv_epnp.c): OpenCV's small-matrix arithmetic on the PnP
path restated operation for operation in C.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference calls cv2.solvePnPRansac (sfm.py:67, test.py:319);
its 5-point minimal solver is OpenCV's EPnP, whose decompositions are all smaller than the 25 rows from which OpenCV
uses LAPACK, so they run OpenCV's own one-sided Jacobi — plain IEEE double arithmetic in a fixed order — and can be
restated exactly.  tests/test_oracle.py pins every function here against in-process cv2, bit for bit.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libcvoracle.so")


def build() -> str:
    res = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the oracle's C restatement failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def _load():
    if not os.path.isfile(LIB_PATH):
        build()
    return C.CDLL(LIB_PATH)


_lib = _load()
_P = lambda a: a.ctypes.data_as(C.c_void_p)


def svd(A: np.ndarray):
    """cv2.SVDecomp(A) for a tall or square float64 matrix with fewer than 25 rows: w (n,), U (m,n), Vt (n,n)."""
    A = np.ascontiguousarray(A, np.float64)
    m, n = A.shape
    assert m >= n and m < 25
    w, U, Vt = np.zeros(n), np.zeros((m, n)), np.zeros((n, n))
    _lib.cvo_svd(_P(A), m, n, _P(w), _P(U), _P(Vt))
    return w, U, Vt


def mul_transposed(M: np.ndarray) -> np.ndarray:
    """cv2.mulTransposed(M, aTa=True) below OpenCV's GEMM threshold."""
    M = np.ascontiguousarray(M, np.float64)
    out = np.zeros((M.shape[1], M.shape[1]))
    _lib.cvo_mul_transposed(_P(M), M.shape[0], M.shape[1], _P(out))
    return out


def invert_svd(A: np.ndarray) -> np.ndarray:
    """cv2.invert(A, flags=cv2.DECOMP_SVD)[1]"""
    A = np.ascontiguousarray(A, np.float64)
    out = np.zeros_like(A)
    _lib.cvo_invert_svd(_P(A), A.shape[0], _P(out))
    return out


def solve_svd(A: np.ndarray, b: np.ndarray) -> np.ndarray:
    """cv2.solve(A, b, flags=cv2.DECOMP_SVD)[1] for one right-hand side."""
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64).ravel()
    x = np.zeros(A.shape[1])
    _lib.cvo_solve_svd(_P(A), A.shape[0], A.shape[1], _P(b), _P(x))
    return x


def epnp(X32: np.ndarray, p32: np.ndarray, K: np.ndarray):
    """The raw output (R (3,3), t (3,)) of cv2.solvePnP(X, p, K, zeros, flags=SOLVEPNP_EPNP) before cv2.Rodrigues,
    for up to 16 correspondences; also returns a dict with the left singular vectors of M^T M, its singular values,
    the number of Jacobi sweeps and the winning beta approximation (1..3)."""
    X = np.ascontiguousarray(X32, np.float32).reshape(-1, 3)
    p = np.ascontiguousarray(p32, np.float32).reshape(-1, 2)
    K = np.ascontiguousarray(K, np.float64).reshape(9)
    R, t, dbg = np.zeros(9), np.zeros(3), np.zeros(160)
    n = _lib.cvo_epnp(_P(X), _P(p), len(X), _P(K), _P(R), _P(t), _P(dbg))
    assert n > 0, "cvo_epnp: more than 16 correspondences"
    return R.reshape(3, 3), t, dict(Ut=dbg[:144].reshape(12, 12).copy(), D=dbg[144:156].copy(), sweeps=int(dbg[156]), N=int(dbg[157]))
