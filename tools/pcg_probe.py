"""How many block-Jacobi PCG iterations would the reduced camera system of BASELINE configs[3] need?  (A probe for a
chain-free reduced solver: the tile Cholesky is a chain of 3000 dependent columns.)  Builds S, g on the engine at the
perturbed start and after a few LM iterations, then runs PCG on the host in float64.
  python tools/pcg_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import synth

ctx = sfm.Context(0)
pb = synth.ba_problem(500, 100000, 10, seed=0)
prob = sfm.BAProblem(ctx, 500, 100000, pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
prob.set_params(pb["cams0"], pb["pts0"])


def pcg(S, b, tol, maxit=3000):
    n = len(b)
    C = n // 6
    Minv = np.linalg.inv(S.reshape(C, 6, C, 6)[np.arange(C), :, np.arange(C), :])       # (C,6,6) diagonal blocks
    prec = lambda r: np.einsum("cij,cj->ci", Minv, r.reshape(C, 6)).ravel()
    x = np.zeros(n); r = b.copy(); z = prec(r); p = z.copy(); rz = r @ z
    b0 = np.linalg.norm(b)
    for k in range(1, maxit + 1):
        Sp = S @ p
        a = rz / (p @ Sp)
        x += a * p; r -= a * Sp
        if np.linalg.norm(r) <= tol * b0:
            return x, k
        z = prec(r); rz2 = r @ z
        p = z + (rz2 / rz) * p; rz = rz2
    return x, maxit


lam = 1e-3
for it in range(6):
    S, g, hd = prob.build_system(lam)
    S = S.astype(np.float64); S = np.tril(S) + np.tril(S, -1).T
    b = -g.astype(np.float64)
    ref = np.linalg.solve(S, b)
    ev = np.linalg.eigvalsh(S)
    line = [f"LM iteration {it}, lambda {lam:.0e}: cond(S) {ev[-1] / ev[0]:.2e}"]
    for tol in (1e-2, 1e-4, 1e-6, 1e-8):
        x, k = pcg(S, b, tol)
        line.append(f"tol {tol:.0e}: {k} it, |x - x*|/|x*| {np.linalg.norm(x - ref) / np.linalg.norm(ref):.1e}")
    print(" | ".join(line), flush=True)
    st = prob.gn_step(lam)
    lam = st["lambda_next"]
