"""compute-sanitizer target: every kernel family of the path once, at tiny sizes.
usage (GPU box): compute-sanitizer --tool {memcheck,racecheck,synccheck} python tools/sanitize.py [part ...]
parts: match pnp chain ba init (default: all), pcg (the CG solver alone, several block rows per CTA)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import pipeline, synth

parts = sys.argv[1:] or ["match", "pnp", "chain", "ba", "init"]
ctx = sfm.Context(0)
scene = synth.orbit_scene(5, 700, seed=9)
K = scene["K"]
v0, v1 = scene["views"][0], scene["views"][1]
if "match" in parts:       # K1b, K1 (cluster pair, TMEM double buffering), K1c, gather; and the float32 fallback
    idx, dist, good, ng = ctx.knn2(v0["des"], v1["des"], 0.70)
    idx2, *_ = ctx.knn2(v0["des"] + 0.25, v1["des"], 0.70)
    print("match ok", ng, idx.shape, idx2.shape)
if "pnp" in parts:         # minimal solver (Jacobi wavefront), scoring, replay + inliers + LM
    rng = np.random.default_rng(0)
    n = 300
    X = np.c_[rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(5, 11, n)].astype(np.float32)
    R, t = synth.orbit_pose(0.2)
    uv, _ = synth.project(K, R, t, X.astype(np.float64))
    p = (uv + rng.normal(0, 0.5, uv.shape)).astype(np.float32)
    ok, rvec, tvec, inl, info = ctx.pnp_ransac(X, p, K)
    print("pnp ok", ok, len(inl))
    ok4, r4, t4, inl4, _ = ctx.pnp_ransac(X[:4], p[:4], K)            # npoints == 4: the P3P kernel
    print("p3p ok", ok4, None if inl4 is None else len(inl4))
if "chain" in parts:       # the registration loop: three streams, clustered LM, hash association
    outs = pipeline.register_chain(scene, ctx=ctx)
    print("chain ok", len(outs), [o["n_inl"] for o in outs])
if "ba" in parts:          # K5, K6, tile Cholesky graph, back substitution, update
    pb = synth.ba_problem(40, 1500, 5, seed=1)
    prob = sfm.BAProblem(ctx, 40, 1500, pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
    prob.set_params(pb["cams0"], pb["pts0"])
    out = prob.eval(0)
    hist = prob.solve(max_iters=3)
    print("ba ok", hist[0]["cost_before"], hist[-1]["cost_after"])
    B = np.random.default_rng(2).normal(size=(132, 140))
    x, info = ctx.reduced_solve(B @ B.T / 132 + 0.5 * np.eye(132), np.ones(132))       # three tile columns
    print("solve ok", info)
    x2, solved, its = ctx.reduced_solve(B @ B.T / 132 + 0.5 * np.eye(132), np.ones(132), method="pcg")     # CG kernel: grid barrier
    print("pcg ok", solved, its, float(np.abs(x - x2).max()))
    prob.close()
if "pcg" in parts:         # CG kernel at 300 cameras: two or three block rows per CTA, ranges summed in shared memory
    rng = np.random.default_rng(4)
    B = rng.normal(size=(1800, 1808))
    S = (B @ B.T / 1800 + 0.5 * np.eye(1800)).astype(np.float32)
    g = rng.normal(size=1800).astype(np.float32)
    x, solved, its = ctx.reduced_solve(S, g, method="pcg")
    ref = np.linalg.solve(np.tril(S).astype(np.float64) + np.tril(S, -1).T.astype(np.float64), -g.astype(np.float64))
    print("pcg ok", solved, its, float(np.abs(x - ref).max() / np.abs(ref).max()))
if "init" in parts:        # five-point RANSAC + recoverPose
    tv0, tv1, _, _ = synth.two_view_pair(300, seed=3)
    init = pipeline.two_view_init(tv0, tv1, K, ctx=ctx)
    print("init ok", init["n_essential"], init["n_pose"])
