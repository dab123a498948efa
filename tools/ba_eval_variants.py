"""Diagnostic: time the K5 evaluation kernel for SFM_BA_EVAL_VARIANT in {0,1,2,3} (one subprocess each)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r)
import numpy as np, torch
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import synth
ctx = sfm.Context(0)
pb = synth.ba_problem(500, 100000, 10, seed=0)
prob = sfm.BAProblem(ctx, 500, 100000, pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
prob.set_params(pb["cams0"], pb["pts0"])
O = prob.n_obs
ts = ctx.torch_stream()
with torch.cuda.stream(ts):
    r = torch.empty((O, 2), dtype=torch.float32, device="cuda"); Jc = torch.empty((O, 2, 6), dtype=torch.float32, device="cuda")
    Jp = torch.empty((O, 2, 3), dtype=torch.float32, device="cuda"); cost = torch.zeros((1,), dtype=torch.float64, device="cuda")
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
for _ in range(3): prob.eval_into(0, r, Jc, Jp, cost)
ctx.sync(); ctx.set_profiling(True); ctx.reset_profile()
for _ in range(10):
    with torch.cuda.stream(ts): flush.zero_()
    prob.eval_into(0, r, Jc, Jp, cost)
p = ctx.profile()["ba_eval"]; t = p["ms"] / p["launches"] * 1e-3
print("variant", os.environ.get("SFM_BA_EVAL_VARIANT"), "us %%.2f  GB/s %%.0f" %% (t * 1e6, 96.0 * O / t / 1e9))
''' % ROOT
for v in ("0", "1", "2", "3"):
    env = dict(os.environ, SFM_BA_EVAL_VARIANT=v)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(out.stdout.strip() or out.stderr[-500:])
