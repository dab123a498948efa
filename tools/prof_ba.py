"""Profiling driver: BA evaluation (K5) and one LM iteration on the 500 / 100k / 1M problem (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
for _ in range(4): prob.eval_into(0, r, Jc, Jp, cost)
ctx.sync()
lam = 1e-3
for _ in range(2): lam = prob.gn_step(lam)["lambda_next"]
ctx.set_profiling(True); ctx.reset_profile()
st = prob.gn_step(lam)
print(st)
for k, v in ctx.profile().items(): print(k, "ms %.3f" % v["ms"], "launches", v["launches"])
