"""LM iterations on BASELINE configs[3] (500 / 100k / 1M) from the perturbed start under the linear-solver tolerance in
SFM_BA_CG_TOL (one process per value: the library reads it once): cost trajectory, acceptance, time of 10 iterations.
  SFM_BA_CG_TOL=1e-8 python tools/ba_tol_probe.py; python tools/ba_tol_probe.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import synth

ctx = sfm.Context(0)
pb = synth.ba_problem(500, 100000, 10, seed=0)
prob = sfm.BAProblem(ctx, 500, 100000, pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
out = {"tol": os.environ.get("SFM_BA_CG_TOL", "default")}
for rep in range(2):                                   # the second pass is the timed one
    prob.set_params(pb["cams0"], pb["pts0"])
    ctx.sync()
    lam, costs, acc = 1e-3, [], []
    t0 = time.perf_counter()
    for _ in range(10):
        st = prob.gn_step(lam)
        lam = st["lambda_next"]
        costs.append(round(st["cost_after"], 4)); acc.append(bool(st["accepted"]))
    ctx.sync()
    out["ms_per_iter"] = (time.perf_counter() - t0) * 100.0
out["costs"], out["accepted"] = costs, acc
out["final_cost"] = prob.eval(0, want_r=False, want_J=False)["cost"]
print(json.dumps(out))
