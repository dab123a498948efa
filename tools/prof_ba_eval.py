"""Profiling driver: K5 (ba_eval_kernel) on the 500 / 100k / 1M problem, timed exactly as bench.py does it
(256 MB L2 flush between launches) and, for comparison, back to back without a flush.
Under ncu use:  --cache-control none -k regex:ba_eval -c 1 --launch-skip 8"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import synth
ctx = sfm.Context(0)
NPTS = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
pb = synth.ba_problem(500, NPTS, 10, seed=0)
prob = sfm.BAProblem(ctx, 500, NPTS, pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
prob.set_params(pb["cams0"], pb["pts0"])
O = prob.n_obs
ts = ctx.torch_stream()
with torch.cuda.stream(ts):
    r = torch.empty((O, 2), dtype=torch.float32, device="cuda"); Jc = torch.empty((O, 2, 6), dtype=torch.float32, device="cuda")
    Jp = torch.empty((O, 2, 3), dtype=torch.float32, device="cuda"); cost = torch.zeros((1,), dtype=torch.float64, device="cuda")
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
for _ in range(4): prob.eval_into(0, r, Jc, Jp, cost)
ctx.sync()
for mode in ("flush", "noflush"):
    ctx.set_profiling(True); ctx.reset_profile()
    for _ in range(10):
        if mode == "flush":
            with torch.cuda.stream(ts):
                flush.zero_()
        prob.eval_into(0, r, Jc, Jp, cost)
    p = ctx.profile()["ba_eval"]
    ctx.set_profiling(False)
    t = p["ms"] * 1e-3 / p["launches"]
    print(f"{mode}: ba_eval {t * 1e6:.2f} us/launch, {96.0 * O / t / 1e9:.0f} GB/s (96 B/obs, {O} obs)")
