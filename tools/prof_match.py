"""Profiling driver: K1 match kernel on resident descriptor sets (ncu target).  usage: prof_match.py [n] [pairs]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ctx = sfm.Context(0)
sets = [ctx.descriptors(synth.sift_like_descriptors(n, seed=s)) for s in range(pairs + 1)]
ts = ctx.torch_stream()
with torch.cuda.stream(ts):
    idx = torch.empty((n, 2), dtype=torch.int32, device="cuda"); good = torch.empty((n,), dtype=torch.uint8, device="cuda")
from sfm_mvs_b200._lib import lib, check
def run():
    for k in range(pairs):
        check(lib.sfm_desc_match(ctx._h, sets[k]._h, sets[k + 1]._h, 0.7, idx.data_ptr(), None, good.data_ptr(), None, 0))
for _ in range(3): run()
ctx.sync()
ctx.set_profiling(True); ctx.reset_profile()
for _ in range(5): run()
p = ctx.profile()
for k, v in p.items(): print(k, "us/launch %.2f" % (1e3 * v["ms"] / v["launches"]), "launches", v["launches"])
t = p["match_tc"]["ms"] / p["match_tc"]["launches"] * 1e-3
print("n", n, "match_tc TFLOP/s %.1f" % (2.0 * n * n * 128 / t / 1e12))
