"""Profiling driver: K1 match kernel on resident descriptor sets (ncu target).
usage: prof_match.py [n:pairs ...]   e.g. prof_match.py 5000:8 16384:2"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import synth
from sfm_mvs_b200._lib import lib, check
specs = [a for a in sys.argv[1:] if ":" in a] or ["5000:8"]
ctx = sfm.Context(0)
ts = ctx.torch_stream()
for spec in specs:
    n, pairs = (int(x) for x in spec.split(":"))
    sets = [ctx.descriptors(synth.sift_like_descriptors(n, seed=s)) for s in range(pairs + 1)]
    with torch.cuda.stream(ts):
        idx = torch.empty((n, 2), dtype=torch.int32, device="cuda"); good = torch.empty((n,), dtype=torch.uint8, device="cuda")
    def run():
        for k in range(pairs):
            check(lib.sfm_desc_match(ctx._h, sets[k]._h, sets[k + 1]._h, 0.7, idx.data_ptr(), None, good.data_ptr(), None, 0))
    for _ in range(3): run()
    ctx.sync()
    ctx.set_profiling(True); ctx.reset_profile()
    for _ in range(5): run()
    p = ctx.profile()
    ctx.set_profiling(False)
    t = p["match_tc"]["ms"] / p["match_tc"]["launches"] * 1e-3
    print("n", n, "match_tc us/launch %.2f" % (t * 1e6), "TFLOP/s %.1f" % (2.0 * n * n * 128 / t / 1e12),
          "| match_final us %.2f" % (1e3 * p["match_final"]["ms"] / p["match_final"]["launches"]))
    # batched: ONE K1 launch over the items of all pairs (+ one K1c grid)
    import ctypes as C
    hq = np.array([sets[k]._h.value for k in range(pairs)], np.uint64); ht = np.array([sets[k + 1]._h.value for k in range(pairs)], np.uint64)
    with torch.cuda.stream(ts):
        bidx = [torch.empty((n, 2), dtype=torch.int32, device="cuda") for _ in range(pairs)]
        bgood = [torch.empty((n,), dtype=torch.uint8, device="cuda") for _ in range(pairs)]
    pi = np.array([x.data_ptr() for x in bidx], np.uint64); pg = np.array([x.data_ptr() for x in bgood], np.uint64)
    def runb():
        check(lib.sfm_desc_match_batched(ctx._h, pairs, hq.ctypes.data, ht.ctypes.data, 0.7, pi.ctypes.data, None, pg.ctypes.data, None))
    for _ in range(3): runb()
    ctx.sync()
    ctx.set_profiling(True); ctx.reset_profile()
    for _ in range(5): runb()
    p = ctx.profile()
    ctx.set_profiling(False)
    t = p["match_tc"]["ms"] / p["match_tc"]["launches"] * 1e-3
    print("n", n, "BATCHED x%d: match_tc us/launch %.2f = %.2f us/pair" % (pairs, t * 1e6, t * 1e6 / pairs),
          "TFLOP/s %.1f" % (2.0 * n * n * 128 * pairs / t / 1e12),
          "| match_final us/launch %.2f" % (1e3 * p["match_final"]["ms"] / p["match_final"]["launches"]))
    del sets
