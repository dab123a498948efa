"""register_host (chunked upload + registration) wall time for several chunk sizes / loop variants."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import pipeline, synth
V, n = 200, 5000
ctx = sfm.Context(0)
scene = synth.orbit_scene(V, n, seed=0)
K = scene["K"]
Rt0 = np.hstack([scene["views"][0]["R"], scene["views"][0]["t"]]); Rt1 = np.hstack([scene["views"][1]["R"], scene["views"][1]["t"]])
kp = [torch.from_numpy(v["kp"]).pin_memory() for v in scene["views"]]
des = [torch.from_numpy(v["des"]).pin_memory() for v in scene["views"]]
def run(chunk, sync):
    if sync: os.environ["SFM_CHAIN_SYNC"] = "1"
    else: os.environ.pop("SFM_CHAIN_SYNC", None)
    ts = []
    for rep in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        outs = pipeline.register_host(ctx, K, kp, des, Rt0, Rt1, chunk=chunk)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        del outs
    print(f"chunk {chunk:3d} {'sync-loop' if sync else 'sync-free'}: " + " ".join(f"{1e3*t:.1f}" for t in ts) + " ms")
for chunk in (200, 50, 25, 10):
    for sync in (True, False):
        run(chunk, sync)
