"""Wall-clock split of one end-to-end registration step (pinned host inputs): upload+prep, match, chain, fetch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import pipeline, synth
V, n = int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 5000
ctx = sfm.Context(0); ts = ctx.torch_stream()
scene = synth.orbit_scene(V, n, seed=0)
K = scene["K"]
Rt0 = np.hstack([scene["views"][0]["R"], scene["views"][0]["t"]]); Rt1 = np.hstack([scene["views"][1]["R"], scene["views"][1]["t"]])
kp_host = [torch.from_numpy(v["kp"]).pin_memory() for v in scene["views"]]
des_host = [torch.from_numpy(v["des"]).pin_memory() for v in scene["views"]]
def T():
    ctx.sync(); torch.cuda.synchronize(); return time.perf_counter()
for rep in range(4):
    t0 = T()
    views = [pipeline.DeviceView(ctx, k, d) for k, d in zip(kp_host, des_host)]
    t1 = T()
    chain = pipeline.RegistrationChain(ctx, K)
    matches = chain.match_pairs(views, [(i, i + 1) for i in range(V - 1)])
    t2 = T()
    outs = chain.run(views, Rt0, Rt1, matches=matches)
    t3 = T()
    with torch.cuda.stream(ts):
        clouds = [o["X_new"][:o["n_new"]].to("cpu", non_blocking=True) for o in outs]
    t4 = T()
    print(f"rep {rep}: upload+prep {1e3*(t1-t0):.1f} ms | match {1e3*(t2-t1):.1f} | chain {1e3*(t3-t2):.1f} | fetch {1e3*(t4-t3):.1f} | total {1e3*(t4-t0):.1f}")
