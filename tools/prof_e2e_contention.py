"""Does host->device traffic by itself slow the registration loop?  register_device on resident inputs, alone and
with the 520 MB of a step's uploads running on another stream at the same time (into buffers nobody reads).
  python tools/prof_e2e_contention.py [views] [descriptors]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import pipeline, synth
V = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
ctx = sfm.Context(0)
ts = ctx.torch_stream()
scene = synth.orbit_scene(V, n, seed=0)
K = scene["K"]
Rt0 = np.hstack([scene["views"][0]["R"], scene["views"][0]["t"]]); Rt1 = np.hstack([scene["views"][1]["R"], scene["views"][1]["t"]])
des_host = [torch.from_numpy(v["des"]).pin_memory() for v in scene["views"]]
with torch.cuda.stream(ts):
    kp_dev = [torch.from_numpy(v["kp"]).to("cuda") for v in scene["views"]]
    des_dev = [d.to("cuda", non_blocking=True) for d in des_host]
ctx.sync()
big = torch.cat(des_host).pin_memory()
sink = torch.empty_like(big, device="cuda")
cs = torch.cuda.Stream()
def now():
    ctx.sync(); torch.cuda.synchronize(); return time.perf_counter()
for rep in range(5):
    t0 = now()
    pipeline.register_device(ctx, K, kp_dev, des_dev, Rt0, Rt1)
    t1 = now()
    with torch.cuda.stream(cs):
        sink.copy_(big, non_blocking=True)
    pipeline.register_device(ctx, K, kp_dev, des_dev, Rt0, Rt1)
    t2 = now()
    print(f"rep {rep}: register_device alone {1e3 * (t1 - t0):.2f} ms | with a {big.numel() * 4 / 1e6:.0f} MB upload in flight {1e3 * (t2 - t1):.2f} ms")
