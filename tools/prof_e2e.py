"""Wall-clock split of one end-to-end registration step (pinned host inputs) against the resident driver.

  python tools/prof_e2e.py [views] [descriptors]

Prints, per repetition: the host time spent only SUBMITTING the host->device copies of all views (what
pipeline.register_host does before its chunk loop starts), register_host + fetch_clouds (the e2e figure of bench.py),
and register_device on resident copies (the `value` figure).  SFM_REGISTER_EDGES=a,b,... overrides the resident
driver's chunk boundaries."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import pipeline, synth

V = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
ctx = sfm.Context(0)
ts = ctx.torch_stream()
scene = synth.orbit_scene(V, n, seed=0)
K = scene["K"]
Rt0 = np.hstack([scene["views"][0]["R"], scene["views"][0]["t"]])
Rt1 = np.hstack([scene["views"][1]["R"], scene["views"][1]["t"]])
kp_host = [torch.from_numpy(v["kp"]).pin_memory() for v in scene["views"]]
des_host = [torch.from_numpy(v["des"]).pin_memory() for v in scene["views"]]
with torch.cuda.stream(ts):
    kp_dev = [k.to("cuda", non_blocking=True) for k in kp_host]
    des_dev = [d.to("cuda", non_blocking=True) for d in des_host]
ctx.sync()


def now():
    ctx.sync()
    torch.cuda.synchronize()
    return time.perf_counter()


cs = torch.cuda.Stream()
for rep in range(5):
    t0 = now()
    with torch.cuda.stream(cs):
        tmp = [(k.to("cuda", non_blocking=True), d.to("cuda", non_blocking=True)) for k, d in zip(kp_host, des_host)]
    t_submit = time.perf_counter() - t0          # host time to queue the copies (not their duration)
    t1 = now()
    del tmp
    outs = pipeline.register_host(ctx, K, kp_host, des_host, Rt0, Rt1)
    clouds = pipeline.fetch_clouds(ctx, outs)
    t2 = now()
    res = pipeline.register_device(ctx, K, kp_dev, des_dev, Rt0, Rt1)
    t3 = now()
    o2 = pipeline.register_host(ctx, K, kp_host, des_host, Rt0, Rt1, ahead=10 ** 6)      # every copy submitted up front
    pipeline.fetch_clouds(ctx, o2)
    t_il = now() - t3
    assert [o["n_match"] for o in o2] == [o["n_match"] for o in outs]
    print(f"rep {rep}: copy submission {1e3 * t_submit:.2f} ms (copies done after {1e3 * (t1 - t0):.2f}) | "
          f"register_host + fetch {1e3 * (t2 - t1):.2f} ms | register_device {1e3 * (t3 - t2):.2f} ms | "
          f"all copies submitted up front {1e3 * t_il:.2f} ms | {len(outs)} views")
