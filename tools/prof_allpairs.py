"""Host-side profile of the pair-sharded all-pairs matching step (bench.py match_sweep (b)): cProfile of three steps."""
import cProfile, pstats, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sfm_mvs_b200 as sfm
from sfm_mvs_b200 import pipeline, sharding, synth
V = int(sys.argv[1]) if len(sys.argv) > 1 else 64
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = sfm.Context(0)
scene = synth.orbit_scene(V, 5000, seed=0)
dev = ctx.torch_device
with torch.cuda.stream(ctx.torch_stream()):
    kps = [torch.from_numpy(v["kp"]).to(dev) for v in scene["views"]]
    dess = [torch.from_numpy(v["des"]).to(dev) for v in scene["views"]]
    pairs = sharding.all_pairs(V)
    def step():
        views = pipeline.DeviceView.batch(ctx, kps, dess)
        return pipeline.match_pairs_sharded(ctx, views, pairs, 0, world)
    for _ in range(2): step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(3): step()
    torch.cuda.synchronize()
    print("ms per step", (time.perf_counter() - t0) / 3 * 1e3)
    ctx.set_profiling(True); ctx.reset_profile(); step(); print({k: round(v["ms"], 3) for k, v in ctx.profile().items()}); ctx.set_profiling(False)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(3): step()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
