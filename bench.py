#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: views registered/sec (match + PnP + triangulate) on the
synthetic 200-view x 5000-descriptor set (configs[2], the set north_star quotes its target on), plus
BA Gauss-Newton iterations/sec on the 500-camera / 100k-point / 1M-observation problem (configs[3]).

    python bench.py --gpus N --steps K --warmup W            # this engine (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

A step = one pass of the per-view registration loop (sfm.py:341-409, imread/SIFT/GUI removed) over the
whole scene: V-2 views registered.  N > 1: every rank registers its own scene (the chain is sequential —
"replicas only"), weak scaling, no data-path collective; the BA section shards points
over ranks and all-reduces the reduced camera system over NCCL.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def log(msg):
    """Progress on stderr (stdout carries the one JSON line)."""
    if os.environ.get("RANK", "0") == "0":
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--views", type=int, default=200)
    ap.add_argument("--desc", type=int, default=5000)
    ap.add_argument("--cpu-views", type=int, default=26, help="view prefix the CPU baseline registers")
    ap.add_argument("--no-ba", action="store_true")
    ap.add_argument("--ba-cams", type=int, default=500)
    ap.add_argument("--ba-points", type=int, default=100_000)
    ap.add_argument("--ba-obs-per-point", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-match-sweep", action="store_true")
    ap.add_argument("--sweep-views", type=int, default=64, help="views of the all-pairs matching workload (isfm.py:68-87)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]), bf16_sustained=float(p["bf16_tflops_sustained"]),
                    source="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.summary = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if not self.proc:
            return
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            self.summary = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                                samples=len(sm))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# ====================================================================================== reference arm
WORKLOAD = "synthetic {V} views x {n} desc: full incremental register (PnP+triangulate), BASELINE configs[2]"


def cpu_register(scene, n_views, keep=False):
    """The reference's per-view loop on the host cores (oracle port of sfm.py:341-409, the same cv2 calls;
    OpenCV threads = all cores, Python driver single-threaded as in the reference)."""
    from oracle import cvpath
    t0 = time.perf_counter()
    outs = quiet(cvpath.register_chain, scene, n_views)
    dt = time.perf_counter() - t0
    return (len(outs), dt, outs) if keep else (len(outs), dt)


def parity_check(ctx, K, engine_outs, cpu_outs):
    """The timed engine path against the CPU arm's outputs for the same views (the oracle as CHECKER, not timed):
    (1) per call — the reference's own solvePnPRansac inputs of every view through the engine's default
    solvePnPRansac: inlier lists must be identical (bit-exact masks), poses within 1e-4;
    (2) the chains — counts, poses, errors and new points of the engine's loop beside the CPU loop's.  The two
    loops cannot stay bit-identical: cv2 refines each pose with LM whose normal equations go through OpenBLAS
    (summation order not reproducible), the ~1e-9 pose difference moves some float32 3-D points of the next view by
    an ulp, and a 5-point hypothesis changes completely with its input bits (tests/test_oracle.py)."""
    import cv2
    D0 = np.zeros((5, 1), np.float32)
    n = min(len(engine_outs), len(cpu_outs))
    same_mask = pose_ok = 0
    max_call_dpose = 0.0
    for r in cpu_outs[:n]:
        ok_ref, rv_ref, tv_ref, inl_ref = cv2.solvePnPRansac(r["pnp_X"], r["pnp_p"], K, D0, cv2.SOLVEPNP_ITERATIVE)
        ok, rv, tv, inl, _ = ctx.pnp_ransac(r["pnp_X"], r["pnp_p"], K)
        if ok == ok_ref and (not ok or np.array_equal(inl, inl_ref[:, 0])):
            same_mask += 1
        if ok and ok_ref:
            d = max(np.abs(rv - rv_ref.ravel()).max(), np.abs(tv - tv_ref.ravel()).max())
            max_call_dpose = max(max_call_dpose, float(d))
            pose_ok += d <= 1e-4
    cnt = sum((e["n_match"], e["n_pnp"]) == (r["n_match"], r["n_pnp"]) for e, r in zip(engine_outs, cpu_outs))
    inl = sum(e["n_inl"] == r["n_inl"] for e, r in zip(engine_outs, cpu_outs))
    new = sum(e["n_new"] == len(r["X_new"]) for e, r in zip(engine_outs, cpu_outs))
    dRt = max(float(np.abs(e["Rt"] - r["Rt"]).max()) for e, r in zip(engine_outs, cpu_outs))
    derr = max(abs(e["err_new"] - r["err_new"]) / r["err_new"] for e, r in zip(engine_outs, cpu_outs))
    dX = 0.0
    for e, r in zip(engine_outs, cpu_outs):
        if e["n_new"] == len(r["X_new"]) and e["n_new"]:
            x = e["X_new"][:e["n_new"]]
            x = x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)
            dX = max(dX, float(np.abs(x - r["X_new"]).max() / np.abs(r["X_new"]).max()))
    return {"views": n,
            "per_call_same_inputs": {"inlier_mask_identical": same_mask, "pose_within_1e-4": int(pose_ok),
                                     "max_abs_pose_diff": max_call_dpose, "of": n},
            "chain": {"n_match_and_n_pnp_equal": cnt, "n_inliers_equal": inl, "n_new_equal": new, "max_abs_dRt": dRt,
                      "max_rel_derr_new": float(derr), "max_rel_dX_new": dX, "of": n},
            "bars": "masks bit-exact per call; 3-D points and errors within 1e-4 relative (north_star)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cv2
    from sfm_mvs_b200 import synth
    nv = min(args.cpu_views, args.views)
    scene = synth.orbit_scene(nv, args.desc, seed=0)       # sfm_mvs_b200.synth is host-only: the engine is never loaded here
    assert "sfm_mvs_b200._lib" not in sys.modules
    for _ in range(max(args.warmup, 1)):
        cpu_register(scene, min(4, nv))
    t_tot, reg = 0.0, 0
    for _ in range(args.steps):
        r, t = cpu_register(scene, nv)
        reg += r; t_tot += t
    v = reg / t_tot
    sample = f"{nv}-view prefix of the {args.views}x{args.desc} scene ({nv - 2} views registered per step)"
    print(json.dumps({
        "impl": "reference", "metric": "views registered/sec (match+PnP+tri)", "value": v, "unit": "views/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64 (OpenCV)",
        "data": "synthetic", "config": {"workload": WORKLOAD.format(V=args.views, n=args.desc)},
        "cpu_baseline": {"value": v, "unit": "views/s", "cores": cv2.getNumThreads(), "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "cv2": cv2.__version__},
        "e2e": {"value": v, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ====================================================================================== engine arm
def run_engine(args):
    import torch
    import torch.distributed as dist

    import sfm_mvs_b200 as sfm
    from sfm_mvs_b200 import pipeline, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = sfm.Context(local)
    ts = ctx.torch_stream()
    pk = peaks()
    V, n = args.views, args.desc
    scene = synth.orbit_scene(V, n, seed=rank)                 # every rank its own scene (replicas)
    K = scene["K"]
    Rt0 = np.hstack([scene["views"][0]["R"], scene["views"][0]["t"]])
    Rt1 = np.hstack([scene["views"][1]["R"], scene["views"][1]["t"]])
    # host inputs in pinned memory (e2e) and resident copies in HBM (value)
    kp_host = [torch.from_numpy(v["kp"]).pin_memory() for v in scene["views"]]
    des_host = [torch.from_numpy(v["des"]).pin_memory() for v in scene["views"]]
    with torch.cuda.stream(ts):
        kp_dev = [k.to("cuda", non_blocking=True) for k in kp_host]
        des_dev = [d.to("cuda", non_blocking=True) for d in des_host]
    ctx.sync()
    h2d_bytes = sum(k.numel() * 4 + d.numel() * 4 for k, d in zip(kp_host, des_host))
    input_bytes = h2d_bytes

    def step(kps, dess, fetch: bool):
        if fetch:
            # end to end through the public host-array entry: chunked upload on a copy stream overlapping
            # descriptor prep + batched match + registration (pipeline.register_host), clouds read back
            outs = pipeline.register_host(ctx, K, kps, dess, Rt0, Rt1)
            clouds = pipeline.fetch_clouds(ctx, outs)
            return outs, sum(c.size * 4 for c in clouds) + len(outs) * (16 + 96)
        # resident inputs: K1b descriptor prep + batched match on the matching context, chunk by chunk, while the
        # registration loop of the previous chunk runs on the main one (pipeline.register_device)
        outs = pipeline.register_device(ctx, K, kps, dess, Rt0, Rt1)
        return outs, 0

    registered = V - 2

    def timed(kps, dess, fetch, steps):
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        w0 = time.perf_counter()
        e0.record(ts)
        d2h = 0
        for _ in range(steps):
            _, d2h = step(kps, dess, fetch)
        e1.record(ts)
        torch.cuda.synchronize()
        barrier()
        wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        return max_over_ranks(ms), wall, ctx.launch_count() - l0, d2h

    log(f"scene ready: {V} views x {n} descriptors, world {world}")
    for _ in range(args.warmup):
        outs, _ = step(kp_dev, des_dev, False)
    with ClockSampler(local) as cs:
        ms, wall, launches, _ = timed(kp_dev, des_dev, False, args.steps)
    value = world * registered * args.steps / (ms * 1e-3)
    # end to end: pinned host buffers in, clouds + poses out, through the public API
    for _ in range(args.warmup):
        step(kp_host, des_host, True)
    ms_e2e, _, _, d2h_bytes = timed(kp_host, des_host, True, args.steps)
    e2e = world * registered * args.steps / (ms_e2e * 1e-3)

    log(f"registration timed: {value:.0f} views/s resident, {e2e:.0f} end to end")
    # ---- per-kernel device time (CUDA events around every launch, same K steps) and the roofline
    ctx.set_profiling(True)
    ctx.reset_profile()
    for _ in range(args.steps):
        step(kp_dev, des_dev, False)
    prof = ctx.profile()
    ctx.set_profiling(False)
    tot_ms = sum(p["ms"] for p in prof.values()) or 1.0
    shares = {k: round(p["ms"] / tot_ms, 4) for k, p in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    per_launch_us = {k: round(1e3 * p["ms"] / p["launches"], 3) for k, p in prof.items()}
    mt = prof.get("match_tc")
    roofline = None
    if mt:
        # 2*Nq*Nt*128 flops per pair; the V-1 consecutive pairs of the scene are matched by a few batched K1
        # launches per step (one per chunk of pipeline.register_device: 11 + 48 + 140 pairs at V = 200)
        lps = mt["launches"] / args.steps                     # K1 launches per step
        flops = 2.0 * n * n * 128 * (V - 1) / lps             # average algorithmic flops per launch
        t_launch = mt["ms"] * 1e-3 / mt["launches"]
        ach = flops / t_launch / 1e12
        roofline = {"kernel": "match_tc_kernel (K1 tcgen05 distance GEMM + top-2 epilogue)", "bound": "tensor",
                    "achieved": ach, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": ach / pk["bf16"],
                    "frac_of_sustained_peak": ach / pk["bf16_sustained"],
                    # dram__bytes_read.sum + dram__bytes_write.sum of this very launch (199 pairs x 5000 descriptors) from
                    # one ncu --set full capture: profiles/r1v_match_tc_bench_199x5k.txt (327.9 + 84.1 MB); other sizes: null
                    "traffic": 412.06e6 / lps if (V == 200 and n == 5000) else None,
                    "traffic_unit": "bytes per launch, averaged like `achieved`: the ncu capture holds the scene's 199 pairs in one "
                                    "launch (profiles/r1v_match_tc_bench_199x5k.txt, 2.07 MB per pair); divided by the launches per step",
                    "peak_source": pk["source"] + ", burst bf16: K1 is ~1 ms of a ~40 ms step at full clocks, not a sustained load",
                    "algorithmic_flops_per_launch": flops, "pairs_per_launch": (V - 1) / lps, "launches_per_step": lps,
                    "avg_launch_us": 1e6 * t_launch,
                    "launches": mt["launches"],
                    "note": "K1 is the path's dense contraction (the kernel north_star sets a tensor-pipe target for). By time the "
                            "registration loop is dominated by the latency-bound PnP kernels (one warp per EPnP hypothesis, a "
                            "single-CTA LM refinement): see kernel_time_share / DESIGN.md.",
                    "share_of_kernel_time": shares.get("match_tc")}

    out = {
        "metric": "views registered/sec (match+PnP+tri)", "value": value, "unit": "views/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16 tensor-core distances (exact on integer SIFT) + f64 geometry, f32 I/O",
        "data": "synthetic",
        "config": {"workload": WORKLOAD.format(V=V, n=n),
                   "views_registered_per_step": registered, "parallelism": "replicas (one scene per GPU)" if world > 1 else "1 GPU",
                   "l2_policy": f"inputs larger than L2 ({input_bytes / 1e6:.0f} MB of descriptors+keypoints per step vs 126 MB L2)",
                   "pnp_minimal_solver": "EPnP on the GPU in OpenCV's exact arithmetic: hypotheses bit-identical to cv2's "
                                         "(csrc/pnp_epnp.cu) - the parity configuration IS the timed configuration"},
        "e2e": {"value": e2e, "unit": "views/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "wall_s": wall, "clocks": cs.summary,
        "kernel_time_share": shares, "kernel_us_per_launch": per_launch_us, "roofline": roofline,
    }

    # ---- two-view initialisation (sfm.py:307-316: findEssentialMat + recoverPose), once per scene in the reference;
    # public API with host arrays, so copies are inside the timed calls.  Reported beside the headline, not in it.
    if world == 1:
        from sfm_mvs_b200 import pipeline as _pl
        tv0, tv1, _, _ = synth.two_view_pair(int(0.6 * n), seed=7, outliers=0.3)
        for _ in range(3):
            init = _pl.two_view_init(tv0, tv1, scene["K"], ctx=ctx)
        t0 = time.perf_counter()
        for _ in range(10):
            init = _pl.two_view_init(tv0, tv1, scene["K"], ctx=ctx)
        out["two_view_init"] = {"engine_ms": (time.perf_counter() - t0) / 10 * 1e3, "correspondences": len(tv0),
                                "survivors": int(init["n_pose"]), "cpu_ms": None,
                                "what": "findEssentialMat(RANSAC, 0.999, 0.4) + recoverPose, host arrays in / out"}

    # ---- descriptor-match sweep (configs[4]) and pair-sharded all-pairs matching (isfm.py:68-87) over the ranks
    if not args.no_match_sweep:
        out["match_sweep"] = bench_match(args, ctx, world, rank, pk, barrier, max_over_ranks)
        log("match sweep done")

    # ---- BA: GN iterations / s (configs[3]); points sharded over ranks, NCCL all-reduce of the reduced system
    if not args.no_ba:
        out["ba"] = bench_ba(args, ctx, world, rank, pk, barrier, max_over_ranks)

    # ---- CPU baseline (rank 0, N = 1 only): the reference path on a bounded sample of the same scene
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import cv2
        nv = min(args.cpu_views, V)
        cpu_register(scene, 4)
        r, t, cpu_outs = cpu_register(scene, nv, keep=True)
        out["cpu_baseline"] = {"value": r / t, "unit": "views/s", "cores": cv2.getNumThreads(), "kind": "port",
                               "sample": f"first {nv} views of the same scene ({r} views registered, {t:.1f} s)",
                               "host_cpus": os.cpu_count(), "cv2": cv2.__version__}
        out["parity_check"] = parity_check(ctx, K, outs, cpu_outs)
        if "two_view_init" in out:           # the same two lines of the reference on cv2 (sfm.py:307, :311)
            t0 = time.perf_counter()
            for _ in range(3):
                E, m = cv2.findEssentialMat(tv0, tv1, scene["K"], method=cv2.RANSAC, prob=0.999, threshold=0.4, mask=None)
                cv2.recoverPose(E, tv0[m.ravel() == 1], tv1[m.ravel() == 1], scene["K"])
            out["two_view_init"]["cpu_ms"] = (time.perf_counter() - t0) / 3 * 1e3
    elif rank == 0:
        out["cpu_baseline"] = None
    log("done")
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_match(args, ctx, world, rank, pk, barrier, max_over_ranks):
    """BASELINE configs[4]: (a) one pair of n x n 128-D descriptors, n = 1k .. 64k, on one GPU: time of the public
    match call (K1 + K1c, outputs left in HBM), useful TFLOP/s = 2 n^2 128 / t and the operand bytes per second;
    (b) weak scaling over pairs — the all-previous-views pair list of isfm.py:68-87 for `--sweep-views` views of the
    synthetic scene split over the ranks by sharding.shard_pairs, every rank prepares the views and matches its shard
    in one batched launch, counts all-reduced (NCCL), matches stay resident; (c) strong scaling — one 64k x 64k pair
    split by query rows over the ranks.  All times are CUDA-event times on the engine stream, max over ranks."""
    import torch
    import torch.distributed as dist

    from sfm_mvs_b200 import pipeline, sharding, synth

    ts, dev = ctx.torch_stream(), ctx.torch_device
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234)

    def rand_desc(n):            # integer-valued float32 with SIFT's norm (|d| ~ 512), identical on every rank
        return torch.randint(0, 80, (n, 128), device=dev, generator=gen, dtype=torch.int32).to(torch.float32)

    def timed_ms(fn, reps):
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for _ in range(reps):
            fn()
        e1.record(ts)
        torch.cuda.synchronize()
        return max_over_ranks(e0.elapsed_time(e1)) / reps

    def timed_ms_median(fn, reps):
        """Each repetition timed on its own (device events, max over ranks); the median, so that one slow repetition —
        an allocator growth, a collective's first use — does not decide the figure."""
        vals = []
        for _ in range(reps):
            vals.append(timed_ms(fn, 1))
        return sorted(vals)[len(vals) // 2]

    res = {}
    with torch.cuda.stream(ts):
        # (a) single pair, this GPU
        single = []
        for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
            q, t = rand_desc(n), rand_desc(n)
            dq, dt = ctx.descriptors(q), ctx.descriptors(t)
            assert dq.exact and dt.exact
            call = lambda: ctx.knn2(dq, dt, 0.70, device_out=True)
            for _ in range(3):
                call()
            ms = timed_ms(call, 10 if n <= 16384 else 4)
            tf = 2.0 * n * n * 128 / (ms * 1e-3) / 1e12
            single.append({"n": n, "ms": ms, "tflops": tf, "frac_of_burst_bf16_peak": tf / pk["bf16"],
                           "operand_gbs": 2 * n * 160 * 2 / (ms * 1e-3) / 1e9})
            del dq, dt, q, t
        res["single_pair"] = single
        # (b) all pairs of V views, sharded over the ranks
        V = args.sweep_views
        scene = synth.orbit_scene(V, args.desc, seed=0)
        kps = [torch.from_numpy(v["kp"]).to(dev) for v in scene["views"]]
        dess = [torch.from_numpy(v["des"]).to(dev) for v in scene["views"]]
        pairs = sharding.all_pairs(V)
        state = {}

        def all_pairs_step():
            views = pipeline.DeviceView.batch(ctx, kps, dess)
            state["out"] = pipeline.match_pairs_sharded(ctx, views, pairs, rank, world, dist=dist if world > 1 else None)

        for _ in range(2):
            all_pairs_step()
        ms = timed_ms_median(all_pairs_step, 5)
        mine, matches, counts = state["out"]
        res["all_pairs"] = {"views": V, "desc": args.desc, "pairs": len(pairs), "pairs_this_rank": len(mine), "ms": ms,
                            "timing": "median of 5 repetitions", "pairs_per_s": len(pairs) / (ms * 1e-3), "tflops": 2.0 * args.desc ** 2 * 128 * len(pairs) / (ms * 1e-3) / 1e12,
                            "survivors_total": int(counts.sum().item()), "scaling": "strong (fixed pair list split over the ranks)",
                            "exchange": "per-pair survivor counts all-reduced (int32 x pairs); matches stay on the owning GPU"}
        state.clear()
        del matches
        # (c) one large pair split by query rows
        n = 65536
        q, t = rand_desc(n), rand_desc(n)
        split = lambda: state.__setitem__("r", pipeline.match_rows_split(ctx, q, t, rank, world))
        for _ in range(2):
            split()
        ms = timed_ms_median(split, 5)
        lo, hi = state["r"]["lo"], state["r"]["hi"]
        res["row_split"] = {"n": n, "rows_this_rank": hi - lo, "ms": ms, "tflops": 2.0 * n * n * 128 / (ms * 1e-3) / 1e12,
                            "includes": "descriptor preparation (K1b) of this rank's query rows and of the train set",
                            "scaling": "strong (query rows split over the ranks, no collective)"}
    return res


def bench_ba(args, ctx, world, rank, pk, barrier, max_over_ranks):
    import torch
    import torch.distributed as dist

    import sfm_mvs_b200 as sfm
    from sfm_mvs_b200 import synth

    C_, P_, opp = args.ba_cams, args.ba_points, args.ba_obs_per_point
    pb = synth.ba_problem(C_, P_, opp, seed=0)                      # identical on every rank
    from sfm_mvs_b200 import sharding
    sh = sharding.ba_shard(pb, rank, world)                         # contiguous point ranges, all cameras
    prob = sfm.BAProblem(ctx, C_, len(sh["pts0"]), sh["cam_idx"], sh["pt_idx"], sh["obs"], sh["K"], totals=sh["totals"])
    prob.set_params(sh["cams0"], sh["pts0"])
    if world > 1:
        uid = [sfm.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        prob.comm_init(uid[0], rank, world)
    ts = ctx.torch_stream()
    O = prob.n_obs
    # K5 materialised evaluation: the HBM-roofline kernel (96 B / observation)
    with torch.cuda.stream(ts):
        r = torch.empty((O, 2), dtype=torch.float32, device="cuda")
        Jc = torch.empty((O, 2, 6), dtype=torch.float32, device="cuda")
        Jp = torch.empty((O, 2, 3), dtype=torch.float32, device="cuda")
        cost = torch.zeros((1,), dtype=torch.float64, device="cuda")
        flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        prob.eval_into(0, r, Jc, Jp, cost)
    ctx.sync()
    ctx.set_profiling(True)
    ctx.reset_profile()
    reps = 10
    for _ in range(reps):
        with torch.cuda.stream(ts):
            flush.zero_()                                            # flush L2 between timed launches (256 MB > 126 MB)
        prob.eval_into(0, r, Jc, Jp, cost)
    prof = ctx.profile()
    ctx.set_profiling(False)
    t_eval = prof["ba_eval"]["ms"] * 1e-3 / prof["ba_eval"]["launches"]
    gbs = 96.0 * O / t_eval / 1e9
    # LM iterations from the perturbed start (warm-up iterations first, then the parameters are reset): the timed
    # region is the first `iters` iterations of the actual descent, not iterations at the converged point
    lam = 1e-3
    for _ in range(3):
        lam = prob.gn_step(lam)["lambda_next"]
    prob.set_params(sh["cams0"], sh["pts0"])
    barrier()
    torch.cuda.synchronize()
    iters = 10
    lam = 1e-3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    hist = []
    for _ in range(iters):
        st = prob.gn_step(lam)
        lam = st["lambda_next"]
        hist.append(st)
    e1.record(ts)
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1))
    # ... and on to convergence (untimed if it takes longer than the timed iterations): the first iteration whose step —
    # accepted or not — moves the cost by less than 1e-6 relative
    conv = None
    full = list(hist)
    for k in range(40 + len(hist)):
        if k >= len(full):
            st = prob.gn_step(lam)
            lam = st["lambda_next"]
            full.append(st)
        if abs(full[k]["cost_before"] - full[k]["cost_after"]) <= 1e-6 * full[k]["cost_before"]:
            conv = k + 1
            break
    ctx.set_profiling(True)
    ctx.reset_profile()
    prob.gn_step(lam)
    prof2 = ctx.profile()
    ctx.set_profiling(False)
    res = {"metric": "BA GN-iters/sec", "value": iters / (ms * 1e-3), "unit": "iters/s", "ms_per_iter": ms / iters,
           "config": {"workload": f"BA {C_} cams / {P_} points / {P_ * opp} obs synthetic, LM iterations, BASELINE configs[3]",
                      "sharding": f"points over {world} rank(s); NCCL all-reduce of S|g|diag(Hcc) (lower block triangle, {C_ * (C_ + 1) // 2 * 36 * 4 / 1e6:.0f} MB f32) per iteration",
                      "linear_solver": "block-Jacobi PCG on the reduced camera system (one persistent kernel), relative residual "
                                       + os.environ.get("SFM_BA_CG_TOL", "1e-5") + " per LM step"},
           "cost_first": hist[0]["cost_before"], "cost_last": hist[-1]["cost_after"],
           "cost_trajectory": [round(h["cost_after"], 3) for h in hist],
           "accepted": [bool(h["accepted"]) for h in hist],
           "timed": f"the first {iters} LM iterations from the perturbed start",
           "iterations_to_converge": conv, "cost_converged": min(full[-1]["cost_before"], full[-1]["cost_after"]),
           "iter_kernel_ms": {k: round(v["ms"], 3) for k, v in prof2.items()},
           "roofline_eval": {"kernel": "ba_eval_kernel (K5 residual + Jacobian blocks, materialised)", "bound": "hbm",
                             "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                             # ncu capture under the same flush conditions: profiles/r1d_ba_eval_kernel.txt (5.0 MB read + 74.0 MB written)
                             "traffic": 79.0e6 if (world == 1 and O == 1_000_000) else None,
                             "algorithmic_bytes_per_launch": 96.0 * O, "avg_launch_us": 1e6 * t_eval,
                             "l2_policy": "256 MB flush between timed launches", "peak_source": pk["source"]}}
    prob.close()
    if world == 1:
        res["fixture"] = bench_ba_fixture(ctx)
        if not args.no_cpu_baseline:
            res["cpu_comparators"] = bench_ba_cpu(ctx)
    return res


def bench_ba_fixture(ctx):
    """The reference's own reconstruction as a BA problem (57 cameras of pose.csv x 19 282 points of sparse.ply,
    1 061 813 observations; tests/golden/gustav_scene.npz): LM from the perturbed start to convergence."""
    import sfm_mvs_b200 as sfm
    from sfm_mvs_b200 import synth
    path = os.path.join(ROOT, "tests", "golden", "gustav_scene.npz")
    if not os.path.isfile(path):
        return None
    g = np.load(path)
    pb = synth.ba_problem_from_scene(g["K"], g["cams"], g["pts"])
    prob = sfm.BAProblem(ctx, len(pb["cams0"]), len(pb["pts0"]), pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
    prob.set_params(pb["cams0"], pb["pts0"])
    prob.solve(max_iters=3)
    prob.set_params(pb["cams0"], pb["pts0"])
    ctx.sync()
    t0 = time.perf_counter()
    hist = prob.solve(max_iters=30, ftol=1e-6)
    ctx.sync()
    dt = time.perf_counter() - t0
    prob.close()
    return {"workload": f"{len(pb['cams0'])} cams / {len(pb['pts0'])} points / {len(pb['obs'])} obs (pose.csv + sparse.ply of the reference)",
            "iterations": len(hist), "seconds": dt, "iters_per_s": len(hist) / dt, "cost_first": hist[0]["cost_before"],
            "cost_last": hist[-1]["cost_after"], "rms_px_last": float(np.sqrt(hist[-1]["cost_after"] / len(pb["obs"])))}


def bench_ba_cpu(ctx):
    """CPU arms beside the BA numbers: (a) the reference's BundleAdjustment (sfm.py:138-157, oracle port:
    scipy TRF over a dense finite-difference Jacobian of a Python-loop residual) against the drop-in
    sfm_mvs_b200.BundleAdjustment (same formulation, same optimiser, residual + Jacobian on the GPU) on the same
    single-camera problems; (b) scipy's sparse TRF on the multi-camera formulation (notebook cell 6: jac_sparsity,
    x_scale='jac', ftol=1e-4) against the engine's LM on the same small problem."""
    from scipy.optimize import least_squares
    from scipy.sparse import lil_matrix

    import sfm_mvs_b200 as sfm
    from oracle import cvpath
    from sfm_mvs_b200 import synth
    out = {"reference_BundleAdjustment": []}
    K = synth.K_GUSTAV
    for N in (50, 200):
        rng = np.random.default_rng(N)
        R, t = synth.orbit_pose(0.1)
        Xw = np.c_[rng.uniform(-2, 2, N), rng.uniform(-1.5, 1.5, N), rng.uniform(5, 11, N)]
        uv, _ = synth.project(K, R, t, Xw)
        obs = (uv + rng.normal(0, 0.7, uv.shape)).astype(np.float32).T.copy()
        X0 = (Xw + rng.normal(0, 0.03, Xw.shape)).astype(np.float32).reshape(N, 1, 3)
        Rt0 = np.hstack([R, t])
        sfm.BundleAdjustment(X0, obs, Rt0, K, 0.5, ctx=ctx)                      # warm-up
        t0 = time.perf_counter()
        Xe, pe, Rte = sfm.BundleAdjustment(X0, obs, Rt0, K, 0.5, ctx=ctx)
        t_gpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        Xr, pr, Rtr = quiet(cvpath.BundleAdjustment, X0, obs, Rt0, K, 0.5)
        t_cpu = time.perf_counter() - t0
        out["reference_BundleAdjustment"].append({"N": N, "cpu_s": t_cpu, "engine_s": t_gpu, "speedup": t_cpu / t_gpu,
                                                  "max_rel_dX": float(np.abs(Xe - Xr).max() / np.abs(Xr).max()),
                                                  "max_abs_dRt": float(np.abs(Rte - Rtr).max())})
    # (b) sparse TRF on [cams (rvec, tvec) | points], residual proj - obs
    pb = synth.ba_problem(20, 2000, 5, seed=2)
    C_, P_, O = 20, 2000, len(pb["obs"])
    ci, pi = pb["cam_idx"], pb["pt_idx"]

    Kb, obs64 = pb["K"], pb["obs"].astype(np.float64)

    def fun(x):                      # notebook cells 3-5: rotate (Rodrigues' formula, vectorised) + project, residual proj - obs
        cams, pts = x[:6 * C_].reshape(C_, 6), x[6 * C_:].reshape(P_, 3)
        rv, X = cams[ci, :3], pts[pi]
        th = np.linalg.norm(rv, axis=1)[:, None]
        with np.errstate(invalid="ignore"):
            v = np.nan_to_num(rv / th)
        dot = np.sum(X * v, axis=1)[:, None]
        Y = np.cos(th) * X + np.sin(th) * np.cross(v, X) + dot * (1 - np.cos(th)) * v + cams[ci, 3:]
        uv = np.stack([Kb[0, 0] * Y[:, 0] / Y[:, 2] + Kb[0, 2], Kb[1, 1] * Y[:, 1] / Y[:, 2] + Kb[1, 2]], 1)
        return (uv - obs64).ravel()

    A = lil_matrix((2 * O, 6 * C_ + 3 * P_), dtype=int)
    rows = np.arange(O)
    for k in range(6):
        A[2 * rows, 6 * ci + k] = 1
        A[2 * rows + 1, 6 * ci + k] = 1
    for k in range(3):
        A[2 * rows, 6 * C_ + 3 * pi + k] = 1
        A[2 * rows + 1, 6 * C_ + 3 * pi + k] = 1
    x0 = np.hstack([pb["cams0"].ravel(), pb["pts0"].ravel()])
    t0 = time.perf_counter()
    sol = least_squares(fun, x0, jac_sparsity=A, x_scale="jac", ftol=1e-4, method="trf")
    t_cpu = time.perf_counter() - t0
    prob = sfm.BAProblem(ctx, C_, P_, ci, pi, pb["obs"], pb["K"])
    prob.set_params(pb["cams0"], pb["pts0"])
    prob.solve(max_iters=2)
    prob.set_params(pb["cams0"], pb["pts0"])
    ctx.sync()
    t0 = time.perf_counter()
    hist = prob.solve(max_iters=30, ftol=1e-4)
    ctx.sync()
    t_gpu = time.perf_counter() - t0
    prob.close()
    out["scipy_sparse_trf"] = {"workload": f"{C_} cams / {P_} points / {O} obs", "cpu_s": t_cpu, "cpu_nfev": int(sol.nfev),
                               "cpu_cost": float(sol.cost), "engine_s": t_gpu, "engine_iters": len(hist),
                               "engine_cost": hist[-1]["cost_after"], "speedup": t_cpu / t_gpu}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)
