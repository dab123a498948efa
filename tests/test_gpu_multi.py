"""Two-GPU parity (skipped with fewer than two devices; run with `gpurun --gpus 2 -- pytest tests/test_gpu_multi.py -m gpu`):
point-sharded bundle adjustment through csrc/comm.cu (NCCL all-reduce of the reduced camera system) equals the
single-rank solve, and pair-sharded matching (pipeline.match_pairs_sharded) reports the single-rank counts."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    import sfm_mvs_b200 as sfm
    from sfm_mvs_b200 import pipeline, sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = sfm.Context(rank)
    # ---- bundle adjustment: this rank's points, all cameras, the exchange of csrc/comm.cu
    pb = synth.ba_problem(24, 3000, 6, seed=3)
    sh = sharding.ba_shard(pb, rank, world)
    prob = sfm.BAProblem(ctx, 24, len(sh["pts0"]), sh["cam_idx"], sh["pt_idx"], sh["obs"], sh["K"], totals=sh["totals"])
    prob.set_params(sh["cams0"], sh["pts0"])
    uid = [sfm.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    prob.comm_init(uid[0], rank, world)
    lam, hist = 1e-3, []
    for _ in range(4):
        st = prob.gn_step(lam)
        lam = st["lambda_next"]
        hist.append((st["cost_before"], st["cost_after"], st["accepted"]))
    cams, pts = prob.get_params()
    prob.close()
    # ---- matching: the all-previous-views pair list split over the ranks
    scene = synth.orbit_scene(7, 900, seed=2)
    views = [pipeline.DeviceView(ctx, v["kp"], v["des"]) for v in scene["views"]]
    pairs = sharding.all_pairs(7)
    mine, matches, counts = pipeline.match_pairs_sharded(ctx, views, pairs, rank, world, dist=dist)
    q.put((rank, hist, cams, pts, sh["p_lo"], mine, [pm.n for pm in matches], counts.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ba_and_sharded_matching_equal_single_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    import sfm_mvs_b200 as sfm
    from sfm_mvs_b200 import pipeline, sharding, synth
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = _free_port()
    procs = [mpctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-rank references
    ctx = sfm.default_context()
    pb = synth.ba_problem(24, 3000, 6, seed=3)
    prob = sfm.BAProblem(ctx, 24, 3000, pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
    prob.set_params(pb["cams0"], pb["pts0"])
    lam, hist = 1e-3, []
    for _ in range(4):
        st = prob.gn_step(lam)
        lam = st["lambda_next"]
        hist.append((st["cost_before"], st["cost_after"], st["accepted"]))
    cams, pts = prob.get_params()
    prob.close()
    for rank, h2, cams2, pts2, p_lo, *_ in res:
        for (a0, a1, acc), (b0, b1, acc2) in zip(hist, h2):
            # (S is accumulated in float32 by atomics: the order differs between one and two ranks)
            assert acc == acc2 and abs(a0 - b0) <= 1e-4 * a0 and abs(a1 - b1) <= 1e-4 * a1
        assert np.abs(cams2 - cams).max() < 1e-4
        assert np.abs(pts2 - pts[p_lo:p_lo + len(pts2)]).max() < 1e-3
    scene = synth.orbit_scene(7, 900, seed=2)
    views = [pipeline.DeviceView(ctx, v["kp"], v["des"]) for v in scene["views"]]
    pairs = sharding.all_pairs(7)
    _, matches, counts = pipeline.match_pairs_sharded(ctx, views, pairs, 0, 1)
    want = np.array([pm.n for pm in matches], np.int32)
    assert np.array_equal(counts.cpu().numpy(), want)
    assert sorted(res[0][5] + res[1][5]) == list(range(len(pairs)))
    for rank, *_, mine, local, allc in res:
        assert np.array_equal(allc, want) and local == [int(want[k]) for k in mine]
