"""Host-side partitioning of the hot path over ranks (one process per GPU).

  * matching: independent (query view, train view) pairs — sfm.py:347 consecutive pairs or the
    all-previous-pairs loop of isfm.py:68-87 — assigned by longest-processing-time-first on the
    cost Nq*Nt.  No data-path collective.
  * bundle adjustment: points (with all their observations) in contiguous ranges balanced by
    observation count; every rank holds all cameras.  The only exchange per Gauss-Newton iteration
    is the all-reduce(sum) of the partial reduced camera systems (csrc/comm.cu).
  * the per-view registration chain is sequential (view i+1 needs view i's pose, sfm.py:399-409):
    replicas only — independent scenes per rank.
"""
from __future__ import annotations

import numpy as np


def shard_pairs(pairs, costs, world: int):
    """LPT assignment (longest processing time first: every pair, in order of decreasing cost, goes to the least
    loaded rank; ties to the lowest rank).  Returns a list (per rank) of lists of pair indices, each ascending.
    Equal costs — every view has the same number of descriptors — make LPT a round robin, taken without the loop
    (the list has V^2 / 2 entries and this runs inside the timed matching call)."""
    import heapq
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if world <= 1:
        return [list(range(n))]
    if n and costs.min() == costs.max():
        return [list(range(r, n, world)) for r in range(world)]
    order = np.argsort(-costs, kind="stable")
    heap = [(0.0, r) for r in range(world)]
    out = [[] for _ in range(world)]
    for k in order.tolist():
        load, r = heapq.heappop(heap)
        out[r].append(k)
        heapq.heappush(heap, (load + costs[k], r))
    return [sorted(o) for o in out]


def shard_points(pt_idx: np.ndarray, n_pt: int, world: int):
    """Contiguous point ranges with ~equal observation counts.  pt_idx must be non-decreasing.
    Returns [(p_lo, p_hi, o_lo, o_hi)] per rank."""
    pt_idx = np.asarray(pt_idx)
    counts = np.bincount(pt_idx, minlength=n_pt)
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        p = int(np.searchsorted(cum, target, side="left"))
        bounds.append(min(max(p, bounds[-1]), n_pt))
    bounds.append(n_pt)
    return [(bounds[r], bounds[r + 1], int(cum[bounds[r]]), int(cum[bounds[r + 1]])) for r in range(world)]


def ba_shard(problem: dict, rank: int, world: int) -> dict:
    """Slice a synth.ba_problem-style dict for one rank (local point indices start at 0)."""
    n_pt = len(problem["pts0"])
    p_lo, p_hi, o_lo, o_hi = shard_points(problem["pt_idx"], n_pt, world)[rank]
    return dict(K=problem["K"], cams0=problem["cams0"], pts0=problem["pts0"][p_lo:p_hi],
                cam_idx=problem["cam_idx"][o_lo:o_hi], pt_idx=problem["pt_idx"][o_lo:o_hi] - p_lo,
                obs=problem["obs"][o_lo:o_hi], p_lo=p_lo, p_hi=p_hi, totals=(n_pt, len(problem["pt_idx"])))


def all_pairs(n_views: int):
    """The pair list of isfm.py:68-87 in the reference's order: for every view i, every earlier view j."""
    return [(j, i) for i in range(n_views) for j in range(i)]


def split_rows(n_rows: int, world: int, align: int = 128):
    """One big pair split by query rows (BASELINE configs[4], strong scaling): contiguous row ranges, boundaries
    aligned to the matcher's 128-row query tile.  Returns [(lo, hi)] per rank (possibly empty ranges)."""
    tiles = (n_rows + align - 1) // align
    out = []
    for r in range(world):
        lo = min(n_rows, (tiles * r // world) * align)
        hi = min(n_rows, (tiles * (r + 1) // world) * align)
        out.append((lo, hi))
    return out


def gather_pair_counts(local_counts, my_pairs, n_pairs: int, dist=None, device=None):
    """Survivor counts of ALL pairs on every rank.  The matches themselves stay resident on the rank that computed
    them (nothing later in isfm.py needs them elsewhere: the essential matrix of a pair is estimated where its
    matches are); only the per-pair counts — what isfm.py:86 prints — are exchanged: each rank writes its counts into
    a zero vector at its pairs' slots and the vectors are summed (one all-reduce of n_pairs int32, NCCL on the GPUs,
    gloo in the CPU tests).  dist: torch.distributed or None for a single rank."""
    import torch
    buf = torch.zeros((n_pairs,), dtype=torch.int32, device=device)
    if len(my_pairs):
        idx = torch.as_tensor(list(my_pairs), dtype=torch.long, device=device)
        buf[idx] = torch.as_tensor(local_counts, dtype=torch.int32, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf)
    return buf
