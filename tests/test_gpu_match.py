"""GPU parity: hot path 1 (2-NN matching + Lowe ratio) against cv2 / the oracle — bit-exact."""
import cv2
import numpy as np
import pytest

import sfm_mvs_b200 as sfm
from oracle import cvpath, restated
from sfm_mvs_b200 import synth

pytestmark = pytest.mark.gpu


def _check_pair(engine, q, t, mode):
    idx, dist, good, ng = engine.knn2(q, t, 0.70, mode=mode)
    cidx, cdist = cvpath.knn2_arrays(q.astype(np.float32), t.astype(np.float32))
    k = cidx.shape[1]
    assert np.array_equal(idx[:, :k], cidx), f"indices differ in {np.sum(idx[:, :k] != cidx)} places"
    assert np.array_equal(dist[:, :k], cdist)
    if k == 2:
        ref_good = restated.ratio_mask(cdist)
        assert np.array_equal(good.astype(bool), ref_good)
        assert ng == int(ref_good.sum())


@pytest.mark.parametrize("mode", [2, 1])
@pytest.mark.parametrize("nq,nt,seed", [(257, 301, 0), (1000, 900, 1), (128, 256, 2), (2000, 2000, 3),
                                        (5000, 5000, 4), (1, 700, 5), (300, 2, 6), (129, 513, 7)])
def test_knn2_equals_cv2(engine, nq, nt, seed, mode):
    q, t, _ = synth.matching_pair(nq, nt, seed=seed)
    _check_pair(engine, q, t, mode)


def test_tensor_core_accumulator_is_the_exact_squared_distance(engine):
    q, t, _ = synth.matching_pair(200, 300, seed=11)
    dq, dt = engine.descriptors(q), engine.descriptors(t)
    assert dq.exact and dt.exact
    acc = engine.debug_tc_accumulators(dq, dt)[:200, :300]
    d2 = ((q[:, None, :].astype(np.float64) - t[None, :, :].astype(np.float64)) ** 2).sum(-1)
    assert np.array_equal(acc.astype(np.float64), -(4194304.0 + d2 / 2.0))


def test_ties_go_to_lower_train_index(engine):
    q = synth.sift_like_descriptors(300, 5)
    t = np.vstack([q[::-1], q[::-1]])
    for mode in (1, 2):
        idx, dist, _, _ = engine.knn2(q, t, mode=mode)
        cidx, cdist = cvpath.knn2_arrays(q, t)
        assert np.array_equal(idx, cidx) and np.array_equal(dist, cdist)
        assert np.all(idx[:, 0] < idx[:, 1]) and np.all(dist == 0)


def test_real_sift_pair_golden(engine, golden):
    g = golden("real_pair")
    for des0, des1 in ((g["des0"], g["des1"]), (g["des0"].astype(np.float32), g["des1"].astype(np.float32))):
        p0, p1 = sfm.match_keypoints(g["kp0"], des0, g["kp1"], des1, ctx=engine)
        assert np.array_equal(p0, g["pts0"]) and np.array_equal(p1, g["pts1"])


def test_non_integer_descriptors_take_the_fp32_kernel(engine):
    rng = np.random.default_rng(0)
    q = rng.random((300, 128), dtype=np.float32)
    t = rng.random((400, 128), dtype=np.float32)
    d = engine.descriptors(q)
    assert not d.exact
    idx, dist, _, _ = engine.knn2(q, t)
    cidx, cdist = cvpath.knn2_arrays(q, t)
    # float accumulation order differs from OpenCV's SIMD lanes: indices must agree wherever the two
    # nearest candidates are separated by more than rounding noise
    sep = np.abs(cdist[:, 1] - cdist[:, 0]) > 1e-4
    assert np.array_equal(idx[sep, 0], cidx[sep, 0])
    assert np.allclose(dist, cdist, rtol=1e-5)
    with pytest.raises(sfm.error):
        engine.knn2(q, t, mode=2)


def test_knnmatch_dmatch_surface_and_edge_cases(engine):
    q, t, _ = synth.matching_pair(50, 60, seed=3)
    ours = sfm.BFMatcher(ctx=engine).knnMatch(q, t, k=2)
    ref = cv2.BFMatcher().knnMatch(q, t, k=2)
    assert len(ours) == len(ref)
    for a, b in zip(ours, ref):
        assert len(a) == len(b) == 2
        for m, n in zip(a, b):
            assert (m.queryIdx, m.trainIdx, m.imgIdx, m.distance) == (n.queryIdx, n.trainIdx, n.imgIdx, n.distance)
    good = [m for m, n in ours if m.distance < 0.70 * n.distance]          # the reference's loop, sfm.py:262-265
    good_ref = [m for m, n in ref if m.distance < 0.70 * n.distance]
    assert [(m.queryIdx, m.trainIdx) for m in good] == [(m.queryIdx, m.trainIdx) for m in good_ref]
    # nt == 1 -> 1-tuples ; nt == 0 -> empty tuples ; nq == 0 -> ()
    one = sfm.BFMatcher(ctx=engine).knnMatch(q, t[:1], k=2)
    assert all(len(x) == 1 for x in one) and [x[0].distance for x in one] == [x[0].distance for x in cv2.BFMatcher().knnMatch(q, t[:1], k=2)]
    assert sfm.BFMatcher(ctx=engine).knnMatch(q, t[:0], k=2) == tuple(() for _ in range(50))
    assert sfm.BFMatcher(ctx=engine).knnMatch(q[:0], t, k=2) == ()
    with pytest.raises(sfm.error):
        sfm.BFMatcher(ctx=engine).knnMatch(q.astype(np.float64), t.astype(np.float64), k=2)
    with pytest.raises(sfm.error):
        sfm.BFMatcher(ctx=engine).knnMatch(q, t[:, :64].copy(), k=2)


def test_large_pair_properties(engine):
    """BASELINE sweep size (16k x 16k): too slow for the numpy oracle in full, so (a) a random sample of
    query rows is checked exhaustively and (b) matching a set against itself returns the identity."""
    q, t, _ = synth.matching_pair(16384, 16384, seed=9)
    idx, dist, good, _ = engine.knn2(q, t, mode=2)
    rows = np.random.default_rng(0).choice(16384, 256, replace=False)
    ridx, rdist = restated.knn2_l2(q[rows], t)
    assert np.array_equal(idx[rows], ridx) and np.array_equal(dist[rows], rdist)
    sidx, sdist, _, _ = engine.knn2(t, t, mode=2)
    # duplicates inside t are possible in principle; distance 0 with the lowest index is the contract
    assert np.all(sdist[:, 0] == 0) and np.all(sidx[:, 0] <= np.arange(16384))


def test_batched_pairs_equal_single_pair_calls(engine):
    """sfm_desc_match_gather_batched: one K1 launch over the items of all pairs (different sizes, one pair
    that must take the fp32 kernel) == per-pair knn2 + gather, bit for bit."""
    import ctypes as C
    import torch
    from sfm_mvs_b200._lib import check, lib
    rng = np.random.default_rng(3)
    sizes = [(700, 900), (257, 129), (1500, 1500), (300, 400), (1, 50), (2100, 640)]
    sets = []
    for k, (nq, nt) in enumerate(sizes):
        q, t, _ = synth.matching_pair(nq, nt, seed=20 + k)
        if k == 3:                                   # non-integer descriptors -> fp32 kernel inside the batch
            q = q + rng.random(q.shape, dtype=np.float32) * 0.25
        sets.append((q, t, rng.random((nq, 2), dtype=np.float32) * 900, rng.random((nt, 2), dtype=np.float32) * 900))
    dq = [engine.descriptors(s[0]) for s in sets]
    dt = [engine.descriptors(s[1]) for s in sets]
    dev = engine.torch_device
    with torch.cuda.stream(engine.torch_stream()):
        kq = [torch.from_numpy(s[2]).to(dev) for s in sets]
        kt = [torch.from_numpy(s[3]).to(dev) for s in sets]
        pq = [torch.zeros((s[0].shape[0], 2), dtype=torch.float32, device=dev) for s in sets]
        pt = [torch.zeros((s[0].shape[0], 2), dtype=torch.float32, device=dev) for s in sets]
        qi = [torch.zeros((s[0].shape[0],), dtype=torch.int32, device=dev) for s in sets]
        ti = [torch.zeros((s[0].shape[0],), dtype=torch.int32, device=dev) for s in sets]
        idx = [torch.zeros((s[0].shape[0], 2), dtype=torch.int32, device=dev) for s in sets]
        good = [torch.zeros((s[0].shape[0],), dtype=torch.uint8, device=dev) for s in sets]
        n_out = torch.zeros((len(sets),), dtype=torch.int32, device=dev)
    arr = lambda xs: np.array(xs, np.uint64)
    a = [arr([d._h.value for d in dq]), arr([d._h.value for d in dt]), arr([x.data_ptr() for x in kq]),
         arr([x.data_ptr() for x in kt]), arr([x.data_ptr() for x in idx]), arr([x.data_ptr() for x in good]),
         arr([x.data_ptr() for x in pq]), arr([x.data_ptr() for x in pt]), arr([x.data_ptr() for x in qi]),
         arr([x.data_ptr() for x in ti])]
    check(lib.sfm_desc_match_gather_batched(engine._h, len(sets), a[0].ctypes.data, a[1].ctypes.data, 0.70, a[2].ctypes.data,
                                            a[3].ctypes.data, a[4].ctypes.data, a[5].ctypes.data, a[6].ctypes.data,
                                            a[7].ctypes.data, a[8].ctypes.data, a[9].ctypes.data, n_out.data_ptr()))
    engine.sync()
    counts = n_out.cpu().numpy()
    for k, (q, t, kpq, kpt) in enumerate(sets):
        ridx, rdist, rgood, rng_ = engine.knn2(q, t, 0.70)
        assert np.array_equal(idx[k].cpu().numpy(), ridx), f"pair {k}: indices differ"
        assert np.array_equal(good[k].cpu().numpy().astype(bool), rgood.astype(bool))
        m = int(counts[k])
        assert m == int(rgood.sum())
        sel = np.flatnonzero(rgood)
        assert np.array_equal(qi[k].cpu().numpy()[:m], sel)
        assert np.array_equal(ti[k].cpu().numpy()[:m], ridx[sel, 0])
        assert np.array_equal(pq[k].cpu().numpy()[:m], kpq[sel])
        assert np.array_equal(pt[k].cpu().numpy()[:m], kpt[ridx[sel, 0]])
    if sizes[0][1] > 2:
        cidx, _ = cvpath.knn2_arrays(sets[0][0], sets[0][1])
        assert np.array_equal(idx[0].cpu().numpy(), cidx)


def test_batched_descriptor_preparation_equals_per_view(engine):
    """sfm_desc_create_batched (one K1b launch for many views, mixed sizes) gives the same matches as per-view sets;
    a non-integer set inside the batch is flagged and still matched by the fp32 kernel."""
    import torch
    rng = np.random.default_rng(5)
    sets = [synth.sift_like_descriptors(n, seed=30 + k) for k, n in enumerate((700, 129, 1500, 256, 1))]
    sets[3] = sets[3] + rng.random(sets[3].shape, dtype=np.float32) * 0.3
    dev = engine.torch_device
    with torch.cuda.stream(engine.torch_stream()):
        tens = [torch.from_numpy(s).to(dev) for s in sets]
    batch = sfm.Descriptors.batch(engine, tens)
    single = [engine.descriptors(s) for s in sets]
    assert [d.exact for d in batch] == [d.exact for d in single] == [True, True, True, False, True]
    from sfm_mvs_b200._lib import check, lib
    for a in range(len(sets) - 1):
        n = sets[a].shape[0]
        out = []
        for q, t in ((batch[a], batch[a + 1]), (single[a], single[a + 1])):
            idx = torch.zeros((n, 2), dtype=torch.int32, device=dev)
            good = torch.zeros((n,), dtype=torch.uint8, device=dev)
            check(lib.sfm_desc_match(engine._h, q._h, t._h, 0.7, idx.data_ptr(), None, good.data_ptr(), None, 0))
            engine.sync()
            out.append((idx.cpu().numpy(), good.cpu().numpy()))
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("n", [16384, 65536])
def test_full_size_pairs_by_properties(engine, n):
    """BASELINE configs[4] sizes (up to 64k x 64k per pair; the oracle would take minutes): size-independent
    properties.  train = a permutation of query plus distinct extra rows => every query's nearest neighbour is its own
    copy at distance 0 with the index the permutation says; the runner-up is strictly farther; and the 2-NN of a
    sample of query rows equals an exact brute-force evaluation (float64 on integer data) done with torch."""
    import torch
    rng = np.random.default_rng(n)
    q = synth.sift_like_descriptors(n, seed=n % 97)
    # make rows unique (the generator can repeat a row at this size): stamp the row index into two components
    q[:, 0] = (np.arange(n) % 251).astype(np.float32)
    q[:, 1] = ((np.arange(n) // 251) % 251).astype(np.float32)
    q[:, 2] = ((np.arange(n) // (251 * 251)) % 251).astype(np.float32)
    perm = rng.permutation(n)
    t = q[perm]
    inv = np.empty(n, np.int64); inv[perm] = np.arange(n)
    idx, dist, good, ng = engine.knn2(q, t, 0.70, mode=2)
    assert np.array_equal(idx[:, 0], inv), "nearest neighbour is not the permuted copy"
    assert np.all(dist[:, 0] == 0) and np.all(dist[:, 1] > 0)
    assert ng == n and good.all()                        # 0 < 0.7 * d2 everywhere
    rows = rng.choice(n, 48, replace=False)
    tq = torch.from_numpy(q[rows]).to("cuda", torch.float64)
    tt = torch.from_numpy(t).to("cuda", torch.float64)
    d2 = ((tq * tq).sum(1)[:, None] + (tt * tt).sum(1)[None, :] - 2.0 * tq @ tt.T)       # exact: integers < 2^53
    key = d2 * float(n) + torch.arange(n, device="cuda", dtype=torch.float64)[None, :]     # (distance, index) order
    top = torch.topk(key, 2, dim=1, largest=False).indices.cpu().numpy()
    assert np.array_equal(idx[rows], top)
    ref_d = np.sqrt(torch.gather(d2, 1, torch.from_numpy(top).to("cuda")).cpu().numpy().astype(np.float32))
    assert np.array_equal(dist[rows], ref_d)


def test_closing_a_context_closes_what_lives_on_it():
    """Descriptor sets, BA problems and chains hand their buffers back to pools of their context when destroyed: an
    explicit Context.close() therefore closes them first, and their later destruction is a no-op (not a use after free)."""
    ctx = sfm.Context(0)
    rng = np.random.default_rng(0)
    d = sfm.Descriptors(ctx, rng.integers(0, 200, (300, 128)).astype(np.uint8))
    pb = synth.ba_problem(4, 60, 3, seed=1)
    prob = sfm.BAProblem(ctx, 4, 60, pb["cam_idx"], pb["pt_idx"], pb["obs"], pb["K"])
    ctx.close()
    assert not d._h and not prob._h and not ctx._h
    d.close()
    prob.close()
    del d, prob
