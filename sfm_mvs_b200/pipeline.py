"""The reference's per-view registration loop (sfm.py:341-409) on the engine, device-resident.

One *registered view* = one iteration of that loop with imread / SIFT / GUI removed:
match(prev, new) + ratio test, re-triangulation of the previous pair's matches, data association
(common_points), PnP-RANSAC, reprojection error, triangulation of the new points, reprojection
error.  Keypoints, descriptors, matches, 3-D points and masks live in HBM (torch CUDA tensors are
used as plain device buffers); per view the host learns three integers (match / association
counts) and the pose.  Matching does not depend on poses, so consecutive pairs are matched in batches
(one K1 launch per batch; this is also the unit that shards over GPUs, isfm.py:68-87) on a second context
while the registration loop of the previous batch runs (register_device / register_host).  two_view_init is
the reference's initialisation (sfm.py:307-316), pairwise_init the all-earlier-views loop of isfm.py:68-87.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import engine as _e
from ._lib import check, lib


def _on_ctx_stream(fn):
    """Run a method with torch's current stream set to the engine context's stream."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        import torch
        with torch.cuda.stream(self.ctx.torch_stream()):
            return fn(self, *a, **k)
    return wrapper


def _on_ctx_stream_fn(fn):
    """The same for a function whose first argument is the context."""
    import functools

    @functools.wraps(fn)
    def wrapper(ctx, *a, **k):
        import torch
        with torch.cuda.stream(ctx.torch_stream()):
            return fn(ctx, *a, **k)
    return wrapper


class DeviceView:
    """A view's keypoints and prepared descriptors resident in HBM."""

    def __init__(self, ctx: _e.Context, kp, des, desc=None):
        import torch
        self.ctx = ctx
        self.n = int(len(kp))
        if _e._is_torch(kp) and kp.is_cuda and kp.dtype == torch.float32 and kp.is_contiguous() and kp.device == ctx.torch_device:
            self.kp = kp                      # resident already: nothing to enqueue (thousands of views per call otherwise pay
        else:                                 # a stream-context switch each)
            with torch.cuda.stream(ctx.torch_stream()):
                if _e._is_torch(kp):
                    self.kp = kp.to(device=ctx.torch_device, dtype=torch.float32).contiguous()
                else:
                    self.kp = torch.from_numpy(np.ascontiguousarray(kp, np.float32)).to(ctx.torch_device)
        self.desc = desc if desc is not None else ctx.descriptors(des)

    @classmethod
    def batch(cls, ctx: _e.Context, kps, dess):
        """Views whose descriptors are CUDA tensors: ONE K1b launch prepares all of them."""
        if dess and all(_e._is_torch(d) and d.is_cuda for d in dess):
            import torch
            with torch.cuda.stream(ctx.torch_stream()):
                descs = _e.Descriptors.batch(ctx, list(dess))
            return [cls(ctx, k, None, desc=d) for k, d in zip(kps, descs)]
        return [cls(ctx, k, d) for k, d in zip(kps, dess)]


# sfm_view_out (include/sfm_b200.h)
VIEW_OUT = np.dtype([("Rt", np.float64, (12,)), ("err_pnp", np.float64), ("err_new", np.float64),
                     ("n_new", np.int32), ("n_pnp", np.int32), ("n_inl", np.int32), ("n_match", np.int32)])


class PairMatches:
    """The surviving matches of one pair: row range [lo, hi) of the batch's output slabs (views are cut on demand —
    thousands of pairs come out of one launch and most callers only need the pointers)."""
    __slots__ = ("_slabs", "lo", "hi", "k", "n")

    def __init__(self, slabs, lo, hi, k, n):
        self._slabs, self.lo, self.hi, self.k, self.n = slabs, lo, hi, k, n

    pts_q = property(lambda self: self._slabs[0][self.lo:self.hi])
    pts_t = property(lambda self: self._slabs[1][self.lo:self.hi])
    qidx = property(lambda self: self._slabs[2][self.lo:self.hi])
    tidx = property(lambda self: self._slabs[3][self.lo:self.hi])
    n_dev = property(lambda self: self._slabs[4][self.k:self.k + 1])
    ptr_q = property(lambda self: self._slabs[0].data_ptr() + 8 * self.lo)
    ptr_t = property(lambda self: self._slabs[1].data_ptr() + 8 * self.lo)


class RegistrationChain:
    def __init__(self, ctx: _e.Context, K, ratio: float = 0.70, hypothesis_fn=None):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.K = np.ascontiguousarray(K, np.float64)
        self.ratio = float(ratio)
        self.hypothesis_fn = hypothesis_fn     # parity mode: minimal solutions supplied by the caller
        self.dev = ctx.torch_device

    # ------------------------------------------------------------------ matching (batched over pairs)
    @_on_ctx_stream
    def match_pairs(self, views, pairs):
        """knn2 + ratio + survivor gather for every (q, t) view-index pair: ONE K1 launch over the items of
        all pairs, one K1c grid, one gather grid (sfm_desc_match_gather_batched); one host sync at the end."""
        torch = self.torch
        n = len(pairs)
        if n == 0:
            return []
        nq = np.array([v.n for v in views], np.int64)[[a for a, _ in pairs]]
        off = np.concatenate([[0], np.cumsum(nq)])
        tot = int(max(off[-1], 1))
        counts = torch.empty((n,), dtype=torch.int32, device=self.dev)
        pts_q = torch.empty((tot, 2), dtype=torch.float32, device=self.dev)
        pts_t = torch.empty((tot, 2), dtype=torch.float32, device=self.dev)
        qidx = torch.empty((tot,), dtype=torch.int32, device=self.dev)
        tidx = torch.empty((tot,), dtype=torch.int32, device=self.dev)

        def ptrs(base, stride):
            return np.ascontiguousarray(base + off[:-1] * stride, np.uint64)

        # per-view handles and pointers once, then indexed by the pair list
        hv = np.array([v.desc._h.value if hasattr(v.desc._h, "value") else v.desc._h for v in views], np.uint64)
        kv = np.array([v.kp.data_ptr() for v in views], np.uint64)
        pa = np.array([a for a, _ in pairs], np.int64)
        pb = np.array([b for _, b in pairs], np.int64)
        hq, ht, kq, kt = (np.ascontiguousarray(x) for x in (hv[pa], hv[pb], kv[pa], kv[pb]))
        a_pq, a_pt = ptrs(pts_q.data_ptr(), 8), ptrs(pts_t.data_ptr(), 8)
        a_qi, a_ti = ptrs(qidx.data_ptr(), 4), ptrs(tidx.data_ptr(), 4)
        check(lib.sfm_desc_match_gather_batched(self.ctx._h, n, hq.ctypes.data, ht.ctypes.data, self.ratio, kq.ctypes.data,
                                                kt.ctypes.data, None, None, a_pq.ctypes.data, a_pt.ctypes.data,
                                                a_qi.ctypes.data, a_ti.ctypes.data, counts.data_ptr()))
        self.ctx.sync()
        host = counts.cpu().numpy().tolist()
        slabs = (pts_q, pts_t, qidx, tidx, counts)
        offs = off.tolist()
        return [PairMatches(slabs, offs[k], offs[k + 1], k, host[k]) for k in range(n)]

    # ------------------------------------------------------------------ geometry helpers (device buffers)
    def _triangulate(self, P1, P2, x1, x2, n, out_layout):
        torch = self.torch
        width = 3 if out_layout == 2 else 4
        X = torch.empty((max(n, 1), width), dtype=torch.float32, device=self.dev)
        if n == 0:
            return X
        P1 = np.ascontiguousarray(P1, np.float64)
        P2 = np.ascontiguousarray(P2, np.float64)
        check(lib.sfm_triangulate(self.ctx._h, _e._dptr(P1), _e._dptr(P2), x1.data_ptr(), x2.data_ptr(), n, 1,
                                  X.data_ptr(), out_layout, 1))
        return X

    def _reproj(self, X, x_layout, px, n, Rt, err_slot, X3=None):
        if n == 0:
            return
        Rt = np.ascontiguousarray(Rt, np.float64)
        check(lib.sfm_reproj_error(self.ctx._h, X.data_ptr(), x_layout, px.data_ptr(), 1, n, _e._dptr(Rt),
                                   _e._dptr(self.K), err_slot.data_ptr(), None, None if X3 is None else X3.data_ptr()))

    def _gather(self, src, width, idx, n):
        dst = self.torch.empty((max(n, 1), width), dtype=self.torch.float32, device=self.dev)
        check(lib.sfm_gather_rows(self.ctx._h, src.data_ptr(), width, idx.data_ptr(), n, dst.data_ptr()))
        return dst

    def _pnp(self, X, px, n):
        """-> ok, rvec, tvec, n_inliers, inlier index buffer (device)."""
        torch = self.torch
        rvec, tvec = np.zeros(3), np.zeros(3)
        inl = torch.empty((max(n, 1),), dtype=torch.int32, device=self.dev)
        ni, ok = C.c_int32(0), C.c_int32(0)
        if self.hypothesis_fn is None:
            check(lib.sfm_pnp_ransac(self.ctx._h, X.data_ptr(), px.data_ptr(), n, _e._dptr(self.K.reshape(9)), 100, 8.0,
                                     0.99, _e._dptr(rvec), _e._dptr(tvec), inl.data_ptr(), C.byref(ni), C.byref(ok), None))
        else:
            self.ctx.sync()
            hyp, valid = self.hypothesis_fn(X[:n].cpu().numpy(), px[:n].cpu().numpy())
            hyp = np.ascontiguousarray(hyp, np.float64)
            valid = np.ascontiguousarray(valid, np.uint8)
            check(lib.sfm_pnp_ransac_hyp(self.ctx._h, X.data_ptr(), px.data_ptr(), n, _e._dptr(self.K.reshape(9)),
                                         _e._dptr(hyp), _e._dptr(valid), len(hyp), 8.0, 0.99, _e._dptr(rvec),
                                         _e._dptr(tvec), inl.data_ptr(), C.byref(ni), C.byref(ok), None))
        return bool(ok.value), rvec, tvec, int(ni.value), inl

    # ------------------------------------------------------------------ the loop
    @_on_ctx_stream
    def bootstrap(self, views, Rt0, Rt1, pm01: PairMatches):
        """State when the reference's loop starts (sfm.py:304-339) for a given second pose Rt1 — from
        two_view_init (the reference's E-matrix / recoverPose initialisation) or from the scene (as in oracle.cvpath)."""
        torch = self.torch
        P1, P2 = self.K @ Rt0, self.K @ Rt1
        M = pm01.n
        X = self._triangulate(P1, P2, pm01.pts_q, pm01.pts_t, M, 1)                  # (M,4), w == 1
        errs = torch.zeros((2,), dtype=torch.float64, device=self.dev)
        X3 = torch.empty((max(M, 1), 3), dtype=torch.float32, device=self.dev)
        self._reproj(X, 2, pm01.pts_t, M, Rt1, errs[0:1], X3)
        ok, rvec, tvec, k, inl = self._pnp(X3, pm01.pts_t, M)
        pts1 = self._gather(pm01.pts_t, 2, inl, k)
        pts3d = self._gather(X3, 3, inl, k)
        self.state = dict(P1=P1, P2=P2, pts1=pts1, n1=k, points_3d=pts3d, prev_pair=None, err0=errs)
        return self.state

    @_on_ctx_stream
    def register(self, pm_prev: PairMatches | None, pm: PairMatches, first: bool):
        """One iteration of sfm.py:341-409.  pm = matches (prev view -> new view); pm_prev = the
        previous pair's matches (re-triangulated when not first, sfm.py:348-352)."""
        torch = self.torch
        st = self.state
        P1, P2 = st["P1"], st["P2"]
        if first:
            pts1, n1, points_3d = st["pts1"], st["n1"], st["points_3d"]
        else:
            n1 = pm_prev.n
            pts1 = pm_prev.pts_t
            points_3d = self._triangulate(P1, P2, pm_prev.pts_q, pm_prev.pts_t, n1, 2)     # (n1,3)
        M = pm.n
        # data association (sfm.py:356) and the complement (new points)
        i1 = torch.empty((max(n1, 1),), dtype=torch.int32, device=self.dev)
        i2 = torch.empty((max(n1, 1),), dtype=torch.int32, device=self.dev)
        keep = torch.empty((max(M, 1),), dtype=torch.uint8, device=self.dev)
        cnt = torch.empty((2,), dtype=torch.int32, device=self.dev)
        check(lib.sfm_common_points(self.ctx._h, pts1.data_ptr(), n1, pm.pts_q.data_ptr(), M, i1.data_ptr(),
                                    i2.data_ptr(), cnt[0:1].data_ptr(), keep.data_ptr()))
        temp1 = torch.empty((max(M, 1), 2), dtype=torch.float32, device=self.dev)
        temp2 = torch.empty((max(M, 1), 2), dtype=torch.float32, device=self.dev)
        check(lib.sfm_compact_pairs(self.ctx._h, pm.pts_q.data_ptr(), pm.pts_t.data_ptr(), keep.data_ptr(), M,
                                    temp1.data_ptr(), temp2.data_ptr(), cnt[1:2].data_ptr()))
        self.ctx.sync()
        c, m = (int(v) for v in cnt.cpu().numpy())
        if c < 6:          # the same rule as the library's loop (chain.cu): the minimal problems draw 5 of at least 6 points
            raise _e.error(-1, f"registration failed: the new view shares {c} points with the model (6 needed)")
        X_common = self._gather(points_3d, 3, i1, c)
        com2 = self._gather(pm.pts_t, 2, i2, c)
        ok, rvec, tvec, k, inl = self._pnp(X_common, com2, c)                           # sfm.py:362
        if not ok:
            raise _e.error(-1, "registration failed: solvePnPRansac found no consensus")
        Rt = np.hstack([_e.rodrigues_to_matrix(rvec), tvec.reshape(3, 1)])
        Pnew = self.K @ Rt
        errs = torch.zeros((2,), dtype=torch.float64, device=self.dev)
        X_in = self._gather(X_common, 3, inl, k)
        p_in = self._gather(com2, 2, inl, k)
        self._reproj(X_in, 0, p_in, k, Rt, errs[0:1])                                   # sfm.py:368
        X_new4 = self._triangulate(P2, Pnew, temp1, temp2, m, 1)                        # sfm.py:371
        X_new = torch.empty((max(m, 1), 3), dtype=torch.float32, device=self.dev)
        self._reproj(X_new4, 2, temp2, m, Rt, errs[1:2], X_new)                         # sfm.py:372
        self.state = dict(P1=P2.copy(), P2=Pnew.copy())
        return dict(Rt=Rt, errs=errs, X_new=X_new, n_new=m, n_pnp=c, n_inl=k, n_match=M)

    @_on_ctx_stream
    def run(self, views, Rt0, Rt1, matches=None):
        """Register views[2:] sequentially.  `views` are DeviceView.  Returns per-view dicts with host
        values (Rt, err_pnp, err_new, counts) and device X_new.  The loop itself runs in the library
        (sfm_chain_run); the Python loop below is kept for the parity mode in which the caller supplies
        the minimal solutions (hypothesis_fn)."""
        V = len(views)
        if matches is None:
            matches = self.match_pairs(views, [(i, i + 1) for i in range(V - 1)])
        if self.hypothesis_fn is None and V >= 3:
            return self._run_native(matches, Rt0, Rt1)
        self.bootstrap(views, Rt0, Rt1, matches[0])
        outs = []
        for i in range(V - 2):
            outs.append(self.register(matches[i] if i > 0 else None, matches[i + 1], first=(i == 0)))
        self.ctx.sync()
        if outs:
            errs = self.torch.stack([o["errs"] for o in outs]).cpu().numpy()
            for o, e in zip(outs, errs):
                o["err_pnp"], o["err_new"] = float(e[0]), float(e[1])
        return outs

    def _run_native(self, matches, Rt0, Rt1):
        nc = NativeChain(self.ctx, self.K, Rt0, Rt1, max(pm.n for pm in matches))
        try:
            return nc.extend(matches)
        finally:
            nc.close()


class NativeChain:
    """sfm_chain_create / sfm_chain_extend / sfm_chain_destroy: the per-view loop running in the library,
    fed with consecutive pairs' matches (PairMatches) as they become available."""

    def __init__(self, ctx: _e.Context, K, Rt0, Rt1, max_matches: int):
        self.ctx = ctx
        self._h = C.c_void_p()
        K = np.ascontiguousarray(K, np.float64)
        Rt0 = np.ascontiguousarray(Rt0, np.float64)
        Rt1 = np.ascontiguousarray(Rt1, np.float64)
        check(lib.sfm_chain_create(ctx._h, _e._dptr(K), _e._dptr(Rt0), _e._dptr(Rt1), int(max(max_matches, 6)), C.byref(self._h)))
        self._alive = []          # match arrays of the last pair must outlive the next extend()
        self._fed = 0
        ctx._children.add(self)

    def launch(self, matches):
        """Queue the registration of the views these consecutive pairs add (sfm_chain_extend_async) and return;
        collect() waits and returns the records.  One call in flight."""
        import torch
        n_pairs = len(matches)
        if getattr(self, "_inflight", None) is not None:
            raise _e.error(-1, "NativeChain.launch: the previous call has not been collected")
        if n_pairs == 0:
            return
        first = self._fed == 0
        n = np.array([pm.n for pm in matches], np.int32)
        reg_n = n[1:] if first else n                      # matches of the pairs that register a view
        cap = np.maximum(reg_n, 1).astype(np.int64)
        off = np.concatenate([[0], np.cumsum(cap)])
        with torch.cuda.stream(self.ctx.torch_stream()):
            X_all = torch.empty((int(max(off[-1], 1)), 3), dtype=torch.float32, device=self.ctx.torch_device)
        pq = np.array([pm.ptr_q for pm in matches], np.uint64)
        pt = np.array([pm.ptr_t for pm in matches], np.uint64)
        xn = np.ascontiguousarray(X_all.data_ptr() + off[:-1] * 12, np.uint64)
        check(lib.sfm_chain_extend_async(self._h, n_pairs, pq.ctypes.data, pt.ctypes.data, n.ctypes.data, xn.ctypes.data))
        # the queued call still re-triangulates the PREVIOUS call's last pair: it stays referenced until this call's
        # collect(), together with this call's last pair and output slab
        self._alive = [getattr(self, "_last_pair", None), matches[-1], X_all]
        self._last_pair = matches[-1]
        self._fed += n_pairs
        self._inflight = (X_all, off, len(reg_n))

    def collect(self):
        if getattr(self, "_inflight", None) is None:
            return []
        X_all, off, n_views = self._inflight
        self._inflight = None
        out = np.zeros((max(n_views, 1),), dtype=VIEW_OUT)
        nreg = C.c_int32(0)
        check(lib.sfm_chain_collect(self._h, out.ctypes.data, C.byref(nreg)))
        outs = []
        for v in range(nreg.value):
            o = out[v]
            outs.append(dict(Rt=o["Rt"].reshape(3, 4).copy(), X_new=X_all[int(off[v]):int(off[v + 1])], n_new=int(o["n_new"]),
                             n_pnp=int(o["n_pnp"]), n_inl=int(o["n_inl"]), n_match=int(o["n_match"]),
                             err_pnp=float(o["err_pnp"]), err_new=float(o["err_new"]),
                             _slab=(X_all, int(off[v]))))       # fetch_clouds copies a call's points back in one piece
        return outs

    def extend(self, matches):
        self.launch(matches)
        return self.collect()

    def close(self):
        if getattr(self, "_h", None):
            lib.sfm_chain_destroy(self._h)
            self._h = None
        self._alive = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fetch_clouds(ctx: _e.Context, outs):
    """New 3-D points of every registered view as host arrays: one device->host copy per sfm_chain_extend call (the
    points of a call are one slab), not one per view."""
    import torch
    slabs, host = {}, {}
    for o in outs:
        if "_slab" in o:
            slabs[id(o["_slab"][0])] = o["_slab"][0]
    with torch.cuda.stream(ctx.torch_stream()):
        for k, t in slabs.items():
            host[k] = t.to("cpu", non_blocking=True)
    ctx.sync()
    clouds = []
    for o in outs:
        if "_slab" in o:
            t, lo = o["_slab"]
            clouds.append(host[id(t)][lo:lo + o["n_new"]].numpy())
        else:
            clouds.append(o["X_new"][:o["n_new"]].cpu().numpy())
    return clouds


def _chunk_bounds(V: int, edges):
    edges = sorted(set([0] + [e for e in edges if 0 < e < V] + [V]))
    return list(zip(edges[:-1], edges[1:]))


def _register_pipelined(ctx: _e.Context, K, kp_d, des_d, Rt0, Rt1, bounds, events, ratio: float, nmax: int,
                        after_chunk=None):
    """Chunks of device-resident views -> registered views.  Matching runs on a second context (own stream,
    workspace and descriptor pool) so that a chunk is prepared and matched — including the host reads of its match
    counts — while the previous chunk's views register on the main context; the registration loop is launched
    asynchronously (sfm_chain_extend_async) and collected one chunk later."""
    mctx = getattr(ctx, "_match_ctx", None)
    if mctx is None:
        mctx = ctx._match_ctx = _e.Context(ctx.device)
        if ctx_profiling(ctx):
            mctx.set_profiling(True)
    ms = mctx.torch_stream()
    chain = RegistrationChain(mctx, K, ratio=ratio)
    native = NativeChain(ctx, K, Rt0, Rt1, nmax)
    views, outs, keep = [], [], []
    try:
        for k, (lo, hi) in enumerate(bounds):
            if events is not None:
                ms.wait_event(events[k])
            views += DeviceView.batch(mctx, kp_d[lo:hi], des_d[lo:hi])
            pairs = [(i, i + 1) for i in range(max(lo - 1, 0), hi - 1)]
            if not pairs:
                continue
            matches = chain.match_pairs(views, pairs)     # synchronises the matching context: survivors are in HBM
            keep.append(matches)
            outs += native.collect()                      # previous chunk's views
            native.launch(matches)
            if after_chunk is not None:
                after_chunk(k)
        outs += native.collect()
        ctx.sync()
    finally:
        native.close()
    return outs, keep


def ctx_profiling(ctx: _e.Context) -> bool:
    return bool(getattr(ctx, "_profiling_on", False))


def register_device(ctx: _e.Context, K, kps, dess, Rt0, Rt1, ratio: float = 0.70):
    """Device-resident views (torch CUDA tensors: keypoints (n,2) f32, descriptors (n,128)) -> registered views.
    A short first chunk starts the registration loop; the remaining pairs are matched while it runs.  Boundaries
    measured on the 200-view x 5000 set (ms per step): (8) 37.1, (16) 36.5, (8,25) 35.7, (12,60) 34.5, (16,64,128) 34.7,
    (12,50,100,150) 34.8, (10,40,80,120,160) 35.2 — a chunk boundary costs a host round trip, a long match launch
    running beside the loop slows it."""
    import os
    V = len(kps)
    edges = os.environ.get("SFM_REGISTER_EDGES")          # tuning aid: comma-separated chunk boundaries
    bounds = _chunk_bounds(V, [int(e) for e in edges.split(",")] if edges else (12, 60))
    outs, keep = _register_pipelined(ctx, K, kps, dess, Rt0, Rt1, bounds, None, ratio, max(int(k.shape[0]) for k in kps))
    for o in outs:
        o["_keep"] = keep
    return outs


def register_host(ctx: _e.Context, K, kps, dess, Rt0, Rt1, chunk: int = 25, ratio: float = 0.70, ahead: int = 2,
                  edges=None):
    """Host arrays in, registered views out — the call a user makes with a sequence of views whose keypoints
    (n,2) and descriptors (n,128) sit in host memory (numpy arrays or torch CPU tensors; pinned memory lets the
    upload overlap).  Views are uploaded on a copy stream in chunks (a short first chunk gets the loop going);
    descriptor preparation and the batched match of a chunk's pairs run on a matching context, the registration of
    its views on the engine stream, while later chunks are still crossing PCIe.  The copies of chunk k + `ahead` are
    submitted only after chunk k's registration has been queued, so the first chunk's work does not wait for the host
    to enqueue every copy of the sequence; a chunk's arrays land in one device slab per kind and are queued by one
    native call (sfm_upload_batch).  What separates this from the resident driver (35.8 against 32.8 ms on 200 x 5000):
    ~1.1 ms because the loop's latency-bound kernels run slower while 520 MB cross PCIe (tools/prof_e2e_contention.py),
    the first chunk's upload, the read-back of the clouds, and the extra chunk boundaries."""
    import torch
    V = len(kps)
    dev = ctx.torch_device
    as_t = lambda a: a if _e._is_torch(a) else torch.from_numpy(np.ascontiguousarray(a))
    kps = [as_t(a) for a in kps]
    dess = [as_t(a) for a in dess]
    cs = getattr(ctx, "_copy_stream", None)          # one upload stream per context: torch's caching allocator keeps a
    if cs is None:                                    # pool per stream, a fresh stream per call would cudaMalloc every buffer
        cs = ctx._copy_stream = torch.cuda.Stream(device=dev)
    import os
    env = os.environ.get("SFM_HOST_EDGES")              # tuning aid: comma-separated chunk boundaries
    if env:
        edges = [int(e) for e in env.split(",")]
    # chunk k + 1 must cross PCIe (and be matched) while chunk k registers: ~60 us per view of upload against ~166 us per
    # view of registration lets the chunks grow threefold; a boundary costs a host round trip.  Measured on 200 x 5000
    # (ms per step, tools/prof_e2e.py): (8,25,50,...,175) 36.6, (12,60) 36.5, (8,40,100) 36.3, (8,32,96) 35.8
    if edges is None:
        edges = [8, 32, 96] + list(range(224, V, 128)) if chunk == 25 else [min(8, chunk), chunk] + list(range(2 * chunk, V, chunk))
    bounds = _chunk_bounds(V, list(edges))
    kp_d, des_d, events = [None] * V, [None] * V, [None] * len(bounds)

    def upload(k):
        if k >= len(bounds) or events[k] is not None:
            return
        lo, hi = bounds[k]
        with torch.cuda.stream(cs):
            for src, dst in ((kps, kp_d), (dess, des_d)):
                part = src[lo:hi]
                t0 = part[0]
                if all((not t.is_cuda) and t.is_contiguous() and t.dtype == t0.dtype and t.shape[1:] == t0.shape[1:] for t in part):
                    # one device slab per chunk, every view a row range of it; the copies are queued by ONE native call
                    # (sfm_upload_batch) instead of one interpreter round trip per array
                    rows = [int(t.shape[0]) for t in part]
                    slab = torch.empty((sum(rows),) + tuple(t0.shape[1:]), dtype=t0.dtype, device=dev)
                    row_bytes = slab.element_size() * int(np.prod(t0.shape[1:], dtype=np.int64))
                    offs = np.concatenate([[0], np.cumsum(rows)])
                    nbytes = np.asarray(rows, np.int64) * row_bytes
                    dptr = (slab.data_ptr() + offs[:-1] * row_bytes).astype(np.uint64)
                    sptr = np.fromiter((t.data_ptr() for t in part), np.uint64, len(part))
                    _e.check(_e.lib.sfm_upload_batch(ctx._h, int(cs.cuda_stream), len(part), _e._dptr(dptr), _e._dptr(sptr), _e._dptr(nbytes)))
                    for j, i in enumerate(range(lo, hi)):
                        dst[i] = slab[int(offs[j]):int(offs[j + 1])]
                else:
                    for i in range(lo, hi):
                        dst[i] = src[i].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        events[k] = ev

    for k in range(len(bounds)):                      # chunks without pairs never reach after_chunk: upload them up front
        lo, hi = bounds[k]
        if k < ahead or hi - 1 <= max(lo - 1, 0):
            upload(k)
    outs, keep = _register_pipelined(ctx, K, kp_d, des_d, Rt0, Rt1, bounds, events, ratio, max(int(k.shape[0]) for k in kps),
                                     after_chunk=lambda k: upload(k + ahead))
    for o in outs:
        o["_keep"] = (kp_d, des_d, keep)      # uploaded on the copy stream: stay referenced until the caller drops the result
    return outs


def register_chain(scene, ctx: _e.Context | None = None, n_views: int | None = None, hypothesis_fn=None,
                   des_dtype=np.float32):
    """Host-array entry (the call a user makes): scene dict as sfm_mvs_b200.synth.orbit_scene produces —
    K and per view kp (n,2) f32, des (n,128), plus the two bootstrap poses.  Uploads, matches, registers."""
    from .cv2_compat import default_context
    ctx = ctx or default_context()
    views = scene["views"] if n_views is None else scene["views"][:n_views]
    Rt0 = np.hstack([views[0]["R"], views[0]["t"].reshape(3, 1)])
    Rt1 = np.hstack([views[1]["R"], views[1]["t"].reshape(3, 1)])
    import torch
    if hypothesis_fn is None and len(views) >= 3:
        outs = register_host(ctx, scene["K"], [v["kp"] for v in views],
                             [v["des"].astype(des_dtype, copy=False) for v in views], Rt0, Rt1)
    else:
        dviews = [DeviceView(ctx, v["kp"], v["des"].astype(des_dtype, copy=False)) for v in views]
        chain = RegistrationChain(ctx, scene["K"], hypothesis_fn=hypothesis_fn)
        outs = chain.run(dviews, Rt0, Rt1)
    clouds = fetch_clouds(ctx, outs)
    for o, c in zip(outs, clouds):
        o["X_new"] = c
        o.pop("_keep", None)
        o.pop("_slab", None)
    return outs


def two_view_init(pts0, pts1, K, Rt0=None, ctx=None):
    """The reference's two-view initialisation, sfm.py:307-316, on the engine: findEssentialMat(RANSAC, 0.999, 0.4)
    -> keep mask == 1 -> recoverPose -> keep mask > 0 -> second pose composed onto the first
    (R1 = R R0, t1 = t0 + R0 t, exactly as sfm.py:314-315 writes it).
    -> dict(E, Rt0, Rt1, pts0, pts1 (the surviving correspondences), n_essential, n_pose)."""
    from .cv2_compat import RANSAC, findEssentialMat, recoverPose
    K = np.asarray(K, np.float64)
    Rt0 = np.hstack([np.eye(3), np.zeros((3, 1))]) if Rt0 is None else np.asarray(Rt0, np.float64)
    pts0, pts1 = np.asarray(pts0), np.asarray(pts1)
    E, mask = findEssentialMat(pts0, pts1, K, method=RANSAC, prob=0.999, threshold=0.4, mask=None, ctx=ctx)
    if E is None:
        raise _e.error(-1, "two_view_init: findEssentialMat found no model")
    a, b = pts0[mask.ravel() == 1], pts1[mask.ravel() == 1]
    n_e = len(a)
    _, R, t, mask2 = recoverPose(E[:3], a, b, K, ctx=ctx)
    a, b = a[mask2.ravel() > 0], b[mask2.ravel() > 0]
    Rt1 = np.empty((3, 4))
    Rt1[:3, :3] = R @ Rt0[:3, :3]
    Rt1[:3, 3] = Rt0[:3, 3] + Rt0[:3, :3] @ t.ravel()
    return dict(E=E[:3], Rt0=Rt0, Rt1=Rt1, pts0=a, pts1=b, n_essential=n_e, n_pose=len(a))


def pairwise_init(ctx: _e.Context, views, K, pairs=None, ratio: float = 0.70, batched: bool = False):
    """isfm.py:68-87 — every earlier view j against the new view i: 2-NN + ratio matches (ALL pairs in one batched
    K1 launch), then per pair findEssentialMat(RANSAC, 0.999, 0.4) -> mask -> recoverPose -> mask, as the reference
    does it.  `views` are DeviceView; `pairs` defaults to all (j, i), j < i, in the reference's order (i outer).
    -> list of dict(pair, n_match, n_essential, n_pose (what isfm.py prints), R, t, pts0, pts1).
    batched=True (EXPERIMENTAL, not yet run on a GPU) estimates all essential matrices with
    sfm_find_essential_mat_batched instead of one call per pair."""
    V = len(views)
    if pairs is None:
        pairs = [(j, i) for i in range(V) for j in range(i)]
    chain = RegistrationChain(ctx, K, ratio=ratio)
    matches = chain.match_pairs(views, pairs)
    Kc = np.ascontiguousarray(K, np.float64)
    pre = None
    if batched and matches:
        import torch
        P = len(matches)
        n = np.array([pm.n for pm in matches], np.int32)
        a1 = np.array([pm.ptr_q for pm in matches], np.uint64)
        a2 = np.array([pm.ptr_t for pm in matches], np.uint64)
        with torch.cuda.stream(ctx.torch_stream()):
            mask_all = torch.zeros((int(max(n.sum(), 1)),), dtype=torch.uint8, device=ctx.torch_device)
        off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
        am = np.ascontiguousarray(mask_all.data_ptr() + off[:-1], np.uint64)
        E = np.zeros((P, 3, 3))
        info = np.zeros((P, 6), np.int32)
        check(lib.sfm_find_essential_mat_batched(ctx._h, P, a1.ctypes.data, a2.ctypes.data, n.ctypes.data, _e._dptr(Kc),
                                                 0.999, 0.4, 1000, _e._dptr(E), am.ctypes.data, _e._dptr(info)))
        ctx.sync()
        pre = (E, info, mask_all.cpu().numpy(), off)
    out = []
    for idx, ((j, i), pm) in enumerate(zip(pairs, matches)):
        p0 = pm.pts_q[:pm.n].cpu().numpy()
        p1 = pm.pts_t[:pm.n].cpu().numpy()
        rec = dict(pair=(j, i), n_match=pm.n, n_essential=0, n_pose=0, R=None, t=None, pts0=p0[:0], pts1=p1[:0])
        if pre is not None and pm.n >= 6:
            E, info, mask_all, off = pre
            if info[idx, 0] == 1:
                from .cv2_compat import recoverPose
                m = mask_all[off[idx]:off[idx + 1]] == 1
                a, b = p0[m], p1[m]
                _, R, t, m2 = recoverPose(E[idx], a, b, Kc, ctx=ctx)
                a, b = a[m2.ravel() > 0], b[m2.ravel() > 0]
                rec.update(n_essential=int(m.sum()), n_pose=len(a), R=R, t=t, pts0=a, pts1=b)
        elif pm.n >= 5:
            try:
                init = two_view_init(p0, p1, K, ctx=ctx)
                rec.update(n_essential=init["n_essential"], n_pose=init["n_pose"], R=init["Rt1"][:, :3].copy(),
                           t=init["Rt1"][:, 3:].copy(), pts0=init["pts0"], pts1=init["pts1"])
            except _e.error:
                pass                      # no essential matrix for this pair (the reference would raise on E = None)
        out.append(rec)
    return out


def match_pairs_sharded(ctx: _e.Context, views, pairs, rank: int = 0, world: int = 1, ratio: float = 0.70, dist=None):
    """isfm.py:68-87's matching over `world` GPUs (one process per GPU): the pair list is split by
    sharding.shard_pairs (longest-processing-time-first on Nq * Nt), this rank matches its shard in ONE batched K1
    launch, the matches stay resident here and the per-pair survivor counts are all-reduced so that every rank knows
    them.  No other data-path collective.  `views` are DeviceView on this rank's context (every rank holds the
    descriptors of the views its pairs touch).  -> (my pair indices, their PairMatches, counts of all pairs)."""
    from . import sharding
    nv = np.array([v.n for v in views], np.float64)
    pa = np.array(pairs, np.int64).reshape(-1, 2)
    mine = sharding.shard_pairs(pairs, nv[pa[:, 0]] * nv[pa[:, 1]], world)[rank]
    chain = RegistrationChain(ctx, np.eye(3), ratio=ratio)
    matches = chain.match_pairs(views, [pairs[k] for k in mine])
    counts = sharding.gather_pair_counts([pm.n for pm in matches], mine, len(pairs), dist=dist, device=ctx.torch_device)
    return mine, matches, counts


@_on_ctx_stream_fn
def match_rows_split(ctx: _e.Context, q_des, t_des, rank: int = 0, world: int = 1, ratio: float = 0.70):
    """One large pair split by QUERY rows over `world` GPUs (BASELINE configs[4], strong scaling): a query row's two
    nearest train rows do not depend on the other query rows, so rank r matches rows [lo, hi) of q_des (CUDA tensor
    (nq,128)) against the whole of t_des and keeps its block of the result; no collective.
    -> dict(lo, hi, idx (hi-lo,2), dist, good, n_good) with device tensors."""
    from . import sharding
    lo, hi = sharding.split_rows(int(q_des.shape[0]), world)[rank]
    if hi <= lo:
        return dict(lo=lo, hi=hi, idx=None, dist=None, good=None, n_good=0, q=None, t=None)
    dq = ctx.descriptors(q_des[lo:hi])
    dt = ctx.descriptors(t_des)
    idx, dist_, good, ng = ctx.knn2(dq, dt, ratio, device_out=True)
    return dict(lo=lo, hi=hi, idx=idx, dist=dist_, good=good, n_good=ng, q=dq, t=dt)
