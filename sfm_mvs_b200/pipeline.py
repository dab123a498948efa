"""The reference's per-view registration loop (sfm.py:341-409) on the engine, device-resident.

One *registered view* = one iteration of that loop with imread / SIFT / GUI removed:
match(prev, new) + ratio test, re-triangulation of the previous pair's matches, data association
(common_points), PnP-RANSAC, reprojection error, triangulation of the new points, reprojection
error.  Keypoints, descriptors, matches, 3-D points and masks live in HBM (torch CUDA tensors are
used as plain device buffers); per view the host learns three integers (match / association
counts) and the pose.  Matching does not depend on poses, so all consecutive pairs are matched up
front in one stream-ordered batch (this is also the unit that shards over GPUs, isfm.py:68-87).
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


class DeviceView:
    """A view's keypoints and prepared descriptors resident in HBM."""

    def __init__(self, ctx: _e.Context, kp, des):
        import torch
        self.ctx = ctx
        self.n = int(len(kp))
        with torch.cuda.stream(ctx.torch_stream()):
            if _e._is_torch(kp):
                self.kp = kp.to(device=ctx.torch_device, dtype=torch.float32).contiguous()
            else:
                self.kp = torch.from_numpy(np.ascontiguousarray(kp, np.float32)).to(ctx.torch_device)
        self.desc = ctx.descriptors(des)


class PairMatches:
    __slots__ = ("pts_q", "pts_t", "qidx", "tidx", "n_dev", "n")


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
        nq = np.array([views[a].n for a, _ in pairs], np.int64)
        off = np.concatenate([[0], np.cumsum(nq)])
        tot = int(max(off[-1], 1))
        counts = torch.empty((n,), dtype=torch.int32, device=self.dev)
        pts_q = torch.empty((tot, 2), dtype=torch.float32, device=self.dev)
        pts_t = torch.empty((tot, 2), dtype=torch.float32, device=self.dev)
        qidx = torch.empty((tot,), dtype=torch.int32, device=self.dev)
        tidx = torch.empty((tot,), dtype=torch.int32, device=self.dev)

        def ptrs(base, stride):
            return np.ascontiguousarray(base + off[:-1] * stride, np.uint64)

        hq = np.array([views[a].desc._h.value if hasattr(views[a].desc._h, "value") else views[a].desc._h for a, _ in pairs], np.uint64)
        ht = np.array([views[b].desc._h.value if hasattr(views[b].desc._h, "value") else views[b].desc._h for _, b in pairs], np.uint64)
        kq = np.array([views[a].kp.data_ptr() for a, _ in pairs], np.uint64)
        kt = np.array([views[b].kp.data_ptr() for _, b in pairs], np.uint64)
        a_pq, a_pt = ptrs(pts_q.data_ptr(), 8), ptrs(pts_t.data_ptr(), 8)
        a_qi, a_ti = ptrs(qidx.data_ptr(), 4), ptrs(tidx.data_ptr(), 4)
        check(lib.sfm_desc_match_gather_batched(self.ctx._h, n, hq.ctypes.data, ht.ctypes.data, self.ratio, kq.ctypes.data,
                                                kt.ctypes.data, None, None, a_pq.ctypes.data, a_pt.ctypes.data,
                                                a_qi.ctypes.data, a_ti.ctypes.data, counts.data_ptr()))
        self.ctx.sync()
        host = counts.cpu().numpy()
        out = []
        for k in range(n):
            pm = PairMatches()
            lo, hi = int(off[k]), int(off[k + 1])
            pm.pts_q, pm.pts_t, pm.qidx, pm.tidx = pts_q[lo:hi], pts_t[lo:hi], qidx[lo:hi], tidx[lo:hi]
            pm.n_dev = counts[k:k + 1]
            pm.n = int(host[k])
            out.append(pm)
        return out

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
        """State when the reference's loop starts (sfm.py:304-339); the E-matrix / recoverPose
        initialisation is out of scope, the second pose is given (as in oracle.cvpath)."""
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
        values (Rt, err_pnp, err_new, counts) and device X_new."""
        V = len(views)
        if matches is None:
            matches = self.match_pairs(views, [(i, i + 1) for i in range(V - 1)])
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


def register_chain(scene, ctx: _e.Context | None = None, n_views: int | None = None, hypothesis_fn=None,
                   des_dtype=np.float32):
    """Host-array entry (the call a user makes): scene dict as sfm_mvs_b200.synth.orbit_scene produces —
    K and per view kp (n,2) f32, des (n,128), plus the two bootstrap poses.  Uploads, matches, registers."""
    from .cv2_compat import default_context
    ctx = ctx or default_context()
    views = scene["views"] if n_views is None else scene["views"][:n_views]
    dviews = [DeviceView(ctx, v["kp"], v["des"].astype(des_dtype, copy=False)) for v in views]
    Rt0 = np.hstack([views[0]["R"], views[0]["t"].reshape(3, 1)])
    Rt1 = np.hstack([views[1]["R"], views[1]["t"].reshape(3, 1)])
    chain = RegistrationChain(ctx, scene["K"], hypothesis_fn=hypothesis_fn)
    outs = chain.run(dviews, Rt0, Rt1)
    import torch
    with torch.cuda.stream(ctx.torch_stream()):
        for o in outs:
            o["X_new"] = o["X_new"][:o["n_new"]].cpu().numpy()
    return outs
