"""Array-level host API over the C ABI (include/sfm_b200.h).

Every method takes NumPy arrays (host buffers; the call is synchronous like the cv2 call it
replaces) or torch CUDA tensors (device buffers; the call only enqueues kernels on the context's
stream and results are returned as torch CUDA tensors).  Torch is used for device memory and
streams only.  There is no CPU path: constructing a Context without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C

import weakref

import numpy as np

from . import _lib
from ._lib import BaStats, PnpInfo, check, error, lib

KERNEL_IDS = {lib.sfm_kernel_name(i).decode(): i for i in range(17)}


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


class _Arg:
    """A contiguous input buffer: keeps the owner alive and exposes its raw pointer."""

    __slots__ = ("owner", "ptr", "device", "shape")

    def __init__(self, x, dtype, allow_none=False):
        if x is None:
            if not allow_none:
                raise error(-1, "missing array argument")
            self.owner, self.ptr, self.device, self.shape = None, None, False, ()
            return
        if _is_torch(x):
            import torch
            want = {np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32,
                    np.uint8: torch.uint8}[dtype]
            if x.dtype != want:
                raise error(-1, f"device tensor has dtype {x.dtype}, expected {want}")
            x = x.contiguous()
            self.owner, self.ptr, self.device, self.shape = x, x.data_ptr(), x.is_cuda, tuple(x.shape)
            if not x.is_cuda:
                self.ptr = x.data_ptr()
        else:
            a = np.ascontiguousarray(x, dtype=dtype)
            self.owner, self.ptr, self.device, self.shape = a, a.ctypes.data, False, a.shape


def _out(shape, dtype, like_device, torch_device=None):
    """Allocate an output buffer on the same side as the inputs."""
    if like_device:
        import torch
        tdt = {np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32,
               np.uint8: torch.uint8}[dtype]
        t = torch.empty(shape, dtype=tdt, device=torch_device)
        return t, t.data_ptr()
    a = np.empty(shape, dtype=dtype)
    return a, a.ctypes.data


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _stream_ordered(fn):
    """Methods that accept or return torch CUDA tensors: the kernels run on the CONTEXT's stream, torch allocates
    (and frees `.contiguous()` temporaries) on torch's CURRENT stream.  Run the call with the context's stream
    current, after everything the caller has queued so far, and make the caller's stream wait for it — inputs written
    on the caller's stream are complete before the kernels read them, outputs are complete before the caller's next
    operation, and the caching allocator sees every buffer on the stream that uses it."""
    import functools
    import sys

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        torch = sys.modules.get("torch")
        if torch is None or not torch.cuda.is_initialized():
            return fn(self, *a, **k)
        ctx = self if isinstance(self, Context) else self.ctx
        cur = torch.cuda.current_stream(ctx.torch_device)
        own = ctx.torch_stream()
        if cur == own:
            return fn(self, *a, **k)
        own.wait_stream(cur)
        with torch.cuda.stream(own):
            out = fn(self, *a, **k)
        cur.wait_stream(own)
        return out
    return wrapper


class Descriptors:
    """A view's descriptors prepared once (K1b) and resident in HBM (sfm_desc)."""

    def __init__(self, ctx: "Context", des):
        self.ctx = ctx
        if _is_torch(des):
            import torch
            dtype = 1 if des.dtype == torch.uint8 else 0
            arg = _Arg(des, np.uint8 if dtype else np.float32)
        else:
            des = np.asarray(des)
            if des.dtype == np.uint8:
                dtype = 1
            elif des.dtype == np.float32:
                dtype = 0
            else:
                # cv2: batch_distance.cpp asserts type == CV_32F || CV_8U for NORM_L2
                raise error(-1, f"knnMatch: descriptor dtype {des.dtype} not supported (float32 or uint8, as cv2)")
            arg = _Arg(des, des.dtype.type)
        if len(arg.shape) != 2:
            raise error(-1, f"descriptors must be 2-D (n, dim), got shape {arg.shape}")
        self.n, self.dim = int(arg.shape[0]), int(arg.shape[1])
        h = C.c_void_p()
        check(lib.sfm_desc_create(ctx._h, arg.ptr, dtype, self.n, self.dim, C.byref(h)))
        self._h = h
        ctx._children.add(self)

    @classmethod
    def _from_handle(cls, ctx: "Context", handle, n: int, dim: int):
        self = cls.__new__(cls)
        self.ctx, self._h, self.n, self.dim = ctx, handle, int(n), int(dim)
        ctx._children.add(self)
        return self

    @classmethod
    def batch(cls, ctx: "Context", sets):
        """Descriptors of many views (CUDA tensors, (n,128) float32 or uint8, one dtype) prepared by ONE K1b launch
        (sfm_desc_create_batched).  Anything else falls back to one call per set."""
        import torch
        ok = len(sets) > 0 and all(_is_torch(d) and d.is_cuda and d.dim() == 2 and d.shape[1] == 128 and d.shape[0] > 0
                                   and d.is_contiguous() for d in sets)
        ok = ok and len({d.dtype for d in sets}) == 1 and sets[0].dtype in (torch.float32, torch.uint8)
        if not ok:
            return [cls(ctx, d) for d in sets]
        count = len(sets)
        ptrs = np.array([d.data_ptr() for d in sets], np.uint64)
        ns = np.array([d.shape[0] for d in sets], np.int32)
        out = np.zeros((count,), np.uint64)
        check(lib.sfm_desc_create_batched(ctx._h, count, ptrs.ctypes.data, 1 if sets[0].dtype == torch.uint8 else 0,
                                          ns.ctypes.data, 128, out.ctypes.data))
        return [cls._from_handle(ctx, C.c_void_p(int(h)), int(k), 128) for h, k in zip(out, ns)]

    @property
    def exact(self) -> bool:
        return bool(lib.sfm_desc_is_exact(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib.sfm_desc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One engine context = one CUDA device + one stream (sfm_ctx)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = C.c_void_p()
        check(lib.sfm_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self.device = int(device)
        self.sm_count = lib.sfm_ctx_sm_count(h)
        # descriptor sets, BA problems and chains return their buffers to pools of the context when they are destroyed:
        # an explicit close() of the context closes what is still alive on it first
        self._children = weakref.WeakSet()

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            for child in list(getattr(self, "_children", ())):
                try:
                    child.close()
                except Exception:
                    pass
            for name in ("_match_ctx",):
                sub = getattr(self, name, None)
                if sub is not None:
                    sub.close()
            lib.sfm_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib.sfm_ctx_sync(self._h))

    @property
    def stream(self) -> int:
        return lib.sfm_ctx_stream(self._h) or 0

    @property
    def torch_device(self):
        import torch
        return torch.device("cuda", self.device)

    def torch_stream(self):
        """The context's stream as a torch ExternalStream.  The methods that take or return CUDA tensors order
        themselves against torch's current stream (_stream_ordered); code that calls the C ABI directly with torch
        buffers (pipeline.py) runs under `torch.cuda.stream(ctx.torch_stream())`."""
        import torch
        if getattr(self, "_tstream", None) is None:
            self._tstream = torch.cuda.ExternalStream(self.stream, device=self.torch_device)
            # torch objects tied to this stream (pinned host buffers of non_blocking copies record an event on it
            # when they are freed) may outlive the context: the stream is never destroyed once torch has seen it
            check(lib.sfm_ctx_detach_stream(self._h))
        return self._tstream

    def _aux(self):
        """Auxiliary contexts working for this one (pipeline.register_*: the matching context); their launches and
        kernel times are reported with this context's."""
        m = getattr(self, "_match_ctx", None)
        return [m] if m is not None and getattr(m, "_h", None) else []

    def launch_count(self) -> int:
        return int(lib.sfm_ctx_launch_count(self._h)) + sum(a.launch_count() for a in self._aux())

    def set_profiling(self, on: bool):
        check(lib.sfm_ctx_set_profiling(self._h, 1 if on else 0))
        self._profiling_on = bool(on)
        for a in self._aux():
            a.set_profiling(on)

    def reset_profile(self):
        check(lib.sfm_ctx_reset_profile(self._h))
        for a in self._aux():
            a.reset_profile()

    def profile(self) -> dict:
        out = {}
        for name, i in KERNEL_IDS.items():
            ms, n = C.c_double(), C.c_int64()
            check(lib.sfm_ctx_get_profile(self._h, i, C.byref(ms), C.byref(n)))
            if n.value:
                out[name] = dict(ms=ms.value, launches=n.value)
        for a in self._aux():
            for name, p in a.profile().items():
                q = out.setdefault(name, dict(ms=0.0, launches=0))
                q["ms"] += p["ms"]
                q["launches"] += p["launches"]
        return out

    # ------------------------------------------------------------------ hot path 1: matching
    def descriptors(self, des) -> Descriptors:
        return des if isinstance(des, Descriptors) else Descriptors(self, des)

    @_stream_ordered
    def knn2(self, q, t, ratio: float = 0.70, mode: int = 0, device_out: bool = False):
        """2-NN under L2 + Lowe ratio (sfm.py:259-265).  Returns idx (nq,2) i32, dist (nq,2) f32,
        good (nq,) uint8, n_good.  idx=-1 / dist=inf mark missing neighbours (nt < 2)."""
        dq, dt = self.descriptors(q), self.descriptors(t)
        if dq.dim != dt.dim:
            raise error(-1, f"knnMatch: descriptor dims differ ({dq.dim} vs {dt.dim})")
        nq = dq.n
        idx, pidx = _out((nq, 2), np.int32, device_out, self.torch_device if device_out else None)
        dist, pdist = _out((nq, 2), np.float32, device_out, self.torch_device if device_out else None)
        good, pgood = _out((nq,), np.uint8, device_out, self.torch_device if device_out else None)
        ng, png = _out((1,), np.int32, device_out, self.torch_device if device_out else None)
        check(lib.sfm_desc_match(self._h, dq._h, dt._h, float(ratio), pidx, pdist, pgood, png, int(mode)))
        return idx, dist, good, (ng if device_out else int(ng[0]))

    @_stream_ordered
    def match_gather(self, idx, good, kp_q, kp_t, n_hint: int | None = None):
        """sfm.py:267-268 on the device: survivors' keypoints, ascending queryIdx.  All arguments are
        torch CUDA tensors; returns (pts_q, pts_t, qidx, tidx, n) with capacity-nq tensors."""
        import torch
        nq = int(idx.shape[0])
        dev = self.torch_device
        pts_q = torch.empty((nq, 2), dtype=torch.float32, device=dev)
        pts_t = torch.empty((nq, 2), dtype=torch.float32, device=dev)
        qi = torch.empty((nq,), dtype=torch.int32, device=dev)
        ti = torch.empty((nq,), dtype=torch.int32, device=dev)
        n = torch.empty((1,), dtype=torch.int32, device=dev)
        a_idx, a_good = _Arg(idx, np.int32), _Arg(good, np.uint8)
        a_kq, a_kt = _Arg(kp_q, np.float32), _Arg(kp_t, np.float32)
        check(lib.sfm_match_gather(self._h, a_idx.ptr, a_good.ptr, nq, a_kq.ptr, a_kt.ptr, pts_q.data_ptr(),
                                   pts_t.data_ptr(), qi.data_ptr(), ti.data_ptr(), n.data_ptr()))
        return pts_q, pts_t, qi, ti, n

    def debug_tc_accumulators(self, q: Descriptors, t: Descriptors) -> np.ndarray:
        nqt = (q.n + 127) // 128
        ntt = (t.n + 127) // 128
        out = np.empty((nqt * 128, ntt * 128), np.float32)
        check(lib.sfm_debug_match_tc_dump(self._h, q._h, t._h, _dptr(out), out.size))
        return out

    # ------------------------------------------------------------------ hot path 2: triangulation
    @_stream_ordered
    def triangulate(self, P1, P2, x1, x2, pts_layout: int = 0, out_layout: int = 0, normalize_w: bool = False):
        """cv2.triangulatePoints (sfm.py:53) [+ cloud/cloud[3] (sfm.py:54) if normalize_w].
        pts_layout 0: (2,N), 1: (N,2).  out_layout 0: (4,N), 1: (N,4), 2: (N,3)."""
        P1 = np.ascontiguousarray(P1, np.float64).reshape(12)
        P2 = np.ascontiguousarray(P2, np.float64).reshape(12)
        a1, a2 = _Arg(x1, np.float32), _Arg(x2, np.float32)
        if a1.shape != a2.shape or len(a1.shape) != 2:
            raise error(-1, f"triangulatePoints: point arrays must have equal 2-D shapes, got {a1.shape} / {a2.shape}")
        n = a1.shape[1] if pts_layout == 0 else a1.shape[0]
        if (a1.shape[0] if pts_layout == 0 else a1.shape[1]) != 2:
            raise error(-1, f"triangulatePoints: expected {'(2,N)' if pts_layout == 0 else '(N,2)'}, got {a1.shape}")
        shape = {0: (4, n), 1: (n, 4), 2: (n, 3)}[out_layout]
        dev = a1.device
        X, pX = _out(shape, np.float32, dev, self.torch_device if dev else None)
        check(lib.sfm_triangulate(self._h, _dptr(P1), _dptr(P2), a1.ptr, a2.ptr, int(n), pts_layout, pX, out_layout,
                                  1 if normalize_w else 0))
        return X

    @_stream_ordered
    def reproj_error(self, X, x_layout: int, px, px_layout: int, Rt, K, want_proj: bool = False,
                     want_X3: bool = False, device_err: bool = False):
        """ReprojectionError core (sfm.py:84-95).  x_layout 0: (N,3), 1: (4,N), 2: (N,4);
        px_layout 0: (2,N), 1: (N,2).  Returns (err, proj|None, X3|None)."""
        Rt = np.ascontiguousarray(Rt, np.float64).reshape(12)
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        aX, ap = _Arg(X, np.float32), _Arg(px, np.float32)
        n = aX.shape[1] if x_layout == 1 else aX.shape[0]
        dev = aX.device
        tdev = self.torch_device if dev else None
        proj, pproj = _out((n, 2), np.float32, dev, tdev) if want_proj else (None, None)
        X3, pX3 = _out((n, 3), np.float32, dev, tdev) if want_X3 else (None, None)
        if device_err:
            err, perr = _out((1,), np.float64, True, self.torch_device)
        else:
            err, perr = _out((1,), np.float64, False)
        check(lib.sfm_reproj_error(self._h, aX.ptr, x_layout, ap.ptr, px_layout, int(n), _dptr(Rt), _dptr(K), perr,
                                   pproj, pX3))
        return (err if device_err else float(err[0])), proj, X3

    @_stream_ordered
    def common_points(self, pts1, pts2):
        """common_points (sfm.py:215-239) association.  Returns idx1, idx2 (trimmed on the host path),
        keep2 (uint8 mask over pts2 rows never chosen), n_common."""
        a1, a2 = _Arg(pts1, np.float32), _Arg(pts2, np.float32)
        n1, n2 = int(a1.shape[0]), int(a2.shape[0])
        dev = a1.device
        tdev = self.torch_device if dev else None
        i1, p1 = _out((max(n1, 1),), np.int32, dev, tdev)
        i2, p2 = _out((max(n1, 1),), np.int32, dev, tdev)
        keep, pk = _out((max(n2, 1),), np.uint8, dev, tdev)
        nc, pn = _out((1,), np.int32, dev, tdev)
        check(lib.sfm_common_points(self._h, a1.ptr, n1, a2.ptr, n2, p1, p2, pn, pk))
        if dev:
            return i1, i2, keep[:n2], nc
        c = int(nc[0])
        return i1[:c], i2[:c], keep[:n2], c

    # ------------------------------------------------------------------ hot path 3a: PnP
    @_stream_ordered
    def pnp_score(self, X, px, K, Rt, thr: float = 8.0, want_masks: bool = True):
        """K4: inlier counts (H,) and masks (H,N) of H poses Rt (H,3,4) — PnPRansacCallback::computeError."""
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        Rt = np.ascontiguousarray(Rt, np.float64).reshape(-1, 12)
        H = Rt.shape[0]
        aX, ap = _Arg(X, np.float32), _Arg(px, np.float32)
        n = int(aX.shape[0])
        dev = aX.device
        tdev = self.torch_device if dev else None
        counts, pc = _out((H,), np.int32, dev, tdev)
        masks, pm = _out((H, n), np.uint8, dev, tdev) if want_masks else (None, None)
        check(lib.sfm_pnp_score(self._h, aX.ptr, ap.ptr, n, _dptr(K), _dptr(Rt), H, float(thr), pc, pm))
        return counts, masks

    @_stream_ordered
    def pnp_ransac(self, X, px, K, max_iters: int = 100, thr: float = 8.0, confidence: float = 0.99,
                   hypotheses=None, hyp_valid=None):
        """cv2.solvePnPRansac with OpenCV defaults (sfm.py:67).  Returns ok, rvec (3,), tvec (3,),
        inliers (n,) int32 ascending (empty when not ok), info dict.  `hypotheses` (max_iters,6)
        optionally supplies the minimal solutions (rvec|tvec) for the subsets of ransac_subsets()."""
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        aX, ap = _Arg(X, np.float32), _Arg(px, np.float32)
        n = int(aX.shape[0])
        if len(aX.shape) != 2 or aX.shape[1] != 3 or ap.shape != (n, 2):
            raise error(-1, f"solvePnPRansac: expected X (N,3) and p (N,2), got {aX.shape} / {ap.shape}")
        rvec, tvec = np.zeros(3), np.zeros(3)
        inl = np.empty(max(n, 1), np.int32)
        ni, ok = C.c_int32(0), C.c_int32(0)
        info = PnpInfo()
        if hypotheses is None:
            check(lib.sfm_pnp_ransac(self._h, aX.ptr, ap.ptr, n, _dptr(K), int(max_iters), float(thr), float(confidence),
                                     _dptr(rvec), _dptr(tvec), _dptr(inl), C.byref(ni), C.byref(ok), C.byref(info)))
        else:
            hyp = np.ascontiguousarray(hypotheses, np.float64).reshape(-1, 6)
            if hyp.shape[0] != max_iters:
                raise error(-1, f"hypotheses must have {max_iters} rows, got {hyp.shape[0]}")
            hv = None if hyp_valid is None else np.ascontiguousarray(hyp_valid, np.uint8)
            check(lib.sfm_pnp_ransac_hyp(self._h, aX.ptr, ap.ptr, n, _dptr(K), _dptr(hyp),
                                         None if hv is None else _dptr(hv), int(max_iters), float(thr),
                                         float(confidence), _dptr(rvec), _dptr(tvec), _dptr(inl), C.byref(ni),
                                         C.byref(ok), C.byref(info)))
        d = dict(iters_run=info.iters_run, best_iter=info.best_iter, hyp_solved=info.hyp_solved,
                 refine_iters=info.refine_iters, rvec_ransac=np.array(info.rvec_ransac[:]),
                 tvec_ransac=np.array(info.tvec_ransac[:]))
        return bool(ok.value), rvec, tvec, inl[:ni.value].copy(), d


    def ba_reference_fd(self, x, n_points: int, want_jac: bool = True):
        """OptimReprojectionError (sfm.py:104-136) at x = [Rt 12 | K 9 | p (2,N) | X (N,3)] and, if wanted, its
        forward-difference Jacobian with scipy's steps: (f0 (2N,), J (2N, 21+5N) or None), float64 host arrays."""
        x = np.ascontiguousarray(x, np.float64).ravel()
        f0 = np.empty(2 * n_points)
        J = np.empty((2 * n_points, len(x))) if want_jac else None
        check(lib.sfm_ba_reference_fd(self._h, _dptr(x), len(x), int(n_points), _dptr(f0), None if J is None else _dptr(J)))
        return f0, J

    def reduced_solve(self, S, g, method: str = "cholesky"):
        """S x = -g for a dense symmetric positive definite S (6C x 6C, only its lower triangle is read) — the reduced
        camera system solve of an LM step on its own.  method "cholesky" (csrc/solve.cu) returns (x float64, info);
        "pcg" (csrc/pcg.cu, the LM step's default) returns (x, solved, iterations)."""
        from .layout import pack_lower_blocks
        blocks = pack_lower_blocks(S)
        n = int(np.asarray(S).shape[0])
        C = n // 6
        g = np.ascontiguousarray(g, np.float32).ravel()
        x = np.empty(n)
        info = np.zeros(1, np.int32)
        if method == "pcg":
            its = np.zeros(1, np.int32)
            check(lib.sfm_reduced_solve_pcg(self._h, _dptr(blocks), _dptr(g), C, _dptr(x), _dptr(info), _dptr(its)))
            return x, bool(info[0]), int(its[0])
        check(lib.sfm_reduced_solve(self._h, _dptr(blocks), _dptr(g), C, _dptr(x), _dptr(info)))
        return x, int(info[0])

    @_stream_ordered
    def epnp_batch(self, X, px, K, subsets):
        """The device minimal solver on explicit 5-point subsets (H,5): (R (H,3,3), t (H,3)), the raw output of
        cv2.solvePnP(X[s], px[s], K, 0, flags=SOLVEPNP_EPNP) before cv2.Rodrigues — bit-identical to OpenCV."""
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        aX, ap = _Arg(X, np.float32), _Arg(px, np.float32)
        subs = np.ascontiguousarray(subsets, np.int32).reshape(-1, 5)
        out = np.empty((len(subs), 12))
        check(lib.sfm_epnp_batch(self._h, aX.ptr, ap.ptr, int(aX.shape[0]), _dptr(K), _dptr(subs), len(subs), _dptr(out)))
        return out[:, :9].reshape(-1, 3, 3).copy(), out[:, 9:].copy()


# ---------------------------------------------------------------------- host utilities (no GPU needed)
def rodrigues_to_matrix(rvec) -> np.ndarray:
    r = np.ascontiguousarray(rvec, np.float64).reshape(3)
    R = np.empty(9)
    check(lib.sfm_rodrigues_to_matrix(_dptr(r), _dptr(R)))
    return R.reshape(3, 3)


def rodrigues_to_vector(R) -> np.ndarray:
    R = np.ascontiguousarray(R, np.float64).reshape(9)
    r = np.empty(3)
    check(lib.sfm_rodrigues_to_vector(_dptr(R), _dptr(r)))
    return r


def ransac_subsets(n: int, iters: int = 100) -> np.ndarray:
    out = np.empty((iters, 5), np.int32)
    check(lib.sfm_ransac_subsets(int(n), int(iters), _dptr(out)))
    return out


def epnp(X, px, K):
    X = np.ascontiguousarray(X, np.float32).reshape(-1, 3)
    px = np.ascontiguousarray(px, np.float32).reshape(-1, 2)
    K = np.ascontiguousarray(K, np.float64).reshape(9)
    R, t = np.empty(9), np.empty(3)
    check(lib.sfm_epnp(_dptr(X), _dptr(px), len(X), _dptr(K), _dptr(R), _dptr(t)))
    return R.reshape(3, 3), t


def five_point(q1, q2):
    """Nister's five-point solver on one minimal sample of normalised coordinates (5,2) — the hypothesis kernel's
    code compiled for the host (sfm_five_point; no GPU needed).  -> (k,3,3) essential matrices, k <= 10."""
    q1 = np.ascontiguousarray(q1, np.float64).reshape(-1, 2)
    q2 = np.ascontiguousarray(q2, np.float64).reshape(-1, 2)
    if q1.shape != (5, 2) or q2.shape != (5, 2):
        raise error(-1, "five_point: needs exactly five correspondences")
    E = np.zeros((10, 3, 3))
    k = C.c_int32(0)
    check(lib.sfm_five_point(_dptr(q1), _dptr(q2), _dptr(E), C.byref(k)))
    return E[:k.value].copy()


# ---------------------------------------------------------------------- hot path 3b: bundle adjustment
class BAProblem:
    """A bundle-adjustment problem resident in HBM (sfm_ba).  Observations must be point-major."""

    def __init__(self, ctx: Context, n_cam: int, n_pt: int, cam_idx, pt_idx, obs, K, totals=None):
        self.ctx = ctx
        cam_idx = np.ascontiguousarray(cam_idx, np.int32)
        pt_idx = np.ascontiguousarray(pt_idx, np.int32)
        obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 2)
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        self.n_cam, self.n_pt, self.n_obs = int(n_cam), int(n_pt), int(len(cam_idx))
        h = C.c_void_p()
        check(lib.sfm_ba_create(ctx._h, self.n_cam, self.n_pt, self.n_obs, _dptr(cam_idx), _dptr(pt_idx), _dptr(obs),
                                _dptr(K), C.byref(h)))
        self._h = h
        ctx._children.add(self)
        if totals is not None:
            check(lib.sfm_ba_set_totals(h, int(totals[0]), int(totals[1])))

    def close(self):
        if getattr(self, "_h", None):
            lib.sfm_ba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, cams=None, pts=None):
        c = None if cams is None else np.ascontiguousarray(cams, np.float64).reshape(self.n_cam, 6)
        p = None if pts is None else np.ascontiguousarray(pts, np.float64).reshape(self.n_pt, 3)
        check(lib.sfm_ba_set_params(self._h, None if c is None else _dptr(c), None if p is None else _dptr(p)))

    def get_params(self):
        c, p = np.empty((self.n_cam, 6)), np.empty((self.n_pt, 3))
        check(lib.sfm_ba_get_params(self._h, _dptr(c), _dptr(p)))
        return c, p

    @_stream_ordered
    def eval(self, mode: int = 0, want_r=True, want_J=True, device_out: bool = False):
        """K5.  Returns dict(r, Jc, Jp, cost)."""
        O = self.n_obs
        tdev = self.ctx.torch_device if device_out else None
        r, pr = _out((O, 2) if mode != 2 else (O,), np.float32, device_out, tdev) if want_r else (None, None)
        wj = want_J and mode == 0
        Jc, pjc = _out((O, 2, 6), np.float32, device_out, tdev) if wj else (None, None)
        Jp, pjp = _out((O, 2, 3), np.float32, device_out, tdev) if wj else (None, None)
        cost, pc = _out((1,), np.float64, device_out, tdev)
        check(lib.sfm_ba_eval(self._h, int(mode), pr, pjc, pjp, pc))
        return dict(r=r, Jc=Jc, Jp=Jp, cost=cost if device_out else float(cost[0]))

    @_stream_ordered
    def eval_into(self, mode, r, Jc, Jp, cost):
        """K5 into caller-owned device tensors (the benchmarked call: nothing allocated, nothing synchronised)."""
        check(lib.sfm_ba_eval(self._h, int(mode), None if r is None else r.data_ptr(),
                              None if Jc is None else Jc.data_ptr(), None if Jp is None else Jp.data_ptr(),
                              None if cost is None else cost.data_ptr()))

    def build_system(self, lam: float = 0.0):
        check(lib.sfm_ba_build_system(self._h, float(lam)))
        n = 6 * self.n_cam
        S, g, hd = np.empty((n, n), np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
        check(lib.sfm_ba_read(self._h, 0, _dptr(S), S.size))
        check(lib.sfm_ba_read(self._h, 1, _dptr(g), n))
        check(lib.sfm_ba_read(self._h, 2, _dptr(hd), n))
        return S, g, hd

    def gn_step(self, lam: float) -> dict:
        st = BaStats()
        check(lib.sfm_ba_gn_step(self._h, float(lam), C.byref(st)))
        return dict(cost_before=st.cost_before, cost_after=st.cost_after, step_norm=st.step_norm,
                    accepted=bool(st.accepted), solve_info=st.solve_info, lambda_next=st.lambda_next)

    def solve(self, max_iters: int = 20, lam: float = 1e-3, ftol: float = 1e-8, verbose: bool = False):
        """Levenberg-Marquardt loop over gn_step.  Returns the list of per-iteration stats."""
        hist = []
        for _ in range(max_iters):
            st = self.gn_step(lam)
            hist.append(dict(st, lam=lam))
            if verbose:
                print(st)
            lam = st["lambda_next"]
            if st["accepted"] and st["cost_before"] - st["cost_after"] <= ftol * st["cost_before"]:
                break
            if not st["accepted"] and lam >= 1e12:
                break
        return hist

    # C1
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(unique_id, 128)
        check(lib.sfm_ba_comm_init(self._h, buf, int(rank), int(world)))


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib.sfm_nccl_unique_id(buf))
    return buf.raw
