"""ctypes binding of libsfm_b200.so (C ABI declared in include/sfm_b200.h).

There is no CPU implementation behind this module: if the shared library is missing the
import fails loudly, and without a CUDA device `sfm_ctx_create` returns an error that is
raised as `sfm_mvs_b200.error`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsfm_b200.so")


class error(RuntimeError):
    """Raised where the reference's cv2 call would raise cv2.error (bad shape/dtype) and for CUDA
    failures.  Carries the C-ABI status code in `.status`."""

    def __init__(self, status: int, msg: str):
        super().__init__(f"sfm_b200 error {status}: {msg}")
        self.status = status


class PnpInfo(C.Structure):
    _fields_ = [("iters_run", C.c_int32), ("best_iter", C.c_int32), ("hyp_solved", C.c_int32),
                ("refine_iters", C.c_int32), ("rvec_ransac", C.c_double * 3), ("tvec_ransac", C.c_double * 3)]


class BaStats(C.Structure):
    _fields_ = [("cost_before", C.c_double), ("cost_after", C.c_double), ("step_norm", C.c_double),
                ("grad_norm", C.c_double), ("accepted", C.c_int32), ("solve_info", C.c_int32),
                ("lambda_next", C.c_double)]


def _load() -> C.CDLL:
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C sfm_mvs_b200/csrc`). sfm_mvs_b200 has no CPU fallback.")
    return C.CDLL(LIB_PATH)


lib = _load()

_vp, _i, _i64, _d, _f = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_float
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes).  Every symbol declared in include/sfm_b200.h appears here.
PROTOTYPES = {
    "sfm_version": (_i, []),
    "sfm_last_error": (C.c_char_p, []),
    "sfm_ctx_create": (_i, [_i, _vp, _pp]),
    "sfm_ctx_destroy": (None, [_vp]),
    "sfm_ctx_sync": (_i, [_vp]),
    "sfm_ctx_stream": (_vp, [_vp]),
    "sfm_ctx_detach_stream": (_i, [_vp]),
    "sfm_ctx_sm_count": (_i, [_vp]),
    "sfm_kernel_name": (C.c_char_p, [_i]),
    "sfm_ctx_set_profiling": (_i, [_vp, _i]),
    "sfm_ctx_reset_profile": (_i, [_vp]),
    "sfm_ctx_get_profile": (_i, [_vp, _i, C.POINTER(_d), C.POINTER(_i64)]),
    "sfm_ctx_launch_count": (_i64, [_vp]),
    "sfm_knn2_l2_ratio": (_i, [_vp, _vp, _i, _vp, _i, _i, _d, _vp, _vp, _vp, _vp, _i]),
    "sfm_desc_create": (_i, [_vp, _vp, _i, _i, _i, _pp]),
    "sfm_desc_destroy": (None, [_vp]),
    "sfm_desc_rows": (_i, [_vp]),
    "sfm_desc_is_exact": (_i, [_vp]),
    "sfm_desc_match": (_i, [_vp, _vp, _vp, _d, _vp, _vp, _vp, _vp, _i]),
    "sfm_desc_match_batched": (_i, [_vp, _i, _vp, _vp, _d, _vp, _vp, _vp, _vp]),
    "sfm_match_gather": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sfm_triangulate": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i]),
    "sfm_reproj_error": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "sfm_common_points": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "sfm_gather_rows": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "sfm_compact_pairs": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "sfm_pnp_score": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _f, _vp, _vp]),
    "sfm_pnp_ransac": (_i, [_vp, _vp, _vp, _i, _vp, _i, _f, _d, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sfm_pnp_ransac_hyp": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _f, _d, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sfm_rodrigues_to_matrix": (_i, [_vp, _vp]),
    "sfm_rodrigues_to_vector": (_i, [_vp, _vp]),
    "sfm_epnp": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "sfm_ransac_subsets": (_i, [_i, _i, _vp]),
    "sfm_epnp_batch": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "sfm_ba_create": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _pp]),
    "sfm_ba_destroy": (None, [_vp]),
    "sfm_ba_set_totals": (_i, [_vp, _i64, _i64]),
    "sfm_debug_match_tc_dump": (_i, [_vp, _vp, _vp, _vp, _i64]),
    "sfm_find_essential_mat_batched": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _d, _d, _i, _vp, _vp, _vp]),
    "sfm_five_point": (_i, [_vp, _vp, _vp, _vp]),
    "sfm_find_essential_mat": (_i, [_vp, _vp, _vp, _i, _i, _vp, _d, _d, _i, _vp, _vp, _vp]),
    "sfm_recover_pose": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _d, _vp, _vp, _vp, _vp, _vp]),
    "sfm_desc_create_batched": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp]),
    "sfm_chain_run": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "sfm_chain_create": (_i, [_vp, _vp, _vp, _vp, _i, _pp]),
    "sfm_chain_destroy": (None, [_vp]),
    "sfm_chain_extend": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sfm_chain_extend_async": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "sfm_chain_collect": (_i, [_vp, _vp, _vp]),
    "sfm_desc_match_gather_batched": (_i, [_vp, _i, _vp, _vp, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sfm_debug_match_tc_timeline": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "sfm_ba_set_params": (_i, [_vp, _vp, _vp]),
    "sfm_ba_get_params": (_i, [_vp, _vp, _vp]),
    "sfm_ba_eval": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "sfm_ba_gn_step": (_i, [_vp, _d, _vp]),
    "sfm_ba_reference_fd": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "sfm_reduced_solve": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "sfm_upload_batch": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "sfm_reduced_solve_pcg": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "sfm_ba_build_system": (_i, [_vp, _d]),
    "sfm_ba_read": (_i, [_vp, _i, _vp, _i64]),
    "sfm_nccl_unique_id": (_i, [_vp]),
    "sfm_ba_comm_init": (_i, [_vp, _vp, _i, _i]),
    "sfm_ba_comm_destroy": (_i, [_vp]),
}

_missing = []
for _name, (_res, _args) in PROTOTYPES.items():
    try:
        _fn = getattr(lib, _name)
    except AttributeError:
        _missing.append(_name)
        continue
    _fn.restype = _res
    _fn.argtypes = _args
if _missing:
    raise ImportError(f"{LIB_PATH} does not export: {', '.join(_missing)} (stale build?)")


def last_error() -> str:
    return (lib.sfm_last_error() or b"").decode("utf-8", "replace")


def check(status: int) -> None:
    if status != 0:
        raise error(status, last_error())
