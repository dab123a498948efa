"""sfm_mvs_b200 — B200-native geometry engine for the hot paths of FlagArihant2000/sfm-mvs.

Host side is Python (as the reference's is) over a ctypes C ABI (include/sfm_b200.h) into
hand-written sm_100a CUDA kernels (sfm_mvs_b200/csrc).  Every engine name needs the built shared library
(the first access loads it and fails loudly if it is missing) and a CUDA device to run.  There is no CPU
fallback.  The host-only helpers — `synth` (seeded synthetic inputs), `io` (PLY / pose.csv wire formats) and
`sharding` (work splits) — import without the library, so the CPU reference arm of bench.py and the host
tests never load the engine.
"""
import importlib

__version__ = "0.2.0"

_ENGINE = {
    "_lib": ("LIB_PATH", "error"),
    "engine": ("BAProblem", "Context", "Descriptors", "epnp", "five_point", "nccl_unique_id", "ransac_subsets",
               "rodrigues_to_matrix", "rodrigues_to_vector"),
    "cv2_compat": ("NORM_L2", "RATIO", "SOLVEPNP_ITERATIVE", "BFMatcher", "BundleAdjustment", "BundleAdjustmentSE3", "DMatch", "PnP",
                   "ReprojectionError", "Triangulation", "common_points", "default_context", "findEssentialMat", "knn2",
                   "match_keypoints", "patch_cv2", "recoverPose", "set_default_context", "solvePnPRansac",
                   "triangulatePoints", "unpatch_cv2"),
    "io": ("to_ply", "save_poses", "load_poses", "load_ply", "lookup_colors"),            # sfm.py:169-201, :423
}
_WHERE = {name: mod for mod, names in _ENGINE.items() for name in names}
_SUBMODULES = ("_lib", "engine", "cv2_compat", "pipeline", "ba", "io", "layout", "sharding", "synth")
__all__ = sorted(_WHERE) + ["ba", "pipeline", "io", "layout", "sharding", "synth"]


def __getattr__(name):
    if name in _WHERE:
        value = getattr(importlib.import_module("." + _WHERE[name], __name__), name)
    elif name in _SUBMODULES:
        value = importlib.import_module("." + name, __name__)
    else:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    globals()[name] = value
    return value
