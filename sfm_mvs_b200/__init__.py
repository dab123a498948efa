"""sfm_mvs_b200 — B200-native geometry engine for the hot paths of FlagArihant2000/sfm-mvs.

Host side is Python (as the reference's is) over a ctypes C ABI (include/sfm_b200.h) into
hand-written sm_100a CUDA kernels (sfm_mvs_b200/csrc).  Importing this package requires the built
shared library; using it requires a CUDA device.  There is no CPU fallback.
"""
from ._lib import LIB_PATH, error  # noqa: F401  (import fails loudly if the library is missing)
from .engine import (BAProblem, Context, Descriptors, epnp, five_point, nccl_unique_id, ransac_subsets,  # noqa: F401
                     rodrigues_to_matrix, rodrigues_to_vector)
from .cv2_compat import (NORM_L2, RATIO, SOLVEPNP_ITERATIVE, BFMatcher, BundleAdjustment, DMatch, PnP,  # noqa: F401
                         ReprojectionError, Triangulation, common_points, default_context, findEssentialMat, knn2,
                         match_keypoints, patch_cv2, recoverPose, set_default_context, solvePnPRansac, triangulatePoints,
                         unpatch_cv2)
from . import ba  # noqa: F401

__version__ = "0.1.0"
from .io import to_ply, save_poses, load_poses, load_ply  # noqa: E402,F401  (sfm.py:169-201, :423)
