"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/sfm_b200.h declares, fails loudly without a GPU, and its host utilities agree with the
oracle / cv2.  No kernel is launched here."""
import os
import re
import subprocess

import cv2
import numpy as np
import pytest

import sfm_mvs_b200 as sfm
from oracle import restated
from sfm_mvs_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "sfm_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return set(re.findall(r"\b(sfm_[a-z0-9_]+)\s*\(", h))


def test_library_exports_every_declared_symbol():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (sfm_\w+)", out))
    declared = _declared()
    assert declared, "header parse failed"
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    assert exported <= declared, f"exported but not declared: {sorted(exported - declared)}"


def test_ctypes_prototypes_cover_the_header():
    assert set(_lib.PROTOTYPES) == _declared()
    assert _lib.lib.sfm_version() == 100


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sfm.error) as ei:
        sfm.Context(0)
    assert "no CPU path" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sfm_mvs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "import cv2" not in src or f == "cv2_compat.py", f"{f} imports cv2"


def test_ransac_subset_stream_equals_restated():
    for n in (5, 6, 37, 1000, 4999):
        assert np.array_equal(sfm.ransac_subsets(n, 100), restated.ransac_subsets(n, 100))


def test_rodrigues_host_utilities_equal_cv2():
    rng = np.random.default_rng(0)
    for _ in range(100):
        r = rng.normal(0, 1.2, 3)
        Rc, _ = cv2.Rodrigues(r)
        assert np.abs(sfm.rodrigues_to_matrix(r) - Rc).max() < 1e-15
        rc, _ = cv2.Rodrigues(Rc)
        assert np.abs(sfm.rodrigues_to_vector(Rc) - rc.ravel()).max() < 1e-12
    assert np.array_equal(sfm.rodrigues_to_matrix(np.zeros(3)), np.eye(3))


def test_epnp_host_solver_recovers_noise_free_pose():
    rng = np.random.default_rng(1)
    K = synth.K_GUSTAV
    worst = 0.0
    for trial in range(50):
        R, t = synth.orbit_pose(rng.uniform(-0.6, 0.6))
        n = 5 if trial % 2 == 0 else 12
        X = np.c_[rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n), rng.uniform(5, 11, n)].astype(np.float32)
        uv, _ = synth.project(K, R, t, X.astype(np.float64))
        Re, te = sfm.epnp(X, uv.astype(np.float32), K)
        uv2, _ = synth.project(K, Re, te.reshape(3, 1), X.astype(np.float64))
        worst = max(worst, np.abs(uv2 - uv).max())
        assert abs(np.linalg.det(Re) - 1) < 1e-9
    assert worst < 0.05, worst       # float32 pixel quantisation only


def test_library_sass_is_blackwell_native():
    """What proves a Blackwell-native build (profiling guide): the match kernel issues tcgen05.mma on CTA pairs
    (SASS `UTCHMMA.2CTA`), reads accumulators out of TMEM (`LDTM`), stages operands with the bulk-copy engine
    (`UBLKCP`), and nothing in the library falls back to the legacy tensor path (`HMMA`, i.e. mma.sync / wmma).
    The library holds sm_100a code only."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lst = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.", lst))
    assert archs == {"sm_100a"}, archs
    full = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    k1 = full[full.index("match_tc_kernel"):]
    assert "UTCHMMA.2CTA" in k1 and "LDTM" in k1 and "UBLKCP" in k1 and "UTCBAR" in k1
    assert not re.search(r"\bHMMA\b", full) and not re.search(r"\b[HQI]GMMA\b", full)
