"""Wire formats: to_ply / pose.csv byte-identical to the reference's writers."""
import os

import numpy as np
import pytest

from oracle import refload
from sfm_mvs_b200 import io as sio


def _cloud(seed=0, n=400):
    rng = np.random.default_rng(seed)
    X = rng.normal(0, 1.5, (n, 3))
    X[::37] *= 40.0                                    # a few far outliers for the cull
    return X.astype(np.float32).reshape(n, 1, 3), rng.integers(0, 256, (n, 3)).astype(np.uint8)


def test_ply_and_pose_round_trip(tmp_path):
    X, col = _cloud()
    os.makedirs(tmp_path / "Point_Cloud")
    kept = sio.to_ply(str(tmp_path), X, col, False)
    pts, c = sio.load_ply(str(tmp_path / "Point_Cloud" / "sparse.ply"))
    verts = sio.cull_and_scale(X, col)
    assert kept == len(verts) == len(pts) < len(X)
    assert np.allclose(pts * 200.0, verts[:, :3], atol=1e-6) and np.array_equal(c, verts[:, 3:].astype(np.uint8))
    K = np.array([[1196.98, 0, 466.19], [0, 1199.06, 314.13], [0, 0, 1.0]])
    Ps = [K @ np.hstack([np.eye(3), np.array([[i], [0.5], [2.0]])]) for i in range(4)]
    sio.save_poses(str(tmp_path / "pose.csv"), K, Ps)
    K2, P2 = sio.load_poses(str(tmp_path / "pose.csv"))
    assert np.array_equal(K2, K) and np.array_equal(P2, np.array(Ps))
    assert len(open(tmp_path / "pose.csv").read().splitlines()) == 9 + 12 * 4     # one value per line (sfm.py:423)


@pytest.mark.skipif(not refload.available(), reason="build container only: needs /root/reference")
def test_to_ply_equals_the_reference_writer(tmp_path):
    import contextlib, io
    ref = refload.load_reference_defs()
    for seed, densify in ((1, False), (2, True)):
        X, col = _cloud(seed)
        a, b = tmp_path / f"a{seed}", tmp_path / f"b{seed}"
        os.makedirs(a / "Point_Cloud"); os.makedirs(b / "Point_Cloud")
        with contextlib.redirect_stdout(io.StringIO()):
            ref["to_ply"](str(a), X, col, densify)
        sio.to_ply(str(b), X, col, densify)
        name = "dense.ply" if densify else "sparse.ply"
        assert open(a / "Point_Cloud" / name, "rb").read() == open(b / "Point_Cloud" / name, "rb").read()


@pytest.mark.skipif(not os.path.isfile("/root/reference/pose.csv"), reason="build container only")
def test_reads_the_reference_artifacts():
    K, Ps = sio.load_poses("/root/reference/pose.csv")
    assert Ps.shape == (57, 3, 4) and abs(K[0, 0] - 1196.98) < 1.0
    Rt = np.linalg.inv(K) @ Ps[5]
    assert abs(np.linalg.det(Rt[:, :3]) - 1.0) < 1e-9
    pts, col = sio.load_ply("/root/reference/Point_Cloud/sparse.ply")
    assert len(pts) == 19282 and col.dtype == np.uint8


def test_lookup_colors_equals_the_reference_expression():
    """sfm.py:392-395: pts1_reg = np.array(temp2, dtype=np.int32); colors = np.array([img2[l[1], l[0]] for l in pts1_reg.T])"""
    from sfm_mvs_b200 import io as sio
    rng = np.random.default_rng(3)
    img2 = rng.integers(0, 256, (648, 968, 3), dtype=np.uint8)
    temp2 = np.stack([rng.uniform(0, 967.9, 500), rng.uniform(0, 647.9, 500)]).astype(np.float32)     # (2,N) as in the loop
    pts1_reg = np.array(temp2, dtype=np.int32)
    want = np.array([img2[l[1], l[0]] for l in pts1_reg.T])
    assert np.array_equal(sio.lookup_colors(img2, temp2), want)
    assert np.array_equal(sio.lookup_colors(img2, temp2.T.copy()), want)
