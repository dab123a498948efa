"""Wire / disk formats of the reference — host-side I/O, not part of the data-parallel path.

  to_ply(path, point_cloud, colors, densify)   sfm.py:169-201   ASCII PLY, points x200, outlier cull
                                               dist < mean(dist) + 300, tab-indented header, '%f %f %f %d %d %d'
  save_poses(fname, K, Ps)                      sfm.py:276,334-335,375,423   pose.csv: the 9 entries of K, then every
                                               3x4 projection matrix P = K[R|t] row-major, ONE value per line
  load_poses / load_ply                         readers for the artifacts the reference ships (pose.csv,
                                               Point_Cloud/sparse.ply) — used to build fixture-derived scenes

The files are byte-identical to what the reference writes for the same arrays (tests/test_io.py).
"""
from __future__ import annotations

import os

import numpy as np

_PLY_HEADER = "ply\n\t\tformat ascii 1.0\n\t\telement vertex %(vert_num)d\n\t\tproperty float x\n\t\tproperty float y\n" \
              "\t\tproperty float z\n\t\tproperty uchar blue\n\t\tproperty uchar green\n\t\tproperty uchar red\n" \
              "\t\tend_header\n\t\t"


def cull_and_scale(point_cloud, colors):
    """The vertex table to_ply writes: points x200 | colours, rows with dist-to-mean >= mean(dist)+300 dropped."""
    out_points = np.asarray(point_cloud).reshape(-1, 3) * 200
    out_colors = np.asarray(colors).reshape(-1, 3)
    verts = np.hstack([out_points, out_colors])
    mean = np.mean(verts[:, :3], axis=0)
    temp = verts[:, :3] - mean
    dist = np.sqrt(temp[:, 0] ** 2 + temp[:, 1] ** 2 + temp[:, 2] ** 2)
    return verts[np.where(dist < np.mean(dist) + 300)]


def to_ply(path, point_cloud, colors, densify=False):
    """Drop-in for the reference's to_ply (same arguments, same file: <path>/Point_Cloud/sparse.ply or dense.ply)."""
    verts = cull_and_scale(point_cloud, colors)
    name = "dense.ply" if densify else "sparse.ply"
    with open(os.path.join(path, "Point_Cloud", name), "w") as f:
        f.write(_PLY_HEADER % dict(vert_num=len(verts)))
        np.savetxt(f, verts, "%f %f %f %d %d %d")
    return len(verts)


def save_poses(fname, K, Ps):
    """pose.csv as sfm.py:423 writes it: K.ravel() followed by each P.ravel(), one '%.18e' value per line."""
    arr = np.asarray(K, np.float64).ravel()
    for P in Ps:
        arr = np.hstack((arr, np.asarray(P, np.float64).ravel()))
    np.savetxt(fname, arr, delimiter="\n")
    return arr


def load_poses(fname):
    """-> K (3,3), Ps (n,3,4) from a pose.csv."""
    v = np.loadtxt(fname)
    if v.size < 9 or (v.size - 9) % 12:
        raise ValueError(f"{fname}: {v.size} values is not 9 + 12*n")
    return v[:9].reshape(3, 3), v[9:].reshape(-1, 3, 4)


def load_ply(fname):
    """-> points (n,3) float64 in the reference's world units (file / 200), colours (n,3) uint8 (b,g,r)."""
    with open(fname, "r") as f:
        n = None
        for line in f:
            s = line.strip()
            if s.startswith("element vertex"):
                n = int(s.split()[-1])
            if s == "end_header":
                break
        data = np.loadtxt(f, ndmin=2)
    if n is not None and len(data) != n:
        raise ValueError(f"{fname}: header says {n} vertices, file has {len(data)}")
    return data[:, :3] / 200.0, data[:, 3:6].astype(np.uint8)


def lookup_colors(img, pts):
    """sfm.py:393-395 — the colour of every new point: pixel coordinates truncated to int32 (np.array(temp2,
    dtype=np.int32)), then img[y, x] per point.  pts: (2,N) as the loop holds temp2, or (N,2); img: (H,W,3) uint8
    (BGR as cv2.imread gives it).  -> (N,3), the rows the reference stacks into `colorstot` for to_ply."""
    p = np.asarray(pts)
    if p.ndim != 2 or 2 not in p.shape:
        raise ValueError(f"lookup_colors: points must be (2,N) or (N,2), got {p.shape}")
    if p.shape[0] != 2:
        p = p.T
    reg = np.array(p, dtype=np.int32)                 # the reference's cast: truncation toward zero
    return np.asarray(img)[reg[1], reg[0]]
