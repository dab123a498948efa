"""Load the reference's OWN function definitions from /root/reference/sfm.py.

TEST INFRASTRUCTURE ONLY; build-container only (/root/reference does not exist on
the GPU box, so nothing in the `-m gpu` tests, smoke() or bench.py calls this).

`import sfm` cannot work: module-level code opens a GUI window (sfm.py:274), lists a
hard-coded directory (sfm.py:30,288) and imports open3d/matplotlib (sfm.py:11,13).
So the file is parsed with `ast`, only top-level `def`s are kept and executed in a
namespace that supplies the modules they use, plus the one shim the authors'
opencv-contrib build needs (cv2.xfeatures2d.SIFT_create -> cv2.SIFT_create).
No reference source text is copied into this repository.
"""
from __future__ import annotations

import ast
import copy
import os
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sfm.py"))


def load_reference_defs(filename: str = "sfm.py") -> dict:
    import cv2
    import numpy as np
    from scipy.optimize import least_squares

    path = os.path.join(REFERENCE_ROOT, filename)
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    tree.body = [node for node in tree.body if isinstance(node, ast.FunctionDef)]
    if not hasattr(cv2, "xfeatures2d"):
        cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=cv2.SIFT_create)
    ns = dict(cv2=cv2, np=np, copy=copy, os=os, least_squares=least_squares)
    exec(compile(tree, path, "exec"), ns)
    return {k: v for k, v in ns.items() if isinstance(v, types.FunctionType)}
