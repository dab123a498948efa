"""Host-side views of the engine's packed layouts (no GPU, no extension needed).

The reduced camera system S (6C x 6C, symmetric) lives in HBM as the lower triangle of its 6x6 camera blocks: block
(a, b), b <= a, is 36 contiguous row-major float32 at ((a (a + 1)) / 2 + b) * 36 (csrc/ba.cu; what the Schur kernel
accumulates, the all-reduce sums and both linear solvers read)."""
from __future__ import annotations

import numpy as np


def pack_lower_blocks(S) -> np.ndarray:
    """(6C, 6C) -> (C (C+1) / 2, 6, 6) float32: the blocks of the lower block triangle in row-major (a, b <= a) order.
    Only the lower triangle of S is read, diagonal blocks are taken whole."""
    S = np.asarray(S)
    n = S.shape[0]
    C = n // 6
    if S.ndim != 2 or S.shape != (n, n) or n != 6 * C or C < 1:
        raise ValueError("pack_lower_blocks: S must be (6C, 6C)")
    ia, ib = np.tril_indices(C)
    return np.ascontiguousarray(S.reshape(C, 6, C, 6).transpose(0, 2, 1, 3)[ia, ib], np.float32)


def unpack_lower_blocks(blocks, symmetric: bool = True) -> np.ndarray:
    """The inverse: packed blocks -> dense (6C, 6C) float64, the upper block triangle filled by symmetry (or left zero)."""
    blocks = np.asarray(blocks, np.float64).reshape(-1, 6, 6)
    nb = len(blocks)
    C = int((np.sqrt(8 * nb + 1) - 1) // 2)
    if C * (C + 1) // 2 != nb:
        raise ValueError("unpack_lower_blocks: not a triangular number of blocks")
    S = np.zeros((C, 6, C, 6))
    ia, ib = np.tril_indices(C)
    S[ia, :, ib, :] = blocks
    if symmetric:
        off = ia != ib
        S[ib[off], :, ia[off], :] = blocks[off].transpose(0, 2, 1)
    return S.reshape(6 * C, 6 * C)
