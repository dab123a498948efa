"""The index arithmetic of the CG product (csrc/pcg.cu: lane_map, row_times_p, the dealing of block rows to CTAs and
of column ranges to warps) replayed lane by lane in numpy against a dense symmetric product — a design aid that runs
without a GPU: every load address is asserted to lie inside the packed lower block triangle, the upper triangle of the
stored diagonal blocks is poisoned (it must never be read), and S p must come out to rounding.
  python tools/pcg_lane_map_check.py"""
import numpy as np

BATCH, WARPS = 8, 8          # PCG_BATCH, PCG_WARPS


def lane_map(lane):
    act = lane < 27
    slot = lane // 9 if act else 0
    c = lane - 9 * slot if act else 0
    e0 = 4 * c
    ilo, ihi = e0 // 6, (e0 + 3) // 6
    j0 = e0 - 6 * ilo
    return act, slot, c, ilo, ihi, j0, (j0 + 2) % 6


def row_times_p(S4, ps, a, b0, b1):
    """y_a over the column blocks [b0, b1): what one warp computes (the sum over lanes is the shuffle reduction)."""
    y = np.zeros(6)
    n4 = len(S4)
    for lane in range(32):
        act, slot, c, ilo, ihi, j0, j2 = lane_map(lane)
        lo = hi = t0 = t1 = t2 = t3 = 0.0
        # stored blocks (a, b), b < a
        l0, nL = b0, min(b1, a) - b0
        rowq, pq = (a * (a + 1) // 2 + l0 + slot) * 9 + c, 6 * (l0 + slot)
        s0 = 0
        while s0 + 3 * BATCH <= nL:
            for u in range(BATCH):
                k = rowq + (s0 + 3 * u) * 9
                assert 0 <= k < n4
                v, pp = S4[k], pq + 6 * (s0 + 3 * u)
                lo += v[0] * ps[pp + j0] + v[1] * ps[pp + j0 + 1]
                hi += v[2] * ps[pp + j2] + v[3] * ps[pp + j2 + 1]
            s0 += 3 * BATCH
        if s0 < nL:
            for u in range(BATCH):
                bi = s0 + 3 * u + slot
                k = rowq + (min(bi, nL - 1) - slot) * 9
                assert 0 <= k < n4
                v = S4[k] if bi < nL else np.zeros(4)
                pp = pq + 6 * (min(bi, nL - 1) - slot)
                lo += v[0] * ps[pp + j0] + v[1] * ps[pp + j0 + 1]
                hi += v[2] * ps[pp + j2] + v[3] * ps[pp + j2 + 1]
        # stored blocks (b, a), b > a
        l0 = max(b0, a + 1)
        nT, colq, s0 = b1 - l0, a * 9 + c, 0
        while s0 + 3 * BATCH <= nT:
            for u in range(BATCH):
                b = l0 + s0 + 3 * u + slot
                k = colq + (b * (b + 1) // 2) * 9
                assert 0 <= k < n4
                v = S4[k]
                t0 += v[0] * ps[6 * b + ilo]; t1 += v[1] * ps[6 * b + ilo]
                t2 += v[2] * ps[6 * b + ihi]; t3 += v[3] * ps[6 * b + ihi]
            s0 += 3 * BATCH
        if s0 < nT:
            for u in range(BATCH):
                bi = s0 + 3 * u + slot
                b = l0 + min(bi, nT - 1)
                k = colq + (b * (b + 1) // 2) * 9
                assert 0 <= k < n4
                v = S4[k] if bi < nT else np.zeros(4)
                t0 += v[0] * ps[6 * b + ilo]; t1 += v[1] * ps[6 * b + ilo]
                t2 += v[2] * ps[6 * b + ihi]; t3 += v[3] * ps[6 * b + ihi]
        # the diagonal block through its lower triangle
        if b0 <= a < b1:
            v = S4[(a * (a + 1) // 2 + a) * 9 + c] if slot == 0 else np.zeros(4)
            pa = ps[6 * a:6 * a + 6]
            lo += (v[0] if j0 <= ilo else 0) * pa[j0] + (v[1] if j0 + 1 <= ilo else 0) * pa[j0 + 1]
            hi += (v[2] if j2 <= ihi else 0) * pa[j2] + (v[3] if j2 + 1 <= ihi else 0) * pa[j2 + 1]
            t0 += (v[0] if j0 < ilo else 0) * pa[ilo]; t1 += (v[1] if j0 + 1 < ilo else 0) * pa[ilo]
            t2 += (v[2] if j2 < ihi else 0) * pa[ihi]; t3 += (v[3] if j2 + 1 < ihi else 0) * pa[ihi]
        if not act:
            continue
        for j, val in ((ilo, lo), (ihi, hi), (j0, t0), (j0 + 1, t1), (j2, t2), (j2 + 1, t3)):
            y[j] += val
    return y


def product(S, p, grid):
    n = len(p)
    C = n // 6
    blocks = np.zeros((C * (C + 1) // 2, 6, 6))
    for a in range(C):
        for b in range(a + 1):
            blk = S[6 * a:6 * a + 6, 6 * b:6 * b + 6].copy()
            if a == b:
                blk[np.triu_indices(6, 1)] = 1e30
            blocks[a * (a + 1) // 2 + b] = blk
    S4 = blocks.reshape(-1, 4)
    y = np.full(n, np.nan)
    for cta in range(grid):
        rows = (C - 1 - cta) // grid + 1 if cta < C else 0
        parts = max(1, WARPS // rows) if rows else 1
        assert rows * parts <= WARPS
        ysm = np.zeros((WARPS, 6))
        for warp in range(rows * parts):
            slot, pi = divmod(warp, parts)
            ysm[warp] = row_times_p(S4, p, cta + slot * grid, C * pi // parts, C * (pi + 1) // parts)
        for tid in range(6 * rows):
            slot, i = divmod(tid, 6)
            y[6 * (cta + slot * grid) + i] = sum(ysm[slot * parts + pi][i] for pi in range(parts))
    return y


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for C, grid in [(1, 148), (2, 148), (7, 148), (11, 3), (53, 7), (30, 4), (100, 148), (64, 8)]:
        n = 6 * C
        A = rng.normal(size=(n, n))
        S = A + A.T
        p = rng.normal(size=n)
        err = np.abs(product(S, p, grid) - S @ p).max()
        print(f"C = {C:4d} on {grid:3d} CTAs: max |S p - dense| = {err:.1e}")
        assert err < 1e-11
