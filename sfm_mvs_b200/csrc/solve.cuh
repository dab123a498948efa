// solve.cuh — dense SPD solve used by the bundle-adjustment step (solve.cu).
#pragma once
#include <algorithm>

#include "common.cuh"

// Solves S x = -g.  S (float32, the lower triangle of 6x6 blocks: block (a, b), b <= a, = 36 contiguous floats at (a (a+1) / 2 + b) * 36;
// n is a multiple of 6), g (n float32), A: (n+1) x n float64
// scratch (factor L is left in its lower triangle), x (n float64), info: device int, 0 or the
// 1-based index of the first non-positive pivot.  Stream-ordered, no host synchronisation.
// doubles of scratch `A` must provide: (n+1) x n matrix + the inverses of the 32x32 diagonal blocks
size_t sfm_spd_scratch_doubles(int n);
int sfm_spd_solve(sfm_ctx* ctx, const float* S, const float* g, int n, double* A, double* x, int* info);
