// solve.cuh — dense SPD solve used by the bundle-adjustment step (solve.cu).
#pragma once
#include <algorithm>

#include "common.cuh"

// Solves S x = -g.  S (n x n float32 stored by 6x6 blocks — block (a, b) = 36 contiguous floats —, lower triangle read;
// n is a multiple of 6), g (n float32), A: (n+1) x n float64
// scratch (factor L is left in its lower triangle), x (n float64), info: device int, 0 or the
// 1-based index of the first non-positive pivot.  Stream-ordered, no host synchronisation.
// doubles of scratch `A` must provide: (n+1) x n matrix + the inverses of the 32x32 diagonal blocks
size_t sfm_spd_scratch_doubles(int n);
int sfm_spd_solve(sfm_ctx* ctx, const float* S, const float* g, int n, double* A, double* x, int* info);
