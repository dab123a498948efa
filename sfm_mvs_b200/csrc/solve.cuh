// solve.cuh — dense SPD solve used by the bundle-adjustment step (solve.cu).
#pragma once
#include <algorithm>

#include "common.cuh"

// Solves S x = -g.  S (float32, the lower triangle of 6x6 blocks: block (a, b), b <= a, = 36 contiguous floats at (a (a+1) / 2 + b) * 36;
// n is a multiple of 6), g (n float32), A: (n+1) x n float64
// scratch (factor L is left in its lower triangle), x (n float64), info: device int, 0 or the
// 1-based index of the first non-positive pivot.  Stream-ordered, no host synchronisation.
// doubles of scratch `A` must provide: (n+1) x n matrix + the inverses of the 32x32 diagonal blocks
// skip_if (device int, may be null): when it reads 1 at run time — the conjugate-gradient solver below has already
// produced x — every kernel of this solve returns at once.
size_t sfm_spd_scratch_doubles(int n);
int sfm_spd_solve(sfm_ctx* ctx, const float* S, const float* g, int n, double* A, double* x, int* info, const int* skip_if = nullptr);

// The same system by block-Jacobi preconditioned conjugate gradients in one persistent kernel (pcg.cu): the default of
// the LM step when the vectors fit in shared memory (n <= ~3200).  *status_dev = 1: x holds the solution (relative
// residual 1e-8) and *info = 0; 0: not solved (not positive definite, no convergence) — run sfm_spd_solve with
// skip_if = status_dev behind it.  scratch: sfm_pcg_scratch_doubles(n) doubles.
size_t sfm_pcg_scratch_doubles(int n);
bool sfm_spd_pcg_fits(sfm_ctx* ctx, int n);
int sfm_spd_pcg(sfm_ctx* ctx, const float* S, const float* g, int n, double* scratch, double* x, int* status_dev, int* info,
                int* iters_dev, double rel_tol);
