// solve.cu — dense symmetric-positive-definite solve of the reduced camera system (n = 6C; 3000 at
// C = 500), float64, hand-written right-looking blocked Cholesky with NB = 32:
//
//   A is (n+1) x n row-major, lower triangle used; row n carries the right-hand side, so the
//   forward substitution L y = b falls out of the panel solves for free (row n ends up as y^T).
//   per block column k:
//     chol_panel_kernel  every CTA factors the 32x32 diagonal block redundantly in warp 0 with
//                        register rows + shuffles (no block-level sync per column), then each
//                        thread solves one row of the panel below against it (x L_kk^T = a)
//     chol_syrk_kernel   trailing update A22 -= L21 L21^T, 64x64 tile per CTA, 4x4 per thread,
//                        K = 32 staged in shared memory, lower-triangular tiles only
//   backward substitution L^T x = y, one kernel per block from the bottom up: warp 0 of every CTA
//   solves the transposed 32x32 triangle with shuffles, then each thread updates one earlier entry.
#include <math.h>

#include "common.cuh"
#include "solve.cuh"

namespace {

constexpr int NB = 32;

// Cholesky of a 32x32 SPD block held one row per lane (r[j], j <= lane meaningful).  Rows >= nb are
// treated as identity.  On return lane i holds row i of L.
__device__ __forceinline__ void warp_chol32(double (&r)[NB], int lane, int nb, int* info, int k0) {
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    double d = __shfl_sync(0xffffffffu, r[j], j);
    if (j < nb && !(d > 0.0)) {
      if (lane == 0 && info && *info == 0) *info = k0 + j + 1;
      d = 1.0;
    }
    const double ljj = sqrt(d);
    const double lij = (lane == j) ? ljj : r[j] / ljj;      // column j of L, valid for lane >= j
    r[j] = lij;
#pragma unroll
    for (int c = j + 1; c < NB; ++c) {
      const double lcj = __shfl_sync(0xffffffffu, lij, c);
      if (lane >= c) r[c] -= lij * lcj;
    }
  }
}

__global__ void __launch_bounds__(64) chol_panel_kernel(double* __restrict__ A, int n, int k0, int* __restrict__ info) {
  __shared__ double L[NB][NB + 1];
  __shared__ double inv_d[NB];
  const int nb = min(NB, n - k0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    double r[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j)
      r[j] = (lane < nb && j <= lane) ? A[(size_t)(k0 + lane) * n + k0 + j] : ((j == lane) ? 1.0 : 0.0);
    warp_chol32(r, lane, nb, blockIdx.x == 0 ? info : nullptr, k0);
#pragma unroll
    for (int j = 0; j < NB; ++j) L[lane][j] = (j <= lane) ? r[j] : 0.0;
#pragma unroll
    for (int j = 0; j < NB; ++j)
      if (j == lane) inv_d[lane] = 1.0 / r[j];
    if (blockIdx.x == 0 && lane < nb) {
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (j <= lane) A[(size_t)(k0 + lane) * n + k0 + j] = r[j];
    }
  }
  __syncthreads();
  const int row = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;      // rows below the block, incl. the rhs row n
  if (row > n) return;
  double* a = A + (size_t)row * n + k0;
  double x[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) x[j] = (j < nb) ? a[j] : 0.0;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    double s = x[j];
#pragma unroll
    for (int k = 0; k < j; ++k) s -= x[k] * L[j][k];
    x[j] = s * inv_d[j];
  }
#pragma unroll
  for (int j = 0; j < NB; ++j)
    if (j < nb) a[j] = x[j];
}

// rows [base, n] (n+1-base of them, the last is the rhs row), columns [base, n)
__global__ void __launch_bounds__(256) chol_syrk_kernel(double* __restrict__ A, int n, int k0, int nb, int n_row_tiles) {
  __shared__ double Pi[64][NB + 1];
  __shared__ double Pj[64][NB + 1];
  int t = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  const int base = k0 + nb;
  const int i0 = base + ti * 64, j0 = base + tj * 64;
  if (j0 >= n) return;
  for (int e = threadIdx.x; e < 64 * NB; e += blockDim.x) {
    int r = e / NB, c = e % NB;
    Pi[r][c] = (i0 + r <= n && c < nb) ? A[(size_t)(i0 + r) * n + k0 + c] : 0.0;
    Pj[r][c] = (j0 + r < n && c < nb) ? A[(size_t)(j0 + r) * n + k0 + c] : 0.0;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
  for (int k = 0; k < NB; ++k) {
    double av[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) av[a] = Pi[ty + 16 * a][k];
#pragma unroll
    for (int b = 0; b < 4; ++b) bv[b] = Pj[tx + 16 * b][k];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
      if (i <= n && j < n && j <= i) A[(size_t)i * n + j] -= acc[a][b];
    }
}

// y (row n of A after the factorisation) is consumed in place; the solution goes to a separate
// array so that CTAs starting late still read the unsolved y_k.
// Block k: x_k = L_kk^-T y_k ; y[0:k0] -= L[k-block rows, 0:k0]^T x_k.
__global__ void __launch_bounds__(256) chol_back_kernel(const double* __restrict__ A, int n, int k0, double* __restrict__ y,
                                                        double* __restrict__ x) {
  __shared__ double xs[NB];
  __shared__ double Ls[NB][NB + 1];
  const int nb = min(NB, n - k0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    int r = e / NB, c = e % NB;
    Ls[r][c] = (r < nb && c <= r) ? A[(size_t)(k0 + r) * n + k0 + c] : ((r == c) ? 1.0 : 0.0);
  }
  __syncthreads();
  if (warp == 0) {
    double v = (lane < nb) ? y[k0 + lane] : 0.0;
    const double inv = 1.0 / Ls[lane][lane];
#pragma unroll
    for (int j = NB - 1; j >= 0; --j) {
      double xj = __shfl_sync(0xffffffffu, v * inv, j);
      if (lane == j) v = xj;
      if (lane < j) v -= Ls[j][lane] * xj;
    }
    xs[lane] = v;
    if (blockIdx.x == 0 && lane < nb) x[k0 + lane] = v;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k0) return;
  double s = 0.0;
  for (int r = 0; r < nb; ++r) s = fma(A[(size_t)(k0 + r) * n + i], xs[r], s);
  y[i] -= s;
}

__global__ void widen_kernel(const float* __restrict__ S, const float* __restrict__ g, int n, double* __restrict__ A) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)(n + 1) * n;
  if (idx >= total) return;
  int r = (int)(idx / n), c = (int)(idx % n);
  A[idx] = (r == n) ? -(double)g[c] : ((c <= r) ? (double)S[idx] : 0.0);
}

}  // namespace

int sfm_spd_solve(sfm_ctx* ctx, const float* S, const float* g, int n, double* A, double* x, int* info) {
  size_t total = (size_t)(n + 1) * n;
  SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (widen_kernel<<<(unsigned)div_up64((int64_t)total, 256), 256, 0, ctx->stream>>>(S, g, n, A)));
  SFM_CUDA(cudaMemsetAsync(info, 0, sizeof(int), ctx->stream));
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int nb = std::min(NB, n - k0);
    const int rows_below = n + 1 - (k0 + nb);                 // >= 1: the rhs row
    SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (chol_panel_kernel<<<div_up(rows_below, 64), 64, 0, ctx->stream>>>(A, n, k0, info)));
    const int cols = n - (k0 + nb);
    if (cols > 0) {
      const int tiles = div_up(rows_below, 64);
      SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (chol_syrk_kernel<<<tiles * (tiles + 1) / 2, 256, 0, ctx->stream>>>(A, n, k0, nb, tiles)));
    }
  }
  double* y = A + (size_t)n * n;
  for (int k0 = ((n - 1) / NB) * NB; k0 >= 0; k0 -= NB)
    SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (chol_back_kernel<<<std::max(1, div_up(k0, 256)), 256, 0, ctx->stream>>>(A, n, k0, y, x)));
  return SFM_OK;
}
