// solve.cu — dense symmetric-positive-definite solve of the reduced camera system (n = 6C; 3000 at
// C = 500), float64, hand-written right-looking blocked Cholesky with NB = 32 and look-ahead.
//
//   A is (n+1) x n row-major, lower triangle used; row n carries the right-hand side, so the
//   forward substitution L y = b falls out of the panel products for free (row n ends up as y^T).
//   Every 32x32 diagonal block is factored by ONE warp with its rows in registers (shuffles, no
//   block-level sync per column) and its inverse L_kk^-1 is kept in a side buffer, which turns the
//   two triangular solves into plain products without serial dependence chains:
//     chol_diag_kernel   block 0 only: factor + invert
//     chol_panel_kernel  rows below block k (incl. the rhs row): x = a * L_kk^-T, one row per thread
//     chol_syrk_kernel   trailing update A22 -= L21 L21^T, 64x64 tile per CTA, 4x4 per thread, K = 32
//                        staged in shared memory, lower-triangular tiles only; the CTA that owns the
//                        first tile then factors + inverts diagonal block k+1 (look-ahead), so the
//                        serial part of step k+1 overlaps the rest of step k's update
//   backward substitution L^T x = y in 256-row super-blocks from the bottom up:
//     back_diag_kernel   one CTA: eight 32-row sub-steps, x_s = L_ss^-T y_s then update of the
//                        super-block's earlier rows
//     back_update_kernel all earlier entries: y[i] -= sum_r L[k0+r][i] x[k0+r]
#include <math.h>

#include "common.cuh"
#include "solve.cuh"

namespace {

constexpr int NB = 32;
constexpr int SB = 256;   // backward-substitution super-block

// Factor + invert the 32x32 diagonal block starting at k0 with ONE warp, entirely in shared memory
// (row stride 33 doubles: lane-per-row accesses are conflict-free, same-address reads broadcast), so it
// costs few registers and can ride inside the update kernel.  Rows >= nb are treated as identity.
// Reads the lower triangle of A, writes L back and L^-1 to Linv + (k0/NB)*NB*NB.
__device__ __forceinline__ void factor_diag_block(double* __restrict__ A, int n, int k0, double* __restrict__ Linv,
                                                  int* __restrict__ info, int lane, double (*Ls)[NB + 1],
                                                  double (*Li)[NB + 1], double* __restrict__ Ld) {
  const int nb = min(NB, n - k0);
  for (int j = 0; j < NB; ++j)
    Ls[lane][j] = (lane < nb && j <= lane) ? A[(size_t)(k0 + lane) * n + k0 + j] : ((j == lane) ? 1.0 : 0.0);
  __syncwarp();
  for (int j = 0; j < NB; ++j) {
    double d = Ls[j][j];
    if (j < nb && !(d > 0.0)) {
      if (lane == 0 && info && *info == 0) *info = k0 + j + 1;   // not positive definite
      d = 1.0;
    }
    const double rinv = rsqrt(d);
    const double lij = (lane == j) ? d * rinv : Ls[lane][j] * rinv;   // column j of L (lanes >= j)
    __syncwarp();
    Ls[lane][j] = (lane >= j) ? lij : 0.0;
    if (lane == j) Ld[j] = rinv;
    __syncwarp();
    for (int c = j + 1; c < NB; ++c)
      if (lane >= c) Ls[lane][c] = fma(-lij, Ls[c][j], Ls[lane][c]);
    __syncwarp();
  }
  // lane c solves L z = e_c  ->  column c of L^-1 (kept in its own column of Li)
  for (int i = 0; i < NB; ++i) {
    double s = (i == lane) ? 1.0 : 0.0;
    for (int k = lane; k < i; ++k) s = fma(-Ls[i][k], Li[k][lane], s);
    Li[i][lane] = (i >= lane) ? s * Ld[i] : 0.0;
  }
  __syncwarp();
  double* out = Linv + (size_t)(k0 / NB) * NB * NB;
  for (int i = 0; i < NB; ++i) {
    out[i * NB + lane] = Li[i][lane];
    if (i < nb && lane <= i) A[(size_t)(k0 + i) * n + k0 + lane] = Ls[i][lane];
  }
}

__global__ void __launch_bounds__(32) chol_diag_kernel(double* __restrict__ A, int n, int k0, double* __restrict__ Linv,
                                                       int* __restrict__ info) {
  __shared__ double Ls[NB][NB + 1];
  __shared__ double Li[NB][NB + 1];
  __shared__ double Ld[NB];
  factor_diag_block(A, n, k0, Linv, info, threadIdx.x, Ls, Li, Ld);
}

// rows k0+nb .. n (the last one is the rhs row): a <- a * L_kk^-T
__global__ void __launch_bounds__(128) chol_panel_kernel(double* __restrict__ A, int n, int k0,
                                                         const double* __restrict__ Linv) {
  __shared__ double Li[NB][NB + 1];
  const int nb = min(NB, n - k0);
  const double* src = Linv + (size_t)(k0 / NB) * NB * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) Li[e / NB][e % NB] = src[e];
  __syncthreads();
  const int row = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (row > n) return;
  double* a = A + (size_t)row * n + k0;
  double v[NB], x[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) v[j] = (j < nb) ? a[j] : 0.0;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c <= j; ++c) s = fma(v[c], Li[j][c], s);
    x[j] = s;
  }
#pragma unroll
  for (int j = 0; j < NB; ++j)
    if (j < nb) a[j] = x[j];
}

// rows [base, n] (n+1-base of them, the last is the rhs row), columns [base, n)
__global__ void __launch_bounds__(256) chol_syrk_kernel(double* __restrict__ A, int n, int k0, int nb,
                                                        double* __restrict__ Linv, int* __restrict__ info) {
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
  // look-ahead: the first tile now holds the final values of diagonal block k+1
  if (t == 0) {
    __syncthreads();
    if (threadIdx.x < 32)
      factor_diag_block(A, n, base, Linv, info, threadIdx.x, reinterpret_cast<double(*)[NB + 1]>(&Pi[0][0]),
                        reinterpret_cast<double(*)[NB + 1]>(&Pj[0][0]), &Pi[32][0]);
  }
}

// One CTA per call: rows [k0, k0+sb) of L^T x = y, eight 32-row sub-steps from the bottom.
__global__ void __launch_bounds__(256) back_diag_kernel(const double* __restrict__ A, int n, int k0, int sb,
                                                        const double* __restrict__ Linv, double* __restrict__ y,
                                                        double* __restrict__ x) {
  __shared__ double ys[SB];
  __shared__ double xs[NB];
  for (int i = threadIdx.x; i < SB; i += blockDim.x) ys[i] = (i < sb) ? y[k0 + i] : 0.0;
  __syncthreads();
  const int nsub = (sb + NB - 1) / NB;
  for (int s = nsub - 1; s >= 0; --s) {
    const int r0 = k0 + s * NB;                       // global row of the sub-block
    const int nb = min(NB, n - r0);
    if (threadIdx.x < 32) {
      // x_s = L_ss^-T y_s : lane j sums Linv[i][j] * y[i] over i >= j
      const double* Li = Linv + (size_t)(r0 / NB) * NB * NB;
      const int lane = threadIdx.x;
      double acc = 0.0;
      for (int i = 0; i < nb; ++i) acc = fma(Li[i * NB + lane], ys[s * NB + i], acc);
      xs[lane] = acc;
      if (lane < nb) x[r0 + lane] = acc;
    }
    __syncthreads();
    // earlier rows of this super-block: y[i] -= sum_r L[r0+r][k0+i] x_s[r]
    for (int i = threadIdx.x; i < s * NB; i += blockDim.x) {
      double acc = 0.0;
      for (int r = 0; r < nb; ++r) acc = fma(A[(size_t)(r0 + r) * n + k0 + i], xs[r], acc);
      ys[i] -= acc;
    }
    __syncthreads();
  }
}

// y[i] -= sum_{r < sb} L[k0+r][i] x[k0+r] for i < k0
__global__ void __launch_bounds__(128) back_update_kernel(const double* __restrict__ A, int n, int k0, int sb,
                                                          const double* __restrict__ x, double* __restrict__ y) {
  __shared__ double xs[SB];
  for (int i = threadIdx.x; i < sb; i += blockDim.x) xs[i] = x[k0 + i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k0) return;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int r = 0;
  for (; r + 3 < sb; r += 4) {
    a0 = fma(A[(size_t)(k0 + r) * n + i], xs[r], a0);
    a1 = fma(A[(size_t)(k0 + r + 1) * n + i], xs[r + 1], a1);
    a2 = fma(A[(size_t)(k0 + r + 2) * n + i], xs[r + 2], a2);
    a3 = fma(A[(size_t)(k0 + r + 3) * n + i], xs[r + 3], a3);
  }
  for (; r < sb; ++r) a0 = fma(A[(size_t)(k0 + r) * n + i], xs[r], a0);
  y[i] -= (a0 + a1) + (a2 + a3);
}

__global__ void widen_kernel(const float* __restrict__ S, const float* __restrict__ g, int n, double* __restrict__ A) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)(n + 1) * n;
  if (idx >= total) return;
  int r = (int)(idx / n), c = (int)(idx % n);
  A[idx] = (r == n) ? -(double)g[c] : ((c <= r) ? (double)S[idx] : 0.0);
}

}  // namespace

size_t sfm_spd_scratch_doubles(int n) {
  return ((size_t)n + 1) * n + (size_t)div_up(n, NB) * NB * NB;
}

int sfm_spd_solve(sfm_ctx* ctx, const float* S, const float* g, int n, double* A, double* x, int* info) {
  size_t total = (size_t)(n + 1) * n;
  double* Linv = A + total;
  SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (widen_kernel<<<(unsigned)div_up64((int64_t)total, 256), 256, 0, ctx->stream>>>(S, g, n, A)));
  SFM_CUDA(cudaMemsetAsync(info, 0, sizeof(int), ctx->stream));
  SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (chol_diag_kernel<<<1, 32, 0, ctx->stream>>>(A, n, 0, Linv, info)));
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int nb = std::min(NB, n - k0);
    const int rows_below = n + 1 - (k0 + nb);                 // >= 1: the rhs row
    SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (chol_panel_kernel<<<div_up(rows_below, 128), 128, 0, ctx->stream>>>(A, n, k0, Linv)));
    const int cols = n - (k0 + nb);
    if (cols > 0) {
      const int tiles = div_up(rows_below, 64);
      SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (chol_syrk_kernel<<<tiles * (tiles + 1) / 2, 256, 0, ctx->stream>>>(A, n, k0, nb, Linv, info)));
    }
  }
  double* y = A + (size_t)n * n;
  for (int k0 = ((n - 1) / SB) * SB; k0 >= 0; k0 -= SB) {
    const int sb = std::min(SB, n - k0);
    SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (back_diag_kernel<<<1, 256, 0, ctx->stream>>>(A, n, k0, sb, Linv, y, x)));
    if (k0 > 0)
      SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (back_update_kernel<<<div_up(k0, 128), 128, 0, ctx->stream>>>(A, n, k0, sb, x, y)));
  }
  return SFM_OK;
}
