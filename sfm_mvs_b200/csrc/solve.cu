// solve.cu — dense symmetric-positive-definite solve of the reduced camera system (n = 6C; 3000 at
// C = 500), float64, hand-written right-looking blocked Cholesky, NB = 64, with look-ahead.
//
//   A is (n+1) x n row-major, lower triangle used; row n carries the right-hand side, so the
//   forward substitution L y = b falls out of the panel solves for free (row n ends up as y^T).
//   The sequential part of a step — the 64x64 diagonal block — is factored by a whole CTA with the
//   block distributed over registers (thread = one row x 16 columns) and ONE block barrier per
//   column; it runs as a look-ahead inside the trailing update of the previous step (the CTA that
//   owns the first tile of the update owns exactly the next diagonal block), so it overlaps the
//   rest of that update.
//     chol_diag_kernel   block 0 only
//     chol_panel_kernel  rows below block k (incl. the rhs row): x L_kk^T = a, one row per thread,
//                        L_kk and 1/diag staged in shared memory
//     chol_syrk_kernel   A22 -= L21 L21^T, 64x64 tile per CTA, 4x4 per thread, K = 64 in two
//                        32-wide shared-memory stages, lower-triangular tiles only; tile 0 then
//                        factors diagonal block k+1
//   backward substitution L^T x = y, 256-row super-blocks from the bottom up:
//     back_diag_kernel   one CTA: four 64-row sub-steps (operands staged in shared memory with
//                        coalesced loads; one barrier per unknown inside a sub-step)
//     back_update_kernel earlier entries: y[i] -= sum_r L[k0+r][i] x[k0+r], 8 threads per entry
#include <math.h>

#include "common.cuh"
#include "solve.cuh"

namespace {

constexpr int NB = 64;
constexpr int SB = 256;   // backward-substitution super-block

// Cholesky of the diagonal block at k0 by a 256-thread CTA.  Thread t holds row t%64, columns
// 16*(t/64) .. +15 in registers.  The block is factored in four 16-column panels:
//   inside a panel only its 64 owner threads work — per column two 64-thread named barriers (publish the
//   diagonal, publish the scaled column) and <= 15 FMAs per thread for the columns of the same panel;
//   after a panel ONE block barrier, then the threads holding columns to its right apply the whole rank-16
//   update from shared memory (256 independent FMAs per thread).
// 4 block barriers instead of 64, and the serial part is 64 x (two cheap barriers + a handful of FMAs).
// Rows/columns >= nb are identity.  Writes L (lower) back to A and 1/diag(L) to dinv[k0 .. k0+63].
// colbuf: 16 x (NB+2) doubles of shared memory.
__device__ __forceinline__ void factor_diag_block(double* __restrict__ A, int n, int k0, double* __restrict__ dinv,
                                                  int* __restrict__ info, double (*colbuf)[NB + 2]) {
  const int nb = min(NB, n - k0);
  const int row = threadIdx.x & 63, cseg = threadIdx.x >> 6;
  double a[16];
#pragma unroll
  for (int cc = 0; cc < 16; ++cc) {
    const int c = 16 * cseg + cc;
    a[cc] = (row < nb && c <= row) ? A[(size_t)(k0 + row) * n + k0 + c] : ((c == row) ? 1.0 : 0.0);
  }
#pragma unroll 1
  for (int seg = 0; seg < 4; ++seg) {
    if (cseg == seg) {                                  // the panel's owners: two warps, rows 0..63
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const int j = 16 * seg + jj;
        double* cb = colbuf[jj];                        // cb[0..63] scaled column j, cb[64] raw diagonal
        if (row == j) cb[NB] = a[jj];
        asm volatile("bar.sync 1, 64;" ::: "memory");
        double d = cb[NB];
        if (j < nb && !(d > 0.0)) {
          if (row == j && info && *info == 0) *info = k0 + j + 1;   // not positive definite
          d = 1.0;
        }
        const double rinv = rsqrt(d);
        const double li = (row == j) ? d * rinv : ((row > j) ? a[jj] * rinv : 0.0);   // L[row][j]
        a[jj] = li;
        cb[row] = li;
        if (row == j) dinv[k0 + j] = rinv;
        asm volatile("bar.sync 1, 64;" ::: "memory");
#pragma unroll
        for (int cc = jj + 1; cc < 16; ++cc) {          // remaining columns of this panel
          const int c = 16 * seg + cc;
          if (row >= c) a[cc] = fma(-li, cb[c], a[cc]);
        }
      }
    }
    __syncthreads();
    if (cseg > seg) {                                   // rank-16 update of the columns to the right
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const double li = colbuf[jj][row];
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {
          const int c = 16 * cseg + cc;
          if (row >= c) a[cc] = fma(-li, colbuf[jj][c], a[cc]);
        }
      }
    }
    __syncthreads();                                    // colbuf is rewritten by the next panel
  }
#pragma unroll
  for (int cc = 0; cc < 16; ++cc) {
    const int c = 16 * cseg + cc;
    if (row < nb && c <= row) A[(size_t)(k0 + row) * n + k0 + c] = a[cc];
  }
}

__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ A, int n, int k0, double* __restrict__ dinv,
                                                        int* __restrict__ info) {
  __shared__ double colbuf[16][NB + 2];
  factor_diag_block(A, n, k0, dinv, info, colbuf);
}

// rows k0+nb .. n (the last one is the rhs row): solve x L_kk^T = a, one row per thread, the running
// solution in registers (fully unrolled), L_kk (strictly lower) and 1/diag read as shared broadcasts.
constexpr int PANEL_THREADS = 128;
__global__ void __launch_bounds__(PANEL_THREADS) chol_panel_kernel(double* __restrict__ A, int n, int k0,
                                                                   const double* __restrict__ dinv) {
  __shared__ double L[NB][NB + 1];
  __shared__ double di[NB];
  const int nb = min(NB, n - k0);
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    L[r][c] = (r < nb && c < r) ? A[(size_t)(k0 + r) * n + k0 + c] : 0.0;
  }
  if (threadIdx.x < NB) di[threadIdx.x] = (threadIdx.x < nb) ? dinv[k0 + threadIdx.x] : 1.0;
  __syncthreads();
  const int row = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (row > n) return;
  double* a = A + (size_t)row * n + k0;
  double x[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) x[j] = (j < nb) ? a[j] : 0.0;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    double s0 = x[j], s1 = 0.0, s2 = 0.0, s3 = 0.0;    // four chains: the sum over k < j is latency-bound otherwise
#pragma unroll
    for (int k = 0; k < j; ++k) {
      if ((k & 3) == 0) s0 = fma(-x[k], L[j][k], s0);
      else if ((k & 3) == 1) s1 = fma(-x[k], L[j][k], s1);
      else if ((k & 3) == 2) s2 = fma(-x[k], L[j][k], s2);
      else s3 = fma(-x[k], L[j][k], s3);
    }
    x[j] = ((s0 + s1) + (s2 + s3)) * di[j];
  }
#pragma unroll
  for (int j = 0; j < NB; ++j)
    if (j < nb) a[j] = x[j];
}

// rows [base, n] (n+1-base of them, the last is the rhs row), columns [base, n)
__global__ void __launch_bounds__(256) chol_syrk_kernel(double* __restrict__ A, int n, int k0, int nb,
                                                        double* __restrict__ dinv, int* __restrict__ info) {
  constexpr int KC = 32;
  __shared__ double Pi[64][KC + 1];
  __shared__ double Pj[64][KC + 1];
  int t = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  const int base = k0 + nb;
  const int i0 = base + ti * 64, j0 = base + tj * 64;
  if (j0 >= n) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int kc = 0; kc < nb; kc += KC) {
    if (kc) __syncthreads();
    for (int e = threadIdx.x; e < 64 * KC; e += blockDim.x) {
      const int r = e / KC, c = e % KC;
      Pi[r][c] = (i0 + r <= n && kc + c < nb) ? A[(size_t)(i0 + r) * n + k0 + kc + c] : 0.0;
      Pj[r][c] = (j0 + r < n && kc + c < nb) ? A[(size_t)(j0 + r) * n + k0 + kc + c] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < KC; ++k) {
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
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
      if (i <= n && j < n && j <= i) A[(size_t)i * n + j] -= acc[a][b];
    }
  // look-ahead: tile 0 of the update is exactly diagonal block k+1 and now holds its final values
  if (t == 0) {
    __syncthreads();
    factor_diag_block(A, n, base, dinv, info, reinterpret_cast<double(*)[NB + 2]>(&Pi[0][0]));
  }
}

// One CTA per call: rows [k0, k0+sb) of L^T x = y in 64-row sub-steps from the bottom.
__global__ void __launch_bounds__(256) back_diag_kernel(const double* __restrict__ A, int n, int k0, int sb,
                                                        const double* __restrict__ dinv, double* __restrict__ y,
                                                        double* __restrict__ x) {
  __shared__ double ys[SB];
  __shared__ double Ls[NB][NB + 1];
  __shared__ double xs[NB];
  for (int i = threadIdx.x; i < SB; i += blockDim.x) ys[i] = (i < sb) ? y[k0 + i] : 0.0;
  const int nsub = (sb + NB - 1) / NB;
  for (int s = nsub - 1; s >= 0; --s) {
    const int r0 = k0 + s * NB;                       // global row of the sub-block
    const int nb = min(NB, n - r0);
    __syncthreads();
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
      const int r = e / NB, c = e % NB;
      Ls[r][c] = (r < nb && c < r) ? A[(size_t)(r0 + r) * n + r0 + c] : 0.0;
    }
    __syncthreads();
    // x_j = (v_j - sum_{i>j} L[i][j] x_i) / L_jj, unknown j owned by thread j (two warps)
    double v = 0.0, di = 1.0;
    if (threadIdx.x < NB) {
      v = ys[s * NB + threadIdx.x];
      di = (threadIdx.x < nb) ? dinv[r0 + threadIdx.x] : 1.0;
    }
    // 64 unknowns, bottom up.  Unknown j lives in lane j%32 of warp j/32; inside a 32-unknown half the solved value
    // travels by shuffle (no barrier), the upper half's effect on the lower one is a 32x32 matvec.
    if (threadIdx.x < NB) xs[threadIdx.x] = 0.0;
    __syncthreads();
    if (threadIdx.x >= 32 && threadIdx.x < NB) {       // upper half (unknowns 32..63), warp 1
      const int l = threadIdx.x - 32;
      for (int j = 31; j >= 0; --j) {
        const double xj = __shfl_sync(0xffffffffu, v * di, j);
        if (l == j) xs[32 + j] = xj;
        if (l < j) v = fma(-Ls[32 + j][32 + l], xj, v);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {                            // lower half (unknowns 0..31), warp 0
      const int l = threadIdx.x;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
      for (int r = 0; r < 32; r += 2) {
        s0 = fma(Ls[32 + r][l], xs[32 + r], s0);
        s1 = fma(Ls[33 + r][l], xs[33 + r], s1);
      }
      v -= s0 + s1;
      for (int j = 31; j >= 0; --j) {
        const double xj = __shfl_sync(0xffffffffu, v * di, j);
        if (l == j) xs[j] = xj;
        if (l < j) v = fma(-Ls[j][l], xj, v);
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < nb) x[r0 + threadIdx.x] = xs[threadIdx.x];
    // earlier rows of this super-block: y[i] -= sum_r L[r0+r][k0+i] x_s[r]
    for (int i = threadIdx.x; i < s * NB; i += blockDim.x) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int r = 0;
      for (; r + 3 < nb; r += 4) {
        a0 = fma(A[(size_t)(r0 + r) * n + k0 + i], xs[r], a0);
        a1 = fma(A[(size_t)(r0 + r + 1) * n + k0 + i], xs[r + 1], a1);
        a2 = fma(A[(size_t)(r0 + r + 2) * n + k0 + i], xs[r + 2], a2);
        a3 = fma(A[(size_t)(r0 + r + 3) * n + k0 + i], xs[r + 3], a3);
      }
      for (; r < nb; ++r) a0 = fma(A[(size_t)(r0 + r) * n + k0 + i], xs[r], a0);
      ys[i] -= (a0 + a1) + (a2 + a3);
    }
  }
}

// y[i] -= sum_{r < sb} L[k0+r][i] x[k0+r] for i < k0; 8 threads per entry (rows r = q, q+8, ...),
// 32 entries per CTA, so a warp-level load covers 32 consecutive doubles of one row of L.
__global__ void __launch_bounds__(256) back_update_kernel(const double* __restrict__ A, int n, int k0, int sb,
                                                          const double* __restrict__ x, double* __restrict__ y) {
  __shared__ double xs[SB];
  __shared__ double part[8][33];
  for (int i = threadIdx.x; i < sb; i += blockDim.x) xs[i] = x[k0 + i];
  __syncthreads();
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  double a0 = 0.0, a1 = 0.0;
  if (i < k0) {
    int r = q;
    for (; r + 8 < sb; r += 16) {
      a0 = fma(A[(size_t)(k0 + r) * n + i], xs[r], a0);
      a1 = fma(A[(size_t)(k0 + r + 8) * n + i], xs[r + 8], a1);
    }
    for (; r < sb; r += 8) a0 = fma(A[(size_t)(k0 + r) * n + i], xs[r], a0);
  }
  part[q][lane] = a0 + a1;
  __syncthreads();
  if (q == 0 && i < k0) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][lane];
    y[i] -= s;
  }
}

__global__ void widen_kernel(const float* __restrict__ S, const float* __restrict__ g, int n, double* __restrict__ A) {
  const int r = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const size_t idx = (size_t)r * n + c;
  A[idx] = (r == n) ? -(double)g[c] : ((c <= r) ? (double)S[idx] : 0.0);
}

}  // namespace

size_t sfm_spd_scratch_doubles(int n) {
  return ((size_t)n + 1) * n + (size_t)div_up(n, NB) * NB;
}

int sfm_spd_solve(sfm_ctx* ctx, const float* S, const float* g, int n, double* A, double* x, int* info) {
  const size_t total = (size_t)(n + 1) * n;
  double* dinv = A + total;
  SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (widen_kernel<<<dim3(div_up(n, 256), n + 1), 256, 0, ctx->stream>>>(S, g, n, A)));
  SFM_CUDA(cudaMemsetAsync(info, 0, sizeof(int), ctx->stream));
  SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (chol_diag_kernel<<<1, 256, 0, ctx->stream>>>(A, n, 0, dinv, info)));
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int nb = std::min(NB, n - k0);
    const int rows_below = n + 1 - (k0 + nb);                 // >= 1: the rhs row
    SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (chol_panel_kernel<<<div_up(rows_below, PANEL_THREADS), PANEL_THREADS, 0, ctx->stream>>>(A, n, k0, dinv)));
    const int cols = n - (k0 + nb);
    if (cols > 0) {
      const int tiles = div_up(rows_below, 64);
      SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (chol_syrk_kernel<<<tiles * (tiles + 1) / 2, 256, 0, ctx->stream>>>(A, n, k0, nb, dinv, info)));
    }
  }
  double* y = A + (size_t)n * n;
  for (int k0 = ((n - 1) / SB) * SB; k0 >= 0; k0 -= SB) {
    const int sb = std::min(SB, n - k0);
    SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (back_diag_kernel<<<1, 256, 0, ctx->stream>>>(A, n, k0, sb, dinv, y, x)));
    if (k0 > 0)
      SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (back_update_kernel<<<div_up(k0, 32), 256, 0, ctx->stream>>>(A, n, k0, sb, x, y)));
  }
  return SFM_OK;
}
