// solve.cu — dense symmetric-positive-definite solve of the reduced camera system (n = 6C; 3000 at C = 500) in
// float64: a tile Cholesky run as a dependency graph inside ONE persistent kernel, and the back substitution
// inside a second one — no launch per block step, no grid-wide barrier.
//
// Layout: the lower triangle of S in 64 x 64 tiles, each tile contiguous and COLUMN-major (element (r, c) at
// c * 64 + r), n padded to a multiple of 64 with an identity diagonal, plus one extra tile row whose row 0 is the
// right-hand side -g: the forward substitution L y = b is then just the bottom row of the factorisation.
// Column-major tiles make a finished tile L(i,k) directly the k-major operand of every later update.
//
// spd_factor_kernel (one CTA per SM, 256 threads).  Task = output tile (i, j), j <= i, claimed from an atomic
// counter in column-major order — every task depends only on tasks earlier in that order, which are finished or
// held by a running CTA, so spinning on their ready flags cannot deadlock.  A task is left-looking:
//   acc (4 x 4 per thread, registers) <- A(i,j);  for k < j: wait L(i,k), L(j,k);  acc -= L(i,k) L(j,k)^T
//   i == j: Cholesky of the 64 x 64 block (four 16-column panels, as before) and 1/diag
//   i >  j: wait L(j,j);  X L(j,j)^T = acc, one row per thread
//   store, __threadfence, release the tile's flag.
// Each tile is written once and read as an operand afterwards; the trailing matrix never makes the
// read-modify-write round trips through L2 that the right-looking version made at every step, and the ~100
// launches (47 block steps x (panel, update)) with their drains are gone.
//
// spd_backsolve_kernel (one CTA per tile column + one "diagonal" CTA).  L^T x = y from the bottom in super-blocks
// of 4 tile columns: column CTA c owns y_c (64 entries, in shared memory), subtracts L(rows of super-block s, c)^T
// x_s as soon as x_s is published, for every super-block below its own, then publishes y_c; the diagonal CTA
// waits for the four y_c of its super-block, solves the 256 x 256 triangle (64-row sub-steps by warp shuffles)
// and publishes x_s.  No atomics: a column's entries are only ever touched by its own CTA.
#include <math.h>

#include "common.cuh"
#include "solve.cuh"

namespace {

constexpr int T = 64;           // tile edge
constexpr int TT = T * T;
constexpr int SBT = 4;          // tile columns per back-substitution super-block

struct SpdPlan {
  int n, ntc, ntasks;
  double* tiles;                // tile-major, tiles column-major
  double* inv;                  // ntc tiles: inverses of the diagonal blocks L(c,c) (lower triangular, column-major)
  double* dinv;                 // 1 / diag(L), ntc * 64
  double* ybuf;                 // ntc * 64: y_c as published by the column CTAs
  double* xbuf;                 // ntc * 64: the solution
  int* flags;                   // one per tile (+ one per inverse): 1 = the final tile is in memory
  int* sync;                    // [0] task counter; [1 ..] back substitution flags: ycol_ready[ntc], x_ready[nsb]
  int* info;
  long long* stamps;            // diagnostics (SFM_SPD_TIMELINE): cycle counts of the phases of column `probe`'s critical tasks
  int probe;
  const int* skip;              // device word: 1 = the system is already solved (pcg.cu), every kernel returns at once
};

__host__ __device__ inline int tile_id(int i, int j, int ntc) { return (i < ntc ? i * (i + 1) / 2 : ntc * (ntc + 1) / 2) + j; }

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const int* p) {
  while (ld_acquire(p) == 0) __nanosleep(40);
}

// S (float32, lower triangle) and g -> float64 tiles.  grid = (tile id, 16 column groups), 256 threads.
__global__ void __launch_bounds__(256) spd_pack_kernel(const float* __restrict__ S, const float* __restrict__ g, SpdPlan p) {
  if (p.skip && *p.skip == 1) return;
  const int ntc = p.ntc, n = p.n;
  int t = blockIdx.x, i, j;
  const int ntri = ntc * (ntc + 1) / 2;
  if (t < ntri) {
    i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    j = t - i * (i + 1) / 2;
  } else {
    i = ntc;
    j = t - ntri;
  }
  double* dst = p.tiles + (size_t)t * TT;
  for (int e = threadIdx.x; e < TT; e += blockDim.x) {
    const int c = e >> 6, r = e & 63;                   // column-major within the tile
    const int gr = T * i + r, gc = T * j + c;
    double v;
    if (i == ntc) v = (r == 0 && gc < n) ? -(double)g[gc] : 0.0;
    else if (gr < n && gc < n)       // S arrives as the lower triangle of 6x6 blocks: block (a, b), b <= a, at (a (a+1) / 2 + b) * 36
      v = (gc <= gr) ? (double)S[((size_t)(gr / 6) * (gr / 6 + 1) / 2 + gc / 6) * 36 + 6 * (gr % 6) + gc % 6] : 0.0;
    else v = (gr == gc) ? 1.0 : 0.0;
    dst[e] = v;
  }
}

// 1 / sqrt(d) for a pivot known to be a positive normal number: the hardware seed and two Newton steps, without the
// range handling of rsqrt() — this sits on the factorisation's critical chain once per column.
__device__ __forceinline__ double rsqrt_normal(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double h = 0.5 * d;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}

// Cholesky of a 64 x 64 block held column-major in shared memory (C), by 256 threads: thread t holds row t%64,
// columns 16*(t/64) .. +15 in registers; four 16-column panels (inside a panel only its 64 owner threads work, two
// 64-thread named barriers per column; after a panel one block barrier and the rank-16 update of the columns to its
// right).  Writes L (lower, zeros above) column-major to `out` and 1/diag to dinv.  colbuf: 16 x (T+2) doubles.
__device__ __forceinline__ void factor_block(const double* __restrict__ C, double* __restrict__ out, double* __restrict__ dinv,
                                             int pivot_base, int* __restrict__ info, double (*colbuf)[T + 2]) {
  const int row = threadIdx.x & 63, cseg = threadIdx.x >> 6;
  double a[16];
#pragma unroll
  for (int cc = 0; cc < 16; ++cc) {
    const int c = 16 * cseg + cc;
    a[cc] = (c <= row) ? C[c * T + row] : 0.0;
  }
#pragma unroll 1
  for (int seg = 0; seg < 4; ++seg) {
    if (cseg == seg) {                                  // the panel's owners: two warps, rows 0..63
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const int j = 16 * seg + jj;
        double* cb = colbuf[jj];
        // ONE 64-thread barrier per column: every owner publishes its UNSCALED entry of column j, then reads the
        // diagonal and the entries it needs and scales them itself (1 / sqrt(d) is formed redundantly by all 64)
        cb[row] = a[jj];
        asm volatile("bar.sync 1, 64;" ::: "memory");
        double d = cb[j];
        if (!(d > 0.0)) {
          if (row == j && info && *info == 0) *info = pivot_base + j + 1;   // not positive definite
          d = 1.0;
        }
        const double rinv = rsqrt_normal(d);
        const double li = (row == j) ? d * rinv : ((row > j) ? a[jj] * rinv : 0.0);   // L[row][j]
        a[jj] = li;
        if (row == j) dinv[j] = rinv;
        const double lir = li * rinv;
#pragma unroll
        for (int cc = jj + 1; cc < 16; ++cc) {          // remaining columns of this panel: a_rc -= l_rj l_cj, l_cj = cb[c] rinv
          const int c = 16 * seg + cc;
          if (row >= c) a[cc] = fma(-lir, cb[c], a[cc]);
        }
      }
      // the scaled columns of the panel for the rank-16 update of the columns to the right
      asm volatile("bar.sync 1, 64;" ::: "memory");     // everybody is done reading the unscaled entries
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) colbuf[jj][row] = a[jj];
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
    out[c * T + row] = (c <= row) ? a[cc] : 0.0;
  }
}

constexpr int FACTOR_SMEM = (2 * TT + 16 * (T + 2) + T) * (int)sizeof(double);

__global__ void __launch_bounds__(256, 1) spd_factor_kernel(SpdPlan p) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                  // operand L(i,k), k-major: As[k * 64 + row]   (== the tile as stored)
  double* Bs = smem + TT;             // operand L(j,k)
  double (*colbuf)[T + 2] = reinterpret_cast<double(*)[T + 2]>(smem + 2 * TT);
  double* di = smem + 2 * TT + 16 * (T + 2);
  __shared__ int s_task;
  if (p.skip && *p.skip == 1) return;
  const int ntc = p.ntc;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;        // rows 4 ty .. +3, columns tx + 16 b of the tile
  for (;;) {
    if (threadIdx.x == 0) s_task = atomicAdd(p.sync, 1);
    __syncthreads();
    const int q = s_task;
    __syncthreads();
    if (q >= p.ntasks + ntc) return;
    if (q >= p.ntasks) {
      // ---- inverse of a diagonal block, for the back substitution (off the factorisation's critical path: these
      // tasks come last in the queue).  Column j of L^-1 by thread j: z_j = 1 / L_jj, z_i = -(sum_{m=j}^{i-1} L[i][m] z_m) / L_ii
      const int c = q - p.ntasks;
      const int id = tile_id(c, c, ntc);
      if (threadIdx.x == 0) wait_flag(p.flags + id);
      __syncthreads();
      {
        const double2* ga = reinterpret_cast<const double2*>(p.tiles + (size_t)id * TT);
        double2* sa = reinterpret_cast<double2*>(As);
#pragma unroll
        for (int e = 0; e < TT / 2 / 256; ++e) sa[threadIdx.x + 256 * e] = __ldcg(ga + threadIdx.x + 256 * e);
        if (threadIdx.x < T) di[threadIdx.x] = __ldcg(p.dinv + T * c + threadIdx.x);
      }
      __syncthreads();
      if (threadIdx.x < T) {
        const int j = threadIdx.x;                  // Bs[i * 64 + j] = (L^-1)[i][j]
#pragma unroll 1
        for (int i = 0; i < T; ++i) {
          double z = 0.0;
          if (i == j) z = di[j];
          else if (i > j) {
            double s0 = 0.0, s1 = 0.0;
            int m = j;
            for (; m + 1 < i; m += 2) {
              s0 = fma(As[m * T + i], Bs[m * T + j], s0);
              s1 = fma(As[(m + 1) * T + i], Bs[(m + 1) * T + j], s1);
            }
            if (m < i) s0 = fma(As[m * T + i], Bs[m * T + j], s0);
            z = -(s0 + s1) * di[i];
          }
          Bs[i * T + j] = z;
        }
      }
      __syncthreads();
      double* out = p.inv + (size_t)c * TT;
      for (int e = threadIdx.x; e < TT; e += 256) out[e] = Bs[(e & 63) * T + (e >> 6)];      // column-major: (i, j) at j * 64 + i
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) st_release(p.flags + p.ntasks + c, 1);
      continue;
    }
    // column-major task order: column j holds tiles (j, j) .. (ntc, j)
    int j = 0, rem = q;
    while (rem >= ntc + 1 - j) { rem -= ntc + 1 - j; ++j; }
    const int i = j + rem;
    const bool tl = p.stamps && j == p.probe && i <= j + 1 && threadIdx.x == 0;     // probe: tasks (j, j) and (j + 1, j)
    long long* st = p.stamps + (i == j ? 0 : 8);
    long long t_wait = 0, t_load = 0, t_mma = 0, t0 = 0;
    double* tile = p.tiles + (size_t)tile_id(i, j, ntc) * TT;
    double acc[4][4];                                             // acc[b][a]: column tx + 16 b, row 4 ty + a
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const double2 lo = *reinterpret_cast<const double2*>(tile + (tx + 16 * b) * T + 4 * ty);
      const double2 hi = *reinterpret_cast<const double2*>(tile + (tx + 16 * b) * T + 4 * ty + 2);
      acc[b][0] = lo.x; acc[b][1] = lo.y; acc[b][2] = hi.x; acc[b][3] = hi.y;
    }
#pragma unroll 1
    for (int k = 0; k < j; ++k) {
      const int ia = tile_id(i, k, ntc), ib = tile_id(j, k, ntc);
      if (tl) t0 = clock64();
      if (threadIdx.x == 0) {
        wait_flag(p.flags + ia);
        if (ib != ia) wait_flag(p.flags + ib);
      }
      __syncthreads();
      if (tl) { const long long t = clock64(); t_wait += t - t0; t0 = t; }
      const double* Bop = (ib == ia) ? As : Bs;                 // the diagonal task multiplies a tile by itself: one load
      {
        const double2* ga = reinterpret_cast<const double2*>(p.tiles + (size_t)ia * TT);
        const double2* gb = reinterpret_cast<const double2*>(p.tiles + (size_t)ib * TT);
        double2* sa = reinterpret_cast<double2*>(As);
        double2* sb = reinterpret_cast<double2*>(Bs);
        if (ib == ia) {
#pragma unroll
          for (int e = 0; e < TT / 2 / 256; ++e) sa[threadIdx.x + 256 * e] = __ldcg(ga + threadIdx.x + 256 * e);
        } else {
#pragma unroll
          for (int e = 0; e < TT / 2 / 256; ++e) {
            sa[threadIdx.x + 256 * e] = __ldcg(ga + threadIdx.x + 256 * e);
            sb[threadIdx.x + 256 * e] = __ldcg(gb + threadIdx.x + 256 * e);
          }
        }
      }
      __syncthreads();
      if (tl) { const long long t = clock64(); t_load += t - t0; t0 = t; }
#pragma unroll 4
      for (int kk = 0; kk < T; ++kk) {
        const double2 a01 = *reinterpret_cast<const double2*>(As + kk * T + 4 * ty);
        const double2 a23 = *reinterpret_cast<const double2*>(As + kk * T + 4 * ty + 2);
        const double av[4] = {a01.x, a01.y, a23.x, a23.y};
        const double bv[4] = {Bop[kk * T + tx], Bop[kk * T + tx + 16], Bop[kk * T + tx + 32], Bop[kk * T + tx + 48]};
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int a = 0; a < 4; ++a) acc[b][a] = fma(-av[a], bv[b], acc[b][a]);
      }
      __syncthreads();
      if (tl) { const long long t = clock64(); t_mma += t - t0; t0 = t; }
    }
    if (tl) { st[0] = t_wait; st[1] = t_load; st[2] = t_mma; st[3] = j; t0 = clock64(); }
    // the updated block, column-major, into As
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      *reinterpret_cast<double2*>(As + (tx + 16 * b) * T + 4 * ty) = make_double2(acc[b][0], acc[b][1]);
      *reinterpret_cast<double2*>(As + (tx + 16 * b) * T + 4 * ty + 2) = make_double2(acc[b][2], acc[b][3]);
    }
    if (i == j) {
      __syncthreads();
      factor_block(As, tile, p.dinv + T * j, T * j, p.info, colbuf);
      if (tl) st[4] = clock64() - t0;
    } else {
      const int id = tile_id(j, j, ntc);
      if (threadIdx.x == 0) wait_flag(p.flags + id);
      __syncthreads();
      if (tl) { const long long t = clock64(); st[4] = t - t0; t0 = t; }
      {
        const double2* gb = reinterpret_cast<const double2*>(p.tiles + (size_t)id * TT);
        double2* sb = reinterpret_cast<double2*>(Bs);
#pragma unroll
        for (int e = 0; e < TT / 2 / 256; ++e) sb[threadIdx.x + 256 * e] = __ldcg(gb + threadIdx.x + 256 * e);
        if (threadIdx.x < T) di[threadIdx.x] = __ldcg(p.dinv + T * j + threadIdx.x);
      }
      __syncthreads();
      // X L^T = A: row r of X by thread r, x_c = (a_c - sum_{m<c} x_m L[c][m]) / L[c][c]; the running row in
      // registers, L[c][m] = Bs[m * 64 + c] read as a broadcast; four chains per sum
      if (threadIdx.x < T) {
        const int r = threadIdx.x;
        double x[T];
#pragma unroll
        for (int c = 0; c < T; ++c) x[c] = As[c * T + r];
#pragma unroll
        for (int c = 0; c < T; ++c) {
          double s0 = x[c], s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
          for (int m = 0; m < c; ++m) {
            if ((m & 3) == 0) s0 = fma(-x[m], Bs[m * T + c], s0);
            else if ((m & 3) == 1) s1 = fma(-x[m], Bs[m * T + c], s1);
            else if ((m & 3) == 2) s2 = fma(-x[m], Bs[m * T + c], s2);
            else s3 = fma(-x[m], Bs[m * T + c], s3);
          }
          x[c] = ((s0 + s1) + (s2 + s3)) * di[c];
        }
#pragma unroll
        for (int c = 0; c < T; ++c) tile[c * T + r] = x[c];
      }
      if (tl) st[5] = clock64() - t0;
    }
    if (tl) t0 = clock64();
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release(p.flags + tile_id(i, j, ntc), 1);
    if (tl) st[6] = clock64() - t0;
  }
}

// grid = 1 diagonal CTA (block 0) + ntc column CTAs, all co-resident (ntc + 1 <= SM count is required).
__global__ void __launch_bounds__(256, 1) spd_backsolve_kernel(SpdPlan p) {
  if (p.skip && *p.skip == 1) return;
  const int ntc = p.ntc, nsb = (ntc + SBT - 1) / SBT;
  int* ycol_ready = p.sync + 1;
  int* x_ready = p.sync + 1 + ntc;
  __shared__ double ys[SBT * T];
  __shared__ double xs[SBT * T];
  if (blockIdx.x > 0) {
    // ---- column CTA c: y_c -= L(tr, c)^T x_tr for every tile row tr of every super-block below its own
    const int c = blockIdx.x - 1, mysb = c / SBT;
    const double* rhs = p.tiles + (size_t)tile_id(ntc, c, ntc) * TT;       // row 0 of the rhs tile: element (0, col) at col * 64
    if (threadIdx.x < T) ys[threadIdx.x] = rhs[threadIdx.x * T];
    const int col = threadIdx.x >> 2, part = threadIdx.x & 3;
    for (int sb = nsb - 1; sb > mysb; --sb) {
      if (threadIdx.x == 0) wait_flag(x_ready + sb);
      __syncthreads();
      const int tr0 = sb * SBT, ntr = min(SBT, ntc - tr0);
      for (int e = threadIdx.x; e < ntr * T; e += blockDim.x) xs[e] = __ldcg(p.xbuf + T * tr0 + e);
      __syncthreads();
      double s = 0.0;
      for (int q = 0; q < ntr; ++q) {
        const double* Lt = p.tiles + (size_t)tile_id(tr0 + q, c, ntc) * TT + col * T + part * 16;
        const double* xv = xs + q * T + part * 16;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int r = 0; r < 16; r += 4) {
          const double2 l01 = __ldcg(reinterpret_cast<const double2*>(Lt + r));
          const double2 l23 = __ldcg(reinterpret_cast<const double2*>(Lt + r + 2));
          s0 = fma(l01.x, xv[r], s0); s1 = fma(l01.y, xv[r + 1], s1);
          s0 = fma(l23.x, xv[r + 2], s0); s1 = fma(l23.y, xv[r + 3], s1);
        }
        s += s0 + s1;
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (part == 0) ys[col] -= s;
      __syncthreads();
    }
    if (threadIdx.x < T) p.ybuf[T * c + threadIdx.x] = ys[threadIdx.x];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release(ycol_ready + c, 1);
    return;
  }
  // ---- diagonal CTA: the super-blocks from the bottom up
  for (int sb = nsb - 1; sb >= 0; --sb) {
    const int c0 = sb * SBT, nc = min(SBT, ntc - c0);
    if ((int)threadIdx.x < nc) {
      wait_flag(ycol_ready + c0 + threadIdx.x);
      wait_flag(p.flags + p.ntasks + c0 + threadIdx.x);          // (set by the factorisation kernel, long since)
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nc * T; e += blockDim.x) ys[e] = __ldcg(p.ybuf + T * c0 + e);
    const int col = threadIdx.x >> 2, part = threadIdx.x & 3;
    for (int cc = nc - 1; cc >= 0; --cc) {
      const int c = c0 + cc;
      __syncthreads();
      // x_c = L(c,c)^-T v: x_j = sum_{i >= j} (L^-1)[i][j] v_i, four lanes per unknown, the inverse read straight from
      // L2 (column j of the tile is contiguous)
      {
        const double* Li = p.inv + (size_t)c * TT + col * T;
        const double* v = ys + cc * T;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int m = 0; m < 16; m += 2) {
          s0 = fma(__ldcg(Li + part + 4 * m), v[part + 4 * m], s0);
          s1 = fma(__ldcg(Li + part + 4 * m + 4), v[part + 4 * m + 4], s1);
        }
        double sx = s0 + s1;
        sx += __shfl_xor_sync(0xffffffffu, sx, 1);
        sx += __shfl_xor_sync(0xffffffffu, sx, 2);
        if (part == 0) xs[cc * T + col] = sx;
      }
      __syncthreads();
      // earlier columns of this super-block: y_{c'} -= L(c, c')^T x_c
      const double* xo = xs + cc * T;
      for (int cp = 0; cp < cc; ++cp) {
        const double* Lt = p.tiles + (size_t)tile_id(c, c0 + cp, ntc) * TT + col * T;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int m = 0; m < 16; m += 2) {
          s0 = fma(__ldcg(Lt + part + 4 * m), xo[part + 4 * m], s0);
          s1 = fma(__ldcg(Lt + part + 4 * m + 4), xo[part + 4 * m + 4], s1);
        }
        double sy = s0 + s1;
        sy += __shfl_xor_sync(0xffffffffu, sy, 1);
        sy += __shfl_xor_sync(0xffffffffu, sy, 2);
        if (part == 0) ys[cp * T + col] -= sy;
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nc * T; e += blockDim.x) p.xbuf[T * c0 + e] = xs[e];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release(x_ready + sb, 1);
  }
}

__global__ void spd_copy_x_kernel(const double* __restrict__ xbuf, int n, double* __restrict__ x, const int* __restrict__ skip) {
  if (skip && *skip == 1) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = xbuf[i];
}

static SpdPlan make_plan(int n, double* A, int* info) {
  SpdPlan p;
  p.n = n;
  p.ntc = div_up(n, T);
  p.ntasks = p.ntc * (p.ntc + 1) / 2 + p.ntc;
  const size_t ntiles = (size_t)p.ntasks;
  p.tiles = A;
  p.inv = A + ntiles * TT;
  p.dinv = p.inv + (size_t)p.ntc * TT;
  p.ybuf = p.dinv + (size_t)p.ntc * T;
  p.xbuf = p.ybuf + (size_t)p.ntc * T;
  p.flags = reinterpret_cast<int*>(p.xbuf + (size_t)p.ntc * T);
  p.sync = p.flags + ntiles + p.ntc;
  p.info = info;
  p.stamps = nullptr;
  p.probe = 0;
  p.skip = nullptr;
  return p;
}

}  // namespace

size_t sfm_spd_scratch_doubles(int n) {
  const size_t ntc = (size_t)div_up(n, T);
  const size_t ntiles = ntc * (ntc + 1) / 2 + ntc;
  const size_t ints = ntiles + ntc + 1 + ntc + (ntc + SBT - 1) / SBT + 16;
  return (ntiles + ntc) * TT + 3 * ntc * T + (ints + 1) / 2;
}

int sfm_spd_solve(sfm_ctx* ctx, const float* S, const float* g, int n, double* A, double* x, int* info, const int* skip_if) {
  SpdPlan p = make_plan(n, A, info);
  p.skip = skip_if;
  SFM_REQUIRE(p.ntc + 1 <= ctx->sm_count, "sfm_spd_solve: %d unknowns need %d co-resident CTAs, the device has %d SMs", n,
              p.ntc + 1, ctx->sm_count);
  const size_t nints = (size_t)p.ntasks + p.ntc + 1 + p.ntc + (p.ntc + SBT - 1) / SBT;
  SFM_CUDA(cudaMemsetAsync(p.flags, 0, nints * sizeof(int), ctx->stream));
  if (!skip_if) SFM_CUDA(cudaMemsetAsync(info, 0, sizeof(int), ctx->stream));      // (behind the CG solver: it has written info)
  SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (spd_pack_kernel<<<p.ntasks, 256, 0, ctx->stream>>>(S, g, p)));
  static bool attr_set = false;
  if (!attr_set) {
    SFM_CUDA(cudaFuncSetAttribute(spd_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FACTOR_SMEM));
    attr_set = true;
  }
  const int grid = std::min(ctx->sm_count, p.ntasks + p.ntc);
  long long* dstamps = nullptr;
  if (getenv("SFM_SPD_TIMELINE")) {
    SFM_TRY(ws_alloc_t(ctx, 16, &dstamps));
    SFM_CUDA(cudaMemsetAsync(dstamps, 0, 16 * sizeof(long long), ctx->stream));
    p.stamps = dstamps;
    p.probe = std::min(p.ntc - 2, std::max(1, atoi(getenv("SFM_SPD_TIMELINE"))));
  }
  SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (spd_factor_kernel<<<grid, 256, FACTOR_SMEM, ctx->stream>>>(p)));
  if (dstamps) {
    long long h[16];
    SFM_CUDA(cudaMemcpyAsync(h, dstamps, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(cudaStreamSynchronize(ctx->stream));
    fprintf(stderr, "[spd column %lld] diagonal task: wait %lld | load %lld | update %lld | cholesky %lld | publish %lld  ||  tile below: wait %lld | load %lld | update %lld | "
            "wait for L(k,k) %lld | triangular solve (incl. load) %lld | publish %lld   (cycles, %lld update steps each)\n",
            h[3], h[0], h[1], h[2], h[4], h[6], h[8], h[9], h[10], h[12], h[13], h[14], h[3]);
  }
  SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (spd_backsolve_kernel<<<p.ntc + 1, 256, 0, ctx->stream>>>(p)));
  SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (spd_copy_x_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(p.xbuf, n, x, p.skip)));
  return SFM_OK;
}

extern "C" int sfm_reduced_solve(sfm_ctx* ctx, const float* S_blocks, const float* g, int n_cams, double* x, int32_t* info) {
  SFM_REQUIRE(ctx && S_blocks && g && x && info, "sfm_reduced_solve: null argument");
  SFM_REQUIRE(n_cams >= 1, "sfm_reduced_solve: no cameras");
  SFM_TRY(sfm_ws_begin(ctx));
  const int n = 6 * n_cams;
  const size_t nblk = (size_t)n_cams * (n_cams + 1) / 2;
  const float *dS, *dg;
  SFM_TRY(dev_in(ctx, S_blocks, nblk * 36, &dS));
  SFM_TRY(dev_in(ctx, g, (size_t)n, &dg));
  double *A, *dx;
  int* dinfo;
  SFM_TRY(ws_alloc_t(ctx, sfm_spd_scratch_doubles(n), &A));
  SFM_TRY(ws_alloc_t(ctx, (size_t)n, &dx));
  SFM_TRY(ws_alloc_t(ctx, 1, &dinfo));
  SFM_TRY(sfm_spd_solve(ctx, dS, dg, n, A, dx, dinfo));
  SFM_CUDA(cudaMemcpyAsync(x, dx, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(cudaMemcpyAsync(info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFM_OK;
}

extern "C" int sfm_reduced_solve_pcg(sfm_ctx* ctx, const float* S_blocks, const float* g, int n_cams, double* x, int32_t* solved,
                                     int32_t* iterations) {
  SFM_REQUIRE(ctx && S_blocks && g && x && solved, "sfm_reduced_solve_pcg: null argument");
  SFM_REQUIRE(n_cams >= 1, "sfm_reduced_solve_pcg: no cameras");
  const int n = 6 * n_cams;
  SFM_REQUIRE(sfm_spd_pcg_fits(ctx, n), "sfm_reduced_solve_pcg: %d cameras exceed the shared-memory-resident vectors", n_cams);
  SFM_TRY(sfm_ws_begin(ctx));
  const size_t nblk = (size_t)n_cams * (n_cams + 1) / 2;
  const float *dS, *dg;
  SFM_TRY(dev_in(ctx, S_blocks, nblk * 36, &dS));
  SFM_TRY(dev_in(ctx, g, (size_t)n, &dg));
  double *scratch, *dx;
  int* dwords;
  SFM_TRY(ws_alloc_t(ctx, sfm_pcg_scratch_doubles(n), &scratch));
  SFM_TRY(ws_alloc_t(ctx, (size_t)n, &dx));
  SFM_TRY(ws_alloc_t(ctx, 4, &dwords));
  SFM_CUDA(cudaMemsetAsync(dx, 0, sizeof(double) * n, ctx->stream));
  SFM_CUDA(cudaMemsetAsync(dwords, 0, 4 * sizeof(int), ctx->stream));
  SFM_TRY(sfm_spd_pcg(ctx, dS, dg, n, scratch, dx, dwords, dwords + 1, dwords + 2, 1e-8));
  int h[4];
  SFM_CUDA(cudaMemcpyAsync(x, dx, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(cudaMemcpyAsync(h, dwords, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  *solved = h[0];
  if (iterations) *iterations = h[2];
  return SFM_OK;
}
