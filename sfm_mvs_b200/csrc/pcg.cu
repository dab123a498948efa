// pcg.cu — the reduced camera system S x = -g by block-Jacobi preconditioned conjugate gradients, in ONE persistent
// kernel.  The tile Cholesky (solve.cu) is a chain of n = 6C dependent columns (3000 at C = 500: ~1.1 ms however many
// SMs wait on it); S of a damped LM step is well served by its 6x6 diagonal blocks as a preconditioner — measured on
// BASELINE configs[3] (tools/pcg_probe.py): cond(S) 7e5 .. 1e9 along the LM iterations, 26 - 64 iterations to a relative
// residual of 1e-8 (solution within 1e-7 of LAPACK's) — and an iteration is one symmetric matrix-vector product over the
// 18 MB of S (L2-resident) plus 3000-long vector updates.
//
// S arrives as the packed lower triangle of 6x6 blocks (block (a, b), b <= a, 36 row-major float32 at (a (a+1) / 2 + b) * 36),
// exactly what the Schur kernel wrote and the all-reduce summed; nothing is repacked.
//
//   matrix-vector product: block rows are dealt round-robin over the CTAs (row a -> CTA a mod grid: three or four rows
//     per SM at C = 500) and a CTA cuts each of its rows into column ranges, one (row, range) per WARP: y_a += B_ab p_b
//     for the stored blocks b < a (contiguous in memory) and y_a += B_ba^T p_b for b > a (every stored block is read
//     twice per product, from L2).  The lanes of a warp lie over the 16-byte pieces of three blocks per step, so a
//     step of the row part is one contiguous 432-byte request; loads are issued eight steps deep.  One shuffle
//     reduction per range, the ranges of a row summed in shared memory in a fixed order, S p written as ONE vector of n
//     doubles (parity-double-buffered) — no atomics, bit-reproducible.
//   ONE grid barrier per iteration (arrival counter + generation word, acquire / release, both monotonic).
//   vector part: x, p and the inverted diagonal blocks live in EVERY CTA's shared memory, and every CTA performs the
//     identical updates and the identical (fixed-tree) dot products — a thread OWNS the block rows tid, tid + 256, ..:
//     its residual stays in registers, x and p in its own shared-memory slots — so alpha, beta and the convergence
//     verdict need no second grid barrier and no broadcast: all CTAs hold bit-identical state and leave the loop together.
//
// Failure (a diagonal block or p^T S p not positive, no convergence in max_iter iterations, n too large for the
// vectors to fit in shared memory) is reported in *status; the caller then runs the Cholesky path (sfm_spd_solve takes
// the status word and its kernels return at once when it says "solved").
#include <math.h>

#include "common.cuh"
#include "solve.cuh"

namespace {

constexpr int PCG_THREADS = 256;
constexpr int PCG_WARPS = PCG_THREADS / 32;
constexpr int PCG_NB = 3;                 // block rows of the vectors a thread owns: C <= 256 * 3

struct PcgPlan {
  int n, C, max_iter;
  double tol2;                  // (relative residual)^2
  const float* S;
  const float* g;
  double* x;
  double* w;                    // [2][n]: S p, parity-double-buffered
  unsigned int* bar;            // [0] arrivals, [1] barriers completed
  int* status;                  // 1: x holds the solution; 0: not solved
  int* info;                    // the LM step's solve_info word: zeroed on success
  int* iters;
  long long* stamps;            // diagnostics (SFM_PCG_TIMELINE): cycles of CTA 0 in the product, the barrier, the vector part; iterations
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Every CTA of the (co-resident) grid: one arrival counter (bar[0]) and one generation word (bar[1]) that thread 0
// polls; both are monotonic within a launch (zeroed before it), nothing is reset.  (Measured against per-CTA flag words
// that every CTA polls — no atomics, but 148 CTAs polling five lines: 11.0 k cycles per barrier against 4.7 k.)
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(&bar[0], 1u);
    if (prev == gen * gridDim.x - 1) st_release_u32(&bar[1], gen);
    else
      while (ld_acquire_u32(&bar[1]) < gen) {}
  }
  __syncthreads();
}

// the same value in every thread, fixed summation order; `red` (PCG_WARPS doubles per value) must not be in use by an
// earlier call that some warp has not left yet: the call sites alternate between disjoint areas
__device__ __forceinline__ double block_sum1(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < PCG_WARPS; ++w) s += red[w];
  return s;
}
__device__ __forceinline__ void block_sum2(double v0, double v1, double* red /*2 PCG_WARPS*/, double& s0, double& s1) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v0 += __shfl_xor_sync(0xffffffffu, v0, o);
    v1 += __shfl_xor_sync(0xffffffffu, v1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = v0;
    red[PCG_WARPS + (threadIdx.x >> 5)] = v1;
  }
  __syncthreads();
  s0 = 0.0;
  s1 = 0.0;
#pragma unroll
  for (int w = 0; w < PCG_WARPS; ++w) {
    s0 += red[w];
    s1 += red[PCG_WARPS + w];
  }
}

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i

// inverse of a symmetric positive definite 6 x 6 block given by its lower triangle (row-major 6 x 6 float32): Cholesky,
// inverse of the factor, M = L^-T L^-1.  Returns false on a non-positive pivot.
__device__ inline bool invert_block6(const float* __restrict__ blk, double* __restrict__ M /*21*/) {
  double L[21], Li[21];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = (double)blk[6 * j + j];
#pragma unroll
    for (int m = 0; m < j; ++m) d -= L[tri(j, m)] * L[tri(j, m)];
    ok = ok && (d > 0.0) && isfinite(d);
    const double rj = rsqrt(ok ? d : 1.0);
    L[tri(j, j)] = rj;                                  // the diagonal is stored inverted
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double v = (double)blk[6 * i + j];
#pragma unroll
      for (int m = 0; m < j; ++m) v -= L[tri(i, m)] * L[tri(j, m)];
      L[tri(i, j)] = v * rj;
    }
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {                          // column j of L^-1
    Li[tri(j, j)] = L[tri(j, j)];
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double v = 0.0;
#pragma unroll
      for (int m = j; m < i; ++m) v -= L[tri(i, m)] * Li[tri(m, j)];
      Li[tri(i, j)] = v * L[tri(i, i)];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double v = 0.0;
#pragma unroll
      for (int m = i; m < 6; ++m) v += Li[tri(m, i)] * Li[tri(m, j)];
      M[tri(i, j)] = v;
    }
  return ok;
}

// the lanes of a warp over one 6 x 6 float32 block: lane = 9 * slot + c reads the c-th 16-byte piece of the block in
// `slot` (three blocks per step, lanes 27 .. 31 idle), so a step of the row part is ONE contiguous 432-byte request —
// with a lane per block the same bytes were 32 pieces 144 bytes apart (ncu: the L1 the busiest unit of the kernel).
// Piece c holds the flat elements 4c .. 4c + 3 of the row-major block: two of row ilo (columns j0, j0 + 1) and two of
// row ihi (columns j2, j2 + 1), ihi = ilo or ilo + 1.
struct LaneMap {
  bool act;
  int slot, c, ilo, ihi, j0, j2;
};
__device__ __forceinline__ LaneMap lane_map(int lane) {
  LaneMap m;
  m.act = lane < 27;
  m.slot = m.act ? lane / 9 : 0;
  m.c = m.act ? lane - 9 * m.slot : 0;
  const int e0 = 4 * m.c;
  m.ilo = e0 / 6;
  m.ihi = (e0 + 3) / 6;
  m.j0 = e0 - 6 * m.ilo;
  m.j2 = (m.j0 + 2) % 6;
  return m;
}

constexpr int PCG_BATCH = 8;              // 16-byte loads of a lane in flight

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// y_a restricted to the column blocks [b0, b1) of block row a, in every lane (fixed summation order).  Full batches
// carry no predicates (the idle lanes 27 .. 31 repeat lanes 0 .. 4 and are discarded at the end), so the PCG_BATCH
// loads of a lane are issued back to back; the ragged end of a range is one batch with clamped addresses and zeroed values.
__device__ __forceinline__ void row_times_p(const float* __restrict__ S, const double* __restrict__ ps, int a, int b0, int b1,
                                            const LaneMap& m, double* __restrict__ y /*6*/) {
  const float4* S4 = reinterpret_cast<const float4*>(S);
  double lo = 0.0, hi = 0.0, t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  // ---- stored blocks (a, b), b < a: contiguous in memory; y_a += B p_b
  {
    const int l0 = b0, nL = min(b1, a) - b0;
    const float4* rowq = S4 + ((size_t)a * (a + 1) / 2 + l0 + m.slot) * 9 + m.c;
    const double* pq = ps + 6 * (l0 + m.slot);
    int s0 = 0;
    for (; s0 + 3 * PCG_BATCH <= nL; s0 += 3 * PCG_BATCH) {
      float4 v[PCG_BATCH];
#pragma unroll
      for (int u = 0; u < PCG_BATCH; ++u) v[u] = __ldg(rowq + (long long)(s0 + 3 * u) * 9);
#pragma unroll
      for (int u = 0; u < PCG_BATCH; ++u) {
        const double* pb_ = pq + 6 * (s0 + 3 * u);
        const double2 pa = lds2(pb_ + m.j0), pb = lds2(pb_ + m.j2);
        lo += (double)v[u].x * pa.x + (double)v[u].y * pa.y;
        hi += (double)v[u].z * pb.x + (double)v[u].w * pb.y;
      }
    }
    if (s0 < nL) {
      float4 v[PCG_BATCH];
#pragma unroll
      for (int u = 0; u < PCG_BATCH; ++u) {
        const int bi = s0 + 3 * u + m.slot;
        v[u] = __ldg(rowq + (long long)(min(bi, nL - 1) - m.slot) * 9);
        if (bi >= nL) v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < PCG_BATCH; ++u) {
        const double* pb_ = pq + 6 * (min(s0 + 3 * u + m.slot, nL - 1) - m.slot);
        const double2 pa = lds2(pb_ + m.j0), pb = lds2(pb_ + m.j2);
        lo += (double)v[u].x * pa.x + (double)v[u].y * pa.y;
        hi += (double)v[u].z * pb.x + (double)v[u].w * pb.y;
      }
    }
  }
  // ---- stored blocks (b, a), b > a: one 144-byte block per slot; y_a += B^T p_b
  {
    const int l0 = max(b0, a + 1), nT = b1 - l0;
    const float4* colq = S4 + (size_t)a * 9 + m.c;
    int s0 = 0;
    for (; s0 + 3 * PCG_BATCH <= nT; s0 += 3 * PCG_BATCH) {
      float4 v[PCG_BATCH];
#pragma unroll
      for (int u = 0; u < PCG_BATCH; ++u) {
        const size_t b = (size_t)(l0 + s0 + 3 * u + m.slot);
        v[u] = __ldg(colq + (b * (b + 1) / 2) * 9);
      }
#pragma unroll
      for (int u = 0; u < PCG_BATCH; ++u) {
        const double* pb_ = ps + 6 * (l0 + s0 + 3 * u + m.slot);
        const double pl = pb_[m.ilo], ph = pb_[m.ihi];
        t0 += (double)v[u].x * pl;
        t1 += (double)v[u].y * pl;
        t2 += (double)v[u].z * ph;
        t3 += (double)v[u].w * ph;
      }
    }
    if (s0 < nT) {
      float4 v[PCG_BATCH];
#pragma unroll
      for (int u = 0; u < PCG_BATCH; ++u) {
        const int bi = s0 + 3 * u + m.slot;
        const size_t b = (size_t)(l0 + min(bi, nT - 1));
        v[u] = __ldg(colq + (b * (b + 1) / 2) * 9);
        if (bi >= nT) v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < PCG_BATCH; ++u) {
        const double* pb_ = ps + 6 * (l0 + min(s0 + 3 * u + m.slot, nT - 1));
        const double pl = pb_[m.ilo], ph = pb_[m.ihi];
        t0 += (double)v[u].x * pl;
        t1 += (double)v[u].y * pl;
        t2 += (double)v[u].z * ph;
        t3 += (double)v[u].w * ph;
      }
    }
  }
  // ---- the diagonal block through its lower triangle: element (i, j), j <= i, serves row i, and column j when j < i
  if (b0 <= a && a < b1) {
    float4 v = __ldg(S4 + ((size_t)a * (a + 1) / 2 + a) * 9 + m.c);
    if (m.slot != 0) v = make_float4(0.f, 0.f, 0.f, 0.f);
    const double* pa = ps + 6 * a;
    const double pl = pa[m.ilo], ph = pa[m.ihi];
    const double2 q0 = lds2(pa + m.j0), q2 = lds2(pa + m.j2);
    const double vx = v.x, vy = v.y, vz = v.z, vw = v.w;
    lo += (m.j0 <= m.ilo ? vx : 0.0) * q0.x + (m.j0 + 1 <= m.ilo ? vy : 0.0) * q0.y;
    hi += (m.j2 <= m.ihi ? vz : 0.0) * q2.x + (m.j2 + 1 <= m.ihi ? vw : 0.0) * q2.y;
    t0 += (m.j0 < m.ilo ? vx : 0.0) * pl;
    t1 += (m.j0 + 1 < m.ilo ? vy : 0.0) * pl;
    t2 += (m.j2 < m.ihi ? vz : 0.0) * ph;
    t3 += (m.j2 + 1 < m.ihi ? vw : 0.0) * ph;
  }
  if (!m.act) lo = hi = t0 = t1 = t2 = t3 = 0.0;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double v = (m.ilo == j ? lo : 0.0) + (m.ihi == j ? hi : 0.0) + (m.j0 == j ? t0 : 0.0) + (m.j0 + 1 == j ? t1 : 0.0) +
               (m.j2 == j ? t2 : 0.0) + (m.j2 + 1 == j ? t3 : 0.0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    y[j] = v;
  }
}

// z = M^-1 r for one block (M^-1 as its lower triangle in shared memory); returns r . z
__device__ __forceinline__ double precondition6(const double* __restrict__ M, const double* r, double* z) {
  double rz = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double v = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) v += M[j <= i ? tri(i, j) : tri(j, i)] * r[j];
    z[i] = v;
    rz += r[i] * v;
  }
  return rz;
}

__global__ void __launch_bounds__(PCG_THREADS, 1) spd_pcg_kernel(PcgPlan P) {
  extern __shared__ __align__(16) double sm[];
  const int n = P.n, C = P.C, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* xs = sm;                 // x and p, identical in every CTA (p is what the product reads)
  double* ps = xs + n;
  double* Ms = ps + n;             // 21 C: inverses of the diagonal blocks, lower triangles
  double* red = Ms + 21 * (size_t)C;      // three reduction areas
  __shared__ int s_fail;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  // A thread OWNS the block rows a = tid + 256 k (k < PCG_NB) of every vector, in every CTA alike: its residual stays
  // in registers, x and p in its own shared-memory slots, and the updates need no index arithmetic and no barrier
  // other than the one that publishes p.
  double r[PCG_NB][6], z[PCG_NB][6];
  double bb_l = 0.0, rz_l = 0.0;
  // ---- preconditioner and start: x = 0, r = b = -g, z = M^-1 r, p = z
  for (int a = tid; a < C; a += PCG_THREADS) {
    const float* blk = P.S + ((size_t)a * (a + 1) / 2 + a) * 36;
    float b36[36];
#pragma unroll
    for (int q4 = 0; q4 < 9; ++q4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(blk) + q4);
      b36[4 * q4] = q.x; b36[4 * q4 + 1] = q.y; b36[4 * q4 + 2] = q.z; b36[4 * q4 + 3] = q.w;
    }
    double M[21];
    if (!invert_block6(b36, M)) s_fail = 1;
#pragma unroll
    for (int q = 0; q < 21; ++q) Ms[21 * (size_t)a + q] = M[q];
  }
#pragma unroll
  for (int k = 0; k < PCG_NB; ++k) {
    const int a = tid + PCG_THREADS * k;
    if (a < C) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        r[k][i] = -(double)P.g[6 * a + i];
        bb_l += r[k][i] * r[k][i];
        xs[6 * a + i] = 0.0;
      }
      rz_l += precondition6(Ms + 21 * (size_t)a, r[k], z[k]);      // (this thread's own blocks: no barrier needed)
#pragma unroll
      for (int i = 0; i < 6; ++i) ps[6 * a + i] = z[k][i];
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) r[k][i] = z[k][i] = 0.0;
    }
  }
  double bb, rz;
  block_sum2(bb_l, rz_l, red + PCG_WARPS, bb, rz);           // (its barriers publish p and s_fail)
  int state = 0;                   // 0 running, 1 converged, 2 failed
  if (s_fail || !isfinite(bb) || !isfinite(rz)) state = 2;
  else if (bb == 0.0) state = 1;
  int it = 0;
  unsigned int gen = 0;            // grid barriers passed
  bool verifying = false;
  // Block rows are dealt round-robin over the CTAs (row a -> CTA a mod grid: three or four rows per SM at C = 500), and
  // a CTA cuts each of its rows into `parts` column ranges, one per warp — so the ranges of a row are summed inside the
  // CTA (shared memory, fixed order) and S p leaves as ONE vector of n doubles: every CTA reads 24 KB of it after the
  // barrier, where per-range partial sums from all over the grid were 48 KB (and 4.4 k cycles of 148 SMs on the same lines).
  const int grid = gridDim.x;
  const int rows = (int)blockIdx.x < C ? (C - 1 - (int)blockIdx.x) / grid + 1 : 0;
  const int parts = rows ? max(1, PCG_WARPS / rows) : 1;
  __shared__ double ysm[PCG_WARPS][6];
  const LaneMap lm = lane_map(lane);
  const bool tl = P.stamps && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && tid == 0;
  long long t_mv = 0, t_bar = 0, t_vec = 0, t_v1 = 0, t0 = tl ? clock64() : 0;
  while (state == 0 && (it < P.max_iter || verifying)) {
    // ---- S p: one (row, column range) per warp
    double* wout = P.w + (size_t)(gen & 1) * n;
    if (warp < rows * parts) {
      const int slot = warp / parts, pi = warp - slot * parts;
      const int a = (int)blockIdx.x + slot * grid;
      const int b0 = (int)((long long)C * pi / parts), b1 = (int)((long long)C * (pi + 1) / parts);
      double y[6];
      row_times_p(P.S, ps, a, b0, b1, lm, y);
      if (lane < 6) {
        const double v = lane == 0 ? y[0] : lane == 1 ? y[1] : lane == 2 ? y[2] : lane == 3 ? y[3] : lane == 4 ? y[4] : y[5];
        ysm[warp][lane] = v;
      }
    }
    __syncthreads();
    if (tid < 6 * rows) {
      const int slot = tid / 6, i = tid - 6 * slot;
      double v = 0.0;
      for (int pi = 0; pi < parts; ++pi) v += ysm[slot * parts + pi][i];
      wout[6 * ((int)blockIdx.x + slot * grid) + i] = v;
    }
    if (tl) { const long long t = clock64(); t_mv += t - t0; t0 = t; }
    grid_barrier(P.bar, ++gen);
    if (tl) { const long long t = clock64(); t_bar += t - t0; t0 = t; }
    // ---- the vector part, replicated: w = S p for the rows this thread owns
    double w[PCG_NB][6];
#pragma unroll
    for (int k = 0; k < PCG_NB; ++k) {
      const int a = tid + PCG_THREADS * k;
      const double2* src = reinterpret_cast<const double2*>(wout + 6 * (a < C ? a : 0));
#pragma unroll
      for (int h = 0; h < 3; ++h) {
        const double2 q = a < C ? __ldcg(src + h) : make_double2(0.0, 0.0);
        w[k][2 * h] = q.x;
        w[k][2 * h + 1] = q.y;
      }
    }
    if (tl) { const long long t = clock64(); t_v1 += t - t0; }
    if (verifying) {
      // p held x: w = S x.  The recursively updated residual has converged; the TRUE residual b - S x decides — on a
      // numerically singular system (lambda ~ 1e-7 on float32 data) the two part ways and x is not a solution.
      double tr_l = 0.0;
#pragma unroll
      for (int k = 0; k < PCG_NB; ++k) {
        const int a = tid + PCG_THREADS * k;
        if (a < C) {
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const double d = -(double)P.g[6 * a + i] - w[k][i];
            tr_l += d * d;
          }
        }
      }
      double tr, unused;
      block_sum2(tr_l, 0.0, red + PCG_WARPS, tr, unused);
      state = (isfinite(tr) && tr <= fmin(1e4 * P.tol2, 1e-8) * bb) ? 1 : 2;       // within 100 x the tolerance, and 1e-4
      break;
    }
    double pw_l = 0.0;
#pragma unroll
    for (int k = 0; k < PCG_NB; ++k) {
      const int a = tid + PCG_THREADS * k;
      if (a < C) {
#pragma unroll
        for (int i = 0; i < 6; ++i) pw_l += ps[6 * a + i] * w[k][i];
      }
    }
    const double pw = block_sum1(pw_l, red);
    if (!(pw > 0.0) || !isfinite(pw)) { state = 2; break; }
    const double alpha = rz / pw;
    double rr_l = 0.0, rz_l2 = 0.0;
#pragma unroll
    for (int k = 0; k < PCG_NB; ++k) {
      const int a = tid + PCG_THREADS * k;
      if (a < C) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          xs[6 * a + i] += alpha * ps[6 * a + i];
          r[k][i] -= alpha * w[k][i];
          rr_l += r[k][i] * r[k][i];
        }
        rz_l2 += precondition6(Ms + 21 * (size_t)a, r[k], z[k]);
      }
    }
    double rr, rz_new;
    block_sum2(rr_l, rz_l2, red + PCG_WARPS, rr, rz_new);
    ++it;
    if (!isfinite(rr) || !isfinite(rz_new)) { state = 2; break; }
    if (rr <= P.tol2 * bb) {             // one more product, with x in the place of p
      verifying = true;
#pragma unroll
      for (int k = 0; k < PCG_NB; ++k) {
        const int a = tid + PCG_THREADS * k;
        if (a < C) {
#pragma unroll
          for (int i = 0; i < 6; ++i) ps[6 * a + i] = xs[6 * a + i];
        }
      }
      __syncthreads();
      continue;
    }
    const double beta = rz_new / rz;
    rz = rz_new;
#pragma unroll
    for (int k = 0; k < PCG_NB; ++k) {
      const int a = tid + PCG_THREADS * k;
      if (a < C) {
#pragma unroll
        for (int i = 0; i < 6; ++i) ps[6 * a + i] = z[k][i] + beta * ps[6 * a + i];
      }
    }
    __syncthreads();
    if (tl) { const long long t = clock64(); t_vec += t - t0; t0 = t; }
  }
  __syncthreads();
  if (tl) {
    long long* st = P.stamps + (blockIdx.x == 0 ? 0 : 5);
    st[0] = t_mv; st[1] = t_bar; st[2] = t_vec; st[3] = it; st[4] = t_v1;
  }
  if (blockIdx.x == 0) {
    const bool ok = state == 1;
    if (ok)
      for (int i = tid; i < n; i += PCG_THREADS) P.x[i] = xs[i];
    if (tid == 0) {
      *P.status = ok ? 1 : 0;
      if (ok && P.info) *P.info = 0;
      if (P.iters) *P.iters = it;
    }
  }
}

size_t pcg_smem_bytes(int n) { return sizeof(double) * ((size_t)2 * n + (size_t)21 * (n / 6) + 3 * PCG_WARPS + 2); }

}  // namespace

size_t sfm_pcg_scratch_doubles(int n) {
  const size_t C = (size_t)n / 6;
  return 2 * C * 6 + 16;                   // S p twice (parity), barrier words, counters
}

bool sfm_spd_pcg_fits(sfm_ctx* ctx, int n) {
  const int C = n / 6;
  return n % 6 == 0 && n >= 6 && C <= PCG_THREADS * PCG_NB && (C + ctx->sm_count - 1) / ctx->sm_count <= PCG_WARPS &&
         pcg_smem_bytes(n) <= 216 * 1024;
}

// Queues the solve on the context's stream.  scratch: sfm_pcg_scratch_doubles(n) doubles.  status_dev (device int):
// 1 when x holds the solution, 0 when the caller's fallback has to run.  iters_dev: optional device int.  rel_tol:
// |b - S x| <= rel_tol |b| of the recursively updated residual ends the iteration (the true residual is then checked).
int sfm_spd_pcg(sfm_ctx* ctx, const float* S, const float* g, int n, double* scratch, double* x, int* status_dev, int* info,
                int* iters_dev, double rel_tol) {
  SFM_REQUIRE(rel_tol > 0.0 && rel_tol < 1.0, "sfm_spd_pcg: relative tolerance %g", rel_tol);
  SFM_REQUIRE(sfm_spd_pcg_fits(ctx, n), "sfm_spd_pcg: %d unknowns do not fit the shared-memory-resident vectors", n);
  PcgPlan P;
  P.n = n;
  P.C = n / 6;
  const int grid = ctx->sm_count;
  P.max_iter = 400;
  if (const char* e = getenv("SFM_PCG_MAX_ITER")) P.max_iter = std::max(1, atoi(e));      // (tests: 1 forces the fallback path)
  P.tol2 = rel_tol * rel_tol;
  P.S = S;
  P.g = g;
  P.x = x;
  P.w = scratch;
  P.bar = reinterpret_cast<unsigned int*>(scratch + 2 * (size_t)n);
  P.status = status_dev;
  P.info = info;
  P.iters = iters_dev;
  P.stamps = getenv("SFM_PCG_TIMELINE") ? reinterpret_cast<long long*>(scratch + 2 * (size_t)n + 2) : nullptr;
  SFM_CUDA(cudaMemsetAsync(P.bar, 0, 2 * sizeof(unsigned int), ctx->stream));
  SFM_CUDA(cudaMemsetAsync(status_dev, 0, sizeof(int), ctx->stream));
  if (info) SFM_CUDA(cudaMemsetAsync(info, 0, sizeof(int), ctx->stream));      // (the fallback behind this solve sets it on a bad pivot)
  const size_t smem = pcg_smem_bytes(n);
  static size_t attr_set = 0;
  if (smem > attr_set) {
    SFM_CUDA(cudaFuncSetAttribute(spd_pcg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    attr_set = 216 * 1024;
  }
  SFM_LAUNCH(ctx, SFM_K_BA_SOLVE, (spd_pcg_kernel<<<grid, PCG_THREADS, smem, ctx->stream>>>(P)));
  if (P.stamps) {
    long long h[10];
    SFM_CUDA(cudaMemcpyAsync(h, P.stamps, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(cudaStreamSynchronize(ctx->stream));
    const long long k = h[3] ? h[3] : 1;
    fprintf(stderr, "[pcg n=%d] %lld iterations: matrix-vector product %lld | grid barrier %lld | vector part %lld (reading S p %lld) cycles per iteration on CTA 0; %lld | %lld | %lld (%lld) on the last CTA\n",
            n, h[3], h[0] / k, h[1] / k, h[2] / k, h[4] / k, h[5] / k, h[6] / k, h[7] / k, h[9] / k);
  }
  return SFM_OK;
}
