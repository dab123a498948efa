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
//   matrix-vector product: unit = (block row a, one of `parts` column ranges), one unit per WARP (2C = 1000 units on 1184
//     warps at C = 500): y_a += B_ab p_b for the stored blocks b <= a and y_a += B_ba^T p_b for b > a (every stored block is
//     read twice per product, from L2), lanes over the column blocks, one shuffle reduction, six doubles per unit into
//     a parity-double-buffered array — no atomics, fixed summation order.
//   ONE grid barrier per iteration (arrive counter + generation word, acquire / release).
//   vector part: x, r, p, z live in EVERY CTA's shared memory and every CTA performs the identical updates and the
//     identical (fixed-tree) dot products on them — 12 elements per thread — so alpha, beta and the convergence verdict
//     need no second barrier and no broadcast: all CTAs hold bit-identical state and leave the loop together.
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
constexpr int PCG_KMAX = 14;              // elements per thread of a vector: n <= 256 * 14 (the shared-memory limit is ~3200)

struct PcgPlan {
  int n, C, parts, units, max_iter;
  double tol2;                  // (relative residual)^2
  const float* S;
  const float* g;
  double* x;
  double* partial;              // [2][units][6]
  unsigned int* bar;            // [0] arrivals, [1] generation
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

// every CTA of the (co-resident) grid; `gen` is thread 0's copy of the generation word
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(&bar[0], 1u);
    if (prev == gridDim.x - 1) {
      bar[0] = 0;
      st_release_u32(&bar[1], gen + 1);
    } else {
      while (ld_acquire_u32(&bar[1]) == gen) {}
    }
    gen += 1;
  }
  __syncthreads();
}

// the same value in every thread, fixed summation order
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                                   // red is free again
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < PCG_WARPS; ++w) s += red[w];
  return s;
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

__global__ void __launch_bounds__(PCG_THREADS, 1) spd_pcg_kernel(PcgPlan P) {
  extern __shared__ __align__(16) double sm[];
  const int n = P.n, C = P.C, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* xs = sm;                 // the four vectors, identical in every CTA
  double* rs = xs + n;
  double* ps = rs + n;
  double* ws = ps + n;             // S p, then z = M^-1 r
  double* Ms = ws + n;             // 21 C: inverses of the diagonal blocks, lower triangles
  double* red = Ms + 21 * (size_t)C;
  __shared__ int s_fail;
  unsigned int gen = 0;
  if (tid == 0) {
    gen = ld_acquire_u32(&P.bar[1]);
    s_fail = 0;
  }
  __syncthreads();
  // ---- preconditioner and start: x = 0, r = b = -g, z = M^-1 r, p = z
  for (int a = tid; a < C; a += PCG_THREADS) {
    const float* blk = P.S + ((size_t)a * (a + 1) / 2 + a) * 36;
    float b36[36];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(blk) + k);
      b36[4 * k] = q.x; b36[4 * k + 1] = q.y; b36[4 * k + 2] = q.z; b36[4 * k + 3] = q.w;
    }
    double M[21];
    if (!invert_block6(b36, M)) s_fail = 1;
#pragma unroll
    for (int k = 0; k < 21; ++k) Ms[21 * (size_t)a + k] = M[k];
  }
  for (int i = tid; i < n; i += PCG_THREADS) {
    xs[i] = 0.0;
    rs[i] = -(double)P.g[i];
  }
  __syncthreads();
  double bb_l = 0.0, rz_l = 0.0;
  for (int i = tid; i < n; i += PCG_THREADS) bb_l += rs[i] * rs[i];
  for (int a = tid; a < C; a += PCG_THREADS) {
    const double* M = Ms + 21 * (size_t)a;
    const double* r = rs + 6 * a;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double v = 0.0;
#pragma unroll
      for (int j = 0; j < 6; ++j) v += M[j <= i ? tri(i, j) : tri(j, i)] * r[j];
      ws[6 * a + i] = v;
      ps[6 * a + i] = v;
      rz_l += r[i] * v;
    }
  }
  const double bb = block_sum(bb_l, red);
  double rz = block_sum(rz_l, red);
  int state = 0;                   // 0 running, 1 converged, 2 failed
  if (s_fail || !isfinite(bb) || !isfinite(rz)) state = 2;
  else if (bb == 0.0) state = 1;
  int it = 0;
  bool verifying = false;
  // units are dealt round-robin over the CTAs (unit u -> CTA u mod grid): 1000 units on 148 SMs are 6 or 7 per SM, where
  // filling the CTAs one after the other gave 125 SMs eight units each and left 23 idle
  const int gw = warp * gridDim.x + blockIdx.x, total_warps = gridDim.x * PCG_WARPS;
  const bool tl = P.stamps && blockIdx.x == 0 && tid == 0;
  long long t_mv = 0, t_bar = 0, t_vec = 0, t_v1 = 0, t_v2 = 0, t_v3 = 0, t0 = tl ? clock64() : 0, t1 = 0;
  while (state == 0 && (it < P.max_iter || verifying)) {
    // ---- S p: one unit per warp
    double* part = P.partial + (size_t)(it & 1) * P.units * 6;
    for (int u = gw; u < P.units; u += total_warps) {
      const int a = u / P.parts, pi = u - a * P.parts;
      const int b0 = (int)((long long)C * pi / P.parts), b1 = (int)((long long)C * (pi + 1) / P.parts);
      double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      for (int b = b0 + lane; b < b1; b += 32) {
        const size_t blk = (b <= a) ? ((size_t)a * (a + 1) / 2 + b) : ((size_t)b * (b + 1) / 2 + a);
        const float4* q = reinterpret_cast<const float4*>(P.S + blk * 36);
        float B[36];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float4 v = __ldg(q + k);
          B[4 * k] = v.x; B[4 * k + 1] = v.y; B[4 * k + 2] = v.z; B[4 * k + 3] = v.w;
        }
        const double* pb = ps + 6 * b;
        const double p0 = pb[0], p1 = pb[1], p2 = pb[2], p3 = pb[3], p4 = pb[4], p5 = pb[5];
        if (b == a) {              // the diagonal block through its lower triangle
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const double pv[6] = {p0, p1, p2, p3, p4, p5};
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j) v += (double)(j <= i ? B[6 * i + j] : B[6 * j + i]) * pv[j];
            acc[i] += v;
          }
        } else if (b < a) {
#pragma unroll
          for (int i = 0; i < 6; ++i)
            acc[i] += (double)B[6 * i] * p0 + (double)B[6 * i + 1] * p1 + (double)B[6 * i + 2] * p2 + (double)B[6 * i + 3] * p3 +
                      (double)B[6 * i + 4] * p4 + (double)B[6 * i + 5] * p5;
        } else {
#pragma unroll
          for (int i = 0; i < 6; ++i)
            acc[i] += (double)B[i] * p0 + (double)B[6 + i] * p1 + (double)B[12 + i] * p2 + (double)B[18 + i] * p3 +
                      (double)B[24 + i] * p4 + (double)B[30 + i] * p5;
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[i] = v;
      }
      if (lane < 6) {
        const double v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : lane == 3 ? acc[3] : lane == 4 ? acc[4] : acc[5];
        part[(size_t)u * 6 + lane] = v;
      }
    }
    if (tl) { const long long t = clock64(); t_mv += t - t0; t0 = t; }
    grid_barrier(P.bar, gen);
    if (tl) { const long long t = clock64(); t_bar += t - t0; t0 = t; }
    // ---- the vector part, replicated: w = S p (from the partial sums), alpha, x, r, z, beta, p
    double pw_l = 0.0;
    {
      // a thread's elements i = tid + 256 k together: their loads of the partial sums are all in flight at once
      double v[PCG_KMAX];
#pragma unroll
      for (int k = 0; k < PCG_KMAX; ++k) v[k] = 0.0;
#pragma unroll 4
      for (int pi = 0; pi < P.parts; ++pi) {
#pragma unroll
        for (int k = 0; k < PCG_KMAX; ++k) {
          const int i = tid + PCG_THREADS * k;
          if (i < n) {
            const int a = i / 6, c = i - 6 * a;
            v[k] += __ldcg(part + ((size_t)a * P.parts + pi) * 6 + c);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < PCG_KMAX; ++k) {
        const int i = tid + PCG_THREADS * k;
        if (i < n) {
          ws[i] = v[k];
          pw_l += ps[i] * v[k];
        }
      }
    }
    if (tl) { t1 = clock64(); t_v1 += t1 - t0; }
    if (verifying) {
      // p held x: w = S x.  The recursively updated residual has converged; the TRUE residual b - S x decides — on a
      // numerically singular system (lambda ~ 1e-7 on float32 data) the two part ways and x is not a solution.
      double tr_l = 0.0;
      for (int i = tid; i < n; i += PCG_THREADS) {
        const double d = -(double)P.g[i] - ws[i];
        tr_l += d * d;
      }
      const double tr = block_sum(tr_l, red);
      state = (isfinite(tr) && tr <= 1e4 * P.tol2 * bb) ? 1 : 2;       // within 100 x the tolerance
      break;
    }
    const double pw = block_sum(pw_l, red);
    if (!(pw > 0.0) || !isfinite(pw)) { state = 2; break; }
    const double alpha = rz / pw;
    double rr_l = 0.0;
    for (int i = tid; i < n; i += PCG_THREADS) {
      xs[i] += alpha * ps[i];
      const double r = rs[i] - alpha * ws[i];
      rs[i] = r;
      rr_l += r * r;
    }
    __syncthreads();
    double rz_l2 = 0.0;
    for (int a = tid; a < C; a += PCG_THREADS) {
      const double* M = Ms + 21 * (size_t)a;
      const double* r = rs + 6 * a;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) v += M[j <= i ? tri(i, j) : tri(j, i)] * r[j];
        ws[6 * a + i] = v;
        rz_l2 += r[i] * v;
      }
    }
    if (tl) { const long long t = clock64(); t_v2 += t - t1; t1 = t; }
    const double rr = block_sum(rr_l, red);
    const double rz_new = block_sum(rz_l2, red);
    if (tl) { const long long t = clock64(); t_v3 += t - t1; t1 = t; }
    ++it;
    if (!isfinite(rr) || !isfinite(rz_new)) { state = 2; break; }
    if (rr <= P.tol2 * bb) {             // one more product, with x in the place of p
      verifying = true;
      for (int i = tid; i < n; i += PCG_THREADS) ps[i] = xs[i];
      __syncthreads();
      continue;
    }
    const double beta = rz_new / rz;
    rz = rz_new;
    for (int i = tid; i < n; i += PCG_THREADS) ps[i] = ws[i] + beta * ps[i];
    __syncthreads();
    if (tl) { const long long t = clock64(); t_vec += t - t0; t0 = t; }
  }
  __syncthreads();
  if (tl) { P.stamps[0] = t_mv; P.stamps[1] = t_bar; P.stamps[2] = t_vec; P.stamps[3] = it; P.stamps[4] = t_v1; P.stamps[5] = t_v2; P.stamps[6] = t_v3; }
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

size_t pcg_smem_bytes(int n) { return sizeof(double) * ((size_t)4 * n + (size_t)21 * (n / 6) + PCG_WARPS + 2); }

}  // namespace

size_t sfm_pcg_scratch_doubles(int n) {
  const size_t C = (size_t)n / 6;
  return 2 * (C * 64) * 6 + 16;            // partial sums for up to 64 parts per block row, barrier words, counters
}

bool sfm_spd_pcg_fits(sfm_ctx* ctx, int n) {
  (void)ctx;
  return n % 6 == 0 && n >= 6 && n <= PCG_THREADS * PCG_KMAX && pcg_smem_bytes(n) <= 216 * 1024;
}

// Queues the solve on the context's stream.  scratch: sfm_pcg_scratch_doubles(n) doubles.  status_dev (device int):
// 1 when x holds the solution, 0 when the caller's fallback has to run.  iters_dev: optional device int.
int sfm_spd_pcg(sfm_ctx* ctx, const float* S, const float* g, int n, double* scratch, double* x, int* status_dev, int* info,
                int* iters_dev) {
  SFM_REQUIRE(sfm_spd_pcg_fits(ctx, n), "sfm_spd_pcg: %d unknowns do not fit the shared-memory-resident vectors", n);
  PcgPlan P;
  P.n = n;
  P.C = n / 6;
  const int grid = ctx->sm_count;
  const int total_warps = grid * PCG_WARPS;
  P.parts = std::max(1, std::min(std::min(64, P.C), total_warps / P.C));
  if (const char* e = getenv("SFM_PCG_PARTS")) P.parts = std::max(1, std::min(std::min(64, P.C), atoi(e)));      // (tuning aid)
  P.units = P.C * P.parts;
  P.max_iter = 400;
  if (const char* e = getenv("SFM_PCG_MAX_ITER")) P.max_iter = std::max(1, atoi(e));      // (tests: 1 forces the fallback path)
  P.tol2 = 1e-8 * 1e-8;
  P.S = S;
  P.g = g;
  P.x = x;
  P.partial = scratch;
  P.bar = reinterpret_cast<unsigned int*>(scratch + 2 * (size_t)P.units * 6);
  P.status = status_dev;
  P.info = info;
  P.iters = iters_dev;
  P.stamps = getenv("SFM_PCG_TIMELINE") ? reinterpret_cast<long long*>(scratch + 2 * (size_t)P.units * 6 + 2) : nullptr;
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
    long long h[7];
    SFM_CUDA(cudaMemcpyAsync(h, P.stamps, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(cudaStreamSynchronize(ctx->stream));
    const long long k = h[3] ? h[3] : 1;
    fprintf(stderr, "[pcg n=%d] %lld iterations: matrix-vector product %lld | grid barrier %lld | vector part %lld (partial sums %lld, first dot + x, r, z %lld, two dots %lld) cycles per iteration (CTA 0)\n",
            n, h[3], h[0] / k, h[1] / k, h[2] / k, h[4] / k, h[5] / k, h[6] / k);
  }
  return SFM_OK;
}
