// wave_jacobi.cuh — OpenCV's one-sided Jacobi SVD (JacobiSVDImpl_<double>, modules/core/src/lapack.cpp) run by one warp as
// a wavefront over the row pairs; bit-identical to the sequential loop.  Included by pnp_epnp.cu (compiled with
// -fmad=false) and by tools/jacobi_bench.cu.
#pragma once
#include <float.h>
#include <math.h>

#include "hostmath.h"

namespace {

// One-sided Jacobi SVD (OpenCV JacobiSVDImpl_<double>) of n rows of length M (At, row stride M) by one warp, as a
// wavefront over the pairs.  Vt (n x n, identity on entry) may be null.
// sched: n x 8 words of scratch.  Returns the number of sweeps that rotated something.  On return the rows are
// orthogonal, NOT yet sorted / normalised.
template <int M, bool TIMED = false>
__device__ __noinline__ int wave_jacobi(double* __restrict__ At, double* __restrict__ Vt,
                                        int* __restrict__ sched, const int n, const int lane,
                                        long long* __restrict__ stamps = nullptr) {
  static_assert(M % 2 == 0, "rows are read as double2");
  constexpr int CPL = (M + 3) / 4;      // columns of a row pair each of the 4 lanes of a group rotates
  const int grp = lane >> 2, sub = lane & 3;
  const int max_iter = M > 30 ? M : 30;
  const double eps = DBL_EPSILON * 10;
  long long acc0 = 0, acc1 = 0, acc2 = 0, tprev = 0, nsteps = 0, nskip = 0;
  // the pairs of step phi (of every period): sweep sigma, i + j = phi + 1 ("first"), and the tail of sweep
  // sigma - 1, i + j = phi + 1 + n; group g of four lanes takes the g-th of them
  for (int e = lane; e < 8 * n; e += 32) {
    const int phi = e >> 3, g = e & 7;
    const int s1 = phi + 1, s2 = s1 + n;
    const int lo1 = max(0, s1 - (n - 1)), hi1 = (s1 - 1) >> 1;
    const int cnt1 = max(0, hi1 - lo1 + 1);
    const int lo2 = s2 - (n - 1), hi2 = (s2 - 1) >> 1;
    const int cnt2 = (s2 <= 2 * n - 3) ? max(0, hi2 - lo2 + 1) : 0;
    int code = -1;
    if (g < cnt1) { const int i = lo1 + g; code = i | ((s1 - i) << 8) | (1 << 16); }
    else if (g - cnt1 < cnt2) { const int i = lo2 + g - cnt1; code = i | ((s2 - i) << 8); }
    sched[e] = code;
  }
  __syncwarp();
  if (TIMED) tprev = clock64();
  bool chg_prev = true, chg_cur = false;
  int sweeps = 0;
  int code = sched[grp];
  for (int sigma = 0;; ++sigma) {
#pragma unroll 1
    for (int phi = 0; phi < n; ++phi) {
      const bool first = code >= 0 && ((code >> 16) & 1);
      const bool act = code >= 0 && (first ? sigma < max_iter : sigma >= 1);
      const int i = code & 0xff, j = (code >> 8) & 0xff;
      code = sched[8 * (phi + 1 < n ? phi + 1 : 0) + grp];      // next step's pair: the load latency hides behind this step
      int state = 0;                      // 0: the pair passes OpenCV's orthogonality test (no rotation), 1: it rotates, 2: borderline
      double p = 0.0, a = 0.0, b = 0.0;
      double mi[CPL], mj[CPL], vi[3], vj[3];
      if (act) {
        // OpenCV's running |Ai|^2 is the ordered sum of the squares formed at the row's last rotation (or of the
        // initial entries): the squares of the entries as they stand — recomputed here rather than kept in a second array
        const double2* Ai = reinterpret_cast<const double2*>(At + i * M);
        const double2* Aj = reinterpret_cast<const double2*>(At + j * M);
#pragma unroll
        for (int k = 0; k < M / 2; ++k) {
          const double2 x = Ai[k], y = Aj[k];
          p += x.x * y.x; p += x.y * y.y;
          a += x.x * x.x; a += x.y * x.y;
          b += y.x * y.x; b += y.y * y.y;
        }
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          const int k = sub + 4 * q;
          mi[q] = (k < M) ? At[i * M + k] : 0.0;
          mj[q] = (k < M) ? At[j * M + k] : 0.0;
        }
        if (Vt) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const int k = sub + 4 * q;
            vi[q] = (k < n) ? Vt[i * n + k] : 0.0;
            vj[q] = (k < n) ? Vt[j * n + k] : 0.0;
          }
        }
        // |p| <= eps sqrt(a b) decided on the squares wherever that is certain (hostmath.h cv_jacobi_skip)
        const double pp = p * p, lim = (eps * eps) * (a * b);
        state = 2;
        if (pp > 1e-280 && pp < 1e280 && lim > 1e-280 && lim < 1e280) {
          if (pp > lim * (1.0 + 1e-9)) state = 1;
          else if (pp < lim * (1.0 - 1e-9)) state = 0;
        } else if (p == 0.0 && lim >= 0.0) {
          state = 0;
        }
      }
      if (TIMED) { const long long t = clock64(); acc0 += t - tprev; tprev = t; ++nsteps; }
      if (!__any_sync(0xffffffffu, state != 0)) {          // nothing rotates in this step
        if (TIMED) ++nskip;
      } else {
        bool rot = false;
        double c = 1.0, s = 0.0;
        if (state != 0) {
          hm::cv_jacobi_cs(p * 2, a - b, c, s);
          rot = true;
          if (state == 2) rot = !(fabs(p) <= eps * sqrt(a * b));
        }
        if (TIMED) { const long long t = clock64(); acc1 += t - tprev + (long long)(c == 123.0); tprev = t; }
        __syncwarp();                      // every read of this step precedes every write
        if (rot) {
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            const int k = sub + 4 * q;
            if (k < M) {
              const double t0 = c * mi[q] + s * mj[q];
              const double t1 = -s * mi[q] + c * mj[q];
              At[i * M + k] = t0; At[j * M + k] = t1;
            }
          }
          if (Vt) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const int k = sub + 4 * q;
              if (k < n) {
                const double t0 = c * vi[q] + s * vj[q];
                const double t1 = -s * vi[q] + c * vj[q];
                Vt[i * n + k] = t0; Vt[j * n + k] = t1;
              }
            }
          }
        }
        __syncwarp();                      // ... and every write precedes the next step's reads
        const unsigned any = __ballot_sync(0xffffffffu, rot);
        const unsigned any_first = __ballot_sync(0xffffffffu, rot && first);
        chg_cur |= any_first != 0u;
        chg_prev |= (any & ~any_first) != 0u;
        if (TIMED) { const long long t = clock64(); acc2 += t - tprev; tprev = t; }
      }
      // OpenCV's loop ends after the first sweep that rotates nothing (or after max_iter sweeps).  Pairs of the
      // next sweep that already ran saw the same rows and the same sums as in that sweep, so they skipped too.
      bool done = false;
      if (n >= 4) {
        if (sigma >= 1 && phi == n - 4) {          // sweep sigma - 1 is complete
          if (chg_prev) ++sweeps;
          done = !chg_prev || sigma >= max_iter;
        }
      } else if (phi == n - 1) {                   // n < 4: no overlap, sweep sigma is complete
        if (chg_cur) ++sweeps;
        done = !chg_cur || sigma + 1 >= max_iter;
      }
      if (done) {
        if (TIMED && stamps && lane == 0) { stamps[0] = acc0; stamps[1] = acc1; stamps[2] = acc2; stamps[3] = nskip; stamps[4] = nsteps; }
        return sweeps;
      }
    }
    chg_prev = chg_cur;
    chg_cur = false;
  }
}

// Singular values (descending) and OpenCV's selection sort on them as a row permutation: W[pos] belongs to row
// ord[pos].  Distinct values have one descending order, found by ranking in parallel; equal values (never seen on
// this path) take the serial loop with OpenCV's swaps.
template <int M>
__device__ __forceinline__ void wave_sort(const double* __restrict__ At, int n, double* __restrict__ W, int* __restrict__ ord,
                                          double* __restrict__ Wtmp, int lane) {
  __syncwarp();
  double w = 0.0;
  if (lane < n) {
    double sd = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) { const double t = At[lane * M + k]; sd += t * t; }
    w = sqrt(sd);
    Wtmp[lane] = w;
  }
  __syncwarp();
  int rank = 0;
  bool tie = false;
  if (lane < n) {
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
      const double o = Wtmp[k];
      rank += (o > w) ? 1 : 0;
      tie |= (k != lane) && (o == w);
    }
  }
  if (__any_sync(0xffffffffu, tie)) {
    if (lane == 0) {
      for (int i = 0; i < n; ++i) { W[i] = Wtmp[i]; ord[i] = i; }
      for (int i = 0; i < n - 1; ++i) {
        int j = i;
        for (int k = i + 1; k < n; ++k)
          if (W[j] < W[k]) j = k;
        if (i != j) {
          const double tw = W[i]; W[i] = W[j]; W[j] = tw;
          const int to = ord[i]; ord[i] = ord[j]; ord[j] = to;
        }
      }
    }
  } else if (lane < n) {
    W[rank] = w;
    ord[rank] = lane;
  }
  __syncwarp();
}

}  // namespace
