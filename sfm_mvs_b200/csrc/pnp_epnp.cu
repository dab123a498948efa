// pnp_epnp.cu — the minimal solver of cv2.solvePnPRansac (reference call site sfm.py:67, test.py:319): EPnP on the
// five correspondences of each RANSAC iteration, all iterations at once, bit-identical to OpenCV's solver.
//
// THIS FILE IS COMPILED WITH -fmad=false.  OpenCV's solver is plain IEEE double arithmetic in a fixed order
// (SSE3 baseline build, no fused multiply-add), and every decomposition in it is small enough to run OpenCV's own
// one-sided Jacobi instead of LAPACK (epnp.h, hostmath.h).  For 5 points the 12x12 M^T M has a 2-dimensional null
// space; which basis of it the Jacobi sweeps end in is decided by rounding, and the three beta initialisations
// depend on that basis — so the hypothesis equals OpenCV's only if every operation up to there produces the same
// bits.  Device double add / mul / div / sqrt are IEEE round-to-nearest: with contraction off, the same sequence
// gives the same bits.
//
// Latency, not throughput, is what matters (100 hypotheses on 148 SMs, on the pose-critical path of the
// registration loop), and OpenCV's Jacobi is a sequential loop over the pairs (i, j).  It is NOT sequential in
// its data: rotation (i, j) touches rows i and j only, so it can run as soon as (i, j-1) and (i-1, j) are done.
// wave_jacobi runs that wavefront: pair (i, j) of sweep s executes at step s n + i + j - 1 — up to six pairs per
// step for n = 12, a sweep every n steps with consecutive sweeps overlapping — each pair on a group of four lanes
// that redundantly accumulate the three ordered sums (<Ai, Aj>, |Ai|^2, |Aj|^2 in column order, as OpenCV's loops
// do) and rotate three columns each.  66 dependent rotations per sweep become 12 dependent steps.
//   CTA = one hypothesis, 3 warps: warp 0 does the control points, alphas, M^T M and the 12x12 decomposition,
//   then each warp takes one beta initialisation (SVD least squares on a 6 x {4, 3, 5} system by the same
//   wavefront, Gauss-Newton, absolute orientation), and thread 0 picks the pose with the smallest error.
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "epnp.h"
#include "pnp_dev.cuh"
#include "ransac.cuh"
#include "wave_jacobi.cuh"

namespace {

struct CandShared {        // one beta initialisation (one warp); At first: its rows are read as double2
  double At[30];
  double Vt[25], W[5], Wtmp[5];
  double ut[30], vts[25];  // sorted, normalised
  double pcs[15];
  double R[9], t[3], err;
  double gnA[24], gnb[6], gnbeta[4];   // Gauss-Newton: the 6 x 4 system, its right-hand side, the betas
  int ord[5], pad[3];
  int sched[40];
};
static_assert(sizeof(CandShared) % 16 == 0, "CandShared rows are read as double2");

struct EpnpShared {
  double At[144];
  CandShared cand[3];
  double pw[15], us[10], alphas[20], cws[12];
  double M[120], W[12], Wtmp[12];
  double v4[48], L[60], rho[6];
  int ord[12];
  int sched[96];
};

// ---- the solver's Gauss-Newton refinement of the betas (five iterations of a 6 x 4 Householder least squares,
// hm::epnp_gauss_newton) by one warp.  Every number is the result of the same operations in the same order as in the
// serial routine; what runs side by side is what does not depend on each other: the 30 entries of the system, and,
// at Householder step k, the columns to its right AND the right-hand side (the serial code reflects b in a second
// pass; reflections 0..k-1 have been applied to it by then either way).  The scalar part of a step (pivot scan with
// its off-by-one, scaling, sigma) is computed redundantly by every lane.
__device__ __forceinline__ double gn_A_entry(const double* l, const double* b, int c) {
  switch (c) {
    case 0: return 2 * l[0] * b[0] + l[1] * b[1] + l[3] * b[2] + l[6] * b[3];
    case 1: return l[1] * b[0] + 2 * l[2] * b[1] + l[4] * b[2] + l[7] * b[3];
    case 2: return l[3] * b[0] + l[4] * b[1] + 2 * l[5] * b[2] + l[8] * b[3];
    default: return l[6] * b[0] + l[7] * b[1] + l[8] * b[2] + 2 * l[9] * b[3];
  }
}
__device__ __forceinline__ double gn_r_entry(const double* l, const double* b, double rho) {
  return rho - (l[0] * b[0] * b[0] + l[1] * b[0] * b[1] + l[2] * b[1] * b[1] + l[3] * b[0] * b[2] +
                l[4] * b[1] * b[2] + l[5] * b[2] * b[2] + l[6] * b[0] * b[3] + l[7] * b[1] * b[3] +
                l[8] * b[2] * b[3] + l[9] * b[3] * b[3]);
}

__device__ __forceinline__ void gauss_newton_warp(const double* __restrict__ L, const double* __restrict__ rho, double* gA,
                                               double* gb, double* gbeta, int lane) {
  double x[4] = {0, 0, 0, 0};                       // (every lane carries the same x)
#pragma unroll 1
  for (int it = 0; it < 5; ++it) {
    {
      const double b[4] = {gbeta[0], gbeta[1], gbeta[2], gbeta[3]};
      if (lane < 24) gA[lane] = gn_A_entry(L + 10 * (lane >> 2), b, lane & 3);
      else if (lane < 30) gb[lane - 24] = gn_r_entry(L + 10 * (lane - 24), b, rho[lane - 24]);
    }
    __syncwarp();
    double A1[4], A2[4];
    bool dead = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!dead) {
        double colk[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) colk[i] = (i >= k) ? gA[i * 4 + k] : 0.0;
        double eta = fabs(colk[k]);
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
          const double elt = fabs(colk[i - 1]);
          if (eta < elt) eta = elt;
        }
        if (eta == 0) dead = true;
        if (!dead) {
          const double inv_eta = 1. / eta;
          double sum2 = 0.0;
#pragma unroll
          for (int i = k; i < 6; ++i) {
            colk[i] *= inv_eta;
            sum2 += colk[i] * colk[i];
          }
          double sigma = sqrt(sum2);
          if (colk[k] < 0) sigma = -sigma;
          colk[k] += sigma;
          A1[k] = sigma * colk[k];
          A2[k] = -eta * sigma;
          // lane j - (k+1) reflects column j > k of A; the next lane reflects b
          const int j = k + 1 + lane;
          if (j <= 4) {
            double* col = (j < 4) ? gA + j : gb;
            const int stride = (j < 4) ? 4 : 1;
            double sum = 0;
#pragma unroll
            for (int i = k; i < 6; ++i) sum += colk[i] * col[i * stride];
            const double tau = sum / A1[k];
#pragma unroll
            for (int i = k; i < 6; ++i) col[i * stride] -= tau * colk[i];
          }
          __syncwarp();
        }
      }
    }
    if (!dead) {
      x[3] = gb[3] / A2[3];
#pragma unroll
      for (int i = 2; i >= 0; --i) {
        double sum = 0;
#pragma unroll
        for (int j = i + 1; j < 4; ++j) sum += gA[i * 4 + j] * x[j];
        x[i] = (gb[i] - sum) / A2[i];
      }
    }
    __syncwarp();
    if (lane < 4) {
      const double xv = lane == 0 ? x[0] : (lane == 1 ? x[1] : (lane == 2 ? x[2] : x[3]));
      gbeta[lane] += xv;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(96, 1) pnp_epnp_kernel(const float* __restrict__ X, const float* __restrict__ px, int n,
                                                      int H, PnpCam cam, const __grid_constant__ PnpSubsets subs,
                                                      double* __restrict__ poses,
                                                      unsigned char* __restrict__ valid, long long* __restrict__ dbg,
                                                      const int* __restrict__ n_dev, const int* __restrict__ subs_dev,
                                                      int* __restrict__ counts, float thr2) {
  __shared__ __align__(16) EpnpShared sh;
  const int h = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (h >= H) return;
  __shared__ double s_pose[12];
  __shared__ int s_ok, s_cnt;
  if (n_dev) {
    n = *n_dev;
    if (n < 6) {                             // no minimal problem to solve: every hypothesis invalid -> ok = 0 downstream
      if (tid == 0) { valid[h] = 0; if (counts) counts[h] = 0; }
      return;
    }
  }
  auto tick = [&](int k) { if (dbg && h == 0 && tid == 0) dbg[k] = clock64(); };
  tick(0);
  const hm::EpnpCam ec = {cam.fx, cam.fy, cam.cx, cam.cy};
  if (warp == 0) {
    if (lane == 0) {
      int sub[5] = {0, 1, 2, 3, 4};
      if (subs_dev) {
        for (int k = 0; k < 5; ++k) sub[k] = subs_dev[5 * h + k];
      } else if (n > 5 && h < subs.count) {
        for (int k = 0; k < 5; ++k) sub[k] = subs.idx[5 * h + k];
      } else if (n > 5) {                            // the subset iteration h of OpenCV's RANSAC draws
        unsigned long long state = 0xFFFFFFFFFFFFFFFFull;
        for (int it = 0; it <= h; ++it)
          for (int i = 0; i < 5; ++i)
            for (;;) {
              state = (state & 0xFFFFFFFFull) * 4164903690ull + (state >> 32);
              int j = (int)((unsigned int)state % (unsigned int)n);
              bool dup = false;
              for (int k = 0; k < i; ++k) dup |= (sub[k] == j);
              if (!dup) { sub[i] = j; break; }
            }
      }
      for (int k = 0; k < 5; ++k) {
        const int j = sub[k];
        sh.pw[3 * k] = (double)X[3 * (size_t)j]; sh.pw[3 * k + 1] = (double)X[3 * (size_t)j + 1]; sh.pw[3 * k + 2] = (double)X[3 * (size_t)j + 2];
        hm::epnp_roundtrip_pixel(px[2 * (size_t)j], px[2 * (size_t)j + 1], ec, sh.us + 2 * k);
      }
      // control points and barycentric coordinates: two 3 x 3 decompositions, inherently serial (epnp.h)
      hm::epnp_control_alphas(sh.pw, 5, sh.alphas, reinterpret_cast<double(*)[3]>(sh.cws));
    }
    __syncwarp();
    tick(1);
    // M (10 x 12), then the upper triangle of M^T M: every entry its own ordered sum over the rows of M
    for (int e = lane; e < 120; e += 32) sh.M[e] = hm::epnp_M_entry(sh.alphas, sh.us, ec, e / 12, e % 12);
    __syncwarp();
    for (int e = lane; e < 78; e += 32) {
      int r = 0, rem = e;
      while (rem >= 12 - r) { rem -= 12 - r; ++r; }
      const int c = r + rem;
      double s0 = 0.0;
#pragma unroll
      for (int k = 0; k < 10; ++k) s0 += sh.M[12 * k + r] * sh.M[12 * k + c];
      sh.At[12 * r + c] = s0; sh.At[12 * c + r] = s0;
    }
    __syncwarp();
    tick(2);
    const int sweeps = (dbg && h == 0 && dbg[47] == 2) ? wave_jacobi<12, true>(sh.At, nullptr, sh.sched, 12, lane, dbg + 11)
                                       : wave_jacobi<12>(sh.At, nullptr, sh.sched, 12, lane);
    wave_sort<12>(sh.At, 12, sh.W, sh.ord, sh.Wtmp, lane);
    tick(3);
    if (dbg && h == 0 && lane == 0) dbg[10] = sweeps;
    // v[q] = left singular vector of the (q+1)-th smallest singular value: sorted row 11 - q, scaled by 1 / sigma
    for (int e = lane; e < 48; e += 32) {
      const int q = e / 12, k = e - 12 * q;
      const double sd = sh.W[11 - q];
      const double sc = sd > DBL_MIN ? 1 / sd : 0.;
      sh.v4[e] = sh.At[12 * sh.ord[11 - q] + k] * sc;
    }
    __syncwarp();
    const double* v[4] = {sh.v4, sh.v4 + 12, sh.v4 + 24, sh.v4 + 36};
    for (int e = lane; e < 66; e += 32) {
      if (e < 60) sh.L[e] = hm::epnp_L_entry(v, e / 10, e % 10);
      else sh.rho[e - 60] = hm::epnp_rho_entry(reinterpret_cast<const double(*)[3]>(sh.cws), e - 60);
    }
  }
  // lane 0 of warp 0 comes out of long serial sections: the warps (and the lanes of a warp) reach this point far
  // apart, so the barrier is the non-aligned form, which does not require a warp to arrive converged
  __syncwarp();
  asm volatile("barrier.sync 0;" ::: "memory");
  tick(4);
  {
    // warp = beta initialisation: least squares on nc columns of L by SVD (cv::solve DECOMP_SVD)
    CandShared& cs = sh.cand[warp];
    const int ap = warp, nc = hm::epnp_approx_cols(ap);
    for (int e = lane; e < 6 * nc; e += 32) {
      const int c = e / 6, i = e - 6 * c;
      const double x = sh.L[10 * i + hm::epnp_approx_col(ap, c)];
      cs.At[e] = x;
    }
    for (int e = lane; e < nc * nc; e += 32) cs.Vt[e] = (e / nc == e % nc) ? 1.0 : 0.0;
    __syncwarp();
    wave_jacobi<6>(cs.At, cs.Vt, cs.sched, nc, lane);
    wave_sort<6>(cs.At, nc, cs.W, cs.ord, cs.Wtmp, lane);
    for (int e = lane; e < 6 * nc; e += 32) {
      const int r = e / 6, k = e - 6 * r;
      const double sd = cs.W[r];
      const double sc = sd > DBL_MIN ? 1 / sd : 0.;
      cs.ut[e] = cs.At[6 * cs.ord[r] + k] * sc;
    }
    for (int e = lane; e < nc * nc; e += 32) cs.vts[e] = cs.Vt[nc * cs.ord[e / nc] + e % nc];
    __syncwarp();
    if (dbg && h == 0 && lane == 0) dbg[20 + warp] = clock64();
    if (lane == 0) {
      double b[5], betas[4] = {0, 0, 0, 0};
      if (ap == 0) hm::cv_svd_backsubst<6, 4>(cs.W, cs.ut, cs.vts, sh.rho, b);
      else if (ap == 1) hm::cv_svd_backsubst<6, 3>(cs.W, cs.ut, cs.vts, sh.rho, b);
      else hm::cv_svd_backsubst<6, 5>(cs.W, cs.ut, cs.vts, sh.rho, b);
      hm::epnp_betas_from_ls(ap, b, betas);
      for (int k = 0; k < 4; ++k) cs.gnbeta[k] = betas[k];
    }
    __syncwarp();
    gauss_newton_warp(sh.L, sh.rho, cs.gnA, cs.gnb, cs.gnbeta, lane);
    if (dbg && h == 0 && lane == 0) dbg[23 + warp] = clock64();
    if (lane == 0) {
      const double betas[4] = {cs.gnbeta[0], cs.gnbeta[1], cs.gnbeta[2], cs.gnbeta[3]};
      const double* v[4] = {sh.v4, sh.v4 + 12, sh.v4 + 24, sh.v4 + 36};
      cs.err = hm::epnp_pose_from_betas(v, betas, sh.alphas, sh.pw, sh.us, 5, ec, cs.pcs, cs.R, cs.t);
      if (dbg && h == 0) dbg[26 + warp] = clock64();
    }
  }
  __syncwarp();
  asm volatile("barrier.sync 0;" ::: "memory");
  tick(5);
  if (tid == 0) {
    const double errs[3] = {sh.cand[0].err, sh.cand[1].err, sh.cand[2].err};
    const int N = hm::epnp_pick(errs);
    if (dbg && h == 0) dbg[7] = clock64();
    const double* R = sh.cand[N].R;
    const double* t = sh.cand[N].t;
    // OpenCV hands the model on as (rvec, tvec) = (Rodrigues(R), t) and scores with R' = Rodrigues(rvec): the round trip
    // of an R = U V^T that is orthonormal to rounding moves it by ~1e-16, the same order as the last-place differences
    // between this device's sin / cos / acos and the host libm's — and three transcendental calls executed once, cold,
    // cost ~6 us of this kernel's critical path.  The scoring kernel projects with R itself; the winner's rvec (LM
    // start, info) is formed by the replay, for the winner alone.
    double* P = poses + 12 * (size_t)h;
    bool ok = true;
    for (int k = 0; k < 9; ++k) { P[k] = R[k]; ok &= isfinite(R[k]); }
    for (int k = 0; k < 3; ++k) { P[9 + k] = t[k]; ok &= isfinite(t[k]); }
    valid[h] = ok ? 1 : 0;
    if (dbg && h == 0) dbg[8] = clock64();
    if (counts) {
      for (int k = 0; k < 12; ++k) s_pose[k] = P[k];
      s_ok = ok ? 1 : 0;
      s_cnt = 0;
    }
    if (dbg && h == 0) {                 // diagnostics: the raw solver output of hypothesis 0
      double* d = reinterpret_cast<double*>(dbg + 32);
      for (int k = 0; k < 9; ++k) d[k] = R[k];
      for (int k = 0; k < 3; ++k) d[9 + k] = t[k];
    }
  }
  tick(6);
  if (counts) {
    // K4 inside the solver: this hypothesis's consensus over all points, by the CTA that produced it
    __syncwarp();
    asm volatile("barrier.sync 0;" ::: "memory");
    int c = 0;
    if (s_ok) {
#pragma unroll 4
      for (int i = tid; i < n; i += 96) {        // (unrolled: four points' division chains in flight per thread)
        const float2 o = __ldg(reinterpret_cast<const float2*>(px) + i);
        c += is_inlier(s_pose, cam, __ldg(X + 3 * (size_t)i), __ldg(X + 3 * (size_t)i + 1), __ldg(X + 3 * (size_t)i + 2), o.x, o.y, thr2) ? 1 : 0;
      }
    }
    __syncwarp();
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0 && c) atomicAdd(&s_cnt, c);
    __syncwarp();
    asm volatile("barrier.sync 0;" ::: "memory");
    if (tid == 0) counts[h] = s_cnt;
  }
}

}  // namespace

int sfm_pnp_epnp_launch(sfm_ctx* ctx, const float* X, const float* px, int n, int H, const PnpCam& cam, const PnpSubsets& subs,
                        double* poses, unsigned char* valid, long long* dbg, const int* n_dev,
                        const int* subs_dev, int* counts, float thr2) {
  SFM_LAUNCH(ctx, SFM_K_PNP_EPNP, (pnp_epnp_kernel<<<H, 96, 0, ctx->stream>>>(X, px, n, H, cam, subs, poses, valid, dbg,
                                                                              n_dev, subs_dev, counts, thr2)));
  return SFM_OK;
}

// The raw solver output (R row-major, t) of the batched kernel for explicit 5-point subsets — the hook the parity
// tests use to compare the DEVICE arithmetic with cv2.solvePnP(EPNP) bit for bit.  X (n,3), px (n,2) host or device;
// subsets (H,5) host; R9t3 (H,12) host.
extern "C" int sfm_epnp_batch(sfm_ctx* ctx, const float* X, const float* px, int n, const double* K, const int32_t* subsets,
                              int H, double* R9t3) {
  SFM_REQUIRE(ctx && X && px && K && subsets && R9t3, "sfm_epnp_batch: null argument");
  SFM_REQUIRE(n >= 5 && H >= 1 && H <= 100, "sfm_epnp_batch: need n >= 5 and 1 <= H <= 100");
  SFM_TRY(sfm_ws_begin(ctx));
  const float *dX, *dpx;
  SFM_TRY(dev_in(ctx, X, (size_t)3 * n, &dX));
  SFM_TRY(dev_in(ctx, px, (size_t)2 * n, &dpx));
  PnpSubsets subs;
  subs.count = H;
  for (int k = 0; k < 5 * H; ++k) {
    SFM_REQUIRE(subsets[k] >= 0 && subsets[k] < n, "sfm_epnp_batch: subset index out of range");
    subs.idx[k] = subsets[k];
  }
  const PnpCam cam = {K[0], K[4], K[2], K[5]};
  double *dposes, *hout;
  unsigned char* dvalid;
  long long* dbg;
  SFM_TRY(ws_alloc_t(ctx, (size_t)12 * H, &dposes));
  SFM_TRY(ws_alloc_t(ctx, (size_t)H, &dvalid));
  SFM_TRY(hs_alloc_t(ctx, (size_t)12, &hout));
  // one launch per hypothesis slot 0 with the debug block carrying the raw (R, t): H is small in the tests
  for (int h = 0; h < H; ++h) {
    PnpSubsets one;
    one.count = 1;
    for (int k = 0; k < 5; ++k) one.idx[k] = subs.idx[5 * h + k];
    SFM_TRY(ws_alloc_t(ctx, 48, &dbg));
    SFM_CUDA(cudaMemsetAsync(dbg, 0, 48 * sizeof(long long), ctx->stream));
    SFM_TRY(sfm_pnp_epnp_launch(ctx, dX, dpx, n, 1, cam, one, dposes, dvalid, dbg, nullptr, nullptr, nullptr, 0.f));
    SFM_CUDA(cudaMemcpyAsync(hout, dbg + 32, sizeof(double) * 12, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(R9t3 + 12 * (size_t)h, hout, sizeof(double) * 12);
  }
  return SFM_OK;
}
