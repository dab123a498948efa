// pnp.cu — hot path 3a: PnP-RANSAC (K4 hypothesis scoring, minimal solver, replay, refinement).
//
// Replaces cv2.solvePnPRansac(X, p, K, d, cv2.SOLVEPNP_ITERATIVE) as the reference calls it
// (sfm.py:67, test.py:319; the 5th positional binds to `rvec`, so OpenCV's defaults apply:
// 100 iterations, reprojection threshold 8 px, confidence 0.99, 5-point EPnP minimal solver,
// Levenberg-Marquardt refinement on the inliers of the winning hypothesis).
//
// OpenCV's RANSAC loop is sequential only in its *stopping rule*: the subset drawn at iteration i
// depends on nothing but N and i (RNG seeded with 2^64-1), and a hypothesis' score does not depend
// on earlier hypotheses.  So the whole loop is evaluated as a few stream-ordered kernels with no
// host round trip in between:
//   pnp_epnp_kernel      H minimal problems, one CTA each (pnp_epnp.cu: EPnP in OpenCV's exact arithmetic, the
//                        Jacobi decompositions as wavefronts over a warp), bit-identical hypotheses
//   pnp_score_kernel     K4: H x N reprojection tests, poses staged in shared memory, one point
//                        per thread held in registers, ballot/popc warp counts
//   pnp_replay_kernel    the accept / RANSACUpdateNumIters recursion over the count vector; then
//                        the winner's inlier list (stable compaction) and the LM refinement
//                        (pnp_refine_kernel, one CTA, normal equations by block reduction)
// and a single device->host copy of (rvec, tvec, inliers, info) at the end.
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <vector>

#include "common.cuh"
#include "chain_dev.cuh"
#include "epnp.h"
#include "pnp_dev.cuh"

#include "ransac.cuh"

namespace {

constexpr int PNP_HG = 4;          // hypotheses per CTA of the scoring kernel
constexpr int PNP_MAX_H = 1024;

// ------------------------------------------------------------------ K4 scoring
// poses: (H, 12) doubles = R (9, row-major) then t (3).  counts must be zero on entry.
__global__ void __launch_bounds__(256) pnp_score_kernel(const float* __restrict__ X, const float* __restrict__ px,
                                                        int n, const double* __restrict__ poses,
                                                        const unsigned char* __restrict__ valid, int H, PnpCam cam,
                                                        float thr2, int* __restrict__ counts,
                                                        unsigned char* __restrict__ masks,
                                                        const int* __restrict__ n_dev = nullptr) {
  if (n_dev) n = min(n, *n_dev);
  __shared__ double s_pose[PNP_HG][12];
  __shared__ int s_count[PNP_HG];
  __shared__ unsigned char s_valid[PNP_HG];
  const int h0 = blockIdx.y * PNP_HG;
  const int nh = min(PNP_HG, H - h0);
  if (threadIdx.x < nh * 12) s_pose[threadIdx.x / 12][threadIdx.x % 12] = poses[(size_t)h0 * 12 + threadIdx.x];
  if (threadIdx.x < PNP_HG) {
    s_count[threadIdx.x] = 0;
    s_valid[threadIdx.x] = (threadIdx.x < nh) && (!valid || valid[h0 + threadIdx.x]);
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float x = 0.f, y = 0.f, z = 0.f, ox = 0.f, oy = 0.f;
  if (i < n) {
    x = __ldg(X + 3 * (size_t)i); y = __ldg(X + 3 * (size_t)i + 1); z = __ldg(X + 3 * (size_t)i + 2);
    float2 o = __ldg(reinterpret_cast<const float2*>(px) + i);
    ox = o.x; oy = o.y;
  }
  for (int h = 0; h < nh; ++h) {
    bool in = (i < n) && s_valid[h] && is_inlier(s_pose[h], cam, x, y, z, ox, oy, thr2);
    if (masks && i < n) masks[(size_t)(h0 + h) * n + i] = in ? 1 : 0;
    unsigned m = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_count[h], __popc(m));
  }
  __syncthreads();
  if (threadIdx.x < nh && s_count[threadIdx.x]) atomicAdd(&counts[h0 + threadIdx.x], s_count[threadIdx.x]);
}

// ------------------------------------------------------------------ replay of the stopping rule
struct PnpResult {          // device-resident, copied to the host once
  double rvec[3], tvec[3];  // refined pose
  double rvec0[3], tvec0[3];
  int best_iter, iters_run, best_count, n_inliers, refine_iters, ok, pad0, pad1;
};

// rt6 (rvec | tvec per hypothesis) is given when the caller supplied the minimal solutions; the engine's own solver
// leaves only the poses (R | t) and the winner's rvec — OpenCV's Rodrigues(R) — is formed here, for the winner alone.
__device__ inline void pnp_replay(const int* __restrict__ counts, const unsigned char* __restrict__ valid, int n,
                                  int max_iters, double conf, const double* __restrict__ rt6,
                                  const double* __restrict__ poses, PnpResult* __restrict__ res) {
  int niters = max_iters, best = 0, best_it = -1, it = 0;
  while (it < niters) {
    if ((!valid || valid[it]) && counts[it] > max(best, 4)) {
      best = counts[it];
      best_it = it;
      niters = update_num_iters(conf, (double)(n - best) / n, 5, niters);
    }
    ++it;
  }
  res->best_iter = best_it;
  res->iters_run = it;
  res->best_count = best;
  res->ok = best_it >= 0 ? 1 : 0;
  res->n_inliers = 0;
  res->refine_iters = 0;
  double rv[3] = {0.0, 0.0, 0.0}, tv[3] = {0.0, 0.0, 0.0};
  if (best_it >= 0) {
    if (rt6) {
      for (int k = 0; k < 3; ++k) { rv[k] = rt6[6 * best_it + k]; tv[k] = rt6[6 * best_it + 3 + k]; }
    } else {
      double R[9];
      for (int k = 0; k < 9; ++k) R[k] = poses[12 * (size_t)best_it + k];
      hm::rotation_log(R, rv);
      for (int k = 0; k < 3; ++k) tv[k] = poses[12 * (size_t)best_it + 9 + k];
    }
  }
  for (int k = 0; k < 3; ++k) { res->rvec0[k] = rv[k]; res->tvec0[k] = tv[k]; res->rvec[k] = rv[k]; res->tvec[k] = tv[k]; }
}

// Winner's inliers, ascending (single CTA, stable compaction); the test is re-evaluated with the
// same arithmetic as the scoring kernel, so no H x N mask matrix has to exist.
__device__ inline void pnp_inliers(const float* __restrict__ X, const float* __restrict__ px,
                                   int n, const double* __restrict__ poses, PnpCam cam,
                                   float thr2, PnpResult* __restrict__ res,
                                   int* __restrict__ inliers) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  __shared__ double P[12];
  const int best = res->best_iter;
  if (best < 0) return;                       // uniform over the CTA
  if (threadIdx.x < 12) P[threadIdx.x] = poses[12 * (size_t)best + threadIdx.x];
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int start = 0; start < n; start += blockDim.x) {
    int i = start + threadIdx.x;
    bool f = false;
    if (i < n) {
      float2 o = __ldg(reinterpret_cast<const float2*>(px) + i);
      f = is_inlier(P, cam, __ldg(X + 3 * (size_t)i), __ldg(X + 3 * (size_t)i + 1), __ldg(X + 3 * (size_t)i + 2), o.x, o.y, thr2);
    }
    unsigned m = __ballot_sync(0xffffffffu, f);
    int pre = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[w] = __popc(m);
    __syncthreads();
    int off = 0;
    for (int k = 0; k < w; ++k) off += warp_tot[k];
    int base = base_s;
    if (f) inliers[base + off + pre] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int k = 0; k < nw; ++k) tot += warp_tot[k];
      base_s = base + tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) res->n_inliers = base_s;
  __syncthreads();
}

// ------------------------------------------------------------------ LM refinement (SOLVEPNP_ITERATIVE)
// cv2.solvePnP(inliers as float64, useExtrinsicGuess=True, flags=ITERATIVE): Levenberg-Marquardt
// with OpenCV's schedule — lambda = 10^-3 initially, (J^T J) with its diagonal scaled by
// (1+lambda), accept when the residual norm does not increase (lambda /= 10) else lambda *= 10
// and re-step from the same linearisation, stop after 20 accepted iterations or when the
// relative parameter change drops below FLT_EPSILON.
struct PoseJac {
  double R[9];
  double dR[3][9];   // dR/drvec_k
};

// The same, split over three threads of three different warps (K is a compile-time constant per call site, so
// nothing is indexed dynamically): each forms R in registers and its own dR/dr_K; the K = 0 caller publishes R.
// exp map for the LM linearisation: one sincos, 1 / theta by rsqrt beside the square root instead of behind it
// (the refinement is not a bit-exact path: OpenCV sums its normal equations through a GEMM)
__device__ inline void rodrigues_lm(const double* r, double th2, double* R) {
  if (th2 < DBL_EPSILON * DBL_EPSILON) {
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    return;
  }
  const double it = rsqrt(th2), theta = th2 * it;
  double s, c;
  sincos(theta, &s, &c);
  const double c1 = 1.0 - c;
  const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
  R[0] = c + c1 * (x * x);     R[1] = c1 * (x * y) - s * z; R[2] = c1 * (x * z) + s * y;
  R[3] = c1 * (x * y) + s * z; R[4] = c + c1 * (y * y);     R[5] = c1 * (y * z) - s * x;
  R[6] = c1 * (x * z) - s * y; R[7] = c1 * (y * z) + s * x; R[8] = c + c1 * (z * z);
}

template <int K>
__device__ inline void pose_jacobian_setup_k(const double* rv, PoseJac* pj) {
  double R[9];
  const double th2 = rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2];
  rodrigues_lm(rv, th2, R);
  if (K == 0)
    for (int i = 0; i < 9; ++i) pj->R[i] = R[i];
  const double e[3] = {K == 0 ? 1.0 : 0.0, K == 1 ? 1.0 : 0.0, K == 2 ? 1.0 : 0.0};
  double S[9];
  if (th2 < 1e-24) {
    S[0] = 0; S[1] = -e[2]; S[2] = e[1]; S[3] = e[2]; S[4] = 0; S[5] = -e[0]; S[6] = -e[1]; S[7] = e[0]; S[8] = 0;
    for (int i = 0; i < 9; ++i) pj->dR[K][i] = S[i];
    return;
  }
  const double inv_th2 = 1.0 / th2;
  // dR/dr_k = ( r_k [r]x + [ r x (I - R) e_k ]x ) R / |r|^2      (Gallego & Yezzi 2015)
  const double m[3] = {e[0] - R[0 + K], e[1] - R[3 + K], e[2] - R[6 + K]};   // (I - R) e_k
  const double c[3] = {rv[1] * m[2] - rv[2] * m[1], rv[2] * m[0] - rv[0] * m[2], rv[0] * m[1] - rv[1] * m[0]};
  const double a[3] = {rv[K] * rv[0] + c[0], rv[K] * rv[1] + c[1], rv[K] * rv[2] + c[2]};
  S[0] = 0; S[1] = -a[2]; S[2] = a[1]; S[3] = a[2]; S[4] = 0; S[5] = -a[0]; S[6] = -a[1]; S[7] = a[0]; S[8] = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      pj->dR[K][3 * i + j] = (S[3 * i] * R[j] + S[3 * i + 1] * R[3 + j] + S[3 * i + 2] * R[6 + j]) * inv_th2;
}

__device__ inline void pose_jacobian_setup(const double* rv, PoseJac* pj) {
  hm::rodrigues_to_matrix(rv, pj->R);
  double th2 = rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2];
  const double inv_th2 = th2 < 1e-24 ? 0.0 : 1.0 / th2;
  const double* R = pj->R;
  for (int k = 0; k < 3; ++k) {
    double e[3] = {k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0};
    double S[9];
    if (th2 < 1e-24) {
      S[0] = 0; S[1] = -e[2]; S[2] = e[1]; S[3] = e[2]; S[4] = 0; S[5] = -e[0]; S[6] = -e[1]; S[7] = e[0]; S[8] = 0;
      for (int i = 0; i < 9; ++i) pj->dR[k][i] = S[i];
      continue;
    }
    // dR/dr_k = ( r_k [r]x + [ r x (I - R) e_k ]x ) R / |r|^2      (Gallego & Yezzi 2015)
    double m[3] = {e[0] - R[0 + k], e[1] - R[3 + k], e[2] - R[6 + k]};   // (I - R) e_k
    double c[3] = {rv[1] * m[2] - rv[2] * m[1], rv[2] * m[0] - rv[0] * m[2], rv[0] * m[1] - rv[1] * m[0]};
    double a[3] = {rv[k] * rv[0] + c[0], rv[k] * rv[1] + c[1], rv[k] * rv[2] + c[2]};
    S[0] = 0; S[1] = -a[2]; S[2] = a[1]; S[3] = a[2]; S[4] = 0; S[5] = -a[0]; S[6] = -a[1]; S[7] = a[0]; S[8] = 0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        pj->dR[k][3 * i + j] = (S[3 * i] * R[j] + S[3 * i + 1] * R[3 + j] + S[3 * i + 2] * R[6 + j]) * inv_th2;
  }
}

constexpr int REFINE_THREADS = 256;
constexpr int REFINE_NACC = 28;   // 21 (upper JtJ) + 6 (Jt e) + 1 (|e|^2)
constexpr int REFINE_CACHED = 2;  // inliers per thread kept in registers across the LM passes

// keep one of (a, b) according to `up`, add the partner lane's other one: after the exchange the lane pair holds
// the pairwise sums of a (lower lane) and b (upper lane) — one shuffle for two values
__device__ __forceinline__ double fold_pair(double a, double b, bool up, int o) {
  const double send = up ? a : b, keep = up ? b : a;
  return keep + __shfl_xor_sync(0xffffffffu, send, o);
}

// Sum of the 28 accumulators over the CTA.  Inside a warp the reduction is a reduce-scatter (the values are halved
// with the lanes: 28 -> 14 -> 7 -> 4 -> 2 -> 1 per lane, 28 shuffles of a double instead of 140); lane l ends up
// with the warp's total of value vidx(l).
__device__ inline void block_reduce_acc(double* acc, double (*sh)[REFINE_NACC], double* out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
  double v14[14], v7[7], v4[4], v2[2];
#pragma unroll
  for (int k = 0; k < 14; ++k) v14[k] = fold_pair(acc[k], acc[k + 14], b4, 16);
#pragma unroll
  for (int k = 0; k < 7; ++k) v7[k] = fold_pair(v14[k], v14[k + 7], b3, 8);
#pragma unroll
  for (int k = 0; k < 3; ++k) v4[k] = fold_pair(v7[k], v7[k + 4], b2, 4);
  v4[3] = v7[3] + __shfl_xor_sync(0xffffffffu, v7[3], 4);
  v2[0] = fold_pair(v4[0], v4[2], b1, 2);
  v2[1] = fold_pair(v4[1], v4[3], b1, 2);
  const double v = fold_pair(v2[0], v2[1], b0, 1);
  const int s3 = (b0 ? 1 : 0) + (b1 ? 2 : 0);
  const int i7 = s3 < 3 ? s3 + (b2 ? 4 : 0) : 3;
  const int idx = i7 + (b3 ? 7 : 0) + (b4 ? 14 : 0);
  if (!(s3 == 3 && b2)) sh[w][idx] = v;              // value 3 of the 7 is held by both halves: the lower one writes
  __syncthreads();
  if (threadIdx.x < REFINE_NACC) {
    double s = 0.0;
    for (int ww = 0; ww < REFINE_THREADS / 32; ++ww) s += sh[ww][threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

// Solve (A with diag *= 1+lambda) x = b for the symmetric 6x6 normal matrix (upper packed row-major
// in a21): Cholesky; a non-positive pivot (rank-deficient configuration) falls back to the
// eigen-decomposition pseudo-inverse, which is what OpenCV's DECOMP_SVD solve amounts to.
__device__ __noinline__ void solve6_pinv(double* A, const double* b, double* x) {
  double w[6], V[36];
  hm::eig_sym<6>(A, w, V);
  double wmax = fabs(w[5]) > fabs(w[0]) ? fabs(w[5]) : fabs(w[0]);
  double thr = wmax * 6 * DBL_EPSILON;
  for (int i = 0; i < 6; ++i) x[i] = 0.0;
  for (int e = 0; e < 6; ++e) {
    if (fabs(w[e]) <= thr) continue;
    double d = 0.0;
    for (int i = 0; i < 6; ++i) d += V[6 * e + i] * b[i];
    d /= w[e];
    for (int i = 0; i < 6; ++i) x[i] += d * V[6 * e + i];
  }
}

// Every loop has constant bounds and is unrolled: the 6 x 6 factor lives in registers (this runs on one thread, on
// the refinement's critical path, once per LM pass).
__device__ inline void solve6_damped(const double* a21, const double* b, double lambda, double* x) {
  double A[6][6], Lc[6][6];
  {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = i; j < 6; ++j) { A[i][j] = a21[k]; A[j][i] = a21[k]; ++k; }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) A[i][i] *= 1.0 + lambda;
  bool pd = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = A[j][j];
#pragma unroll
    for (int m = 0; m < j; ++m) d -= Lc[j][m] * Lc[j][m];
    pd = pd && (d > 1e-14 * fabs(A[j][j]));
    const double rj = rsqrt(pd ? d : 1.0);           // Lc[j][j] = d * rj; its reciprocal rj serves every later division
    Lc[j][j] = rj;                                   // (the diagonal is stored inverted)
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double v = A[i][j];
#pragma unroll
      for (int m = 0; m < j; ++m) v -= Lc[i][m] * Lc[j][m];
      Lc[i][j] = v * rj;
    }
  }
  if (pd) {
    double y[6], xx[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double v = b[i];
#pragma unroll
      for (int m = 0; m < i; ++m) v -= Lc[i][m] * y[m];
      y[i] = v * Lc[i][i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
      double v = y[i];
#pragma unroll
      for (int m = i + 1; m < 6; ++m) v -= Lc[m][i] * xx[m];
      xx[i] = v * Lc[i][i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = xx[i];
    return;
  }
  double Af[36];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) Af[6 * i + j] = A[i][j];
  solve6_pinv(Af, b, x);
}

// 10^k for the integer-valued log10(lambda) of OpenCV's LM schedule
__device__ inline double pow10_int(double k) {
  int n = __double2int_rn(k);
  double r = 1.0, b = n < 0 ? 0.1 : 10.0;
  for (int i = n < 0 ? -n : n; i > 0; --i) r *= b;
  return r;
}

__device__ __forceinline__ uint32_t pnp_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void pnp_st_dsmem(double* p, uint32_t rank, double v) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}

// NCTA > 1: the CTAs of a cluster split the inliers; every pass their 28 partial sums meet through distributed
// shared memory (one cluster barrier per pass, partials double-buffered) and EVERY CTA then runs the identical
// scalar LM logic on the identical totals, so no parameter broadcast is needed.  The accumulation is bound by
// one SM's float64 throughput (measured 7.7 k of ~14 k cycles per pass at 2000 inliers); 8 SMs take that to ~1 k.
template <int NCTA>
__device__ inline void pnp_refine(const float* __restrict__ X,
                                                                    const float* __restrict__ px,
                                                                    const int* __restrict__ inliers, PnpCam cam,
                                                                    int max_iter, PnpResult* __restrict__ res,
                                                                    long long* __restrict__ dbg = nullptr,
                                                                    const int* __restrict__ local_list = nullptr,
                                                                    int local_n = 0) {
  __shared__ double sh[REFINE_THREADS / 32][REFINE_NACC];
  __shared__ double part[2][NCTA][REFINE_NACC];      // [parity][source CTA]: written by the peers
  const uint32_t crank = NCTA > 1 ? pnp_cluster_rank() : 0u;
  __shared__ double red[REFINE_NACC];
  __shared__ PoseJac pj;
  __shared__ double param[6], prev_param[6], JtJ[21], JtE[6];
  __shared__ int s_state;   // 0 = CALC_J requested, 1 = CHECK_ERR requested, 2 = DONE
  __shared__ double prev_err, lambda_lg10;
  __shared__ int iters;
  if (!res->ok) return;
  const int m = res->n_inliers;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; ++k) { param[k] = res->rvec0[k]; param[3 + k] = res->tvec0[k]; }
    s_state = 0;
    lambda_lg10 = -3.0;
    iters = 0;
    prev_err = 0.0;
  }
  __syncthreads();
  // the inliers this CTA sums over: its own slice of the list (fused tail kernel) or every NCTA-th block of the global
  // one.  The list does not change between the passes: a thread's first REFINE_CACHED points stay in registers, so a
  // pass starts computing at once instead of behind an index -> point chain of global loads.
  const int q0 = local_list ? (int)threadIdx.x : (int)crank * REFINE_THREADS + (int)threadIdx.x;
  const int q1 = local_list ? local_n : m;
  const int qs = local_list ? REFINE_THREADS : NCTA * REFINE_THREADS;
  double cX[REFINE_CACHED][3], cpx[REFINE_CACHED][2];
#pragma unroll
  for (int slot = 0; slot < REFINE_CACHED; ++slot) {
    const int q = q0 + slot * qs;
    cX[slot][0] = cX[slot][1] = cX[slot][2] = cpx[slot][0] = cpx[slot][1] = 0.0;
    if (q < q1) {
      const int i = local_list ? local_list[q] : inliers[q];
      cX[slot][0] = (double)X[3 * (size_t)i]; cX[slot][1] = (double)X[3 * (size_t)i + 1]; cX[slot][2] = (double)X[3 * (size_t)i + 2];
      cpx[slot][0] = (double)px[2 * (size_t)i]; cpx[slot][1] = (double)px[2 * (size_t)i + 1];
    }
  }
  // Every pass evaluates the residual AND the normal equations at `param`: when the candidate is
  // accepted its linearisation is already there (OpenCV's CHECK_ERR -> CALC_J pair in one pass).
  for (int guard = 0; guard < 2000; ++guard) {
    const int state = s_state;       // 0: first linearisation, 1: candidate check, 2: done
    if (state == 2) break;
    if (dbg && guard == 1 && threadIdx.x == 0) dbg[4] = clock64();
    if (threadIdx.x == 0) pose_jacobian_setup_k<0>(param, &pj);          // three warps, one dR/dr_k each
    else if (threadIdx.x == 32) pose_jacobian_setup_k<1>(param, &pj);
    else if (threadIdx.x == 64) pose_jacobian_setup_k<2>(param, &pj);
    __syncthreads();
    if (dbg && guard == 1 && threadIdx.x == 0) dbg[5] = clock64();
    double acc[REFINE_NACC];
#pragma unroll
    for (int k = 0; k < REFINE_NACC; ++k) acc[k] = 0.0;
    const double tx = param[3], ty = param[4], tz = param[5];
    for (int q = q0, slot = 0; q < q1; q += qs, ++slot) {
      double Xw[3], ox, oy;
      static_assert(REFINE_CACHED == 2, "the cached slots are selected by hand (static register indices)");
      if (slot < REFINE_CACHED) {
        const bool s0 = slot == 0;
        Xw[0] = s0 ? cX[0][0] : cX[1][0]; Xw[1] = s0 ? cX[0][1] : cX[1][1]; Xw[2] = s0 ? cX[0][2] : cX[1][2];
        ox = s0 ? cpx[0][0] : cpx[1][0]; oy = s0 ? cpx[0][1] : cpx[1][1];
      } else {
        const int i = local_list ? local_list[q] : inliers[q];
        Xw[0] = (double)X[3 * (size_t)i]; Xw[1] = (double)X[3 * (size_t)i + 1]; Xw[2] = (double)X[3 * (size_t)i + 2];
        ox = (double)px[2 * (size_t)i]; oy = (double)px[2 * (size_t)i + 1];
      }
      const double* R = pj.R;
      double x = R[0] * Xw[0] + R[1] * Xw[1] + R[2] * Xw[2] + tx;
      double y = R[3] * Xw[0] + R[4] * Xw[1] + R[5] * Xw[2] + ty;
      double z = R[6] * Xw[0] + R[7] * Xw[1] + R[8] * Xw[2] + tz;
      double iz = 1.0;
      if (fabs(z) > 1e-30 && fabs(z) < 1e30) {   // reciprocal by float seed + Newton steps (double accuracy, no division chain)
        iz = (double)(1.0f / (float)z);
        iz = iz * (2.0 - z * iz);
        iz = iz * (2.0 - z * iz);
        iz = iz * (2.0 - z * iz);
      } else if (z != 0.0) {
        iz = 1.0 / z;
      }
      double xn = x * iz, yn = y * iz;
      double ex = xn * cam.fx + cam.cx - ox, ey = yn * cam.fy + cam.cy - oy;
      acc[27] += ex * ex + ey * ey;
      double J[2][6];
      double a0 = cam.fx * iz, a2 = -cam.fx * xn * iz, b1 = cam.fy * iz, b2 = -cam.fy * yn * iz;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double* D = pj.dR[k];
        double dx = D[0] * Xw[0] + D[1] * Xw[1] + D[2] * Xw[2];
        double dy = D[3] * Xw[0] + D[4] * Xw[1] + D[5] * Xw[2];
        double dz = D[6] * Xw[0] + D[7] * Xw[1] + D[8] * Xw[2];
        J[0][k] = a0 * dx + a2 * dz;
        J[1][k] = b1 * dy + b2 * dz;
      }
      J[0][3] = a0; J[0][4] = 0.0; J[0][5] = a2;
      J[1][3] = 0.0; J[1][4] = b1; J[1][5] = b2;
      int k = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int b = a; b < 6; ++b) acc[k++] += J[0][a] * J[0][b] + J[1][a] * J[1][b];
      }
#pragma unroll
      for (int a = 0; a < 6; ++a) acc[21 + a] += J[0][a] * ex + J[1][a] * ey;
    }
    if (dbg && guard == 1 && threadIdx.x == 0) dbg[6] = clock64();
    block_reduce_acc(acc, sh, red);
    if (NCTA > 1) {
      // every CTA PUSHES its 28 partial sums into every peer's shared memory (remote stores do not wait), one cluster
      // barrier, then the totals are summed from local memory in rank order — identical in every CTA
      const int par = guard & 1;
      if (threadIdx.x < REFINE_NACC * NCTA) {
        const int k = threadIdx.x % REFINE_NACC;
        const uint32_t r = threadIdx.x / REFINE_NACC;
        pnp_st_dsmem(&part[par][crank][k], r, red[k]);
      }
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
      if (threadIdx.x < REFINE_NACC) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < NCTA; ++r) t += part[par][r][threadIdx.x];
        red[threadIdx.x] = t;
      }
      __syncthreads();
    }
    if (dbg && guard == 1 && threadIdx.x == 0) dbg[7] = clock64();
    if (dbg && threadIdx.x == 0) dbg[11] = guard + 1;
    if (threadIdx.x == 0) {
      const double err_norm = sqrt(red[27]);
      bool relinearise = (state == 0);
      if (state == 1) {
        bool retry = false;
        if (err_norm > prev_err) {
          lambda_lg10 += 1.0;
          if (lambda_lg10 <= 16.0) {               // reject: larger damping, same linearisation
            double dx[6];
            solve6_damped(JtJ, JtE, pow10_int(lambda_lg10), dx);
            for (int k = 0; k < 6; ++k) param[k] = prev_param[k] - dx[k];
            retry = true;
          }
        }
        if (!retry) {                               // accept
          lambda_lg10 = fmax(lambda_lg10 - 1.0, -16.0);
          double dn = 0.0, pn = 0.0;
          for (int k = 0; k < 6; ++k) { double d = param[k] - prev_param[k]; dn += d * d; pn += prev_param[k] * prev_param[k]; }
          iters += 1;
          if (iters >= max_iter || sqrt(dn) / (sqrt(pn) + DBL_EPSILON) < (double)FLT_EPSILON) s_state = 2;
          else relinearise = true;
        }
      }
      if (relinearise) {
        for (int k = 0; k < 21; ++k) JtJ[k] = red[k];
        for (int k = 0; k < 6; ++k) { JtE[k] = red[21 + k]; prev_param[k] = param[k]; }
        prev_err = err_norm;
        double dx[6];
        solve6_damped(JtJ, JtE, pow10_int(lambda_lg10), dx);
        for (int k = 0; k < 6; ++k) param[k] = prev_param[k] - dx[k];
        s_state = 1;
      }
    }
    if (dbg && guard == 1 && threadIdx.x == 0) dbg[10] = clock64();
    __syncthreads();
  }
  if (threadIdx.x == 0 && crank == 0) {
    for (int k = 0; k < 3; ++k) { res->rvec[k] = param[k]; res->tvec[k] = param[3 + k]; }
    res->refine_iters = iters;
  }
  if (NCTA > 1) {        // nobody leaves while a peer may still read its partial sums
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

// The RNG index stream of OpenCV's RANSAC for a row count that only the device knows.  The raw
// multiply-with-carry states do not depend on n (only how many of them an iteration consumes does, through
// the duplicate redraws), so they come from a table computed once on the host; the kernel reduces them
// modulo n in parallel and one thread replays the "five distinct indices" bookkeeping on the results.
constexpr int PNP_RAW = 2048;

// five distinct values starting at r[pos]; returns how many entries were consumed (0: ran off the table)
__device__ __forceinline__ int take_five(const int* r, int pos, int* five) {
  int got = 0, used = 0;
  while (got < 5) {
    if (pos + used >= PNP_RAW) return 0;
    const int j = r[pos + used];
    ++used;
    bool dup = false;
    for (int k = 0; k < got; ++k) dup |= (five[k] == j);
    if (!dup) five[got++] = j;
  }
  return used;
}

// Iteration `it` starts where the previous ones stopped, which depends on their redraws — rare events.  So
// every iteration is evaluated in parallel from a guessed start (5 entries each at first), the consumed counts
// are prefix-summed into new starts, and the round repeats until no start moves (iteration k is exact after at
// most k rounds; with a handful of redraws in 100 iterations that is a handful of rounds).
__global__ void __launch_bounds__(1024) pnp_subsets_kernel(const int* __restrict__ n_dev, int iters,
                                                           const unsigned int* __restrict__ raw, int* __restrict__ out) {
  __shared__ int r[PNP_RAW];
  __shared__ int cnt[128], warp_tot[4];
  __shared__ int overflow;
  const int n = *n_dev;
  if (n < 6) return;
  const int t = threadIdx.x;
  for (int k = t; k < PNP_RAW; k += blockDim.x) r[k] = (int)(raw[k] % (unsigned int)n);
  if (t < 128) cnt[t] = (t < iters) ? 5 : 0;
  if (t == 0) overflow = 0;
  __syncthreads();
  int five[5] = {0, 0, 0, 0, 0};
  for (int round = 0; round <= iters; ++round) {
    // exclusive prefix sum of cnt over the first 128 threads
    int start = 0;
    if (t < 128) {
      int v = cnt[t], x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((t & 31) >= o) x += y; }
      if ((t & 31) == 31) warp_tot[t >> 5] = x;
      start = x - v;
    }
    __syncthreads();
    bool changed = false;
    if (t < 128 && t < iters) {
      for (int w = 0; w < (t >> 5); ++w) start += warp_tot[w];
      const int used = take_five(r, start, five);
      if (used == 0) overflow = 1;
      else if (used != cnt[t]) { changed = true; }
      if (used) cnt[t] = used;       // own slot only; read again after the barrier below
    }
    if (!__syncthreads_or(changed ? 1 : 0)) break;
  }
  if (overflow) {                     // tiny n: the table is too short — one thread replays the recurrence
    if (t == 0) ransac_subsets(n, iters, out);
    return;
  }
  if (t < iters && t < 128) {
    int* o = out + 5 * t;
    o[0] = five[0]; o[1] = five[1]; o[2] = five[2]; o[3] = five[3]; o[4] = five[4];
  }
}

struct PoseOut {              // registration loop: what the next launches need, formed on the device (chain.cu)
  double* pose6;              // rvec | tvec
  int* n_inl;
  int* ok;
  double* Rt;                 // 12: [R|t]
  double* P;                  // 12: K [R|t]
  CamParams* cam;             // projectPoints operands of the reference's ReprojectionError (sfm.py:84,88)
  const double* K;            // 9 (device)
  long long* dbg;             // diagnostics: clock64 at the phase boundaries of the tail kernel
};

// registration loop: the result stays in HBM — pose, counts, [R|t], K[R|t], projectPoints operands
__device__ inline void pnp_publish(const PnpResult* __restrict__ res, const PoseOut& po) {
  {
    double p6[6];
    for (int k = 0; k < 3; ++k) { p6[k] = res->rvec[k]; p6[3 + k] = res->tvec[k]; po.pose6[k] = p6[k]; po.pose6[3 + k] = p6[3 + k]; }
    *po.n_inl = res->ok ? res->n_inliers : 0;
    *po.ok = res->ok;
    if (po.Rt) {
      // [R|t], P = K [R|t] for the next triangulations, and R' = Rodrigues(log(R)): the reference's
      // ReprojectionError passes R through cv2.Rodrigues twice.  R is orthonormal by construction here, so the
      // log map is taken directly (cv2's SVD re-orthonormalisation would change it by ~1e-16).
      double R[9], Rt[12];
      hm::rodrigues_to_matrix(p6, R);
      for (int i = 0; i < 3; ++i) { Rt[4 * i] = R[3 * i]; Rt[4 * i + 1] = R[3 * i + 1]; Rt[4 * i + 2] = R[3 * i + 2]; Rt[4 * i + 3] = p6[3 + i]; }
      for (int k = 0; k < 12; ++k) po.Rt[k] = Rt[k];
      const double* K = po.K;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) po.P[4 * i + j] = K[3 * i] * Rt[j] + K[3 * i + 1] * Rt[4 + j] + K[3 * i + 2] * Rt[8 + j];
      double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
      double sn = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
      double cs = (R[0] + R[4] + R[8] - 1.0) * 0.5;
      cs = cs > 1.0 ? 1.0 : (cs < -1.0 ? -1.0 : cs);
      double rv[3] = {p6[0], p6[1], p6[2]};
      if (sn >= 1e-5) {
        const double vth = acos(cs) / (2.0 * sn);
        rv[0] = rx * vth; rv[1] = ry * vth; rv[2] = rz * vth;
      }
      hm::rodrigues_to_matrix(rv, po.cam->R);
      po.cam->t[0] = p6[3]; po.cam->t[1] = p6[4]; po.cam->t[2] = p6[5];
      po.cam->fx = K[0]; po.cam->fy = K[4]; po.cam->cx = K[2]; po.cam->cy = K[5];
    }
  }

}

// Tail of the RANSAC in ONE single-CTA launch: replay of the stopping rule (thread 0), the winner's
// inlier list (stable compaction), the LM refinement.
__global__ void __launch_bounds__(REFINE_THREADS) pnp_finish_kernel(const float* __restrict__ X, const float* __restrict__ px,
                                                                    int n, const int* __restrict__ counts,
                                                                    const unsigned char* __restrict__ valid, int H,
                                                                    double conf, const double* __restrict__ poses,
                                                                    const double* __restrict__ rt6, PnpCam cam, float thr2,
                                                                    int refine_iters, int* __restrict__ inliers,
                                                                    PnpResult* __restrict__ res,
                                                                    const int* __restrict__ n_dev = nullptr,
                                                                    PoseOut po = PoseOut()) {
  if (n_dev) n = min(n, *n_dev);
  auto tick = [&](int k) { if (po.dbg && threadIdx.x == 0) po.dbg[k] = clock64(); };
  tick(0);
  if (threadIdx.x == 0) pnp_replay(counts, valid, n, H, conf, rt6, poses, res);
  __threadfence_block();
  __syncthreads();
  tick(1);
  pnp_inliers(X, px, n, poses, cam, thr2, res, inliers);
  tick(2);
  if (refine_iters > 0) pnp_refine<1>(X, px, inliers, cam, refine_iters, res, po.dbg);
  __syncthreads();
  tick(3);
  if (threadIdx.x == 0 && po.pose6) pnp_publish(res, po);
}

// LM refinement over a cluster of CTAs (registration loop; the tail kernel has produced the inlier list).
constexpr int REFINE_CLUSTER = 8;
__global__ void __cluster_dims__(REFINE_CLUSTER, 1, 1) __launch_bounds__(REFINE_THREADS)
    pnp_refine_cluster_kernel(const float* __restrict__ X, const float* __restrict__ px, const int* __restrict__ inliers,
                              PnpCam cam, int refine_iters, PnpResult* __restrict__ res, PoseOut po) {
  pnp_refine<REFINE_CLUSTER>(X, px, inliers, cam, refine_iters, res, nullptr);
  if (threadIdx.x == 0 && pnp_cluster_rank() == 0 && po.pose6) pnp_publish(res, po);
}

// Tail of the RANSAC for the registration loop in ONE cluster launch: replay of the stopping rule (every CTA, same
// result), the winner's inlier list — each CTA tests its contiguous slice of the points, compacts it in order and the
// slice counts meet through distributed shared memory, so the global list is ascending — and the LM refinement over
// the cluster, each CTA summing over its own slice.
constexpr int TAIL_CHUNK = 2048;
__device__ __forceinline__ int pnp_ld_dsmem_i32(const int* p, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), ra;
  int v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(ra) : "memory");
  return v;
}

__global__ void __cluster_dims__(REFINE_CLUSTER, 1, 1) __launch_bounds__(REFINE_THREADS)
    pnp_tail_cluster_kernel(const float* __restrict__ X, const float* __restrict__ px, int n, const int* __restrict__ counts,
                            const unsigned char* __restrict__ valid, int H, double conf, const double* __restrict__ poses,
                            PnpCam cam, float thr2, int refine_iters, int* __restrict__ inliers, PnpResult* __restrict__ res_out,
                            const int* __restrict__ n_dev, PoseOut po) {
  __shared__ PnpResult s_res;
  __shared__ int s_list[TAIL_CHUNK];
  __shared__ int s_cnt, s_warp_tot[REFINE_THREADS / 32];
  __shared__ double s_P[12];
  if (n_dev) n = min(n, *n_dev);
  const uint32_t crank = pnp_cluster_rank();
  auto tick = [&](int k) { if (po.dbg && threadIdx.x == 0 && crank == 0) po.dbg[k] = clock64(); };
  tick(0);
  // the replay walks the iterations one by one: counts and validity come into shared memory in one coalesced load
  // first, instead of one global-memory latency per iteration of the walk
  __shared__ int s_counts[128];
  __shared__ unsigned char s_valid[128];
  const bool staged = H <= 128;
  if (staged && (int)threadIdx.x < H) {
    s_counts[threadIdx.x] = counts[threadIdx.x];
    s_valid[threadIdx.x] = valid ? valid[threadIdx.x] : (unsigned char)1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    pnp_replay(staged ? s_counts : counts, staged ? s_valid : valid, n, H, conf, nullptr, poses, &s_res);
    s_cnt = 0;
  }
  __syncthreads();
  tick(1);
  const int best = s_res.best_iter;                     // identical in every CTA of the cluster
  if (best >= 0) {
    if (threadIdx.x < 12) s_P[threadIdx.x] = poses[12 * (size_t)best + threadIdx.x];
    __syncthreads();
    const int chunk = (n + REFINE_CLUSTER - 1) / REFINE_CLUSTER;
    const int lo = (int)crank * chunk, hi = min(n, lo + chunk);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int start = lo; start < hi; start += REFINE_THREADS) {
      const int i = start + (int)threadIdx.x;
      bool f = false;
      if (i < hi) {
        const float2 o = __ldg(reinterpret_cast<const float2*>(px) + i);
        f = is_inlier(s_P, cam, __ldg(X + 3 * (size_t)i), __ldg(X + 3 * (size_t)i + 1), __ldg(X + 3 * (size_t)i + 2), o.x, o.y, thr2);
      }
      const unsigned m = __ballot_sync(0xffffffffu, f);
      if (lane == 0) s_warp_tot[w] = __popc(m);
      __syncthreads();
      int off = s_cnt;
      for (int k = 0; k < w; ++k) off += s_warp_tot[k];
      if (f) s_list[off + __popc(m & ((1u << lane) - 1u))] = i;
      __syncthreads();
      if (threadIdx.x == 0) {
        int tot = 0;
        for (int k = 0; k < REFINE_THREADS / 32; ++k) tot += s_warp_tot[k];
        s_cnt += tot;
      }
      __syncthreads();
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    int offset = 0, total = 0;
    for (uint32_t r = 0; r < REFINE_CLUSTER; ++r) {
      const int c = pnp_ld_dsmem_i32(&s_cnt, r);
      offset += (r < crank) ? c : 0;
      total += c;
    }
    for (int k = threadIdx.x; k < s_cnt; k += REFINE_THREADS) inliers[offset + k] = s_list[k];
    if (threadIdx.x == 0) s_res.n_inliers = total;
    __syncthreads();
    tick(2);
    if (refine_iters > 0) pnp_refine<REFINE_CLUSTER>(X, px, nullptr, cam, refine_iters, &s_res, crank == 0 ? po.dbg : nullptr, s_list, s_cnt);
    __syncthreads();
    tick(3);
    if (po.dbg && threadIdx.x == 0 && crank == 0) { po.dbg[8] = s_res.refine_iters; po.dbg[9] = total; }
  }
  if (threadIdx.x == 0 && crank == 0) {
    *res_out = s_res;
    if (po.pose6) pnp_publish(&s_res, po);
  }
  // nobody leaves while a peer may still read its shared memory (the counts; pnp_refine has its own closing barrier)
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

static PnpCam make_pnp_cam(const double* K) {
  PnpCam c = {K[0], K[4], K[2], K[5]};
  return c;
}

}  // namespace

// ============================================================================ C ABI
extern "C" int sfm_pnp_score(sfm_ctx* ctx, const float* X, const float* px, int n, const double* K,
                             const double* Rt, int H, float thr, int32_t* counts, uint8_t* masks) {
  SFM_REQUIRE(ctx && X && px && K && Rt, "sfm_pnp_score: null argument");
  SFM_REQUIRE(n >= 1 && H >= 1, "sfm_pnp_score: need n >= 1 and H >= 1");
  SFM_REQUIRE(counts || masks, "sfm_pnp_score: no output requested");
  SFM_TRY(sfm_ws_begin(ctx));
  const float *dX, *dpx;
  SFM_TRY(dev_in(ctx, X, (size_t)3 * n, &dX));
  SFM_TRY(dev_in(ctx, px, (size_t)2 * n, &dpx));
  // (H,12) [R|t] rows of a 3x4 -> R9 | t3
  double* hp;
  SFM_TRY(hs_alloc_t(ctx, (size_t)12 * H, &hp));
  for (int h = 0; h < H; ++h) {
    const double* s = Rt + 12 * (size_t)h;
    double* d = hp + 12 * (size_t)h;
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[4]; d[4] = s[5]; d[5] = s[6]; d[6] = s[8]; d[7] = s[9]; d[8] = s[10];
    d[9] = s[3]; d[10] = s[7]; d[11] = s[11];
  }
  double* dposes;
  SFM_TRY(ws_alloc_t(ctx, (size_t)12 * H, &dposes));
  SFM_CUDA(cudaMemcpyAsync(dposes, hp, sizeof(double) * 12 * H, cudaMemcpyHostToDevice, ctx->stream));
  bool host_out = false;
  DevOut<int32_t> oc;
  DevOut<uint8_t> om;
  SFM_TRY(dev_out(ctx, counts, (size_t)H, &oc, &host_out));
  SFM_TRY(dev_out(ctx, masks, (size_t)H * n, &om, &host_out));
  int32_t* dcounts = oc.dev;
  if (!dcounts) SFM_TRY(ws_alloc_t(ctx, (size_t)H, &dcounts));
  SFM_CUDA(cudaMemsetAsync(dcounts, 0, sizeof(int32_t) * H, ctx->stream));
  const float thr2 = (float)((double)thr * (double)thr);
  dim3 grid(div_up(n, 256), div_up(H, PNP_HG));
  SFM_LAUNCH(ctx, SFM_K_PNP_SCORE, (pnp_score_kernel<<<grid, 256, 0, ctx->stream>>>(dX, dpx, n, dposes, nullptr, H, make_pnp_cam(K), thr2, dcounts, om.dev)));
  SFM_TRY(dev_out_finish(ctx, &oc));
  SFM_TRY(dev_out_finish(ctx, &om));
  if (host_out) SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFM_OK;
}

// ---- npoints == 4: cv2.solvePnPRansac does not iterate — model_points == npoints, so it calls
// solvePnP(flags = SOLVEPNP_P3P) once: the perspective-three-point problem on the first three correspondences, the
// fourth choosing among its (up to four) solutions by reprojection error, every point reported as an inlier, no
// refinement (calib3d solvepnp.cpp).  OpenCV's p3p.cpp is not restated operation for operation (its result is the
// root of a quartic, not a rounding-order-dependent decision): this is Grunert's formulation — the three cosine-law
// equations in the depths s1, s2 = u s1, s3 = v s1 reduced to a quartic in v — with the four roots found by one lane
// each (Durand-Kerner, then Newton on the real axis) and the pose from the two triangles' orthonormal frames.
// Agreement with cv2: R, t to ~1e-5 (cv2's own solution leaves ~1e-5 px on the three points), same root chosen.
__device__ inline void p3p_frame(const double* A, const double* B, const double* C, double* F /*3x3, columns e1 e2 e3*/) {
  double e1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
  double w[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
  double n1 = rsqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
  for (int k = 0; k < 3; ++k) e1[k] *= n1;
  double e3[3] = {e1[1] * w[2] - e1[2] * w[1], e1[2] * w[0] - e1[0] * w[2], e1[0] * w[1] - e1[1] * w[0]};
  double n3 = rsqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
  for (int k = 0; k < 3; ++k) e3[k] *= n3;
  const double e2[3] = {e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]};
  for (int k = 0; k < 3; ++k) { F[3 * k] = e1[k]; F[3 * k + 1] = e2[k]; F[3 * k + 2] = e3[k]; }
}

__global__ void __launch_bounds__(32) pnp_p3p4_kernel(const float* __restrict__ X, const float* __restrict__ px, PnpCam cam,
                                                      int32_t* __restrict__ inl, PnpResult* __restrict__ res) {
  const int lane = threadIdx.x;
  double P[4][3], f[4][3], q[4][2];
  for (int i = 0; i < 4; ++i) {
    for (int k = 0; k < 3; ++k) P[i][k] = (double)X[3 * i + k];
    q[i][0] = (double)px[2 * i]; q[i][1] = (double)px[2 * i + 1];
    const double x = (q[i][0] - cam.cx) / cam.fx, y = (q[i][1] - cam.cy) / cam.fy;
    const double nn = rsqrt(x * x + y * y + 1.0);
    f[i][0] = x * nn; f[i][1] = y * nn; f[i][2] = nn;
  }
  auto d2 = [&](int i, int j) { double s = 0; for (int k = 0; k < 3; ++k) s += (P[i][k] - P[j][k]) * (P[i][k] - P[j][k]); return s; };
  auto dt = [&](int i, int j) { return f[i][0] * f[j][0] + f[i][1] * f[j][1] + f[i][2] * f[j][2]; };
  const double a2 = d2(1, 2), b2 = d2(0, 2), c2 = d2(0, 1);
  const double ca = dt(1, 2), cb = dt(0, 2), cg = dt(0, 1);
  bool good = a2 > 0.0 && b2 > 0.0 && c2 > 0.0;
  const double ib2 = good ? 1.0 / b2 : 0.0;
  const double qq = (a2 - c2) * ib2, rr = (a2 + c2) * ib2;
  double A[5];
  A[4] = (qq - 1) * (qq - 1) - 4 * c2 * ib2 * ca * ca;
  A[3] = 4 * (qq * (1 - qq) * cb - (1 - rr) * ca * cg + 2 * c2 * ib2 * ca * ca * cb);
  A[2] = 2 * (qq * qq - 1 + 2 * qq * qq * cb * cb + 2 * (b2 - c2) * ib2 * ca * ca - 4 * rr * ca * cb * cg + 2 * (b2 - a2) * ib2 * cg * cg);
  A[1] = 4 * (-qq * (1 + qq) * cb + 2 * a2 * ib2 * cg * cg * cb - (1 - rr) * ca * cg);
  A[0] = (1 + qq) * (1 + qq) - 4 * a2 * ib2 * cg * cg;
  good = good && fabs(A[4]) > 1e-14 * (fabs(A[3]) + fabs(A[2]) + fabs(A[1]) + fabs(A[0]));
  // monic quartic v^4 + m3 v^3 + m2 v^2 + m1 v + m0: lanes 0..3 hold one root each (Durand-Kerner, simultaneous updates)
  const double i4 = good ? 1.0 / A[4] : 0.0;
  const double m3 = A[3] * i4, m2 = A[2] * i4, m1 = A[1] * i4, m0 = A[0] * i4;
  const double bound = 1.0 + fmax(fmax(fabs(m3), fabs(m2)), fmax(fabs(m1), fabs(m0)));
  const int r4 = lane & 3;
  // starting points on a circle inside the Cauchy bound, not symmetric about the real axis
  double zr = 0.5 * bound * cos(0.4 + 1.5707963267948966 * r4), zi = 0.5 * bound * sin(0.4 + 1.5707963267948966 * r4);
#pragma unroll 1
  for (int it = 0; it < 200; ++it) {
    // p(z) by Horner, complex
    double pr = 1.0, pi = 0.0;
    const double cs[4] = {m3, m2, m1, m0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double tr = pr * zr - pi * zi + cs[k], ti = pr * zi + pi * zr;
      pr = tr; pi = ti;
    }
    double dr = 1.0, di = 0.0;
#pragma unroll
    for (int o = 1; o < 4; ++o) {
      const int src = (lane & ~3) | ((r4 + o) & 3);
      const double orr = __shfl_sync(0xffffffffu, zr, src), oi = __shfl_sync(0xffffffffu, zi, src);
      const double er = zr - orr, ei = zi - oi;
      const double tr = dr * er - di * ei, ti = dr * ei + di * er;
      dr = tr; di = ti;
    }
    const double den = dr * dr + di * di;
    double sr = 0.0, si = 0.0;
    if (den > 0.0) { sr = (pr * dr + pi * di) / den; si = (pi * dr - pr * di) / den; }
    zr -= sr; zi -= si;
    const bool conv = fabs(sr) + fabs(si) <= 1e-15 * (fabs(zr) + fabs(zi));
    if (__all_sync(0xffffffffu, conv)) break;
  }
  // real roots (a double root carries ~1e-8 of imaginary noise), polished on the real axis
  double v = zr;
  bool sol = good && lane < 4 && isfinite(zr) && isfinite(zi) && fabs(zi) <= 1e-6 * fmax(1.0, fabs(zr));
  if (sol) {
    for (int it = 0; it < 3; ++it) {
      const double pv = (((v + m3) * v + m2) * v + m1) * v + m0;
      const double dv = ((4 * v + 3 * m3) * v + 2 * m2) * v + m1;
      if (dv != 0.0 && isfinite(pv / dv)) v -= pv / dv;
    }
    sol = v > 0.0;
  }
  double R[9], t[3], err = INFINITY;
  if (sol) {
    const double den = 2 * (cg - v * ca);
    const double u = ((qq - 1) * v * v - 2 * qq * cb * v + 1 + qq) / den;
    const double s1sq = b2 / (1 + v * v - 2 * v * cb);
    sol = den != 0.0 && u > 0.0 && s1sq > 0.0 && isfinite(u) && isfinite(s1sq);
    if (sol) {
      const double s1 = sqrt(s1sq), s2 = u * s1, s3 = v * s1;
      double C[3][3];
      for (int k = 0; k < 3; ++k) { C[0][k] = s1 * f[0][k]; C[1][k] = s2 * f[1][k]; C[2][k] = s3 * f[2][k]; }
      double Fw[9], Fc[9];
      p3p_frame(P[0], P[1], P[2], Fw);
      p3p_frame(C[0], C[1], C[2], Fc);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = Fc[3 * i] * Fw[3 * j] + Fc[3 * i + 1] * Fw[3 * j + 1] + Fc[3 * i + 2] * Fw[3 * j + 2];
      for (int i = 0; i < 3; ++i) t[i] = C[0][i] - (R[3 * i] * P[0][0] + R[3 * i + 1] * P[0][1] + R[3 * i + 2] * P[0][2]);
      const double xc = R[0] * P[3][0] + R[1] * P[3][1] + R[2] * P[3][2] + t[0];
      const double yc = R[3] * P[3][0] + R[4] * P[3][1] + R[5] * P[3][2] + t[1];
      const double zc = R[6] * P[3][0] + R[7] * P[3][1] + R[8] * P[3][2] + t[2];
      const double eu = xc / zc * cam.fx + cam.cx - q[3][0], ev = yc / zc * cam.fy + cam.cy - q[3][1];
      err = eu * eu + ev * ev;
      sol = isfinite(err);
      if (!sol) err = INFINITY;
    }
  }
  // the solution with the smallest error on the fourth point; ties to the lowest lane
  double best = err;
  int who = sol ? lane : 64;
  for (int o = 16; o > 0; o >>= 1) {
    const double oe = __shfl_xor_sync(0xffffffffu, best, o);
    const int ow = __shfl_xor_sync(0xffffffffu, who, o);
    if (oe < best || (oe == best && ow < who)) { best = oe; who = ow; }
  }
  if (who == 64) {
    if (lane == 0) {
      res->ok = 0; res->n_inliers = 0; res->best_iter = -1; res->iters_run = 0; res->best_count = 0; res->refine_iters = 0;
      for (int k = 0; k < 3; ++k) { res->rvec[k] = res->tvec[k] = res->rvec0[k] = res->tvec0[k] = 0.0; }
    }
    return;
  }
  if (lane == who) {
    double rv[3];
    hm::rotation_log(R, rv);
    for (int k = 0; k < 3; ++k) { res->rvec[k] = res->rvec0[k] = rv[k]; res->tvec[k] = res->tvec0[k] = t[k]; }
    res->ok = 1; res->n_inliers = 4; res->best_iter = 0; res->iters_run = 1; res->best_count = 4; res->refine_iters = 0;
    for (int k = 0; k < 4; ++k) inl[k] = k;
  }
}

static int pnp_ransac_impl(sfm_ctx* ctx, const float* X, const float* px, int n, const double* K,
                           const double* hyp_rt6, const uint8_t* hyp_valid, int max_iters, float thr,
                           double confidence, double* rvec, double* tvec, int32_t* inliers, int32_t* n_inliers,
                           int32_t* ok, sfm_pnp_info* info) {
  SFM_REQUIRE(ctx && X && px && K && rvec && tvec && ok, "sfm_pnp_ransac: null argument");
  SFM_REQUIRE(n >= 4, "sfm_pnp_ransac: needs at least 4 correspondences (cv2 asserts npoints >= 4), got %d", n);
  if (n == 4) {          // model_points == npoints: one P3P solve, all four points inliers (pnp_p3p4_kernel)
    SFM_TRY(sfm_ws_begin(ctx));
    const float *dX4, *dpx4;
    SFM_TRY(dev_in(ctx, X, (size_t)12, &dX4));
    SFM_TRY(dev_in(ctx, px, (size_t)8, &dpx4));
    PnpResult *dres4, *hres4;
    int32_t *dinl4, *hinl4;
    SFM_TRY(ws_alloc_t(ctx, 1, &dres4));
    SFM_TRY(ws_alloc_t(ctx, 4, &dinl4));
    SFM_TRY(hs_alloc_t(ctx, 1, &hres4));
    SFM_TRY(hs_alloc_t(ctx, 4, &hinl4));
    SFM_LAUNCH(ctx, SFM_K_PNP_EPNP, (pnp_p3p4_kernel<<<1, 32, 0, ctx->stream>>>(dX4, dpx4, make_pnp_cam(K), dinl4, dres4)));
    SFM_CUDA(cudaMemcpyAsync(hres4, dres4, sizeof(PnpResult), cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(cudaStreamSynchronize(ctx->stream));
    *ok = hres4->ok;
    for (int k = 0; k < 3; ++k) { rvec[k] = hres4->rvec[k]; tvec[k] = hres4->tvec[k]; }
    const int ni4 = hres4->ok ? 4 : 0;
    if (n_inliers) *n_inliers = ni4;
    if (inliers && ni4) {
      if (sfm_is_device_ptr(inliers)) SFM_CUDA(cudaMemcpyAsync(inliers, dinl4, sizeof(int32_t) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      else for (int k = 0; k < 4; ++k) inliers[k] = k;
    }
    if (info) {
      info->iters_run = hres4->iters_run; info->best_iter = hres4->best_iter; info->hyp_solved = 1; info->refine_iters = 0;
      for (int k = 0; k < 3; ++k) { info->rvec_ransac[k] = hres4->rvec0[k]; info->tvec_ransac[k] = hres4->tvec0[k]; }
    }
    (void)hinl4;
    return SFM_OK;
  }
  SFM_REQUIRE(max_iters >= 1 && max_iters <= PNP_MAX_H, "sfm_pnp_ransac: iterationsCount %d out of range", max_iters);
  SFM_TRY(sfm_ws_begin(ctx));
  const PnpCam cam = make_pnp_cam(K);
  const float thr2 = (float)((double)thr * (double)thr);
  const float *dX, *dpx;
  SFM_TRY(dev_in(ctx, X, (size_t)3 * n, &dX));
  SFM_TRY(dev_in(ctx, px, (size_t)2 * n, &dpx));
  const int H = (n == 5) ? 1 : max_iters;
  double *dposes, *drt6;
  unsigned char* dvalid;
  int32_t* dcounts;
  int32_t* dinl;
  PnpResult* dres;
  SFM_TRY(ws_alloc_t(ctx, (size_t)12 * H, &dposes));
  SFM_TRY(ws_alloc_t(ctx, (size_t)6 * H, &drt6));
  SFM_TRY(ws_alloc_t(ctx, (size_t)H, &dvalid));
  SFM_TRY(ws_alloc_t(ctx, (size_t)H, &dcounts));
  SFM_TRY(ws_alloc_t(ctx, (size_t)n, &dinl));
  SFM_TRY(ws_alloc_t(ctx, 1, &dres));
  if (hyp_rt6) {
    // externally supplied minimal solutions (rvec|tvec per iteration): convert like cv2 does when it
    // projects with a model, R = Rodrigues(rvec)
    double* hp;
    unsigned char* hv;
    SFM_TRY(hs_alloc_t(ctx, (size_t)18 * H, &hp));
    SFM_TRY(hs_alloc_t(ctx, (size_t)H, &hv));
    for (int h = 0; h < H; ++h) {
      const double* s = hyp_rt6 + 6 * (size_t)h;
      double* d = hp + 12 * (size_t)h;
      bool good = !hyp_valid || hyp_valid[h];
      for (int k = 0; k < 6; ++k) good = good && isfinite(s[k]);
      if (good) {
        hm::rodrigues_to_matrix(s, d);
        d[9] = s[3]; d[10] = s[4]; d[11] = s[5];
      } else {
        for (int k = 0; k < 12; ++k) d[k] = 0.0;
      }
      hv[h] = good ? 1 : 0;
      memcpy(hp + 12 * (size_t)H + 6 * (size_t)h, s, 6 * sizeof(double));
    }
    SFM_CUDA(cudaMemcpyAsync(dposes, hp, sizeof(double) * 12 * H, cudaMemcpyHostToDevice, ctx->stream));
    SFM_CUDA(cudaMemcpyAsync(drt6, hp + 12 * (size_t)H, sizeof(double) * 6 * H, cudaMemcpyHostToDevice, ctx->stream));
    SFM_CUDA(cudaMemcpyAsync(dvalid, hv, H, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    PnpSubsets subs;
    subs.count = (n > 5 && H <= 100) ? H : 0;
    if (subs.count) ransac_subsets(n, subs.count, subs.idx);
    long long* dbg = nullptr;
    if (getenv("SFM_PNP_TIMELINE")) {          // =2: also per-step stamps inside the 12x12 decomposition (they cost ~100 cycles a step)
      SFM_TRY(ws_alloc_t(ctx, 48, &dbg));
      long long* hflag;
      SFM_TRY(hs_alloc_t(ctx, 48, &hflag));
      memset(hflag, 0, 48 * sizeof(long long));
      hflag[47] = atoi(getenv("SFM_PNP_TIMELINE"));
      SFM_CUDA(cudaMemcpyAsync(dbg, hflag, 48 * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    }
    SFM_TRY(sfm_pnp_epnp_launch(ctx, dX, dpx, n, H, cam, subs, dposes, dvalid, dbg, nullptr, nullptr, nullptr, 0.f));
    if (dbg) {   // diagnostics: phase boundaries of hypothesis 0 in SM clocks
      long long hs[32];
      SFM_CUDA(cudaMemcpyAsync(hs, dbg, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
      SFM_CUDA(cudaStreamSynchronize(ctx->stream));
      fprintf(stderr, "[12x12 jacobi] %lld steps (%lld without a rotation): loads+sums+test %lld | rotation parameters %lld | writes+barriers %lld (cycles, lane 0)\n",
              hs[15], hs[14], hs[11], hs[12], hs[13]);
      fprintf(stderr, "[candidates N=4lin / N=2 / N=3] svd+sort %lld %lld %lld | backsubst+gauss-newton %lld %lld %lld | pose %lld %lld %lld\n",
              hs[20] - hs[4], hs[21] - hs[4], hs[22] - hs[4], hs[23] - hs[20], hs[24] - hs[21], hs[25] - hs[22], hs[26] - hs[23],
              hs[27] - hs[24], hs[28] - hs[25]);
      fprintf(stderr, "[epnp cycles] subset+control points+alphas %lld | MtM %lld | 12x12 jacobi+sort %lld (%lld sweeps) | L,rho %lld | candidates %lld | pick %lld, store %lld, rest %lld\n",
              hs[1] - hs[0], hs[2] - hs[1], hs[3] - hs[2], hs[10], hs[4] - hs[3], hs[5] - hs[4], hs[7] - hs[5], hs[8] - hs[7], hs[6] - hs[8]);
    }
  }
  SFM_CUDA(cudaMemsetAsync(dcounts, 0, sizeof(int32_t) * H, ctx->stream));
  dim3 grid(div_up(n, 256), div_up(H, PNP_HG));
  // model_points == npoints (n == 5): OpenCV keeps all five points whatever their error
  const float thr2_eff = (n == 5) ? INFINITY : thr2;
  SFM_LAUNCH(ctx, SFM_K_PNP_SCORE, (pnp_score_kernel<<<grid, 256, 0, ctx->stream>>>(dX, dpx, n, dposes, dvalid, H, cam, thr2_eff, dcounts, nullptr)));
  PnpResult* hres;
  int32_t* hinl;
  SFM_TRY(hs_alloc_t(ctx, 1, &hres));
  SFM_TRY(hs_alloc_t(ctx, (size_t)n, &hinl));
  // model_points == npoints (n == 5): OpenCV returns the EPnP pose of all five points, all inliers, no refinement
  PoseOut po0 = PoseOut();
  if (getenv("SFM_PNP_TIMELINE")) SFM_TRY(ws_alloc_t(ctx, 8, &po0.dbg));
  SFM_LAUNCH(ctx, SFM_K_PNP_REFINE, (pnp_finish_kernel<<<1, REFINE_THREADS, 0, ctx->stream>>>(
                                        dX, dpx, n, dcounts, n == 5 ? nullptr : dvalid, n == 5 ? 1 : H, confidence, dposes, hyp_rt6 ? drt6 : nullptr,
                                        cam, thr2_eff, n == 5 ? 0 : 20, dinl, dres, nullptr, po0)));
  if (po0.dbg) {
    long long hs[8];
    SFM_CUDA(cudaMemcpyAsync(hs, po0.dbg, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(cudaStreamSynchronize(ctx->stream));
    fprintf(stderr, "[pnp tail cycles] replay %lld | inliers %lld | refine %lld  (second pass: setup %lld, accumulate %lld, reduce %lld)\n",
            hs[1] - hs[0], hs[2] - hs[1], hs[3] - hs[2], hs[5] - hs[4], hs[6] - hs[5], hs[7] - hs[6]);
  }
  // ---- single copy back (the inlier list only when the caller wants it on the host)
  SFM_CUDA(cudaMemcpyAsync(hres, dres, sizeof(PnpResult), cudaMemcpyDeviceToHost, ctx->stream));
  const bool inl_host = inliers && !sfm_is_device_ptr(inliers);
  if (inl_host) SFM_CUDA(cudaMemcpyAsync(hinl, dinl, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  *ok = hres->ok;
  for (int k = 0; k < 3; ++k) { rvec[k] = hres->rvec[k]; tvec[k] = hres->tvec[k]; }
  int ni = hres->ok ? hres->n_inliers : 0;
  if (n_inliers) *n_inliers = ni;
  if (inliers && ni > 0) {
    if (!inl_host) SFM_CUDA(cudaMemcpyAsync(inliers, dinl, sizeof(int32_t) * ni, cudaMemcpyDeviceToDevice, ctx->stream));
    else memcpy(inliers, hinl, sizeof(int32_t) * ni);
  }
  if (info) {
    info->iters_run = hres->iters_run;
    info->best_iter = hres->best_iter;
    info->hyp_solved = H;
    info->refine_iters = hres->refine_iters;
    for (int k = 0; k < 3; ++k) { info->rvec_ransac[k] = hres->rvec0[k]; info->tvec_ransac[k] = hres->tvec0[k]; }
  }
  return SFM_OK;
}

static const unsigned int* pnp_raw_table(sfm_ctx* ctx) {
  // low words of the first PNP_RAW states of cv::RNG(2^64-1): one table per device, uploaded once
  static unsigned int* tab[64] = {nullptr};
  const int dev = ctx->device & 63;
  if (!tab[dev]) {
    std::vector<unsigned int> h(PNP_RAW);
    unsigned long long state = 0xFFFFFFFFFFFFFFFFull;
    for (int k = 0; k < PNP_RAW; ++k) {
      state = (state & 0xFFFFFFFFull) * 4164903690ull + (state >> 32);
      h[k] = (unsigned int)state;
    }
    unsigned int* d = nullptr;
    if (cudaMalloc(&d, sizeof(unsigned int) * PNP_RAW) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), sizeof(unsigned int) * PNP_RAW, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    tab[dev] = d;
  }
  return tab[dev];
}

int sfm_pnp_subsets_dev(sfm_ctx* ctx, const int* n_dev, int32_t* subs_dev) {
  const unsigned int* raw = pnp_raw_table(ctx);
  SFM_REQUIRE(raw, "sfm_pnp_subsets_dev: RNG table allocation failed");
  SFM_LAUNCH(ctx, SFM_K_MISC, (pnp_subsets_kernel<<<1, 1024, 0, ctx->stream>>>(n_dev, 100, raw, subs_dev)));
  return SFM_OK;
}

int sfm_pnp_ransac_dev(sfm_ctx* ctx, const float* X, const float* px, int n_cap, const int* n_dev, const double* K,
                       const double* K_dev, double* pose6_dev, int32_t* inliers_dev, int32_t* n_inl_dev, int32_t* ok_dev,
                       double* Rt_dev, double* P_dev, CamParams* cam_dev, const int32_t* subs_dev) {
  SFM_REQUIRE(n_cap >= 1, "sfm_pnp_ransac_dev: empty capacity");
  SFM_TRY(sfm_ws_begin(ctx));
  const unsigned int* raw = subs_dev ? nullptr : pnp_raw_table(ctx);
  SFM_REQUIRE(subs_dev || raw, "sfm_pnp_ransac_dev: RNG table allocation failed");
  const PnpCam cam = make_pnp_cam(K);
  const int H = 100;
  const float thr2 = 64.0f;
  double* dposes;
  unsigned char* dvalid;
  int32_t *dcounts, *dsubs;
  PnpResult* dres;
  SFM_TRY(ws_alloc_t(ctx, (size_t)12 * H, &dposes));
  SFM_TRY(ws_alloc_t(ctx, (size_t)H, &dvalid));
  SFM_TRY(ws_alloc_t(ctx, (size_t)H, &dcounts));
  SFM_TRY(ws_alloc_t(ctx, (size_t)5 * H, &dsubs));
  SFM_TRY(ws_alloc_t(ctx, 1, &dres));
  PnpSubsets subs;
  subs.count = 0;
  if (!subs_dev) SFM_LAUNCH(ctx, SFM_K_MISC, (pnp_subsets_kernel<<<1, 1024, 0, ctx->stream>>>(n_dev, H, raw, dsubs)));
  // every hypothesis CTA scores its own pose over all points when it is done (K4's arithmetic, pnp_dev.cuh): no
  // separate scoring launch on the pose-critical path
  SFM_TRY(sfm_pnp_epnp_launch(ctx, X, px, n_cap, H, cam, subs, dposes, dvalid, nullptr, n_dev, subs_dev ? subs_dev : dsubs, dcounts, thr2));
  PoseOut po;
  po.pose6 = pose6_dev; po.n_inl = n_inl_dev; po.ok = ok_dev; po.Rt = Rt_dev; po.P = P_dev; po.cam = cam_dev; po.K = K_dev;
  po.dbg = nullptr;
  if (getenv("SFM_PNP_SINGLE_CTA_REFINE")) {
    SFM_LAUNCH(ctx, SFM_K_PNP_REFINE, (pnp_finish_kernel<<<1, REFINE_THREADS, 0, ctx->stream>>>(
                                          X, px, n_cap, dcounts, dvalid, H, 0.99, dposes, nullptr, cam, thr2, 20, inliers_dev, dres,
                                          n_dev, po)));
    return SFM_OK;
  }
  if (div_up(n_cap, REFINE_CLUSTER) <= TAIL_CHUNK && !getenv("SFM_PNP_SPLIT_TAIL")) {
    if (getenv("SFM_PNP_TIMELINE")) SFM_TRY(ws_alloc_t(ctx, 16, &po.dbg));
    SFM_LAUNCH(ctx, SFM_K_PNP_REFINE, (pnp_tail_cluster_kernel<<<REFINE_CLUSTER, REFINE_THREADS, 0, ctx->stream>>>(
                                          X, px, n_cap, dcounts, dvalid, H, 0.99, dposes, cam, thr2, 20, inliers_dev, dres, n_dev, po)));
    if (po.dbg) {      // diagnostics only: this synchronises the otherwise asynchronous chain
      long long hs[16];
      SFM_CUDA(cudaMemcpyAsync(hs, po.dbg, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
      SFM_CUDA(cudaStreamSynchronize(ctx->stream));
      fprintf(stderr, "[pnp cluster tail cycles] replay %lld | inliers %lld | refine %lld (%lld LM iterations in %lld passes, %lld inliers; pass 2: setup %lld, accumulate %lld, reduce %lld, scalar LM step %lld)\n",
              hs[1] - hs[0], hs[2] - hs[1], hs[3] - hs[2], hs[8], hs[11], hs[9], hs[5] - hs[4], hs[6] - hs[5], hs[7] - hs[6], hs[10] - hs[7]);
    }
    return SFM_OK;
  }
  SFM_LAUNCH(ctx, SFM_K_PNP_REFINE, (pnp_finish_kernel<<<1, REFINE_THREADS, 0, ctx->stream>>>(
                                        X, px, n_cap, dcounts, dvalid, H, 0.99, dposes, nullptr, cam, thr2, 0, inliers_dev, dres, n_dev,
                                        PoseOut())));
  SFM_LAUNCH(ctx, SFM_K_PNP_REFINE, (pnp_refine_cluster_kernel<<<REFINE_CLUSTER, REFINE_THREADS, 0, ctx->stream>>>(
                                        X, px, inliers_dev, cam, 20, dres, po)));
  return SFM_OK;
}

extern "C" int sfm_pnp_ransac(sfm_ctx* ctx, const float* X, const float* px, int n, const double* K,
                              int max_iters, float thr, double confidence, double* rvec, double* tvec,
                              int32_t* inliers, int32_t* n_inliers, int32_t* ok, sfm_pnp_info* info) {
  return pnp_ransac_impl(ctx, X, px, n, K, nullptr, nullptr, max_iters, thr, confidence, rvec, tvec, inliers,
                         n_inliers, ok, info);
}

extern "C" int sfm_pnp_ransac_hyp(sfm_ctx* ctx, const float* X, const float* px, int n, const double* K,
                                  const double* hyp_rt6, const uint8_t* hyp_valid, int max_iters, float thr,
                                  double confidence, double* rvec, double* tvec, int32_t* inliers,
                                  int32_t* n_inliers, int32_t* ok, sfm_pnp_info* info) {
  SFM_REQUIRE(hyp_rt6, "sfm_pnp_ransac_hyp: hypotheses missing");
  SFM_REQUIRE(n > 5, "sfm_pnp_ransac_hyp: needs more than 5 correspondences");
  return pnp_ransac_impl(ctx, X, px, n, K, hyp_rt6, hyp_valid, max_iters, thr, confidence, rvec, tvec, inliers,
                         n_inliers, ok, info);
}

// ---- host utilities (parameter marshalling; usable without a GPU)
extern "C" int sfm_ransac_subsets(int n, int iters, int32_t* out) {
  SFM_REQUIRE(n >= 5 && iters >= 0 && out, "sfm_ransac_subsets: need n >= 5");
  ransac_subsets(n, iters, out);
  return SFM_OK;
}

extern "C" int sfm_epnp(const float* X, const float* px, int n, const double* K, double* R9, double* t3) {
  SFM_REQUIRE(X && px && K && R9 && t3, "sfm_epnp: null argument");
  SFM_REQUIRE(n >= 4, "sfm_epnp: needs at least 4 points");
  hm::EpnpCam ec = {K[0], K[4], K[2], K[5]};
  std::vector<double> pw((size_t)3 * n), us((size_t)2 * n), work((size_t)7 * n);
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) pw[3 * i + k] = (double)X[3 * i + k];
    hm::epnp_roundtrip_pixel(px[2 * i], px[2 * i + 1], ec, &us[2 * i]);
  }
  hm::epnp_solve(pw.data(), us.data(), n, ec, work.data(), R9, t3);
  return SFM_OK;
}
