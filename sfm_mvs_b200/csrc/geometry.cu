// geometry.cu — K2 DLT triangulation, K3 reprojection error, common_points association.
//
// Reference call sites (FlagArihant2000/sfm-mvs):
//   Triangulation      sfm.py:45-56   (cv2.triangulatePoints sfm.py:53, w-normalise sfm.py:54)
//   ReprojectionError  sfm.py:79-100  (Rodrigues :84, convertPointsFromHomogeneous :86,
//                                      projectPoints :88, cv2.norm/N :91-95)
//   common_points      sfm.py:215-239
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "chain_dev.cuh"
#include "hostmath.h"

// ============================================================================ K2 triangulate
struct ProjPair {
  double P1[12];
  double P2[12];
};

// One-sided (Hestenes) Jacobi SVD of the 4x4 DLT matrix in float64, one point per thread, all
// state in registers.  The rotation schedule (cyclic i<j, rotate while |p| > eps*sqrt(a*b), at
// most 30 sweeps) and the rotation formula are those of the Jacobi SVD OpenCV runs for a 4x4
// double matrix, so the returned singular vector normally carries the same sign as cv2's.
__device__ __forceinline__ void dlt_null_vector(double (&At)[4][4], double (&out)[4]) {
  double Vt[4][4];
  double W[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double sd = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sd += At[i][k] * At[i][k];
      Vt[i][k] = (i == k) ? 1.0 : 0.0;
    }
    W[i] = sd;
  }
  const double eps = DBL_EPSILON * 10.0;
  for (int iter = 0; iter < 30; ++iter) {
    bool changed = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = i + 1; j < 4; ++j) {
        double a = W[i], b = W[j], p = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) p += At[i][k] * At[j][k];
        // This kernel re-triangulates on the pose-critical path of the registration loop, one thread per point: its
        // time is (rotations) x (dependent latency of one rotation).  OpenCV's test is decided on the squares where
        // that is certain (hostmath.h), and OpenCV's (c, s) — a hypot, two quotients and two square roots, ~550
        // cycles of dependent float64 latency — is formed from two reciprocal square roots instead:
        //   gamma = hypot(p, beta);  q = (gamma + |beta|) / (2 gamma) = (1 + |beta| / gamma) / 2;
        //   r1 = sqrt(q);  r2 = p / (2 gamma r1)   ->   ig = rsqrt(p^2 + beta^2), q = (1 + |beta| ig) / 2,
        //   iq = rsqrt(q), r1 = q iq, r2 = p ig iq / 2          (~180 cycles; same rotation to a few ulp)
        if (hm::cv_jacobi_skip(p, a, b, eps)) continue;
        p *= 2.0;
        const double beta = a - b;
        const double ig = rsqrt(p * p + beta * beta);
        const double q = 0.5 + 0.5 * fabs(beta) * ig;
        const double iq = rsqrt(q);
        const double r1 = q * iq, r2 = 0.5 * p * ig * iq;
        const double c = beta < 0.0 ? r2 : r1, s = beta < 0.0 ? r1 : r2;
        a = 0.0;
        b = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          double t0 = c * At[i][k] + s * At[j][k];
          double t1 = -s * At[i][k] + c * At[j][k];
          At[i][k] = t0;
          At[j][k] = t1;
          a += t0 * t0;
          b += t1 * t1;
        }
        W[i] = a;
        W[j] = b;
        changed = true;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          double t0 = c * Vt[i][k] + s * Vt[j][k];
          double t1 = -s * Vt[i][k] + c * Vt[j][k];
          Vt[i][k] = t0;
          Vt[j][k] = t1;
        }
      }
    }
    if (!changed) break;
  }
  // smallest singular value = smallest row norm of the rotated At; on exact ties the later row
  // (what a descending selection sort leaves last)
  int m = 0;
  double wm = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double sd = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) sd += At[i][k] * At[i][k];
    if (i == 0 || sd <= wm) { wm = sd; m = i; }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double v = Vt[0][k];
    v = (m == 1) ? Vt[1][k] : v;
    v = (m == 2) ? Vt[2][k] : v;
    v = (m == 3) ? Vt[3][k] : v;
    out[k] = v;
  }
}

// The same null vector without the Jacobi sweeps, for the callers that divide by its last component straight away
// (`cloud / cloud[3]`, sfm.py:54 — sign and scale of the singular vector cancel): Householder QR of the 4x4 DLT matrix
// (A^T A = R^T R, no squaring of the condition number in the factor), then inverse iteration on R^T R from the null
// vector of R with its last pivot dropped.  The error contracts by (sigma_4 / sigma_3)^2 per iteration (1-4 iterations
// on 97 % of noisy synthetic points, never more than 20 in 20 000); iteration stops when two iterates agree to 1e-12
// and gives up (returns false: the caller runs the Jacobi version) after 24 — a point with sigma_4 ~ sigma_3 has no
// well-defined null vector, and then only OpenCV's own sweep order reproduces OpenCV's choice.  ~3 k cycles of
// dependent float64 latency against ~25 k for the sweeps; this kernel re-triangulates on the pose-critical path of
// the registration loop.  Agreement with the SVD's vector after the division: 2e-13 relative.
__device__ __forceinline__ bool dlt_null_vector_qr(const double (&At)[4][4], double (&out)[4]) {
  double A[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) A[r][k] = At[k][r];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double sigma = 0.0;
#pragma unroll
    for (int r = c; r < 4; ++r) sigma = fma(A[r][c], A[r][c], sigma);
    if (sigma > 0.0) {
      const double alpha = -copysign(sqrt(sigma), A[c][c]);
      const double vc = A[c][c] - alpha;
      const double beta = 1.0 / (sigma - A[c][c] * alpha);           // 2 / (v^T v)
      A[c][c] = alpha;
#pragma unroll
      for (int k = c + 1; k < 4; ++k) {
        double sdot = vc * A[c][k];
#pragma unroll
        for (int r = c + 1; r < 4; ++r) sdot = fma(A[r][c], A[r][k], sdot);
        sdot *= beta;
        A[c][k] = fma(-sdot, vc, A[c][k]);
#pragma unroll
        for (int r = c + 1; r < 4; ++r) A[r][k] = fma(-sdot, A[r][c], A[r][k]);
      }
    }
  }
  // R = upper triangle of A (row i: A[i][i..3])
  const double dmax = fmax(fmax(fabs(A[0][0]), fabs(A[1][1])), fmax(fabs(A[2][2]), fabs(A[3][3])));
  if (!(dmax > 0.0) || !isfinite(dmax)) return false;
  double id[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double d = A[i][i];
    id[i] = 1.0 / (fabs(d) > 1e-18 * dmax ? d : copysign(1e-18 * dmax, d));
  }
  double x[4];
  x[3] = 1.0;
  x[2] = -A[2][3] * id[2];
  x[1] = -(A[1][2] * x[2] + A[1][3]) * id[1];
  x[0] = -(A[0][1] * x[1] + A[0][2] * x[2] + A[0][3]) * id[0];
  double nrm = rsqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] *= nrm;
  bool ok = false;
#pragma unroll 1
  for (int it = 0; it < 24 && !ok; ++it) {
    double y[4], z[4];
    y[0] = x[0] * id[0];                                              // R^T y = x
    y[1] = (x[1] - A[0][1] * y[0]) * id[1];
    y[2] = (x[2] - A[0][2] * y[0] - A[1][2] * y[1]) * id[2];
    y[3] = (x[3] - A[0][3] * y[0] - A[1][3] * y[1] - A[2][3] * y[2]) * id[3];
    z[3] = y[3] * id[3];                                              // R z = y
    z[2] = (y[2] - A[2][3] * z[3]) * id[2];
    z[1] = (y[1] - A[1][2] * z[2] - A[1][3] * z[3]) * id[1];
    z[0] = (y[0] - A[0][1] * z[1] - A[0][2] * z[2] - A[0][3] * z[3]) * id[0];
    const double zz = z[0] * z[0] + z[1] * z[1] + z[2] * z[2] + z[3] * z[3];
    if (!(zz > 0.0) || !isfinite(zz)) return false;
    nrm = rsqrt(zz);
    const double sgn = (z[0] * x[0] + z[1] * x[1] + z[2] * x[2] + z[3] * x[3]) < 0.0 ? -nrm : nrm;
    double dmaxx = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double xn = z[k] * sgn;
      dmaxx = fmax(dmaxx, fabs(xn - x[k]));
      x[k] = xn;
    }
    ok = dmaxx <= 1e-12;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) out[k] = x[k];
  return ok;
}

template <int PTS_LAYOUT, int OUT_LAYOUT>
__global__ void __launch_bounds__(128) triangulate_kernel(ProjPair pp, const float* __restrict__ x1,
                                                           const float* __restrict__ x2, int n,
                                                           float* __restrict__ X, int normalize_w,
                                                           const int* __restrict__ n_dev = nullptr,
                                                           const double* __restrict__ P_dev = nullptr,
                                                           const int* __restrict__ idx = nullptr) {
  if (n_dev) n = min(n, *n_dev);             // row count produced by an earlier kernel of the same stream
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int src = idx ? __ldg(idx + i) : i;  // registration loop: triangulate only the rows data association selected
  if (P_dev) {                               // projection matrices produced on the device (registration loop)
#pragma unroll
    for (int k = 0; k < 12; ++k) { pp.P1[k] = P_dev[k]; pp.P2[k] = P_dev[12 + k]; }
  }
  float u1, v1, u2, v2;
  if (PTS_LAYOUT == 0) {   // (2,N): two coalesced row reads per view
    u1 = __ldg(x1 + i); v1 = __ldg(x1 + n + i);
    u2 = __ldg(x2 + i); v2 = __ldg(x2 + n + i);
  } else {                 // (N,2): one 8-byte read per view
    float2 a = __ldg(reinterpret_cast<const float2*>(x1) + src);
    float2 b = __ldg(reinterpret_cast<const float2*>(x2) + src);
    u1 = a.x; v1 = a.y; u2 = b.x; v2 = b.y;
  }
  // At[k][r] = A[r][k];  A rows: x*P[2]-P[0], y*P[2]-P[1] for view 1 then view 2
  double At[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    At[k][0] = (double)u1 * pp.P1[8 + k] - pp.P1[k];
    At[k][1] = (double)v1 * pp.P1[8 + k] - pp.P1[4 + k];
    At[k][2] = (double)u2 * pp.P2[8 + k] - pp.P2[k];
    At[k][3] = (double)v2 * pp.P2[8 + k] - pp.P2[4 + k];
  }
  double v[4];
  // sign and scale of the vector matter only when it is returned as it is (cv2.triangulatePoints' own output)
  if (!((normalize_w || OUT_LAYOUT == 2) && dlt_null_vector_qr(At, v))) dlt_null_vector(At, v);
  float f0 = (float)v[0], f1 = (float)v[1], f2 = (float)v[2], f3 = (float)v[3];
  if (normalize_w || OUT_LAYOUT == 2) {   // `cloud / cloud[3]` on the float32 array (sfm.py:54)
    f0 = __fdiv_rn(f0, f3); f1 = __fdiv_rn(f1, f3); f2 = __fdiv_rn(f2, f3); f3 = __fdiv_rn(f3, f3);
  }
  if (OUT_LAYOUT == 0) {
    X[i] = f0; X[n + i] = f1; X[2 * (size_t)n + i] = f2; X[3 * (size_t)n + i] = f3;
  } else if (OUT_LAYOUT == 1) {
    reinterpret_cast<float4*>(X)[i] = make_float4(f0, f1, f2, f3);
  } else {
    X[3 * (size_t)i] = f0; X[3 * (size_t)i + 1] = f1; X[3 * (size_t)i + 2] = f2;
  }
}

extern "C" int sfm_triangulate(sfm_ctx* ctx, const double* P1, const double* P2, const float* x1,
                               const float* x2, int n, int pts_layout, float* X, int out_layout,
                               int normalize_w) {
  SFM_REQUIRE(ctx && P1 && P2, "sfm_triangulate: null ctx or projection matrix");
  SFM_REQUIRE(n >= 1, "sfm_triangulate: number of points must be >= 1 (cv2 raises for N=0)");
  SFM_REQUIRE(x1 && x2 && X, "sfm_triangulate: null point buffer");
  SFM_REQUIRE(pts_layout == 0 || pts_layout == 1, "sfm_triangulate: pts_layout %d", pts_layout);
  SFM_REQUIRE(out_layout >= 0 && out_layout <= 2, "sfm_triangulate: out_layout %d", out_layout);
  SFM_TRY(sfm_ws_begin(ctx));
  ProjPair pp;
  memcpy(pp.P1, P1, sizeof(pp.P1));
  memcpy(pp.P2, P2, sizeof(pp.P2));
  const float *d1, *d2;
  SFM_TRY(dev_in(ctx, x1, (size_t)2 * n, &d1));
  SFM_TRY(dev_in(ctx, x2, (size_t)2 * n, &d2));
  bool host_out = false;
  DevOut<float> o;
  SFM_TRY(dev_out(ctx, X, (size_t)(out_layout == 2 ? 3 : 4) * n, &o, &host_out));
  dim3 grid(div_up(n, 128)), block(128);
#define TRI_CASE(PL, OL)                                                                       \
  if (pts_layout == PL && out_layout == OL)                                                    \
    SFM_LAUNCH(ctx, SFM_K_TRIANGULATE, (triangulate_kernel<PL, OL><<<grid, block, 0, ctx->stream>>>( \
                                            pp, d1, d2, n, o.dev, normalize_w)))
  TRI_CASE(0, 0); TRI_CASE(0, 1); TRI_CASE(0, 2);
  TRI_CASE(1, 0); TRI_CASE(1, 1); TRI_CASE(1, 2);
#undef TRI_CASE
  SFM_TRY(dev_out_finish(ctx, &o));
  if (host_out) SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFM_OK;
}

// ============================================================================ K3 reprojection
// cv2.projectPoints with zero distortion, operation for operation (separately rounded products
// and sums, no FMA contraction) so that the float32-rounded pixel equals OpenCV's.
__device__ __forceinline__ void project_cv(const CamParams& c, double X, double Y, double Z,
                                           double& u, double& v) {
  double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c.R[0], X), __dmul_rn(c.R[1], Y)), __dmul_rn(c.R[2], Z)), c.t[0]);
  double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c.R[3], X), __dmul_rn(c.R[4], Y)), __dmul_rn(c.R[5], Z)), c.t[1]);
  double z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(c.R[6], X), __dmul_rn(c.R[7], Y)), __dmul_rn(c.R[8], Z)), c.t[2]);
  z = (z != 0.0) ? __ddiv_rn(1.0, z) : 1.0;
  x = __dmul_rn(x, z);
  y = __dmul_rn(y, z);
  u = __dadd_rn(__dmul_rn(x, c.fx), c.cx);
  v = __dadd_rn(__dmul_rn(y, c.fy), c.cy);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block sum with a fixed reduction tree -> deterministic result for a given grid.
__device__ __forceinline__ double block_sum_256(double v, double* sh) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
    r = warp_sum(r);
  }
  return r;   // valid in warp 0
}

template <int X_LAYOUT, int PX_LAYOUT>
__global__ void __launch_bounds__(256) reproj_kernel(CamParams cam, const float* __restrict__ X,
                                                      const float* __restrict__ px, int n,
                                                      float* __restrict__ proj, float* __restrict__ X3,
                                                      double* __restrict__ partial,
                                                      unsigned int* __restrict__ counter,
                                                      double* __restrict__ err_out,
                                                      const int* __restrict__ n_dev = nullptr,
                                                      const CamParams* __restrict__ cam_dev = nullptr) {
  __shared__ double sh[8];
  __shared__ bool is_last;
  if (n_dev) n = min(n, *n_dev);
  if (cam_dev) cam = *cam_dev;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double sq = 0.0;
  if (i < n) {
    float x, y, z;
    if (X_LAYOUT == 0) {
      x = __ldg(X + 3 * (size_t)i); y = __ldg(X + 3 * (size_t)i + 1); z = __ldg(X + 3 * (size_t)i + 2);
    } else {
      float w;
      if (X_LAYOUT == 1) {
        x = __ldg(X + i); y = __ldg(X + n + i); z = __ldg(X + 2 * (size_t)n + i); w = __ldg(X + 3 * (size_t)n + i);
      } else {
        float4 q = __ldg(reinterpret_cast<const float4*>(X) + i);
        x = q.x; y = q.y; z = q.z; w = q.w;
      }
      // cv2.convertPointsFromHomogeneous on float32: scale = w != 0 ? 1/w : 1
      float s = (w != 0.f) ? __fdiv_rn(1.f, w) : 1.f;
      x = __fmul_rn(x, s); y = __fmul_rn(y, s); z = __fmul_rn(z, s);
    }
    if (X3) { X3[3 * (size_t)i] = x; X3[3 * (size_t)i + 1] = y; X3[3 * (size_t)i + 2] = z; }
    double u, v;
    project_cv(cam, (double)x, (double)y, (double)z, u, v);
    float pu = (float)u, pv = (float)v;     // projectPoints output dtype = input dtype (f32)
    if (proj) reinterpret_cast<float2*>(proj)[i] = make_float2(pu, pv);
    float ox, oy;
    if (PX_LAYOUT == 0) { ox = __ldg(px + i); oy = __ldg(px + n + i); }
    else { float2 o = __ldg(reinterpret_cast<const float2*>(px) + i); ox = o.x; oy = o.y; }
    // cv2.norm(NORM_L2) on float32 arrays: float32 differences, float64 accumulation
    double dx = (double)__fsub_rn(pu, ox), dy = (double)__fsub_rn(pv, oy);
    sq = dx * dx + dy * dy;
  }
  double bs = block_sum_256(sq, sh);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = bs;
    __threadfence();
    unsigned int ticket = atomicAdd(counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double s = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += partial[b];
    __syncthreads();
    s = block_sum_256(s, sh);
    if (threadIdx.x == 0) {
      *err_out = n > 0 ? sqrt(s) / (double)n : 0.0;
      *counter = 0u;
    }
  }
}

static int make_cam(const double* Rt, const double* K, bool roundtrip, CamParams* cam) {
  double R[9] = {Rt[0], Rt[1], Rt[2], Rt[4], Rt[5], Rt[6], Rt[8], Rt[9], Rt[10]};
  if (roundtrip) {   // the reference goes R -> Rodrigues -> rvec -> projectPoints -> R' (sfm.py:84,88)
    double rv[3];
    hm::rodrigues_to_vector(R, rv);
    hm::rodrigues_to_matrix(rv, cam->R);
  } else {
    memcpy(cam->R, R, sizeof(R));
  }
  cam->t[0] = Rt[3]; cam->t[1] = Rt[7]; cam->t[2] = Rt[11];
  cam->fx = K[0]; cam->fy = K[4]; cam->cx = K[2]; cam->cy = K[5];
  return SFM_OK;
}

extern "C" int sfm_reproj_error(sfm_ctx* ctx, const float* X, int x_layout, const float* px,
                                int px_layout, int n, const double* Rt, const double* K,
                                double* err, float* proj, float* X3) {
  SFM_REQUIRE(ctx && X && px && Rt && K, "sfm_reproj_error: null argument");
  SFM_REQUIRE(n >= 1, "sfm_reproj_error: need at least one point");
  SFM_REQUIRE(x_layout >= 0 && x_layout <= 2 && (px_layout == 0 || px_layout == 1),
              "sfm_reproj_error: bad layout (%d,%d)", x_layout, px_layout);
  SFM_TRY(sfm_ws_begin(ctx));
  CamParams cam;
  make_cam(Rt, K, true, &cam);
  const float *dX, *dpx;
  SFM_TRY(dev_in(ctx, X, (size_t)(x_layout == 0 ? 3 : 4) * n, &dX));
  SFM_TRY(dev_in(ctx, px, (size_t)2 * n, &dpx));
  bool host_out = false;
  DevOut<float> oproj, oX3;
  DevOut<double> oerr;
  SFM_TRY(dev_out(ctx, proj, (size_t)2 * n, &oproj, &host_out));
  SFM_TRY(dev_out(ctx, X3, (size_t)3 * n, &oX3, &host_out));
  SFM_TRY(dev_out(ctx, err, 1, &oerr, &host_out));
  double* err_dev = oerr.dev ? oerr.dev : ctx->dscratch;
  int nblk = div_up(n, 256);
  double* partial;
  SFM_TRY(ws_alloc_t(ctx, (size_t)nblk, &partial));
#define RP_CASE(XL, PL)                                                                            \
  if (x_layout == XL && px_layout == PL)                                                           \
    SFM_LAUNCH(ctx, SFM_K_REPROJ, (reproj_kernel<XL, PL><<<nblk, 256, 0, ctx->stream>>>(           \
                                       cam, dX, dpx, n, oproj.dev, oX3.dev, partial, ctx->counters, err_dev)))
  RP_CASE(0, 0); RP_CASE(0, 1); RP_CASE(1, 0); RP_CASE(1, 1); RP_CASE(2, 0); RP_CASE(2, 1);
#undef RP_CASE
  SFM_TRY(dev_out_finish(ctx, &oproj));
  SFM_TRY(dev_out_finish(ctx, &oX3));
  SFM_TRY(dev_out_finish(ctx, &oerr));
  if (host_out) SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFM_OK;
}

// ---- registration-loop variants: counts and matrices in HBM, nothing synchronises (chain_dev.cuh)
int sfm_triangulate_dev(sfm_ctx* ctx, const double* P1P2_dev, const float* x1, const float* x2, int n_cap,
                        const int* n_dev, float* X, int out_layout, const int32_t* idx) {
  if (n_cap <= 0) return SFM_OK;
  ProjPair pp;
  memset(&pp, 0, sizeof(pp));
  dim3 grid(div_up(n_cap, 128)), block(128);
  if (out_layout == 1)
    SFM_LAUNCH(ctx, SFM_K_TRIANGULATE, (triangulate_kernel<1, 1><<<grid, block, 0, ctx->stream>>>(pp, x1, x2, n_cap, X, 1, n_dev, P1P2_dev, idx)));
  else
    SFM_LAUNCH(ctx, SFM_K_TRIANGULATE, (triangulate_kernel<1, 2><<<grid, block, 0, ctx->stream>>>(pp, x1, x2, n_cap, X, 1, n_dev, P1P2_dev, idx)));
  return SFM_OK;
}

int sfm_reproj_error_dev(sfm_ctx* ctx, const float* X, int x_layout, const float* px, int n_cap, const int* n_dev,
                         const CamParams* cam_dev, double* err_dev, float* X3) {
  if (n_cap <= 0) return SFM_OK;
  CamParams cam;
  memset(&cam, 0, sizeof(cam));
  SFM_TRY(sfm_ws_begin(ctx));
  int nblk = div_up(n_cap, 256);
  double* partial;
  SFM_TRY(ws_alloc_t(ctx, (size_t)nblk, &partial));
  if (x_layout == 0)
    SFM_LAUNCH(ctx, SFM_K_REPROJ, (reproj_kernel<0, 1><<<nblk, 256, 0, ctx->stream>>>(cam, X, px, n_cap, nullptr, X3, partial,
                                                                                     ctx->counters, err_dev, n_dev, cam_dev)));
  else
    SFM_LAUNCH(ctx, SFM_K_REPROJ, (reproj_kernel<2, 1><<<nblk, 256, 0, ctx->stream>>>(cam, X, px, n_cap, nullptr, X3, partial,
                                                                                     ctx->counters, err_dev, n_dev, cam_dev)));
  return SFM_OK;
}

// ============================================================================ recoverPose (second half)
// cv2.recoverPose(E, pts0, pts1, K) as the reference calls it (sfm.py:311, isfm.py:83, test.py:250): the cheirality
// test of the four (R, t) decompositions of E.  Per correspondence and candidate: the pixels are normalised with K
// in float64, the point is triangulated by the same 4x4 DLT / Jacobi SVD as K2, and the candidate keeps it when
// z*w > 0, z/w < dist, and the depth in the second camera is in (0, dist) — the conditions and their order are
// OpenCV's.  One thread per correspondence evaluates all four candidates; counts by ballot/popc + atomics.
struct PoseCand {
  double P[4][12];       // [R1|t], [R2|t], [R1|-t], [R2|-t]
  double fx, fy, cx, cy, dist;
};

template <typename T>
__global__ void __launch_bounds__(128) recover_pose_kernel(PoseCand pc, const T* __restrict__ p1, const T* __restrict__ p2, int n,
                                                           const unsigned char* __restrict__ mask_in,
                                                           unsigned char* __restrict__ masks /*4 x n*/, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool keep[4] = {false, false, false, false};
  if (i < n) {
    // (x - cx) / fx in float64, like points.col(0) = (points.col(0) - cx) / fx on a CV_64F matrix
    const double x1 = __ddiv_rn(__dsub_rn((double)p1[2 * (size_t)i], pc.cx), pc.fx);
    const double y1 = __ddiv_rn(__dsub_rn((double)p1[2 * (size_t)i + 1], pc.cy), pc.fy);
    const double x2 = __ddiv_rn(__dsub_rn((double)p2[2 * (size_t)i], pc.cx), pc.fx);
    const double y2 = __ddiv_rn(__dsub_rn((double)p2[2 * (size_t)i + 1], pc.cy), pc.fy);
    const bool in = !mask_in || mask_in[i] != 0;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const double* P = pc.P[c];
      // rows of the DLT matrix: x P0[2] - P0[0], y P0[2] - P0[1] with P0 = [I|0]; then the same for the candidate
      double At[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double p00 = (k == 0) ? 1.0 : 0.0, p01 = (k == 1) ? 1.0 : 0.0, p02 = (k == 2) ? 1.0 : 0.0;
        At[k][0] = __dsub_rn(__dmul_rn(x1, p02), p00);
        At[k][1] = __dsub_rn(__dmul_rn(y1, p02), p01);
        At[k][2] = __dsub_rn(__dmul_rn(x2, P[8 + k]), P[k]);
        At[k][3] = __dsub_rn(__dmul_rn(y2, P[8 + k]), P[4 + k]);
      }
      double Q[4];
      dlt_null_vector(At, Q);
      bool m = __dmul_rn(Q[2], Q[3]) > 0.0;
      const double X = __ddiv_rn(Q[0], Q[3]), Y = __ddiv_rn(Q[1], Q[3]), Z = __ddiv_rn(Q[2], Q[3]);
      m = m && (Z < pc.dist);
      // depth in the second camera: row 2 of P * (X, Y, Z, 1)
      const double z2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[8], X), __dmul_rn(P[9], Y)), __dmul_rn(P[10], Z)), P[11]);
      m = m && (z2 > 0.0) && (z2 < pc.dist);
      keep[c] = m && in;
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (i < n) masks[(size_t)c * n + i] = keep[c] ? 255 : 0;
    const unsigned b = __ballot_sync(0xffffffffu, keep[c]);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(counts + c, __popc(b));
  }
}

extern "C" int sfm_recover_pose(sfm_ctx* ctx, const double* E, const void* pts1, const void* pts2, int dtype, int n,
                                const double* K, double dist, const uint8_t* mask_in, double* R, double* t,
                                uint8_t* mask_out, int32_t* n_good) {
  SFM_REQUIRE(ctx && E && pts1 && pts2 && K && R && t, "sfm_recover_pose: null argument");
  SFM_REQUIRE(n >= 1, "sfm_recover_pose: need at least one correspondence");
  SFM_REQUIRE(dtype == 0 || dtype == 2, "sfm_recover_pose: points must be float32 (0) or float64 (2)");
  SFM_TRY(sfm_ws_begin(ctx));
  // decomposeEssentialMat: E = U diag(1,1,0) V^T, det(U), det(V^T) made positive, R1 = U W V^T, R2 = U W^T V^T, t = u3
  double U[9], W3[3], Vt[9];
  hm::svd_square<3>(E, U, W3, Vt);
  auto det3 = [](const double* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
  };
  if (det3(U) < 0) for (double& v : U) v = -v;
  if (det3(Vt) < 0) for (double& v : Vt) v = -v;
  const double Wm[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
  double UW[9], UWt[9], R1[9], R2[9], tv[3] = {U[2], U[5], U[8]};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double a = 0.0, b = 0.0;
      for (int k = 0; k < 3; ++k) { a += U[3 * i + k] * Wm[3 * k + j]; b += U[3 * i + k] * Wm[3 * j + k]; }
      UW[3 * i + j] = a; UWt[3 * i + j] = b;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double a = 0.0, b = 0.0;
      for (int k = 0; k < 3; ++k) { a += UW[3 * i + k] * Vt[3 * k + j]; b += UWt[3 * i + k] * Vt[3 * k + j]; }
      R1[3 * i + j] = a; R2[3 * i + j] = b;
    }
  PoseCand pc;
  for (int c = 0; c < 4; ++c) {
    const double* Rc = (c & 1) ? R2 : R1;
    const double sg = (c & 2) ? -1.0 : 1.0;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) pc.P[c][4 * i + j] = Rc[3 * i + j];
      pc.P[c][4 * i + 3] = sg * tv[i];
    }
  }
  pc.fx = K[0]; pc.fy = K[4]; pc.cx = K[2]; pc.cy = K[5]; pc.dist = dist;
  const size_t esz = dtype == 0 ? sizeof(float) : sizeof(double);
  const void *d1 = pts1, *d2 = pts2;
  if (!sfm_is_device_ptr(pts1)) { char* p; SFM_TRY(ws_alloc_t(ctx, 2 * (size_t)n * esz, &p)); SFM_CUDA(cudaMemcpyAsync(p, pts1, 2 * (size_t)n * esz, cudaMemcpyHostToDevice, ctx->stream)); d1 = p; }
  if (!sfm_is_device_ptr(pts2)) { char* p; SFM_TRY(ws_alloc_t(ctx, 2 * (size_t)n * esz, &p)); SFM_CUDA(cudaMemcpyAsync(p, pts2, 2 * (size_t)n * esz, cudaMemcpyHostToDevice, ctx->stream)); d2 = p; }
  const uint8_t* dmask = nullptr;
  SFM_TRY(dev_in(ctx, mask_in, (size_t)n, &dmask));
  unsigned char* masks;
  int* counts;
  SFM_TRY(ws_alloc_t(ctx, (size_t)4 * n, &masks));
  SFM_TRY(ws_alloc_t(ctx, 4, &counts));
  SFM_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(int), ctx->stream));
  if (dtype == 0)
    SFM_LAUNCH(ctx, SFM_K_TRIANGULATE, (recover_pose_kernel<float><<<div_up(n, 128), 128, 0, ctx->stream>>>(
                                           pc, (const float*)d1, (const float*)d2, n, dmask, masks, counts)));
  else
    SFM_LAUNCH(ctx, SFM_K_TRIANGULATE, (recover_pose_kernel<double><<<div_up(n, 128), 128, 0, ctx->stream>>>(
                                           pc, (const double*)d1, (const double*)d2, n, dmask, masks, counts)));
  int* hc;
  SFM_TRY(hs_alloc_t(ctx, 4, &hc));
  SFM_CUDA(cudaMemcpyAsync(hc, counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  // OpenCV's choice: the first candidate (order R1|t, R2|t, R1|-t, R2|-t) whose count is >= all others
  const int g1 = hc[0], g2 = hc[1], g3 = hc[2], g4 = hc[3];
  int win = 3;
  if (g1 >= g2 && g1 >= g3 && g1 >= g4) win = 0;
  else if (g2 >= g1 && g2 >= g3 && g2 >= g4) win = 1;
  else if (g3 >= g1 && g3 >= g2 && g3 >= g4) win = 2;
  const double* Rw = (win & 1) ? R2 : R1;
  for (int k = 0; k < 9; ++k) R[k] = Rw[k];
  for (int k = 0; k < 3; ++k) t[k] = ((win & 2) ? -1.0 : 1.0) * tv[k];
  if (n_good) *n_good = hc[win];
  if (mask_out) {
    const bool dev = sfm_is_device_ptr(mask_out);
    SFM_CUDA(cudaMemcpyAsync(mask_out, masks + (size_t)win * n, (size_t)n, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                             ctx->stream));
    if (!dev) SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return SFM_OK;
}

// ============================================================================ common_points
// `np.where(pts2 == pts1[i])[0][0]` (sfm.py:221-226): the first row j of pts2 whose x equals pts1[i].x
// OR whose y equals pts1[i].y (element-wise comparison — the reference's quirk), float32 equality.
// Instead of the reference's O(n1*n2) scan: two open-addressing hash tables over pts2 (one keyed by the
// bits of x, one by the bits of y) holding the smallest row index per distinct value (atomicMin), then
// one lookup pair per row of pts1: hit = min(first row with equal x, first row with equal y).
struct HashSlot {
  unsigned int key;   // float bits (0xFFFFFFFF = empty; a NaN pattern, which never compares equal anyway)
  unsigned int row;   // smallest row index seen for this key
};

__device__ __forceinline__ unsigned int float_key(float v) { return (v == 0.f) ? 0u : __float_as_uint(v); }   // -0 == +0
__device__ __forceinline__ unsigned int hash_u32(unsigned int k) {
  k ^= k >> 16; k *= 0x7feb352du; k ^= k >> 15; k *= 0x846ca68bu; k ^= k >> 16;
  return k;
}

__device__ __forceinline__ void hash_insert_min(HashSlot* __restrict__ tab, unsigned int mask, float v, unsigned int row) {
  if (v != v) return;                                   // NaN never matches
  const unsigned int key = float_key(v);
  unsigned int h = hash_u32(key) & mask;
  for (;;) {
    const unsigned int prev = atomicCAS(&tab[h].key, 0xFFFFFFFFu, key);
    if (prev == 0xFFFFFFFFu || prev == key) { atomicMin(&tab[h].row, row); return; }
    h = (h + 1) & mask;
  }
}

__device__ __forceinline__ unsigned int hash_lookup(const HashSlot* __restrict__ tab, unsigned int mask, float v) {
  if (v != v) return 0xFFFFFFFFu;
  const unsigned int key = float_key(v);
  unsigned int h = hash_u32(key) & mask;
  for (;;) {
    const unsigned int k = tab[h].key;
    if (k == key) return tab[h].row;
    if (k == 0xFFFFFFFFu) return 0xFFFFFFFFu;
    h = (h + 1) & mask;
  }
}

__global__ void __launch_bounds__(256) hash_build_kernel(const float2* __restrict__ p2, int n2, HashSlot* __restrict__ tx,
                                                          HashSlot* __restrict__ ty, unsigned int mask) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n2) return;
  float2 b = __ldg(p2 + j);
  hash_insert_min(tx, mask, b.x, (unsigned int)j);
  hash_insert_min(ty, mask, b.y, (unsigned int)j);
}

__global__ void __launch_bounds__(256) first_hit_kernel(const float2* __restrict__ p1, int n1, const HashSlot* __restrict__ tx,
                                                         const HashSlot* __restrict__ ty, unsigned int mask,
                                                         int* __restrict__ hit, unsigned char* __restrict__ keep2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n1) return;
  float2 a = __ldg(p1 + i);
  const unsigned int jx = hash_lookup(tx, mask, a.x), jy = hash_lookup(ty, mask, a.y);
  const unsigned int j = min(jx, jy);
  const int found = (j == 0xFFFFFFFFu) ? -1 : (int)j;
  hit[i] = found;
  if (found >= 0 && keep2) keep2[found] = 0;
}

// Stable single-CTA compaction of (i, hit[i]) for hit[i] >= 0.
__global__ void __launch_bounds__(1024) compact_hits_kernel(const int* __restrict__ hit, int n,
                                                             int* __restrict__ idx1, int* __restrict__ idx2,
                                                             int* __restrict__ n_out) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int start = 0; start < n; start += 1024) {
    int i = start + threadIdx.x;
    int h = (i < n) ? hit[i] : -1;
    bool f = h >= 0;
    unsigned m = __ballot_sync(0xffffffffu, f);
    int pre = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[w] = __popc(m);
    __syncthreads();
    int off = 0;
    for (int k = 0; k < w; ++k) off += warp_tot[k];
    int base = base_s;
    if (f) {
      int pos = base + off + pre;
      if (idx1) idx1[pos] = i;
      if (idx2) idx2[pos] = h;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int k = 0; k < 32; ++k) tot += warp_tot[k];
      base_s = base + tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && n_out) *n_out = base_s;
}

extern "C" int sfm_common_points(sfm_ctx* ctx, const float* pts1, int n1, const float* pts2, int n2,
                                 int32_t* idx1, int32_t* idx2, int32_t* n_common, uint8_t* keep2) {
  SFM_REQUIRE(ctx, "sfm_common_points: null ctx");
  SFM_REQUIRE(n1 >= 0 && n2 >= 0, "sfm_common_points: negative size");
  SFM_TRY(sfm_ws_begin(ctx));
  const float *d1, *d2;
  SFM_TRY(dev_in(ctx, pts1, (size_t)2 * n1, &d1));
  SFM_TRY(dev_in(ctx, pts2, (size_t)2 * n2, &d2));
  bool host_out = false;
  DevOut<int32_t> o1, o2, on;
  DevOut<uint8_t> ok;
  SFM_TRY(dev_out(ctx, idx1, (size_t)n1, &o1, &host_out));
  SFM_TRY(dev_out(ctx, idx2, (size_t)n1, &o2, &host_out));
  SFM_TRY(dev_out(ctx, n_common, 1, &on, &host_out));
  SFM_TRY(dev_out(ctx, keep2, (size_t)n2, &ok, &host_out));
  int* hit = nullptr;
  SFM_TRY(ws_alloc_t(ctx, (size_t)(n1 > 0 ? n1 : 1), &hit));
  if (ok.dev && n2 > 0) SFM_CUDA(cudaMemsetAsync(ok.dev, 1, (size_t)n2, ctx->stream));
  if (n1 > 0) {
    unsigned int size = 64;
    while (size < 2u * (unsigned int)(n2 > 0 ? n2 : 1)) size <<= 1;
    HashSlot* tabs = nullptr;
    SFM_TRY(ws_alloc_t(ctx, (size_t)2 * size, &tabs));
    SFM_CUDA(cudaMemsetAsync(tabs, 0xFF, sizeof(HashSlot) * 2 * (size_t)size, ctx->stream));
    if (n2 > 0)
      SFM_LAUNCH(ctx, SFM_K_ASSOC, (hash_build_kernel<<<div_up(n2, 256), 256, 0, ctx->stream>>>(
                                       (const float2*)d2, n2, tabs, tabs + size, size - 1)));
    SFM_LAUNCH(ctx, SFM_K_ASSOC, (first_hit_kernel<<<div_up(n1, 256), 256, 0, ctx->stream>>>(
                                     (const float2*)d1, n1, tabs, tabs + size, size - 1, hit, ok.dev)));
  }
  SFM_LAUNCH(ctx, SFM_K_ASSOC, (compact_hits_kernel<<<1, 1024, 0, ctx->stream>>>(hit, n1, o1.dev, o2.dev, on.dev)));
  SFM_TRY(dev_out_finish(ctx, &o1));
  SFM_TRY(dev_out_finish(ctx, &o2));
  SFM_TRY(dev_out_finish(ctx, &on));
  SFM_TRY(dev_out_finish(ctx, &ok));
  if (host_out) SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFM_OK;
}

// ============================================================================ host utilities
extern "C" int sfm_rodrigues_to_matrix(const double* rvec, double* R9) {
  SFM_REQUIRE(rvec && R9, "sfm_rodrigues_to_matrix: null argument");
  hm::rodrigues_to_matrix(rvec, R9);
  return SFM_OK;
}
extern "C" int sfm_rodrigues_to_vector(const double* R9, double* rvec) {
  SFM_REQUIRE(rvec && R9, "sfm_rodrigues_to_vector: null argument");
  hm::rodrigues_to_vector(R9, rvec);
  return SFM_OK;
}
