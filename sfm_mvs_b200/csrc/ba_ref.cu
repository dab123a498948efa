// ba_ref.cu — the reference's single-camera BundleAdjustment (sfm.py:138-157) as the reference formulates it:
// x = [Rt 12 | K 9 | observed pixels (2,N) | points (N,3)] — the pose as twelve unconstrained numbers, K and the
// observations free — residual OptimReprojectionError(x) = ((p - proj)^2).ravel() / N (sfm.py:104-136), minimised by
// scipy.optimize.least_squares (TRF) under a dense 2-point finite-difference Jacobian: 22 + 5N evaluations of the
// residual per Jacobian, each a Python loop over the points ("close to half a minute per frame", sfm.py:378).
//
// The optimiser stays the reference's (scipy drives the iterations on the host, cv2_compat.BundleAdjustment); what
// moves to the GPU is the part that is data-parallel: ALL 22 + 5N residual vectors of a Jacobian in one launch, with
// scipy's own forward-difference steps (h = sqrt(eps) sign(x) max(1, |x|), made representable), so that the Jacobian
// handed back is the matrix scipy's approx_derivative would have formed from cv2's residuals.
//
// Compiled with -fmad=false: cv2.Rodrigues (3x3 -> rvec, through OpenCV's small-matrix SVD, hostmath.h) and
// cv2.projectPoints are mirrored operation for operation in float64.
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "hostmath.h"

namespace {

constexpr int NFIX = 21;     // Rt (12) + K (9)

// scipy.optimize._numdiff: h = EPS**0.5 * sign(x0) * max(1, |x0|), sign(0) = +1, then h = (x0 + h) - x0
__host__ __device__ inline double fd_step(double x0) {
  const double rel = 1.4901161193847656e-08;          // sqrt(2^-52)
  const double s = x0 >= 0.0 ? 1.0 : -1.0;
  const double h = rel * s * fmax(1.0, fabs(x0));
  return (x0 + h) - x0;
}

// 13 rotation matrices R' = Rodrigues(Rodrigues(R + h e_k)) for k = 0..11 (entries of the 3x4 [R|t] that fall into
// R) and the unperturbed one (slot 12), plus the 13 translations.  One thread per slot.
__global__ void ba_ref_pose_kernel(const double* __restrict__ x, double* __restrict__ poses /*13 x 12: R' 9 | t 3*/) {
  const int k = threadIdx.x;
  if (k > 12) return;
  double Rt[12];
  for (int i = 0; i < 12; ++i) Rt[i] = x[i];
  if (k < 12) Rt[k] = x[k] + fd_step(x[k]);
  const double R[9] = {Rt[0], Rt[1], Rt[2], Rt[4], Rt[5], Rt[6], Rt[8], Rt[9], Rt[10]};
  double rv[3], Rr[9];
  hm::rodrigues_to_vector(R, rv);
  hm::rodrigues_to_matrix(rv, Rr);
  double* P = poses + 12 * k;
  for (int i = 0; i < 9; ++i) P[i] = Rr[i];
  P[9] = Rt[3]; P[10] = Rt[7]; P[11] = Rt[11];
}

// blockIdx.y = column b of the Jacobian (b == nparams: the unperturbed residual f0), threads over the points.
// out: (nparams + 1) rows of 2N doubles.
__global__ void __launch_bounds__(128) ba_ref_residual_kernel(const double* __restrict__ x, int N, int nparams,
                                                              const double* __restrict__ poses, double* __restrict__ out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const bool pert = b < nparams;
  const double h = pert ? fd_step(x[b]) : 0.0;
  const double* P = poses + 12 * ((pert && b < 12) ? b : 12);
  // K as projectPoints reads it: fx = K[0][0], fy = K[1][1], cx = K[0][2], cy = K[1][2]
  double fx = x[12], cx = x[14], fy = x[16], cy = x[17];
  if (pert) {
    if (b == 12) fx += h;
    else if (b == 14) cx += h;
    else if (b == 16) fy += h;
    else if (b == 17) cy += h;
  }
  // observed pixels: x[21 .. 21 + 2N) laid out (2, N); points: x[21 + 2N ..) laid out (N, 3)
  const int op = NFIX, ox = NFIX + 2 * N;
  double pu = x[op + i], pv = x[op + N + i];
  double X = x[ox + 3 * i], Y = x[ox + 3 * i + 1], Z = x[ox + 3 * i + 2];
  if (pert && b >= op) {
    if (b == op + i) pu += h;
    else if (b == op + N + i) pv += h;
    else if (b == ox + 3 * i) X += h;
    else if (b == ox + 3 * i + 1) Y += h;
    else if (b == ox + 3 * i + 2) Z += h;
  }
  // cv2.projectPoints, zero distortion (geometry.cu project_cv)
  double xc = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[0], X), __dmul_rn(P[1], Y)), __dmul_rn(P[2], Z)), P[9]);
  double yc = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[3], X), __dmul_rn(P[4], Y)), __dmul_rn(P[5], Z)), P[10]);
  double zc = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[6], X), __dmul_rn(P[7], Y)), __dmul_rn(P[8], Z)), P[11]);
  zc = (zc != 0.0) ? __ddiv_rn(1.0, zc) : 1.0;
  xc = __dmul_rn(xc, zc);
  yc = __dmul_rn(yc, zc);
  const double u = __dadd_rn(__dmul_rn(xc, fx), cx), v = __dadd_rn(__dmul_rn(yc, fy), cy);
  const double du = pu - u, dv = pv - v;
  double* o = out + (size_t)b * 2 * N;
  o[2 * i] = (du * du) / N;
  o[2 * i + 1] = (dv * dv) / N;
}

// J[r][c] = (f_c[r] - f0[r]) / h_c, row-major (2N) x nparams
__global__ void ba_ref_quotient_kernel(const double* __restrict__ x, int N, int nparams, const double* __restrict__ f,
                                       double* __restrict__ J) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= nparams) return;
  const double h = fd_step(x[c]);
  J[(size_t)r * nparams + c] = (f[(size_t)c * 2 * N + r] - f[(size_t)nparams * 2 * N + r]) / h;
}

}  // namespace

extern "C" int sfm_ba_reference_fd(sfm_ctx* ctx, const double* x, int n_params, int n_points, double* f0, double* J) {
  SFM_REQUIRE(ctx && x && f0, "sfm_ba_reference_fd: null argument");
  SFM_REQUIRE(n_points >= 1 && n_params == NFIX + 5 * n_points, "sfm_ba_reference_fd: %d parameters do not describe %d points (21 + 5 N)",
              n_params, n_points);
  SFM_TRY(sfm_ws_begin(ctx));
  const int N = n_points, rows = J ? n_params + 1 : 1;
  const double* dx;
  SFM_TRY(dev_in(ctx, x, (size_t)n_params, &dx));
  double *poses, *f, *dJ = nullptr;
  SFM_TRY(ws_alloc_t(ctx, (size_t)13 * 12, &poses));
  SFM_TRY(ws_alloc_t(ctx, (size_t)(n_params + 1) * 2 * N, &f));
  SFM_LAUNCH(ctx, SFM_K_BA_EVAL, (ba_ref_pose_kernel<<<1, 32, 0, ctx->stream>>>(dx, poses)));
  if (J) {
    SFM_LAUNCH(ctx, SFM_K_BA_EVAL, (ba_ref_residual_kernel<<<dim3(div_up(N, 128), n_params + 1), 128, 0, ctx->stream>>>(dx, N, n_params, poses, f)));
    SFM_TRY(ws_alloc_t(ctx, (size_t)n_params * 2 * N, &dJ));
    SFM_LAUNCH(ctx, SFM_K_BA_EVAL, (ba_ref_quotient_kernel<<<dim3(div_up(n_params, 256), 2 * N), 256, 0, ctx->stream>>>(dx, N, n_params, f, dJ)));
    if (sfm_is_device_ptr(J)) SFM_CUDA(cudaMemcpyAsync(J, dJ, sizeof(double) * (size_t)n_params * 2 * N, cudaMemcpyDeviceToDevice, ctx->stream));
    else SFM_CUDA(cudaMemcpyAsync(J, dJ, sizeof(double) * (size_t)n_params * 2 * N, cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    // only f0: the last row index is n_params, so launch that single row by offsetting the output base
    SFM_LAUNCH(ctx, SFM_K_BA_EVAL, (ba_ref_residual_kernel<<<dim3(div_up(N, 128), 1), 128, 0, ctx->stream>>>(dx, N, 0, poses, f)));
  }
  const double* f0src = f + (size_t)(rows - 1) * 2 * N;
  SFM_CUDA(cudaMemcpyAsync(f0, f0src, sizeof(double) * 2 * N,
                           sfm_is_device_ptr(f0) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFM_OK;
}
