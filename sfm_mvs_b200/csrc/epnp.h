// epnp.h — EPnP minimal solver (Lepetit, Moreno-Noguer, Fua, "EPnP: An Accurate O(n) Solution to
// the PnP Problem", IJCV 2009) in float64, __host__ __device__: the 5-point kernel that
// cv2.solvePnPRansac runs per RANSAC iteration (reference call site sfm.py:67, test.py:319).
//
// The steps are the published ones in the exact arithmetic OpenCV's solver performs: control points = centroid +
// principal axes (SVD of the 3x3 scatter) scaled by sqrt(sigma/n); barycentric alphas through the SVD inverse of
// the control-point frame; M^T M (12x12) accumulated row by row; its SVD — OpenCV's own one-sided Jacobi (every
// decomposition on this path is smaller than the 25 rows from which OpenCV calls LAPACK), whose rotation order
// and operations are reproduced (hostmath.h jacobi_svd), so the four vectors taken from the null space of M
// (2-dimensional for 5 points, the basis inside it decided by rounding) are bit-identical to OpenCV's; L (6x10)
// and rho; the beta initialisations "N=4 linearised", "N=2", "N=3" by SVD least squares; five Gauss-Newton
// iterations each with the solver's Householder QR; absolute orientation by SVD; the pose with the smallest
// mean reprojection error wins.  Every sum is accumulated in OpenCV's order and nothing may be contracted into a
// fused multiply-add: host code is built with -ffp-contract=off, the device code that includes this header with
// -fmad=false (pnp_epnp.cu).  tests/test_host_logic.py: sfm_epnp == cv2.solvePnP(EPNP) bit for bit.
#pragma once
#include "hostmath.h"

namespace hm {

// Cyclic Jacobi eigen-decomposition of a symmetric N x N matrix (row-major, destroyed).
// On return w[] ascending, V rows = eigenvectors (V[i*N+k] = k-th component of eigenvector i).
template <int N>
HM_HD inline void eig_sym(double* A, double* w, double* V, int max_sweeps = 60) {
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < N; ++j) V[i * N + j] = 0.0;
    V[i * N + i] = 1.0;
  }
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < N; ++i) {
      diag += A[i * N + i] * A[i * N + i];
      for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
    }
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        double apq = A[p * N + q];
        if (apq == 0.0) continue;
        double app = A[p * N + p], aqq = A[q * N + q];
        if (fabs(apq) <= 1e-300) continue;
        // tan(angle) = sgn(al) apq / (|al| + hypot(al, apq)), al = (aqq - app)/2  (the textbook theta form, one
        // division and one square root cheaper)
        double al = 0.5 * (aqq - app);
        double rr = sqrt(al * al + apq * apq);
        double t = apq / (al + (al >= 0.0 ? rr : -rr));
#ifdef __CUDA_ARCH__
        double c = rsqrt(t * t + 1.0), s = t * c;
#else
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#endif
        for (int k = 0; k < N; ++k) {   // columns p,q
          double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {   // rows p,q
          double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        A[p * N + q] = 0.0;
        A[q * N + p] = 0.0;
        for (int k = 0; k < N; ++k) {
          double vp = V[p * N + k], vq = V[q * N + k];
          V[p * N + k] = c * vp - s * vq;
          V[q * N + k] = s * vp + c * vq;
        }
      }
  }
  for (int i = 0; i < N; ++i) w[i] = A[i * N + i];
  for (int i = 0; i < N - 1; ++i) {   // selection sort ascending
    int m = i;
    for (int k = i + 1; k < N; ++k)
      if (w[k] < w[m]) m = k;
    if (m != i) {
      double tw = w[i]; w[i] = w[m]; w[m] = tw;
      for (int k = 0; k < N; ++k) { double tv = V[i * N + k]; V[i * N + k] = V[m * N + k]; V[m * N + k] = tv; }
    }
  }
  // sign convention: largest-magnitude component positive
  for (int i = 0; i < N; ++i) {
    int m = 0;
    for (int k = 1; k < N; ++k)
      if (fabs(V[i * N + k]) > fabs(V[i * N + m])) m = k;
    if (V[i * N + m] < 0.0)
      for (int k = 0; k < N; ++k) V[i * N + k] = -V[i * N + k];
  }
}


struct EpnpCam { double fu, fv, uc, vc; };

HM_HD inline double epnp_dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
HM_HD inline double epnp_dist2(const double* p, const double* q) {
  return (p[0] - q[0]) * (p[0] - q[0]) + (p[1] - q[1]) * (p[1] - q[1]) + (p[2] - q[2]) * (p[2] - q[2]);
}

// cv::SVD::backSubst for one right-hand side: x = V diag(1/w) U^T b, singular values <= 2 eps sum(w) dropped.
// ut: n rows of length m (row i = i-th left singular vector), vt: n x n.
template <int M, int N>
HM_HD inline void cv_svd_backsubst(const double* w, const double* ut, const double* vt, const double* b, double* x) {
  double threshold = 0.0;
  for (int j = 0; j < N; ++j) x[j] = 0.0;
  for (int i = 0; i < N; ++i) threshold += w[i];
  threshold *= DBL_EPSILON * 2;
  for (int i = 0; i < N; ++i) {
    double wi = w[i];
    if (fabs(wi) <= threshold) continue;
    wi = 1 / wi;
    double s = 0.0;
    for (int j = 0; j < M; ++j) s += ut[i * M + j] * b[j];
    s *= wi;
    for (int j = 0; j < N; ++j) x[j] = x[j] + s * vt[i * N + j];
  }
}

// cv::invert(A, DECOMP_SVD) of a 3x3: pseudo-inverse accumulated one singular triplet at a time.
HM_HD inline void cv_invert3_svd(const double* A, double* Ainv) {
  double At[9], w[3], vt[9];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) At[i * 3 + k] = A[k * 3 + i];
  jacobi_svd<3, 3>(At, w, vt);
  double threshold = 0.0;
  for (int i = 0; i < 9; ++i) Ainv[i] = 0.0;
  for (int i = 0; i < 3; ++i) threshold += w[i];
  threshold *= DBL_EPSILON * 2;
  for (int i = 0; i < 3; ++i) {
    double wi = w[i];
    if (fabs(wi) <= threshold) continue;
    wi = 1 / wi;
    double buf[3];
    for (int j = 0; j < 3; ++j) buf[j] = At[i * 3 + j] * wi;                       // u[j][i] * wi
    for (int r = 0; r < 3; ++r)
      for (int j = 0; j < 3; ++j) Ainv[r * 3 + j] = Ainv[r * 3 + j] + vt[i * 3 + r] * buf[j];
  }
}

// ---- stage 1a: control points (cws) and barycentric coordinates (alphas: n x 4); scratch: 3n doubles
HM_HD inline void epnp_control_alphas(const double* pw, int n, double* alphas, double (*cws)[3]) {
  cws[0][0] = cws[0][1] = cws[0][2] = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 3; ++j) cws[0][j] += pw[3 * i + j];
  for (int j = 0; j < 3; ++j) cws[0][j] /= n;
  // PW0^T PW0 (cv::mulTransposed: upper triangle, sums over the points in order, mirrored)
  double S[9];
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) {
      double s0 = 0.0;
      for (int k = 0; k < n; ++k) s0 += (pw[3 * k + i] - cws[0][i]) * (pw[3 * k + j] - cws[0][j]);
      S[3 * i + j] = s0;
      S[3 * j + i] = s0;
    }
  double dc[3], vt[9];
  jacobi_svd<3, 3>(S, dc, vt);          // S symmetric: S^T == S; rows of S are now the left singular vectors
  for (int i = 1; i < 4; ++i) {
    const double k = sqrt(dc[i - 1] / n);
    for (int j = 0; j < 3; ++j) cws[i][j] = cws[0][j] + k * S[3 * (i - 1) + j];
  }
  double cc[9], ci[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
  cv_invert3_svd(cc, ci);
  for (int i = 0; i < n; ++i) {
    const double* pi = pw + 3 * i;
    double* a = alphas + 4 * i;
    for (int j = 0; j < 3; ++j)
      a[1 + j] = ci[3 * j] * (pi[0] - cws[0][0]) + ci[3 * j + 1] * (pi[1] - cws[0][1]) + ci[3 * j + 2] * (pi[2] - cws[0][2]);
    a[0] = 1.0 - a[1] - a[2] - a[3];
  }
}

// one entry of M (2n x 12): row 2i = u-equation of correspondence i, row 2i+1 = v-equation
HM_HD inline double epnp_M_entry(const double* alphas, const double* us, const EpnpCam& cam, int row, int col) {
  const int i = row >> 1, j = col / 3, comp = col - 3 * j;
  const double a = alphas[4 * i + j];
  if ((row & 1) == 0) return comp == 0 ? a * cam.fu : (comp == 1 ? 0.0 : a * (cam.uc - us[2 * i]));
  return comp == 0 ? 0.0 : (comp == 1 ? a * cam.fv : a * (cam.vc - us[2 * i + 1]));
}

// one entry (r <= c) of M^T M: the sum over the rows of M in order (cv::mulTransposed)
HM_HD inline double epnp_MtM_entry(const double* alphas, const double* us, int n, const EpnpCam& cam, int r, int c) {
  double s0 = 0.0;
  for (int k = 0; k < 2 * n; ++k) s0 += epnp_M_entry(alphas, us, cam, k, r) * epnp_M_entry(alphas, us, cam, k, c);
  return s0;
}

// ---- stage 2 (after the SVD): one entry of L (6 x 10) from the four null-space vectors v[0] (smallest) .. v[3]
HM_HD inline double epnp_L_entry(const double* const v[4], int i, int col) {
  // row i <-> control-point pair (a,b) in the order (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
  const int pa = (i < 3) ? 0 : ((i < 5) ? 1 : 2);
  const int pb = (i < 3) ? i + 1 : ((i < 5) ? i - 1 : 3);
  // columns: 0 (0,0) 1 (0,1) 2 (1,1) 3 (0,2) 4 (1,2) 5 (2,2) 6 (0,3) 7 (1,3) 8 (2,3) 9 (3,3)
  const int q = (col < 1) ? 0 : ((col < 3) ? 1 : ((col < 6) ? 2 : 3));
  const int p = col - ((q * (q + 1)) >> 1);
  const double* vp = v[p];
  const double* vq = v[q];
  const double dp[3] = {vp[3 * pa] - vp[3 * pb], vp[3 * pa + 1] - vp[3 * pb + 1], vp[3 * pa + 2] - vp[3 * pb + 2]};
  const double dq[3] = {vq[3 * pa] - vq[3 * pb], vq[3 * pa + 1] - vq[3 * pb + 1], vq[3 * pa + 2] - vq[3 * pb + 2]};
  const double d = epnp_dot3(dp, dq);
  return (p == q) ? d : 2.0 * d;
}

HM_HD inline double epnp_rho_entry(const double (*cws)[3], int i) {
  const int pa = (i < 3) ? 0 : ((i < 5) ? 1 : 2);
  const int pb = (i < 3) ? i + 1 : ((i < 5) ? i - 1 : 3);
  return epnp_dist2(cws[pa], cws[pb]);
}

// columns of L that approximation ap (0: N=4 linearised [B11 B12 B13 B14], 1: N=2 [B11 B12 B22],
// 2: N=3 [B11 B12 B22 B13 B23]) solves for
HM_HD inline int epnp_approx_cols(int ap) { return ap == 0 ? 4 : (ap == 1 ? 3 : 5); }
HM_HD inline int epnp_approx_col(int ap, int c) { return ap == 0 ? (c == 0 ? 0 : (c == 1 ? 1 : (c == 2 ? 3 : 6))) : c; }

// betas from the least-squares solution b of approximation ap
HM_HD inline void epnp_betas_from_ls(int ap, const double* b, double* betas) {
  if (ap == 0) {
    if (b[0] < 0) { betas[0] = sqrt(-b[0]); betas[1] = -b[1] / betas[0]; betas[2] = -b[2] / betas[0]; betas[3] = -b[3] / betas[0]; }
    else { betas[0] = sqrt(b[0]); betas[1] = b[1] / betas[0]; betas[2] = b[2] / betas[0]; betas[3] = b[3] / betas[0]; }
    return;
  }
  if (b[0] < 0) { betas[0] = sqrt(-b[0]); betas[1] = (b[2] < 0) ? sqrt(-b[2]) : 0.0; }
  else { betas[0] = sqrt(b[0]); betas[1] = (b[2] > 0) ? sqrt(b[2]) : 0.0; }
  if (b[1] < 0) betas[0] = -betas[0];
  betas[2] = (ap == 2) ? b[3] / betas[0] : 0.0;
  betas[3] = 0.0;
}

// The solver's Householder QR least squares (6 x 4), including its pivot scan, which starts at row k and stops one
// row early, and its early return on a zero column (x keeps its previous value).
HM_HD inline void epnp_qr_solve_6x4(double* pA, double* pb, double* pX) {
  const int nr = 6, nc = 4;
  double A1[4], A2[4];
  for (int k = 0; k < nc; ++k) {
    double eta = fabs(pA[k * nc + k]);
    for (int i = k + 1; i < nr; ++i) {
      const double elt = fabs(pA[(i - 1) * nc + k]);
      if (eta < elt) eta = elt;
    }
    if (eta == 0) return;
    const double inv_eta = 1. / eta;
    double sum2 = 0.0;
    for (int i = k; i < nr; ++i) {
      pA[i * nc + k] *= inv_eta;
      sum2 += pA[i * nc + k] * pA[i * nc + k];
    }
    double sigma = sqrt(sum2);
    if (pA[k * nc + k] < 0) sigma = -sigma;
    pA[k * nc + k] += sigma;
    A1[k] = sigma * pA[k * nc + k];
    A2[k] = -eta * sigma;
    for (int j = k + 1; j < nc; ++j) {
      double sum = 0;
      for (int i = k; i < nr; ++i) sum += pA[i * nc + k] * pA[i * nc + j];
      const double tau = sum / A1[k];
      for (int i = k; i < nr; ++i) pA[i * nc + j] -= tau * pA[i * nc + k];
    }
  }
  for (int j = 0; j < nc; ++j) {
    double tau = 0;
    for (int i = j; i < nr; ++i) tau += pA[i * nc + j] * pb[i];
    tau /= A1[j];
    for (int i = j; i < nr; ++i) pb[i] -= tau * pA[i * nc + j];
  }
  pX[nc - 1] = pb[nc - 1] / A2[nc - 1];
  for (int i = nc - 2; i >= 0; --i) {
    double sum = 0;
    for (int j = i + 1; j < nc; ++j) sum += pA[i * nc + j] * pX[j];
    pX[i] = (pb[i] - sum) / A2[i];
  }
}

HM_HD inline void epnp_gauss_newton(const double* L, const double* rho, double* b) {
  double x[4] = {0, 0, 0, 0};
  for (int it = 0; it < 5; ++it) {
    double A[24], r[6];
    for (int i = 0; i < 6; ++i) {
      const double* l = L + 10 * i;
      A[4 * i + 0] = 2 * l[0] * b[0] + l[1] * b[1] + l[3] * b[2] + l[6] * b[3];
      A[4 * i + 1] = l[1] * b[0] + 2 * l[2] * b[1] + l[4] * b[2] + l[7] * b[3];
      A[4 * i + 2] = l[3] * b[0] + l[4] * b[1] + 2 * l[5] * b[2] + l[8] * b[3];
      A[4 * i + 3] = l[6] * b[0] + l[7] * b[1] + l[8] * b[2] + 2 * l[9] * b[3];
      r[i] = rho[i] - (l[0] * b[0] * b[0] + l[1] * b[0] * b[1] + l[2] * b[1] * b[1] + l[3] * b[0] * b[2] +
                       l[4] * b[1] * b[2] + l[5] * b[2] * b[2] + l[6] * b[0] * b[3] + l[7] * b[1] * b[3] +
                       l[8] * b[2] * b[3] + l[9] * b[3] * b[3]);
    }
    epnp_qr_solve_6x4(A, r, x);
    for (int k = 0; k < 4; ++k) b[k] += x[k];
  }
}

// Pose from betas: control points in the camera frame, sign fix (first point in front of the camera), absolute
// orientation by SVD, mean reprojection error.  alphas (n,4), pw (n,3), us (n,2); pcs: 3n doubles of scratch.
HM_HD inline double epnp_pose_from_betas(const double* const v[4], const double* betas, const double* alphas,
                                         const double* pw, const double* us, int n, const EpnpCam& cam,
                                         double* pcs, double* R, double* t) {
  double ccs[4][3];
  for (int i = 0; i < 4; ++i) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      for (int k = 0; k < 3; ++k) ccs[j][k] += betas[i] * v[i][3 * j + k];
  for (int i = 0; i < n; ++i) {
    const double* a = alphas + 4 * i;
    for (int j = 0; j < 3; ++j) pcs[3 * i + j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
  }
  if (pcs[2] < 0.0)
    for (int i = 0; i < 3 * n; ++i) pcs[i] = -pcs[i];
  double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 3; ++j) {
      pc0[j] += pcs[3 * i + j];
      pw0[j] += pw[3 * i + j];
    }
  for (int j = 0; j < 3; ++j) { pc0[j] /= n; pw0[j] /= n; }
  double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 3; ++j) {
      abt[3 * j] += (pcs[3 * i + j] - pc0[j]) * (pw[3 * i] - pw0[0]);
      abt[3 * j + 1] += (pcs[3 * i + j] - pc0[j]) * (pw[3 * i + 1] - pw0[1]);
      abt[3 * j + 2] += (pcs[3 * i + j] - pc0[j]) * (pw[3 * i + 2] - pw0[2]);
    }
  double U[9], W[3], Vt[9];
  svd_square<3>(abt, U, W, Vt);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = U[3 * i] * Vt[j] + U[3 * i + 1] * Vt[3 + j] + U[3 * i + 2] * Vt[6 + j];
  const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] - R[1] * R[3] * R[8] -
                     R[0] * R[5] * R[7];
  if (det < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
  for (int k = 0; k < 3; ++k) t[k] = pc0[k] - epnp_dot3(R + 3 * k, pw0);
  double sum2 = 0.0;
  for (int i = 0; i < n; ++i) {
    const double* p = pw + 3 * i;
    const double Xc = epnp_dot3(R, p) + t[0], Yc = epnp_dot3(R + 3, p) + t[1];
    const double inv_Zc = 1.0 / (epnp_dot3(R + 6, p) + t[2]);
    const double ue = cam.uc + cam.fu * Xc * inv_Zc, ve = cam.vc + cam.fv * Yc * inv_Zc;
    const double u = us[2 * i], vv = us[2 * i + 1];
    sum2 += sqrt((u - ue) * (u - ue) + (vv - ve) * (vv - ve));
  }
  return sum2 / n;
}

// ---- stage 3: one of the three beta initialisations, Gauss-Newton, pose, mean reprojection error (serial form)
HM_HD inline double epnp_candidate(int ap, const double* L, const double* rho, const double* const v[4],
                                   const double* alphas, const double* pw, const double* us, int n,
                                   const EpnpCam& cam, double* pcs, double* R, double* t) {
  double betas[4] = {0, 0, 0, 0}, b[5], w[5], vt[25], At[30];
  const int nc = epnp_approx_cols(ap);
  for (int c = 0; c < nc; ++c)
    for (int i = 0; i < 6; ++i) At[c * 6 + i] = L[10 * i + epnp_approx_col(ap, c)];
  if (ap == 0) { jacobi_svd<6, 4>(At, w, vt); cv_svd_backsubst<6, 4>(w, At, vt, rho, b); }
  else if (ap == 1) { jacobi_svd<6, 3>(At, w, vt); cv_svd_backsubst<6, 3>(w, At, vt, rho, b); }
  else { jacobi_svd<6, 5>(At, w, vt); cv_svd_backsubst<6, 5>(w, At, vt, rho, b); }
  epnp_betas_from_ls(ap, b, betas);
  epnp_gauss_newton(L, rho, betas);
  return epnp_pose_from_betas(v, betas, alphas, pw, us, n, cam, pcs, R, t);
}

// OpenCV's selection among the three candidates: N=1; if e2 < e1 N=2; if e3 < e_N N=3 (NaN never wins)
HM_HD inline int epnp_pick(const double* errs) {
  int N = 0;
  if (errs[1] < errs[0]) N = 1;
  if (errs[2] < errs[N]) N = 2;
  return N;
}

// Serial solver (host utility sfm_epnp; the batched kernel in pnp_epnp.cu runs the same operations spread over a
// CTA).  pw (n,3) object points, us (n,2) pixel coordinates (already passed through the float32 normalise /
// de-normalise round trip OpenCV applies), work: 7*n doubles of scratch.
HM_HD inline void epnp_solve(const double* pw, const double* us, int n, const EpnpCam& cam, double* work,
                             double* R, double* t) {
  double* alphas = work;
  double* pcs = work + 4 * n;
  double cws[4][3], At[144], w12[12], Vt[144], L[60], rho[6];
  epnp_control_alphas(pw, n, alphas, cws);
  for (int r = 0; r < 12; ++r)
    for (int c = r; c < 12; ++c) {
      const double s0 = epnp_MtM_entry(alphas, us, n, cam, r, c);
      At[12 * r + c] = s0;
      At[12 * c + r] = s0;
    }
  jacobi_svd<12, 12>(At, w12, Vt);       // rows of At: left singular vectors, singular values descending
  const double* v[4] = {At + 12 * 11, At + 12 * 10, At + 12 * 9, At + 12 * 8};   // v[0] = smallest
  for (int i = 0; i < 6; ++i) {
    for (int c = 0; c < 10; ++c) L[10 * i + c] = epnp_L_entry(v, i, c);
    rho[i] = epnp_rho_entry(cws, i);
  }
  double Rs[3][9], ts[3][3], errs[3];
  for (int ap = 0; ap < 3; ++ap) errs[ap] = epnp_candidate(ap, L, rho, v, alphas, pw, us, n, cam, pcs, Rs[ap], ts[ap]);
  const int N = epnp_pick(errs);
  for (int k = 0; k < 9; ++k) R[k] = Rs[N][k];
  for (int k = 0; k < 3; ++k) t[k] = ts[N][k];
}

// OpenCV hands EPnP the image points after cv::undistortPoints (zero distortion: x_n = (u-cx)/fx,
// stored in the input dtype, float32) and the solver maps them back to pixels.
HM_HD inline void epnp_roundtrip_pixel(float u, float v, const EpnpCam& cam, double* out) {
  float xn = (float)(((double)u - cam.uc) * (1.0 / cam.fu));
  float yn = (float)(((double)v - cam.vc) * (1.0 / cam.fv));
  out[0] = (double)xn * cam.fu + cam.uc;
  out[1] = (double)yn * cam.fv + cam.vc;
}

}  // namespace hm
