// epnp.h — EPnP minimal solver (Lepetit, Moreno-Noguer, Fua, "EPnP: An Accurate O(n) Solution to
// the PnP Problem", IJCV 2009) in float64, __host__ __device__: the 5-point kernel that
// cv2.solvePnPRansac runs per RANSAC iteration (reference call site sfm.py:67, test.py:319).
//
// The steps are the published ones, in the conventions OpenCV's solver uses (so the three
// candidate poses are the same three candidates): control points = centroid + PCA axes scaled by
// sqrt(lambda/n); barycentric alphas; M^T M (12x12) null-space basis v0..v3 (smallest first);
// L (6x10) and rho; beta initialisations "N=4 linearised", "N=2", "N=3"; five Gauss-Newton
// iterations each; absolute orientation by SVD; the pose with the smallest mean reprojection
// error wins.  Eigen-decompositions are cyclic Jacobi (no LAPACK on a GPU), so the basis chosen
// inside the degenerate null space of M (rank <= 10 for 5 points) is this library's, not
// LAPACK's — see DESIGN.md "PnP parity".
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

// Least squares min |A x - b| for an M x N system (M >= N) by Householder QR; A (row-major) and b
// are destroyed.  Rank-deficient columns produce non-finite x, which the caller's "smallest
// reprojection error" selection then discards (NaN never compares smaller).
template <int M, int N>
HM_HD inline void ls_solve(double* A, double* b, double* x) {
  HM_UNROLL
  for (int k = 0; k < N; ++k) {
    double nrm = 0.0;
    HM_UNROLL
    for (int i = k; i < M; ++i) nrm += A[i * N + k] * A[i * N + k];
    nrm = sqrt(nrm);
    double alpha = A[k * N + k] > 0.0 ? -nrm : nrm;
    double vk = A[k * N + k] - alpha;
    double vnorm2 = vk * vk;
    HM_UNROLL
    for (int i = k + 1; i < M; ++i) vnorm2 += A[i * N + k] * A[i * N + k];
    if (vnorm2 > 0.0) {
      const double two_inv = 2.0 / vnorm2;     // one division per reflector instead of one per column
      HM_UNROLL
      for (int j = k + 1; j < N; ++j) {
        double dot = vk * A[k * N + j];
        HM_UNROLL
        for (int i = k + 1; i < M; ++i) dot += A[i * N + k] * A[i * N + j];
        double f = dot * two_inv;
        A[k * N + j] -= f * vk;
        HM_UNROLL
        for (int i = k + 1; i < M; ++i) A[i * N + j] -= f * A[i * N + k];
      }
      double dot = vk * b[k];
      HM_UNROLL
      for (int i = k + 1; i < M; ++i) dot += A[i * N + k] * b[i];
      double f = dot * two_inv;
      b[k] -= f * vk;
      HM_UNROLL
      for (int i = k + 1; i < M; ++i) b[i] -= f * A[i * N + k];
    }
    A[k * N + k] = alpha;
  }
  HM_UNROLL
  for (int k = N - 1; k >= 0; --k) {
    double s = b[k];
    HM_UNROLL
    for (int j = k + 1; j < N; ++j) s -= A[k * N + j] * x[j];
    x[k] = s / A[k * N + k];
  }
}

struct EpnpCam { double fu, fv, uc, vc; };

HM_HD inline double epnp_dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Pose from betas: control points in the camera frame, sign fix, absolute orientation, mean
// reprojection error.  alphas (n,4), pw (n,3), us (n,2).
HM_HD inline double epnp_pose_from_betas(const double* const v[4], const double* betas, const double* alphas,
                                         const double* pw, const double* us, int n, const EpnpCam& cam,
                                         double* R, double* t) {
  double ccs[4][3];
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 3; ++k)
      ccs[i][k] = betas[0] * v[0][3 * i + k] + betas[1] * v[1][3 * i + k] + betas[2] * v[2][3 * i + k] +
                  betas[3] * v[3][3 * i + k];
  // sign: the first point must be in front of the camera
  {
    const double* a = alphas;
    double z0 = a[0] * ccs[0][2] + a[1] * ccs[1][2] + a[2] * ccs[2][2] + a[3] * ccs[3][2];
    if (z0 < 0.0)
      for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 3; ++k) ccs[i][k] = -ccs[i][k];
  }
  double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i) {
    const double* a = alphas + 4 * i;
    for (int k = 0; k < 3; ++k) {
      pc0[k] += a[0] * ccs[0][k] + a[1] * ccs[1][k] + a[2] * ccs[2][k] + a[3] * ccs[3][k];
      pw0[k] += pw[3 * i + k];
    }
  }
  for (int k = 0; k < 3; ++k) { pc0[k] /= n; pw0[k] /= n; }
  double ABt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    const double* a = alphas + 4 * i;
    double pc[3];
    for (int k = 0; k < 3; ++k) pc[k] = a[0] * ccs[0][k] + a[1] * ccs[1][k] + a[2] * ccs[2][k] + a[3] * ccs[3][k];
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) ABt[3 * j + k] += (pc[j] - pc0[j]) * (pw[3 * i + k] - pw0[k]);
  }
  double U[9], W[3], Vt[9];
  svd_square<3>(ABt, U, W, Vt);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = U[3 * i] * Vt[j] + U[3 * i + 1] * Vt[3 + j] + U[3 * i + 2] * Vt[6 + j];
  double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
  if (det < 0.0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
  for (int k = 0; k < 3; ++k) t[k] = pc0[k] - epnp_dot3(R + 3 * k, pw0);
  double sum = 0.0;
  for (int i = 0; i < n; ++i) {
    const double* p = pw + 3 * i;
    double Xc = epnp_dot3(R, p) + t[0], Yc = epnp_dot3(R + 3, p) + t[1];
    double iz = 1.0 / (epnp_dot3(R + 6, p) + t[2]);
    double du = us[2 * i] - (cam.uc + cam.fu * Xc * iz), dv = us[2 * i + 1] - (cam.vc + cam.fv * Yc * iz);
    sum += sqrt(du * du + dv * dv);
  }
  return sum / n;
}

HM_HD inline void epnp_gauss_newton(const double* L, const double* rho, double* b) {
  for (int it = 0; it < 5; ++it) {
    double A[24], r[6], dx[4];
    HM_UNROLL
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
    ls_solve<6, 4>(A, r, dx);
    HM_UNROLL
    for (int k = 0; k < 4; ++k) b[k] += dx[k];
  }
}

HM_HD inline void epnp_MtM(const double* alphas, const double* us, int n, const EpnpCam& cam, double* MtM);

// ---- stage 1a: control points and barycentric coordinates (alphas: n x 4)
HM_HD inline void epnp_control_alphas(const double* pw, int n, double* alphas, double (*cws)[3], int pca_sweeps = 60);

// ---- stage 1: control points, barycentric coordinates, M^T M.  alphas: n x 4 scratch.
HM_HD inline void epnp_build(const double* pw, const double* us, int n, const EpnpCam& cam, double* alphas,
                             double (*cws)[3], double* MtM) {
  epnp_control_alphas(pw, n, alphas, cws);
  epnp_MtM(alphas, us, n, cam, MtM);
}

// pca_sweeps bounds the Jacobi sweeps of the 3x3 covariance: the control points only have to be a well-conditioned
// affine frame that the barycentric coordinates are computed from consistently — they do not have to be the exact
// principal axes (4 sweeps bring the off-diagonal to ~1e-12; the batched kernel uses that).
HM_HD inline void epnp_control_alphas(const double* pw, int n, double* alphas, double (*cws)[3], int pca_sweeps) {
  for (int k = 0; k < 3; ++k) cws[0][k] = 0.0;
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) cws[0][k] += pw[3 * i + k];
  for (int k = 0; k < 3; ++k) cws[0][k] /= n;
  double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    double d[3] = {pw[3 * i] - cws[0][0], pw[3 * i + 1] - cws[0][1], pw[3 * i + 2] - cws[0][2]};
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) C[3 * j + k] += d[j] * d[k];
  }
  double dc[3], uct[9];
  eig_sym<3>(C, dc, uct, pca_sweeps);   // ascending; OpenCV's SVD order is descending
  for (int i = 1; i < 4; ++i) {
    int e = 3 - i;
    double lam = dc[e] > 0.0 ? dc[e] : 0.0;
    double k = sqrt(lam / n);
    for (int j = 0; j < 3; ++j) cws[i][j] = cws[0][j] + k * uct[3 * e + j];
  }
  // barycentric coordinates: inverse of CC = [c1-c0 | c2-c0 | c3-c0]
  double cc[9], ci[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
  {
    double d = cc[0] * (cc[4] * cc[8] - cc[5] * cc[7]) - cc[1] * (cc[3] * cc[8] - cc[5] * cc[6]) +
               cc[2] * (cc[3] * cc[7] - cc[4] * cc[6]);
    double id = 1.0 / d;
    ci[0] = (cc[4] * cc[8] - cc[5] * cc[7]) * id; ci[1] = (cc[2] * cc[7] - cc[1] * cc[8]) * id; ci[2] = (cc[1] * cc[5] - cc[2] * cc[4]) * id;
    ci[3] = (cc[5] * cc[6] - cc[3] * cc[8]) * id; ci[4] = (cc[0] * cc[8] - cc[2] * cc[6]) * id; ci[5] = (cc[2] * cc[3] - cc[0] * cc[5]) * id;
    ci[6] = (cc[3] * cc[7] - cc[4] * cc[6]) * id; ci[7] = (cc[1] * cc[6] - cc[0] * cc[7]) * id; ci[8] = (cc[0] * cc[4] - cc[1] * cc[3]) * id;
  }
  for (int i = 0; i < n; ++i) {
    double d[3] = {pw[3 * i] - cws[0][0], pw[3 * i + 1] - cws[0][1], pw[3 * i + 2] - cws[0][2]};
    double* a = alphas + 4 * i;
    for (int j = 0; j < 3; ++j) a[1 + j] = ci[3 * j] * d[0] + ci[3 * j + 1] * d[1] + ci[3 * j + 2] * d[2];
    a[0] = 1.0 - a[1] - a[2] - a[3];
  }
}

// M^T M (12x12, symmetric) from the barycentric coordinates: two rows of M per correspondence.
HM_HD inline void epnp_MtM(const double* alphas, const double* us, int n, const EpnpCam& cam, double* MtM) {
  for (int i = 0; i < 144; ++i) MtM[i] = 0.0;
  for (int i = 0; i < n; ++i) {
    const double* a = alphas + 4 * i;
    double m1[12], m2[12];
    for (int j = 0; j < 4; ++j) {
      m1[3 * j] = a[j] * cam.fu; m1[3 * j + 1] = 0.0;            m1[3 * j + 2] = a[j] * (cam.uc - us[2 * i]);
      m2[3 * j] = 0.0;           m2[3 * j + 1] = a[j] * cam.fv;  m2[3 * j + 2] = a[j] * (cam.vc - us[2 * i + 1]);
    }
    for (int r = 0; r < 12; ++r)
      for (int c = r; c < 12; ++c) MtM[12 * r + c] += m1[r] * m1[c] + m2[r] * m2[c];
  }
  for (int r = 0; r < 12; ++r)
    for (int c = 0; c < r; ++c) MtM[12 * r + c] = MtM[12 * c + r];
}

// ---- stage 2 (after the eigen-decomposition): L (6x10) and rho from the four null-space vectors
HM_HD inline void epnp_L_rho(const double* const v[4], const double (*cws)[3], double* L, double* rho) {
  double dv[4][6][3];
  for (int i = 0; i < 4; ++i) {
    int a = 0, b = 1;
    for (int j = 0; j < 6; ++j) {
      for (int k = 0; k < 3; ++k) dv[i][j][k] = v[i][3 * a + k] - v[i][3 * b + k];
      if (++b > 3) { ++a; b = a + 1; }
    }
  }
  for (int i = 0; i < 6; ++i) {
    double* row = L + 10 * i;
    row[0] = epnp_dot3(dv[0][i], dv[0][i]);
    row[1] = 2.0 * epnp_dot3(dv[0][i], dv[1][i]);
    row[2] = epnp_dot3(dv[1][i], dv[1][i]);
    row[3] = 2.0 * epnp_dot3(dv[0][i], dv[2][i]);
    row[4] = 2.0 * epnp_dot3(dv[1][i], dv[2][i]);
    row[5] = epnp_dot3(dv[2][i], dv[2][i]);
    row[6] = 2.0 * epnp_dot3(dv[0][i], dv[3][i]);
    row[7] = 2.0 * epnp_dot3(dv[1][i], dv[3][i]);
    row[8] = 2.0 * epnp_dot3(dv[2][i], dv[3][i]);
    row[9] = epnp_dot3(dv[3][i], dv[3][i]);
  }
  int a = 0, b = 1;
  for (int j = 0; j < 6; ++j) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) { double d = cws[a][k] - cws[b][k]; s += d * d; }
    rho[j] = s;
    if (++b > 3) { ++a; b = a + 1; }
  }
}

// ---- stage 3: one of the three beta initialisations (ap = 0: N=4 linearised, 1: N=2, 2: N=3),
// Gauss-Newton, pose, mean reprojection error.
HM_HD inline double epnp_candidate(int ap, const double* L, const double* rho, const double* const v[4],
                                   const double* alphas, const double* pw, const double* us, int n,
                                   const EpnpCam& cam, double* R, double* t) {
  double betas[4] = {0, 0, 0, 0};
  if (ap == 0) {          // [B11 B12 B13 B14]
    double A[24], r[6], b4[4];
    for (int i = 0; i < 6; ++i) {
      A[4 * i] = L[10 * i]; A[4 * i + 1] = L[10 * i + 1]; A[4 * i + 2] = L[10 * i + 3]; A[4 * i + 3] = L[10 * i + 6];
      r[i] = rho[i];
    }
    ls_solve<6, 4>(A, r, b4);
    if (b4[0] < 0) { betas[0] = sqrt(-b4[0]); betas[1] = -b4[1] / betas[0]; betas[2] = -b4[2] / betas[0]; betas[3] = -b4[3] / betas[0]; }
    else { betas[0] = sqrt(b4[0]); betas[1] = b4[1] / betas[0]; betas[2] = b4[2] / betas[0]; betas[3] = b4[3] / betas[0]; }
  } else if (ap == 1) {   // [B11 B12 B22]
    double A[18], r[6], b3[3];
    for (int i = 0; i < 6; ++i) {
      A[3 * i] = L[10 * i]; A[3 * i + 1] = L[10 * i + 1]; A[3 * i + 2] = L[10 * i + 2];
      r[i] = rho[i];
    }
    ls_solve<6, 3>(A, r, b3);
    if (b3[0] < 0) { betas[0] = sqrt(-b3[0]); betas[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0; }
    else { betas[0] = sqrt(b3[0]); betas[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0; }
    if (b3[1] < 0) betas[0] = -betas[0];
  } else {                // [B11 B12 B22 B13 B23]
    double A[30], r[6], b5[5];
    for (int i = 0; i < 6; ++i) {
      for (int k = 0; k < 5; ++k) A[5 * i + k] = L[10 * i + k];
      r[i] = rho[i];
    }
    ls_solve<6, 5>(A, r, b5);
    if (b5[0] < 0) { betas[0] = sqrt(-b5[0]); betas[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0; }
    else { betas[0] = sqrt(b5[0]); betas[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0; }
    if (b5[1] < 0) betas[0] = -betas[0];
    betas[2] = b5[3] / betas[0];
  }
  epnp_gauss_newton(L, rho, betas);
  return epnp_pose_from_betas(v, betas, alphas, pw, us, n, cam, R, t);
}

// OpenCV's selection among the three candidates: N=1; if e2 < e1 N=2; if e3 < e_N N=3 (NaN never wins)
HM_HD inline int epnp_pick(const double* errs) {
  int N = 0;
  if (errs[1] < errs[0]) N = 1;
  if (errs[2] < errs[N]) N = 2;
  return N;
}

// Serial solver (host utility sfm_epnp; also usable in a single device thread).
// pw (n,3) object points, us (n,2) pixel coordinates (already passed through the float32
// normalise / de-normalise round trip OpenCV applies), work: 4*n doubles of scratch.
HM_HD inline void epnp_solve(const double* pw, const double* us, int n, const EpnpCam& cam, double* work,
                             double* R, double* t) {
  double* alphas = work;
  double cws[4][3], MtM[144], w12[12], V12[144], L[60], rho[6];
  epnp_build(pw, us, n, cam, alphas, cws, MtM);
  eig_sym<12>(MtM, w12, V12);
  const double* v[4] = {V12, V12 + 12, V12 + 24, V12 + 36};   // v[0] = smallest eigenvalue
  epnp_L_rho(v, cws, L, rho);
  double Rs[3][9], ts[3][3], errs[3];
  for (int ap = 0; ap < 3; ++ap) errs[ap] = epnp_candidate(ap, L, rho, v, alphas, pw, us, n, cam, Rs[ap], ts[ap]);
  int N = epnp_pick(errs);
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
