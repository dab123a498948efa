/* cv_epnp.c — operation-for-operation C restatement of the arithmetic OpenCV 4.x runs for
 * cv2.solvePnP(flags=SOLVEPNP_EPNP), the 5-point minimal solver inside cv2.solvePnPRansac
 * (reference call site /root/reference/sfm.py:67, test.py:319).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): built by oracle/Makefile into
 * oracle/_build/libcvoracle.so, loaded by oracle/cv_exact.py, used by tests/ and the bench's parity leg.
 *
 * Why it can be exact: the algorithm lives in OpenCV (un-vendored; the image has cv2 4.13.0), and every
 * decomposition on this path is SMALL — OpenCV only hands SVDs to LAPACK from 25 rows up, below that it runs its
 * own one-sided Jacobi (modules/core/src/lapack.cpp JacobiSVDImpl_), which is plain IEEE double arithmetic in a
 * fixed order, compiled for the SSE3 baseline (no FMA).  Restating that order gives the same bits, including
 * the basis OpenCV ends up with inside the 2-dimensional null space of the 5-point M^T M.  Pinned by
 * tests/test_oracle.py against in-process cv2: cv2.SVDecomp, cv2.mulTransposed, cv2.invert(DECOMP_SVD),
 * cv2.solve(DECOMP_SVD) and cv2.solvePnP(EPNP) itself, bit for bit.
 *
 * Published algorithm: Lepetit, Moreno-Noguer, Fua, "EPnP: An Accurate O(n) Solution to the PnP Problem",
 * IJCV 2009, in the form of OpenCV's modules/calib3d/src/epnp.cpp.
 *
 * Build with -ffp-contract=off (no fused multiply-add anywhere).
 */
#include <float.h>
#include <math.h>
#include <string.h>

/* OpenCV's own hypot (lapack.cpp), NOT libm's */
static double cv_hypot(double a, double b) {
  a = fabs(a);
  b = fabs(b);
  if (a > b) {
    b /= a;
    return a * sqrt(1 + b * b);
  }
  if (b > 0) {
    a /= b;
    return b * sqrt(1 + a * a);
  }
  return 0;
}

/* JacobiSVDImpl_<double>: At = n rows of length m (the COLUMNS of the decomposed matrix), row stride m.
 * On return: W[n] descending, At rows = left singular vectors, Vt (n x n) rows = right singular vectors.
 * Returns the number of sweeps that rotated something. */
int cvo_jacobi_svd(double* At, double* Wout, double* Vt, int m, int n) {
  double W[32];
  const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
  int i, j, k, iter, max_iter = m > 30 ? m : 30, sweeps = 0;
  for (i = 0; i < n; i++) {
    double sd = 0;
    for (k = 0; k < m; k++) {
      double t = At[i * m + k];
      sd += t * t;
    }
    W[i] = sd;
    for (k = 0; k < n; k++) Vt[i * n + k] = 0;
    Vt[i * n + i] = 1;
  }
  for (iter = 0; iter < max_iter; iter++) {
    int changed = 0;
    for (i = 0; i < n - 1; i++)
      for (j = i + 1; j < n; j++) {
        double *Ai = At + i * m, *Aj = At + j * m;
        double a = W[i], p = 0, b = W[j];
        for (k = 0; k < m; k++) p += Ai[k] * Aj[k];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2;
        double beta = a - b, gamma = cv_hypot(p, beta), c, s;
        if (beta < 0) {
          double delta = (gamma - beta) * 0.5;
          s = sqrt(delta / gamma);
          c = p / (gamma * s * 2);
        } else {
          c = sqrt((gamma + beta) / (gamma * 2));
          s = p / (gamma * c * 2);
        }
        a = b = 0;
        for (k = 0; k < m; k++) {
          double t0 = c * Ai[k] + s * Aj[k];
          double t1 = -s * Ai[k] + c * Aj[k];
          Ai[k] = t0;
          Aj[k] = t1;
          a += t0 * t0;
          b += t1 * t1;
        }
        W[i] = a;
        W[j] = b;
        changed = 1;
        double *Vi = Vt + i * n, *Vj = Vt + j * n;
        for (k = 0; k < n; k++) {
          double t0 = c * Vi[k] + s * Vj[k];
          double t1 = -s * Vi[k] + c * Vj[k];
          Vi[k] = t0;
          Vj[k] = t1;
        }
      }
    if (!changed) break;
    ++sweeps;
  }
  for (i = 0; i < n; i++) {
    double sd = 0;
    for (k = 0; k < m; k++) {
      double t = At[i * m + k];
      sd += t * t;
    }
    W[i] = sqrt(sd);
  }
  for (i = 0; i < n - 1; i++) {
    j = i;
    for (k = i + 1; k < n; k++)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      double t = W[i];
      W[i] = W[j];
      W[j] = t;
      for (k = 0; k < m; k++) {
        t = At[i * m + k];
        At[i * m + k] = At[j * m + k];
        At[j * m + k] = t;
      }
      for (k = 0; k < n; k++) {
        t = Vt[i * n + k];
        Vt[i * n + k] = Vt[j * n + k];
        Vt[j * n + k] = t;
      }
    }
  }
  for (i = 0; i < n; i++) {
    Wout[i] = W[i];
    /* OpenCV regenerates a left vector from its RNG when sd <= DBL_MIN; never met on this path (the rows of a
     * rank-deficient M^T M end as rounding noise ~1e-16 |A|, far above DBL_MIN) — flagged by a zero vector */
    double sd = W[i];
    double s = sd > minval ? 1 / sd : 0.;
    for (k = 0; k < m; k++) At[i * m + k] *= s;
  }
  return sweeps;
}

/* cv::SVD::compute for a square or tall matrix A (m x n, row-major, m >= n): w[n], U (m x n, row-major),
 * Vt (n x n). */
static void cv_svd(const double* A, int m, int n, double* w, double* U, double* Vt) {
  double At[24 * 24];
  for (int i = 0; i < n; i++)
    for (int k = 0; k < m; k++) At[i * m + k] = A[k * n + i];
  cvo_jacobi_svd(At, w, Vt, m, n);
  for (int i = 0; i < n; i++)
    for (int k = 0; k < m; k++) U[k * n + i] = At[i * m + k];
}

void cvo_svd(const double* A, int m, int n, double* w, double* U, double* Vt) { cv_svd(A, m, n, w, U, Vt); }

/* cv::mulTransposed(src, dst, aTa = true) for a (rows x cols) double matrix below the GEMM threshold:
 * MulTransposedR — upper triangle by ascending-row sums, then completeSymm. */
void cvo_mul_transposed(const double* src, int rows, int cols, double* dst) {
  for (int i = 0; i < cols; i++)
    for (int j = i; j < cols; j++) {
      double s0 = 0;
      for (int k = 0; k < rows; k++) s0 += src[k * cols + i] * src[k * cols + j];
      dst[i * cols + j] = s0 * 1.0;
    }
  for (int i = 0; i < cols; i++)
    for (int j = 0; j < i; j++) dst[i * cols + j] = dst[j * cols + i];
}

/* cv::SVD::backSubst == SVBkSbImpl_<double>(m, n, w, u (m x nm), vt (nm x n), b (m x nb) or NULL -> identity) */
static void cv_svbksb(int m, int n, const double* w, const double* u, const double* vt, const double* b, int nb,
                      double* x) {
  int nm = m < n ? m : n;
  double threshold = 0, buffer[32];
  if (!b) nb = m;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < nb; j++) x[i * nb + j] = 0;
  for (int i = 0; i < nm; i++) threshold += w[i];
  threshold *= DBL_EPSILON * 2;
  for (int i = 0; i < nm; i++) {
    double wi = w[i];
    if (fabs(wi) <= threshold) continue;
    wi = 1 / wi;
    if (nb == 1) {
      double s = 0;
      if (b)
        for (int j = 0; j < m; j++) s += u[j * nm + i] * b[j];
      else
        s = u[i];
      s *= wi;
      for (int j = 0; j < n; j++) x[j] = x[j] + s * vt[i * n + j];
    } else {
      if (b) {
        for (int j = 0; j < nb; j++) buffer[j] = 0;
        for (int r = 0; r < m; r++) /* MatrAXPY(m, nb, b, ldb, u, udelta1, buffer, 0) */
          for (int j = 0; j < nb; j++) buffer[j] = buffer[j] + u[r * nm + i] * b[r * nb + j];
        for (int j = 0; j < nb; j++) buffer[j] *= wi;
      } else {
        for (int j = 0; j < nb; j++) buffer[j] = u[j * nm + i] * wi;
      }
      for (int r = 0; r < n; r++) /* MatrAXPY(n, nb, buffer, 0, v, vdelta1, x, ldx) */
        for (int j = 0; j < nb; j++) x[r * nb + j] = x[r * nb + j] + vt[i * n + r] * buffer[j];
    }
  }
}

/* cv::invert(A, DECOMP_SVD), n x n */
void cvo_invert_svd(const double* A, int n, double* Ainv) {
  double w[24], U[24 * 24], Vt[24 * 24];
  cv_svd(A, n, n, w, U, Vt);
  cv_svbksb(n, n, w, U, Vt, 0, n, Ainv);
}

/* cv::solve(A (m x n), b (m), DECOMP_SVD) */
void cvo_solve_svd(const double* A, int m, int n, const double* b, double* x) {
  double w[24], U[24 * 24], Vt[24 * 24];
  cv_svd(A, m, n, w, U, Vt);
  cv_svbksb(m, n, w, U, Vt, b, 1, x);
}

/* ------------------------------------------------------------------------------------ epnp.cpp */
typedef struct {
  double fu, fv, uc, vc;
  int n;
  const double* pws; /* n x 3 */
  const double* us;  /* n x 2 */
  double alphas[4 * 16], pcs[3 * 16];
  double cws[4][3], ccs[4][3];
} Epnp;

static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double dist2(const double* p1, const double* p2) {
  return (p1[0] - p2[0]) * (p1[0] - p2[0]) + (p1[1] - p2[1]) * (p1[1] - p2[1]) + (p1[2] - p2[2]) * (p1[2] - p2[2]);
}

static void choose_control_points(Epnp* e) {
  const int n = e->n;
  e->cws[0][0] = e->cws[0][1] = e->cws[0][2] = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 3; j++) e->cws[0][j] += e->pws[3 * i + j];
  for (int j = 0; j < 3; j++) e->cws[0][j] /= n;
  double pw0[3 * 16], pw0tpw0[9], dc[3], u[9], vt[9];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 3; j++) pw0[3 * i + j] = e->pws[3 * i + j] - e->cws[0][j];
  cvo_mul_transposed(pw0, n, 3, pw0tpw0);
  cv_svd(pw0tpw0, 3, 3, dc, u, vt); /* UCt = U^T: row i-1 of UCt = column i-1 of U */
  for (int i = 1; i < 4; i++) {
    double k = sqrt(dc[i - 1] / n);
    for (int j = 0; j < 3; j++) e->cws[i][j] = e->cws[0][j] + k * u[3 * j + (i - 1)];
  }
}

static void compute_barycentric_coordinates(Epnp* e) {
  double cc[9], ci[9];
  for (int i = 0; i < 3; i++)
    for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = e->cws[j][i] - e->cws[0][i];
  cvo_invert_svd(cc, 3, ci);
  for (int i = 0; i < e->n; i++) {
    const double* pi = e->pws + 3 * i;
    double* a = e->alphas + 4 * i;
    for (int j = 0; j < 3; j++)
      a[1 + j] = ci[3 * j] * (pi[0] - e->cws[0][0]) + ci[3 * j + 1] * (pi[1] - e->cws[0][1]) +
                 ci[3 * j + 2] * (pi[2] - e->cws[0][2]);
    a[0] = 1.0f - a[1] - a[2] - a[3];
  }
}

static void fill_M(const Epnp* e, double* M, int row, const double* as, double u, double v) {
  double* M1 = M + row * 12;
  double* M2 = M1 + 12;
  for (int i = 0; i < 4; i++) {
    M1[3 * i] = as[i] * e->fu;
    M1[3 * i + 1] = 0.0;
    M1[3 * i + 2] = as[i] * (e->uc - u);
    M2[3 * i] = 0.0;
    M2[3 * i + 1] = as[i] * e->fv;
    M2[3 * i + 2] = as[i] * (e->vc - v);
  }
}

static void compute_L_6x10(const double* ut, double* l) {
  const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
  double dv[4][6][3];
  for (int i = 0; i < 4; i++) {
    int a = 0, b = 1;
    for (int j = 0; j < 6; j++) {
      dv[i][j][0] = v[i][3 * a] - v[i][3 * b];
      dv[i][j][1] = v[i][3 * a + 1] - v[i][3 * b + 1];
      dv[i][j][2] = v[i][3 * a + 2] - v[i][3 * b + 2];
      b++;
      if (b > 3) {
        a++;
        b = a + 1;
      }
    }
  }
  for (int i = 0; i < 6; i++) {
    double* row = l + 10 * i;
    row[0] = dot3(dv[0][i], dv[0][i]);
    row[1] = 2.0f * dot3(dv[0][i], dv[1][i]);
    row[2] = dot3(dv[1][i], dv[1][i]);
    row[3] = 2.0f * dot3(dv[0][i], dv[2][i]);
    row[4] = 2.0f * dot3(dv[1][i], dv[2][i]);
    row[5] = dot3(dv[2][i], dv[2][i]);
    row[6] = 2.0f * dot3(dv[0][i], dv[3][i]);
    row[7] = 2.0f * dot3(dv[1][i], dv[3][i]);
    row[8] = 2.0f * dot3(dv[2][i], dv[3][i]);
    row[9] = dot3(dv[3][i], dv[3][i]);
  }
}

static void compute_rho(const Epnp* e, double* rho) {
  rho[0] = dist2(e->cws[0], e->cws[1]);
  rho[1] = dist2(e->cws[0], e->cws[2]);
  rho[2] = dist2(e->cws[0], e->cws[3]);
  rho[3] = dist2(e->cws[1], e->cws[2]);
  rho[4] = dist2(e->cws[1], e->cws[3]);
  rho[5] = dist2(e->cws[2], e->cws[3]);
}

static void find_betas_approx_1(const double* L, const double* rho, double* betas) {
  double l[24], b4[4];
  for (int i = 0; i < 6; i++) {
    l[4 * i] = L[10 * i];
    l[4 * i + 1] = L[10 * i + 1];
    l[4 * i + 2] = L[10 * i + 3];
    l[4 * i + 3] = L[10 * i + 6];
  }
  cvo_solve_svd(l, 6, 4, rho, b4);
  if (b4[0] < 0) {
    betas[0] = sqrt(-b4[0]);
    betas[1] = -b4[1] / betas[0];
    betas[2] = -b4[2] / betas[0];
    betas[3] = -b4[3] / betas[0];
  } else {
    betas[0] = sqrt(b4[0]);
    betas[1] = b4[1] / betas[0];
    betas[2] = b4[2] / betas[0];
    betas[3] = b4[3] / betas[0];
  }
}

static void find_betas_approx_2(const double* L, const double* rho, double* betas) {
  double l[18], b3[3];
  for (int i = 0; i < 6; i++) {
    l[3 * i] = L[10 * i];
    l[3 * i + 1] = L[10 * i + 1];
    l[3 * i + 2] = L[10 * i + 2];
  }
  cvo_solve_svd(l, 6, 3, rho, b3);
  if (b3[0] < 0) {
    betas[0] = sqrt(-b3[0]);
    betas[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0;
  } else {
    betas[0] = sqrt(b3[0]);
    betas[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0;
  }
  if (b3[1] < 0) betas[0] = -betas[0];
  betas[2] = 0.0;
  betas[3] = 0.0;
}

static void find_betas_approx_3(const double* L, const double* rho, double* betas) {
  double l[30], b5[5];
  for (int i = 0; i < 6; i++)
    for (int k = 0; k < 5; k++) l[5 * i + k] = L[10 * i + k];
  cvo_solve_svd(l, 6, 5, rho, b5);
  if (b5[0] < 0) {
    betas[0] = sqrt(-b5[0]);
    betas[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0;
  } else {
    betas[0] = sqrt(b5[0]);
    betas[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0;
  }
  if (b5[1] < 0) betas[0] = -betas[0];
  betas[2] = b5[3] / betas[0];
  betas[3] = 0.0;
}

static void compute_A_and_b_gauss_newton(const double* l_6x10, const double* rho, const double* betas, double* A,
                                         double* b) {
  for (int i = 0; i < 6; i++) {
    const double* rowL = l_6x10 + i * 10;
    double* rowA = A + i * 4;
    rowA[0] = 2 * rowL[0] * betas[0] + rowL[1] * betas[1] + rowL[3] * betas[2] + rowL[6] * betas[3];
    rowA[1] = rowL[1] * betas[0] + 2 * rowL[2] * betas[1] + rowL[4] * betas[2] + rowL[7] * betas[3];
    rowA[2] = rowL[3] * betas[0] + rowL[4] * betas[1] + 2 * rowL[5] * betas[2] + rowL[8] * betas[3];
    rowA[3] = rowL[6] * betas[0] + rowL[7] * betas[1] + rowL[8] * betas[2] + 2 * rowL[9] * betas[3];
    b[i] = rho[i] - (rowL[0] * betas[0] * betas[0] + rowL[1] * betas[0] * betas[1] + rowL[2] * betas[1] * betas[1] +
                     rowL[3] * betas[0] * betas[2] + rowL[4] * betas[1] * betas[2] + rowL[5] * betas[2] * betas[2] +
                     rowL[6] * betas[0] * betas[3] + rowL[7] * betas[1] * betas[3] + rowL[8] * betas[2] * betas[3] +
                     rowL[9] * betas[3] * betas[3]);
  }
}

/* epnp::qr_solve (Householder), including its pivot scan that starts at row k and stops one row early */
static void qr_solve(double* pA, int nr, int nc, double* pb, double* pX) {
  double A1[8], A2[8];
  double* ppAkk = pA;
  for (int k = 0; k < nc; k++) {
    double *ppAik1 = ppAkk, eta = fabs(*ppAik1);
    for (int i = k + 1; i < nr; i++) {
      double elt = fabs(*ppAik1);
      if (eta < elt) eta = elt;
      ppAik1 += nc;
    }
    if (eta == 0) {
      A1[k] = A2[k] = 0.0;
      return;
    } else {
      double *ppAik2 = ppAkk, sum2 = 0.0, inv_eta = 1. / eta;
      for (int i = k; i < nr; i++) {
        *ppAik2 *= inv_eta;
        sum2 += *ppAik2 * *ppAik2;
        ppAik2 += nc;
      }
      double sigma = sqrt(sum2);
      if (*ppAkk < 0) sigma = -sigma;
      *ppAkk += sigma;
      A1[k] = sigma * *ppAkk;
      A2[k] = -eta * sigma;
      for (int j = k + 1; j < nc; j++) {
        double *ppAik = ppAkk, sum = 0;
        for (int i = k; i < nr; i++) {
          sum += *ppAik * ppAik[j - k];
          ppAik += nc;
        }
        double tau = sum / A1[k];
        ppAik = ppAkk;
        for (int i = k; i < nr; i++) {
          ppAik[j - k] -= tau * *ppAik;
          ppAik += nc;
        }
      }
    }
    ppAkk += nc + 1;
  }
  double* ppAjj = pA;
  for (int j = 0; j < nc; j++) {
    double *ppAij = ppAjj, tau = 0;
    for (int i = j; i < nr; i++) {
      tau += *ppAij * pb[i];
      ppAij += nc;
    }
    tau /= A1[j];
    ppAij = ppAjj;
    for (int i = j; i < nr; i++) {
      pb[i] -= tau * *ppAij;
      ppAij += nc;
    }
    ppAjj += nc + 1;
  }
  pX[nc - 1] = pb[nc - 1] / A2[nc - 1];
  for (int i = nc - 2; i >= 0; i--) {
    double *ppAij = pA + i * nc + (i + 1), sum = 0;
    for (int j = i + 1; j < nc; j++) {
      sum += *ppAij * pX[j];
      ppAij++;
    }
    pX[i] = (pb[i] - sum) / A2[i];
  }
}

static void gauss_newton(const double* L, const double* rho, double* betas) {
  double a[24], b[6], x[4] = {0, 0, 0, 0};
  for (int k = 0; k < 5; k++) {
    compute_A_and_b_gauss_newton(L, rho, betas, a, b);
    qr_solve(a, 6, 4, b, x);
    for (int i = 0; i < 4; i++) betas[i] += x[i];
  }
}

static void compute_ccs(Epnp* e, const double* betas, const double* ut) {
  for (int i = 0; i < 4; i++) e->ccs[i][0] = e->ccs[i][1] = e->ccs[i][2] = 0.0f;
  for (int i = 0; i < 4; i++) {
    const double* v = ut + 12 * (11 - i);
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 3; k++) e->ccs[j][k] += betas[i] * v[3 * j + k];
  }
}

static void compute_pcs(Epnp* e) {
  for (int i = 0; i < e->n; i++) {
    const double* a = e->alphas + 4 * i;
    double* pc = e->pcs + 3 * i;
    for (int j = 0; j < 3; j++) pc[j] = a[0] * e->ccs[0][j] + a[1] * e->ccs[1][j] + a[2] * e->ccs[2][j] + a[3] * e->ccs[3][j];
  }
}

static void solve_for_sign(Epnp* e) {
  if (e->pcs[2] < 0.0) {
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 3; j++) e->ccs[i][j] = -e->ccs[i][j];
    for (int i = 0; i < e->n; i++) {
      e->pcs[3 * i] = -e->pcs[3 * i];
      e->pcs[3 * i + 1] = -e->pcs[3 * i + 1];
      e->pcs[3 * i + 2] = -e->pcs[3 * i + 2];
    }
  }
}

static void estimate_R_and_t(Epnp* e, double R[3][3], double t[3]) {
  double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
  const int n = e->n;
  for (int i = 0; i < n; i++) {
    const double* pc = e->pcs + 3 * i;
    const double* pw = e->pws + 3 * i;
    for (int j = 0; j < 3; j++) {
      pc0[j] += pc[j];
      pw0[j] += pw[j];
    }
  }
  for (int j = 0; j < 3; j++) {
    pc0[j] /= n;
    pw0[j] /= n;
  }
  double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, abt_d[3], abt_u[9], abt_vt[9];
  for (int i = 0; i < n; i++) {
    const double* pc = e->pcs + 3 * i;
    const double* pw = e->pws + 3 * i;
    for (int j = 0; j < 3; j++) {
      abt[3 * j] += (pc[j] - pc0[j]) * (pw[0] - pw0[0]);
      abt[3 * j + 1] += (pc[j] - pc0[j]) * (pw[1] - pw0[1]);
      abt[3 * j + 2] += (pc[j] - pc0[j]) * (pw[2] - pw0[2]);
    }
  }
  cv_svd(abt, 3, 3, abt_d, abt_u, abt_vt);
  /* R[i][j] = dot(U row i, V row j); V row j = column j of Vt */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      R[i][j] = abt_u[3 * i] * abt_vt[j] + abt_u[3 * i + 1] * abt_vt[3 + j] + abt_u[3 * i + 2] * abt_vt[6 + j];
  const double det = R[0][0] * R[1][1] * R[2][2] + R[0][1] * R[1][2] * R[2][0] + R[0][2] * R[1][0] * R[2][1] -
                     R[0][2] * R[1][1] * R[2][0] - R[0][1] * R[1][0] * R[2][2] - R[0][0] * R[1][2] * R[2][1];
  if (det < 0) {
    R[2][0] = -R[2][0];
    R[2][1] = -R[2][1];
    R[2][2] = -R[2][2];
  }
  t[0] = pc0[0] - dot3(R[0], pw0);
  t[1] = pc0[1] - dot3(R[1], pw0);
  t[2] = pc0[2] - dot3(R[2], pw0);
}

static double reprojection_error(const Epnp* e, double R[3][3], const double t[3]) {
  double sum2 = 0.0;
  for (int i = 0; i < e->n; i++) {
    const double* pw = e->pws + 3 * i;
    double Xc = dot3(R[0], pw) + t[0];
    double Yc = dot3(R[1], pw) + t[1];
    double inv_Zc = 1.0 / (dot3(R[2], pw) + t[2]);
    double ue = e->uc + e->fu * Xc * inv_Zc;
    double ve = e->vc + e->fv * Yc * inv_Zc;
    double u = e->us[2 * i], v = e->us[2 * i + 1];
    sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
  }
  return sum2 / e->n;
}

static double compute_R_and_t(Epnp* e, const double* ut, const double* betas, double R[3][3], double t[3]) {
  compute_ccs(e, betas, ut);
  compute_pcs(e);
  solve_for_sign(e);
  estimate_R_and_t(e, R, t);
  return reprojection_error(e, R, t);
}

/* epnp::compute_pose for n <= 16 correspondences.  X (n,3) and px (n,2) float32 as solvePnPRansac hands them over;
 * K row-major 3x3.  Outputs R (9), t (3); optional dbg (>= 12*12 + 12 + 4 doubles): Ut | D | sweeps, N.
 * Returns the index (1..3) of the winning beta approximation. */
int cvo_epnp(const float* X, const float* px, int n, const double* K, double* Rout, double* tout, double* dbg) {
  Epnp e;
  double pws[3 * 16], us[2 * 16];
  if (n > 16) return -1;
  e.fu = K[0];
  e.fv = K[4];
  e.uc = K[2];
  e.vc = K[5];
  e.n = n;
  const double ifx = 1. / e.fu, ify = 1. / e.fv;
  for (int i = 0; i < n; i++) {
    pws[3 * i] = X[3 * i];
    pws[3 * i + 1] = X[3 * i + 1];
    pws[3 * i + 2] = X[3 * i + 2];
    /* cv::undistortPoints (zero distortion) keeps the input dtype: normalised coordinates rounded to float32 */
    float xn = (float)(((double)px[2 * i] - e.uc) * ifx);
    float yn = (float)(((double)px[2 * i + 1] - e.vc) * ify);
    us[2 * i] = xn * e.fu + e.uc;
    us[2 * i + 1] = yn * e.fv + e.vc;
  }
  e.pws = pws;
  e.us = us;
  choose_control_points(&e);
  compute_barycentric_coordinates(&e);
  double M[2 * 16 * 12], mtm[144], d[12], u[144], vt[144], ut[144];
  for (int i = 0; i < n; i++) fill_M(&e, M, 2 * i, e.alphas + 4 * i, us[2 * i], us[2 * i + 1]);
  cvo_mul_transposed(M, 2 * n, 12, mtm);
  {
    double At[144];
    for (int i = 0; i < 12; i++)
      for (int k = 0; k < 12; k++) At[i * 12 + k] = mtm[k * 12 + i];
    int sweeps = cvo_jacobi_svd(At, d, vt, 12, 12);
    memcpy(ut, At, sizeof(ut)); /* Ut: row i = i-th left singular vector */
    if (dbg) dbg[156] = sweeps;
  }
  (void)u;
  double L[60], rho[6];
  compute_L_6x10(ut, L);
  compute_rho(&e, rho);
  double Betas[4][4], rep_errors[4], Rs[4][3][3], ts[4][3];
  memset(Betas, 0, sizeof(Betas));
  find_betas_approx_1(L, rho, Betas[1]);
  gauss_newton(L, rho, Betas[1]);
  rep_errors[1] = compute_R_and_t(&e, ut, Betas[1], Rs[1], ts[1]);
  find_betas_approx_2(L, rho, Betas[2]);
  gauss_newton(L, rho, Betas[2]);
  rep_errors[2] = compute_R_and_t(&e, ut, Betas[2], Rs[2], ts[2]);
  find_betas_approx_3(L, rho, Betas[3]);
  gauss_newton(L, rho, Betas[3]);
  rep_errors[3] = compute_R_and_t(&e, ut, Betas[3], Rs[3], ts[3]);
  int N = 1;
  if (rep_errors[2] < rep_errors[1]) N = 2;
  if (rep_errors[3] < rep_errors[N]) N = 3;
  for (int i = 0; i < 3; i++) {
    tout[i] = ts[N][i];
    for (int j = 0; j < 3; j++) Rout[3 * i + j] = Rs[N][i][j];
  }
  if (dbg) {
    memcpy(dbg, ut, sizeof(ut));
    memcpy(dbg + 144, d, sizeof(d));
    dbg[157] = N;
  }
  return N;
}
