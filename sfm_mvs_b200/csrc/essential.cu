// essential.cu — two-view initialisation: cv2.findEssentialMat(pts0, pts1, K, RANSAC, prob, threshold)
// (reference: sfm.py:307, isfm.py:80, test.py:247).
//
// OpenCV's loop (five-point.cpp + ptsetreg.cpp) draws a five-index subset per iteration, solves Nister's
// five-point problem (up to ten essential matrices per subset), scores every model with the Sampson error over
// all N correspondences and shrinks the iteration budget whenever a model beats the best count.  The subsets
// depend only on N (RNG seeded with 2^64-1), so all iterations are independent until the accept/stop recursion:
//   e5_normalize_kernel   pixels -> normalised float64 coordinates, once
//   e5_solve_kernel       one thread per iteration: null space of the 5x9 epipolar system, the ten cubic
//                         constraints as a 10x20 matrix built by polynomial arithmetic, Gauss-Jordan, the 3x3
//                         polynomial matrix B(z), det B(z) (degree 10), its real roots, one model per root
//   e5_score_kernel       one CTA per iteration: Sampson error of every model x every point, inlier counts
//   e5_replay_kernel      the sequential accept / RANSACUpdateNumIters recursion over the count table
//   e5_mask_kernel        inlier mask of the winner
// No refit follows in OpenCV: the returned E is the winning minimal-sample model.
//
// Model order inside an iteration (which only matters when two models of one subset tie on the count): OpenCV
// visits roots in the order its Durand-Kerner iteration leaves them, for a null-space basis its SVD happens to
// return; neither is reproducible from the published algorithm.  Here models are ordered by ascending E[0][0]^2
// of the unit-norm matrix, which is independent of the basis (oracle/restated.py five_point does the same).
#include <float.h>

#include "common.cuh"
#include "ransac.cuh"

namespace {

constexpr int E5_MAXM = 10;       // models per iteration
constexpr int E5_MODEL_POINTS = 5;

// ---- monomial bookkeeping.  Linear terms: 0:x 1:y 2:z 3:1.
// Quadratic monomials (products of two linear terms), index by the sorted pair.
__host__ __device__ constexpr int quad_index(int a, int b) {
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return lo == 0 ? hi : lo == 1 ? 3 + hi : lo == 2 ? 5 + hi : 9;
}
// exponents (x, y, z) of the linear / quadratic monomials packed as ex*16 + ey*4 + ez
__host__ __device__ constexpr int lin_exp(int a) { return a == 0 ? 16 : a == 1 ? 4 : a == 2 ? 1 : 0; }
__host__ __device__ constexpr int quad_exp(int q) {
  return q == 0 ? 32 : q == 1 ? 20 : q == 2 ? 17 : q == 3 ? 16 : q == 4 ? 8 : q == 5 ? 5 : q == 6 ? 4 : q == 7 ? 2 : q == 8 ? 1 : 0;
}
// Column of a cubic monomial in Nister's elimination order:
//   x^3 y^3 x^2y xy^2 x^2z x^2 y^2z y^2 xyz xy | xz^2 xz x yz^2 yz y z^3 z^2 z 1
__host__ __device__ constexpr int cubic_col_exp(int e) {
  return e == 48 ? 0 : e == 12 ? 1 : e == 36 ? 2 : e == 24 ? 3 : e == 33 ? 4 : e == 32 ? 5 : e == 9 ? 6 : e == 8 ? 7 :
         e == 21 ? 8 : e == 20 ? 9 : e == 18 ? 10 : e == 17 ? 11 : e == 16 ? 12 : e == 6 ? 13 : e == 5 ? 14 :
         e == 4 ? 15 : e == 3 ? 16 : e == 2 ? 17 : e == 1 ? 18 : 19;
}
__host__ __device__ constexpr int cubic_col(int q, int l) { return cubic_col_exp(quad_exp(q) + lin_exp(l)); }

// q += s * a * b   (linear x linear -> quadratic)
__host__ __device__ __forceinline__ void mul_ll(const double* a, const double* b, double* q, double s) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) q[quad_index(i, j)] += s * a[i] * b[j];
}
// c += q * l   (quadratic x linear -> cubic, scattered to the elimination order)
__host__ __device__ __forceinline__ void mul_ql(const double* q, const double* l, double* c) {
#pragma unroll
  for (int i = 0; i < 10; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[cubic_col(i, j)] += q[i] * l[j];
}

// ---- null space of the 5x9 epipolar system: Gauss-Jordan with complete pivoting, then two rounds of modified
// Gram-Schmidt so the four basis vectors are orthonormal (the solutions do not depend on the basis; an
// orthonormal one keeps the cubic system well scaled).  basis[k][9], k = X, Y, Z, W.
__host__ __device__ inline bool null_space_5x9(double (*Q)[9], double (*basis)[9]) {
  int perm[9];
  for (int j = 0; j < 9; ++j) perm[j] = j;
  for (int r = 0; r < 5; ++r) {
    int pi = r, pj = r;
    double best = -1.0;
    for (int i = r; i < 5; ++i)
      for (int j = r; j < 9; ++j) {
        const double v = fabs(Q[i][j]);
        if (v > best) { best = v; pi = i; pj = j; }
      }
    if (!(best > 1e-300)) return false;
    if (pi != r)
      for (int j = 0; j < 9; ++j) { const double t = Q[r][j]; Q[r][j] = Q[pi][j]; Q[pi][j] = t; }
    if (pj != r) {
      for (int i = 0; i < 5; ++i) { const double t = Q[i][r]; Q[i][r] = Q[i][pj]; Q[i][pj] = t; }
      const int t = perm[r]; perm[r] = perm[pj]; perm[pj] = t;
    }
    const double inv = 1.0 / Q[r][r];
    for (int j = r; j < 9; ++j) Q[r][j] *= inv;
    for (int i = 0; i < 5; ++i) {
      if (i == r) continue;
      const double f = Q[i][r];
      for (int j = r; j < 9; ++j) Q[i][j] -= f * Q[r][j];
    }
  }
  for (int k = 0; k < 4; ++k) {
    for (int j = 0; j < 9; ++j) basis[k][j] = 0.0;
    basis[k][perm[5 + k]] = 1.0;
    for (int i = 0; i < 5; ++i) basis[k][perm[i]] = -Q[i][5 + k];
  }
  for (int k = 0; k < 4; ++k) {
    for (int pass = 0; pass < 2; ++pass)
      for (int m = 0; m < k; ++m) {
        double d = 0.0;
        for (int j = 0; j < 9; ++j) d += basis[k][j] * basis[m][j];
        for (int j = 0; j < 9; ++j) basis[k][j] -= d * basis[m][j];
      }
    double nn = 0.0;
    for (int j = 0; j < 9; ++j) nn += basis[k][j] * basis[k][j];
    if (!(nn > 1e-300)) return false;
    const double inv = 1.0 / sqrt(nn);
    for (int j = 0; j < 9; ++j) basis[k][j] *= inv;
  }
  return true;
}

// ---- the ten cubic constraints: det(E) = 0 and (E E^T - tr(E E^T)/2 I) E = 0, E = x X + y Y + z Z + W
__host__ __device__ inline void build_constraints(const double (*basis)[9], double (*M)[20]) {
  double e[9][4];
#pragma unroll
  for (int c = 0; c < 9; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) e[c][k] = basis[k][c];
#pragma unroll
  for (int r = 0; r < 10; ++r)
#pragma unroll
    for (int c = 0; c < 20; ++c) M[r][c] = 0.0;
  {  // determinant, expanded along the first row
    double m0[10], m1[10], m2[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) { m0[k] = 0.0; m1[k] = 0.0; m2[k] = 0.0; }
    mul_ll(e[4], e[8], m0, 1.0); mul_ll(e[5], e[7], m0, -1.0);
    mul_ll(e[5], e[6], m1, 1.0); mul_ll(e[3], e[8], m1, -1.0);
    mul_ll(e[3], e[7], m2, 1.0); mul_ll(e[4], e[6], m2, -1.0);
    mul_ql(m0, e[0], M[0]); mul_ql(m1, e[1], M[0]); mul_ql(m2, e[2], M[0]);
  }
  double L[6][10];     // Lambda = E E^T - tr/2 I, symmetric: (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
#pragma unroll
  for (int s = 0; s < 6; ++s)
#pragma unroll
    for (int k = 0; k < 10; ++k) L[s][k] = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    mul_ll(e[0 + k], e[0 + k], L[0], 1.0);
    mul_ll(e[0 + k], e[3 + k], L[1], 1.0);
    mul_ll(e[0 + k], e[6 + k], L[2], 1.0);
    mul_ll(e[3 + k], e[3 + k], L[3], 1.0);
    mul_ll(e[3 + k], e[6 + k], L[4], 1.0);
    mul_ll(e[6 + k], e[6 + k], L[5], 1.0);
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const double h = 0.5 * (L[0][k] + L[3][k] + L[5][k]);
    L[0][k] -= h; L[3][k] -= h; L[5][k] -= h;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double* row = M[1 + 3 * i + j];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int lo = i < k ? i : k, hi = i < k ? k : i;
        const int s = lo == 0 ? hi : lo == 1 ? 2 + hi : 5;
        mul_ql(L[s], e[3 * k + j], row);
      }
    }
}

// Gauss-Jordan on the first ten columns (partial pivoting); afterwards row i reads
// monomial_i + sum_j M[i][10+j] * tail_j = 0.
__host__ __device__ inline bool reduce_constraints(double (*M)[20]) {
  for (int c = 0; c < 10; ++c) {
    int p = c;
    double best = fabs(M[c][c]);
    for (int r = c + 1; r < 10; ++r) {
      const double v = fabs(M[r][c]);
      if (v > best) { best = v; p = r; }
    }
    if (!(best > 1e-300)) return false;
    if (p != c)
      for (int j = c; j < 20; ++j) { const double t = M[c][j]; M[c][j] = M[p][j]; M[p][j] = t; }
    const double inv = 1.0 / M[c][c];
    for (int j = c; j < 20; ++j) M[c][j] *= inv;
    for (int r = 0; r < 10; ++r) {
      if (r == c) continue;
      const double f = M[r][c];
      if (f != 0.0)
        for (int j = c; j < 20; ++j) M[r][j] -= f * M[c][j];
    }
  }
  return true;
}

__host__ __device__ __forceinline__ double horner(const double* c, int deg, double x) {   // ascending coefficients
  double v = c[deg];
  for (int k = deg - 1; k >= 0; --k) v = fma(v, x, c[k]);
  return v;
}

// A root of q (degree m, ascending coefficients) bracketed by [lo, hi] with sign(q(lo)) = slo != sign(q(hi)):
// Newton steps kept inside the bracket, bisection otherwise.
__host__ __device__ inline double bracketed_root(const double* q, int m, double lo, double hi, int slo) {
  double x = 0.5 * (lo + hi), dxold = hi - lo, dx = dxold;
  for (int it = 0; it < 200; ++it) {
    double f = q[m], df = 0.0;
    for (int k = m - 1; k >= 0; --k) { df = fma(df, x, f); f = fma(f, x, q[k]); }
    if (f == 0.0) return x;
    if ((f < 0.0) == (slo < 0)) lo = x; else hi = x;
    const double xn = x - f / df;
    dxold = dx;
    if (!(xn > lo && xn < hi) || fabs(2.0 * f) > fabs(dxold * df)) {
      dx = 0.5 * (hi - lo);
      const double xm = lo + dx;
      if (xm == lo || xm == hi) return xm;
      x = xm;
    } else {
      dx = xn - x;
      x = xn;
      if (fabs(dx) <= 2.3e-16 * fabs(x)) return x;
    }
    if (hi - lo <= 2.3e-16 * fmax(fabs(lo), fabs(hi))) return x;
  }
  return x;
}

// All real roots of p (ascending coefficients, degree n <= 10), ascending: the real roots of the k-th derivative
// split the line into intervals on which the (k-1)-th derivative is monotonic, so each level's roots are
// bracketed by the previous level's.  Roots of even multiplicity (no sign change) are not reported.
__host__ __device__ inline int real_roots(const double* p, int n, double* roots) {
  // Fujiwara's bound on |root|
  double R = 0.0;
  for (int k = 1; k <= n; ++k) {
    double a = fabs(p[n - k] / p[n]);
    if (k == n) a *= 0.5;
    if (a > 0.0) R = fmax(R, pow(a, 1.0 / k));
  }
  R = 2.0 * R * (1.0 + 1e-9) + 1e-300;
  if (!(R < 1e30)) R = 1e30;
  double prev[E5_MAXM], cur[E5_MAXM], q[E5_MAXM + 1];
  int np = 0;
  for (int m = 1; m <= n; ++m) {
    const int d = n - m;                       // q = d-th derivative of p
    for (int i = 0; i <= m; ++i) {
      double c = p[i + d];
      for (int k = 1; k <= d; ++k) c *= (double)(i + k);
      q[i] = c;
    }
    const int slead = q[m] > 0.0 ? 1 : -1;
    int nc = 0;
    double lo = -R;
    int slo = (m & 1) ? -slead : slead;
    for (int k = 0; k <= np; ++k) {
      const double hi = k < np ? prev[k] : R;
      int shi;
      if (k < np) {
        const double v = horner(q, m, hi);
        shi = v > 0.0 ? 1 : (v < 0.0 ? -1 : 0);
      } else {
        shi = slead;
      }
      if (shi == 0) {
        // the bracket end is itself a root (a stationary point of q on the axis); the sign to its right is read
        // a hair inside the next interval
        if (nc < E5_MAXM) cur[nc++] = hi;
        const double v = horner(q, m, hi + 1e-8 * (1.0 + fabs(hi)));
        lo = hi;
        slo = v > 0.0 ? 1 : (v < 0.0 ? -1 : 0);
        continue;
      }
      if (slo != 0 && slo != shi && hi > lo) {
        if (nc < E5_MAXM) cur[nc++] = bracketed_root(q, m, lo, hi, slo);
      }
      lo = hi;
      slo = shi;
    }
    np = nc;
    for (int k = 0; k < nc; ++k) prev[k] = cur[k];
  }
  for (int k = 0; k < np; ++k) roots[k] = prev[k];
  return np;
}

__global__ void e5_normalize_kernel(const void* __restrict__ p1, const void* __restrict__ p2, int is_f64, int n,
                                    double fx, double fy, double cx, double cy, double* __restrict__ qn /*n x 4*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a, b, c, d;
  if (is_f64) {
    const double* s1 = (const double*)p1; const double* s2 = (const double*)p2;
    a = s1[2 * i]; b = s1[2 * i + 1]; c = s2[2 * i]; d = s2[2 * i + 1];
  } else {
    const float* s1 = (const float*)p1; const float* s2 = (const float*)p2;
    a = (double)s1[2 * i]; b = (double)s1[2 * i + 1]; c = (double)s2[2 * i]; d = (double)s2[2 * i + 1];
  }
  qn[4 * i + 0] = (a - cx) / fx;
  qn[4 * i + 1] = (b - cy) / fy;
  qn[4 * i + 2] = (c - cx) / fx;
  qn[4 * i + 3] = (d - cy) / fy;
}

// One minimal sample: Q = the 5x9 epipolar system (destroyed).  Writes the essential matrices (unit Frobenius norm,
// ascending E00^2) to out[10][9] and returns how many.  Host-callable too (sfm_five_point), so the CPU tests exercise
// the very code the kernel runs.
__host__ __device__ inline int five_point_solve(double (*Q)[9], double* out) {
  int nm = 0;
  double basis[4][9];
  double M[10][20];
  if (!null_space_5x9(Q, basis)) return 0;
  build_constraints(basis, M);
  if (!reduce_constraints(M)) return 0;
  // B(z): rows <x^2 z> - z <x^2>, <y^2 z> - z <y^2>, <xyz> - z <xy>; each is x p3(z) + y q3(z) + r4(z).
  // Ascending coefficients; entries 0, 1 have degree 3, entry 2 degree 4.
  double B[3][3][5];
  for (int i = 0; i < 3; ++i) {
    const double* ra = &M[2 * i + 4][10];
    const double* rb = &M[2 * i + 5][10];
    for (int c = 0; c < 2; ++c) {
      // tail order: z^2 z 1  ->  ascending index 2 1 0
      B[i][c][0] = ra[3 * c + 2];
      B[i][c][1] = ra[3 * c + 1] - rb[3 * c + 2];
      B[i][c][2] = ra[3 * c + 0] - rb[3 * c + 1];
      B[i][c][3] = -rb[3 * c + 0];
      B[i][c][4] = 0.0;
    }
    B[i][2][0] = ra[9];
    B[i][2][1] = ra[8] - rb[9];
    B[i][2][2] = ra[7] - rb[8];
    B[i][2][3] = ra[6] - rb[7];
    B[i][2][4] = -rb[6];
  }
  double det[11];
  for (int k = 0; k < 11; ++k) det[k] = 0.0;
  {
    // minors of rows 1, 2: m12 (deg 7), m02 (deg 7), m01 (deg 6)
    double m12[8], m02[8], m01[7];
    for (int k = 0; k < 8; ++k) { m12[k] = 0.0; m02[k] = 0.0; }
    for (int k = 0; k < 7; ++k) m01[k] = 0.0;
    for (int a = 0; a <= 3; ++a) {
      for (int b = 0; b <= 4; ++b) {
        m12[a + b] += B[1][1][a] * B[2][2][b] - B[2][1][a] * B[1][2][b];
        m02[a + b] += B[1][0][a] * B[2][2][b] - B[2][0][a] * B[1][2][b];
      }
      for (int b = 0; b <= 3; ++b) m01[a + b] += B[1][0][a] * B[2][1][b] - B[1][1][a] * B[2][0][b];
    }
    for (int a = 0; a <= 3; ++a)
      for (int b = 0; b <= 7; ++b) det[a + b] += B[0][0][a] * m12[b] - B[0][1][a] * m02[b];
    for (int a = 0; a <= 4; ++a)
      for (int b = 0; b <= 6; ++b) det[a + b] += B[0][2][a] * m01[b];
  }
  bool finite = true;
  for (int k = 0; k < 11; ++k) finite = finite && isfinite(det[k]);
  if (!finite) return 0;
  int deg = 10;
  while (deg > 1 && fabs(det[deg]) <= DBL_EPSILON) --deg;     // solvePoly's leading-coefficient trim
  if (fabs(det[deg]) == 0.0) return 0;
  double roots[E5_MAXM];
  const int nr = real_roots(det, deg, roots);
  double key[E5_MAXM];
  for (int r = 0; r < nr; ++r) {
    const double z = roots[r];
    double bz[3][3];
    for (int i = 0; i < 3; ++i) {
      bz[i][0] = horner(B[i][0], 3, z);
      bz[i][1] = horner(B[i][1], 3, z);
      bz[i][2] = horner(B[i][2], 4, z);
    }
    // null vector of the (rank-2) matrix: the largest of the three row cross products
    double best = -1.0, v[3] = {0, 0, 0};
    for (int a = 0; a < 3; ++a) {
      const int i0 = a == 2 ? 1 : 0, i1 = a == 0 ? 1 : 2;
      const double c0 = bz[i0][1] * bz[i1][2] - bz[i0][2] * bz[i1][1];
      const double c1 = bz[i0][2] * bz[i1][0] - bz[i0][0] * bz[i1][2];
      const double c2 = bz[i0][0] * bz[i1][1] - bz[i0][1] * bz[i1][0];
      const double nn = c0 * c0 + c1 * c1 + c2 * c2;
      if (nn > best) { best = nn; v[0] = c0; v[1] = c1; v[2] = c2; }
    }
    if (!(best > 0.0)) continue;
    const double vn = sqrt(best);
    if (fabs(v[2]) < 1e-10 * vn) continue;           // OpenCV: |xy1(2)| < 1e-10 on the unit vector
    const double x = v[0] / v[2], y = v[1] / v[2];
    double E[9], nn = 0.0;
    for (int c = 0; c < 9; ++c) {
      E[c] = x * basis[0][c] + y * basis[1][c] + z * basis[2][c] + basis[3][c];
      nn += E[c] * E[c];
    }
    if (!(nn > 0.0) || !isfinite(nn)) continue;
    const double inv = 1.0 / sqrt(nn);
    for (int c = 0; c < 9; ++c) E[c] *= inv;
    // insertion by ascending E00^2
    const double kk = E[0] * E[0];
    int pos = nm;
    while (pos > 0 && key[pos - 1] > kk) {
      key[pos] = key[pos - 1];
      for (int c = 0; c < 9; ++c) out[9 * pos + c] = out[9 * (pos - 1) + c];
      --pos;
    }
    key[pos] = kk;
    for (int c = 0; c < 9; ++c) out[9 * pos + c] = E[c];
    ++nm;
  }
  return nm;
}

constexpr int E5_SOLVE_THREADS = 8;   // few lanes per warp: the 1000 independent solves spread over every SM and a
                                      // warp waits only for the slowest of 8 data-dependent root searches
__global__ void __launch_bounds__(32) e5_solve_kernel(const double* __restrict__ qn, const int* __restrict__ subsets,
                                                      int iters, double* __restrict__ models /*iters x 10 x 9*/,
                                                      int* __restrict__ nmodels) {
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= iters) return;
  double Q[5][9];
  for (int r = 0; r < 5; ++r) {
    const int i = subsets[5 * it + r];
    const double x1 = qn[4 * i], y1 = qn[4 * i + 1], x2 = qn[4 * i + 2], y2 = qn[4 * i + 3];
    // x2^T E x1 = 0, E row-major
    Q[r][0] = x2 * x1; Q[r][1] = x2 * y1; Q[r][2] = x2;
    Q[r][3] = y2 * x1; Q[r][4] = y2 * y1; Q[r][5] = y2;
    Q[r][6] = x1;      Q[r][7] = y1;      Q[r][8] = 1.0;
  }
  nmodels[it] = five_point_solve(Q, models + (size_t)it * E5_MAXM * 9);
}

// EMEstimatorCallback::computeError: (x2^T E x1)^2 / (|E x1|_xy^2 + |E^T x2|_xy^2) in float64 with separately
// rounded products (OpenCV's Matx arithmetic), stored as float32 and compared with (float)(thr^2).
__device__ __forceinline__ float sampson(const double* __restrict__ E, double x1, double y1, double x2, double y2) {
  const double ex0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], x1), __dmul_rn(E[1], y1)), E[2]);
  const double ex1 = __dadd_rn(__dadd_rn(__dmul_rn(E[3], x1), __dmul_rn(E[4], y1)), E[5]);
  const double ex2 = __dadd_rn(__dadd_rn(__dmul_rn(E[6], x1), __dmul_rn(E[7], y1)), E[8]);
  const double et0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], x2), __dmul_rn(E[3], y2)), E[6]);
  const double et1 = __dadd_rn(__dadd_rn(__dmul_rn(E[1], x2), __dmul_rn(E[4], y2)), E[7]);
  const double s = __dadd_rn(__dadd_rn(__dmul_rn(x2, ex0), __dmul_rn(y2, ex1)), ex2);
  const double den = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(ex0, ex0), __dmul_rn(ex1, ex1)), __dmul_rn(et0, et0)),
                               __dmul_rn(et1, et1));
  return (float)(__dmul_rn(s, s) / den);
}

__global__ void __launch_bounds__(256) e5_score_kernel(const double* __restrict__ qn, int n,
                                                       const double* __restrict__ models,
                                                       const int* __restrict__ nmodels, float t2,
                                                       int* __restrict__ counts /*iters x 10*/) {
  __shared__ double sE[E5_MAXM * 9];
  __shared__ int sc[E5_MAXM];
  const int it = blockIdx.x;
  const int nm = nmodels[it];
  if (threadIdx.x < E5_MAXM) { sc[threadIdx.x] = 0; counts[it * E5_MAXM + threadIdx.x] = 0; }
  if (nm == 0) return;
  for (int k = threadIdx.x; k < nm * 9; k += blockDim.x) sE[k] = models[(size_t)it * E5_MAXM * 9 + k];
  __syncthreads();
  int cnt[E5_MAXM];
#pragma unroll
  for (int m = 0; m < E5_MAXM; ++m) cnt[m] = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double4 q = reinterpret_cast<const double4*>(qn)[i];
#pragma unroll
    for (int m = 0; m < E5_MAXM; ++m)
      if (m < nm) cnt[m] += (sampson(sE + 9 * m, q.x, q.y, q.z, q.w) <= t2) ? 1 : 0;
  }
#pragma unroll
  for (int m = 0; m < E5_MAXM; ++m) {
    if (m < nm) {
      int v = cnt[m];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sc[m], v);
    }
  }
  __syncthreads();
  if (threadIdx.x < nm) counts[it * E5_MAXM + threadIdx.x] = sc[threadIdx.x];
}

struct E5Result {
  double E[9];
  int ok, best_iter, best_model, best_count, iters_run, models_total, pad0, pad1;
};

// RANSACPointSetRegistrator::run's accept / stop recursion over the table of counts
__global__ void e5_replay_kernel(const int* __restrict__ counts, const int* __restrict__ nmodels,
                                 const double* __restrict__ models, int n, int max_iters, double prob,
                                 E5Result* __restrict__ res) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int niters = max_iters, best = 0, bi = -1, bm = -1, it = 0, total = 0;
  while (it < niters) {
    const int nm = nmodels[it];
    total += nm;
    for (int m = 0; m < nm; ++m) {
      const int c = counts[it * E5_MAXM + m];
      if (c > max(best, E5_MODEL_POINTS - 1)) {
        best = c; bi = it; bm = m;
        niters = update_num_iters(prob, (double)(n - c) / n, E5_MODEL_POINTS, niters);
      }
    }
    ++it;
  }
  res->ok = bi >= 0;
  res->best_iter = bi; res->best_model = bm; res->best_count = best; res->iters_run = it; res->models_total = total;
  for (int c = 0; c < 9; ++c) res->E[c] = bi >= 0 ? models[((size_t)bi * E5_MAXM + bm) * 9 + c] : 0.0;
}

__global__ void e5_mask_kernel(const double* __restrict__ qn, int n, const E5Result* __restrict__ res, float t2,
                               unsigned char* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!res->ok) { mask[i] = 0; return; }
  const double4 q = reinterpret_cast<const double4*>(qn)[i];
  mask[i] = sampson(res->E, q.x, q.y, q.z, q.w) <= t2 ? 1 : 0;
}

}  // namespace

// ---- host utility (usable without a GPU): the minimal solver on one sample of normalised coordinates
extern "C" int sfm_five_point(const double* q1, const double* q2, double* E, int32_t* n_models) {
  SFM_REQUIRE(q1 && q2 && E && n_models, "sfm_five_point: null argument");
  double Q[5][9];
  for (int r = 0; r < 5; ++r) {
    const double x1 = q1[2 * r], y1 = q1[2 * r + 1], x2 = q2[2 * r], y2 = q2[2 * r + 1];
    Q[r][0] = x2 * x1; Q[r][1] = x2 * y1; Q[r][2] = x2;
    Q[r][3] = y2 * x1; Q[r][4] = y2 * y1; Q[r][5] = y2;
    Q[r][6] = x1;      Q[r][7] = y1;      Q[r][8] = 1.0;
  }
  *n_models = five_point_solve(Q, E);
  return SFM_OK;
}

extern "C" int sfm_find_essential_mat(sfm_ctx* ctx, const void* pts1, const void* pts2, int dtype, int n,
                                      const double* K, double prob, double threshold, int max_iters,
                                      double* E, uint8_t* mask, int32_t* info) {
  SFM_REQUIRE(ctx && pts1 && pts2 && K && E && info, "sfm_find_essential_mat: null argument");
  SFM_REQUIRE(dtype == 0 || dtype == 2, "sfm_find_essential_mat: points must be float32 (0) or float64 (2)");
  SFM_REQUIRE(n >= 0, "sfm_find_essential_mat: negative point count");
  SFM_REQUIRE(prob > 0.0 && prob < 1.0, "sfm_find_essential_mat: confidence must lie in (0, 1)");
  SFM_REQUIRE(max_iters <= 100000, "sfm_find_essential_mat: maxIters above 100000");
  for (int k = 0; k < 6; ++k) info[k] = 0;
  info[3] = -1; info[4] = -1;
  if (n < E5_MODEL_POINTS) return SFM_OK;            // cv2 returns (None, None)
  SFM_TRY(sfm_ws_begin(ctx));
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const double thr = threshold / ((fx + fy) / 2);
  const float t2 = (float)(thr * thr);
  const size_t esz = dtype == 0 ? sizeof(float) : sizeof(double);
  const void *d1 = pts1, *d2 = pts2;
  if (!sfm_is_device_ptr(pts1)) { char* p; SFM_TRY(ws_alloc_t(ctx, 2 * (size_t)n * esz, &p)); SFM_CUDA(cudaMemcpyAsync(p, pts1, 2 * (size_t)n * esz, cudaMemcpyHostToDevice, ctx->stream)); d1 = p; }
  if (!sfm_is_device_ptr(pts2)) { char* p; SFM_TRY(ws_alloc_t(ctx, 2 * (size_t)n * esz, &p)); SFM_CUDA(cudaMemcpyAsync(p, pts2, 2 * (size_t)n * esz, cudaMemcpyHostToDevice, ctx->stream)); d2 = p; }
  const int iters = n == E5_MODEL_POINTS ? 1 : (max_iters > 1 ? max_iters : 1);
  double* qn; double* models; int* nmodels; int* counts; int* subs; E5Result* res; unsigned char* dmask;
  SFM_TRY(ws_alloc_t(ctx, (size_t)4 * n, &qn));
  SFM_TRY(ws_alloc_t(ctx, (size_t)iters * E5_MAXM * 9, &models));
  SFM_TRY(ws_alloc_t(ctx, (size_t)iters, &nmodels));
  SFM_TRY(ws_alloc_t(ctx, (size_t)iters * E5_MAXM, &counts));
  SFM_TRY(ws_alloc_t(ctx, (size_t)iters * 5, &subs));
  SFM_TRY(ws_alloc_t(ctx, 1, &res));
  SFM_TRY(ws_alloc_t(ctx, (size_t)n, &dmask));
  int* hsubs;
  SFM_TRY(hs_alloc_t(ctx, (size_t)iters * 5, &hsubs));
  if (n == E5_MODEL_POINTS) { for (int k = 0; k < 5; ++k) hsubs[k] = k; }
  else ransac_subsets(n, iters, hsubs);
  SFM_CUDA(cudaMemcpyAsync(subs, hsubs, sizeof(int) * 5 * (size_t)iters, cudaMemcpyHostToDevice, ctx->stream));
  SFM_LAUNCH(ctx, SFM_K_MISC, (e5_normalize_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(
                                  d1, d2, dtype == 2, n, fx, fy, cx, cy, qn)));
  SFM_LAUNCH(ctx, SFM_K_ESSENTIAL, (e5_solve_kernel<<<div_up(iters, E5_SOLVE_THREADS), E5_SOLVE_THREADS, 0, ctx->stream>>>(
                                       qn, subs, iters, models, nmodels)));
  if (n == E5_MODEL_POINTS) {
    // cv2 returns every model of the single minimal sample, stacked, and an all-ones mask
    int* hn; double* hm;
    SFM_TRY(hs_alloc_t(ctx, 1, &hn));
    SFM_TRY(hs_alloc_t(ctx, (size_t)E5_MAXM * 9, &hm));
    SFM_CUDA(cudaMemcpyAsync(hn, nmodels, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(cudaMemcpyAsync(hm, models, sizeof(double) * E5_MAXM * 9, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < *hn * 9; ++k) E[k] = hm[k];
    info[0] = *hn; info[1] = *hn ? n : 0; info[3] = *hn ? 0 : -1; info[4] = *hn ? 0 : -1; info[5] = *hn;
    if (mask) {
      if (sfm_is_device_ptr(mask)) SFM_CUDA(cudaMemsetAsync(mask, *hn ? 1 : 0, (size_t)n, ctx->stream));
      else memset(mask, *hn ? 1 : 0, (size_t)n);
    }
    return SFM_OK;
  }
  SFM_LAUNCH(ctx, SFM_K_ESSENTIAL, (e5_score_kernel<<<iters, 256, 0, ctx->stream>>>(qn, n, models, nmodels, t2, counts)));
  SFM_LAUNCH(ctx, SFM_K_MISC, (e5_replay_kernel<<<1, 32, 0, ctx->stream>>>(counts, nmodels, models, n, iters, prob, res)));
  SFM_LAUNCH(ctx, SFM_K_MISC, (e5_mask_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(qn, n, res, t2, dmask)));
  E5Result* hres;
  SFM_TRY(hs_alloc_t(ctx, 1, &hres));
  SFM_CUDA(cudaMemcpyAsync(hres, res, sizeof(E5Result), cudaMemcpyDeviceToHost, ctx->stream));
  if (mask) {
    const bool dev = sfm_is_device_ptr(mask);
    SFM_CUDA(cudaMemcpyAsync(mask, dmask, (size_t)n, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
  }
  SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  info[0] = hres->ok ? 1 : 0;
  info[1] = hres->ok ? hres->best_count : 0;
  info[2] = hres->iters_run;
  info[3] = hres->best_iter;
  info[4] = hres->best_model;
  info[5] = hres->models_total;
  for (int k = 0; k < 9; ++k) E[k] = hres->E[k];
  return SFM_OK;
}

// ================================================================================================================
// EXPERIMENTAL — batched over pairs (isfm.py:68-87 runs findEssentialMat once per pair).  Written at the end of
// round 1 after the GPU budget was spent: compiles, mirrors the kernels above with a pair table in blockIdx.y, but
// has NOT been run on a GPU yet; nothing in bench.py, smoke() or the default tests uses it
// (tests/test_gpu_essential.py::test_find_essential_mat_batched is skipped unless SFM_TEST_EXPERIMENTAL=1).
// The single-pair call is latency-bound (1000 solver threads, ~1.1 ms); P pairs x 1000 solves fill the machine.
// ================================================================================================================
namespace {

struct E5Pair {
  const float* p1;            // (n,2) float32 pixels, device
  const float* p2;
  int n, pad;
  double* qn;                 // n x 4 normalised
  const int* subs;            // iters x 5
  double* models;             // iters x 10 x 9
  int* nmodels;               // iters
  int* counts;                // iters x 10
  E5Result* res;
  unsigned char* mask;        // n
};

__global__ void e5b_normalize_kernel(const E5Pair* __restrict__ tab, double fx, double fy, double cx, double cy) {
  const E5Pair pr = tab[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pr.n) return;
  pr.qn[4 * i + 0] = ((double)pr.p1[2 * i] - cx) / fx;
  pr.qn[4 * i + 1] = ((double)pr.p1[2 * i + 1] - cy) / fy;
  pr.qn[4 * i + 2] = ((double)pr.p2[2 * i] - cx) / fx;
  pr.qn[4 * i + 3] = ((double)pr.p2[2 * i + 1] - cy) / fy;
}

__global__ void __launch_bounds__(32) e5b_solve_kernel(const E5Pair* __restrict__ tab, int iters) {
  const E5Pair pr = tab[blockIdx.y];
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= iters) return;
  double Q[5][9];
  for (int r = 0; r < 5; ++r) {
    const int i = pr.subs[5 * it + r];
    const double x1 = pr.qn[4 * i], y1 = pr.qn[4 * i + 1], x2 = pr.qn[4 * i + 2], y2 = pr.qn[4 * i + 3];
    Q[r][0] = x2 * x1; Q[r][1] = x2 * y1; Q[r][2] = x2;
    Q[r][3] = y2 * x1; Q[r][4] = y2 * y1; Q[r][5] = y2;
    Q[r][6] = x1;      Q[r][7] = y1;      Q[r][8] = 1.0;
  }
  pr.nmodels[it] = five_point_solve(Q, pr.models + (size_t)it * E5_MAXM * 9);
}

__global__ void __launch_bounds__(256) e5b_score_kernel(const E5Pair* __restrict__ tab, float t2) {
  __shared__ double sE[E5_MAXM * 9];
  __shared__ int sc[E5_MAXM];
  const E5Pair pr = tab[blockIdx.y];
  const int it = blockIdx.x;
  const int nm = pr.nmodels[it];
  if (threadIdx.x < E5_MAXM) { sc[threadIdx.x] = 0; pr.counts[it * E5_MAXM + threadIdx.x] = 0; }
  if (nm == 0) return;
  for (int k = threadIdx.x; k < nm * 9; k += blockDim.x) sE[k] = pr.models[(size_t)it * E5_MAXM * 9 + k];
  __syncthreads();
  int cnt[E5_MAXM];
#pragma unroll
  for (int m = 0; m < E5_MAXM; ++m) cnt[m] = 0;
  for (int i = threadIdx.x; i < pr.n; i += blockDim.x) {
    const double4 q = reinterpret_cast<const double4*>(pr.qn)[i];
#pragma unroll
    for (int m = 0; m < E5_MAXM; ++m)
      if (m < nm) cnt[m] += (sampson(sE + 9 * m, q.x, q.y, q.z, q.w) <= t2) ? 1 : 0;
  }
#pragma unroll
  for (int m = 0; m < E5_MAXM; ++m) {
    if (m < nm) {
      int v = cnt[m];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sc[m], v);
    }
  }
  __syncthreads();
  if (threadIdx.x < nm) pr.counts[it * E5_MAXM + threadIdx.x] = sc[threadIdx.x];
}

__global__ void e5b_replay_kernel(const E5Pair* __restrict__ tab, int npairs, int max_iters, double prob) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const E5Pair pr = tab[p];
  int niters = max_iters, best = 0, bi = -1, bm = -1, it = 0, total = 0;
  while (it < niters) {
    const int nm = pr.nmodels[it];
    total += nm;
    for (int m = 0; m < nm; ++m) {
      const int c = pr.counts[it * E5_MAXM + m];
      if (c > max(best, E5_MODEL_POINTS - 1)) {
        best = c; bi = it; bm = m;
        niters = update_num_iters(prob, (double)(pr.n - c) / pr.n, E5_MODEL_POINTS, niters);
      }
    }
    ++it;
  }
  E5Result* res = pr.res;
  res->ok = bi >= 0;
  res->best_iter = bi; res->best_model = bm; res->best_count = best; res->iters_run = it; res->models_total = total;
  for (int c = 0; c < 9; ++c) res->E[c] = bi >= 0 ? pr.models[((size_t)bi * E5_MAXM + bm) * 9 + c] : 0.0;
}

__global__ void e5b_mask_kernel(const E5Pair* __restrict__ tab, float t2) {
  const E5Pair pr = tab[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pr.n || pr.mask == nullptr) return;
  if (!pr.res->ok) { pr.mask[i] = 0; return; }
  const double4 q = reinterpret_cast<const double4*>(pr.qn)[i];
  pr.mask[i] = sampson(pr.res->E, q.x, q.y, q.z, q.w) <= t2 ? 1 : 0;
}

}  // namespace

extern "C" int sfm_find_essential_mat_batched(sfm_ctx* ctx, int npairs, const float* const* pts1, const float* const* pts2,
                                              const int32_t* n, const double* K, double prob, double threshold,
                                              int max_iters, double* E, uint8_t* const* masks, int32_t* info) {
  SFM_REQUIRE(ctx && npairs >= 0 && K && E && info && (npairs == 0 || (pts1 && pts2 && n)),
              "sfm_find_essential_mat_batched: null argument");
  SFM_REQUIRE(prob > 0.0 && prob < 1.0, "sfm_find_essential_mat_batched: confidence must lie in (0, 1)");
  SFM_REQUIRE(max_iters >= 1 && max_iters <= 100000, "sfm_find_essential_mat_batched: maxIters outside 1..100000");
  for (int k = 0; k < npairs; ++k) {
    for (int j = 0; j < 6; ++j) info[6 * k + j] = 0;
    info[6 * k + 3] = info[6 * k + 4] = -1;
    for (int j = 0; j < 9; ++j) E[9 * k + j] = 0.0;
    SFM_REQUIRE(n[k] >= 0, "sfm_find_essential_mat_batched: negative point count (pair %d)", k);
    SFM_REQUIRE(n[k] < 6 || (sfm_is_device_ptr(pts1[k]) && sfm_is_device_ptr(pts2[k])),
                "sfm_find_essential_mat_batched: point arrays must be device pointers (pair %d)", k);
  }
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const double thr = threshold / ((fx + fy) / 2);
  const float t2 = (float)(thr * thr);
  const int iters = max_iters;
  constexpr int GROUP = 48;                                  // pairs per launch group: ~0.8 MB of tables per pair
  for (int g0 = 0; g0 < npairs; g0 += GROUP) {
    // pairs of this group that take the RANSAC path (n >= 6; smaller ones are left to sfm_find_essential_mat)
    std::vector<int> idx;
    for (int k = g0; k < npairs && k < g0 + GROUP; ++k)
      if (n[k] >= 6) idx.push_back(k);
    const int P = (int)idx.size();
    if (P == 0) continue;
    SFM_TRY(sfm_ws_begin(ctx));
    std::vector<E5Pair> host((size_t)P);
    std::vector<int> hsubs((size_t)P * iters * 5);
    int maxn = 0;
    E5Result* res_all = nullptr;
    SFM_TRY(ws_alloc_t(ctx, (size_t)P, &res_all));
    int* subs_all = nullptr;
    SFM_TRY(ws_alloc_t(ctx, (size_t)P * iters * 5, &subs_all));
    for (int j = 0; j < P; ++j) {
      const int k = idx[j];
      E5Pair& pr = host[j];
      pr.p1 = pts1[k]; pr.p2 = pts2[k]; pr.n = n[k]; pr.pad = 0;
      maxn = n[k] > maxn ? n[k] : maxn;
      SFM_TRY(ws_alloc_t(ctx, (size_t)4 * n[k], &pr.qn));
      SFM_TRY(ws_alloc_t(ctx, (size_t)iters * E5_MAXM * 9, &pr.models));
      SFM_TRY(ws_alloc_t(ctx, (size_t)iters, &pr.nmodels));
      SFM_TRY(ws_alloc_t(ctx, (size_t)iters * E5_MAXM, &pr.counts));
      pr.subs = subs_all + (size_t)j * iters * 5;
      pr.res = res_all + j;
      pr.mask = nullptr;
      if (masks && masks[k]) {
        if (sfm_is_device_ptr(masks[k])) pr.mask = masks[k];
        else SFM_TRY(ws_alloc_t(ctx, (size_t)n[k], &pr.mask));
      }
      ransac_subsets(n[k], iters, hsubs.data() + (size_t)j * iters * 5);
    }
    E5Pair* tab = nullptr;
    SFM_TRY(ws_alloc_t(ctx, (size_t)P, &tab));
    // pageable sources: cudaMemcpyAsync returns after staging them, so the vectors may die at the end of the scope
    SFM_CUDA(cudaMemcpyAsync(tab, host.data(), sizeof(E5Pair) * (size_t)P, cudaMemcpyHostToDevice, ctx->stream));
    SFM_CUDA(cudaMemcpyAsync(subs_all, hsubs.data(), sizeof(int) * hsubs.size(), cudaMemcpyHostToDevice, ctx->stream));
    const dim3 gpt((unsigned)div_up(maxn, 256), (unsigned)P);
    SFM_LAUNCH(ctx, SFM_K_MISC, (e5b_normalize_kernel<<<gpt, 256, 0, ctx->stream>>>(tab, fx, fy, cx, cy)));
    SFM_LAUNCH(ctx, SFM_K_ESSENTIAL, (e5b_solve_kernel<<<dim3((unsigned)div_up(iters, E5_SOLVE_THREADS), (unsigned)P),
                                                        E5_SOLVE_THREADS, 0, ctx->stream>>>(tab, iters)));
    SFM_LAUNCH(ctx, SFM_K_ESSENTIAL, (e5b_score_kernel<<<dim3((unsigned)iters, (unsigned)P), 256, 0, ctx->stream>>>(tab, t2)));
    SFM_LAUNCH(ctx, SFM_K_MISC, (e5b_replay_kernel<<<div_up(P, 32), 32, 0, ctx->stream>>>(tab, P, iters, prob)));
    SFM_LAUNCH(ctx, SFM_K_MISC, (e5b_mask_kernel<<<gpt, 256, 0, ctx->stream>>>(tab, t2)));
    std::vector<E5Result> hres((size_t)P);
    SFM_CUDA(cudaMemcpyAsync(hres.data(), res_all, sizeof(E5Result) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    for (int j = 0; j < P; ++j) {
      const int k = idx[j];
      if (masks && masks[k] && !sfm_is_device_ptr(masks[k]))
        SFM_CUDA(cudaMemcpyAsync(masks[k], host[j].mask, (size_t)n[k], cudaMemcpyDeviceToHost, ctx->stream));
    }
    SFM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < P; ++j) {
      const int k = idx[j];
      const E5Result& r = hres[j];
      info[6 * k + 0] = r.ok ? 1 : 0;
      info[6 * k + 1] = r.ok ? r.best_count : 0;
      info[6 * k + 2] = r.iters_run;
      info[6 * k + 3] = r.best_iter;
      info[6 * k + 4] = r.best_model;
      info[6 * k + 5] = r.models_total;
      for (int c = 0; c < 9; ++c) E[9 * k + c] = r.E[c];
    }
  }
  return SFM_OK;
}
