// pnp_dev.cuh — declarations shared by pnp.cu (RANSAC scoring / replay / refinement) and pnp_epnp.cu (the minimal
// solver, a separate translation unit because it is compiled with -fmad=false).
#pragma once
#include "common.cuh"

struct PnpCam {
  double fx, fy, cx, cy;
};

// cv2.projectPoints with zero distortion, operation for operation (see geometry.cu project_cv).  Shared by the
// scoring kernel (pnp.cu) and the minimal solver's own scoring pass (pnp_epnp.cu): explicit roundings, so the result
// does not depend on the translation unit's -fmad setting.
__device__ __forceinline__ void project_pose(const double* __restrict__ P /*R row-major 9 | t 3*/, const PnpCam& c,
                                             double X, double Y, double Z, double& u, double& v) {
  double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[0], X), __dmul_rn(P[1], Y)), __dmul_rn(P[2], Z)), P[9]);
  double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[3], X), __dmul_rn(P[4], Y)), __dmul_rn(P[5], Z)), P[10]);
  double z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[6], X), __dmul_rn(P[7], Y)), __dmul_rn(P[8], Z)), P[11]);
  z = (z != 0.0) ? __ddiv_rn(1.0, z) : 1.0;
  x = __dmul_rn(x, z);
  y = __dmul_rn(y, z);
  u = __dadd_rn(__dmul_rn(x, c.fx), c.cx);
  v = __dadd_rn(__dmul_rn(y, c.fy), c.cy);
}

// PnPRansacCallback::computeError + findInliers: projection in float64 stored as float32,
// err = dx*dx + dy*dy in float32 with separately rounded products, inlier iff err <= thr^2.
__device__ __forceinline__ bool is_inlier(const double* __restrict__ P, const PnpCam& c, float X, float Y, float Z,
                                          float ox, float oy, float thr2) {
  double u, v;
  project_pose(P, c, (double)X, (double)Y, (double)Z, u, v);
  float dx = __fsub_rn(ox, (float)u), dy = __fsub_rn(oy, (float)v);
  float e = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  return e <= thr2;
}

// The RNG index stream depends only on N: for the default 100 iterations the host draws it (a few
// microseconds) and passes it by value in the kernel parameter block — no copy, no extra launch.
struct PnpSubsets {
  int count;            // iterations covered by idx (0: draw in the kernel)
  int idx[500];
};

// H minimal problems, one CTA each.  poses: (H,12) R (row-major 9) | t, the solver's output; valid: (H).
// n_dev / subs_dev: row count and subsets known to the device only.
// counts (optional, H): every hypothesis CTA also SCORES its pose over all n points (K4's test, squared threshold
// thr2) and writes its inlier count — the registration loop then needs no separate scoring launch.
// dbg (optional, >= 48 words): clock64 stamps of hypothesis 0 at the phase boundaries, sweeps, raw (R, t).
int sfm_pnp_epnp_launch(sfm_ctx* ctx, const float* X, const float* px, int n, int H, const PnpCam& cam, const PnpSubsets& subs,
                        double* poses, unsigned char* valid, long long* dbg, const int* n_dev,
                        const int* subs_dev, int* counts, float thr2);
