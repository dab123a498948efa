// pnp_dev.cuh — declarations shared by pnp.cu (RANSAC scoring / replay / refinement) and pnp_epnp.cu (the minimal
// solver, a separate translation unit because it is compiled with -fmad=false).
#pragma once
#include "common.cuh"

struct PnpCam {
  double fx, fy, cx, cy;
};

// The RNG index stream depends only on N: for the default 100 iterations the host draws it (a few
// microseconds) and passes it by value in the kernel parameter block — no copy, no extra launch.
struct PnpSubsets {
  int count;            // iterations covered by idx (0: draw in the kernel)
  int idx[500];
};

// H minimal problems, one CTA each.  poses: (H,12) R (row-major 9) | t, the solver's output; valid: (H).
// n_dev / subs_dev: row count and subsets known to the device only.
// dbg (optional, >= 48 words): clock64 stamps of hypothesis 0 at the phase boundaries, sweeps, raw (R, t).
int sfm_pnp_epnp_launch(sfm_ctx* ctx, const float* X, const float* px, int n, int H, const PnpCam& cam, const PnpSubsets& subs,
                        double* poses, unsigned char* valid, long long* dbg, const int* n_dev,
                        const int* subs_dev);
