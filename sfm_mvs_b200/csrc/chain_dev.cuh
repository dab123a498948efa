// chain_dev.cuh — internal entry points of the synchronisation-free registration loop (chain.cu):
// the same kernels as the public calls, but with the row counts and the camera matrices read from
// DEVICE memory (written by the kernels that produced them), so that a whole view sequence is
// enqueued without the host ever waiting for a count or a pose.  n_cap bounds the launch grid.
#pragma once
#include "common.cuh"

struct CamParams {            // cv2.projectPoints operands: R' = Rodrigues(Rodrigues(R)) as the reference round-trips it
  double R[9];
  double t[3];
  double fx, fy, cx, cy;
};

// geometry.cu
// idx (optional): output row j is the point of input row idx[j]
int sfm_triangulate_dev(sfm_ctx* ctx, const double* P1P2_dev /*24 doubles*/, const float* x1, const float* x2, int n_cap,
                        const int* n_dev, float* X, int out_layout /*1: (N,4), 2: (N,3)*/, const int32_t* idx = nullptr);
// pnp.cu: the RNG index stream (100 x 5) for a device-side row count; needs no pose, so the loop runs it early
int sfm_pnp_subsets_dev(sfm_ctx* ctx, const int* n_dev, int32_t* subs_dev /*500*/);
int sfm_reproj_error_dev(sfm_ctx* ctx, const float* X, int x_layout /*0: (N,3), 2: (N,4)*/, const float* px, int n_cap,
                         const int* n_dev, const CamParams* cam_dev, double* err_dev, float* X3);
// chain.cu
int sfm_gather_rows_dev(sfm_ctx* ctx, const float* src, int width, const int32_t* idx, int n_cap, const int* n_dev, float* dst);
// pnp.cu: cv2.solvePnPRansac defaults; everything stays in HBM.  pose6_dev: rvec|tvec (refined).
// Rt_dev / P_dev / cam_dev (all or none): [R|t], K[R|t] and the projectPoints operands of the solved view, formed
// by the last kernel of the call.
int sfm_pnp_ransac_dev(sfm_ctx* ctx, const float* X, const float* px, int n_cap, const int* n_dev, const double* K,
                       const double* K_dev, double* pose6_dev, int32_t* inliers_dev, int32_t* n_inl_dev, int32_t* ok_dev,
                       double* Rt_dev, double* P_dev, CamParams* cam_dev, const int32_t* subs_dev = nullptr);
