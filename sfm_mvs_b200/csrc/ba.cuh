// ba.cuh — the bundle-adjustment problem object shared by ba.cu (kernels) and comm.cu (C1 exchange).
#pragma once
#include <algorithm>

#include "common.cuh"

struct sfm_ba {
  sfm_ctx* ctx = nullptr;
  int n_cam = 0, n_pt = 0, n_obs = 0;          // this rank's shard: all cameras, its points/observations
  int64_t n_pt_total = 0, n_obs_total = 0;     // whole problem (for the reference's 1/N residual modes)
  double K[9] = {0};
  // observations (point-major) and CSR offsets
  float2* uv = nullptr;
  int* cam_idx = nullptr;
  int* pt_idx = nullptr;
  int* pt_start = nullptr;
  int max_deg = 0;                             // most observations of one point (sizes the per-warp scratch of the fused kernels)
  // parameters and step candidates
  double *cams = nullptr, *cams_new = nullptr, *pts = nullptr, *pts_new = nullptr;
  double* cam_pre = nullptr;                   // [C] 144-byte records: R, t float64 | Jl float32
  // reduced camera system: one allocation S | g | hdiag (so one memset and one all-reduce cover it)
  float* S = nullptr;
  float* g = nullptr;
  float* hdiag = nullptr;
  size_t sys_count = 0;
  double* A64 = nullptr;                       // float64 copy factored in place
  double* dc = nullptr;                        // camera step (6C)
  float* Tbuf = nullptr;                       // [O][18] T_a = W_a Hpp^-1 of the current linearisation (Schur kernel -> back substitution)
  double* qp = nullptr;                        // [P][3] -Hpp^-1 bp of the current linearisation
  double* scal = nullptr;                      // [0] cost at linearisation, [1] cost at candidate, [2] |dp|^2, [3] |dc|^2
  int* info = nullptr;                         // solve status: [0] Cholesky info (0 = ok), [1] 1 = the CG solver produced dc, [2] its iterations
  double* pcg = nullptr;                       // scratch of the conjugate-gradient solver (pcg.cu), null when it does not apply
  // C1 exchange
  void* comm = nullptr;                        // ncclComm_t
  int rank = 0, world = 1;
};

// comm.cu
int sfm_ba_allreduce_system(sfm_ba* ba);       // sum over ranks of S|g|hdiag (float32) and scal[0]
int sfm_ba_allreduce_scalars(sfm_ba* ba);      // sum over ranks of scal[1..2]
