// match_common.cuh — shared pieces of the 2-NN matching kernels.
#pragma once
#include "common.cuh"

// A candidate is a 64-bit key: (float bits of squared distance) << 32 | train index.  Squared
// distances are non-negative, so unsigned integer order on the key is (distance, index) lexicographic
// order — exactly OpenCV's tie rule (equal distance -> lower trainIdx first).
typedef unsigned long long mkey_t;
#define MKEY_INF 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ mkey_t make_key(float d2, int idx) {
  return ((mkey_t)__float_as_uint(d2) << 32) | (unsigned int)idx;
}
__device__ __forceinline__ void key_insert(mkey_t k, mkey_t& k1, mkey_t& k2) {
  if (k < k2) {
    if (k < k1) { k2 = k1; k1 = k; } else { k2 = k; }
  }
}

struct sfm_desc {
  sfm_ctx* ctx = nullptr;
  int n = 0, dim = 0;
  bool exact = false;        // every value is an integer in [0,255] -> bf16/tensor path is exact
  float* f32 = nullptr;      // (n, dim) row-major copy for the fp32 kernel (always present)
  void* tiles = nullptr;     // UMMA operand image: [n_tiles][TILE bytes] (only when exact && dim==128)
  int n_tiles = 0;           // 128-row tiles
  float* sqnorm = nullptr;   // (n_pad) exact |d|^2 as float (only when exact)
  unsigned int* flag = nullptr;  // device word: non-zero if any value is not an integer in [0,255]
};

// match_exact.cu
int sfm_match_exact_launch(sfm_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim,
                           mkey_t* cand, int nsplit);
int sfm_match_exact_splits(sfm_ctx* ctx, int nq, int nt);
// match_tc.cu
int sfm_match_tc_splits(sfm_ctx* ctx, int nq, int nt);
int sfm_match_tc_launch(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, mkey_t* cand, int nsplit);
int sfm_desc_prepare_tiles(sfm_ctx* ctx, sfm_desc* d);
// match.cu
int sfm_match_finalize(sfm_ctx* ctx, const mkey_t* cand, int nq, int nt, int nsplit, double ratio,
                       int32_t* idx, float* dist, uint8_t* good, int32_t* n_good);
