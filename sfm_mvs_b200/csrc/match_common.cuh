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

// Device storage of one view's descriptors; recycled through a per-context pool so that creating /
// destroying descriptor sets in a loop performs no cudaMalloc / cudaFree and no device synchronisation.
struct DescBuf {
  float* f32 = nullptr;            // (cap_rows, dim) float32 copy (fp32 kernel operand)
  unsigned char* tiles = nullptr;  // UMMA operand image (dim == 128 only)
  float* sqnorm = nullptr;         // (cap_rows) exact |d|^2
  unsigned int* flag = nullptr;    // device word: bit0 = a value is not an integer in [0,255], bit1 = |d|^2 > 2^21
  unsigned int* hflag = nullptr;   // pinned host copy of `flag`, valid once `ready` has completed
  cudaEvent_t ready = nullptr;
  size_t cap_rows = 0;             // multiple of 256
  int dim = 0;
};

struct sfm_desc {
  sfm_ctx* ctx = nullptr;
  int n = 0, dim = 0;
  DescBuf* buf = nullptr;
  bool resolved = false;     // hflag has been read
  bool exact = false;        // every value is an integer in [0,255] -> bf16/tensor path is exact
  float* f32 = nullptr;      // aliases into buf
  void* tiles = nullptr;     // non-null only when the tensor-core path may be used (after resolve)
  int n_tiles = 0;           // 128-row tiles
  float* sqnorm = nullptr;
};

int sfm_desc_resolve(sfm_desc* d);   // waits for the flag word (first use only)

// match_exact.cu
int sfm_match_exact_launch(sfm_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim,
                           mkey_t* cand, int nsplit);
int sfm_match_exact_splits(sfm_ctx* ctx, int nq, int nt);
// match_tc.cu
int sfm_match_tc_splits(sfm_ctx* ctx, int nq, int nt);
int sfm_match_tc_launch(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, mkey_t* cand, int nsplit);
int sfm_match_tc_launch_batched(sfm_ctx* ctx, int npairs, const sfm_desc* const* q, const sfm_desc* const* t,
                                mkey_t** cand_out, int* nsub_out);
int sfm_desc_prepare_launch(sfm_ctx* ctx, sfm_desc* d, const void* src, int dtype);
int sfm_desc_prepare_launch_batched(sfm_ctx* ctx, int count, sfm_desc* const* d, const void* const* src, int dtype,
                                    unsigned int* flags_dev);
// match.cu
// qf/tf non-null: candidates come from the tensor-core kernel ([nq][nsplit][3]); null: [nq][nsplit][2]
int sfm_match_finalize(sfm_ctx* ctx, const mkey_t* cand, int nq, int nt, int nsplit, double ratio,
                       const float* qf, const float* tf, int32_t* idx, float* dist, uint8_t* good, int32_t* n_good);
