// common.cuh — context, workspace, error and launch plumbing shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/sfm_b200.h"

void sfm_set_error(const char* fmt, ...);

#define SFM_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess) {                                                            \
      sfm_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return SFM_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define SFM_REQUIRE(cond, ...)         \
  do {                                 \
    if (!(cond)) {                     \
      sfm_set_error(__VA_ARGS__);      \
      return SFM_ERR_INVALID;          \
    }                                  \
  } while (0)

#define SFM_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != SFM_OK) return _s; \
  } while (0)

struct sfm_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // grow-only device workspace, bump-allocated per API call
  char* ws = nullptr;
  size_t ws_cap = 0, ws_off = 0;
  std::vector<void*> retired;       // old blocks kept alive until the next call begins
  // grow-only pinned host staging
  char* hs = nullptr;
  size_t hs_cap = 0, hs_off = 0;
  std::vector<void*> hs_retired;
  // profiling
  bool profiling = false;
  struct Ev { cudaEvent_t a, b; int id; };
  std::vector<Ev> pending;
  std::vector<cudaEvent_t> pool;
  double ms[SFM_K_COUNT] = {0};
  int64_t launches[SFM_K_COUNT] = {0};
  int64_t total_launches = 0;
  cudaEvent_t cur_a = nullptr, cur_b = nullptr;
  // persistent zero-initialised device words for last-block-done reductions (self-resetting)
  unsigned int* counters = nullptr;
  double* dscratch = nullptr;       // 64 doubles of persistent device scratch
  std::vector<void*> desc_pool;     // recycled DescBuf storage (match.cu)
  void* chain_parked = nullptr;     // a destroyed registration chain's buffers / second context, kept for the next
                                    // sfm_chain_create (pinned allocations and context creation cost milliseconds)
};

void sfm_desc_pool_free(sfm_ctx* c);   // match.cu
void sfm_chain_parked_free(sfm_ctx* c);   // chain.cu
void sfm_ctx_merge_profile(sfm_ctx* dst, sfm_ctx* src);   // api.cu

// ---- workspace -------------------------------------------------------------------------
int sfm_ws_begin(sfm_ctx* c);                         // start of an API call: reset bump pointers
int sfm_ws_alloc(sfm_ctx* c, size_t bytes, void** out);   // 256-byte aligned device scratch
int sfm_hs_alloc(sfm_ctx* c, size_t bytes, void** out);   // pinned host scratch

template <typename T>
static inline int ws_alloc_t(sfm_ctx* c, size_t count, T** out) {
  void* p = nullptr;
  int s = sfm_ws_alloc(c, count * sizeof(T), &p);
  *out = (T*)p;
  return s;
}
template <typename T>
static inline int hs_alloc_t(sfm_ctx* c, size_t count, T** out) {
  void* p = nullptr;
  int s = sfm_hs_alloc(c, count * sizeof(T), &p);
  *out = (T*)p;
  return s;
}

bool sfm_is_device_ptr(const void* p);
bool sfm_is_pinned_ptr(const void* p);

// An input that must be readable by kernels: device pointers pass through, host buffers are
// copied into the workspace on the ctx stream.
template <typename T>
static inline int dev_in(sfm_ctx* c, const T* p, size_t count, const T** out) {
  if (p == nullptr || count == 0) { *out = p; return SFM_OK; }
  if (sfm_is_device_ptr(p)) { *out = p; return SFM_OK; }
  T* d = nullptr;
  SFM_TRY(ws_alloc_t(c, count, &d));
  SFM_CUDA(cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  *out = d;
  return SFM_OK;
}

// An output: device pointers are written in place; host buffers get a workspace twin that
// dev_out_finish copies back.  `any_host` tells the caller whether it must synchronise.
template <typename T>
struct DevOut {
  T* dev = nullptr;
  T* host = nullptr;
  size_t count = 0;
};
template <typename T>
static inline int dev_out(sfm_ctx* c, T* p, size_t count, DevOut<T>* o, bool* any_host) {
  o->count = count;
  if (p == nullptr) return SFM_OK;
  if (count == 0) { o->dev = p; return SFM_OK; }
  if (sfm_is_device_ptr(p)) { o->dev = p; return SFM_OK; }
  o->host = p;
  *any_host = true;
  return ws_alloc_t(c, count, &o->dev);
}
template <typename T>
static inline int dev_out_finish(sfm_ctx* c, DevOut<T>* o, size_t count = (size_t)-1) {
  if (o->host == nullptr || o->dev == nullptr) return SFM_OK;
  size_t n = count == (size_t)-1 ? o->count : count;
  if (n == 0) return SFM_OK;
  SFM_CUDA(cudaMemcpyAsync(o->host, o->dev, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
  return SFM_OK;
}

// ---- launch bookkeeping ----------------------------------------------------------------
int sfm_launch_begin(sfm_ctx* c, int kernel_id);
int sfm_launch_end(sfm_ctx* c, int kernel_id);

#define SFM_LAUNCH(ctx, id, ...)                 \
  do {                                           \
    SFM_TRY(sfm_launch_begin((ctx), (id)));      \
    __VA_ARGS__;                                 \
    SFM_CUDA(cudaGetLastError());                \
    SFM_TRY(sfm_launch_end((ctx), (id)));        \
  } while (0)

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline int64_t div_up64(int64_t a, int64_t b) { return (a + b - 1) / b; }
