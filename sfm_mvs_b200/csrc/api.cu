// api.cu — context lifetime, workspace, error reporting, profiling (C ABI in include/sfm_b200.h).
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void sfm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* sfm_last_error(void) { return g_err; }
extern "C" int sfm_version(void) { return SFM_B200_VERSION; }

static const char* kKernelNames[SFM_K_COUNT] = {
    "desc_prep",   "match_tc",  "match_exact", "match_final", "match_gather",
    "triangulate", "reproj",    "pnp_score",   "pnp_refine",  "common_points",
    "ba_eval",     "ba_schur",  "ba_update",   "ba_solve",    "misc",
    "pnp_epnp", "essential"};

extern "C" const char* sfm_kernel_name(int id) {
  return (id >= 0 && id < SFM_K_COUNT) ? kKernelNames[id] : "?";
}

extern "C" int sfm_ctx_create(int device, void* stream, sfm_ctx** out) {
  SFM_REQUIRE(out != nullptr, "sfm_ctx_create: out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    sfm_set_error("sfm_ctx_create: no CUDA device (%s); this engine has no CPU path",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return SFM_ERR_CUDA;
  }
  SFM_REQUIRE(device >= 0 && device < ndev, "sfm_ctx_create: device %d out of range [0,%d)", device, ndev);
  SFM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SFM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    sfm_set_error("sfm_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                  device, prop.major, prop.minor);
    return SFM_ERR_UNSUPPORTED;
  }
  sfm_ctx* c = new sfm_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    SFM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  SFM_CUDA(cudaMalloc(&c->counters, 64 * sizeof(unsigned int)));
  SFM_CUDA(cudaMemset(c->counters, 0, 64 * sizeof(unsigned int)));
  SFM_CUDA(cudaMalloc(&c->dscratch, 64 * sizeof(double)));
  SFM_CUDA(cudaMemset(c->dscratch, 0, 64 * sizeof(double)));
  *out = c;
  return SFM_OK;
}

extern "C" void sfm_ctx_destroy(sfm_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  sfm_desc_pool_free(c);
  sfm_chain_parked_free(c);
  for (void* p : c->retired) cudaFree(p);
  for (void* p : c->hs_retired) cudaFreeHost(p);
  if (c->ws) cudaFree(c->ws);
  if (c->hs) cudaFreeHost(c->hs);
  if (c->counters) cudaFree(c->counters);
  if (c->dscratch) cudaFree(c->dscratch);
  for (auto& ev : c->pending) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  for (auto& ev : c->pool) cudaEventDestroy(ev);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int sfm_ctx_sync(sfm_ctx* c) {
  SFM_REQUIRE(c, "sfm_ctx_sync: ctx is NULL");
  SFM_CUDA(cudaStreamSynchronize(c->stream));
  return SFM_OK;
}
extern "C" void* sfm_ctx_stream(sfm_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int sfm_ctx_detach_stream(sfm_ctx* c) {
  SFM_REQUIRE(c, "sfm_ctx_detach_stream: ctx is NULL");
  c->own_stream = false;
  return SFM_OK;
}
extern "C" int sfm_ctx_sm_count(sfm_ctx* c) { return c ? c->sm_count : 0; }

// ---------------------------------------------------------------------------- workspace
int sfm_ws_begin(sfm_ctx* c) {
  SFM_CUDA(cudaSetDevice(c->device));
  if (!c->retired.empty() || !c->hs_retired.empty()) {
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    for (void* p : c->retired) cudaFree(p);
    for (void* p : c->hs_retired) cudaFreeHost(p);
    c->retired.clear();
    c->hs_retired.clear();
  }
  c->ws_off = 0;
  c->hs_off = 0;
  return SFM_OK;
}

int sfm_ws_alloc(sfm_ctx* c, size_t bytes, void** out) {
  size_t need = (bytes + 255) & ~(size_t)255;
  if (c->ws_off + need > c->ws_cap) {
    size_t cap = c->ws_cap ? c->ws_cap : ((size_t)8 << 20);
    while (cap < need) cap *= 2;
    if (c->ws_cap && cap < 2 * c->ws_cap) cap = 2 * c->ws_cap;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, cap);
    if (e != cudaSuccess) {
      sfm_set_error("workspace cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
      return SFM_ERR_NOMEM;
    }
    if (c->ws) c->retired.push_back(c->ws);
    c->ws = (char*)p;
    c->ws_cap = cap;
    c->ws_off = 0;
  }
  *out = c->ws + c->ws_off;
  c->ws_off += need;
  return SFM_OK;
}

int sfm_hs_alloc(sfm_ctx* c, size_t bytes, void** out) {
  size_t need = (bytes + 63) & ~(size_t)63;
  if (c->hs_off + need > c->hs_cap) {
    size_t cap = c->hs_cap ? c->hs_cap : ((size_t)1 << 20);
    while (cap < need) cap *= 2;
    if (c->hs_cap && cap < 2 * c->hs_cap) cap = 2 * c->hs_cap;
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, cap);
    if (e != cudaSuccess) {
      sfm_set_error("pinned cudaMallocHost(%zu) failed: %s", cap, cudaGetErrorString(e));
      return SFM_ERR_NOMEM;
    }
    if (c->hs) c->hs_retired.push_back(c->hs);
    c->hs = (char*)p;
    c->hs_cap = cap;
    c->hs_off = 0;
  }
  *out = c->hs + c->hs_off;
  c->hs_off += need;
  return SFM_OK;
}

bool sfm_is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Pinned (page-locked, UVA-mapped) host memory: kernels can store to it directly.
bool sfm_is_pinned_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost && a.devicePointer != nullptr;
}

// ---------------------------------------------------------------------------- profiling
static int drain_events(sfm_ctx* c) {
  if (c->pending.empty()) return SFM_OK;
  SFM_CUDA(cudaStreamSynchronize(c->stream));
  for (auto& ev : c->pending) {
    float ms = 0.f;
    SFM_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    c->ms[ev.id] += ms;
    c->pool.push_back(ev.a);
    c->pool.push_back(ev.b);
  }
  c->pending.clear();
  return SFM_OK;
}

static int get_event(sfm_ctx* c, cudaEvent_t* e) {
  if (!c->pool.empty()) { *e = c->pool.back(); c->pool.pop_back(); return SFM_OK; }
  SFM_CUDA(cudaEventCreate(e));
  return SFM_OK;
}

int sfm_launch_begin(sfm_ctx* c, int id) {
  c->launches[id]++;
  c->total_launches++;
  if (c->profiling) {
    if (c->pending.size() > 4096) SFM_TRY(drain_events(c));
    SFM_TRY(get_event(c, &c->cur_a));
    SFM_TRY(get_event(c, &c->cur_b));
    SFM_CUDA(cudaEventRecord(c->cur_a, c->stream));
  }
  return SFM_OK;
}

int sfm_launch_end(sfm_ctx* c, int id) {
  if (c->profiling) {
    SFM_CUDA(cudaEventRecord(c->cur_b, c->stream));
    c->pending.push_back({c->cur_a, c->cur_b, id});
  }
  return SFM_OK;
}

extern "C" int sfm_ctx_set_profiling(sfm_ctx* c, int on) {
  SFM_REQUIRE(c, "ctx is NULL");
  SFM_TRY(drain_events(c));
  c->profiling = on != 0;
  return SFM_OK;
}

extern "C" int sfm_ctx_reset_profile(sfm_ctx* c) {
  SFM_REQUIRE(c, "ctx is NULL");
  SFM_TRY(drain_events(c));
  for (int i = 0; i < SFM_K_COUNT; ++i) { c->ms[i] = 0; c->launches[i] = 0; }
  return SFM_OK;
}

extern "C" int sfm_ctx_get_profile(sfm_ctx* c, int id, double* ms_total, int64_t* launches) {
  SFM_REQUIRE(c && id >= 0 && id < SFM_K_COUNT, "bad kernel id %d", id);
  SFM_TRY(drain_events(c));
  if (ms_total) *ms_total = c->ms[id];
  if (launches) *launches = c->launches[id];
  return SFM_OK;
}

extern "C" int64_t sfm_ctx_launch_count(sfm_ctx* c) { return c ? c->total_launches : 0; }

// A side context (the registration loop's association / output streams) hands its launch counts and, when
// profiling, its per-kernel event times to the context the caller sees.
void sfm_ctx_merge_profile(sfm_ctx* dst, sfm_ctx* src) {
  if (!dst || !src) return;
  drain_events(src);
  for (int i = 0; i < SFM_K_COUNT; ++i) {
    dst->ms[i] += src->ms[i];
    dst->launches[i] += src->launches[i];
    src->ms[i] = 0;
    src->launches[i] = 0;
  }
  dst->total_launches += src->total_launches;
  src->total_launches = 0;
}

// n host -> device copies queued on `stream` in one call: the upload of a chunk of views (keypoints and descriptors of
// every view are separate host arrays) costs the host one native loop instead of one interpreter round trip per array.
extern "C" int sfm_upload_batch(sfm_ctx* c, void* stream, int n, void* const* dst, const void* const* src, const int64_t* bytes) {
  SFM_REQUIRE(c && (n == 0 || (dst && src && bytes)), "sfm_upload_batch: null argument");
  SFM_CUDA(cudaSetDevice(c->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  for (int i = 0; i < n; ++i) {
    SFM_REQUIRE(bytes[i] >= 0 && (bytes[i] == 0 || (dst[i] && src[i])), "sfm_upload_batch: bad entry %d", i);
    if (bytes[i]) SFM_CUDA(cudaMemcpyAsync(dst[i], src[i], (size_t)bytes[i], cudaMemcpyHostToDevice, s));
  }
  return SFM_OK;
}
