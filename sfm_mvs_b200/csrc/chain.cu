// chain.cu — device-side glue of the per-view registration loop (sfm.py:341-409): the fancy
// indexing the reference does in NumPy between its cv2 calls, kept on the GPU so that matched
// keypoints, 3-D points and masks never travel to the host.
//   pts2[indx2], points_3d[indx1]                     sfm.py:358-362   -> sfm_gather_rows
//   temp_array1/2 = pts2[~mask], pts3[~mask]          sfm.py:229-237   -> sfm_compact_pairs
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "chain_dev.cuh"
#include "hostmath.h"

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, int width,
                                                          const int* __restrict__ idx, int n, float* __restrict__ dst,
                                                          const int* __restrict__ n_dev = nullptr) {
  if (n_dev) n = min(n, *n_dev);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  int r = i / width, c = i - r * width;
  dst[i] = __ldg(src + (size_t)__ldg(idx + r) * width + c);
}

extern "C" int sfm_gather_rows(sfm_ctx* ctx, const float* src, int width, const int32_t* idx, int n, float* dst) {
  SFM_REQUIRE(ctx && width >= 1 && n >= 0, "sfm_gather_rows: bad argument");
  if (n == 0) return SFM_OK;
  SFM_REQUIRE(src && idx && dst, "sfm_gather_rows: null buffer");
  SFM_REQUIRE(sfm_is_device_ptr(src) && sfm_is_device_ptr(idx) && sfm_is_device_ptr(dst),
              "sfm_gather_rows: device pointers only (host arrays are indexed by the caller)");
  SFM_LAUNCH(ctx, SFM_K_GATHER, (gather_rows_kernel<<<div_up(n * width, 256), 256, 0, ctx->stream>>>(src, width, idx, n, dst)));
  return SFM_OK;
}

int sfm_gather_rows_dev(sfm_ctx* ctx, const float* src, int width, const int32_t* idx, int n_cap, const int* n_dev, float* dst) {
  if (n_cap <= 0) return SFM_OK;
  SFM_LAUNCH(ctx, SFM_K_GATHER, (gather_rows_kernel<<<div_up(n_cap * width, 256), 256, 0, ctx->stream>>>(src, width, idx, n_cap, dst, n_dev)));
  return SFM_OK;
}

// Stable compaction of the rows of two (n,2) arrays where keep[i] != 0 (single CTA: n is a few
// thousand keypoints).
__global__ void __launch_bounds__(1024) compact_pairs_kernel(const float2* __restrict__ a, const float2* __restrict__ b,
                                                             const unsigned char* __restrict__ keep, int n,
                                                             float2* __restrict__ a_out, float2* __restrict__ b_out,
                                                             int* __restrict__ n_out) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int start = 0; start < n; start += 1024) {
    int i = start + threadIdx.x;
    bool f = (i < n) && keep[i];
    unsigned m = __ballot_sync(0xffffffffu, f);
    int pre = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[w] = __popc(m);
    __syncthreads();
    int off = 0;
    for (int k = 0; k < w; ++k) off += warp_tot[k];
    int base = base_s;
    if (f) {
      int pos = base + off + pre;
      if (a_out) a_out[pos] = a[i];
      if (b_out) b_out[pos] = b[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int k = 0; k < 32; ++k) tot += warp_tot[k];
      base_s = base + tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && n_out) *n_out = base_s;
}

extern "C" int sfm_compact_pairs(sfm_ctx* ctx, const float* a, const float* b, const uint8_t* keep, int n,
                                 float* a_out, float* b_out, int32_t* n_out) {
  SFM_REQUIRE(ctx && n >= 0, "sfm_compact_pairs: bad argument");
  SFM_REQUIRE(n == 0 || (keep && sfm_is_device_ptr(keep)), "sfm_compact_pairs: device pointers only");
  SFM_REQUIRE((!a_out || a) && (!b_out || b), "sfm_compact_pairs: source missing");
  SFM_TRY(sfm_ws_begin(ctx));
  bool host_out = false;
  DevOut<int32_t> on;
  SFM_TRY(dev_out(ctx, n_out, 1, &on, &host_out));
  SFM_LAUNCH(ctx, SFM_K_GATHER, (compact_pairs_kernel<<<1, 1024, 0, ctx->stream>>>(
                                    (const float2*)a, (const float2*)b, keep, n, (float2*)a_out, (float2*)b_out, on.dev)));
  SFM_TRY(dev_out_finish(ctx, &on));
  if (host_out) SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFM_OK;
}

// ============================================================================ the per-view loop, native
// sfm.py:341-409 (imread / SIFT / GUI removed) for a whole view sequence in ONE C call: the host side of
// the loop — sizing, launching, reading back the two association counts and the PnP pose each view
// needs before it can size the next launches — runs here instead of in the Python caller, so the GPU
// is not left idle between a view's ~14 launches while an interpreter marshals arguments.
// Matches of pair k = (view k, view k+1) are device arrays pts_q[k], pts_t[k] (n_match[k] rows,
// ascending queryIdx: the output of sfm_desc_match_gather_batched).
namespace {

struct Arena {                 // chain-owned device scratch, sized for the largest pair
  char* base = nullptr;
  size_t cap = 0, off = 0;
  template <typename T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    if (off + bytes > cap) return nullptr;
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
};

void matmul_K_Rt(const double* K, const double* Rt, double* P) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) P[4 * i + j] = K[3 * i] * Rt[j] + K[3 * i + 1] * Rt[4 + j] + K[3 * i + 2] * Rt[8 + j];
}

}  // namespace

struct sfm_chain {
  sfm_ctx* ctx = nullptr;
  double K[9];
  double P1[12], P2[12], Rt1[12];
  int nmax = 0;
  Arena ar;
  float *X4 = nullptr, *pts3d_a = nullptr, *boot_pts1 = nullptr, *boot_p3d = nullptr, *temp1 = nullptr, *temp2 = nullptr,
        *Xc = nullptr, *com2 = nullptr, *X_in = nullptr, *p_in = nullptr;
  int32_t *i1 = nullptr, *i2 = nullptr, *inl = nullptr, *cnt = nullptr;
  uint8_t* keep = nullptr;
  double* errs = nullptr;          // device: 2 per call slot, ERR_SLOTS slots
  // synchronisation-free path: everything a view produces stays in HBM until the end of the call
  struct ViewRec* recs = nullptr;  // [ERR_SLOTS]   per registered view of a call
  double* P_view = nullptr;        // [ERR_SLOTS+2][12]  projection matrix K[R|t] of every view so far
  CamParams* cams = nullptr;       // [ERR_SLOTS]   projectPoints operands of the view registered in that slot
  double* pose6 = nullptr;         // rvec | tvec of the PnP just run
  double* K_dev = nullptr;
  // Data association of view v+1 (common_points, complement, pixel gather) needs no pose, so it runs on a
  // second stream (own context = own workspace) while view v is inside PnP: second set of its outputs + events.
  sfm_ctx* ctxB = nullptr;
  // The outputs of a view that nothing later depends on (its two reprojection errors, its new 3-D points) run on
  // a third stream/context after the pose is known; the main stream only carries the pose dependency:
  // re-triangulation -> PnP.  Buffers both sides touch exist twice (view parity).
  sfm_ctx* ctxC = nullptr;
  float* Xcb = nullptr;
  int32_t* inlb = nullptr;
  int32_t* subs[2] = {nullptr, nullptr};   // RNG subsets of the PnP (needs only the association count: stream B)
  cudaEvent_t ev_core[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  int32_t *i1b = nullptr, *i2b = nullptr;
  uint8_t* keepb = nullptr;
  float *temp1b = nullptr, *temp2b = nullptr, *com2b = nullptr;
  cudaEvent_t ev_assoc[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_start = nullptr;
  bool set_used[2] = {false, false};
  struct ViewRec* hrecs = nullptr; // pinned
  int32_t* hcnt = nullptr;         // pinned
  double* herrs = nullptr;         // pinned
  // a call queued by sfm_chain_extend_async and not collected yet
  bool pending = false, pending_parsed = false;
  int pend_reg = 0;
  std::vector<int32_t> pend_nmatch;
  std::vector<sfm_view_out> pend_out;
  // loop state (sfm.py:399-409)
  bool started = false;
  bool dead = false;                // a view failed to register: later views were computed from an undefined pose
  const float* prev_q = nullptr;   // previous pair's matches (re-triangulated for the next view)
  const float* prev_t = nullptr;
  int prev_n = 0;
  const float* pts1 = nullptr;
  const float* points_3d = nullptr;
  int n1 = 0;
  int views_done = 0;
};
constexpr int ERR_SLOTS = 4096;

struct ViewRec {                  // device-side result record of one registered view
  double Rt[12];
  double err_pnp, err_new;
  int n_new, n_pnp, n_inl, ok;
};

static void sfm_chain_release(sfm_chain* c);

extern "C" int sfm_chain_create(sfm_ctx* ctx, const double* K, const double* Rt0, const double* Rt1, int max_matches,
                                sfm_chain** out) {
  SFM_REQUIRE(ctx && K && Rt0 && Rt1 && out && max_matches >= 6, "sfm_chain_create: bad argument");
  if (ctx->chain_parked) {                               // reuse the buffers of the last chain on this context
    sfm_chain* p = static_cast<sfm_chain*>(ctx->chain_parked);
    ctx->chain_parked = nullptr;
    if (p->nmax >= max_matches) {
      memcpy(p->K, K, sizeof(p->K));
      memcpy(p->Rt1, Rt1, sizeof(p->Rt1));
      matmul_K_Rt(K, Rt0, p->P1);
      matmul_K_Rt(K, Rt1, p->P2);
      p->started = false; p->prev_q = p->prev_t = nullptr; p->prev_n = 0; p->pts1 = p->points_3d = nullptr; p->n1 = 0;
      p->views_done = 0; p->set_used[0] = p->set_used[1] = false;
      p->pending = p->pending_parsed = false; p->pend_reg = 0;     // an uncollected call died with its chain
      p->dead = false;
      SFM_CUDA(cudaMemcpyAsync(p->K_dev, p->K, 9 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      SFM_CUDA(cudaMemcpyAsync(p->P_view, p->P1, 12 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      SFM_CUDA(cudaMemcpyAsync(p->P_view + 12, p->P2, 12 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      *out = p;
      return SFM_OK;
    }
    p->ctx->chain_parked = nullptr;
    sfm_chain_release(p);
  }
  sfm_chain* c = new sfm_chain();
  c->ctx = ctx;
  memcpy(c->K, K, sizeof(c->K));
  memcpy(c->Rt1, Rt1, sizeof(c->Rt1));
  matmul_K_Rt(K, Rt0, c->P1);
  matmul_K_Rt(K, Rt1, c->P2);
  const int nmax = c->nmax = max_matches;
  c->ar.cap = (size_t)nmax * 220 + (size_t)ERR_SLOTS * (16 + sizeof(ViewRec) + 96 + sizeof(CamParams)) + 64 * 1024;
  cudaError_t e = cudaMalloc(&c->ar.base, c->ar.cap);
  if (e == cudaSuccess) e = cudaMallocHost(&c->hcnt, 4 * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMallocHost(&c->herrs, sizeof(double) * 2 * ERR_SLOTS);
  if (e == cudaSuccess) e = cudaMallocHost(&c->hrecs, sizeof(ViewRec) * ERR_SLOTS);
  if (e != cudaSuccess) {
    sfm_set_error("sfm_chain_create: %s", cudaGetErrorString(e));
    sfm_chain_destroy(c);
    return SFM_ERR_NOMEM;
  }
  Arena& ar = c->ar;
  c->X4 = ar.take<float>((size_t)4 * nmax);          // triangulated points (N,4)
  c->pts3d_a = ar.take<float>((size_t)3 * nmax);     // points_3d of the previous pair
  c->boot_pts1 = ar.take<float>((size_t)2 * nmax);
  c->boot_p3d = ar.take<float>((size_t)3 * nmax);
  c->i1 = ar.take<int32_t>(nmax);
  c->i2 = ar.take<int32_t>(nmax);
  c->keep = ar.take<uint8_t>(nmax);
  c->temp1 = ar.take<float>((size_t)2 * nmax);
  c->temp2 = ar.take<float>((size_t)2 * nmax);
  c->Xc = ar.take<float>((size_t)3 * nmax);
  c->com2 = ar.take<float>((size_t)2 * nmax);
  c->X_in = ar.take<float>((size_t)3 * nmax);
  c->p_in = ar.take<float>((size_t)2 * nmax);
  c->inl = ar.take<int32_t>(nmax);
  c->cnt = ar.take<int32_t>(4);
  c->errs = ar.take<double>((size_t)2 * ERR_SLOTS);
  c->recs = ar.take<ViewRec>(ERR_SLOTS);
  c->P_view = ar.take<double>((size_t)12 * (ERR_SLOTS + 2));
  c->cams = ar.take<CamParams>(ERR_SLOTS);
  c->pose6 = ar.take<double>(16);
  c->i1b = ar.take<int32_t>(nmax);
  c->i2b = ar.take<int32_t>(nmax);
  c->keepb = ar.take<uint8_t>(nmax);
  c->temp1b = ar.take<float>((size_t)2 * nmax);
  c->temp2b = ar.take<float>((size_t)2 * nmax);
  c->com2b = ar.take<float>((size_t)2 * nmax);
  c->Xcb = ar.take<float>((size_t)3 * nmax);
  c->inlb = ar.take<int32_t>(nmax);
  c->subs[0] = ar.take<int32_t>(512);
  c->subs[1] = ar.take<int32_t>(512);
  if (sfm_ctx_create(ctx->device, nullptr, &c->ctxB) != SFM_OK) c->ctxB = nullptr;
  if (sfm_ctx_create(ctx->device, nullptr, &c->ctxC) != SFM_OK) c->ctxC = nullptr;
  for (int k = 0; k < 2; ++k) {
    cudaEventCreateWithFlags(&c->ev_assoc[k], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_core[k], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_out[k], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming);
  double* Kdev = c->inlb && c->ctxB && c->ctxC ? ar.take<double>(16) : nullptr;
  if (Kdev) {
    c->K_dev = Kdev;
    cudaMemcpy(Kdev, K, 9 * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(c->P_view, c->P1, 12 * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(c->P_view + 12, c->P2, 12 * sizeof(double), cudaMemcpyHostToDevice);
  }
  if (!Kdev) {
    sfm_set_error("sfm_chain_create: arena too small");
    sfm_chain_destroy(c);
    return SFM_ERR_NOMEM;
  }
  *out = c;
  return SFM_OK;
}

static void sfm_chain_release(sfm_chain* c);

extern "C" void sfm_chain_destroy(sfm_chain* c) {
  if (!c) return;
  if (c->ctx) cudaStreamSynchronize(c->ctx->stream);
  for (sfm_ctx* side : {c->ctxB, c->ctxC})
    if (side) {
      cudaStreamSynchronize(side->stream);
      sfm_ctx_merge_profile(c->ctx, side);
    }
  if (c->ctx && !c->ctx->chain_parked && c->ctxB && c->K_dev) {   // park: the next chain on this context reuses everything
    c->ctx->chain_parked = c;
    return;
  }
  sfm_chain_release(c);
}

void sfm_chain_parked_free(sfm_ctx* ctx) {
  if (ctx->chain_parked) sfm_chain_release(static_cast<sfm_chain*>(ctx->chain_parked));
  ctx->chain_parked = nullptr;
}

static void sfm_chain_release(sfm_chain* c) {
  if (c->ctxB) sfm_ctx_destroy(c->ctxB);
  if (c->ctxC) sfm_ctx_destroy(c->ctxC);
  for (int k = 0; k < 2; ++k) {
    if (c->ev_assoc[k]) cudaEventDestroy(c->ev_assoc[k]);
    if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]);
    if (c->ev_core[k]) cudaEventDestroy(c->ev_core[k]);
    if (c->ev_out[k]) cudaEventDestroy(c->ev_out[k]);
  }
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  if (c->ar.base) cudaFree(c->ar.base);
  if (c->hcnt) cudaFreeHost(c->hcnt);
  if (c->herrs) cudaFreeHost(c->herrs);
  if (c->hrecs) cudaFreeHost(c->hrecs);
  delete c;
}

// Feed the next n_pairs consecutive pairs.  The very first pair bootstraps the model (no view is
// registered by it); every further pair registers one view: out / X_new have one entry per registered
// view of THIS call.  The match arrays of the last pair must stay alive until the next call (they are
// re-triangulated against the next view, sfm.py:348-352).
static int chain_collect_impl(sfm_chain* c, sfm_view_out* out, int32_t* n_registered) {
  sfm_ctx* ctx = c->ctx;
  const int reg = c->pend_reg;
  if (n_registered) *n_registered = 0;
  if (!c->pending) return SFM_OK;
  c->pending = false;
  if (c->pending_parsed) {           // the host-synchronised loop already produced the records
    for (int v = 0; v < reg; ++v) out[v] = c->pend_out[v];
    if (n_registered) *n_registered = reg;
    return SFM_OK;
  }
  SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int v = 0; v < reg; ++v) {
    const ViewRec& r = c->hrecs[v];
    sfm_view_out& o = out[v];
    memcpy(o.Rt, r.Rt, sizeof(o.Rt));
    o.err_pnp = r.err_pnp; o.err_new = r.err_new;
    o.n_new = r.n_new; o.n_pnp = r.n_pnp; o.n_inl = r.n_inl; o.n_match = c->pend_nmatch[v];
    if (r.n_pnp < 6 || !r.ok) {
      // The whole call was queued before any record could be looked at: the views after this one were computed from
      // its (undefined) pose.  out[0 .. v) are valid and reported; the chain cannot be extended any further.
      c->dead = true;
      if (n_registered) *n_registered = v;
      if (r.n_pnp < 6)
        sfm_set_error("registration failed: view %d shares %d points with the model (6 needed); %d view(s) of this call are valid",
                      c->views_done - reg + v + 2, r.n_pnp, v);
      else
        sfm_set_error("registration failed: solvePnPRansac found no consensus (view %d); %d view(s) of this call are valid",
                      c->views_done - reg + v + 2, v);
      return SFM_ERR_INVALID;
    }
  }
  if (n_registered) *n_registered = reg;
  return SFM_OK;
}

static int chain_extend_impl(sfm_chain* c, int n_pairs, const float* const* pts_q, const float* const* pts_t,
                             const int32_t* n_match, float* const* X_new, sfm_view_out* out, int32_t* n_registered,
                             bool defer) {
  SFM_REQUIRE(c && n_pairs >= 0 && (n_pairs == 0 || (pts_q && pts_t && n_match)), "sfm_chain_extend: null argument");
  SFM_REQUIRE(!c->pending, "sfm_chain_extend: the previous asynchronous call has not been collected");
  SFM_REQUIRE(!c->dead, "sfm_chain_extend: a view of an earlier call failed to register; the chain cannot be extended");
  sfm_ctx* ctx = c->ctx;
  const double* K = c->K;
  int reg = 0;
  if (n_registered) *n_registered = 0;
  for (int k = 0; k < n_pairs; ++k)
    SFM_REQUIRE(n_match[k] >= 6 && n_match[k] <= c->nmax, "sfm_chain_extend: pair %d has %d matches (need 6..%d)", k, n_match[k], c->nmax);
  const int n_views = n_pairs - (c->started ? 0 : 1);
  if (defer) { c->pend_out.assign((size_t)(n_views > 0 ? n_views : 1), sfm_view_out()); out = c->pend_out.data(); }
  SFM_REQUIRE(n_views <= 0 || (X_new && out), "sfm_chain_extend: output arrays missing");
  SFM_REQUIRE(n_views < ERR_SLOTS, "sfm_chain_extend: at most %d views per call", ERR_SLOTS - 1);
  if (n_pairs == 0) return SFM_OK;
  SFM_CUDA(cudaMemsetAsync(c->errs, 0, sizeof(double) * 2 * (size_t)(n_views > 0 ? n_views : 1), ctx->stream));
  int k0 = 0;
  if (!c->started) {
    // ---- state when the reference's loop starts (sfm.py:304-339; the second pose is given)
    const int M = n_match[0];
    SFM_TRY(sfm_triangulate(ctx, c->P1, c->P2, pts_q[0], pts_t[0], M, 1, c->X4, 1, 1));
    SFM_TRY(sfm_reproj_error(ctx, c->X4, 2, pts_t[0], 1, M, c->Rt1, K, c->errs + 2 * (ERR_SLOTS - 1), nullptr, c->Xc));
    double rvec[3], tvec[3];
    int32_t ni = 0, ok = 0;
    SFM_TRY(sfm_pnp_ransac(ctx, c->Xc, pts_t[0], M, K, 100, 8.0f, 0.99, rvec, tvec, c->inl, &ni, &ok, nullptr));
    SFM_REQUIRE(ok && ni >= 1, "registration failed: solvePnPRansac found no consensus on the first pair");
    SFM_TRY(sfm_gather_rows(ctx, pts_t[0], 2, c->inl, ni, c->boot_pts1));
    SFM_TRY(sfm_gather_rows(ctx, c->Xc, 3, c->inl, ni, c->boot_p3d));
    c->pts1 = c->boot_pts1; c->points_3d = c->boot_p3d; c->n1 = ni;
    c->prev_q = nullptr; c->prev_t = nullptr; c->prev_n = 0;
    c->started = true;
    k0 = 1;
  }
  const bool sync_free = getenv("SFM_CHAIN_SYNC") == nullptr;
  if (sync_free && n_views > 0) {
    // ---- no host round trip inside the loop: counts, poses and matrices are produced and consumed in HBM
    SFM_REQUIRE(c->views_done + n_views < ERR_SLOTS, "sfm_chain_extend: more than %d views in one chain", ERR_SLOTS);
    SFM_CUDA(cudaMemsetAsync(c->recs, 0, sizeof(ViewRec) * (size_t)n_views, ctx->stream));
    sfm_ctx* cb = c->ctxB;
    sfm_ctx* cc = c->ctxC;
    cb->profiling = cc->profiling = ctx->profiling;        // their kernels show up in the caller's profile
    SFM_CUDA(cudaEventRecord(c->ev_start, ctx->stream));
    SFM_CUDA(cudaStreamWaitEvent(cb->stream, c->ev_start, 0));          // records are zeroed before B / C write into them
    SFM_CUDA(cudaStreamWaitEvent(cc->stream, c->ev_start, 0));
    for (int k = k0; k < n_pairs; ++k, ++reg) {
      const int M = n_match[k];
      const float* q = pts_q[k];
      const float* t = pts_t[k];
      const int g = c->views_done;                         // views g, g+1 are the previous pair; g+2 is registered now
      const int set = g & 1;
      ViewRec* rec = c->recs + reg;
      int32_t* i1 = set ? c->i1b : c->i1;
      int32_t* i2 = set ? c->i2b : c->i2;
      uint8_t* keep = set ? c->keepb : c->keep;
      float* temp1 = set ? c->temp1b : c->temp1;
      float* temp2 = set ? c->temp2b : c->temp2;
      float* com2 = set ? c->com2b : c->com2;
      float* Xc = set ? c->Xcb : c->Xc;
      int32_t* inl = set ? c->inlb : c->inl;
      if (c->prev_q) { c->n1 = c->prev_n; c->pts1 = c->prev_t; }
      const int n1 = c->n1;
      // ---- stream B: data association (sfm.py:356) and the complement — 2-D data only, no pose needed.
      // This parity's buffers were last read by the outputs of view g-2.
      if (c->set_used[set]) SFM_CUDA(cudaStreamWaitEvent(cb->stream, c->ev_out[set], 0));
      SFM_TRY(sfm_common_points(cb, c->pts1, n1, q, M, i1, i2, &rec->n_pnp, keep));
      SFM_TRY(sfm_compact_pairs(cb, q, t, keep, M, temp1, temp2, &rec->n_new));
      SFM_TRY(sfm_gather_rows_dev(cb, t, 2, i2, n1, &rec->n_pnp, com2));
      SFM_TRY(sfm_pnp_subsets_dev(cb, &rec->n_pnp, c->subs[set]));     // the RANSAC index stream needs only the count
      SFM_CUDA(cudaEventRecord(c->ev_assoc[set], cb->stream));
      // ---- main stream: the pose dependency — re-triangulation with the previous pose, then PnP
      SFM_CUDA(cudaStreamWaitEvent(ctx->stream, c->ev_assoc[set], 0));
      if (c->set_used[set]) SFM_CUDA(cudaStreamWaitEvent(ctx->stream, c->ev_out[set], 0));   // Xc / inl of this parity free again
      if (c->prev_q) {
        // re-triangulate the previous pair's matches (sfm.py:348-352) — only the rows data association kept
        // (points_3d is used through points_3d[indx1] alone, sfm.py:362), straight into the PnP's point array
        SFM_TRY(sfm_triangulate_dev(ctx, c->P_view + 12 * (size_t)g, c->prev_q, c->prev_t, n1, &rec->n_pnp, Xc, 2, i1));
      } else {
        SFM_TRY(sfm_gather_rows_dev(ctx, c->points_3d, 3, i1, n1, &rec->n_pnp, Xc));
      }
      SFM_TRY(sfm_pnp_ransac_dev(ctx, Xc, com2, n1, &rec->n_pnp, K, c->K_dev, c->pose6, inl, &rec->n_inl, &rec->ok,
                                 rec->Rt, c->P_view + 12 * (size_t)(g + 2), c->cams + reg, c->subs[set]));          // sfm.py:362
      SFM_CUDA(cudaEventRecord(c->ev_core[set], ctx->stream));
      // ---- stream C: what nothing later depends on — the two reprojection errors and the new points
      SFM_CUDA(cudaStreamWaitEvent(cc->stream, c->ev_core[set], 0));
      SFM_TRY(sfm_gather_rows_dev(cc, Xc, 3, inl, n1, &rec->n_inl, c->X_in));
      SFM_TRY(sfm_gather_rows_dev(cc, com2, 2, inl, n1, &rec->n_inl, c->p_in));
      SFM_TRY(sfm_reproj_error_dev(cc, c->X_in, 0, c->p_in, n1, &rec->n_inl, c->cams + reg, &rec->err_pnp, nullptr));   // sfm.py:368
      SFM_TRY(sfm_triangulate_dev(cc, c->P_view + 12 * (size_t)(g + 1), temp1, temp2, M, &rec->n_new, c->X4, 1));        // sfm.py:371
      SFM_TRY(sfm_reproj_error_dev(cc, c->X4, 2, temp2, M, &rec->n_new, c->cams + reg, &rec->err_new, X_new[reg]));      // sfm.py:372
      SFM_CUDA(cudaEventRecord(c->ev_out[set], cc->stream));
      c->set_used[set] = true;
      c->prev_q = q; c->prev_t = t; c->prev_n = M;
      c->views_done += 1;
    }
    for (int k = 0; k < 2; ++k)
      if (c->set_used[k]) SFM_CUDA(cudaStreamWaitEvent(ctx->stream, c->ev_out[k], 0));   // the records are complete
    SFM_CUDA(cudaMemcpyAsync(c->hrecs, c->recs, sizeof(ViewRec) * (size_t)reg, cudaMemcpyDeviceToHost, ctx->stream));
    c->pending = true; c->pending_parsed = false; c->pend_reg = reg;
    c->pend_nmatch.assign(n_match + k0, n_match + k0 + reg);
    if (defer) return SFM_OK;                              // sfm_chain_collect synchronises and reads the records
    return chain_collect_impl(c, out, n_registered);
  }
  for (int k = k0; k < n_pairs; ++k, ++reg) {
    const int M = n_match[k];
    const float* q = pts_q[k];
    const float* t = pts_t[k];
    if (c->prev_q) {                                     // re-triangulate the previous pair's matches (sfm.py:348-352)
      c->n1 = c->prev_n;
      c->pts1 = c->prev_t;
      SFM_TRY(sfm_triangulate(ctx, c->P1, c->P2, c->prev_q, c->prev_t, c->n1, 1, c->pts3d_a, 2, 1));
      c->points_3d = c->pts3d_a;
    }
    // data association (sfm.py:356) and the complement (new points)
    SFM_TRY(sfm_common_points(ctx, c->pts1, c->n1, q, M, c->i1, c->i2, c->cnt, c->keep));
    SFM_TRY(sfm_compact_pairs(ctx, q, t, c->keep, M, c->temp1, c->temp2, c->cnt + 1));
    SFM_CUDA(cudaMemcpyAsync(c->hcnt, c->cnt, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(cudaStreamSynchronize(ctx->stream));
    const int nc = c->hcnt[0], m = c->hcnt[1];
    SFM_REQUIRE(nc >= 6, "registration failed: view %d shares %d points with the model", c->views_done + 2, nc);
    SFM_TRY(sfm_gather_rows(ctx, c->points_3d, 3, c->i1, nc, c->Xc));
    SFM_TRY(sfm_gather_rows(ctx, t, 2, c->i2, nc, c->com2));
    double rvec[3], tvec[3];
    int32_t ki = 0, ok = 0;
    SFM_TRY(sfm_pnp_ransac(ctx, c->Xc, c->com2, nc, K, 100, 8.0f, 0.99, rvec, tvec, c->inl, &ki, &ok, nullptr));   // sfm.py:362
    SFM_REQUIRE(ok, "registration failed: solvePnPRansac found no consensus (view %d)", c->views_done + 2);
    sfm_view_out& o = out[reg];
    double R[9];
    hm::rodrigues_to_matrix(rvec, R);
    for (int i = 0; i < 3; ++i) { o.Rt[4 * i] = R[3 * i]; o.Rt[4 * i + 1] = R[3 * i + 1]; o.Rt[4 * i + 2] = R[3 * i + 2]; o.Rt[4 * i + 3] = tvec[i]; }
    double Pnew[12];
    matmul_K_Rt(K, o.Rt, Pnew);
    if (ki > 0) {
      SFM_TRY(sfm_gather_rows(ctx, c->Xc, 3, c->inl, ki, c->X_in));
      SFM_TRY(sfm_gather_rows(ctx, c->com2, 2, c->inl, ki, c->p_in));
      SFM_TRY(sfm_reproj_error(ctx, c->X_in, 0, c->p_in, 1, ki, o.Rt, K, c->errs + 2 * reg, nullptr, nullptr));     // sfm.py:368
    }
    if (m > 0) {
      SFM_TRY(sfm_triangulate(ctx, c->P2, Pnew, c->temp1, c->temp2, m, 1, c->X4, 1, 1));                            // sfm.py:371
      SFM_TRY(sfm_reproj_error(ctx, c->X4, 2, c->temp2, 1, m, o.Rt, K, c->errs + 2 * reg + 1, nullptr, X_new[reg])); // sfm.py:372
    }
    o.n_new = m; o.n_pnp = nc; o.n_inl = ki; o.n_match = M;
    o.err_pnp = o.err_new = 0.0;
    memcpy(c->P1, c->P2, sizeof(c->P1));
    memcpy(c->P2, Pnew, sizeof(c->P2));
    c->prev_q = q; c->prev_t = t; c->prev_n = M;
    c->views_done += 1;
  }
  if (reg > 0) {
    SFM_CUDA(cudaMemcpyAsync(c->herrs, c->errs, sizeof(double) * 2 * reg, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int v = 0; v < reg; ++v) { out[v].err_pnp = c->herrs[2 * v]; out[v].err_new = c->herrs[2 * v + 1]; }
  }
  if (defer) { c->pending = true; c->pending_parsed = true; c->pend_reg = reg; return SFM_OK; }
  if (n_registered) *n_registered = reg;
  return SFM_OK;
}

extern "C" int sfm_chain_extend(sfm_chain* c, int n_pairs, const float* const* pts_q, const float* const* pts_t,
                                const int32_t* n_match, float* const* X_new, sfm_view_out* out, int32_t* n_registered) {
  return chain_extend_impl(c, n_pairs, pts_q, pts_t, n_match, X_new, out, n_registered, false);
}

// The same call split in two so that a host can prepare the next batch of pairs (upload, match) while this one
// registers: _async queues every launch of the call and returns; _collect waits for it and fills the records.
// One call may be in flight per chain.
extern "C" int sfm_chain_extend_async(sfm_chain* c, int n_pairs, const float* const* pts_q, const float* const* pts_t,
                                      const int32_t* n_match, float* const* X_new) {
  return chain_extend_impl(c, n_pairs, pts_q, pts_t, n_match, X_new, nullptr, nullptr, true);
}

extern "C" int sfm_chain_collect(sfm_chain* c, sfm_view_out* out, int32_t* n_registered) {
  SFM_REQUIRE(c && out, "sfm_chain_collect: null argument");
  return chain_collect_impl(c, out, n_registered);
}

extern "C" int sfm_chain_run(sfm_ctx* ctx, const double* K, const double* Rt0, const double* Rt1, int n_pairs,
                             const float* const* pts_q, const float* const* pts_t, const int32_t* n_match,
                             float* const* X_new, sfm_view_out* out) {
  SFM_REQUIRE(ctx && K && Rt0 && Rt1 && n_pairs >= 1 && pts_q && pts_t && n_match, "sfm_chain_run: null argument");
  int nmax = 6;
  for (int k = 0; k < n_pairs; ++k) nmax = std::max(nmax, n_match[k]);
  sfm_chain* c = nullptr;
  SFM_TRY(sfm_chain_create(ctx, K, Rt0, Rt1, nmax, &c));
  int s = sfm_chain_extend(c, n_pairs, pts_q, pts_t, n_match, X_new, out, nullptr);
  sfm_chain_destroy(c);
  return s;
}
