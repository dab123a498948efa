// match.cu — hot path 1 API: descriptor residency, kernel selection, K1c finalisation (cross-split
// merge + sqrt + Lowe ratio + count) and survivor gather.
//
// Reference: bf.knnMatch(des0, des1, k=2) sfm.py:259-260 / isfm.py:71 / test.py:42,225,352;
//            ratio loop sfm.py:262-265; coordinate gather sfm.py:267-268.
#include <math.h>

#include "match_common.cuh"

// ------------------------------------------------------------------ K1c finalise
// One thread per query: merge nsplit x 2 candidate keys, emit idx/dist/good.
// dist = float32(sqrt(d2)) (correctly rounded, like OpenCV's std::sqrt on the float32 sum);
// the Lowe test is evaluated as Python does it: float32 distances widened to double,
// d1 < ratio*d2 in double, strict (sfm.py:264).
// STRIDE 3 (tensor-core kernel): the third key of a split names a column that has the runner-up's
// distance IF it is real (match_tc.cu, epilogue comment); it is evaluated exactly (integer-valued
// float32 arithmetic on the resident copies) only when it would enter the top-2.
template <int STRIDE>
__device__ __forceinline__ void finalize_row(int i, const mkey_t* __restrict__ cand, int nq, int nt, int nsplit, double ratio,
                                             const float* __restrict__ qf, const float* __restrict__ tf,
                                             int* __restrict__ idx, float* __restrict__ dist,
                                             unsigned char* __restrict__ good, int* __restrict__ n_good) {
  bool g = false;
  mkey_t k1 = MKEY_INF, k2 = MKEY_INF;
  if (i < nq) {
    const mkey_t* c = cand + (size_t)i * nsplit * STRIDE;
    for (int s = 0; s < nsplit; ++s) {
      key_insert(c[s * STRIDE], k1, k2);
      key_insert(c[s * STRIDE + 1], k1, k2);
    }
  }
  if (STRIDE == 3) {
    // "check this column" keys: evaluated by the whole warp, one 128-dimensional difference at a time
    // (coalesced 512-byte rows, shuffle reduction) — a few per warp, instead of a divergent per-lane loop
    const int lane = threadIdx.x & 31;
    const mkey_t* c = cand + (size_t)(i < nq ? i : 0) * nsplit * STRIDE;
    for (int s = 0; s < nsplit; ++s) {
      const mkey_t k3 = (i < nq) ? c[s * STRIDE + 2] : MKEY_INF;
      const int col = (int)(unsigned int)(k3 & 0xFFFFFFFFull);
      unsigned need = __ballot_sync(0xffffffffu, k3 < k2 && col < nt);
      while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const int r = __shfl_sync(0xffffffffu, i, src), cc = __shfl_sync(0xffffffffu, col, src);
        const float4 x = reinterpret_cast<const float4*>(qf + (size_t)r * 128)[lane];
        const float4 y = reinterpret_cast<const float4*>(tf + (size_t)cc * 128)[lane];
        const float e0 = x.x - y.x, e1 = x.y - y.y, e2 = x.z - y.z, e3 = x.w - y.w;
        float d2 = e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;     // integers < 2^24: exact in any order
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        if (lane == src && __float_as_uint(d2) == (unsigned int)(k3 >> 32)) key_insert(k3, k1, k2);
      }
    }
  }
  if (i < nq) {
    int i1 = (int)(unsigned int)(k1 & 0xFFFFFFFFull), i2 = (int)(unsigned int)(k2 & 0xFFFFFFFFull);
    bool v1 = (k1 != MKEY_INF) && i1 < nt, v2 = (k2 != MKEY_INF) && i2 < nt;
    float d1 = v1 ? __fsqrt_rn(__uint_as_float((unsigned int)(k1 >> 32))) : __int_as_float(0x7f800000);
    float d2 = v2 ? __fsqrt_rn(__uint_as_float((unsigned int)(k2 >> 32))) : __int_as_float(0x7f800000);
    if (idx) { idx[2 * i] = v1 ? i1 : -1; idx[2 * i + 1] = v2 ? i2 : -1; }
    if (dist) { dist[2 * i] = d1; dist[2 * i + 1] = d2; }
    g = v1 && v2 && ((double)d1 < ratio * (double)d2);
    if (good) good[i] = g ? 1 : 0;
  }
  if (n_good) {
    unsigned m = __ballot_sync(0xffffffffu, g);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_good, __popc(m));
  }
}

template <int STRIDE>
__global__ void __launch_bounds__(256) match_finalize_kernel(const mkey_t* __restrict__ cand, int nq,
                                                              int nt, int nsplit, double ratio,
                                                              const float* __restrict__ qf, const float* __restrict__ tf,
                                                              int* __restrict__ idx, float* __restrict__ dist,
                                                              unsigned char* __restrict__ good,
                                                              int* __restrict__ n_good) {
  finalize_row<STRIDE>(blockIdx.x * blockDim.x + threadIdx.x, cand, nq, nt, nsplit, ratio, qf, tf, idx, dist, good, n_good);
}

// Per-pair work record of the batched path (device array): K1c finalisation and survivor gather.
struct PairWork {
  const mkey_t* cand; const float* qf; const float* tf;
  int nq, nt, nsub, pad_;
  int* idx; float* dist; unsigned char* good; int* n_good;
  const float2* kp_q; const float2* kp_t; float2* pts_q; float2* pts_t; int* qidx; int* tidx; int* n_out;
};

__global__ void __launch_bounds__(256) match_finalize_batched_kernel(const PairWork* __restrict__ work, double ratio) {
  const PairWork w = work[blockIdx.y];
  if ((int)(blockIdx.x * blockDim.x) >= w.nq) return;
  finalize_row<3>(blockIdx.x * blockDim.x + threadIdx.x, w.cand, w.nq, w.nt, w.nsub, ratio, w.qf, w.tf, w.idx, w.dist, w.good,
                  w.n_good);
}

int sfm_match_finalize(sfm_ctx* ctx, const mkey_t* cand, int nq, int nt, int nsplit, double ratio,
                       const float* qf, const float* tf, int32_t* idx, float* dist, uint8_t* good, int32_t* n_good) {
  if (n_good) SFM_CUDA(cudaMemsetAsync(n_good, 0, sizeof(int32_t), ctx->stream));
  if (nq == 0) return SFM_OK;
  if (qf && tf)
    SFM_LAUNCH(ctx, SFM_K_MATCH_FINAL, (match_finalize_kernel<3><<<div_up(nq, 256), 256, 0, ctx->stream>>>(
                                           cand, nq, nt, nsplit, ratio, qf, tf, idx, dist, good, n_good)));
  else
    SFM_LAUNCH(ctx, SFM_K_MATCH_FINAL, (match_finalize_kernel<2><<<div_up(nq, 256), 256, 0, ctx->stream>>>(
                                           cand, nq, nt, nsplit, ratio, nullptr, nullptr, idx, dist, good, n_good)));
  return SFM_OK;
}

// ------------------------------------------------------------------ descriptors
// Generic ingest (dim != 128): float32 resident copy + integrality flag.  dim == 128 uses the fused
// K1b kernel in match_tc.cu, which also writes the tensor-core operand image in the same pass.
template <typename T>
__global__ void __launch_bounds__(256) desc_ingest_kernel(const T* __restrict__ src, size_t count,
                                                           float* __restrict__ dst,
                                                           unsigned int* __restrict__ flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (; i < count; i += stride) {
    float v = (float)src[i];
    dst[i] = v;
    bad |= !(v >= 0.f && v <= 255.f && v == rintf(v));
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

static void descbuf_free(DescBuf* b) {
  if (!b) return;
  if (b->f32) cudaFree(b->f32);
  if (b->tiles) cudaFree(b->tiles);
  if (b->sqnorm) cudaFree(b->sqnorm);
  if (b->flag) cudaFree(b->flag);
  if (b->hflag) cudaFreeHost(b->hflag);
  if (b->ready) cudaEventDestroy(b->ready);
  delete b;
}

void sfm_desc_pool_free(sfm_ctx* c) {
  for (void* p : c->desc_pool) descbuf_free((DescBuf*)p);
  c->desc_pool.clear();
}

static int descbuf_acquire(sfm_ctx* ctx, int n, int dim, DescBuf** out) {
  size_t need = ((size_t)(n > 0 ? n : 1) + 255) & ~(size_t)255;
  for (size_t i = 0; i < ctx->desc_pool.size(); ++i) {
    DescBuf* b = (DescBuf*)ctx->desc_pool[i];
    if (b->dim == dim && b->cap_rows >= need && b->cap_rows <= 2 * need) {
      ctx->desc_pool.erase(ctx->desc_pool.begin() + i);
      *out = b;
      return SFM_OK;
    }
  }
  DescBuf* b = new DescBuf();
  b->cap_rows = need;
  b->dim = dim;
  cudaError_t e = cudaMalloc(&b->f32, need * dim * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&b->flag, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMallocHost(&b->hflag, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ready, cudaEventDisableTiming);
  if (e == cudaSuccess && dim == 128) {
    e = cudaMalloc(&b->tiles, (need / 128) * (size_t)40960);
    if (e == cudaSuccess) e = cudaMalloc(&b->sqnorm, need * sizeof(float));
  }
  if (e != cudaSuccess) {
    sfm_set_error("descriptor storage: %s", cudaGetErrorString(e));
    descbuf_free(b);
    return SFM_ERR_NOMEM;
  }
  *out = b;
  return SFM_OK;
}

extern "C" int sfm_desc_create(sfm_ctx* ctx, const void* data, int dtype, int n, int dim, sfm_desc** out) {
  SFM_REQUIRE(ctx && out, "sfm_desc_create: null argument");
  SFM_REQUIRE(dtype == 0 || dtype == 1, "sfm_desc_create: dtype must be 0 (float32) or 1 (uint8); cv2 rejects others too");
  SFM_REQUIRE(n >= 0 && dim > 0, "sfm_desc_create: bad shape (%d,%d)", n, dim);
  SFM_REQUIRE(n == 0 || data, "sfm_desc_create: null data");
  SFM_TRY(sfm_ws_begin(ctx));
  DescBuf* b = nullptr;
  SFM_TRY(descbuf_acquire(ctx, n, dim, &b));
  sfm_desc* d = new sfm_desc();
  d->ctx = ctx; d->n = n; d->dim = dim; d->buf = b;
  d->f32 = b->f32;
  int s = SFM_OK;
  do {
    if (cudaMemsetAsync(b->flag, 0, sizeof(unsigned int), ctx->stream) != cudaSuccess) { s = SFM_ERR_CUDA; break; }
    size_t count = (size_t)n * dim;
    if (count) {
      const void* src = nullptr;
      if (dtype == 0) { const float* p; if ((s = dev_in(ctx, (const float*)data, count, &p))) break; src = p; }
      else { const uint8_t* p; if ((s = dev_in(ctx, (const uint8_t*)data, count, &p))) break; src = p; }
      if (dim == 128) {
        if ((s = sfm_desc_prepare_launch(ctx, d, src, dtype))) break;
      } else {
        int grid = (int)((count + 255) / 256);
        if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
        if (dtype == 0)
          s = [&]() -> int { SFM_LAUNCH(ctx, SFM_K_DESC_PREP, (desc_ingest_kernel<float><<<grid, 256, 0, ctx->stream>>>((const float*)src, count, b->f32, b->flag))); return SFM_OK; }();
        else
          s = [&]() -> int { SFM_LAUNCH(ctx, SFM_K_DESC_PREP, (desc_ingest_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>((const uint8_t*)src, count, b->f32, b->flag))); return SFM_OK; }();
        if (s) break;
      }
    }
    // the flag word travels to pinned memory asynchronously; it is only waited for at first use
    if (cudaMemcpyAsync(b->hflag, b->flag, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaEventRecord(b->ready, ctx->stream) != cudaSuccess) {
      sfm_set_error("sfm_desc_create: %s", cudaGetErrorString(cudaGetLastError()));
      s = SFM_ERR_CUDA;
      break;
    }
    // a host source buffer was staged through the workspace: the caller may reuse it after return
    if (count && !sfm_is_device_ptr(data) && cudaStreamSynchronize(ctx->stream) != cudaSuccess) { s = SFM_ERR_CUDA; break; }
  } while (0);
  if (s != SFM_OK) { sfm_desc_destroy(d); return s; }
  *out = d;
  return SFM_OK;
}

static void desc_apply_flags(sfm_desc* d, unsigned int f) {
  d->exact = (f & 1u) == 0;
  bool tc = d->exact && !(f & 2u) && d->dim == 128 && d->n > 0 && d->buf->tiles;
  d->tiles = tc ? d->buf->tiles : nullptr;
  d->sqnorm = tc ? d->buf->sqnorm : nullptr;
  d->n_tiles = tc ? (d->n + 127) / 128 : 0;
  d->resolved = true;
}

// Many 128-dimensional descriptor sets (DEVICE sources) prepared by ONE K1b launch; the domain flags of the
// whole batch come back in one copy and the sets are returned resolved.
extern "C" int sfm_desc_create_batched(sfm_ctx* ctx, int count, const void* const* data, int dtype, const int32_t* n,
                                       int dim, sfm_desc** out) {
  SFM_REQUIRE(ctx && count >= 0 && (count == 0 || (data && n && out)), "sfm_desc_create_batched: null argument");
  SFM_REQUIRE(dtype == 0 || dtype == 1, "sfm_desc_create_batched: dtype must be 0 (float32) or 1 (uint8)");
  SFM_REQUIRE(dim == 128, "sfm_desc_create_batched: 128-dimensional descriptors only (use sfm_desc_create otherwise)");
  if (count == 0) return SFM_OK;
  SFM_TRY(sfm_ws_begin(ctx));
  std::vector<sfm_desc*> ds((size_t)count, nullptr);
  int s = SFM_OK;
  for (int k = 0; k < count && s == SFM_OK; ++k) {
    if (n[k] <= 0 || !data[k] || !sfm_is_device_ptr(data[k])) {
      sfm_set_error("sfm_desc_create_batched: set %d must be a non-empty device array", k);
      s = SFM_ERR_INVALID;
      break;
    }
    DescBuf* b = nullptr;
    s = descbuf_acquire(ctx, n[k], dim, &b);
    if (s != SFM_OK) break;
    sfm_desc* d = new sfm_desc();
    d->ctx = ctx; d->n = n[k]; d->dim = dim; d->buf = b; d->f32 = b->f32;
    ds[k] = d;
  }
  unsigned int* flags = nullptr;
  unsigned int* hflags = nullptr;
  if (s == SFM_OK) s = ws_alloc_t(ctx, (size_t)count, &flags);
  if (s == SFM_OK) s = hs_alloc_t(ctx, (size_t)count, &hflags);
  if (s == SFM_OK && cudaMemsetAsync(flags, 0, sizeof(unsigned int) * count, ctx->stream) != cudaSuccess) s = SFM_ERR_CUDA;
  if (s == SFM_OK) s = sfm_desc_prepare_launch_batched(ctx, count, ds.data(), data, dtype, flags);
  if (s == SFM_OK && (cudaMemcpyAsync(hflags, flags, sizeof(unsigned int) * count, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                      cudaStreamSynchronize(ctx->stream) != cudaSuccess)) {
    sfm_set_error("sfm_desc_create_batched: %s", cudaGetErrorString(cudaGetLastError()));
    s = SFM_ERR_CUDA;
  }
  if (s != SFM_OK) {
    for (sfm_desc* d : ds) sfm_desc_destroy(d);
    return s;
  }
  for (int k = 0; k < count; ++k) {
    desc_apply_flags(ds[k], hflags[k]);
    out[k] = ds[k];
  }
  return SFM_OK;
}

int sfm_desc_resolve(sfm_desc* d) {
  if (d->resolved) return SFM_OK;
  SFM_CUDA(cudaEventSynchronize(d->buf->ready));
  unsigned int f = *d->buf->hflag;
  d->exact = (f & 1u) == 0;
  bool tc = d->exact && !(f & 2u) && d->dim == 128 && d->n > 0 && d->buf->tiles;
  d->tiles = tc ? d->buf->tiles : nullptr;
  d->sqnorm = tc ? d->buf->sqnorm : nullptr;
  d->n_tiles = tc ? (d->n + 127) / 128 : 0;
  d->resolved = true;
  return SFM_OK;
}

extern "C" void sfm_desc_destroy(sfm_desc* d) {
  if (!d) return;
  // storage goes back to the pool; reuse is stream-ordered on the ctx stream, so no synchronisation
  if (d->buf && d->ctx) d->ctx->desc_pool.push_back(d->buf);
  delete d;
}
extern "C" int sfm_desc_rows(const sfm_desc* d) { return d ? d->n : 0; }
extern "C" int sfm_desc_is_exact(const sfm_desc* d) {
  if (!d) return 0;
  if (sfm_desc_resolve(const_cast<sfm_desc*>(d)) != SFM_OK) return 0;
  return d->exact ? 1 : 0;
}

// ------------------------------------------------------------------ one pair
static int match_pairs_batched(sfm_ctx* ctx, int npairs, const sfm_desc* const* q, const sfm_desc* const* t, double ratio,
                               int32_t* const* idx, float* const* dist, uint8_t* const* good, int32_t* n_good_dev,
                               const float* const* kp_q, const float* const* kp_t, float* const* pts_q, float* const* pts_t,
                               int32_t* const* qidx, int32_t* const* tidx, int32_t* n_out_dev);
static int match_pair(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, double ratio, int mode,
                      int32_t* idx, float* dist, uint8_t* good, int32_t* n_good) {
  SFM_REQUIRE(q->dim == t->dim, "knnMatch: descriptor dims differ (%d vs %d)", q->dim, t->dim);
  SFM_TRY(sfm_desc_resolve(const_cast<sfm_desc*>(q)));
  SFM_TRY(sfm_desc_resolve(const_cast<sfm_desc*>(t)));
  const int nq = q->n, nt = t->n;
  bool tc_ok = q->exact && t->exact && q->tiles && t->tiles;
  SFM_REQUIRE(mode != 2 || tc_ok || nq == 0 || nt == 0,
              "tensor-core matcher requested but descriptors are not integer-valued in [0,255] with dim 128");
  bool use_tc = (mode == 2) || (mode == 0 && tc_ok);
  bool host_out = false;
  DevOut<int32_t> oidx, ong;
  DevOut<float> odist;
  DevOut<uint8_t> ogood;
  SFM_TRY(dev_out(ctx, idx, (size_t)2 * nq, &oidx, &host_out));
  SFM_TRY(dev_out(ctx, dist, (size_t)2 * nq, &odist, &host_out));
  SFM_TRY(dev_out(ctx, good, (size_t)nq, &ogood, &host_out));
  SFM_TRY(dev_out(ctx, n_good, 1, &ong, &host_out));
  if (nq > 0) {
    int nsplit = 1;
    mkey_t* cand = nullptr;
    if (nt == 0) {
      SFM_TRY(ws_alloc_t(ctx, (size_t)nq * 2, &cand));
      SFM_CUDA(cudaMemsetAsync(cand, 0xFF, (size_t)nq * 2 * sizeof(mkey_t), ctx->stream));
    } else if (use_tc) {
      nsplit = sfm_match_tc_splits(ctx, nq, nt);
      SFM_TRY(ws_alloc_t(ctx, (size_t)q->n_tiles * 128 * nsplit * 3, &cand));
      SFM_TRY(sfm_match_tc_launch(ctx, q, t, cand, nsplit));
    } else {
      nsplit = sfm_match_exact_splits(ctx, nq, nt);
      SFM_TRY(ws_alloc_t(ctx, (size_t)nq * nsplit * 2, &cand));
      SFM_TRY(sfm_match_exact_launch(ctx, q->f32, nq, t->f32, nt, q->dim, cand, nsplit));
    }
    const bool tc_cand = nt > 0 && use_tc;
    SFM_TRY(sfm_match_finalize(ctx, cand, nq, nt, nsplit, ratio, tc_cand ? q->f32 : nullptr, tc_cand ? t->f32 : nullptr,
                               oidx.dev, odist.dev, ogood.dev, ong.dev));
  } else if (ong.dev) {
    SFM_CUDA(cudaMemsetAsync(ong.dev, 0, sizeof(int32_t), ctx->stream));
  }
  SFM_TRY(dev_out_finish(ctx, &oidx));
  SFM_TRY(dev_out_finish(ctx, &odist));
  SFM_TRY(dev_out_finish(ctx, &ogood));
  SFM_TRY(dev_out_finish(ctx, &ong));
  if (host_out) SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFM_OK;
}

extern "C" int sfm_desc_match(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, double ratio,
                              int32_t* idx, float* dist, uint8_t* good, int32_t* n_good, int mode) {
  SFM_REQUIRE(ctx && q && t, "sfm_desc_match: null argument");
  SFM_REQUIRE(mode >= 0 && mode <= 2, "sfm_desc_match: mode %d", mode);
  SFM_TRY(sfm_ws_begin(ctx));
  return match_pair(ctx, q, t, ratio, mode, idx, dist, good, n_good);
}

extern "C" int sfm_desc_match_batched(sfm_ctx* ctx, int npairs, const sfm_desc* const* q, const sfm_desc* const* t,
                                      double ratio, int32_t* const* idx, float* const* dist, uint8_t* const* good,
                                      int32_t* n_good) {
  SFM_REQUIRE(ctx && npairs >= 0 && (npairs == 0 || (q && t)), "sfm_desc_match_batched: null argument");
  SFM_TRY(sfm_ws_begin(ctx));
  for (int p = 0; p < npairs; ++p) {
    int32_t* pi = idx ? idx[p] : nullptr;
    float* pd = dist ? dist[p] : nullptr;
    uint8_t* pg = good ? good[p] : nullptr;
    SFM_REQUIRE((!pi || sfm_is_device_ptr(pi)) && (!pd || sfm_is_device_ptr(pd)) && (!pg || sfm_is_device_ptr(pg)),
                "sfm_desc_match_batched: per-pair outputs must be device pointers");
  }
  // ONE K1 launch over the items of all pairs + one K1c grid; outputs are device pointers, so nothing
  // synchronises until the caller does (or until a host n_good is copied back).
  const bool ng_dev = n_good && sfm_is_device_ptr(n_good);
  int32_t* ng_stage = n_good;
  if (n_good && !ng_dev && npairs) SFM_TRY(ws_alloc_t(ctx, (size_t)npairs, &ng_stage));
  SFM_TRY(match_pairs_batched(ctx, npairs, q, t, ratio, idx, dist, good, ng_stage, nullptr, nullptr, nullptr, nullptr, nullptr,
                              nullptr, nullptr));
  if (n_good && !ng_dev && npairs) {
    SFM_CUDA(cudaMemcpyAsync(n_good, ng_stage, sizeof(int32_t) * npairs, cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return SFM_OK;
}

extern "C" int sfm_knn2_l2_ratio(sfm_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim,
                                 double ratio, int32_t* idx, float* dist, uint8_t* good,
                                 int32_t* n_good, int mode) {
  SFM_REQUIRE(ctx, "sfm_knn2_l2_ratio: null ctx");
  SFM_REQUIRE(mode >= 0 && mode <= 2, "sfm_knn2_l2_ratio: mode %d", mode);
  SFM_REQUIRE(nq >= 0 && nt >= 0 && dim > 0, "sfm_knn2_l2_ratio: bad shape");
  sfm_desc *dq = nullptr, *dt = nullptr;
  int s = sfm_desc_create(ctx, q, 0, nq, dim, &dq);
  if (s == SFM_OK) s = sfm_desc_create(ctx, t, 0, nt, dim, &dt);
  if (s == SFM_OK) s = sfm_desc_match(ctx, dq, dt, ratio, idx, dist, good, n_good, mode);
  sfm_desc_destroy(dq);
  sfm_desc_destroy(dt);
  return s;
}

// ------------------------------------------------------------------ survivor gather
// Stable single-CTA compaction (ascending queryIdx, like the Python loop's append order).
__device__ __forceinline__ void gather_cta(const int* __restrict__ idx, const unsigned char* __restrict__ good, int nq,
                                           const float2* __restrict__ kp_q, const float2* __restrict__ kp_t,
                                           float2* __restrict__ pts_q, float2* __restrict__ pts_t,
                                           int* __restrict__ qidx, int* __restrict__ tidx, int* __restrict__ n_out) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int start = 0; start < nq; start += 1024) {
    int i = start + threadIdx.x;
    bool f = (i < nq) && good[i];
    unsigned m = __ballot_sync(0xffffffffu, f);
    int pre = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[w] = __popc(m);
    __syncthreads();
    int off = 0;
    for (int k = 0; k < w; ++k) off += warp_tot[k];
    int base = base_s;
    if (f) {
      int pos = base + off + pre;
      int tj = idx[2 * i];
      if (pts_q) pts_q[pos] = kp_q[i];
      if (pts_t) pts_t[pos] = kp_t[tj];
      if (qidx) qidx[pos] = i;
      if (tidx) tidx[pos] = tj;
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

__global__ void __launch_bounds__(1024) match_gather_kernel(const int* __restrict__ idx,
                                                             const unsigned char* __restrict__ good, int nq,
                                                             const float2* __restrict__ kp_q,
                                                             const float2* __restrict__ kp_t,
                                                             float2* __restrict__ pts_q, float2* __restrict__ pts_t,
                                                             int* __restrict__ qidx, int* __restrict__ tidx,
                                                             int* __restrict__ n_out) {
  gather_cta(idx, good, nq, kp_q, kp_t, pts_q, pts_t, qidx, tidx, n_out);
}

__global__ void __launch_bounds__(1024) match_gather_batched_kernel(const PairWork* __restrict__ work) {
  const PairWork w = work[blockIdx.x];
  gather_cta(w.idx, w.good, w.nq, w.kp_q, w.kp_t, w.pts_q, w.pts_t, w.qidx, w.tidx, w.n_out);
}

extern "C" int sfm_match_gather(sfm_ctx* ctx, const int32_t* idx, const uint8_t* good, int nq,
                                const float* kp_q, const float* kp_t, float* pts_q, float* pts_t,
                                int32_t* qidx_out, int32_t* tidx_out, int32_t* n_out) {
  SFM_REQUIRE(ctx && (nq == 0 || (idx && good)), "sfm_match_gather: null argument");
  SFM_REQUIRE((!pts_q || kp_q) && (!pts_t || kp_t), "sfm_match_gather: keypoints missing");
  SFM_REQUIRE(sfm_is_device_ptr(idx) || nq == 0, "sfm_match_gather: idx/good/kp must be device pointers");
  SFM_TRY(sfm_ws_begin(ctx));
  bool host_out = false;
  DevOut<float> oq, ot;
  DevOut<int32_t> oqi, oti, on;
  SFM_TRY(dev_out(ctx, pts_q, (size_t)2 * nq, &oq, &host_out));
  SFM_TRY(dev_out(ctx, pts_t, (size_t)2 * nq, &ot, &host_out));
  SFM_TRY(dev_out(ctx, qidx_out, (size_t)nq, &oqi, &host_out));
  SFM_TRY(dev_out(ctx, tidx_out, (size_t)nq, &oti, &host_out));
  SFM_TRY(dev_out(ctx, n_out, 1, &on, &host_out));
  SFM_LAUNCH(ctx, SFM_K_GATHER, (match_gather_kernel<<<1, 1024, 0, ctx->stream>>>(
                                    idx, good, nq, (const float2*)kp_q, (const float2*)kp_t, (float2*)oq.dev,
                                    (float2*)ot.dev, oqi.dev, oti.dev, on.dev)));
  SFM_TRY(dev_out_finish(ctx, &oq));
  SFM_TRY(dev_out_finish(ctx, &ot));
  SFM_TRY(dev_out_finish(ctx, &oqi));
  SFM_TRY(dev_out_finish(ctx, &oti));
  SFM_TRY(dev_out_finish(ctx, &on));
  if (host_out) SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFM_OK;
}

// ------------------------------------------------------------------ batched pairs
// Match + ratio test (+ optional survivor gather) for many pairs with THREE launches in total: one K1
// kernel over the items of every pair, one K1c grid, one gather grid (one CTA per pair).  All per-pair
// outputs are device pointers; idx/good may be NULL per pair (workspace is used).  Pairs whose
// descriptors are not tensor-core eligible (or empty) go through the single-pair path.
static int match_pairs_batched(sfm_ctx* ctx, int npairs, const sfm_desc* const* q, const sfm_desc* const* t, double ratio,
                               int32_t* const* idx, float* const* dist, uint8_t* const* good, int32_t* n_good_dev,
                               const float* const* kp_q, const float* const* kp_t, float* const* pts_q, float* const* pts_t,
                               int32_t* const* qidx, int32_t* const* tidx, int32_t* n_out_dev) {
  const bool gather = n_out_dev != nullptr;
  std::vector<int> fast, slow;
  for (int p = 0; p < npairs; ++p) {
    SFM_REQUIRE(q[p] && t[p], "batched match: pair %d has a null descriptor set", p);
    SFM_REQUIRE(q[p]->dim == t[p]->dim, "knnMatch: descriptor dims differ (%d vs %d)", q[p]->dim, t[p]->dim);
    SFM_TRY(sfm_desc_resolve(const_cast<sfm_desc*>(q[p])));
    SFM_TRY(sfm_desc_resolve(const_cast<sfm_desc*>(t[p])));
    const bool tc_ok = q[p]->exact && t[p]->exact && q[p]->tiles && t[p]->tiles && q[p]->n > 0 && t[p]->n > 0;
    (tc_ok ? fast : slow).push_back(p);
  }
  if (n_good_dev) SFM_CUDA(cudaMemsetAsync(n_good_dev, 0, sizeof(int32_t) * npairs, ctx->stream));
  // per-pair idx / good buffers (needed by the gather even if the caller does not want them)
  std::vector<int32_t*> idx_p(npairs);
  std::vector<uint8_t*> good_p(npairs);
  for (int p = 0; p < npairs; ++p) {
    idx_p[p] = idx ? idx[p] : nullptr;
    good_p[p] = good ? good[p] : nullptr;
    const size_t nq = (size_t)q[p]->n;
    if (gather && !idx_p[p]) SFM_TRY(ws_alloc_t(ctx, 2 * nq + 2, &idx_p[p]));
    if (gather && !good_p[p]) SFM_TRY(ws_alloc_t(ctx, nq + 1, &good_p[p]));
  }
  for (int p : slow) {
    int32_t* ng = n_good_dev ? n_good_dev + p : nullptr;
    SFM_TRY(match_pair(ctx, q[p], t[p], ratio, 0, idx_p[p], dist ? dist[p] : nullptr, good_p[p], ng));
  }
  // tables are staged from pageable vectors: cudaMemcpyAsync copies a small pageable source before it
  // returns, so nothing here has to outlive the call (the pinned staging area is recycled per API call)
  std::vector<PairWork> host((size_t)(npairs > 0 ? npairs : 1));
  PairWork* dev = nullptr;
  SFM_TRY(ws_alloc_t(ctx, (size_t)(npairs > 0 ? npairs : 1), &dev));
  int max_nq = 0;
  if (!fast.empty()) {
    std::vector<const sfm_desc*> fq(fast.size()), ft(fast.size());
    std::vector<mkey_t*> cand(fast.size());
    std::vector<int> nsub(fast.size());
    for (size_t k = 0; k < fast.size(); ++k) { fq[k] = q[fast[k]]; ft[k] = t[fast[k]]; }
    SFM_TRY(sfm_match_tc_launch_batched(ctx, (int)fast.size(), fq.data(), ft.data(), cand.data(), nsub.data()));
    for (size_t k = 0; k < fast.size(); ++k) {
      const int p = fast[k];
      PairWork& w = host[k];
      memset(&w, 0, sizeof(w));
      w.cand = cand[k]; w.qf = q[p]->f32; w.tf = t[p]->f32;
      w.nq = q[p]->n; w.nt = t[p]->n; w.nsub = nsub[k];
      w.idx = idx_p[p]; w.dist = dist ? dist[p] : nullptr; w.good = good_p[p];
      w.n_good = n_good_dev ? n_good_dev + p : nullptr;
      if (w.nq > max_nq) max_nq = w.nq;
    }
    SFM_CUDA(cudaMemcpyAsync(dev, host.data(), sizeof(PairWork) * fast.size(), cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid(div_up(max_nq, 256), (unsigned)fast.size());
    SFM_LAUNCH(ctx, SFM_K_MATCH_FINAL, (match_finalize_batched_kernel<<<grid, 256, 0, ctx->stream>>>(dev, ratio)));
  }
  if (gather && npairs > 0) {
    std::vector<PairWork> ghost((size_t)npairs);
    PairWork* gdev = nullptr;
    SFM_TRY(ws_alloc_t(ctx, (size_t)npairs, &gdev));
    for (int p = 0; p < npairs; ++p) {
      PairWork& w = ghost[p];
      memset(&w, 0, sizeof(w));
      w.nq = q[p]->n;
      w.idx = idx_p[p]; w.good = good_p[p];
      w.kp_q = (const float2*)(kp_q ? kp_q[p] : nullptr); w.kp_t = (const float2*)(kp_t ? kp_t[p] : nullptr);
      w.pts_q = (float2*)(pts_q ? pts_q[p] : nullptr); w.pts_t = (float2*)(pts_t ? pts_t[p] : nullptr);
      w.qidx = qidx ? qidx[p] : nullptr; w.tidx = tidx ? tidx[p] : nullptr;
      w.n_out = n_out_dev + p;
      SFM_REQUIRE((!w.pts_q || w.kp_q) && (!w.pts_t || w.kp_t), "batched match: keypoints missing for pair %d", p);
    }
    SFM_CUDA(cudaMemcpyAsync(gdev, ghost.data(), sizeof(PairWork) * npairs, cudaMemcpyHostToDevice, ctx->stream));
    SFM_LAUNCH(ctx, SFM_K_GATHER, (match_gather_batched_kernel<<<npairs, 1024, 0, ctx->stream>>>(gdev)));
  }
  return SFM_OK;
}

extern "C" int sfm_desc_match_gather_batched(sfm_ctx* ctx, int npairs, const sfm_desc* const* q, const sfm_desc* const* t,
                                             double ratio, const float* const* kp_q, const float* const* kp_t,
                                             int32_t* const* idx, uint8_t* const* good, float* const* pts_q,
                                             float* const* pts_t, int32_t* const* qidx, int32_t* const* tidx,
                                             int32_t* n_out) {
  SFM_REQUIRE(ctx && npairs >= 0 && (npairs == 0 || (q && t && n_out)), "sfm_desc_match_gather_batched: null argument");
  SFM_REQUIRE(npairs == 0 || sfm_is_device_ptr(n_out), "sfm_desc_match_gather_batched: n_out must be a device array of npairs");
  SFM_TRY(sfm_ws_begin(ctx));
  return match_pairs_batched(ctx, npairs, q, t, ratio, idx, nullptr, good, nullptr, kp_q, kp_t, pts_q, pts_t, qidx, tidx, n_out);
}
