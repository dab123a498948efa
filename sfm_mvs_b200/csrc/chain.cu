// chain.cu — device-side glue of the per-view registration loop (sfm.py:341-409): the fancy
// indexing the reference does in NumPy between its cv2 calls, kept on the GPU so that matched
// keypoints, 3-D points and masks never travel to the host.
//   pts2[indx2], points_3d[indx1]                     sfm.py:358-362   -> sfm_gather_rows
//   temp_array1/2 = pts2[~mask], pts3[~mask]          sfm.py:229-237   -> sfm_compact_pairs
#include "common.cuh"

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, int width,
                                                          const int* __restrict__ idx, int n, float* __restrict__ dst) {
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
