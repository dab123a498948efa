// match_exact.cu — K1': brute-force 2-NN in fp32 on CUDA cores, accumulating (a-b)^2 directly.
//
// This is the path for descriptors that are NOT integer-valued in [0,255] (so the bf16
// tensor-core path would round them); for SIFT descriptors the engine uses match_tc.cu.
// Replaces cv2.BFMatcher(NORM_L2).knnMatch(k=2), reference sfm.py:259-260 / isfm.py:71.
//
// Tiling: a CTA owns 64 query rows and walks a slice of the train rows 64 at a time; the K
// dimension is staged through shared memory 16 columns at a time (k-major, padded), every
// thread keeps a 4x4 block of partial squared distances in registers and a private running
// top-2 (64-bit keys) for each of its 4 query rows, merged across the 16 threads of a row with
// shuffles at the end.
#include "match_common.cuh"

#define BQ 64
#define BT 64
#define KC 16

__global__ void __launch_bounds__(256) match_exact_kernel(const float* __restrict__ Q, int nq,
                                                           const float* __restrict__ T, int nt,
                                                           int dim, mkey_t* __restrict__ cand,
                                                           int nsplit, int rows_per_split) {
  __shared__ float sq[KC][BQ + 4];
  __shared__ float st[KC][BT + 4];
  const int tx = threadIdx.x & 15;   // train micro-column
  const int ty = threadIdx.x >> 4;   // query micro-row
  const int q0 = blockIdx.x * BQ;
  const int split = blockIdx.y;
  const int t_begin = split * rows_per_split;
  const int t_end = min(nt, t_begin + rows_per_split);

  mkey_t k1[4], k2[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) { k1[a] = MKEY_INF; k2[a] = MKEY_INF; }

  // loader mapping: 256 threads load a 64 x 16 panel: row = tid/4, 4 floats at column (tid%4)*4
  const int lrow = threadIdx.x >> 2;
  const int lcol = (threadIdx.x & 3) * 4;

  for (int t0 = t_begin; t0 < t_end; t0 += BT) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    for (int k0 = 0; k0 < dim; k0 += KC) {
      float4 vq = make_float4(0.f, 0.f, 0.f, 0.f), vt = vq;
      int qr = q0 + lrow, tr = t0 + lrow;
      if (qr < nq && k0 + lcol < dim) vq = __ldg(reinterpret_cast<const float4*>(Q + (size_t)qr * dim + k0 + lcol));
      if (tr < t_end && k0 + lcol < dim) vt = __ldg(reinterpret_cast<const float4*>(T + (size_t)tr * dim + k0 + lcol));
      __syncthreads();
      sq[lcol + 0][lrow] = vq.x; sq[lcol + 1][lrow] = vq.y; sq[lcol + 2][lrow] = vq.z; sq[lcol + 3][lrow] = vq.w;
      st[lcol + 0][lrow] = vt.x; st[lcol + 1][lrow] = vt.y; st[lcol + 2][lrow] = vt.z; st[lcol + 3][lrow] = vt.w;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        float a4[4], b4[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) a4[a] = sq[k][ty * 4 + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) b4[b] = st[k][tx + 16 * b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            float d = a4[a] - b4[b];
            acc[a][b] = fmaf(d, d, acc[a][b]);
          }
      }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int tj = t0 + tx + 16 * b;
      if (tj < t_end) {
#pragma unroll
        for (int a = 0; a < 4; ++a) key_insert(make_key(acc[a][b], tj), k1[a], k2[a]);
      }
    }
  }
  // merge the 16 threads (consecutive lanes) that share each query row
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      mkey_t o1 = __shfl_xor_sync(0xffffffffu, k1[a], o);
      mkey_t o2 = __shfl_xor_sync(0xffffffffu, k2[a], o);
      key_insert(o1, k1[a], k2[a]);
      key_insert(o2, k1[a], k2[a]);
    }
    int qi = q0 + ty * 4 + a;
    if (tx == 0 && qi < nq) {
      cand[((size_t)qi * nsplit + split) * 2 + 0] = k1[a];
      cand[((size_t)qi * nsplit + split) * 2 + 1] = k2[a];
    }
  }
}

int sfm_match_exact_splits(sfm_ctx* ctx, int nq, int nt) {
  int qtiles = div_up(nq, BQ);
  int want = div_up(2 * ctx->sm_count * 2, qtiles > 0 ? qtiles : 1);   // ~4 CTAs per SM in flight
  int max_splits = div_up(nt, BT);
  int s = want < 1 ? 1 : want;
  if (s > max_splits) s = max_splits;
  if (s > 64) s = 64;
  return s < 1 ? 1 : s;
}

int sfm_match_exact_launch(sfm_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim,
                           mkey_t* cand, int nsplit) {
  SFM_REQUIRE(dim % 4 == 0, "fp32 matcher needs dim %% 4 == 0 (got %d)", dim);
  int rows_per_split = div_up(div_up(nt, nsplit), BT) * BT;
  dim3 grid(div_up(nq, BQ), nsplit);
  SFM_LAUNCH(ctx, SFM_K_MATCH_EXACT, (match_exact_kernel<<<grid, 256, 0, ctx->stream>>>(
                                         q, nq, t, nt, dim, cand, nsplit, rows_per_split)));
  return SFM_OK;
}
