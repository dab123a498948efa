/* sfm_b200.h — C ABI of the B200-native incremental-SfM geometry engine (libsfm_b200.so).
 *
 * This is the drop-in boundary for the three data-parallel hot paths of
 * FlagArihant2000/sfm-mvs.  The reference has no FFI / plugin layer of its own: its
 * boundary is the OpenCV Python API as called from sfm.py / isfm.py / test.py.  Every
 * entry point below therefore cites the reference call site(s) it replaces; the Python
 * host (package sfm_mvs_b200, ctypes) re-creates the cv2 signatures on top of it, and
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no torch / numpy / OpenCV types.
 *   - every function returns 0 (SFM_OK) or a negative sfm_status; sfm_last_error()
 *     returns a thread-local human-readable message for the last failure.
 *   - data pointers may be HOST or DEVICE pointers, detected per pointer
 *     (cudaPointerGetAttributes).  Host buffers are staged through the context's
 *     workspace and the call returns only after outputs have landed (synchronous, like
 *     the cv2 call it replaces).  When *every* data pointer of a call is a device pointer
 *     the call only enqueues work on the context's stream (asynchronous).
 *   - small parameter blocks (projection matrices, K, poses) are always HOST pointers.
 *   - one sfm_ctx per device and per caller thread; a ctx is not thread-safe.
 *   - there is no CPU fallback anywhere: without a CUDA device sfm_ctx_create fails.
 */
#ifndef SFM_B200_H
#define SFM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFM_B200_VERSION 100

typedef enum sfm_status {
  SFM_OK = 0,
  SFM_ERR_INVALID = -1,      /* bad argument (shape, dtype, null pointer): cv2 raises cv2.error here */
  SFM_ERR_CUDA = -2,         /* a CUDA runtime call failed; message carries cudaGetErrorString */
  SFM_ERR_NOMEM = -3,
  SFM_ERR_UNSUPPORTED = -4,
  SFM_ERR_NCCL = -5
} sfm_status;

typedef struct sfm_ctx sfm_ctx;
typedef struct sfm_desc sfm_desc;   /* a view's descriptors, prepared and resident in HBM */
typedef struct sfm_ba sfm_ba;       /* a bundle-adjustment problem resident in HBM */

/* ------------------------------------------------------------------ context / diagnostics */
int sfm_version(void);
const char* sfm_last_error(void);
/* `stream` is a cudaStream_t to borrow (e.g. torch's current stream) or NULL to own one. */
int sfm_ctx_create(int device, void* stream, sfm_ctx** out);
void sfm_ctx_destroy(sfm_ctx* ctx);
int sfm_ctx_sync(sfm_ctx* ctx);
void* sfm_ctx_stream(sfm_ctx* ctx);
/* The stream outlives the context from now on (sfm_ctx_destroy synchronises but does not destroy it).  For hosts
 * that hand the stream to another runtime whose objects may be released later: torch's pinned-host allocator
 * records an event on every stream a buffer was used on when that buffer is freed, and a destroyed stream there
 * aborts the process. */
int sfm_ctx_detach_stream(sfm_ctx* ctx);
int sfm_ctx_sm_count(sfm_ctx* ctx);
/* n host -> device copies queued on `stream` (NULL: the context's stream) in one call — the upload of a chunk of views
 * (the reference holds keypoints and descriptors of every image as separate arrays, sfm.py:246-252).  Pinned sources
 * copy asynchronously.  dst[i] device, src[i] host, bytes[i] >= 0. */
int sfm_upload_batch(sfm_ctx* ctx, void* stream, int n, void* const* dst, const void* const* src, const int64_t* bytes);

/* Kernel identifiers for the profiling interface. */
enum {
  SFM_K_DESC_PREP = 0,    /* K1b f32/u8 -> bf16 UMMA tiles + exact |d|^2 augmentation */
  SFM_K_MATCH_TC = 1,     /* K1  tcgen05 distance GEMM + fused top-2 epilogue */
  SFM_K_MATCH_EXACT = 2,  /* K1' exact fp32 CUDA-core 2-NN (non-integer descriptors) */
  SFM_K_MATCH_FINAL = 3,  /* K1c cross-split merge + sqrt + Lowe ratio + count */
  SFM_K_GATHER = 4,       /* stable compaction of ratio-test survivors + keypoint gather */
  SFM_K_TRIANGULATE = 5,  /* K2 */
  SFM_K_REPROJ = 6,       /* K3 */
  SFM_K_PNP_SCORE = 7,    /* K4 */
  SFM_K_PNP_REFINE = 8,   /* LM normal equations for SOLVEPNP_ITERATIVE refinement */
  SFM_K_ASSOC = 9,        /* common_points */
  SFM_K_BA_EVAL = 10,     /* K5 residual + Jacobian blocks */
  SFM_K_BA_SCHUR = 11,    /* K6 fused J^T J / Schur accumulation */
  SFM_K_BA_UPDATE = 12,   /* back-substitution + parameter update + cost */
  SFM_K_BA_SOLVE = 13,    /* reduced camera system Cholesky */
  SFM_K_MISC = 14,
  SFM_K_PNP_EPNP = 15,    /* batched 5-point EPnP minimal solver */
  SFM_K_ESSENTIAL = 16,   /* five-point essential-matrix hypotheses + Sampson scoring */
  SFM_K_COUNT = 17
};
const char* sfm_kernel_name(int kernel_id);
/* When on, every kernel launch is bracketed by CUDA events on the ctx stream. */
int sfm_ctx_set_profiling(sfm_ctx* ctx, int on);
int sfm_ctx_reset_profile(sfm_ctx* ctx);
/* Synchronises, then returns accumulated device time and launch count of one kernel id. */
int sfm_ctx_get_profile(sfm_ctx* ctx, int kernel_id, double* ms_total, int64_t* launches);
/* Number of kernels this library launched on this ctx since creation (all ids). */
int64_t sfm_ctx_launch_count(sfm_ctx* ctx);

/* ------------------------------------------------------------------ hot path 1: matching
 * Replaces cv2.BFMatcher().knnMatch(des0, des1, k=2)            sfm.py:259-260, isfm.py:71,
 *                                                               test.py:42,225,352
 * and the Lowe ratio loop `m.distance < 0.70*n.distance`        sfm.py:262-265, isfm.py:73-76,
 *                                                               test.py:226-229,353-356.
 * q (nq,dim) and t (nt,dim) row-major float32.  idx (nq,2) int32, dist (nq,2) float32,
 * ascending distance, ties -> lower train index; dist = float32(sqrt(sum (a-b)^2)).
 * good[i] = (double)dist[i,0] < ratio*(double)dist[i,1].  nt==1 -> idx[:,1]=-1, dist[:,1]=+inf.
 * idx/dist/good/n_good may each be NULL.  mode: 0 auto (tensor cores when every value is an
 * integer in [0,255], as SIFT's are — exact; otherwise the fp32 kernel), 1 force fp32 kernel,
 * 2 force tensor-core kernel (SFM_ERR_INVALID if the descriptors are not bf16-exact). */
int sfm_knn2_l2_ratio(sfm_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim,
                      double ratio, int32_t* idx, float* dist, uint8_t* good, int32_t* n_good,
                      int mode);

/* Diagnostic (used by the tests only): raw tensor-core accumulators -(2^22 + d^2/2) of every
 * (query row, train column), float32 [n_qtiles*128][n_ttiles*128], to a host buffer. */
int sfm_debug_match_tc_dump(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, float* dump_host,
                            int64_t capacity);
/* Diagnostic (tools/match_timeline.py): clock64 stamps of the jobs of CTA pair 0 of one K1 launch,
 * 8 per job (see match_tc.cu); info = {qt, nsplit, n_items, n_pairs, jobs of pair 0}. */
int sfm_debug_match_tc_timeline(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, long long* stamps_host,
                                int max_jobs, int* info);

/* Resident form: prepare a view's descriptors once (K1b), match many times.
 * dtype 0 = float32, 1 = uint8. */
int sfm_desc_create(sfm_ctx* ctx, const void* data, int dtype, int n, int dim, sfm_desc** out);
/* Many 128-D sets in ONE K1b launch: data[k] are DEVICE arrays of n[k] rows; out[k] come back resolved
 * (one synchronisation for the whole batch). */
int sfm_desc_create_batched(sfm_ctx* ctx, int count, const void* const* data, int dtype, const int32_t* n,
                            int dim, sfm_desc** out);
void sfm_desc_destroy(sfm_desc* d);
int sfm_desc_rows(const sfm_desc* d);
int sfm_desc_is_exact(const sfm_desc* d);   /* 1 if integer-valued in [0,255] (tensor path ok) */
int sfm_desc_match(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, double ratio,
                   int32_t* idx, float* dist, uint8_t* good, int32_t* n_good, int mode);
/* Grouped launch over many pairs (isfm.py:68-87 all-pairs loop; sharded over GPUs by the host):
 * ONE persistent kernel walks the tiles of every pair.  Output pointer arrays are host arrays of
 * per-pair DEVICE pointers (any may be NULL); n_good is a device or host array of npairs. */
int sfm_desc_match_batched(sfm_ctx* ctx, int npairs, const sfm_desc* const* q,
                           const sfm_desc* const* t, double ratio, int32_t* const* idx,
                           float* const* dist, uint8_t* const* good, int32_t* n_good);

/* The same plus the survivor gather of sfm.py:267-268 for every pair, three launches in total (K1 over
 * the items of all pairs, K1c, gather).  All arrays are host arrays of per-pair DEVICE pointers;
 * idx / good / qidx / tidx arrays (or single entries) may be NULL; n_out is a DEVICE array of npairs. */
int sfm_desc_match_gather_batched(sfm_ctx* ctx, int npairs, const sfm_desc* const* q,
                                  const sfm_desc* const* t, double ratio, const float* const* kp_q,
                                  const float* const* kp_t, int32_t* const* idx, uint8_t* const* good,
                                  float* const* pts_q, float* const* pts_t, int32_t* const* qidx,
                                  int32_t* const* tidx, int32_t* n_out);

/* sfm.py:267-268 — pts0 = kp0[m.queryIdx], pts1 = kp1[m.trainIdx] for the survivors, ascending
 * queryIdx.  kp_q (nq,2), kp_t (nt,2) float32; outputs (n_good,2) float32, capacity nq rows. */
int sfm_match_gather(sfm_ctx* ctx, const int32_t* idx, const uint8_t* good, int nq,
                     const float* kp_q, const float* kp_t, float* pts_q, float* pts_t,
                     int32_t* qidx_out, int32_t* tidx_out, int32_t* n_out);

/* ------------------------------------------------------------------ hot path 2: triangulation
 * Replaces cv2.triangulatePoints(P1, P2, pts1, pts2)            sfm.py:53, test.py:310,367
 * and `cloud / cloud[3]`                                        sfm.py:54.
 * P1,P2: 12 doubles row-major (host).  pts_layout 0: (2,N) as cv2 takes them, 1: (N,2).
 * out_layout 0: (4,N) like cv2, 1: (N,4), 2: (N,3) (implies normalize_w).  float32 I/O, float64
 * one-sided Jacobi SVD per point, same rotation schedule as OpenCV's. */
int sfm_triangulate(sfm_ctx* ctx, const double* P1, const double* P2, const float* x1,
                    const float* x2, int n, int pts_layout, float* X, int out_layout,
                    int normalize_w);

/* Replaces ReprojectionError's Rodrigues -> projectPoints -> cv2.norm(.., NORM_L2)/N
 *                                                               sfm.py:79-100, ba.pyc L44-63.
 * x_layout 0: (N,3); 1: (4,N) homogeneous; 2: (N,4) homogeneous.  px_layout 0: (2,N), 1: (N,2).
 * Rt 12 doubles, K 9 doubles (host).  err = sqrt(sum |proj-px|^2)/N (host double, synchronises;
 * may be NULL, or a device pointer for asynchronous use).  proj (N,2) f32 and X3 (N,3) f32 (the
 * de-homogenised points, cv2.convertPointsFromHomogeneous) are optional outputs. */
int sfm_reproj_error(sfm_ctx* ctx, const float* X, int x_layout, const float* px, int px_layout,
                     int n, const double* Rt, const double* K, double* err, float* proj,
                     float* X3);

/* Replaces common_points(pts1, pts2, pts3)                      sfm.py:215-239
 * with its exact semantics: row j of pts2 matches pts1[i] when x OR y is equal as float32;
 * the first such j wins.  idx1/idx2 (capacity n1) ascending in i; keep2[j]=1 for rows of pts2
 * never chosen (the reference's temp_array1/2 are pts2[keep2], pts3[keep2]). */
int sfm_common_points(sfm_ctx* ctx, const float* pts1, int n1, const float* pts2, int n2,
                      int32_t* idx1, int32_t* idx2, int32_t* n_common, uint8_t* keep2);

/* Device-side fancy indexing of the per-view loop (sfm.py:341-409), so that matched keypoints and
 * 3-D points stay in HBM between the calls above.  Device pointers only.
 *   dst[i,:] = src[idx[i],:]            (pts2[indx2], points_3d[indx1]           sfm.py:358-362)
 *   a_out, b_out = a[keep], b[keep]     (temp_array1/2 = pts2[~mask], pts3[~mask] sfm.py:229-237), stable;
 *   n_out may be a host pointer (synchronises) or a device pointer. */
int sfm_gather_rows(sfm_ctx* ctx, const float* src, int width, const int32_t* idx, int n, float* dst);
int sfm_compact_pairs(sfm_ctx* ctx, const float* a, const float* b, const uint8_t* keep, int n,
                      float* a_out, float* b_out, int32_t* n_out);

/* The whole per-view loop of sfm.py:341-409 (imread / SIFT / GUI removed) in one call: for the
 * matches of consecutive pairs k = (view k, view k+1) — device arrays pts_q[k], pts_t[k] of
 * n_match[k] rows, ascending queryIdx — bootstrap on pair 0 with the two given poses
 * (sfm.py:304-339), then register views 2 .. n_pairs: re-triangulate, common_points, PnP-RANSAC,
 * reprojection error, triangulate the new points.  X_new[v] (device, capacity n_match[v+1] x 3)
 * receives view v+2's new points; out[v] (host) its pose, errors and counts. */
typedef struct sfm_view_out {
  double Rt[12];          /* [R|t] of the registered view, row-major 3x4 */
  double err_pnp;         /* ReprojectionError on the PnP inliers     sfm.py:368 */
  double err_new;         /* ReprojectionError on the new points      sfm.py:372 */
  int32_t n_new, n_pnp, n_inl, n_match;
} sfm_view_out;
int sfm_chain_run(sfm_ctx* ctx, const double* K, const double* Rt0, const double* Rt1, int n_pairs,
                  const float* const* pts_q, const float* const* pts_t, const int32_t* n_match,
                  float* const* X_new, sfm_view_out* out);
/* The same loop fed incrementally (pairs arrive while later views are still being uploaded and
 * matched): create with the two bootstrap poses and an upper bound on matches per pair, then extend
 * with consecutive pairs.  The first pair ever fed bootstraps the model; every other pair registers
 * one view (X_new / out: one entry per view registered by that call, *n_registered says how many).
 * The match arrays of the last pair fed must stay valid until the next call. */
typedef struct sfm_chain sfm_chain;
int sfm_chain_create(sfm_ctx* ctx, const double* K, const double* Rt0, const double* Rt1, int max_matches,
                     sfm_chain** out);
void sfm_chain_destroy(sfm_chain* chain);
int sfm_chain_extend(sfm_chain* chain, int n_pairs, const float* const* pts_q, const float* const* pts_t,
                     const int32_t* n_match, float* const* X_new, sfm_view_out* out, int32_t* n_registered);
/* sfm_chain_extend split in two, so the host can upload and match the next batch of pairs while this one
 * registers: _async queues the whole call and returns, _collect waits for it and fills out[] (sized like
 * sfm_chain_extend's).  One call in flight per chain; the bootstrap pair of a fresh chain still synchronises. */
int sfm_chain_extend_async(sfm_chain* chain, int n_pairs, const float* const* pts_q, const float* const* pts_t,
                           const int32_t* n_match, float* const* X_new);
int sfm_chain_collect(sfm_chain* chain, sfm_view_out* out, int32_t* n_registered);

/* cv2.recoverPose(E, pts0, pts1, K)                      sfm.py:311, isfm.py:83, test.py:250:
 * the four (R, t) decompositions of E, every correspondence triangulated against each (float64 DLT as K2) and
 * kept when it lies in front of both cameras closer than `dist` (cv2's default 50); returns the decomposition
 * with the most points, OpenCV's tie order.  pts: (n,2) float32 (dtype 0) or float64 (dtype 2), host or device.
 * mask_in (n, nullable) restricts the count like cv2's in/out mask; mask_out (n) holds 255 / 0. */
int sfm_recover_pose(sfm_ctx* ctx, const double* E, const void* pts1, const void* pts2, int dtype, int n,
                     const double* K, double dist, const uint8_t* mask_in, double* R, double* t,
                     uint8_t* mask_out, int32_t* n_good);

/* cv2.findEssentialMat(pts0, pts1, K, method=RANSAC, prob, threshold, maxIters)
 *                                                          sfm.py:307, isfm.py:80, test.py:247.
 * OpenCV's loop restated: pixels normalised with K in float64, threshold / mean focal length, RNG(2^64-1)
 * five-index subsets, Nister's five-point solver (<= 10 models per subset), Sampson error in float64 stored as
 * float32 (inlier iff err <= (float)thr^2), accept when the count beats max(best, 4), RANSACUpdateNumIters.
 * All maxIters subsets are solved and scored in parallel; one device thread replays the accept/stop recursion.
 * pts: (n,2) float32 (dtype 0) or float64 (dtype 2), host or device.  E: 90 doubles (row-major 3x3 models);
 * mask (n, nullable): 1 / 0 like cv2.  info[6] = {models written to E (0: failure / n < 5, 1: RANSAC winner,
 * k <= 10: n == 5, every model of the single sample as cv2 stacks them), inlier count, iterations run,
 * winning iteration, winning model within it, models scored}. */
int sfm_find_essential_mat(sfm_ctx* ctx, const void* pts1, const void* pts2, int dtype, int n, const double* K,
                           double prob, double threshold, int max_iters, double* E, uint8_t* mask, int32_t* info);
/* EXPERIMENTAL (not yet run on a GPU; unused by bench.py and the default tests): sfm_find_essential_mat for many
 * pairs in a few launches — isfm.py:68-87 calls it once per pair, and one pair's 1000 solver threads leave the GPU
 * almost empty.  pts1/pts2[k]: (n[k],2) float32 DEVICE arrays; masks[k]: device or host, nullable; E: npairs x 9;
 * info: npairs x 6 as above.  Pairs with n < 6 are not attempted (info[6k] = 0): use the single-pair call. */
int sfm_find_essential_mat_batched(sfm_ctx* ctx, int npairs, const float* const* pts1, const float* const* pts2,
                                   const int32_t* n, const double* K, double prob, double threshold, int max_iters,
                                   double* E, uint8_t* const* masks, int32_t* info);
/* Host utility (no GPU needed): Nister's five-point solver on ONE minimal sample — the code the hypothesis kernel
 * runs, compiled for the host.  q1, q2: 5 x 2 normalised coordinates; E: 90 doubles (<= 10 row-major 3x3 models,
 * unit Frobenius norm, ascending E00^2); n_models: how many. */
int sfm_five_point(const double* q1, const double* q2, double* E, int32_t* n_models);

/* ------------------------------------------------------------------ hot path 3a: PnP-RANSAC
 * Replaces cv2.solvePnPRansac(X, p, K, d, ...) with OpenCV's defaults, which is what the
 * reference gets                                                sfm.py:67, test.py:319.
 * Hypothesis scoring (PnPRansacCallback::computeError): Rt (H,12) row-major [R|t] doubles (host),
 * X (N,3), px (N,2) float32.  counts[h] = #{ i : (px-proj)^2 summed in float32 <= thr*thr };
 * masks (H,N) uint8 optional. */
int sfm_pnp_score(sfm_ctx* ctx, const float* X, const float* px, int n, const double* K,
                  const double* Rt, int H, float thr, int32_t* counts, uint8_t* masks);

typedef struct sfm_pnp_info {
  int32_t iters_run;      /* RANSAC iterations cv2 would have executed (adaptive stop) */
  int32_t best_iter;      /* index of the winning hypothesis */
  int32_t hyp_solved;     /* minimal (EPnP) problems actually solved */
  int32_t refine_iters;   /* LM iterations of the final refinement */
  double  rvec_ransac[3]; /* winning hypothesis before refinement */
  double  tvec_ransac[3];
} sfm_pnp_info;

/* Full replacement of the call: RNG(2^64-1) subset stream, 5-point EPnP minimal solver, batched
 * scoring on the GPU, replay of the accept / RANSACUpdateNumIters recursion, LM refinement on the
 * inliers (GPU normal equations).  inliers: capacity n, ascending.  *ok = 0 mirrors cv2 returning
 * (False, .., None).  X, px may be device pointers; rvec/tvec/inliers/n_inliers/ok are host. */
int sfm_pnp_ransac(sfm_ctx* ctx, const float* X, const float* px, int n, const double* K,
                   int max_iters, float thr, double confidence, double* rvec, double* tvec,
                   int32_t* inliers, int32_t* n_inliers, int32_t* ok, sfm_pnp_info* info);

/* Same call with the minimal solutions supplied by the caller: hyp_rt6 (max_iters,6) rvec|tvec per
 * RANSAC iteration (the poses a 5-point solver returned for the subsets of sfm_ransac_subsets),
 * hyp_valid (max_iters) or NULL.  Scoring, the stopping rule, the inlier list and the refinement
 * are identical to sfm_pnp_ransac.  The parity tests use it to separate the two halves of the call: OpenCV's own
 * EPnP output fed through the engine's scoring / replay / refinement. */
int sfm_pnp_ransac_hyp(sfm_ctx* ctx, const float* X, const float* px, int n, const double* K,
                       const double* hyp_rt6, const uint8_t* hyp_valid, int max_iters, float thr,
                       double confidence, double* rvec, double* tvec, int32_t* inliers,
                       int32_t* n_inliers, int32_t* ok, sfm_pnp_info* info);

/* Host utilities used by the wrappers (cv2.Rodrigues / cv2.solvePnP(flags=EPNP) on <=  a few
 * points are host-side parameter marshalling, not data-parallel work). */
int sfm_rodrigues_to_matrix(const double* rvec, double* R9);
int sfm_rodrigues_to_vector(const double* R9, double* rvec);
int sfm_epnp(const float* X, const float* px, int n, const double* K, double* R9, double* t3);
int sfm_ransac_subsets(int n, int iters, int32_t* out /* iters x 5 */);

/* The minimal solver of sfm_pnp_ransac run ON THE DEVICE for explicit subsets: cv2.solvePnP(X[s], px[s], K, 0,
 * flags=SOLVEPNP_EPNP) for each row s of subsets (H,5) (host, H <= 100) -> R9t3 (H,12) host: R row-major | t, the
 * solver's raw output before cv2.Rodrigues.  OpenCV's arithmetic operation for operation (csrc/pnp_epnp.cu):
 * bit-identical to cv2 (tests/test_gpu_pnp.py).  X (n,3), px (n,2) float32, host or device. */
int sfm_epnp_batch(sfm_ctx* ctx, const float* X, const float* px, int n, const double* K, const int32_t* subsets,
                   int H, double* R9t3);

/* ------------------------------------------------------------------ hot path 3b: bundle adjustment
 * Formulation: camera = (rvec 3, tvec 3), shared pinhole K, point = 3, residual =
 * cv2.projectPoints(X, rvec, tvec, K, None) - obs, the projection every reference variant calls
 * (sfm.py:121, test.py:101, ba.pyc L24/L51); block structure as notebook cell 6.
 * Observations must be sorted point-major (pt_idx non-decreasing): a point's observations are
 * contiguous, which is the unit the engine shards across GPUs. */
int sfm_ba_create(sfm_ctx* ctx, int n_cam, int n_pt, int n_obs, const int32_t* cam_idx,
                  const int32_t* pt_idx, const float* obs, const double* K, sfm_ba** out);
void sfm_ba_destroy(sfm_ba* ba);
/* Multi-GPU shards: the whole problem's point / observation counts (the reference's residual modes
 * divide by the global N). */
int sfm_ba_set_totals(sfm_ba* ba, int64_t n_pt_total, int64_t n_obs_total);
int sfm_ba_set_params(sfm_ba* ba, const double* cams /*n_cam x 6*/, const double* pts /*n_pt x 3*/);
int sfm_ba_get_params(sfm_ba* ba, double* cams, double* pts);

/* K5 — materialised residuals and Jacobian blocks (the HBM-roofline kernel):
 * r (n_obs,2) f32, Jc (n_obs,2,6) f32, Jp (n_obs,2,3) f32 (any may be NULL; host or device),
 * cost = 0.5*sum r^2 in float64.  mode 0: r = proj - obs.  mode 1: r = (obs-proj)^2/N, the
 * residual of OptimReprojectionError (sfm.py:124-130).  mode 2: one value per observation,
 * sqrt(dx^2+dy^2)/n_obs, the residual of test.py:108-112 (r is then (n_obs,) and Jc/Jp NULL). */
int sfm_ba_eval(sfm_ba* ba, int mode, float* r, float* Jc, float* Jp, double* cost);

/* The reference's own single-camera formulation (sfm.py:104-157): x = [Rt 12 | K 9 | observed pixels (2,N) | points
 * (N,3)], residual OptimReprojectionError(x) = ((p - proj)^2).ravel() / N in float64.  f0 (2N) receives the residual
 * at x; J ((2N) x n_params, row-major; NULL to skip) the 2-point forward-difference Jacobian with
 * scipy.optimize.least_squares' own steps — all 22 + 5N residual vectors in one launch.  cv2_compat.BundleAdjustment
 * hands both to the reference's optimiser (scipy TRF).  x, f0, J: host or device. */
int sfm_ba_reference_fd(sfm_ctx* ctx, const double* x, int n_params, int n_points, double* f0, double* J);

/* The linear solve of one LM step on its own (csrc/solve.cu): S x = -g for the reduced camera system of
 * ba.bundle_adjustment's normal equations — S symmetric positive definite, given as the lower triangle of 6x6 camera
 * blocks (block (a, b), b <= a, = 36 row-major float32 at ((a (a+1)) / 2 + b) * 36), g (6 n_cams) float32, x (6 n_cams)
 * float64, *info = 0 or the 1-based index of the first non-positive pivot.  Host arrays.  The LM step calls the same
 * solver on device buffers; this entry exists so that it can be checked against a dense factorisation directly. */
int sfm_reduced_solve(sfm_ctx* ctx, const float* S_blocks, const float* g, int n_cams, double* x, int32_t* info);

/* The same system by the solver the LM step uses by default (csrc/pcg.cu): block-Jacobi preconditioned conjugate
 * gradients to a relative residual of 1e-8, one persistent kernel.  *solved = 1: x is the solution; 0: the system was
 * not positive definite or did not converge (the LM step then falls back to the factorisation above).  *iterations
 * (may be NULL): iterations used.  Host arrays. */
int sfm_reduced_solve_pcg(sfm_ctx* ctx, const float* S_blocks, const float* g, int n_cams, double* x, int32_t* solved,
                          int32_t* iterations);

typedef struct sfm_ba_stats {
  double cost_before;   /* 0.5*sum r^2 at the linearisation point */
  double cost_after;    /* at the candidate parameters */
  double step_norm;
  double grad_norm;
  int32_t accepted;
  int32_t solve_info;   /* 0 ok; >0 Cholesky failed at that pivot */
  double lambda_next;
} sfm_ba_stats;

/* One damped Gauss-Newton (LM) iteration: K6 fused JtJ/Schur accumulation (no Jacobian in HBM),
 * [all-reduce of the reduced camera system if a communicator is attached], the 6C x 6C system by
 * block-Jacobi preconditioned conjugate gradients to a relative residual of 1e-5 (environment
 * SFM_BA_CG_TOL overrides; the system is float32 data) with the tile Cholesky as fallback and for
 * systems under 600 unknowns, back-substitution, parameter update, new cost; rejected steps are
 * rolled back. */
int sfm_ba_gn_step(sfm_ba* ba, double lambda, sfm_ba_stats* stats);

/* Device views for tests and for a host-side exchange step (float32, row-major):
 * which 0: S (6C x 6C, lower block triangle filled, damped), 1: g (6C), 2: diag(Hcc) (6C). */
int sfm_ba_build_system(sfm_ba* ba, double lambda);
int sfm_ba_read(sfm_ba* ba, int which, float* out, int64_t count);

/* C1 — multi-GPU: points are sharded over ranks by the host; the only exchange per iteration is
 * the all-reduce(sum) of [S | g | cost].  The communicator is created inside the library from an
 * ncclUniqueId that the host distributes (torch.distributed does the rendezvous plumbing). */
int sfm_nccl_unique_id(void* out128);
int sfm_ba_comm_init(sfm_ba* ba, const void* unique_id128, int rank, int world);
int sfm_ba_comm_destroy(sfm_ba* ba);

#ifdef __cplusplus
}
#endif
#endif /* SFM_B200_H */
