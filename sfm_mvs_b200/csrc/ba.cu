// ba.cu — hot path 3b: bundle adjustment (K5 residual/Jacobian evaluation, K6 fused Schur
// accumulation, back-substitution / update; the reduced-camera-system solve is solve.cu).
//
// Reference: the residual every BA variant of the reference evaluates is
//   cv2.projectPoints(X, rvec, tvec, K, None) - observation
// (OptimReprojectionError sfm.py:104-136 -> projectPoints sfm.py:121; test.py:85-113 -> :101;
// ba.pyc L24/L51), driven by scipy.least_squares with a dense finite-difference Jacobian
// (sfm.py:146).  The block structure (one 2x6 camera block and one 2x3 point block per
// observation) is the notebook's bundle_adjustment_sparsity (cell 6).  This engine evaluates the
// same residual with analytic Jacobians and solves the Gauss-Newton / Levenberg-Marquardt step by
// point elimination (Schur complement).
//
// Data layout in HBM (struct sfm_ba):
//   obs_uv  float2[O]   observations, sorted point-major      cam_idx int[O]     pt_idx int[O]
//   pt_start int[P+1]   CSR offsets of each point's observations
//   cams double[C][6] (rvec|tvec)   pts double[P][3]   (+ candidate copies for step rejection)
//   cam_pre 144-byte records [C]: R (9) | t (3) float64, Jl (9 of 12) float32 — rotation matrix and the
//                          left Jacobian of SO(3), recomputed once per linearisation by ba_cam_prep_kernel
//   S float[C (C+1) / 2][36] reduced camera system: the lower triangle of 6x6 blocks, block (a, b), b <= a, at
//                          (a (a+1) / 2 + b) * 36, row-major inside the block; g float[6C]; hdiag float[6C] = diag(Hcc)
//   Tbuf float[O][18], qp double[P][3]   T_a = W_a Hpp^-1 and -Hpp^-1 bp of the current linearisation: written by the
//                          Schur kernel, streamed by the back substitution
// Per observation:  Yr = R X,  Y = Yr + t,  (u,v) = (fx Y.x/Y.z + cx, fy Y.y/Y.z + cy)
//   d(u,v)/dY = [[fx/z, 0, -fx Y.x/z^2], [0, fy/z, -fy Y.y/z^2]]
//   Jc[:,0:3] = d(u,v)/dY * [ Jl e_k x Yr ]_k   (since dR/dr_k = [Jl e_k]x R),  Jc[:,3:6] = d(u,v)/dY
//   Jp = d(u,v)/dY * R
#include <float.h>
#include <stdlib.h>
#include <math.h>

#include "ba.cuh"
#include "hostmath.h"
#include "solve.cuh"

namespace {

constexpr int CAM_PRE = 18;   // doubles per camera record: R (9) | t (3) as float64, Jl (9, padded to 12) as float32 = 144 B
constexpr int BA_MAXO = 64;   // observations per point handled by the fused kernels

// ------------------------------------------------------------------ per-camera precomputation
__global__ void ba_cam_prep_kernel(const double* __restrict__ cams, int n_cam, double* __restrict__ pre) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cam) return;
  const double* rv = cams + 6 * (size_t)c;
  double* o = pre + CAM_PRE * (size_t)c;
  double R[9];
  hm::rodrigues_to_matrix(rv, R);
  for (int k = 0; k < 9; ++k) o[k] = R[k];
  o[9] = rv[3]; o[10] = rv[4]; o[11] = rv[5];
  float* jl = reinterpret_cast<float*>(o + 12);
  double th2 = rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2];
  for (int k = 0; k < 3; ++k) {
    double a[3];
    if (th2 < 1e-24) {
      a[0] = k == 0; a[1] = k == 1; a[2] = k == 2;
    } else {
      // Jl e_k = ( r_k r + r x (I - R) e_k ) / |r|^2     (Gallego & Yezzi 2015, eq. III.7)
      double m[3] = {(k == 0) - R[k], (k == 1) - R[3 + k], (k == 2) - R[6 + k]};
      double cx = rv[1] * m[2] - rv[2] * m[1], cy = rv[2] * m[0] - rv[0] * m[2], cz = rv[0] * m[1] - rv[1] * m[0];
      a[0] = (rv[k] * rv[0] + cx) / th2; a[1] = (rv[k] * rv[1] + cy) / th2; a[2] = (rv[k] * rv[2] + cz) / th2;
    }
    jl[3 * k] = (float)a[0]; jl[3 * k + 1] = (float)a[1]; jl[3 * k + 2] = (float)a[2];
  }
  jl[9] = jl[10] = jl[11] = 0.f;
}

struct Intr {
  double fx, fy, cx, cy;
};

// residual (proj - obs) and, if WANT_J, the Jacobian blocks, in registers.  The camera record is read
// with 16-byte loads (6 x double2 for R|t, 3 x float4 for Jl): from shared memory a 16-byte access is
// served per quarter-warp, which keeps the random per-lane camera gather to ~2-3 wavefronts per load.
template <bool WANT_J>
__device__ __forceinline__ void obs_geometry(const double* __restrict__ cp, const double* __restrict__ X, float2 uv,
                                             const Intr& K, double* r, float (*Jc)[6], float (*Jp)[3]) {
  const double2* q = reinterpret_cast<const double2*>(cp);
  const double2 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4], q5 = q[5];
  // R = [q0.x q0.y q1.x; q1.y q2.x q2.y; q3.x q3.y q4.x], t = (q4.y, q5.x, q5.y)
  const double Yr0 = q0.x * X[0] + q0.y * X[1] + q1.x * X[2];
  const double Yr1 = q1.y * X[0] + q2.x * X[1] + q2.y * X[2];
  const double Yr2 = q3.x * X[0] + q3.y * X[1] + q4.x * X[2];
  const double y0 = Yr0 + q4.y, y1 = Yr1 + q5.x, y2 = Yr2 + q5.y;
  const double iz = (y2 != 0.0) ? 1.0 / y2 : 1.0;
  const double xn = y0 * iz, yn = y1 * iz;
  r[0] = xn * K.fx + K.cx - (double)uv.x;
  r[1] = yn * K.fy + K.cy - (double)uv.y;
  if (WANT_J) {
    // Jacobian blocks are delivered as float32, so they are formed in float32 from the float64
    // projection state (relative error ~1e-7); only the residual path needs float64.
    const float4* f = reinterpret_cast<const float4*>(cp + 12);
    const float4 f0 = f[0], f1 = f[1], f2 = f[2];
    const float jl[9] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x};
    const float fiz = (float)iz, fxn = (float)xn, fyn = (float)yn;
    const float a0 = (float)K.fx * fiz, a2 = -a0 * fxn, b1 = (float)K.fy * fiz, b2 = -b1 * fyn;
    const float yr0 = (float)Yr0, yr1 = (float)Yr1, yr2 = (float)Yr2;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float ax = jl[3 * k], ay = jl[3 * k + 1], az = jl[3 * k + 2];
      const float dx = ay * yr2 - az * yr1, dy = az * yr0 - ax * yr2, dz = ax * yr1 - ay * yr0;
      Jc[0][k] = a0 * dx + a2 * dz;
      Jc[1][k] = b1 * dy + b2 * dz;
    }
    Jc[0][3] = a0; Jc[0][4] = 0.f; Jc[0][5] = a2;
    Jc[1][3] = 0.f; Jc[1][4] = b1; Jc[1][5] = b2;
    const float r00 = (float)q0.x, r01 = (float)q0.y, r02 = (float)q1.x, r10 = (float)q1.y, r11 = (float)q2.x,
                r12 = (float)q2.y, r20 = (float)q3.x, r21 = (float)q3.y, r22 = (float)q4.x;
    Jp[0][0] = a0 * r00 + a2 * r20; Jp[0][1] = a0 * r01 + a2 * r21; Jp[0][2] = a0 * r02 + a2 * r22;
    Jp[1][0] = b1 * r10 + b2 * r20; Jp[1][1] = b1 * r11 + b2 * r21; Jp[1][2] = b1 * r12 + b2 * r22;
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ K5: materialised evaluation
// One persistent CTA of 512 threads per SM (large problems); one observation per thread per trip: 16 B read (uv,
// cam index, point index), 80 B written (r 8 B, Jc 48 B, Jp 24 B).  512 threads x 82 registers measured faster
// than 768 x 72 or 1024 x 64 (28.0 / 31.0 / 33.2 us at 1 M observations): the kernel is bound by the shared-memory
// and store queues, which fewer, longer-lived warps load more evenly, not by latency that more warps would hide.  The per-camera table (C x 144 B,
// 72 KB at C = 500) is staged into shared memory by ONE bulk-copy instruction (cp.async.bulk, the TMA
// engine, completion on an mbarrier) when it fits, so the random per-observation camera gather never
// leaves the SM; otherwise it is read through L1.
__device__ __forceinline__ uint32_t ba_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE, bool SMEM_CAMS, int THREADS = 1024, bool PREFETCH_PT = false>
__global__ void __launch_bounds__(THREADS, 1) ba_eval_kernel(const float2* __restrict__ uv, const int* __restrict__ cam_idx,
                                                          const int* __restrict__ pt_idx, int n_obs,
                                                          const double* __restrict__ cam_pre, int n_cam,
                                                          const double* __restrict__ pts, Intr K, double inv_n,
                                                          float* __restrict__ r_out, float* __restrict__ Jc_out,
                                                          float* __restrict__ Jp_out, double* __restrict__ cost,
                                                          unsigned long long* __restrict__ tl = nullptr) {
  extern __shared__ __align__(128) double s_cam[];
  __shared__ __align__(8) unsigned long long s_bar;
  auto stamp = [&](int k) {          // diagnostics (SFM_BA_EVAL_TIMELINE): %globaltimer per CTA at the phase boundaries
    if (tl && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      tl[5 * blockIdx.x + k] = t;
    }
  };
  stamp(0);
  if (SMEM_CAMS) {
    const uint32_t bar = ba_smem_u32(&s_bar);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const uint32_t bytes = (uint32_t)((n_cam * CAM_PRE * sizeof(double) + 15) & ~(size_t)15);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(ba_smem_u32(s_cam)), "l"(cam_pre), "r"(bytes), "r"(bar) : "memory");
    }
    __syncthreads();   // barrier initialised before anyone polls it
  }
  const double* cams = SMEM_CAMS ? s_cam : cam_pre;
  double local = 0.0;
  // Software pipeline: the (uv, camera index, point index) triple of the next trip is loaded before the
  // current observation is processed, so its DRAM latency overlaps this trip's gather + arithmetic.
  const int stride = gridDim.x * blockDim.x;
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  float2 m_n = make_float2(0.f, 0.f);
  int c_n = 0, p_n = 0;
  if (o < n_obs) { m_n = __ldg(uv + o); c_n = __ldg(cam_idx + o); p_n = __ldg(pt_idx + o); }
  // PREFETCH_PT: the point of the next trip is loaded a trip ahead as well (its index two trips ahead), so no
  // DRAM / L2 latency is left on the per-trip dependent chain.
  double X_n[3] = {0.0, 0.0, 0.0};
  if (PREFETCH_PT) {
    if (o < n_obs) { X_n[0] = __ldg(pts + 3 * (size_t)p_n); X_n[1] = __ldg(pts + 3 * (size_t)p_n + 1); X_n[2] = __ldg(pts + 3 * (size_t)p_n + 2); }
    if (o + stride < n_obs) p_n = __ldg(pt_idx + o + stride);
  }
  if (SMEM_CAMS) {           // wait for the bulk copy only now: the first index loads are already in flight
    const uint32_t bar = ba_smem_u32(&s_bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar) : "memory");
      if (spin > (1u << 22)) asm volatile("trap;");
    }
  }
  stamp(1);
  bool first_trip = true;
  for (; o < n_obs; o += stride) {
    const float2 m = m_n;
    const int c = c_n, p = p_n;
    double X[3];
    if (PREFETCH_PT) {
      X[0] = X_n[0]; X[1] = X_n[1]; X[2] = X_n[2];
      if (o + stride < n_obs) {     // p already is the NEXT trip's point index
        X_n[0] = __ldg(pts + 3 * (size_t)p); X_n[1] = __ldg(pts + 3 * (size_t)p + 1); X_n[2] = __ldg(pts + 3 * (size_t)p + 2);
        m_n = __ldg(uv + o + stride); c_n = __ldg(cam_idx + o + stride);
      }
      if (o + 2 * (size_t)stride < (size_t)n_obs) p_n = __ldg(pt_idx + o + 2 * stride);
    } else {
      X[0] = __ldg(pts + 3 * (size_t)p); X[1] = __ldg(pts + 3 * (size_t)p + 1); X[2] = __ldg(pts + 3 * (size_t)p + 2);
      if (o + stride < n_obs) { m_n = __ldg(uv + o + stride); c_n = __ldg(cam_idx + o + stride); p_n = __ldg(pt_idx + o + stride); }
    }
    double r[2];
    float Jc[2][6], Jp[2][3];
    if (MODE == 0) {
      if (Jc_out != nullptr || Jp_out != nullptr) obs_geometry<true>(cams + CAM_PRE * (size_t)c, X, m, K, r, Jc, Jp);
      else obs_geometry<false>(cams + CAM_PRE * (size_t)c, X, m, K, r, Jc, Jp);
      local += r[0] * r[0] + r[1] * r[1];
      if (r_out) reinterpret_cast<float2*>(r_out)[o] = make_float2((float)r[0], (float)r[1]);
      if (Jc_out) {
        float4* d = reinterpret_cast<float4*>(Jc_out + 12 * (size_t)o);
        d[0] = make_float4((float)Jc[0][0], (float)Jc[0][1], (float)Jc[0][2], (float)Jc[0][3]);
        d[1] = make_float4((float)Jc[0][4], (float)Jc[0][5], (float)Jc[1][0], (float)Jc[1][1]);
        d[2] = make_float4((float)Jc[1][2], (float)Jc[1][3], (float)Jc[1][4], (float)Jc[1][5]);
      }
      if (Jp_out) {
        float2* d = reinterpret_cast<float2*>(Jp_out + 6 * (size_t)o);
        d[0] = make_float2((float)Jp[0][0], (float)Jp[0][1]);
        d[1] = make_float2((float)Jp[0][2], (float)Jp[1][0]);
        d[2] = make_float2((float)Jp[1][1], (float)Jp[1][2]);
      }
    } else {
      obs_geometry<false>(cams + CAM_PRE * (size_t)c, X, m, K, r, Jc, Jp);
      if (MODE == 1) {        // ((p - proj)^2).ravel()/N            sfm.py:124-130
        double a = r[0] * r[0] * inv_n, b = r[1] * r[1] * inv_n;
        local += a * a + b * b;
        if (r_out) reinterpret_cast<float2*>(r_out)[o] = make_float2((float)a, (float)b);
      } else {                // sqrt(dx^2+dy^2)/len(error)            test.py:108-112
        double a = sqrt(r[0] * r[0] + r[1] * r[1]) * inv_n;
        local += a * a;
        if (r_out) r_out[o] = (float)a;
      }
    }
    if (first_trip) { stamp(2); first_trip = false; }
  }
  stamp(3);
  if (cost) {
    local = warp_sum_d(local);
    __shared__ double s_part[32];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += s_part[w];
      atomicAdd(cost, 0.5 * s);
    }
  }
  stamp(4);
}

// ------------------------------------------------------------------ K6: fused Schur accumulation
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// S is stored by 6x6 camera blocks: block (ca, cb) is 36 contiguous floats (row-major inside), so a block's
// contribution is 9 16-byte reductions into 144 contiguous bytes instead of 18 8-byte ones into 6 rows 12 KB apart
// (K6 is bound by the number of reductions the L2 slices retire).
__device__ __forceinline__ void red_add_block36(float* blk, const float* v) {
#pragma unroll
  for (int q = 0; q < 9; ++q) red_add_v4(blk + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// 3x3 symmetric (h = xx,xy,xz,yy,yz,zz) with damped diagonal -> inverse (same packing)
__device__ __forceinline__ void inv_sym3(const double* h, double lambda, double* inv) {
  const double a = h[0] * (1.0 + lambda), b = h[1], c = h[2], d = h[3] * (1.0 + lambda), e = h[4], f = h[5] * (1.0 + lambda);
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double det = a * c00 + b * c01 + c * c02;
  const double id = (det != 0.0) ? 1.0 / det : 0.0;
  inv[0] = c00 * id; inv[1] = c01 * id; inv[2] = c02 * id;
  inv[3] = (a * f - c * c) * id; inv[4] = (b * c - a * e) * id; inv[5] = (a * d - b * b) * id;
}

template <int CAP>
struct WarpPoint {            // per-warp shared scratch of the fused kernels, for points with up to CAP observations
  float W[CAP][18];           // W_a = Jc^T Jp   (6x3)
  float T[CAP][18];           // T_a = W_a * Hpp^-1
  float J[CAP][12];           // Jc of the observation (2x6): its Hcc = Jc^T Jc joins the (a, a) block update of the Schur kernel
  float gc[CAP][6];           // Jc^T r of the observation: joins the one reduction of g per observation
  int cam[CAP];
};

// Per point (one warp): accumulate Hpp/bp over its observations, per-observation W, Hcc/bc via
// atomics; returns (in every lane) Hpp (6), bp (3).  MODE_UPDATE additionally needs dc.
template <int CAP>
__device__ __forceinline__ void point_accumulate(WarpPoint<CAP>& wp, int lane, int o_begin, int nobs,
                                                 const float2* __restrict__ uv, const int* __restrict__ cam_idx,
                                                 const double* __restrict__ cams, const double* X, const Intr& K,
                                                 bool accumulate_cam, float* __restrict__ S, int ld,
                                                 float* __restrict__ g, float* __restrict__ hdiag,
                                                 double* hpp, double* bp, double* cost_local) {
  double h[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
  for (int a = lane; a < nobs; a += 32) {
    const int o = o_begin + a;
    const int c = __ldg(cam_idx + o);
    double r[2];
    float Jc[2][6], Jp[2][3];
    obs_geometry<true>(cams + CAM_PRE * (size_t)c, X, __ldg(uv + o), K, r, Jc, Jp);
    *cost_local += r[0] * r[0] + r[1] * r[1];
    h[0] += Jp[0][0] * Jp[0][0] + Jp[1][0] * Jp[1][0];
    h[1] += Jp[0][0] * Jp[0][1] + Jp[1][0] * Jp[1][1];
    h[2] += Jp[0][0] * Jp[0][2] + Jp[1][0] * Jp[1][2];
    h[3] += Jp[0][1] * Jp[0][1] + Jp[1][1] * Jp[1][1];
    h[4] += Jp[0][1] * Jp[0][2] + Jp[1][1] * Jp[1][2];
    h[5] += Jp[0][2] * Jp[0][2] + Jp[1][2] * Jp[1][2];
#pragma unroll
    for (int k = 0; k < 3; ++k) b[k] += Jp[0][k] * r[0] + Jp[1][k] * r[1];
    wp.cam[a] = c;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) wp.W[a][3 * i + k] = (float)(Jc[0][i] * Jp[0][k] + Jc[1][i] * Jp[1][k]);
    if (accumulate_cam) {
      // Hcc and bc of the observation are NOT reduced here: they ride on the (a, a) block update and on the g update
      // of the Schur kernel's later phases (one set of reductions per observation instead of two)
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        wp.J[a][i] = Jc[0][i];
        wp.J[a][6 + i] = Jc[1][i];
        wp.gc[a][i] = (float)(Jc[0][i] * r[0] + Jc[1][i] * r[1]);
      }
      red_add_v2(hdiag + 6 * c, Jc[0][0] * Jc[0][0] + Jc[1][0] * Jc[1][0], Jc[0][1] * Jc[0][1] + Jc[1][1] * Jc[1][1]);
      red_add_v2(hdiag + 6 * c + 2, Jc[0][2] * Jc[0][2] + Jc[1][2] * Jc[1][2], Jc[0][3] * Jc[0][3] + Jc[1][3] * Jc[1][3]);
      red_add_v2(hdiag + 6 * c + 4, Jc[0][4] * Jc[0][4] + Jc[1][4] * Jc[1][4], Jc[0][5] * Jc[0][5] + Jc[1][5] * Jc[1][5]);
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) hpp[k] = warp_sum_d(h[k]);
#pragma unroll
  for (int k = 0; k < 3; ++k) bp[k] = warp_sum_d(b[k]);
}

// Warps per CTA of the two per-point kernels (one CTA per SM: the camera table takes 72 KB of its shared memory).  Both
// are bound by latency — the index -> camera -> geometry chain of a point, the shared-memory round trips of its pairs
// (ncu: issue slots 28 % busy, 12 of 32 lanes active on average) — and the per-warp scratch is sized by the problem's
// largest point (CAP = 16 or 64 observations), so that more warps fit when the points are small.  Measured at 1 M
// observations, 10 per point: Schur 0.79 / 0.73 / 0.71 ms with 4 / 8 / 12 warps of CAP 64 and 0.63 ms with 12 or 20 of
// CAP 16 (it does not scale further: with the 55 block reductions of a point left out it takes 0.33 ms — the other
// half is the L2 retiring 9 sector reductions per 6x6 block); the update 0.33 ms at 4 warps, 0.16 at 12, 0.11 at 20.
template <int CAP> struct FusedWarps { static constexpr int N = 12; };
template <> struct FusedWarps<16> { static constexpr int N = 20; };

// S -= sum_p W Hpp^-1 W^T (lower block triangle), g += bc - W Hpp^-1 bp, diag blocks += Hcc.
template <bool SMEM_CAMS, int CAP>
__global__ void __launch_bounds__(FusedWarps<CAP>::N * 32) ba_schur_kernel(const float2* __restrict__ uv,
                                                                    const int* __restrict__ cam_idx,
                                                                    const int* __restrict__ pt_start, int n_pt,
                                                                    const double* __restrict__ cam_pre, int n_cam,
                                                                    const double* __restrict__ pts, Intr K,
                                                                    double lambda, float* __restrict__ S, int ld,
                                                                    float* __restrict__ g, float* __restrict__ hdiag,
                                                                    double* __restrict__ cost, float* __restrict__ Tbuf,
                                                                    double* __restrict__ qp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int WARPS = FusedWarps<CAP>::N;
  WarpPoint<CAP>* wps = reinterpret_cast<WarpPoint<CAP>*>(smem_raw);
  double* s_cam = reinterpret_cast<double*>(smem_raw + sizeof(WarpPoint<CAP>) * WARPS);
  if (SMEM_CAMS) {
    for (int i = threadIdx.x; i < n_cam * CAM_PRE; i += blockDim.x) s_cam[i] = cam_pre[i];
    __syncthreads();
  }
  const double* cams = SMEM_CAMS ? s_cam : cam_pre;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpPoint<CAP>& wp = wps[warp];
  // (a, b), a <= b, of the q-th unordered pair, q = b (b + 1) / 2 + a: a table for small points
  __shared__ unsigned char s_tri[CAP <= 16 ? 2 * 136 : 2];
  if (CAP <= 16) {
    for (int q = threadIdx.x; q < 136; q += blockDim.x) {
      int b = 0;
      while ((b + 1) * (b + 2) / 2 <= q) ++b;
      s_tri[2 * q] = (unsigned char)(q - b * (b + 1) / 2);
      s_tri[2 * q + 1] = (unsigned char)b;
    }
    __syncthreads();
  }
  double cost_local = 0.0;
  for (int p = blockIdx.x * WARPS + warp; p < n_pt; p += gridDim.x * WARPS) {
    const int o_begin = __ldg(pt_start + p), nobs = __ldg(pt_start + p + 1) - o_begin;
    if (nobs <= 0) continue;
    const double X[3] = {__ldg(pts + 3 * (size_t)p), __ldg(pts + 3 * (size_t)p + 1), __ldg(pts + 3 * (size_t)p + 2)};
    double hpp[6], bp[3], hinv[6];
    point_accumulate(wp, lane, o_begin, nobs, uv, cam_idx, cams, X, K, true, S, ld, g, hdiag, hpp, bp, &cost_local);
    inv_sym3(hpp, lambda, hinv);
    if (qp && lane == 0) {
      qp[3 * (size_t)p] = -(hinv[0] * bp[0] + hinv[1] * bp[1] + hinv[2] * bp[2]);
      qp[3 * (size_t)p + 1] = -(hinv[1] * bp[0] + hinv[3] * bp[1] + hinv[4] * bp[2]);
      qp[3 * (size_t)p + 2] = -(hinv[2] * bp[0] + hinv[4] * bp[1] + hinv[5] * bp[2]);
    }
    __syncwarp();
    // T_a = W_a Hinv ; g[ca] -= T_a bp
    for (int a = lane; a < nobs; a += 32) {
      float t[18];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double w0 = wp.W[a][3 * i], w1 = wp.W[a][3 * i + 1], w2 = wp.W[a][3 * i + 2];
        t[3 * i] = (float)(w0 * hinv[0] + w1 * hinv[1] + w2 * hinv[2]);
        t[3 * i + 1] = (float)(w0 * hinv[1] + w1 * hinv[3] + w2 * hinv[4]);
        t[3 * i + 2] = (float)(w0 * hinv[2] + w1 * hinv[4] + w2 * hinv[5]);
      }
#pragma unroll
      for (int k = 0; k < 18; ++k) wp.T[a][k] = t[k];
      if (Tbuf) {                      // kept for the back substitution: dp = -Hinv bp - sum_a T_a^T dc[cam_a]
        float2* dst = reinterpret_cast<float2*>(Tbuf + 18 * (size_t)(o_begin + a));
#pragma unroll
        for (int k = 0; k < 9; ++k) dst[k] = make_float2(t[2 * k], t[2 * k + 1]);
      }
      float gc[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) gc[i] = wp.gc[a][i] - (float)(t[3 * i] * bp[0] + t[3 * i + 1] * bp[1] + t[3 * i + 2] * bp[2]);
      float* gp = g + 6 * wp.cam[a];
      red_add_v2(gp, gc[0], gc[1]);
      red_add_v2(gp + 2, gc[2], gc[3]);
      red_add_v2(gp + 4, gc[4], gc[5]);
    }
    __syncwarp();
    // Unordered observation pairs a <= b, oriented by camera index: block (c_hi, c_lo) -= T_hi W_lo^T.  NINE lanes per
    // pair, three pairs per step: lane k of a group forms elements 4k .. 4k+3 of the 6x6 block and issues ONE 16-byte
    // reduction, so the nine reductions of a block come from one instruction and reach L2 as the 5 sectors the block
    // spans — the kernel is bound by the sector reductions L2 retires (9 per block when every lane updates its own
    // block: 0.59 ms, 0.33 ms with the block reductions left out).
    // Two observations of ONE camera (a != b, equal cameras) contribute both orientations to the same diagonal block.
    const int ntri = nobs * (nobs + 1) / 2;
    const int grp = lane / 9, kq = lane - 9 * grp;
    // this lane's four elements m0 .. m0+3 of the row-major 6x6 block: rows i0 (and i0 + 1 when the run crosses a row
    // end), columns j0, j0+1, ... mod 6 with j0 in {0, 4, 2} — even, so the four W rows it needs (3 floats each,
    // wrapping from row 5 to row 0) are six aligned 8-byte reads
    const int m0 = 4 * kq, i0 = m0 / 6, j0 = m0 - 6 * i0;
    const int i1 = i0 < 5 ? i0 + 1 : i0;
    if (grp < 3) {
      for (int t0 = 0; t0 < ntri; t0 += 3) {
        const int q = t0 + grp;
        if (q < ntri) {
          int a, b;
          if (CAP <= 16) {
            a = s_tri[2 * q]; b = s_tri[2 * q + 1];
          } else {
            b = (int)((sqrtf(8.f * (float)q + 1.f) - 1.f) * 0.5f);
            while ((b + 1) * (b + 2) / 2 <= q) ++b;
            while (b * (b + 1) / 2 > q) --b;
            a = q - b * (b + 1) / 2;
          }
          const int ca = wp.cam[a], cb = wp.cam[b];
          const bool sw = ca < cb;
          int hi = sw ? b : a, lo = sw ? a : b;
          const int chi = sw ? cb : ca, clo = sw ? ca : cb;
          float* Sd = S + ((size_t)chi * (chi + 1) / 2 + clo) * 36 + 4 * kq;
          const int reps = (a != b && ca == cb) ? 2 : 1;
          for (int rep = 0; rep < reps; ++rep) {
            const float* Ta = wp.T[hi];
            const float* Wb = wp.W[lo];
            float w[12];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
              int idx = 3 * j0 + 2 * k;
              idx = idx >= 18 ? idx - 18 : idx;
              const float2 t2 = *reinterpret_cast<const float2*>(Wb + idx);
              w[2 * k] = t2.x; w[2 * k + 1] = t2.y;
            }
            const float ta0 = Ta[3 * i0], ta1 = Ta[3 * i0 + 1], ta2 = Ta[3 * i0 + 2];
            const float tb0 = Ta[3 * i1], tb1 = Ta[3 * i1 + 1], tb2 = Ta[3 * i1 + 2];
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const bool nxt = j0 + e >= 6;                   // this element lies in row i0 + 1
              const float x0 = nxt ? tb0 : ta0, x1 = nxt ? tb1 : ta1, x2 = nxt ? tb2 : ta2;
              v[e] = -(x0 * w[3 * e] + x1 * w[3 * e + 1] + x2 * w[3 * e + 2]);
            }
            if (a == b) {                                       // + Hcc = Jc^T Jc of the observation
              const float* Ja = wp.J[a];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int m = m0 + e, i = m / 6, j = m - 6 * i;
                v[e] += Ja[i] * Ja[j] + Ja[6 + i] * Ja[6 + j];
              }
            }
            red_add_v4(Sd, v[0], v[1], v[2], v[3]);
            const int tmp = hi; hi = lo; lo = tmp;
          }
        }
      }
    }
    __syncwarp();
  }
  cost_local = warp_sum_d(cost_local);
  if (lane == 0 && cost) atomicAdd(cost, 0.5 * cost_local);
}

// S[ii] += lambda * Hcc[ii]
__global__ void ba_damp_kernel(float* __restrict__ S, int ld, const float* __restrict__ hdiag, double lambda, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) S[((size_t)(i / 6) * (i / 6 + 1) / 2 + i / 6) * 36 + 7 * (i % 6)] += (float)(lambda * (double)hdiag[i]);
}

// ------------------------------------------------------------------ back-substitution + update
// dp = -Hpp_d^-1 (bp + sum_a W_a^T dc[cam_a]);  candidate point = point + dp
template <bool SMEM_CAMS, int CAP>
__global__ void __launch_bounds__(FusedWarps<CAP>::N * 32) ba_update_points_kernel(
    const float2* __restrict__ uv, const int* __restrict__ cam_idx, const int* __restrict__ pt_start, int n_pt,
    const double* __restrict__ cam_pre, int n_cam, const double* __restrict__ pts, Intr K, double lambda,
    const double* __restrict__ dc, double* __restrict__ pts_new, double* __restrict__ step2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int WARPS = FusedWarps<CAP>::N;
  WarpPoint<CAP>* wps = reinterpret_cast<WarpPoint<CAP>*>(smem_raw);
  double* s_cam = reinterpret_cast<double*>(smem_raw + sizeof(WarpPoint<CAP>) * WARPS);
  if (SMEM_CAMS) {
    for (int i = threadIdx.x; i < n_cam * CAM_PRE; i += blockDim.x) s_cam[i] = cam_pre[i];
    __syncthreads();
  }
  const double* cams = SMEM_CAMS ? s_cam : cam_pre;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpPoint<CAP>& wp = wps[warp];
  double step_local = 0.0;
  for (int p = blockIdx.x * WARPS + warp; p < n_pt; p += gridDim.x * WARPS) {
    const int o_begin = __ldg(pt_start + p), nobs = __ldg(pt_start + p + 1) - o_begin;
    const double X[3] = {__ldg(pts + 3 * (size_t)p), __ldg(pts + 3 * (size_t)p + 1), __ldg(pts + 3 * (size_t)p + 2)};
    if (nobs <= 0) {
      if (lane < 3) pts_new[3 * (size_t)p + lane] = X[lane];
      continue;
    }
    double hpp[6], bp[3], hinv[6], dummy = 0.0;
    point_accumulate(wp, lane, o_begin, nobs, uv, cam_idx, cams, X, K, false, nullptr, 0, nullptr, nullptr, hpp, bp, &dummy);
    inv_sym3(hpp, lambda, hinv);
    __syncwarp();
    double v[3] = {0, 0, 0};
    for (int a = lane; a < nobs; a += 32) {
      const double* d = dc + 6 * wp.cam[a];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        v[0] += (double)wp.W[a][3 * i] * d[i];
        v[1] += (double)wp.W[a][3 * i + 1] * d[i];
        v[2] += (double)wp.W[a][3 * i + 2] * d[i];
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = bp[k] + warp_sum_d(v[k]);
    const double dp[3] = {-(hinv[0] * v[0] + hinv[1] * v[1] + hinv[2] * v[2]),
                          -(hinv[1] * v[0] + hinv[3] * v[1] + hinv[4] * v[2]),
                          -(hinv[2] * v[0] + hinv[4] * v[1] + hinv[5] * v[2])};
    if (lane == 0) {
      pts_new[3 * (size_t)p] = X[0] + dp[0];
      pts_new[3 * (size_t)p + 1] = X[1] + dp[1];
      pts_new[3 * (size_t)p + 2] = X[2] + dp[2];
      step_local += dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2];
    }
    __syncwarp();
  }
  if (lane == 0 && step2 && step_local != 0.0) atomicAdd(step2, step_local);
}

// The same back substitution from what the Schur kernel left behind: T_a = W_a Hpp^-1 of every observation (18 float32)
// and q_p = -Hpp^-1 bp of every point, so that dp = q_p - sum_a T_a^T dc[cam_a] is a stream over 72 bytes per observation
// (one thread per point: its observations are contiguous) instead of a second evaluation of the geometry and its
// Jacobians.  dc (6 C doubles) in shared memory.
__global__ void __launch_bounds__(256) ba_update_points_lin_kernel(const int* __restrict__ cam_idx, const int* __restrict__ pt_start,
                                                                   int n_pt, const float* __restrict__ Tbuf,
                                                                   const double* __restrict__ qp, const double* __restrict__ pts,
                                                                   const double* __restrict__ dc, int n_cam, int dc_in_smem,
                                                                   double* __restrict__ pts_new, double* __restrict__ step2) {
  extern __shared__ __align__(16) double s_dc[];
  if (dc_in_smem) {
    for (int i = threadIdx.x; i < 6 * n_cam; i += blockDim.x) s_dc[i] = dc[i];
    __syncthreads();
  }
  const double* dcs = dc_in_smem ? s_dc : dc;
  double local = 0.0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pt; p += gridDim.x * blockDim.x) {
    const int o0 = __ldg(pt_start + p), o1 = __ldg(pt_start + p + 1);
    const double X0 = __ldg(pts + 3 * (size_t)p), X1 = __ldg(pts + 3 * (size_t)p + 1), X2 = __ldg(pts + 3 * (size_t)p + 2);
    if (o1 <= o0) {
      pts_new[3 * (size_t)p] = X0; pts_new[3 * (size_t)p + 1] = X1; pts_new[3 * (size_t)p + 2] = X2;
      continue;
    }
    double v0 = qp[3 * (size_t)p], v1 = qp[3 * (size_t)p + 1], v2 = qp[3 * (size_t)p + 2];
    for (int o = o0; o < o1; ++o) {
      const float2* T2 = reinterpret_cast<const float2*>(Tbuf + 18 * (size_t)o);
      const double* d = dcs + 6 * __ldg(cam_idx + o);
      float t[18];
#pragma unroll
      for (int k = 0; k < 9; ++k) { const float2 q = __ldg(T2 + k); t[2 * k] = q.x; t[2 * k + 1] = q.y; }
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        v0 -= (double)t[3 * i] * d[i];
        v1 -= (double)t[3 * i + 1] * d[i];
        v2 -= (double)t[3 * i + 2] * d[i];
      }
    }
    pts_new[3 * (size_t)p] = X0 + v0; pts_new[3 * (size_t)p + 1] = X1 + v1; pts_new[3 * (size_t)p + 2] = X2 + v2;
    local += v0 * v0 + v1 * v1 + v2 * v2;
  }
  local = warp_sum_d(local);
  if ((threadIdx.x & 31) == 0 && step2 && local != 0.0) atomicAdd(step2, local);
}

__global__ void ba_update_cams_kernel(const double* __restrict__ cams, const double* __restrict__ dc, int n,
                                      double* __restrict__ cams_new, double* __restrict__ step2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double d = 0.0;
  if (i < n) { d = dc[i]; cams_new[i] = cams[i] + d; }
  d = warp_sum_d(d * d);
  if ((threadIdx.x & 31) == 0 && d != 0.0) atomicAdd(step2, d);
}

// Conjugate gradients first (pcg.cu: 20 - 40 iterations of an L2-resident matrix-vector product instead of a chain of 6C
// dependent columns); the tile Cholesky runs behind it only if it reports failure (its kernels return at once otherwise).
// SFM_BA_SOLVER=cholesky selects the factorisation alone.
int solve_reduced_system(sfm_ba* ba) {
  const int n = 6 * ba->n_cam;
  static const bool chol_only = [] { const char* e = getenv("SFM_BA_SOLVER"); return e && e[0] == 'c'; }();
  // small systems (a few tile columns) are factored faster than iterated: the iteration has a floor of one grid barrier
  // and two L2 round trips whatever n is
  int min_n = 600;
  if (const char* e = getenv("SFM_PCG_MIN_N")) min_n = atoi(e);
  // An LM step is solved to a relative residual of 1e-5, not to the 1e-8 the solver reaches on its own
  // (sfm_reduced_solve_pcg): S and g are accumulated in float32, so the system itself is only known to ~1e-6, and a
  // damped step accepted on "the cost went down" gains nothing from digits beyond that — the iterations it saves
  // (about a third) are the largest single item of the step.  SFM_BA_CG_TOL overrides.
  double cg_tol = 1e-5;
  if (const char* e = getenv("SFM_BA_CG_TOL")) {
    const double t = atof(e);
    if (t > 0.0 && t < 1.0) cg_tol = t;
  }
  if (ba->pcg && !chol_only && n >= min_n) {
    SFM_TRY(sfm_spd_pcg(ba->ctx, ba->S, ba->g, n, ba->pcg, ba->dc, ba->info + 1, ba->info, ba->info + 2, cg_tol));
    return sfm_spd_solve(ba->ctx, ba->S, ba->g, n, ba->A64, ba->dc, ba->info, ba->info + 1);
  }
  return sfm_spd_solve(ba->ctx, ba->S, ba->g, n, ba->A64, ba->dc, ba->info);
}

Intr make_intr(const sfm_ba* ba) {
  Intr k = {ba->K[0], ba->K[4], ba->K[2], ba->K[5]};
  return k;
}

size_t cam_smem_bytes(const sfm_ba* ba) { return (size_t)ba->n_cam * CAM_PRE * sizeof(double); }

int cam_prep(sfm_ba* ba, const double* cams) {
  SFM_LAUNCH(ba->ctx, SFM_K_MISC, (ba_cam_prep_kernel<<<div_up(ba->n_cam, 128), 128, 0, ba->ctx->stream>>>(cams, ba->n_cam, ba->cam_pre)));
  return SFM_OK;
}

// evaluation at (cams already in cam_pre, pts): *cost_dev += 0.5 sum r^2
template <int MODE, bool SM>
int launch_eval_v(sfm_ba* ba, const double* pts, float* r, float* Jc, float* Jp, double* cost_dev) {
  sfm_ctx* ctx = ba->ctx;
  const size_t smem = SM ? ((cam_smem_bytes(ba) + 15) & ~(size_t)15) : 0;
  const double inv_n = 1.0 / (double)(ba->n_obs_total > 0 ? ba->n_obs_total : 1);
  const bool big = ba->n_obs >= 64 * 1024;
  const int threads = big ? 512 : 256;
  int grid = std::max(1, std::min(div_up(ba->n_obs, threads), ctx->sm_count * (SM ? 1 : 2)));
  if (SM) {
    static bool attr = false;
    if (!attr) {
      SFM_CUDA(cudaFuncSetAttribute(ba_eval_kernel<MODE, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      SFM_CUDA(cudaFuncSetAttribute(ba_eval_kernel<MODE, SM, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = true;
    }
  }
  unsigned long long* tl = nullptr;
  if (big && getenv("SFM_BA_EVAL_TIMELINE")) {
    SFM_TRY(ws_alloc_t(ctx, (size_t)5 * grid, &tl));
    SFM_CUDA(cudaMemsetAsync(tl, 0, sizeof(unsigned long long) * 5 * grid, ctx->stream));
  }
  if (big)
    SFM_LAUNCH(ctx, SFM_K_BA_EVAL, (ba_eval_kernel<MODE, SM, 512, true><<<grid, threads, smem, ctx->stream>>>(
                                       ba->uv, ba->cam_idx, ba->pt_idx, ba->n_obs, ba->cam_pre, ba->n_cam, pts,
                                       make_intr(ba), inv_n, r, Jc, Jp, cost_dev, tl)));
  else
    SFM_LAUNCH(ctx, SFM_K_BA_EVAL, (ba_eval_kernel<MODE, SM><<<grid, threads, smem, ctx->stream>>>(
                                       ba->uv, ba->cam_idx, ba->pt_idx, ba->n_obs, ba->cam_pre, ba->n_cam, pts,
                                       make_intr(ba), inv_n, r, Jc, Jp, cost_dev)));
  if (tl) {
    std::vector<unsigned long long> h((size_t)5 * grid);
    SFM_CUDA(cudaMemcpyAsync(h.data(), tl, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    SFM_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned long long t0 = ~0ull;
    for (int b = 0; b < grid; ++b) t0 = std::min(t0, h[5 * b]);
    const char* names[5] = {"CTA start", "camera table in shared memory", "first trip done", "last trip done", "end"};
    for (int k = 0; k < 5; ++k) {
      unsigned long long lo = ~0ull, hi = 0, sum = 0;
      for (int b = 0; b < grid; ++b) { const unsigned long long v = h[5 * b + k] - t0; lo = std::min(lo, v); hi = std::max(hi, v); sum += v; }
      fprintf(stderr, "[ba_eval timeline] %-30s min %6.2f  mean %6.2f  max %6.2f us after the first CTA started\n", names[k], lo * 1e-3,
              sum * 1e-3 / grid, hi * 1e-3);
    }
  }
  return SFM_OK;
}

template <int MODE>
int launch_eval(sfm_ba* ba, const double* pts, float* r, float* Jc, float* Jp, double* cost_dev) {
  if (cam_smem_bytes(ba) <= 200 * 1024 - 64) return launch_eval_v<MODE, true>(ba, pts, r, Jc, Jp, cost_dev);
  return launch_eval_v<MODE, false>(ba, pts, r, Jc, Jp, cost_dev);
}

template <int CAP>
int launch_schur(sfm_ba* ba, double lambda) {
  sfm_ctx* ctx = ba->ctx;
  constexpr int WARPS = FusedWarps<CAP>::N;
  const int n = 6 * ba->n_cam;
  const size_t wp_bytes = sizeof(WarpPoint<CAP>) * WARPS;
  const size_t smem_cams = cam_smem_bytes(ba);
  const bool in_smem = wp_bytes + smem_cams <= 216 * 1024;
  const int grid = std::max(1, std::min(div_up(ba->n_pt, WARPS), ctx->sm_count * (in_smem ? 1 : 4)));
  if (in_smem) {
    static bool attr = false;
    if (!attr) {
      SFM_CUDA(cudaFuncSetAttribute(ba_schur_kernel<true, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
      attr = true;
    }
    SFM_LAUNCH(ctx, SFM_K_BA_SCHUR, (ba_schur_kernel<true, CAP><<<grid, WARPS * 32, wp_bytes + smem_cams, ctx->stream>>>(
                                        ba->uv, ba->cam_idx, ba->pt_start, ba->n_pt, ba->cam_pre, ba->n_cam, ba->pts,
                                        make_intr(ba), lambda, ba->S, n, ba->g, ba->hdiag, ba->scal + 0, ba->Tbuf, ba->qp)));
  } else {
    static bool attr = false;
    if (!attr) {
      SFM_CUDA(cudaFuncSetAttribute(ba_schur_kernel<false, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      attr = true;
    }
    SFM_LAUNCH(ctx, SFM_K_BA_SCHUR, (ba_schur_kernel<false, CAP><<<grid, WARPS * 32, wp_bytes, ctx->stream>>>(
                                        ba->uv, ba->cam_idx, ba->pt_start, ba->n_pt, ba->cam_pre, ba->n_cam, ba->pts,
                                        make_intr(ba), lambda, ba->S, n, ba->g, ba->hdiag, ba->scal + 0, ba->Tbuf, ba->qp)));
  }
  return SFM_OK;
}

template <int CAP>
int launch_update(sfm_ba* ba, double lambda) {
  sfm_ctx* ctx = ba->ctx;
  constexpr int WARPS = FusedWarps<CAP>::N;
  const size_t wp_bytes = sizeof(WarpPoint<CAP>) * WARPS;
  const size_t smem_cams = cam_smem_bytes(ba);
  const bool in_smem = wp_bytes + smem_cams <= 216 * 1024;
  const int grid = std::max(1, std::min(div_up(ba->n_pt, WARPS), ctx->sm_count * (in_smem ? 1 : 4)));
  if (in_smem) {
    static bool attr = false;
    if (!attr) {
      SFM_CUDA(cudaFuncSetAttribute(ba_update_points_kernel<true, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
      attr = true;
    }
    SFM_LAUNCH(ctx, SFM_K_BA_UPDATE, (ba_update_points_kernel<true, CAP><<<grid, WARPS * 32, wp_bytes + smem_cams, ctx->stream>>>(
                                         ba->uv, ba->cam_idx, ba->pt_start, ba->n_pt, ba->cam_pre, ba->n_cam, ba->pts,
                                         make_intr(ba), lambda, ba->dc, ba->pts_new, ba->scal + 2)));
  } else {
    static bool attr = false;
    if (!attr) {
      SFM_CUDA(cudaFuncSetAttribute(ba_update_points_kernel<false, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      attr = true;
    }
    SFM_LAUNCH(ctx, SFM_K_BA_UPDATE, (ba_update_points_kernel<false, CAP><<<grid, WARPS * 32, wp_bytes, ctx->stream>>>(
                                         ba->uv, ba->cam_idx, ba->pt_start, ba->n_pt, ba->cam_pre, ba->n_cam, ba->pts,
                                         make_intr(ba), lambda, ba->dc, ba->pts_new, ba->scal + 2)));
  }
  return SFM_OK;
}

int build_system(sfm_ba* ba, double lambda) {
  sfm_ctx* ctx = ba->ctx;
  const int n = 6 * ba->n_cam;
  SFM_TRY(cam_prep(ba, ba->cams));
  // S | g | hdiag are one allocation (sys_f32) so a single memset / all-reduce covers them
  SFM_CUDA(cudaMemsetAsync(ba->S, 0, ba->sys_count * sizeof(float), ctx->stream));
  SFM_CUDA(cudaMemsetAsync(ba->scal, 0, 8 * sizeof(double), ctx->stream));
  if (ba->n_pt > 0) {
    if (ba->max_deg <= 16) SFM_TRY((launch_schur<16>(ba, lambda)));
    else SFM_TRY((launch_schur<BA_MAXO>(ba, lambda)));
  }
  // exchange step (C1): sum of the partial systems and of the cost over ranks
  SFM_TRY(sfm_ba_allreduce_system(ba));
  SFM_LAUNCH(ctx, SFM_K_MISC, (ba_damp_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(ba->S, n, ba->hdiag, lambda, n)));
  return SFM_OK;
}

int update_points(sfm_ba* ba, double lambda) {
  if (ba->n_pt == 0) return SFM_OK;
  static const bool recompute = getenv("SFM_BA_UPDATE_RECOMPUTE") != nullptr;
  if (ba->Tbuf && ba->qp && !recompute) {
    sfm_ctx* ctx = ba->ctx;
    const size_t smem = sizeof(double) * 6 * (size_t)ba->n_cam;
    const int in_smem = smem <= 48 * 1024;
    const int grid = std::max(1, std::min(div_up(ba->n_pt, 256), ctx->sm_count * 8));
    SFM_LAUNCH(ctx, SFM_K_BA_UPDATE, (ba_update_points_lin_kernel<<<grid, 256, in_smem ? smem : 0, ctx->stream>>>(
                                         ba->cam_idx, ba->pt_start, ba->n_pt, ba->Tbuf, ba->qp, ba->pts, ba->dc, ba->n_cam, in_smem,
                                         ba->pts_new, ba->scal + 2)));
    return SFM_OK;
  }
  if (ba->max_deg <= 16) return launch_update<16>(ba, lambda);
  return launch_update<BA_MAXO>(ba, lambda);
}

}  // namespace

// ============================================================================ C ABI
extern "C" int sfm_ba_create(sfm_ctx* ctx, int n_cam, int n_pt, int n_obs, const int32_t* cam_idx,
                             const int32_t* pt_idx, const float* obs, const double* K, sfm_ba** out) {
  SFM_REQUIRE(ctx && out && K, "sfm_ba_create: null argument");
  SFM_REQUIRE(n_cam >= 1 && n_pt >= 0 && n_obs >= 0, "sfm_ba_create: bad sizes");
  SFM_REQUIRE(n_obs == 0 || (cam_idx && pt_idx && obs), "sfm_ba_create: null observation arrays");
  SFM_REQUIRE(!sfm_is_device_ptr(cam_idx) && !sfm_is_device_ptr(pt_idx) && !sfm_is_device_ptr(obs),
              "sfm_ba_create: observation arrays must be host pointers (they are validated and indexed on the host)");
  std::vector<int> start((size_t)n_pt + 1, 0);
  for (int o = 0; o < n_obs; ++o) {
    SFM_REQUIRE(cam_idx[o] >= 0 && cam_idx[o] < n_cam, "sfm_ba_create: cam_idx[%d]=%d out of range", o, cam_idx[o]);
    SFM_REQUIRE(pt_idx[o] >= 0 && pt_idx[o] < n_pt, "sfm_ba_create: pt_idx[%d]=%d out of range", o, pt_idx[o]);
    SFM_REQUIRE(o == 0 || pt_idx[o] >= pt_idx[o - 1], "sfm_ba_create: observations must be sorted point-major (pt_idx non-decreasing)");
    start[(size_t)pt_idx[o] + 1]++;
  }
  int ba_max_deg = 0;
  for (int p = 0; p < n_pt; ++p) {
    ba_max_deg = std::max(ba_max_deg, start[(size_t)p + 1]);
    if (start[(size_t)p + 1] > BA_MAXO) {
      sfm_set_error("sfm_ba_create: point %d has %d observations; the fused kernels handle at most %d", p, start[(size_t)p + 1], BA_MAXO);
      return SFM_ERR_UNSUPPORTED;
    }
    start[(size_t)p + 1] += start[p];
  }
  SFM_CUDA(cudaSetDevice(ctx->device));
  sfm_ba* ba = new sfm_ba();
  ba->ctx = ctx; ba->n_cam = n_cam; ba->n_pt = n_pt; ba->n_obs = n_obs;
  ba->n_pt_total = n_pt; ba->n_obs_total = n_obs;
  ba->max_deg = ba_max_deg;
  memcpy(ba->K, K, 9 * sizeof(double));
  const int n = 6 * n_cam;
  // S: the lower block triangle only, block (ca, cb), cb <= ca, = 36 contiguous floats at ((ca (ca+1) / 2) + cb) * 36 —
  // half the bytes to clear, to all-reduce and to keep in L2
  const size_t s_floats = (size_t)n_cam * (n_cam + 1) / 2 * 36;
  ba->sys_count = s_floats + 2 * (size_t)n;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes ? bytes : 16); };
  A((void**)&ba->uv, sizeof(float2) * (size_t)n_obs);
  A((void**)&ba->cam_idx, sizeof(int) * (size_t)n_obs);
  A((void**)&ba->pt_idx, sizeof(int) * (size_t)n_obs);
  A((void**)&ba->pt_start, sizeof(int) * ((size_t)n_pt + 1));
  A((void**)&ba->cams, sizeof(double) * 6 * (size_t)n_cam);
  A((void**)&ba->cams_new, sizeof(double) * 6 * (size_t)n_cam);
  A((void**)&ba->pts, sizeof(double) * 3 * (size_t)n_pt);
  A((void**)&ba->pts_new, sizeof(double) * 3 * (size_t)n_pt);
  A((void**)&ba->cam_pre, sizeof(double) * CAM_PRE * (size_t)n_cam + 16);   // +16: bulk copies are 16-byte granular
  A((void**)&ba->S, sizeof(float) * ba->sys_count);
  A((void**)&ba->A64, sizeof(double) * sfm_spd_scratch_doubles(n));
  A((void**)&ba->dc, sizeof(double) * (size_t)n);
  A((void**)&ba->scal, sizeof(double) * 8);
  A((void**)&ba->info, 4 * sizeof(int));
  A((void**)&ba->Tbuf, sizeof(float) * 18 * (size_t)n_obs);
  A((void**)&ba->qp, sizeof(double) * 3 * (size_t)n_pt);
  if (sfm_spd_pcg_fits(ctx, n)) A((void**)&ba->pcg, sizeof(double) * sfm_pcg_scratch_doubles(n));
  if (e != cudaSuccess) {
    sfm_set_error("sfm_ba_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    sfm_ba_destroy(ba);
    return SFM_ERR_NOMEM;
  }
  ba->g = ba->S + s_floats;
  ba->hdiag = ba->g + n;
  cudaStream_t st = ctx->stream;
  if (n_obs) {
    SFM_CUDA(cudaMemcpyAsync(ba->uv, obs, sizeof(float2) * (size_t)n_obs, cudaMemcpyHostToDevice, st));
    SFM_CUDA(cudaMemcpyAsync(ba->cam_idx, cam_idx, sizeof(int) * (size_t)n_obs, cudaMemcpyHostToDevice, st));
    SFM_CUDA(cudaMemcpyAsync(ba->pt_idx, pt_idx, sizeof(int) * (size_t)n_obs, cudaMemcpyHostToDevice, st));
  }
  SFM_CUDA(cudaMemcpyAsync(ba->pt_start, start.data(), sizeof(int) * ((size_t)n_pt + 1), cudaMemcpyHostToDevice, st));
  SFM_CUDA(cudaMemsetAsync(ba->cams, 0, sizeof(double) * 6 * (size_t)n_cam, st));
  SFM_CUDA(cudaMemsetAsync(ba->pts, 0, sizeof(double) * 3 * (size_t)n_pt + (n_pt ? 0 : 16), st));
  SFM_CUDA(cudaStreamSynchronize(st));
  *out = ba;
  return SFM_OK;
}

extern "C" void sfm_ba_destroy(sfm_ba* ba) {
  if (!ba) return;
  if (ba->ctx) { cudaSetDevice(ba->ctx->device); cudaStreamSynchronize(ba->ctx->stream); }
  sfm_ba_comm_destroy(ba);
  void* ptrs[] = {ba->uv, ba->cam_idx, ba->pt_idx, ba->pt_start, ba->cams, ba->cams_new, ba->pts, ba->pts_new,
                  ba->cam_pre, ba->S, ba->A64, ba->dc, ba->scal, ba->info, ba->pcg, ba->Tbuf, ba->qp};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete ba;
}

extern "C" int sfm_ba_set_totals(sfm_ba* ba, int64_t n_pt_total, int64_t n_obs_total) {
  SFM_REQUIRE(ba && n_pt_total >= ba->n_pt && n_obs_total >= ba->n_obs, "sfm_ba_set_totals: bad totals");
  ba->n_pt_total = n_pt_total;
  ba->n_obs_total = n_obs_total;
  return SFM_OK;
}

extern "C" int sfm_ba_set_params(sfm_ba* ba, const double* cams, const double* pts) {
  SFM_REQUIRE(ba, "sfm_ba_set_params: null problem");
  cudaStream_t st = ba->ctx->stream;
  SFM_CUDA(cudaSetDevice(ba->ctx->device));
  if (cams) SFM_CUDA(cudaMemcpyAsync(ba->cams, cams, sizeof(double) * 6 * (size_t)ba->n_cam, cudaMemcpyDefault, st));
  if (pts && ba->n_pt) SFM_CUDA(cudaMemcpyAsync(ba->pts, pts, sizeof(double) * 3 * (size_t)ba->n_pt, cudaMemcpyDefault, st));
  SFM_CUDA(cudaStreamSynchronize(st));
  return SFM_OK;
}

extern "C" int sfm_ba_get_params(sfm_ba* ba, double* cams, double* pts) {
  SFM_REQUIRE(ba, "sfm_ba_get_params: null problem");
  cudaStream_t st = ba->ctx->stream;
  SFM_CUDA(cudaSetDevice(ba->ctx->device));
  if (cams) SFM_CUDA(cudaMemcpyAsync(cams, ba->cams, sizeof(double) * 6 * (size_t)ba->n_cam, cudaMemcpyDefault, st));
  if (pts && ba->n_pt) SFM_CUDA(cudaMemcpyAsync(pts, ba->pts, sizeof(double) * 3 * (size_t)ba->n_pt, cudaMemcpyDefault, st));
  SFM_CUDA(cudaStreamSynchronize(st));
  return SFM_OK;
}

extern "C" int sfm_ba_eval(sfm_ba* ba, int mode, float* r, float* Jc, float* Jp, double* cost) {
  SFM_REQUIRE(ba, "sfm_ba_eval: null problem");
  SFM_REQUIRE(mode >= 0 && mode <= 2, "sfm_ba_eval: mode %d", mode);
  SFM_REQUIRE(mode == 0 || (!Jc && !Jp), "sfm_ba_eval: Jacobians exist for mode 0 only");
  sfm_ctx* ctx = ba->ctx;
  SFM_TRY(sfm_ws_begin(ctx));
  const size_t O = (size_t)ba->n_obs;
  bool host_out = false;
  DevOut<float> orr, ojc, ojp;
  DevOut<double> oc;
  SFM_TRY(dev_out(ctx, r, (mode == 2 ? 1 : 2) * O, &orr, &host_out));
  SFM_TRY(dev_out(ctx, Jc, 12 * O, &ojc, &host_out));
  SFM_TRY(dev_out(ctx, Jp, 6 * O, &ojp, &host_out));
  SFM_TRY(dev_out(ctx, cost, 1, &oc, &host_out));
  SFM_TRY(cam_prep(ba, ba->cams));
  if (oc.dev) SFM_CUDA(cudaMemsetAsync(oc.dev, 0, sizeof(double), ctx->stream));
  if (O > 0) {
    if (mode == 0) SFM_TRY(launch_eval<0>(ba, ba->pts, orr.dev, ojc.dev, ojp.dev, oc.dev));
    else if (mode == 1) SFM_TRY(launch_eval<1>(ba, ba->pts, orr.dev, nullptr, nullptr, oc.dev));
    else SFM_TRY(launch_eval<2>(ba, ba->pts, orr.dev, nullptr, nullptr, oc.dev));
  }
  SFM_TRY(dev_out_finish(ctx, &orr));
  SFM_TRY(dev_out_finish(ctx, &ojc));
  SFM_TRY(dev_out_finish(ctx, &ojp));
  SFM_TRY(dev_out_finish(ctx, &oc));
  if (host_out) SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFM_OK;
}

extern "C" int sfm_ba_build_system(sfm_ba* ba, double lambda) {
  SFM_REQUIRE(ba && lambda >= 0.0, "sfm_ba_build_system: bad argument");
  SFM_TRY(sfm_ws_begin(ba->ctx));
  return build_system(ba, lambda);
}

extern "C" int sfm_ba_read(sfm_ba* ba, int which, float* out, int64_t count) {
  SFM_REQUIRE(ba && out, "sfm_ba_read: null argument");
  const int n = 6 * ba->n_cam;
  const float* src = nullptr;
  int64_t have = 0;
  if (which == 0) { src = ba->S; have = (int64_t)n * n; }
  else if (which == 1) { src = ba->g; have = n; }
  else if (which == 2) { src = ba->hdiag; have = n; }
  SFM_REQUIRE(src, "sfm_ba_read: which=%d", which);
  SFM_REQUIRE(count <= have, "sfm_ba_read: count %lld > %lld", (long long)count, (long long)have);
  if (which == 0) {            // S lives as the lower triangle of 6x6 blocks: hand it out row-major (upper blocks zero)
    SFM_REQUIRE(count == have && !sfm_is_device_ptr(out), "sfm_ba_read: S is read whole, into host memory");
    const size_t s_floats = (size_t)ba->n_cam * (ba->n_cam + 1) / 2 * 36;
    std::vector<float> tiled(s_floats);
    SFM_CUDA(cudaMemcpyAsync(tiled.data(), src, sizeof(float) * s_floats, cudaMemcpyDeviceToHost, ba->ctx->stream));
    SFM_CUDA(cudaStreamSynchronize(ba->ctx->stream));
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < n; ++c) {
        const size_t ca = r / 6, cb = c / 6;
        out[(size_t)r * n + c] = cb <= ca ? tiled[(ca * (ca + 1) / 2 + cb) * 36 + 6 * (r % 6) + c % 6] : 0.f;
      }
    return SFM_OK;
  }
  SFM_CUDA(cudaMemcpyAsync(out, src, sizeof(float) * (size_t)count, cudaMemcpyDefault, ba->ctx->stream));
  SFM_CUDA(cudaStreamSynchronize(ba->ctx->stream));
  return SFM_OK;
}

extern "C" int sfm_ba_gn_step(sfm_ba* ba, double lambda, sfm_ba_stats* stats) {
  SFM_REQUIRE(ba && lambda >= 0.0, "sfm_ba_gn_step: bad argument");
  sfm_ctx* ctx = ba->ctx;
  SFM_TRY(sfm_ws_begin(ctx));
  const int n = 6 * ba->n_cam;
  SFM_TRY(build_system(ba, lambda));                 // scal[0] = cost at the linearisation point (all ranks)
  SFM_TRY(solve_reduced_system(ba));                 // dc
  SFM_TRY(update_points(ba, lambda));                // pts_new, scal[2] += |dp|^2 (cam_pre still at cams)
  SFM_LAUNCH(ctx, SFM_K_BA_UPDATE, (ba_update_cams_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(ba->cams, ba->dc, n, ba->cams_new, ba->scal + 3)));
  SFM_TRY(cam_prep(ba, ba->cams_new));
  if (ba->n_obs > 0) SFM_TRY(launch_eval<0>(ba, ba->pts_new, nullptr, nullptr, nullptr, ba->scal + 1));
  SFM_TRY(sfm_ba_allreduce_scalars(ba));             // scal[1] (new cost), scal[2] (|dp|^2) summed over ranks
  double* h;
  int* hinfo;
  SFM_TRY(hs_alloc_t(ctx, 8, &h));
  SFM_TRY(hs_alloc_t(ctx, 1, &hinfo));
  SFM_CUDA(cudaMemcpyAsync(h, ba->scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(cudaMemcpyAsync(hinfo, ba->info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  const double cost0 = h[0], cost1 = h[1];
  const bool ok = (*hinfo == 0) && isfinite(cost1) && cost1 < cost0;
  if (ok) {
    std::swap(ba->cams, ba->cams_new);
    std::swap(ba->pts, ba->pts_new);
  }
  if (stats) {
    stats->cost_before = cost0;
    stats->cost_after = cost1;
    stats->step_norm = sqrt(h[2] + h[3]);
    stats->grad_norm = 0.0;
    stats->accepted = ok ? 1 : 0;
    stats->solve_info = *hinfo;
    // the damping is not taken below 1e-6: S is accumulated in float32, and a relative damping under its rounding
    // (~1e-7) leaves the gauge directions numerically singular — any solver then returns a step dominated by them
    stats->lambda_next = ok ? fmax(lambda / 10.0, 1e-6) : fmin(fmax(lambda, 1e-6) * 10.0, 1e12);
  }
  return SFM_OK;
}
