// match_tc.cu — K1b descriptor tile preparation and K1 tcgen05 distance GEMM with a fused
// top-2 epilogue (sm_100a only).
//
// Replaces cv2.BFMatcher(NORM_L2).knnMatch(k=2) (reference sfm.py:259-260, isfm.py:71,
// test.py:42,225,352) for integer-valued descriptors in [0,255] (what cv2.SIFT emits): such
// values are exact in bf16 and every partial sum below stays under 2^23 in magnitude, so the
// fp32 tensor-core accumulator holds the EXACT squared distance and match indices are
// bit-identical to OpenCV's.
//
// Data layout in HBM (written once per view by K1b, "UMMA operand image"):
//   a view's descriptors are cut into 128-row tiles; a tile is stored as the exact shared-memory
//   image the tensor core wants — K-major, no swizzle, 8x(16 B) core matrices:
//     byte(r,k) = (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2,  LBO = 128, SBO = 20*128
//   with K = 160 bf16 columns: 128 descriptor values + 16 "A-role" + 16 "B-role" augmentation
//   columns.  Because the image already is the smem layout, a tile (40 KiB) or a two-tile stage
//   (80 KiB) moves HBM->smem with ONE cp.async.bulk (TMA bulk engine, mbarrier complete_tx),
//   no tensor map, no swizzle bookkeeping, fully contiguous reads.
//
// Augmentation (h = |d|^2, integer <= 2^21; c0,c1,c2 = the three bytes of h, each halved — all
// bf16-exact):   A-role cols 128..143 = [-c0,-c1,-c2, 1,1,1,1, 0...]
//                B-role cols 144..159 = [ 1, 1, 1,-c0,-c1,-c2,-2^22, 0...]
//   acc = q.t  - (|q|^2+|t|^2)/2 - 2^22 = -(2^22 + d^2/2)   in (-2^23, -2^22]
//   so the fp32 bit pattern of acc is 0xCA800000 | d^2 : the epilogue never converts or adds, it
//   forms a sortable 32-bit key (bits<<8 | column) with one integer op and keeps a running top-2
//   with three integer min/max ops per element.
//
// Kernel: persistent, warp-specialised, 1 CTA/SM, 320 threads:
//   warp 0  producer  — cp.async.bulk of TWO query tiles (once per work item, 80 KiB) and of one
//                       128-row train tile per stage into a 3-deep smem ring (full/empty mbarriers)
//   warp 1  MMA       — one lane issues 2 x 9 tcgen05.mma (M128 N128 K16, bf16 -> fp32 in TMEM), one
//                       set per resident query tile; tcgen05.commit releases the smem stage and
//                       publishes the accumulator pair
//   warps 2-9 epilogue — tcgen05.ld 32x32b.x32 of their TMEM lane quarter (one query row per thread,
//                       two warps per quarter, one per query tile), top-2 in registers; TMEM
//                       accumulators are double-buffered (2 x 256 columns) so the epilogue of stage s
//                       overlaps the MMAs of s+1.
// A work item is (query tile pair, train split); per-split results are merged by K1c (match.cu).
// Measured motivation for the tile pair: with one query tile per train tile the kernel saturated
// L2 -> SM bandwidth (~6.4 TB/s of train-tile re-reads at 32k x 32k) at 31 % tensor-pipe activity.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "match_common.cuh"

namespace tc {
constexpr int TILE_ROWS = 128;
constexpr int KMAIN = 128;
constexpr int KAUG = 160;
constexpr int CORE_COLS = KAUG / 8;             // 20 core-matrix columns
constexpr int LBO = 128;                        // bytes between K-adjacent core matrices
constexpr int SBO = CORE_COLS * 128;            // bytes between 8-row groups (2560)
constexpr int TILE_BYTES = (TILE_ROWS / 8) * SBO;   // 40960
constexpr int QT_PER_ITEM = 2;                  // query tiles resident per work item: every train tile read
                                                // from L2 feeds 2 x 128 query rows (halves L2 -> SM traffic,
                                                // which is what bounds this kernel, not the tensor pipe)
constexpr int STAGE_COLS = TILE_ROWS;           // one 128-row train tile per stage, MMA N = 128
constexpr int STAGE_BYTES = TILE_BYTES;         // 40960
constexpr int NSTAGE = 3;                       // train-tile ring depth
constexpr int ACC_COLS = QT_PER_ITEM * STAGE_COLS;   // 256 TMEM columns per accumulator buffer (x2 buffers)
constexpr int NTHREADS = 320;                   // producer warp, MMA warp, 8 epilogue warps
constexpr int SMEM_A = 0;
constexpr int SMEM_B = QT_PER_ITEM * TILE_BYTES;            // 81920
constexpr int SMEM_BAR = SMEM_B + NSTAGE * STAGE_BYTES;     // 204800
constexpr int SMEM_BYTES = SMEM_BAR + 128;
constexpr unsigned MAX_SQNORM = 1u << 21;
// barrier indices
enum { A_FULL = 0, A_EMPTY = 1, B_FULL = 2, B_EMPTY = 5, ACC_FULL = 8, ACC_EMPTY = 10, NBAR = 12 };
}  // namespace tc

// ============================================================================ K1b descriptor prep
// One pass over a view's descriptors (float32 or uint8 source): float32 resident copy, bf16 UMMA
// operand image with the |d|^2 augmentation columns, exact |d|^2, and the domain flags.
// One warp per 8-row group: lane l owns row l/4 and the 2-element slice (l%4) of every core
// matrix, so each of the 20 core matrices of the group is written as one coalesced 128-byte
// store and read as 4 x (8 or 2)-byte pieces of one 32-byte (float32) sector per row.
template <typename T>
__global__ void __launch_bounds__(256) desc_prepare_kernel(const T* __restrict__ src, int n,
                                                            float* __restrict__ f32,
                                                            unsigned char* __restrict__ tiles, int n_groups,
                                                            float* __restrict__ sqnorm,
                                                            unsigned int* __restrict__ flag) {
  int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (group >= n_groups) return;
  int row = group * 8 + (lane >> 2);
  int sub = lane & 3;
  bool valid = row < n;
  const T* rp = src + (size_t)row * tc::KMAIN;
  unsigned char* gbase = tiles + (size_t)group * tc::SBO;     // tiles are contiguous: group g at g*SBO
  float h = 0.f;
  bool bad = false;
#pragma unroll
  for (int kc = 0; kc < tc::KMAIN / 8; ++kc) {
    float2 v = make_float2(0.f, 0.f);
    if (valid) {
      v.x = (float)rp[kc * 8 + sub * 2];
      v.y = (float)rp[kc * 8 + sub * 2 + 1];
      *reinterpret_cast<float2*>(f32 + (size_t)row * tc::KMAIN + kc * 8 + sub * 2) = v;
      bad |= !(v.x >= 0.f && v.x <= 255.f && v.x == rintf(v.x)) || !(v.y >= 0.f && v.y <= 255.f && v.y == rintf(v.y));
    }
    h = fmaf(v.x, v.x, h);
    h = fmaf(v.y, v.y, h);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(gbase + kc * tc::LBO + lane * 4) = b;
  }
  h += __shfl_xor_sync(0xffffffffu, h, 1);
  h += __shfl_xor_sync(0xffffffffu, h, 2);   // exact when the values are integers: sums < 2^24
  unsigned int flags = bad ? 1u : 0u;
  if (valid && !(h <= (float)tc::MAX_SQNORM)) flags |= 2u;
  if (__any_sync(0xffffffffu, flags != 0u)) {
    unsigned int f = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0) atomicOr(flag, f);
  }
  unsigned int hi = (h >= 0.f && h <= (float)tc::MAX_SQNORM) ? (unsigned int)h : 0u;
  if (valid && sub == 0) sqnorm[row] = h;
  float c0 = -0.5f * (float)(hi & 0xFFu), c1 = -0.5f * (float)(hi & 0xFF00u), c2 = -0.5f * (float)(hi & 0xFF0000u);
  // augmentation columns: element e (0..7) of core column 16 (A-role) and 18 (B-role); 17,19 zero
  float a0, a1, b0, b1;
  if (sub == 0) { a0 = c0; a1 = c1; b0 = 1.f; b1 = 1.f; }
  else if (sub == 1) { a0 = c2; a1 = 1.f; b0 = 1.f; b1 = c0; }
  else if (sub == 2) { a0 = 1.f; a1 = 1.f; b0 = c1; b1 = c2; }
  else { a0 = 1.f; a1 = 0.f; b0 = -4194304.f; b1 = 0.f; }
  // padding rows: as a query they produce garbage nobody reads; as a train column the accumulator becomes
  // exactly -255 * 2^15 (bits 0xCAFF0000, a "distance" above any real one), so the epilogue needs no masking
  if (!valid) { a0 = a1 = b1 = 0.f; b0 = (sub == 3) ? -8355840.f : 0.f; }
  *reinterpret_cast<__nv_bfloat162*>(gbase + 16 * tc::LBO + lane * 4) = __floats2bfloat162_rn(a0, a1);
  *reinterpret_cast<__nv_bfloat162*>(gbase + 17 * tc::LBO + lane * 4) = __floats2bfloat162_rn(0.f, 0.f);
  *reinterpret_cast<__nv_bfloat162*>(gbase + 18 * tc::LBO + lane * 4) = __floats2bfloat162_rn(b0, b1);
  *reinterpret_cast<__nv_bfloat162*>(gbase + 19 * tc::LBO + lane * 4) = __floats2bfloat162_rn(0.f, 0.f);
}

int sfm_desc_prepare_launch(sfm_ctx* ctx, sfm_desc* d, const void* src, int dtype) {
  // even number of tiles so that a train stage is always two full tiles (padding rows are zero);
  // storage capacity is a multiple of 256 rows, so the padded tile always exists
  int n_tiles_alloc = (div_up(d->n, tc::TILE_ROWS) + 1) & ~1;
  int n_groups = n_tiles_alloc * (tc::TILE_ROWS / 8);
  DescBuf* b = d->buf;
  if (dtype == 0)
    SFM_LAUNCH(ctx, SFM_K_DESC_PREP, (desc_prepare_kernel<float><<<div_up(n_groups * 32, 256), 256, 0, ctx->stream>>>(
                                         (const float*)src, d->n, b->f32, b->tiles, n_groups, b->sqnorm, b->flag)));
  else
    SFM_LAUNCH(ctx, SFM_K_DESC_PREP, (desc_prepare_kernel<unsigned char><<<div_up(n_groups * 32, 256), 256, 0, ctx->stream>>>(
                                         (const unsigned char*)src, d->n, b->f32, b->tiles, n_groups, b->sqnorm, b->flag)));
  return SFM_OK;
}

// ============================================================================ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (an error the host sees), never
// as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 22)) { asm volatile("trap;"); }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Wait for outstanding tcgen05.ld and tie the destination registers to the wait, so that the compiler
// cannot consume (or move) them before the asynchronous load has landed.
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (version 1 = Blackwell).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(tc::LBO >> 4) << 16) |
         ((uint64_t)(tc::SBO >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=256.
__device__ __forceinline__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcParams {
  const unsigned char* q_tiles;   // query view image
  const unsigned char* t_tiles;   // train view image
  int n_qtiles;                   // 128-row query tiles
  int n_qpairs;                   // work-item rows: ceil(n_qtiles / 2)
  int n_stages;                   // 128-row train tiles
  int nt;                         // valid train rows
  int nsplit;
  int stages_per_split;
  int n_items;
  mkey_t* cand;                   // [n_qtiles*128][nsplit][3]: best, runner-up, "check this column" (K1c)
  float* dump;                    // debug: raw accumulators [n_qtiles*128][n_stages*128] or NULL
  unsigned int key_mul;           // = 256, passed at run time so the key build stays an IMAD (FMA pipe)
  unsigned int debug;             // diagnostics (env SFM_MATCH_DEBUG): bit0 skip epilogue math, bit1 skip MMAs
};

// ---------------------------------------------------------------------------- epilogue arithmetic
// The accumulator bits are 0xCA800000 | d^2.  key = bits * 256 + tag (one IMAD, FMA pipe) is
// 0x80000000 | d^2 << 8 | tag with tag = index of the 32-column chunk inside the split (< 256), so
// unsigned order on keys is (d^2, chunk) order.  Per query row (= per thread) the epilogue keeps
//   acc[j], j = column mod 32 : the minimum key of every residue class ("vertical" minima), and
//   (h1, h2)                  : the two smallest per-chunk minima ("horizontal" minima).
// That is ONE integer min/max-pipe operation per element (two 3-input minima per two elements) instead
// of the three a running top-2 needs, and it still determines the exact top-2 of the row: the best
// element e1 is the smallest (acc[j], j); the runner-up e2 is either in another class than e1 — then it
// is the minimum of its class and shows up as the second smallest (acc[j], j) — or in e1's class j1 —
// then it is in another chunk, is the minimum of that chunk, and h2 is its key (column = chunk*32+j1).
// Columns inside one chunk have distinct classes, so no pair of elements can hide in both views.
__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return min(min(a, b), c); }

__device__ __forceinline__ uint32_t chunk_min(const uint32_t (&k)[32]) {
  uint32_t m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) m[i] = umin3(k[3 * i], k[3 * i + 1], k[3 * i + 2]);
  m[10] = min(k[30], k[31]);
  const uint32_t a = umin3(m[0], m[1], m[2]), b = umin3(m[3], m[4], m[5]), c = umin3(m[6], m[7], m[8]);
  return umin3(umin3(a, b, c), m[9], m[10]);
}

// Two 32-column chunks (raw accumulator bits in a, b) folded into the row state.
__device__ __forceinline__ void fold_pair(uint32_t (&a)[32], uint32_t (&b)[32], uint32_t mul256, uint32_t tag_a,
                                          uint32_t (&acc)[32], uint32_t& h1, uint32_t& h2) {
  const uint32_t tag_b = tag_a + 1u;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    a[j] = a[j] * mul256 + tag_a;
    b[j] = b[j] * mul256 + tag_b;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = umin3(acc[j], a[j], b[j]);
  const uint32_t ha = chunk_min(a), hb = chunk_min(b);     // ha != hb (different tags)
  const uint32_t lo = min(ha, hb), hi = max(ha, hb);
  h2 = umin3(max(h1, lo), h2, hi);
  h1 = min(h1, lo);
}

// ============================================================================ K1 kernel
template <bool DUMP>
__global__ void __launch_bounds__(tc::NTHREADS, 1) match_tc_kernel(TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + tc::SMEM_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + tc::SMEM_BAR + tc::NBAR * 8);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (threadIdx.x == 0) {
    mbar_init(bar(tc::A_FULL), 1);
    mbar_init(bar(tc::A_EMPTY), 1);
    for (int s = 0; s < tc::NSTAGE; ++s) {
      mbar_init(bar(tc::B_FULL + s), 1);
      mbar_init(bar(tc::B_EMPTY + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(tc::ACC_FULL + b), 1);
      mbar_init(bar(tc::ACC_EMPTY + b), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: all 512 columns = two buffers of (2 query tiles x 128 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      uint32_t a_phase = 0, b_phase[tc::NSTAGE] = {0, 0, 0};
      int slot = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int qp = item / p.nsplit, split = item % p.nsplit;
        const int s_begin = split * p.stages_per_split;
        const int s_end = min(p.n_stages, s_begin + p.stages_per_split);
        const int nqt = min(tc::QT_PER_ITEM, p.n_qtiles - qp * tc::QT_PER_ITEM);
        mbar_wait(bar(tc::A_EMPTY), a_phase ^ 1);
        mbar_expect_tx(bar(tc::A_FULL), (uint32_t)nqt * tc::TILE_BYTES);
        bulk_g2s(sbase + tc::SMEM_A, p.q_tiles + (size_t)qp * tc::QT_PER_ITEM * tc::TILE_BYTES,
                 (uint32_t)nqt * tc::TILE_BYTES, bar(tc::A_FULL));
        a_phase ^= 1;
        for (int s = s_begin; s < s_end; ++s) {
          mbar_wait(bar(tc::B_EMPTY + slot), b_phase[slot] ^ 1);
          mbar_expect_tx(bar(tc::B_FULL + slot), tc::STAGE_BYTES);
          bulk_g2s(sbase + tc::SMEM_B + slot * tc::STAGE_BYTES, p.t_tiles + (size_t)s * tc::STAGE_BYTES,
                   tc::STAGE_BYTES, bar(tc::B_FULL + slot));
          b_phase[slot] ^= 1;
          slot = (slot + 1 == tc::NSTAGE) ? 0 : slot + 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, tc::STAGE_COLS);
      uint32_t a_phase = 0, b_phase[tc::NSTAGE] = {0, 0, 0}, acc_phase[2] = {0, 0};
      int slot = 0, buf = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int qp = item / p.nsplit, split = item % p.nsplit;
        const int s_begin = split * p.stages_per_split;
        const int s_end = min(p.n_stages, s_begin + p.stages_per_split);
        const int nqt = min(tc::QT_PER_ITEM, p.n_qtiles - qp * tc::QT_PER_ITEM);
        mbar_wait(bar(tc::A_FULL), a_phase);
        a_phase ^= 1;
        for (int s = s_begin; s < s_end; ++s) {
          mbar_wait(bar(tc::ACC_EMPTY + buf), acc_phase[buf] ^ 1);
          mbar_wait(bar(tc::B_FULL + slot), b_phase[slot]);
          b_phase[slot] ^= 1;
          tc_fence_after();
          const uint32_t b_addr = sbase + tc::SMEM_B + slot * tc::STAGE_BYTES;
          for (int t = 0; t < ((p.debug & 2u) ? 0 : nqt); ++t) {
            const uint32_t a_addr = sbase + tc::SMEM_A + t * tc::TILE_BYTES;
            const uint32_t d_addr = tmem_base + (uint32_t)(buf * tc::ACC_COLS + t * tc::STAGE_COLS);
#pragma unroll
            for (int k = 0; k < tc::KMAIN / 16; ++k)
              tc_mma_bf16(d_addr, make_smem_desc(a_addr + k * 2 * tc::LBO), make_smem_desc(b_addr + k * 2 * tc::LBO),
                          idesc, k > 0 ? 1u : 0u);
            tc_mma_bf16(d_addr, make_smem_desc(a_addr + 16 * tc::LBO), make_smem_desc(b_addr + 18 * tc::LBO), idesc, 1u);
          }
          tc_commit(bar(tc::B_EMPTY + slot));     // smem stage reusable once these MMAs retire
          tc_commit(bar(tc::ACC_FULL + buf));     // accumulator pair complete
          acc_phase[buf] ^= 1;
          slot = (slot + 1 == tc::NSTAGE) ? 0 : slot + 1;
          buf ^= 1;
        }
        tc_commit(bar(tc::A_EMPTY));              // query tiles reusable
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM lane quarter (= per SM sub-partition): warps 2-5 reduce query tile 0 of the
    // pair, warps 6-9 query tile 1; one query row per thread.
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;          // query row within a tile
    const int tq = (warp - 2) >> 2;               // which query tile of the pair this warp owns
    uint32_t acc_phase[2] = {0, 0};
    int buf = 0;
    const uint32_t mul256 = p.key_mul;  // 256, opaque to the compiler: the key build stays an IMAD (FMA pipe),
                                        // leaving the integer min/max pipe to the minima
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int qp = item / p.nsplit, split = item % p.nsplit;
      const int s_begin = split * p.stages_per_split;
      const int s_end = min(p.n_stages, s_begin + p.stages_per_split);
      const int nqt = min(tc::QT_PER_ITEM, p.n_qtiles - qp * tc::QT_PER_ITEM);
      const bool active = tq < nqt;
      uint32_t acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0xFFFFFFFFu;
      uint32_t h1 = 0xFFFFFFFFu, h2 = 0xFFFFFFFFu;
      for (int s = s_begin; s < s_end; ++s) {
        mbar_wait(bar(tc::ACC_FULL + buf), acc_phase[buf]);
        acc_phase[buf] ^= 1;
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * tc::ACC_COLS + tq * tc::STAGE_COLS);
        const uint32_t tag0 = (uint32_t)(s - s_begin) * (tc::STAGE_COLS / 32);
        if (active && !(p.debug & 4u)) {
          uint32_t ra[32], rb[32], rc[32];
          tmem_ld32(t_addr, ra);
          tmem_ld32(t_addr + 32, rb);
          tmem_ld_wait_regs(ra);
          tmem_ld_wait_regs(rb);
          tmem_ld32(t_addr + 64, rc);             // third chunk in flight while the first two are folded
          if (DUMP) {
            float* drow = p.dump + ((size_t)((qp * tc::QT_PER_ITEM + tq) * 128 + row) * p.n_stages + s) * tc::STAGE_COLS;
#pragma unroll
            for (int j = 0; j < 32; ++j) { drow[j] = __uint_as_float(ra[j]); drow[32 + j] = __uint_as_float(rb[j]); }
          }
          if (p.debug & 1u) h1 = min(h1, ra[0] ^ rb[31]);
          else fold_pair(ra, rb, mul256, tag0, acc, h1, h2);
          tmem_ld32(t_addr + 96, ra);
          tmem_ld_wait_regs(rc);
          tmem_ld_wait_regs(ra);
          // every column of this accumulator is in registers: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(tc::ACC_EMPTY + buf));
          if (DUMP) {
            float* drow = p.dump + ((size_t)((qp * tc::QT_PER_ITEM + tq) * 128 + row) * p.n_stages + s) * tc::STAGE_COLS + 64;
#pragma unroll
            for (int j = 0; j < 32; ++j) { drow[j] = __uint_as_float(rc[j]); drow[32 + j] = __uint_as_float(ra[j]); }
          }
          if (p.debug & 1u) h1 = min(h1, rc[0] ^ ra[31]);
          else fold_pair(rc, ra, mul256, tag0 + 2u, acc, h1, h2);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(tc::ACC_EMPTY + buf));
        }
        buf ^= 1;
      }
      if (active) {
        // row state -> the split's candidates.  (key, class) order == (d^2, column) order.
        uint32_t b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
        int j1 = 0, j2 = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const uint32_t k = acc[j];
          if (k < b1) { b2 = b1; j2 = j1; b1 = k; j1 = j; }
          else if (k < b2) { b2 = k; j2 = j; }
        }
        const int col_base = s_begin * tc::STAGE_COLS;
        auto key_d2 = [](uint32_t k) { return (k >> 8) & 0x7FFFFFu; };
        auto key_col = [&](uint32_t k, int j) { return col_base + (int)(k & 0xFFu) * 32 + j; };
        mkey_t o1 = MKEY_INF, o2 = MKEY_INF, o3 = MKEY_INF;
        if (key_d2(b1) <= 2u * tc::MAX_SQNORM) o1 = make_key((float)key_d2(b1), key_col(b1, j1));
        if (h2 < b2) {                       // runner-up hidden behind e1 in class j1: it is chunk h2's minimum
          if (key_d2(h2) <= 2u * tc::MAX_SQNORM) o2 = make_key((float)key_d2(h2), key_col(h2, j1));
        } else {
          if (key_d2(b2) <= 2u * tc::MAX_SQNORM) {
            o2 = make_key((float)key_d2(b2), key_col(b2, j2));
            // same d^2 in the same chunk: the chunk may also hold an equal element in class j1 (hidden behind
            // e1); if j1 < j2 it would precede o2.  K1c checks that column when it matters.
            if (h2 == b2 && j1 < j2) o3 = make_key((float)key_d2(b2), key_col(b2, j1));
          }
        }
        mkey_t* out = p.cand + ((size_t)((qp * tc::QT_PER_ITEM + tq) * 128 + row) * p.nsplit + split) * 3;
        out[0] = o1;
        out[1] = o2;
        out[2] = o3;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int sfm_match_tc_splits(sfm_ctx* ctx, int nq, int nt) {
  int qpairs = div_up(div_up(nq, tc::TILE_ROWS), tc::QT_PER_ITEM);
  int stages = div_up(nt, tc::TILE_ROWS);
  int s = ctx->sm_count / (qpairs > 0 ? qpairs : 1);
  int s_min = div_up(stages, 256 / (tc::STAGE_COLS / 32));   // chunk tags are 8 bits: <= 256 chunks per split
  if (s < s_min) s = s_min;
  if (s < 1) s = 1;
  if (s > stages) s = stages;
  return s < 1 ? 1 : s;
}

static int launch_tc(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, mkey_t* cand, int nsplit, float* dump) {
  SFM_REQUIRE(q->tiles && t->tiles, "tensor-core matcher: descriptors have no tile image");
  static bool attr_set = false;
  if (!attr_set) {
    SFM_CUDA(cudaFuncSetAttribute(match_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    SFM_CUDA(cudaFuncSetAttribute(match_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    attr_set = true;
  }
  TcParams p;
  p.q_tiles = (const unsigned char*)q->tiles;
  p.t_tiles = (const unsigned char*)t->tiles;
  p.n_qtiles = q->n_tiles;
  p.n_qpairs = div_up(q->n_tiles, tc::QT_PER_ITEM);
  p.n_stages = t->n_tiles;
  p.nt = t->n;
  p.nsplit = nsplit;
  p.stages_per_split = div_up(p.n_stages, nsplit);
  p.n_items = p.n_qpairs * nsplit;
  p.cand = cand;
  p.dump = dump;
  p.key_mul = 256u;
  { const char* e = getenv("SFM_MATCH_DEBUG"); p.debug = e ? (unsigned)atoi(e) : 0u; }
  int grid = p.n_items < ctx->sm_count ? p.n_items : ctx->sm_count;
  if (dump) SFM_LAUNCH(ctx, SFM_K_MATCH_TC, (match_tc_kernel<true><<<grid, tc::NTHREADS, tc::SMEM_BYTES, ctx->stream>>>(p)));
  else SFM_LAUNCH(ctx, SFM_K_MATCH_TC, (match_tc_kernel<false><<<grid, tc::NTHREADS, tc::SMEM_BYTES, ctx->stream>>>(p)));
  return SFM_OK;
}

int sfm_match_tc_launch(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, mkey_t* cand, int nsplit) {
  return launch_tc(ctx, q, t, cand, nsplit, nullptr);
}

// Debug/self-test entry (not part of the reference-facing surface): raw accumulators of every
// (query row, train column), i.e. -(2^22 + d^2/2), as float32 [n_qtiles*128][n_ttiles*128].
extern "C" int sfm_debug_match_tc_dump(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, float* dump_host,
                                       int64_t capacity) {
  SFM_REQUIRE(ctx && q && t && dump_host, "sfm_debug_match_tc_dump: null argument");
  SFM_TRY(sfm_desc_resolve(const_cast<sfm_desc*>(q)));
  SFM_TRY(sfm_desc_resolve(const_cast<sfm_desc*>(t)));
  SFM_TRY(sfm_ws_begin(ctx));
  int nsplit = 1;
  size_t count = (size_t)q->n_tiles * 128 * t->n_tiles * tc::STAGE_COLS;
  SFM_REQUIRE((int64_t)count <= capacity, "dump buffer too small: need %zu floats", count);
  mkey_t* cand;
  float* dump;
  SFM_TRY(ws_alloc_t(ctx, (size_t)q->n_tiles * 128 * nsplit * 3, &cand));
  SFM_TRY(ws_alloc_t(ctx, count, &dump));
  SFM_TRY(launch_tc(ctx, q, t, cand, nsplit, dump));
  SFM_CUDA(cudaMemcpyAsync(dump_host, dump, count * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFM_OK;
}
