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
//   columns.  Because the image already is the smem layout, a tile (40 KiB) moves HBM->smem with
//   ONE cp.async.bulk (TMA bulk engine, mbarrier complete_tx), no tensor map, no swizzle
//   bookkeeping, fully contiguous reads.
//
// Augmentation (h = |d|^2, integer <= 2^21; c0,c1,c2 = the three bytes of h, each halved — all
// bf16-exact):   A-role cols 128..143 = [-c0,-c1,-c2, 1,1,1,1, 0...]
//                B-role cols 144..159 = [ 1, 1, 1,-c0,-c1,-c2,-2^22, 0...]
//   acc = q.t  - (|q|^2+|t|^2)/2 - 2^22 = -(2^22 + d^2/2)   in (-2^23, -2^22]
//   so the fp32 bit pattern of acc is 0xCA800000 | d^2 : the epilogue never converts or adds.
//   Padding rows of a train tile carry B-role [0,0,0,0,0,0,-255*2^15]: their accumulator is
//   0xCAFF0000, above every real distance, so the epilogue needs no column masking.
//
// Kernel: persistent, warp-specialised, ONE CTA PAIR per TPC (cluster of 2, tcgen05 cta_group::2),
// 352 threads per CTA.  The pair multiplies a 256-row query block (128 rows from each CTA's shared
// memory) with a 256-column train stage (128 columns from each CTA's shared memory): one
// tcgen05.mma M256 N256 K16 per 128 cycles reads only 8 KiB of shared memory per CTA — half of what
// two independent M128 N128 instructions need, which is what bounded the single-CTA version
// (measured: 128 B/clk of operand reads + the TMA writes saturate shared memory at ~57 % tensor
// activity) — and every train tile fetched from L2 feeds 256 (QT=1) or 512 (QT=2) query rows.
//   warp 0   B producer — cp.async.bulk of this CTA's 128-column half of every train stage into a
//                         3-deep ring
//   warp 2   A producer — this CTA's query tiles (2-slot ring: the next item's tile loads while the
//                         last jobs of the current item run)
//   warp 1   leader CTA: one lane issues 9 x tcgen05.mma (8 K-steps + the augmentation step) per job
//                         and multicast tcgen05.commit to both CTAs' barriers;
//            peer CTA:   relays "my TMA data landed" to the leader's barriers (remote mbarrier arrive)
//   warps 3-10 epilogue — tcgen05.ld 32x32b.x32 of their TMEM lane quarter (one query row per thread,
//                         two warps per quarter, each taking 128 of the 256 columns)
// A job = (query tile t of the item, train stage s); its accumulator (128 lanes x 256 columns per
// CTA) lives in TMEM slot job&1, so the epilogue of job j overlaps the MMAs of job j+1.
// A work item is (query group, train split), numbered split-major so that the pairs running at the
// same time stream the SAME train tiles (L2 serves one line to many SMs much faster than many lines);
// per-(split, column half) candidates are merged by K1c.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <vector>

#include "match_common.cuh"

namespace tc {
constexpr int TILE_ROWS = 128;
constexpr int KMAIN = 128;
constexpr int KAUG = 160;
constexpr int CORE_COLS = KAUG / 8;             // 20 core-matrix columns
constexpr int LBO = 128;                        // bytes between K-adjacent core matrices
constexpr int SBO = CORE_COLS * 128;            // bytes between 8-row groups (2560)
constexpr int TILE_BYTES = (TILE_ROWS / 8) * SBO;   // 40960
constexpr int STAGE_COLS = 256;                 // train columns per stage: 128 from each CTA of the pair
constexpr int NA = 2;                           // query-tile ring depth (per CTA)
constexpr int NB = 3;                           // train-tile ring depth (per CTA)
constexpr int NTHREADS = 352;                   // B producer, MMA/relay, A producer, 8 epilogue warps
constexpr int SMEM_A = 0;
constexpr int SMEM_B = NA * TILE_BYTES;                 // 81920
constexpr int SMEM_BAR = SMEM_B + NB * TILE_BYTES;      // 204800
constexpr unsigned MAX_SQNORM = 1u << 21;
constexpr int CHUNK = 16;                       // epilogue chunk width = number of residue classes
constexpr int MAX_STAGES_PER_SPLIT = 256 / (STAGE_COLS / CHUNK);   // chunk tags are 8 bits -> 16 stages
// barrier indices (same layout in both CTAs)
enum { A_FULL = 0, A_EMPTY = 2, A_PEER = 4, B_FULL = 6, B_EMPTY = 9, B_PEER = 12, ACC_FULL = 15, ACC_EMPTY = 17, NBAR = 19 };
constexpr int SMEM_BYTES = SMEM_BAR + 256;
}  // namespace tc

// ============================================================================ K1b descriptor prep
// One pass over a view's descriptors (float32 or uint8 source): float32 resident copy, bf16 UMMA
// operand image with the |d|^2 augmentation columns, exact |d|^2, and the domain flags.
// One warp per 8-row group: lane l owns row l/4 and the 2-element slice (l%4) of every core
// matrix, so each of the 20 core matrices of the group is written as one coalesced 128-byte
// store and read as 4 x (8 or 2)-byte pieces of one 32-byte (float32) sector per row.
template <typename T>
__device__ __forceinline__ void desc_prepare_group(int group, const T* __restrict__ src, int n,
                                                   float* __restrict__ f32,
                                                   unsigned char* __restrict__ tiles, int n_groups,
                                                   float* __restrict__ sqnorm,
                                                   unsigned int* __restrict__ flag) {
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

template <typename T>
__global__ void __launch_bounds__(256) desc_prepare_kernel(const T* __restrict__ src, int n,
                                                            float* __restrict__ f32,
                                                            unsigned char* __restrict__ tiles, int n_groups,
                                                            float* __restrict__ sqnorm,
                                                            unsigned int* __restrict__ flag) {
  desc_prepare_group<T>((blockIdx.x * blockDim.x + threadIdx.x) >> 5, src, n, f32, tiles, n_groups, sqnorm, flag);
}

// Many views in one launch (blockIdx.y = view): a 5000-descriptor view is 2.5 MB, a launch per view is latency-,
// not bandwidth-bound (13 us for what HBM moves in under 1 us).
struct PrepItem {
  const void* src;
  float* f32;
  unsigned char* tiles;
  float* sqnorm;
  unsigned int* flag;
  int n, n_groups;
};
template <typename T>
__global__ void __launch_bounds__(256) desc_prepare_batched_kernel(const PrepItem* __restrict__ items) {
  const PrepItem it = items[blockIdx.y];
  desc_prepare_group<T>((blockIdx.x * blockDim.x + threadIdx.x) >> 5, (const T*)it.src, it.n, it.f32, it.tiles, it.n_groups,
                        it.sqnorm, it.flag);
}

int sfm_desc_prepare_launch_batched(sfm_ctx* ctx, int count, sfm_desc* const* d, const void* const* src, int dtype,
                                    unsigned int* flags_dev) {
  std::vector<PrepItem> host((size_t)count);
  int max_groups = 0;
  for (int k = 0; k < count; ++k) {
    const int n_tiles_alloc = (div_up(d[k]->n, tc::TILE_ROWS) + 1) & ~1;
    PrepItem& it = host[k];
    it.src = src[k]; it.f32 = d[k]->buf->f32; it.tiles = d[k]->buf->tiles; it.sqnorm = d[k]->buf->sqnorm;
    it.flag = flags_dev + k; it.n = d[k]->n; it.n_groups = n_tiles_alloc * (tc::TILE_ROWS / 8);
    if (it.n_groups > max_groups) max_groups = it.n_groups;
  }
  PrepItem* dev = nullptr;
  SFM_TRY(ws_alloc_t(ctx, (size_t)count, &dev));
  SFM_CUDA(cudaMemcpyAsync(dev, host.data(), sizeof(PrepItem) * count, cudaMemcpyHostToDevice, ctx->stream));
  dim3 grid(div_up(max_groups * 32, 256), count);
  if (dtype == 0)
    SFM_LAUNCH(ctx, SFM_K_DESC_PREP, (desc_prepare_batched_kernel<float><<<grid, 256, 0, ctx->stream>>>(dev)));
  else
    SFM_LAUNCH(ctx, SFM_K_DESC_PREP, (desc_prepare_batched_kernel<unsigned char><<<grid, 256, 0, ctx->stream>>>(dev)));
  return SFM_OK;
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

// One (query view, train view) problem of a launch.  A launch covers one or many of them ("batched":
// the items of all pairs form one list, so the fixed cost of a launch — cluster start, TMEM allocation,
// pipeline ramp and tail, ~7 us — is paid once for e.g. the 199 consecutive pairs of a 200-view scene).
struct TcPair {
  const unsigned char* q_tiles;   // query view image
  const unsigned char* t_tiles;   // train view image
  mkey_t* cand;                   // [n_qtiles*128][2*nsplit][3]: best, runner-up, "check this column" (K1c)
  int n_qtiles;                   // 128-row query tiles holding real rows
  int n_qtiles_alloc;             // tiles present in the image (even)
  int n_stages;                   // 256-column train stages
  int n_groups;                   // query groups of 2*QT tiles
  int nsplit;                     // train splits; items of the pair = n_groups * nsplit, split-major
  int item_begin;                 // first item of this pair in the launch's item list
  int n_items;
  int pad_;
};

struct TcParams {
  TcPair one;                     // the pair of a single-pair launch (pairs == nullptr)
  const TcPair* pairs;            // device array of a batched launch, ordered by item_begin
  int n_items;                    // all items of the launch
  float* dump;                    // debug: raw accumulators [n_qtiles*128][dump_cols] or NULL
  int dump_cols;
  long long* timeline;            // debug (MODE 2): per job of pair 0, 8 clock64 stamps taken on the leader SM
  unsigned int key_mul;           // = 256, passed at run time so the key build stays an IMAD (FMA pipe)
  unsigned int debug;             // diagnostics (env SFM_MATCH_DEBUG): bit0 skip epilogue math, bit1 skip MMAs, bit2 skip TMEM loads
};

// Walks the pair list as a role's item index grows (items are visited in increasing order).
struct PairCursor {
  const TcParams& p;
  TcPair pd;
  int pi;
  __device__ __forceinline__ explicit PairCursor(const TcParams& prm) : p(prm), pd(prm.pairs ? prm.pairs[0] : prm.one), pi(0) {}
  __device__ __forceinline__ int seek(int item) {          // -> item index inside its pair
    while (item >= pd.item_begin + pd.n_items) pd = p.pairs[++pi];
    return item - pd.item_begin;
  }
  __device__ __forceinline__ int s_begin_of(int split) const { return (int)(((long long)split * pd.n_stages) / pd.nsplit); }
};

// ---------------------------------------------------------------------------- epilogue arithmetic
// The accumulator bits are 0xCA800000 | d^2.  key = bits * 256 + tag (one IMAD, FMA pipe) is
// 0x80000000 | d^2 << 8 | tag with tag = index of the 16-column chunk inside the split (< 256), so
// unsigned order on keys is (d^2, chunk) order.  Per query row (= per thread) the epilogue keeps
//   acc[j], j = column mod 16 : the minimum key of every residue class ("vertical" minima), and
//   (h1, h2)                  : the two smallest per-chunk minima ("horizontal" minima).
// That is ~ONE integer min/max-pipe operation per element (3-input minima) instead of the three a
// running top-2 needs, and it still determines the exact top-2 of the row: the best element e1 is
// the smallest (acc[j], j); the runner-up e2 is either in another class than e1 — then it is the
// minimum of its class and shows up as the second smallest (acc[j], j) — or in e1's class j1 — then
// it is in another chunk, is the minimum of that chunk, and h2 is its key (column = chunk*16 + j1).
// Columns inside one chunk have distinct classes, so no pair of elements can hide in both views.
__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return min(min(a, b), c); }

__device__ __forceinline__ uint32_t chunk_min16(const uint32_t* k) {
  const uint32_t a = umin3(k[0], k[1], k[2]), b = umin3(k[3], k[4], k[5]), c = umin3(k[6], k[7], k[8]);
  const uint32_t d = umin3(k[9], k[10], k[11]), e = umin3(k[12], k[13], k[14]);
  return umin3(umin3(a, b, c), umin3(d, e, k[15]), 0xFFFFFFFFu);
}

struct RowState {
  uint32_t acc[tc::CHUNK];
  uint32_t h1, h2;
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < tc::CHUNK; ++j) acc[j] = 0xFFFFFFFFu;
    h1 = h2 = 0xFFFFFFFFu;
  }
};

// 32 columns (raw accumulator bits, two 16-column chunks tagged tag and tag+1) folded into the row state.
__device__ __forceinline__ void fold32(uint32_t (&r)[32], uint32_t mul256, uint32_t tag, RowState& st) {
  const uint32_t tag_b = tag + 1u;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    r[j] = r[j] * mul256 + tag;
    r[16 + j] = r[16 + j] * mul256 + tag_b;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) st.acc[j] = umin3(st.acc[j], r[j], r[16 + j]);
  const uint32_t ha = chunk_min16(r), hb = chunk_min16(r + 16);     // ha != hb (different tags)
  const uint32_t lo = min(ha, hb), hi = max(ha, hb);
  st.h2 = umin3(max(st.h1, lo), st.h2, hi);
  st.h1 = min(st.h1, lo);
}

// Row state -> the sub-split's three candidate keys.  (key, class) order == (d^2, column) order.
__device__ __forceinline__ void emit_candidates(const RowState& st, int col_base, mkey_t* out) {
  uint32_t b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
  int j1 = 0, j2 = 0;
#pragma unroll
  for (int j = 0; j < tc::CHUNK; ++j) {
    const uint32_t k = st.acc[j];
    if (k < b1) { b2 = b1; j2 = j1; b1 = k; j1 = j; }
    else if (k < b2) { b2 = k; j2 = j; }
  }
  auto key_d2 = [](uint32_t k) { return (k >> 8) & 0x7FFFFFu; };
  auto key_col = [&](uint32_t k, int j) { return col_base + (int)(k & 0xFFu) * tc::CHUNK + j; };
  mkey_t o1 = MKEY_INF, o2 = MKEY_INF, o3 = MKEY_INF;
  if (key_d2(b1) <= 2u * tc::MAX_SQNORM) o1 = make_key((float)key_d2(b1), key_col(b1, j1));
  if (st.h2 < b2) {                      // runner-up hidden behind e1 in class j1: it is chunk h2's minimum
    if (key_d2(st.h2) <= 2u * tc::MAX_SQNORM) o2 = make_key((float)key_d2(st.h2), key_col(st.h2, j1));
  } else if (key_d2(b2) <= 2u * tc::MAX_SQNORM) {
    o2 = make_key((float)key_d2(b2), key_col(b2, j2));
    // same d^2 in the same chunk: the chunk may also hold an equal element in class j1 (hidden behind
    // e1); if j1 < j2 it would precede o2.  K1c checks that column when it matters.
    if (st.h2 == b2 && j1 < j2) o3 = make_key((float)key_d2(b2), key_col(b2, j1));
  }
  out[0] = o1;
  out[1] = o2;
  out[2] = o3;
}

// ---------------------------------------------------------------------------- cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default
// (.release.cta) semantics on purpose: what the barriers order here travels through the async proxy
// (TMA writes, tcgen05 reads of TMEM/smem) and is already complete when the arrive is issued
// (mbarrier complete_tx / tcgen05.wait::ld); `.release.cluster` would put a MEMBAR.ALL.GPU
// (~1000 cycles, measured) in front of every arrive.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// ============================================================================ K1 kernel
// QT = query tiles per CTA and item (the pair holds 2*QT tiles = 256*QT query rows per train stage).
template <int QT, int MODE>   // MODE 0: product, 1: dump raw accumulators, 2: record a timeline
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tc::NTHREADS, 1) match_tc_kernel(TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + tc::SMEM_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + tc::SMEM_BAR + tc::NBAR * 8);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < tc::NBAR; ++i) mbar_init(bar(i), i >= tc::ACC_EMPTY ? 16u : 1u);   // ACC_EMPTY: 8 warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: all 512 columns = two accumulator slots of 256 columns, allocated for the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // the peer's barriers are initialised before anything can arrive on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr bool DUMP = MODE == 1;
  const bool stamp = MODE == 2 && pair == 0 && rank == 0;
  auto tick = [&](uint32_t job, int k) { if (MODE == 2 && stamp) p.timeline[(size_t)job * 8 + k] = clock64(); };

  if (warp == 0) {
    // ------------------------------------------------------------------ B producer (this CTA's half of every stage)
    if (lane == 0) {
      uint32_t seq = 0;
      PairCursor cur(p);
      for (int item = pair; item < p.n_items; item += n_pairs) {
        const int split = cur.seek(item) / cur.pd.n_groups;
        const int s_begin = cur.s_begin_of(split), s_end = cur.s_begin_of(split + 1);
        const unsigned char* t_tiles = cur.pd.t_tiles;
        for (int s = s_begin; s < s_end; ++s, ++seq) {
          const int slot = (int)(seq % tc::NB);
          const uint32_t ph = (seq / tc::NB) & 1u;
          mbar_wait(bar(tc::B_EMPTY + slot), ph ^ 1u);
          mbar_expect_tx(bar(tc::B_FULL + slot), tc::TILE_BYTES);
          if (QT == 2) tick(2 * seq + 1, 0);      // timeline: when this stage's load was issued
          bulk_g2s(sbase + tc::SMEM_B + slot * tc::TILE_BYTES, t_tiles + (size_t)(2 * s + (int)rank) * tc::TILE_BYTES,
                   tc::TILE_BYTES, bar(tc::B_FULL + slot));
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ A producer (this CTA's query tiles)
    if (lane == 0) {
      uint32_t seq = 0;
      PairCursor cur(p);
      for (int item = pair; item < p.n_items; item += n_pairs) {
        const int g = cur.seek(item) % cur.pd.n_groups;
        for (int t = 0; t < QT; ++t, ++seq) {
          const int slot = (int)(seq & 1u);
          const uint32_t ph = (seq >> 1) & 1u;
          const int qt = min((g * QT + t) * 2 + (int)rank, cur.pd.n_qtiles_alloc - 1);
          mbar_wait(bar(tc::A_EMPTY + slot), ph ^ 1u);
          mbar_expect_tx(bar(tc::A_FULL + slot), tc::TILE_BYTES);
          bulk_g2s(sbase + tc::SMEM_A + slot * tc::TILE_BYTES, cur.pd.q_tiles + (size_t)qt * tc::TILE_BYTES, tc::TILE_BYTES,
                   bar(tc::A_FULL + slot));
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader) / data-landed relay (peer)
    if (lane == 0) {
      const uint32_t idesc = make_idesc(256, tc::STAGE_COLS);
      uint32_t a_seq = 0, b_seq = 0, job = 0;
      PairCursor cur(p);
      for (int item = pair; item < p.n_items; item += n_pairs) {
        const int split = cur.seek(item) / cur.pd.n_groups;
        const int s_begin = cur.s_begin_of(split), s_end = cur.s_begin_of(split + 1);
        for (int s = s_begin; s < s_end; ++s, ++b_seq) {
          const int bslot = (int)(b_seq % tc::NB);
          const uint32_t bph = (b_seq / tc::NB) & 1u;
          mbar_wait(bar(tc::B_FULL + bslot), bph);
          tick(job, 0);
          if (rank == 0) mbar_wait(bar(tc::B_PEER + bslot), bph);
          else mbar_arrive_remote(bar(tc::B_PEER + bslot), 0);
          tick(job, 1);
          const uint32_t b_addr = sbase + tc::SMEM_B + bslot * tc::TILE_BYTES;
          for (int t = 0; t < QT; ++t, ++job) {
            const int aslot = (int)((a_seq + t) & 1u);
            const uint32_t aph = ((a_seq + t) >> 1) & 1u;
            if (s == s_begin) {
              mbar_wait(bar(tc::A_FULL + aslot), aph);
              if (rank == 0) mbar_wait(bar(tc::A_PEER + aslot), aph);
              else mbar_arrive_remote(bar(tc::A_PEER + aslot), 0);
            }
            if (rank != 0) continue;
            const int cslot = (int)(job & 1u);
            tick(job, 2);
            mbar_wait(bar(tc::ACC_EMPTY + cslot), ((job >> 1) & 1u) ^ 1u);
            tc_fence_after();
            tick(job, 3);
            if (!(p.debug & 2u)) {
              const uint32_t a_addr = sbase + tc::SMEM_A + aslot * tc::TILE_BYTES;
              const uint32_t d_addr = tmem_base + (uint32_t)(cslot * tc::STAGE_COLS);
#pragma unroll
              for (int k = 0; k < tc::KMAIN / 16; ++k)
                tc_mma_bf16_pair(d_addr, make_smem_desc(a_addr + k * 2 * tc::LBO), make_smem_desc(b_addr + k * 2 * tc::LBO),
                                 idesc, k > 0 ? 1u : 0u);
              tc_mma_bf16_pair(d_addr, make_smem_desc(a_addr + 16 * tc::LBO), make_smem_desc(b_addr + 18 * tc::LBO), idesc, 1u);
            }
            tc_commit_pair(bar(tc::ACC_FULL + cslot));                     // accumulator complete (both CTAs)
            tick(job, 4);
            if (s == s_end - 1) tc_commit_pair(bar(tc::A_EMPTY + aslot));    // query tile slot reusable
          }
          if (rank == 0) tc_commit_pair(bar(tc::B_EMPTY + bslot));           // train tile slot reusable
        }
        a_seq += QT;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 3..10)
    // Two warps per TMEM lane quarter (= per SM sub-partition); warp set `half` takes columns
    // half*128 .. half*128+127 of every accumulator, one query row per thread.
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;          // query row within a tile
    const int half = (warp - 3) >> 2;
    const uint32_t mul256 = p.key_mul;  // 256, opaque to the compiler: the key build stays an IMAD (FMA pipe),
                                        // leaving the integer min/max pipe to the minima
    uint32_t job = 0;
    PairCursor cur(p);
    for (int item = pair; item < p.n_items; item += n_pairs) {
      const int local = cur.seek(item);
      const int g = local % cur.pd.n_groups, split = local / cur.pd.n_groups;
      const int s_begin = cur.s_begin_of(split), s_end = cur.s_begin_of(split + 1);
      RowState st[QT];
#pragma unroll
      for (int t = 0; t < QT; ++t) st[t].reset();
      for (int s = s_begin; s < s_end; ++s) {
#pragma unroll
        for (int t = 0; t < QT; ++t, ++job) {
          const int cslot = (int)(job & 1u);
          mbar_wait(bar(tc::ACC_FULL + cslot), (job >> 1) & 1u);
          tc_fence_after();
          if (warp == 3 && lane == 0) tick(job, 5);
          const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cslot * tc::STAGE_COLS + half * 128);
          const uint32_t tag0 = (uint32_t)((s - s_begin) * (tc::STAGE_COLS / tc::CHUNK) + half * (128 / tc::CHUNK));
          if (!(p.debug & 4u)) {
            uint32_t ra[32], rb[32];
            float* drow = nullptr;
            if (DUMP) {
              const int qt = (g * QT + t) * 2 + (int)rank;
              drow = p.dump + (size_t)(qt * 128 + row) * p.dump_cols + s * tc::STAGE_COLS + half * 128;
              if (qt >= cur.pd.n_qtiles) drow = nullptr;
            }
            auto dump32 = [&](const uint32_t (&r)[32], int c) {
              if (DUMP && drow && s * tc::STAGE_COLS + half * 128 + c * 32 < p.dump_cols) {
#pragma unroll
                for (int j = 0; j < 32; ++j) drow[c * 32 + j] = __uint_as_float(r[j]);
              }
            };
            tmem_ld32(t_addr, ra);
            tmem_ld_wait_regs(ra);
            tmem_ld32(t_addr + 32, rb);           // next 32 columns in flight while these are folded
            dump32(ra, 0);
            if (p.debug & 1u) st[t].h1 = min(st[t].h1, ra[0] ^ ra[31]); else fold32(ra, mul256, tag0, st[t]);
            tmem_ld_wait_regs(rb);
            tmem_ld32(t_addr + 64, ra);
            dump32(rb, 1);
            if (p.debug & 1u) st[t].h1 = min(st[t].h1, rb[0] ^ rb[31]); else fold32(rb, mul256, tag0 + 2u, st[t]);
            tmem_ld_wait_regs(ra);
            tmem_ld32(t_addr + 96, rb);
            dump32(ra, 2);
            if (p.debug & 1u) st[t].h1 = min(st[t].h1, ra[0] ^ ra[31]); else fold32(ra, mul256, tag0 + 4u, st[t]);
            tmem_ld_wait_regs(rb);
            // every column of this warp's share is in registers: hand the TMEM slot back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(bar(tc::ACC_EMPTY + cslot), 0);
            if (warp == 3 && lane == 0) tick(job, 6);
            dump32(rb, 3);
            if (p.debug & 1u) st[t].h1 = min(st[t].h1, rb[0] ^ rb[31]); else fold32(rb, mul256, tag0 + 6u, st[t]);
            if (warp == 3 && lane == 0) tick(job, 7);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(bar(tc::ACC_EMPTY + cslot), 0);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < QT; ++t) {
        const int qt = (g * QT + t) * 2 + (int)rank;
        if (qt < cur.pd.n_qtiles)
          emit_candidates(st[t], s_begin * tc::STAGE_COLS,
                          cur.pd.cand + ((size_t)(qt * 128 + row) * (2 * cur.pd.nsplit) + 2 * split + half) * 3);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs are done with the pair's TMEM and with each other's barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------- launch plan
// items = query groups x train splits, dealt round-robin to the CTA pairs.  The split count balances
// (a) whole waves over the pairs, (b) per-item ramp (query tile load) against item length, (c) the
// 8-bit chunk tag (<= 16 stages per split).
struct TcPlan { int qt, nsplit, n_groups, n_stages, n_items, n_pairs; };

static int tc_stages(int nt) { return div_up(div_up(nt, tc::TILE_ROWS), 2); }
static int tc_min_split(int n_stages) { return div_up(n_stages, tc::MAX_STAGES_PER_SPLIT); }

static TcPlan tc_plan(sfm_ctx* ctx, int nq, int nt) {
  TcPlan pl;
  const int n_qtiles = div_up(nq, tc::TILE_ROWS);
  pl.n_stages = tc_stages(nt);
  const int pairs = ctx->sm_count / 2 > 0 ? ctx->sm_count / 2 : 1;
  const char* e = getenv("SFM_MATCH_QT");
  int best_qt = 1, best_split = 1;
  double best = -1.0;
  for (int qt = 1; qt <= 2; ++qt) {
    if (e && atoi(e) != qt) continue;
    const int groups = div_up(n_qtiles, 2 * qt);
    const int smin = tc_min_split(pl.n_stages);
    for (int ns = smin; ns <= pl.n_stages && ns <= smin + 4 * pairs; ++ns) {
      const int items = groups * ns;
      const int waves = div_up(items, pairs);
      const double jobs = (double)div_up(pl.n_stages, ns) * qt;          // jobs of the longest item
      const double ramp = 1.5;                                           // query tile load, in job units
      double t = waves * (jobs + ramp) + 0.02 * ns;                      // slight preference for fewer splits
      if (qt == 1) t *= 1.04;                                            // QT=2 halves L2 traffic per job
      if (best < 0.0 || t < best) { best = t; best_qt = qt; best_split = ns; }
    }
  }
  const char* es = getenv("SFM_MATCH_NSPLIT");
  if (es && atoi(es) >= tc_min_split(pl.n_stages) && atoi(es) <= pl.n_stages) best_split = atoi(es);
  pl.qt = best_qt;
  pl.nsplit = best_split;
  pl.n_groups = div_up(n_qtiles, 2 * best_qt);
  pl.n_items = pl.n_groups * pl.nsplit;
  pl.n_pairs = pl.n_items < pairs ? pl.n_items : pairs;
  return pl;
}

// number of candidate sub-splits per query row of a single-pair launch (K1c merges them)
int sfm_match_tc_splits(sfm_ctx* ctx, int nq, int nt) { return 2 * tc_plan(ctx, nq, nt).nsplit; }

template <int QT, int MODE>
static int launch_tc_t(sfm_ctx* ctx, const TcParams& p, int n_pairs) {
  static bool attr_set = false;
  if (!attr_set) {
    SFM_CUDA(cudaFuncSetAttribute(match_tc_kernel<QT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    attr_set = true;
  }
  SFM_LAUNCH(ctx, SFM_K_MATCH_TC, (match_tc_kernel<QT, MODE><<<2 * n_pairs, tc::NTHREADS, tc::SMEM_BYTES, ctx->stream>>>(p)));
  return SFM_OK;
}

static void fill_pair(TcPair& d, const sfm_desc* q, const sfm_desc* t, mkey_t* cand, int qt, int nsplit, int item_begin) {
  d.q_tiles = (const unsigned char*)q->tiles;
  d.t_tiles = (const unsigned char*)t->tiles;
  d.cand = cand;
  d.n_qtiles = q->n_tiles;
  d.n_qtiles_alloc = (q->n_tiles + 1) & ~1;
  d.n_stages = tc_stages(t->n);
  d.n_groups = div_up(q->n_tiles, 2 * qt);
  d.nsplit = nsplit;
  d.item_begin = item_begin;
  d.n_items = d.n_groups * nsplit;
  d.pad_ = 0;
}

static int launch_tc(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, mkey_t* cand, int nsub, float* dump, int dump_cols,
                     long long* timeline = nullptr) {
  SFM_REQUIRE(q->tiles && t->tiles, "tensor-core matcher: descriptors have no tile image");
  TcPlan pl = tc_plan(ctx, q->n, t->n);
  SFM_REQUIRE(nsub == 2 * pl.nsplit, "tensor-core matcher: candidate buffer was sized for another plan");
  TcParams p;
  fill_pair(p.one, q, t, cand, pl.qt, pl.nsplit, 0);
  p.pairs = nullptr;
  p.n_items = pl.n_items;
  p.dump = dump;
  p.dump_cols = dump_cols;
  p.timeline = timeline;
  p.key_mul = 256u;
  { const char* e = getenv("SFM_MATCH_DEBUG"); p.debug = e ? (unsigned)atoi(e) : 0u; }
  if (p.debug & 16u) { p.n_items = 0; p.one.n_items = 0; }     // diagnostics: launch + setup + teardown only
  if (timeline) return pl.qt == 2 ? launch_tc_t<2, 2>(ctx, p, pl.n_pairs) : launch_tc_t<1, 2>(ctx, p, pl.n_pairs);
  if (dump) return pl.qt == 2 ? launch_tc_t<2, 1>(ctx, p, pl.n_pairs) : launch_tc_t<1, 1>(ctx, p, pl.n_pairs);
  return pl.qt == 2 ? launch_tc_t<2, 0>(ctx, p, pl.n_pairs) : launch_tc_t<1, 0>(ctx, p, pl.n_pairs);
}

int sfm_match_tc_launch(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, mkey_t* cand, int nsub) {
  return launch_tc(ctx, q, t, cand, nsub, nullptr, 0);
}

// Batched launch: ONE kernel over the items of all pairs.  nsub_out[k] = candidate sub-splits of pair k,
// cand_out[k] = its candidate array (workspace).  Every pair must be tensor-core eligible.
int sfm_match_tc_launch_batched(sfm_ctx* ctx, int npairs, const sfm_desc* const* q, const sfm_desc* const* t,
                                mkey_t** cand_out, int* nsub_out) {
  const int cta_pairs = ctx->sm_count / 2 > 0 ? ctx->sm_count / 2 : 1;
  // query tiles per CTA: 2 unless the whole batch is too small to fill the machine that way
  long long groups2 = 0;
  for (int k = 0; k < npairs; ++k) groups2 += (long long)div_up(q[k]->n_tiles, 4) * tc_min_split(tc_stages(t[k]->n));
  const int qt = groups2 >= cta_pairs ? 2 : 1;
  long long base_items = 0;
  for (int k = 0; k < npairs; ++k) base_items += (long long)div_up(q[k]->n_tiles, 2 * qt) * tc_min_split(tc_stages(t[k]->n));
  // more splits only while the item list is shorter than ~3 waves
  int scale = 1;
  if (base_items < 3LL * cta_pairs) scale = (int)div_up64(3LL * cta_pairs, base_items > 0 ? base_items : 1);
  std::vector<TcPair> host((size_t)(npairs > 0 ? npairs : 1));   // pageable: copied by cudaMemcpyAsync before it returns
  TcPair* dev = nullptr;
  SFM_TRY(ws_alloc_t(ctx, (size_t)(npairs > 0 ? npairs : 1), &dev));
  int item = 0;
  for (int k = 0; k < npairs; ++k) {
    SFM_REQUIRE(q[k]->tiles && t[k]->tiles, "tensor-core matcher: descriptors have no tile image");
    const int stages = tc_stages(t[k]->n);
    int ns = tc_min_split(stages) * scale;
    if (ns > stages) ns = stages;
    mkey_t* cand = nullptr;
    SFM_TRY(ws_alloc_t(ctx, (size_t)q[k]->n_tiles * 128 * 2 * ns * 3, &cand));
    fill_pair(host[k], q[k], t[k], cand, qt, ns, item);
    item += host[k].n_items;
    cand_out[k] = cand;
    nsub_out[k] = 2 * ns;
  }
  SFM_CUDA(cudaMemcpyAsync(dev, host.data(), sizeof(TcPair) * npairs, cudaMemcpyHostToDevice, ctx->stream));
  TcParams p;
  p.one = host[0];
  p.pairs = dev;
  p.n_items = item;
  p.dump = nullptr;
  p.dump_cols = 0;
  p.timeline = nullptr;
  p.key_mul = 256u;
  { const char* e = getenv("SFM_MATCH_DEBUG"); p.debug = e ? (unsigned)atoi(e) : 0u; }
  const int n_pairs = item < cta_pairs ? item : cta_pairs;
  if (n_pairs == 0) return SFM_OK;
  return qt == 2 ? launch_tc_t<2, 0>(ctx, p, n_pairs) : launch_tc_t<1, 0>(ctx, p, n_pairs);
}

// Debug/self-test entry (not part of the reference-facing surface): raw accumulators of every
// (query row, train column), i.e. -(2^22 + d^2/2), as float32 [n_qtiles*128][n_ttiles*128].
extern "C" int sfm_debug_match_tc_dump(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, float* dump_host,
                                       int64_t capacity) {
  SFM_REQUIRE(ctx && q && t && dump_host, "sfm_debug_match_tc_dump: null argument");
  SFM_TRY(sfm_desc_resolve(const_cast<sfm_desc*>(q)));
  SFM_TRY(sfm_desc_resolve(const_cast<sfm_desc*>(t)));
  SFM_TRY(sfm_ws_begin(ctx));
  const int nsub = sfm_match_tc_splits(ctx, q->n, t->n);
  const int dump_cols = t->n_tiles * 128;
  size_t count = (size_t)q->n_tiles * 128 * dump_cols;
  SFM_REQUIRE((int64_t)count <= capacity, "dump buffer too small: need %zu floats", count);
  mkey_t* cand;
  float* dump;
  SFM_TRY(ws_alloc_t(ctx, (size_t)q->n_tiles * 128 * nsub * 3, &cand));
  SFM_TRY(ws_alloc_t(ctx, count, &dump));
  SFM_TRY(launch_tc(ctx, q, t, cand, nsub, dump, dump_cols));
  SFM_CUDA(cudaMemcpyAsync(dump_host, dump, count * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFM_OK;
}

// Debug entry: clock64 stamps of the first `max_jobs` jobs of CTA pair 0 (leader SM), 8 per job:
// MMA thread: 0 own train tile landed, 1 peer's landed, 2 ready to issue, 3 accumulator slot free, 4 MMAs issued;
// epilogue warp 3: 5 accumulator complete, 6 TMEM slot released, 7 job folded.  Returns the job count of pair 0.
extern "C" int sfm_debug_match_tc_timeline(sfm_ctx* ctx, const sfm_desc* q, const sfm_desc* t, long long* stamps_host,
                                           int max_jobs, int* info /* qt, nsplit, n_items, n_pairs, jobs_pair0 */) {
  SFM_REQUIRE(ctx && q && t && stamps_host && info, "sfm_debug_match_tc_timeline: null argument");
  SFM_TRY(sfm_desc_resolve(const_cast<sfm_desc*>(q)));
  SFM_TRY(sfm_desc_resolve(const_cast<sfm_desc*>(t)));
  SFM_TRY(sfm_ws_begin(ctx));
  TcPlan pl = tc_plan(ctx, q->n, t->n);
  int jobs = 0;
  for (int item = 0; item < pl.n_items; item += pl.n_pairs) {
    const int split = item / pl.n_groups;
    jobs += (int)(((long long)(split + 1) * pl.n_stages) / pl.nsplit - ((long long)split * pl.n_stages) / pl.nsplit) * pl.qt;
  }
  SFM_REQUIRE(jobs <= max_jobs, "timeline buffer too small: %d jobs", jobs);
  mkey_t* cand;
  long long* tl;
  SFM_TRY(ws_alloc_t(ctx, (size_t)q->n_tiles * 128 * 2 * pl.nsplit * 3, &cand));
  SFM_TRY(ws_alloc_t(ctx, (size_t)jobs * 8, &tl));
  SFM_CUDA(cudaMemsetAsync(tl, 0, (size_t)jobs * 8 * sizeof(long long), ctx->stream));
  SFM_TRY(launch_tc(ctx, q, t, cand, 2 * pl.nsplit, nullptr, 0, tl));
  SFM_CUDA(cudaMemcpyAsync(stamps_host, tl, (size_t)jobs * 8 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  SFM_CUDA(cudaStreamSynchronize(ctx->stream));
  info[0] = pl.qt; info[1] = pl.nsplit; info[2] = pl.n_items; info[3] = pl.n_pairs; info[4] = jobs;
  return SFM_OK;
}
