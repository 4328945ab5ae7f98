// gmm_tc.cu -- K2 on the 5th-generation tensor cores: diagonal-GMM log-likelihoods as the dense contraction
//   C[t, m] = [x s, (x s)^2] . [mu/(sigma^2 s), -1/(2 sigma^2 s^2)]^T * log2(e)  (+ g_m log2(e) in the epilogue)      (m = Gaussian)
// issued as tcgen05.mma (kind::f16, fp32 accumulators in TMEM), followed by a per-pdf log-sum-exp computed by the
// epilogue warps straight out of TMEM (tcgen05.ld), one frame per thread.
//
// Replaces DecodableAmDiagGmmScaled::LogLikelihoodZeroBased / gmm_compute_likes (reference call sites:
// montreal_forced_aligner/alignment/multiprocessing.py:846 (inside GmmAligner), :1415); semantics SURVEY.md A.4.
//
// Split precision: both operands are split into fp16 hi + lo parts (22 significand bits) and three products are
// accumulated (hi*hi, hi*lo, lo*hi); features are pre-scaled per dimension by a power of two s_d (folded into the
// weights) so x s and (x s)^2 sit well inside fp16's range.
//
// Data movement: operands live in global memory ALREADY in the UMMA canonical K-major no-swizzle layout
// ([k/8][row/8][8 rows][8 halves], core matrix = 128 contiguous bytes), so a tile is one contiguous image and
// is fetched with a single 1-D bulk copy (cp.async.bulk -> UBLKCP) completing on an mbarrier; no tensor maps.
// Each CTA keeps the A images of two frame tiles (256 frames) resident and streams the Gaussian tiles through a
// ring, so every B image fetched from L2 feeds two accumulators.  TMEM holds 2 stages x 2 accumulators of 128
// columns (all 512 columns): the MMA warp runs one Gaussian tile ahead of the two epilogue warpgroups.
//
// Width-class tiles (round 2).  A Gaussian tile holds pdfs of ONE width class W (4, 8, ..., 32, 48, 64 or 128 columns per pdf:
// the pdf's component count rounded up to the class), floor(128 / W) pdfs per tile, so the column range of every pdf is a
// compile-time constant of the tile's class and the epilogue is straight-line code per class -- no segment masks, no predicates
// per 4-column group, no reconvergence barriers: 1 333 -> ~700 SASS instructions per 128-column tile (the round-1 epilogue was
// what the tensor pipe waited for).  The per-tile side data (gconst per column, output row per pdf slot, the class) travels
// with the accumulator stage as one 672-byte bulk copy.
#include <cuda_fp16.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "cuda_internal.cuh"

using namespace mfa;

namespace {

constexpr int TM = 128, TN = MFA_TILE_N;
// Two geometries.  K = 80 (2 dim <= 80; MFA's 39 / 40-dimensional features): the gconst is added by the epilogue from the tile's side
// data, 5 k-steps x 3 products = 15 MMAs per tile, 40 KB tiles, a THREE-stage B ring.  K = 96 (2 dim + 3 <= 96, or engine option tc_k96):
// the gconst rides as three fp16 columns against ones, 18 MMAs per tile, 48 KB tiles, two-stage ring (the first version of this kernel).
__host__ __device__ constexpr uint32_t img_bytes(int tk) { return (uint32_t)(TM * tk * 2); }   // one fp16 image (hi or lo) of a 128 x tk tile
__host__ __device__ constexpr uint32_t tile_bytes(int tk) { return 2 * img_bytes(tk); }         // hi + lo
constexpr uint32_t LBO_BYTES = (TM / 8) * 128;    // K-adjacent core matrices
constexpr uint32_t SBO_BYTES = 128;               // row-group-adjacent core matrices
constexpr int NTHREADS = 640;   // 4 control warps + 16 epilogue warps (4 per SM sub-partition)
constexpr float kLn2 = 0.69314718055994530942f, kLog2e = 1.44269504088896340736f;
// Timing experiments (WRONG results), compiled only by tools/k2_experiment.py with -DMFA_TC_EXP=n into a scratch library, never shipped:
// 1 no stores, 2 epilogue loads TMEM and releases (no arithmetic, no stores), 3 epilogue releases without reading, 4 one k-step per
// product (3 MMAs per tile instead of 15), 5 B tiles fetched once per item (no B traffic), 6 = 2 + 5, 7 = correct results + cycle counters
// of every role's barrier waits (printed per launch)
#ifndef MFA_TC_EXP
#define MFA_TC_EXP 0
#endif

// Width classes: pdfs of up to W components share tiles of class W.
constexpr int NCLS = 14;
// slot widths 4 6 8 10 12 14 16 | 20 24 28 32 | 48 64 128: even widths up to 16 (a 10-component pdf in a 12-wide slot wasted a sixth of
// the tile: 10 pdfs per tile instead of 12), multiples of 4 up to 32, then the widths whose capacity still differs
__host__ __device__ constexpr int cls_width(int c) { return c < 7 ? 4 + 2 * c : (c < 11 ? 20 + 4 * (c - 7) : (c == 11 ? 48 : (c == 12 ? 64 : 128))); }
__host__ __device__ constexpr int cls_cap(int c) { return TN / cls_width(c); }               // pdfs per tile: 32 21 16 12 10 9 8 6 5 4 4 2 2 1
__host__ __device__ inline int cls_of(int ng) { return ng <= 4 ? 0 : (ng <= 16 ? (ng + 1) / 2 - 2 : (ng <= 32 ? 7 + (ng - 17) / 4 : (ng <= 48 ? 11 : (ng <= 64 ? 12 : 13)))); }

// Side data of one Gaussian tile (bulk-copied into shared memory next to the accumulator stage it belongs to).
struct TcAux {
  float g[TN];        // gconst * log2(e) per column (padding columns: -60000 -> exp2 -> 0); written by gather_b_kernel
  int32_t row[32];    // output row of pdf slot j (utterance-local pdf index, or global pdf id in the dense tiling); -1 = empty slot
  int32_t cls;        // width class of the tile
  int32_t npdf;       // occupied slots
  int32_t lp_base;    // ragged tiles: index of the utterance's first entry in the graphs' lp2pdf list (pdf of slot j = lp2pdf[lp_base + row[j]]); dense: -1 (pdf = row[j])
  int32_t pad[5];
};
static_assert(sizeof(TcAux) == 672 && sizeof(TcAux) % 16 == 0, "TcAux must be 672 bytes");
constexpr uint32_t AUX_BYTES = sizeof(TcAux);
constexpr int IRING = 8;        // slots of the CTA's work-item ring (gmm_tc_kernel)

// ---- PTX helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or the hint (ns) expires, so the loop
// below turns over a few times per wait at most.  The poll counter is the hang guard (a protocol bug must trap, not hang the GPU):
// one integer add per failed poll, no clock reads on the hot path.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, polls = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity), "r"(0x100000u) : "memory");
    if (done) break;
    if (++polls > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
  // layout_type [61,64) = 0 (no swizzle / interleave)
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(LBO_BYTES >> 4) << 16;
  d |= (uint64_t)(SBO_BYTES >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// exp2 on the FMA pipe (Cody-Waite range reduction + degree-5 polynomial, ~2e-7 relative on x <= 0): an alternative to MUFU.EX2 for a
// fraction of the epilogue's exponentials (engine option tc_poly; A/B measured in DESIGN.md section 8).  x <= 0; results below 2^-126 flush to 0.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);           // in [-0.5, 0.5]
  float p = 1.33336498402e-3f;
  p = fmaf(p, f, 9.61812911e-3f);
  p = fmaf(p, f, 5.55041087e-2f);
  p = fmaf(p, f, 2.40226507e-1f);
  p = fmaf(p, f, 6.93147181e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- epilogue of one accumulator (128 frames x 128 columns) whose tile has width class W: thread = frame (TMEM lane).  The tile's
// pdf slot j owns columns [j W, (j + 1) W); the accumulator is read in four 32-column chunks, a pdf that straddles a chunk edge carries
// (max, sum) across it.  Everything about the segmentation is a compile-time constant after unrolling.
template <int W, bool GEPI, bool POLY>
__device__ __forceinline__ void epi_tile(const uint32_t t0, const TcAux *__restrict__ ax, float *__restrict__ out_base, const uint32_t ld,
                                         const bool row_ok, uint64_t *tempty_bar) {
  constexpr int NP = TN / W;                       // pdf slots per tile
  constexpr int NCH = (NP * W + 31) / 32;          // chunks that hold pdf columns
  float cm = 0.0f, cs = 0.0f;
  uint32_t v[32];
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    tmem_ld32(t0 + 32 * c, v);
    tmem_ld_wait();
    if (c == NCH - 1) { tc_fence_before(); mbar_arrive(tempty_bar); }   // accumulator drained: the MMA warp may overwrite it
#if MFA_TC_EXP == 2 || MFA_TC_EXP == 6
    if (__uint_as_float(v[0]) == 1.2345e-30f && row_ok) out_base[0] = 0.0f;
    continue;
#endif
    if (GEPI) {   // gconst of the chunk's columns from shared memory (every thread reads the same words: broadcasts)
#pragma unroll
      for (int q4 = 0; q4 < 8; q4++) {
        if (32 * c + 4 * q4 >= NP * W) break;
        const float4 q = *reinterpret_cast<const float4 *>(ax->g + 32 * c + 4 * q4);
        v[4 * q4] = __float_as_uint(__uint_as_float(v[4 * q4]) + q.x); v[4 * q4 + 1] = __float_as_uint(__uint_as_float(v[4 * q4 + 1]) + q.y);
        v[4 * q4 + 2] = __float_as_uint(__uint_as_float(v[4 * q4 + 2]) + q.z); v[4 * q4 + 3] = __float_as_uint(__uint_as_float(v[4 * q4 + 3]) + q.w);
      }
    }
#pragma unroll
    for (int j = 0; j < NP; j++) {
      const int lo = (j * W > 32 * c ? j * W : 32 * c) - 32 * c, hi = ((j + 1) * W < 32 * c + 32 ? (j + 1) * W : 32 * c + 32) - 32 * c;
      if (lo >= hi) continue;
      const bool starts = j * W >= 32 * c, ends = (j + 1) * W <= 32 * c + 32;
      // (lo, hi and W are even: the loops run over column pairs with literal indices, which keeps v[] in registers)
      float m = fmaxf(__uint_as_float(v[lo]), __uint_as_float(v[lo + 1]));
#pragma unroll
      for (int q2 = 1; q2 < 16; q2++) {
        if (lo + 2 * q2 >= hi) break;
        const int i = lo + 2 * q2;
        m = fmaxf(m, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
      }
      float s0 = 0.0f, s1 = 0.0f;
      if (!starts) { const float nm = fmaxf(cm, m); s0 = cs * ex2(cm - nm); m = nm; }
#pragma unroll
      for (int q2 = 0; q2 < 16; q2++) {
        if (lo + 2 * q2 >= hi) break;
        const int i = lo + 2 * q2;
        s0 += ex2(__uint_as_float(v[i]) - m);
        s1 += (POLY && (q2 & 1)) ? ex2_poly(__uint_as_float(v[i + 1]) - m) : ex2(__uint_as_float(v[i + 1]) - m);
      }
      const float s = s0 + s1;
      if (ends) {
        const int row = ax->row[j];
#if MFA_TC_EXP == 1
        if (row_ok && row == -12345) out_base[(size_t)((uint32_t)row * ld)] = (m + lg2(s)) * kLn2;
#else
        if (row_ok && row >= 0) out_base[(size_t)((uint32_t)row * ld)] = (m + lg2(s)) * kLn2;
#endif
      } else { cm = m; cs = s; }
    }
  }
}

// One work item = one pair of frame tiles (256 frames) against a run of Gaussian tiles.
struct TcItem {
  uint32_t a_tile;       // first of the two A images of the pair
  uint32_t b_tile0, n_b; // Gaussian tiles [b_tile0, b_tile0 + n_b): images in b_img, side data in aux
  uint32_t rows_valid;   // frames of the pair that exist (<= 256)
  uint64_t out_off;      // float offset of the pair's first frame in `out`
  uint32_t ld, pad;      // leading dimension (frames) of this item's output block
};
static_assert(sizeof(TcItem) == 32, "TcItem must be 32 bytes");

struct TcParams {
  const uint8_t *a_img;   // [n_frame_tiles (even)][tile_bytes]
  const uint8_t *b_img;   // [n_b_tiles][tile_bytes]
  const TcAux *aux;       // [n_b_tiles] side data (gconsts, output rows, width class)
  const TcItem *items;    // [n_items]
  int n_items;
  float *out;             // pdf-major blocks: out[item.out_off + row * item.ld + frame]
  int *counter;           // work counter of this launch (zeroed on the stream before it): items beyond the first gridDim.x are handed out dynamically
#if MFA_TC_EXP == 7
  long long *dbg;         // [grid][16] cycle counters
#endif
};
#if MFA_TC_EXP == 7
#define DBG_T0() const long long dbg_t0 = clock64()
#define DBG_ADD(slot) p.dbg[blockIdx.x * 16 + (slot)] += clock64() - dbg_t0
#else
#define DBG_T0()
#define DBG_ADD(slot)
#endif

template <int TKt, int NBt, bool GEPI, bool POLY>
__global__ void __launch_bounds__(NTHREADS, 1)
gmm_tc_kernel(TcParams p) {
  constexpr uint32_t IMG_BYTES = img_bytes(TKt), TILE_BYTES = tile_bytes(TKt);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sA = smem;                       // 2 frame tiles x (hi, lo)
  uint8_t *sB = smem + 2 * TILE_BYTES;      // NBt stages x (hi, lo)
  TcAux *sX = (TcAux *)(smem + (2 + NBt) * TILE_BYTES);   // 2 x side data (one per accumulator stage)
  uint64_t *bars = (uint64_t *)(smem + (2 + NBt) * TILE_BYTES + 2 * AUX_BYTES);
  uint64_t *full_a = bars + 0, *empty_a = bars + 1, *full_b = bars + 2, *empty_b = bars + 2 + NBt, *tfull = bars + 2 + 2 * NBt,
           *tempty = tfull + 4, *gfull = tempty + 4, *gempty = gfull + 2, *ifull = gempty + 2;
  // item ring: the producer thread draws the CTA's next work item from a global counter (the items are sorted longest first, so this is
  // longest-processing-time scheduling) and publishes it to the other roles.  A CTA that starts late -- its SM was still held by another
  // kernel: a Viterbi tail, another job's launch -- then simply finds less work left, where a static round-robin made the whole launch
  // wait for it.  No "empty" barrier: the producer is never more than four items ahead of the slowest reader (A buffer: 1, TMEM stages:
  // 2, fetch-ahead: 1), the ring has eight slots.
  volatile int32_t *s_item = (volatile int32_t *)(ifull + IRING);
  uint32_t *tmem_slot = (uint32_t *)(s_item + IRING);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(full_a, 1); mbar_init(empty_a, 2);    // two MMA issuers (one per frame tile of the pair) release the operand buffers
    for (int s = 0; s < NBt; s++) { mbar_init(full_b + s, 1); mbar_init(empty_b + s, 2); }
    for (int i = 0; i < 4; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 128); }
    for (int i = 0; i < 2; i++) { mbar_init(gfull + i, 1); mbar_init(gempty + i, 256); }
    for (int i = 0; i < IRING; i++) mbar_init(ifull + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_items = p.n_items;

  if (warp == 0) {
    // ===== producer: bulk copies of the A pair (once per item) and of each B tile =====
    if (lane == 0) {
      uint32_t sb = 0, phb = 0;   // sb / phb: slot of the B ring and the phase of its barriers
      int item = blockIdx.x;
      for (uint32_t it = 0;; it++) {
        if (item >= n_items) item = -1;
        s_item[it % IRING] = item;
        mbar_arrive(ifull + it % IRING);
        if (item < 0) break;
        const TcItem I = p.items[item];
        { DBG_T0(); mbar_wait(empty_a, (it & 1) ^ 1); DBG_ADD(0); }
        mbar_expect_tx(full_a, 2 * TILE_BYTES);
        bulk_g2s(sA, p.a_img + (size_t)I.a_tile * TILE_BYTES, 2 * TILE_BYTES, full_a);
        for (uint32_t n = I.b_tile0; n < I.b_tile0 + I.n_b; n++) {
          { DBG_T0(); mbar_wait(empty_b + sb, phb ^ 1); DBG_ADD(1); }
#if MFA_TC_EXP == 5 || MFA_TC_EXP == 6
          if (n >= I.b_tile0 + NBt) { mbar_arrive(full_b + sb); if (++sb == NBt) { sb = 0; phb ^= 1; } continue; }
#endif
          mbar_expect_tx(full_b + sb, TILE_BYTES);
          bulk_g2s(sB + sb * TILE_BYTES, p.b_img + (size_t)n * TILE_BYTES, TILE_BYTES, full_b + sb);
          if (++sb == NBt) { sb = 0; phb ^= 1; }
        }
        item = (int)gridDim.x + atomicAdd(p.counter, 1);
      }
    }
  } else if (warp == 3) {
    // ===== side-data producer: the tile's TcAux travels with the accumulator stage (cnt & 1) and is released by the epilogue warps
    // -- its own thread, so a slow epilogue never delays the B ring =====
    if (lane == 0) {
      uint32_t cnt = 0;
      for (uint32_t it = 0;; it++) {
        mbar_wait(ifull + it % IRING, (it / IRING) & 1);
        const int item = s_item[it % IRING];
        if (item < 0) break;
        const TcItem I = p.items[item];
        for (uint32_t n = I.b_tile0; n < I.b_tile0 + I.n_b; n++, cnt++) {
          const uint32_t sg = cnt & 1;
          { DBG_T0(); mbar_wait(gempty + sg, ((cnt >> 1) & 1) ^ 1); DBG_ADD(2); }
          mbar_expect_tx(gfull + sg, AUX_BYTES);
          bulk_g2s(sX + sg, p.aux + n, AUX_BYTES, gfull + sg);
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===== MMA issuers: one thread per frame tile of the pair (f = 0: warp 1, f = 1: warp 2).  The round-1 kernel issued both
    // accumulators from one thread; its cycle counters (tools/k2_experiment.py, variant 7) showed that thread busy 90 % of the time, ~70
    // cycles per tcgen05.mma of descriptor arithmetic and register -> uniform-register moves, i.e. the ISSUE rate bounded the kernel, not
    // the tensor pipe.  Two threads halve that, and the shared-memory descriptors are now one 32-bit add away from a per-tile base. =====
    if (lane == 0) {
      const int f = warp - 1;
      // InstrDescriptor: c_format=F32 (1<<4), a/b format F16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      // SmemDescriptor: low word = (address >> 4) [0,14) | (LBO >> 4) << 16; high word = (SBO >> 4) | version 1 << 14 -- constant
      const uint32_t desc_hi = (SBO_BYTES >> 4) | (1u << 14);
      const uint32_t a_lo = (((smem_u32(sA) + f * TILE_BYTES) & 0x3FFFF) >> 4) | ((LBO_BYTES >> 4) << 16);
      const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFF) >> 4) | ((LBO_BYTES >> 4) << 16);
      uint32_t cnt = 0, sb = 0, phb = 0;
#if MFA_TC_EXP == 7
      const long long dbg_loop0 = clock64();
#endif
      for (uint32_t it = 0;; it++) {
        mbar_wait(ifull + it % IRING, (it / IRING) & 1);
        const int item = s_item[it % IRING];
        if (item < 0) break;
        const uint32_t n_b = p.items[item].n_b, rows_valid = p.items[item].rows_valid;
        const bool live = f == 0 || rows_valid > TM;   // a pair whose second tile holds no frames skips its MMAs
        { DBG_T0(); mbar_wait(full_a, it & 1); if (f == 0) DBG_ADD(3); }
        tc_fence_after();
        for (uint32_t n = 0; n < n_b; n++, cnt++) {
          const uint32_t s = cnt & 1, ph = (cnt >> 1) & 1;
          { DBG_T0(); mbar_wait(full_b + sb, phb); if (f == 0) DBG_ADD(4); }
          { DBG_T0(); mbar_wait(tempty + s * 2 + f, ph ^ 1); if (f == 0) DBG_ADD(5); }
          tc_fence_after();
          DBG_T0();
          if (live) {
            const uint32_t d = tmem_base + s * 256 + f * 128;
            const uint32_t b_lo = b_lo0 + sb * (TILE_BYTES >> 4);
#pragma unroll
            for (int prod = 0; prod < 3; prod++) {
#pragma unroll
              for (int k = 0; k < (MFA_TC_EXP == 4 ? 1 : TKt / 16); k++) {
                const uint32_t ao = (uint32_t)((k * 2 * LBO_BYTES + (prod == 2 ? IMG_BYTES : 0)) >> 4);
                const uint32_t bo = (uint32_t)((k * 2 * LBO_BYTES + (prod == 1 ? IMG_BYTES : 0)) >> 4);
                umma_f16(d, ((uint64_t)desc_hi << 32) | (a_lo + ao), ((uint64_t)desc_hi << 32) | (b_lo + bo), idesc, (prod | k) != 0);
              }
            }
          }
          umma_commit(tfull + s * 2 + f);
          umma_commit(empty_b + sb);
          if (f == 0) DBG_ADD(6);
#if MFA_TC_EXP == 7
          if (f == 0) p.dbg[blockIdx.x * 16 + 12] += 1;
#endif
          if (++sb == NBt) { sb = 0; phb ^= 1; }
        }
        umma_commit(empty_a);
      }
#if MFA_TC_EXP == 7
      if (f == 0) p.dbg[blockIdx.x * 16 + 7] += clock64() - dbg_loop0;
#endif
    }
  } else if (warp >= 4) {
    // ===== epilogue: 16 warps = 2 accumulator stages x 2 frame tiles x 4 lane quarters; thread = one frame (TMEM lane).
    // The warps of stage s take every other Gaussian tile, so four epilogue warps share each SM sub-partition and the
    // TMEM-load / MUFU latency of one is covered by the others. =====
    const int e = warp - 4, wq = e & 3, f = (e >> 2) & 1, s = e >> 3;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    uint32_t cnt = 0;   // tiles seen by the CTA so far (all roles count alike); this warp serves those with (cnt & 1) == s
    const TcAux *ax = sX + s;
    for (uint32_t it = 0;; it++) {
      mbar_wait(ifull + it % IRING, (it / IRING) & 1);
      const int item = s_item[it % IRING];
      if (item < 0) break;
      const TcItem I = p.items[item];
      const uint32_t row = f * TM + wq * 32 + lane;
      const bool row_ok = row < I.rows_valid;
      const bool tile_live = (uint32_t)(f * TM) < I.rows_valid;
      float *out_base = p.out + I.out_off + row;
      for (uint32_t n = 0; n < I.n_b; n++, cnt++) {
        if ((cnt & 1u) != (uint32_t)s) continue;
        const uint32_t ph = (cnt >> 1) & 1;
#if MFA_TC_EXP == 7
        const bool dbg_me = warp == 4 && lane == 0;
        { const long long t0 = clock64(); mbar_wait(tfull + s * 2 + f, ph); if (dbg_me) p.dbg[blockIdx.x * 16 + 8] += clock64() - t0; }
        tc_fence_after();
        { const long long t0 = clock64(); mbar_wait(gfull + s, ph); if (dbg_me) p.dbg[blockIdx.x * 16 + 9] += clock64() - t0; }
        const long long dbg_e0 = clock64();
#else
        mbar_wait(tfull + s * 2 + f, ph);
        tc_fence_after();
        mbar_wait(gfull + s, ph);
#endif
#if MFA_TC_EXP == 3
        { tc_fence_before(); mbar_arrive(tempty + s * 2 + f); mbar_arrive(gempty + s); continue; }
#endif
        if (!tile_live) { tc_fence_before(); mbar_arrive(tempty + s * 2 + f); mbar_arrive(gempty + s); continue; }
        const uint32_t t0 = tmem_base + lane_base + s * 256 + f * 128;
        uint64_t *tb = tempty + s * 2 + f;
        switch (ax->cls) {   // uniform across the CTA: one class per tile (widths: cls_width)
          case 0: epi_tile<4, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          case 1: epi_tile<6, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          case 2: epi_tile<8, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          case 3: epi_tile<10, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          case 4: epi_tile<12, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          case 5: epi_tile<14, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          case 6: epi_tile<16, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          case 7: epi_tile<20, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          case 8: epi_tile<24, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          case 9: epi_tile<28, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          case 10: epi_tile<32, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          case 11: epi_tile<48, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          case 12: epi_tile<64, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
          default: epi_tile<128, GEPI, POLY>(t0, ax, out_base, I.ld, row_ok, tb); break;
        }
        mbar_arrive(gempty + s);   // this thread is done with the stage's side data
#if MFA_TC_EXP == 7
        if (dbg_me) { p.dbg[blockIdx.x * 16 + 10] += clock64() - dbg_e0; p.dbg[blockIdx.x * 16 + 11] += 1; }
#endif
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// features fp32 -> A images (fp16 hi / lo, canonical layout), one 48 KB image per 128-frame tile.  Tile t covers feature rows
// tile_row0[t] .. tile_row0[t] + tile_rows[t] - 1 (rows beyond that are zero), so tiles may follow utterance boundaries.
__global__ void xsplit_kernel(const float *__restrict__ feats, int dim, const float *__restrict__ colscale, uint8_t *__restrict__ a_img,
                              int64_t n_tiles, const int64_t *__restrict__ tile_row0, const int32_t *__restrict__ tile_rows, int64_t dense_rows,
                              int KC, int ones) {
  const uint32_t IMG_BYTES = img_bytes(KC * 8), TILE_BYTES = 2 * IMG_BYTES;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (tile, kc, row)
  if (idx >= n_tiles * KC * TM) return;
  const int r = (int)(idx % TM), kc = (int)((idx / TM) % KC);
  const int64_t tile = idx / (TM * KC);
  int64_t row; bool live;
  if (tile_row0) { live = r < tile_rows[tile]; row = tile_row0[tile] + r; }
  else { row = tile * TM + r; live = row < dense_rows; }
  __half hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const int k = kc * 8 + e;
    float a = 0.0f;
    if (live) {
      if (k < 2 * dim) {
        float x = feats[row * dim + (k < dim ? k : k - dim)] * colscale[k < dim ? k : k - dim];
        x = fminf(fmaxf(x, -240.0f), 240.0f);
        a = k < dim ? x : x * x;
      } else if (ones && k < 2 * dim + 3) a = 1.0f;
    }
    hi[e] = __float2half_rn(a);
    lo[e] = __float2half_rn(a - __half2float(hi[e]));
  }
  const size_t off = (size_t)tile * TILE_BYTES + ((size_t)(kc * (TM / 8) + r / 8) * 64 + (size_t)(r % 8) * 8) * 2;
  *(uint4 *)(a_img + off) = *(const uint4 *)hi;
  *(uint4 *)(a_img + off + IMG_BYTES) = *(const uint4 *)lo;
}

// B images: column t of a tile of class W belongs to pdf slot j = t / W, component t % W.  One CTA per (tile, hi | lo): the source rows
// (row-major fp16 hi/lo weight rows of the model) are read whole (consecutive threads = consecutive 16-byte pieces of a row: full
// sectors), re-ordered into the canonical [k/8][row/8][8 rows][8 halves] image in shared memory and written out linearly; the tile's
// gconst column (aux.g) is gathered here as well.  Padding columns get zero weights and gconst -60000 (K = 96 geometry: -60000 in the
// first gconst column of the weight row).
__global__ void __launch_bounds__(256)
gather_b_kernel(const __half *__restrict__ w_rows, int64_t num_gauss, uint8_t *__restrict__ b_img, int64_t n_tiles, int gcol, int KC,
                const float *__restrict__ g_src, TcAux *__restrict__ aux, const int32_t *__restrict__ lp2pdf, const int32_t *__restrict__ pdf_off) {
  __shared__ int s_src[TN];
  __shared__ int s_g0[32], s_ng[32];
  __shared__ uint4 s_img[(TN + 1) * 12];   // up to K = 96; one padding unit per k-chunk keeps the transposing stores conflict-free
  const int TK = KC * 8;
  const uint32_t IMG_BYTES = img_bytes(TK), TILE_BYTES = 2 * IMG_BYTES;
  const int64_t tile = blockIdx.x >> 1;
  const int which = blockIdx.x & 1, t = threadIdx.x;
  if (tile >= n_tiles) return;
  TcAux *ax = aux + tile;
  const int W = cls_width(ax->cls), NP = TN / W;
  if (t < 32) {
    int g0 = 0, ng = 0;
    if (t < NP) {
      const int row = ax->row[t];
      if (row >= 0) { const int pdf = ax->lp_base >= 0 ? lp2pdf[ax->lp_base + row] : row; g0 = pdf_off[pdf]; ng = pdf_off[pdf + 1] - g0; }
    }
    s_g0[t] = g0; s_ng[t] = ng;
  }
  __syncthreads();
  if (t < TN) {
    const int j = t / W, c = t - j * W;
    const int g = (j < NP && c < s_ng[j]) ? s_g0[j] + c : -1;
    s_src[t] = g;
    if (which == 0) ax->g[t] = g >= 0 ? g_src[g] : -60000.0f;
  }
  __syncthreads();
  for (int c = t; c < TN * KC; c += 256) {
    const int r = c / KC, kc = c - r * KC;
    const int g = s_src[r];
    uint4 v = make_uint4(0, 0, 0, 0);
    if (g >= 0) v = *(const uint4 *)(w_rows + ((size_t)which * num_gauss + g) * TK + kc * 8);
    else if (gcol >= 0 && which == 0 && kc == gcol / 8) {
      __half pad[8];
#pragma unroll
      for (int e = 0; e < 8; e++) pad[e] = __float2half_rn(e == gcol % 8 ? -60000.0f : 0.0f);
      v = *(const uint4 *)pad;
    }
    s_img[kc * (TN + 1) + r] = v;   // canonical image in 16-byte units: unit index = kc * 128 + r
  }
  __syncthreads();
  uint4 *dst = (uint4 *)(b_img + (size_t)tile * TILE_BYTES + (size_t)which * IMG_BYTES);
  for (int c = t; c < TN * KC; c += 256) dst[c] = s_img[(c >> 7) * (TN + 1) + (c & (TN - 1))];
}

// Per-utterance tile plan of the ragged path, on the device (rebuilt whenever an M-step changed some pdf's component count): one
// thread per utterance walks its pdf list (graphs' lp2pdf) and counts the pdfs per width class.  Tiles are then formed from the WIDEST
// class down: a tile has the class of its widest pdf and cap(class) slots, and the slots its own class leaves free are handed to the
// next narrower pdfs (a pdf fits any slot at least as wide as its component count; the gather pads).  Walking the pdfs in descending
// width and closing a tile only when it is full minimises the number of tiles per utterance -- with one class per tile and no
// hand-down a model whose pdfs spread over all eleven classes (any model after mix-up) paid up to ten partly filled tiles per
// utterance.  fill == 0: tiles per utterance; fill == 1 (given the utterance's first tile): the tiles' class and rows.
__global__ void rag_plan_kernel(int n_utts, const int64_t *__restrict__ lp_off, const int32_t *__restrict__ lp2pdf, const int32_t *__restrict__ pdf_off,
                                int fill, int32_t *__restrict__ n_tiles_out, const int64_t *__restrict__ tile_off, TcAux *__restrict__ aux) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_utts) return;
  const int64_t k0 = lp_off[u], k1 = lp_off[u + 1];
  int cnt[NCLS];
#pragma unroll
  for (int c = 0; c < NCLS; c++) cnt[c] = 0;
  for (int64_t k = k0; k < k1; k++) { const int pdf = lp2pdf[k]; cnt[cls_of(pdf_off[pdf + 1] - pdf_off[pdf])]++; }
  // class c: its first take[c] pdfs (list order) go to slots inh_slot[c].. of the open tile inh_tile[c] of a wider class, the rest to its
  // own tiles base[c], base[c] + 1, ...
  int64_t base[NCLS], inh_tile[NCLS];
  int take[NCLS], inh_slot[NCLS];
  int64_t t = fill ? tile_off[u] : 0, open_tile = -1;
  int room = 0, open_slot = 0;
  for (int c = NCLS - 1; c >= 0; c--) {
    int m = cnt[c];
    take[c] = room < m ? room : m; inh_tile[c] = open_tile; inh_slot[c] = open_slot;
    m -= take[c]; room -= take[c]; open_slot += take[c];
    base[c] = t;
    if (m > 0) {
      const int cap = cls_cap(c), nt = (m + cap - 1) / cap;
      if (fill)
        for (int i = 0; i < nt; i++) {
          TcAux *ax = aux + t + i;
          ax->cls = c; ax->lp_base = (int32_t)k0; ax->npdf = cap;
          for (int j = 0; j < 32; j++) ax->row[j] = -1;
        }
      open_tile = t + nt - 1; open_slot = m - (nt - 1) * cap; room = cap - open_slot;
      t += nt;
    }
    cnt[c] = 0;
  }
  if (!fill) { n_tiles_out[u] = (int)t; return; }
  for (int64_t k = k0; k < k1; k++) {
    const int pdf = lp2pdf[k];
    const int c = cls_of(pdf_off[pdf + 1] - pdf_off[pdf]);
    const int i = cnt[c]++;
    if (i < take[c]) aux[inh_tile[c]].row[inh_slot[c] + i] = (int32_t)(k - k0);
    else { const int j = i - take[c]; aux[base[c] + j / cls_cap(c)].row[j % cls_cap(c)] = (int32_t)(k - k0); }
  }
}

// ---- operand images from the device-resident natural-layout parameters (model creation and after every device M-step) ----
// second moments about zero per dimension (features are not re-centred): m2[d] += mu^2 + sigma^2 over the Gaussians
__global__ void __launch_bounds__(128)
tc_moment_kernel(const float *__restrict__ miv, const float *__restrict__ iv, int G, int D, double *__restrict__ m2) {
  const int d = threadIdx.x;
  if (d >= D) return;
  double acc = 0.0;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const double v = (double)iv[(size_t)g * D + d], mu = (double)miv[(size_t)g * D + d] / v;
    acc += mu * mu + 1.0 / v;
  }
  atomicAdd(&m2[d], acc);
}
// per-dimension power-of-two feature scale: x * s has unit order, so x s and (x s)^2 sit inside fp16's range
__global__ void tc_colscale_kernel(const double *__restrict__ m2, int G, int D, float *__restrict__ colscale) {
  const int d = threadIdx.x;
  if (d >= D) return;
  const double rms = sqrt(fmax(m2[d] / G, 1e-30));
  colscale[d] = (float)exp2(-round(log2(rms)));
}
// fp16 hi / lo weight rows [2][G][TK] (row-major: the source of every tile gather) and gconst * log2(e) per Gaussian
__global__ void tc_rows_kernel(const float *__restrict__ miv, const float *__restrict__ iv, const float *__restrict__ gconsts,
                               const float *__restrict__ colscale, int G, int D, int TK, int k80, __half *__restrict__ rows,
                               float *__restrict__ gl2, int32_t *__restrict__ flag) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)G * TK) return;
  const int g = (int)(idx / TK), k = (int)(idx - (int64_t)g * TK);
  double gc = (double)gconsts[g] * (double)kLog2e;
  if (!(gc > -60000.0)) gc = -60000.0;
  if (k == 0) gl2[g] = (float)gc;
  __half hi = __float2half_rn(0.0f), lo = hi;
  if (k < 2 * D) {
    const int d = k < D ? k : k - D;
    const double s = (double)colscale[d];
    const double w = k < D ? (double)miv[(size_t)g * D + d] / s * (double)kLog2e : -0.5 * (double)iv[(size_t)g * D + d] / (s * s) * (double)kLog2e;
    hi = __float2half_rn((float)w);
    lo = __float2half_rn((float)(w - (double)__half2float(hi)));
    if (fabs(w) > 60000.0) *flag = 1;
  } else if (!k80 && k < 2 * D + 3) {   // K = 96 geometry: the gconst as three fp16 columns against ones
    const float g1 = __half2float(__float2half_rn((float)gc));
    const float g2 = __half2float(__float2half_rn((float)(gc - g1)));
    const float g3 = (float)(gc - g1 - g2);
    hi = __float2half_rn(k == 2 * D ? g1 : (k == 2 * D + 1 ? g2 : g3));
  }
  rows[((size_t)0 * G + g) * TK + k] = hi;
  rows[((size_t)1 * G + g) * TK + k] = lo;
}

}  // namespace

namespace mfa {
int build_tc_device(mfa_model *m, bool layout_changed) {
  const int D = m->dim, G = m->num_gauss;
  mfa_engine *e = m->eng;
  const bool k80 = 2 * D <= 80 && !e->cfg.tc_k96;
  m->tc_ready = false;
  if (!k80 && 2 * D + 3 > 96) { m->tc_unsupported = true; return set_error(MFA_ERR_UNSUPPORTED, "tensor-core GMM kernel needs 2*dim+3 <= 96"); }
  const int TK = k80 ? 80 : 96, KC = TK / 8;
  const uint32_t TILE_BYTES = tile_bytes(TK);
  const bool geometry_changed = m->tc_k != TK;
  m->tc_k = TK;
  // dense tiling (all pdfs) by width class, from the host's copy of the layout
  std::vector<std::vector<int32_t>> by_cls(NCLS);
  for (int p = 0; p < m->num_pdfs; p++) {
    const int ng = m->h_pdf_off[p + 1] - m->h_pdf_off[p];
    if (ng <= 0) return set_error(MFA_ERR_INVALID, "pdf " + std::to_string(p) + " has no Gaussians");
    if (ng > TN) { m->tc_unsupported = true; return set_error(MFA_ERR_UNSUPPORTED, "pdf " + std::to_string(p) + " has more than 128 Gaussians"); }
    by_cls[cls_of(ng)].push_back(p);
  }
  std::vector<TcAux> aux;
  for (int c = 0; c < NCLS; c++)
    for (size_t i = 0; i < by_cls[c].size(); i += (size_t)cls_cap(c)) {
      TcAux a;
      memset(&a, 0, sizeof(a));
      a.cls = c; a.lp_base = -1;
      for (int j = 0; j < 32; j++) a.row[j] = -1;
      a.npdf = (int32_t)std::min<size_t>((size_t)cls_cap(c), by_cls[c].size() - i);
      for (int j = 0; j < a.npdf; j++) a.row[j] = by_cls[c][i + j];
      aux.push_back(a);
    }
  const int nt = (int)aux.size();
  m->tc_n_tiles = nt;
  cudaStream_t s = e->stream;
  if (!m->d_tc_colscale) CUDA_TRY(cudaMalloc((void **)&m->d_tc_colscale, 64 * sizeof(float)));
  if (!m->d_tc_flag) CUDA_TRY(cudaMalloc((void **)&m->d_tc_flag, 64 * sizeof(double) + 16));   // flag | m2[64]
  double *d_m2 = (double *)((uint8_t *)m->d_tc_flag + 16);
  CUDA_TRY(cudaMemsetAsync(m->d_tc_flag, 0, 64 * sizeof(double) + 16, s));
  // capacity-based: a changed layout alone never re-allocates (cudaMalloc / cudaFree have a long latency tail on this platform)
  const size_t img_total = (size_t)nt * TILE_BYTES, aux_bytes = (size_t)nt * sizeof(TcAux);
  if ((size_t)G > m->tc_cap_gauss || geometry_changed || !m->d_tc_rows) {
    CUDA_TRY(cudaStreamSynchronize(s));
    for (void **p : {&m->d_tc_rows, (void **)&m->d_tc_g}) if (*p) { CUDA_TRY(cudaFree(*p)); *p = nullptr; }
    m->tc_cap_gauss = (size_t)G + (size_t)G / 4 + 64;
  }
  if (img_total + aux_bytes > m->tc_w_cap || !m->d_tc_w) {
    CUDA_TRY(cudaStreamSynchronize(s));
    if (m->d_tc_w) { CUDA_TRY(cudaFree(m->d_tc_w)); m->d_tc_w = nullptr; }
    m->tc_w_cap = (img_total + aux_bytes) + (img_total + aux_bytes) / 4 + 4096;
  }
  if (!m->d_tc_rows) CUDA_TRY(cudaMalloc(&m->d_tc_rows, (size_t)2 * m->tc_cap_gauss * TK * sizeof(__half)));
  if (!m->d_tc_g) CUDA_TRY(cudaMalloc((void **)&m->d_tc_g, m->tc_cap_gauss * sizeof(float)));
  if (!m->d_tc_w) CUDA_TRY(cudaMalloc(&m->d_tc_w, m->tc_w_cap));
  m->tc_w_bytes = img_total;
  tc_moment_kernel<<<std::min(G, 4 * e->sm_count), 128, 0, s>>>(m->d_miv, m->d_iv, G, D, d_m2);
  tc_colscale_kernel<<<1, 64, 0, s>>>(d_m2, G, D, m->d_tc_colscale);
  const int64_t tot = (int64_t)G * TK;
  tc_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(m->d_miv, m->d_iv, m->d_gconsts, m->d_tc_colscale, G, D, TK, k80 ? 1 : 0,
                                                              (__half *)m->d_tc_rows, m->d_tc_g, m->d_tc_flag);
  e->launches += 3;
  CUDA_TRY(cudaGetLastError());
  TcAux *d_aux_stage;
  MFA_TRY(e->upload(DB_TC_ITEMS, aux.data(), aux.size(), &d_aux_stage));
  TcAux *d_aux = (TcAux *)((uint8_t *)m->d_tc_w + img_total);
  CUDA_TRY(cudaMemcpyAsync(d_aux, d_aux_stage, aux_bytes, cudaMemcpyDeviceToDevice, s));
  gather_b_kernel<<<(unsigned)(2 * nt), 256, 0, s>>>((const __half *)m->d_tc_rows, G, (uint8_t *)m->d_tc_w, nt, k80 ? -1 : 2 * D, KC, m->d_tc_g, d_aux,
                                                    nullptr, m->d_pdf_off);
  e->launches++;
  int32_t h_flag = 0;
  CUDA_TRY(cudaMemcpyAsync(&h_flag, m->d_tc_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (h_flag) { m->tc_unsupported = true; return set_error(MFA_ERR_UNSUPPORTED, "model weights exceed the fp16 range of the tensor-core kernel"); }
  m->tc_unsupported = false;
  // The per-utterance tile plan cached in mfa_graphs depends only on how many Gaussians each pdf has (and on the geometry), not on the
  // parameter values: key it on a hash of that layout, so that a re-estimated model with an unchanged layout (the usual case between
  // training iterations once pruning and mix-up have settled, and for the .alimdl / .mdl pair of a two-pass alignment) reuses the plan.
  uint64_t h = 1469598103934665603ull ^ (uint64_t)TK;
  for (int32_t v : m->h_pdf_off) { h ^= (uint64_t)(uint32_t)v; h *= 1099511628211ull; }
  m->tc_version = h | 1ull;
  m->tc_ready = true;
  return MFA_OK;
}
}  // namespace mfa

namespace {

template <int TKt, int NBt, bool GEPI, bool POLY>
static int launch_tc_t(mfa_engine *e, const TcParams &p, int grid) {
  const size_t smem = (2 + NBt) * (size_t)tile_bytes(TKt) + 2 * AUX_BYTES + 512;   // barriers + item ring
  CUDA_TRY(cudaFuncSetAttribute(gmm_tc_kernel<TKt, NBt, GEPI, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gmm_tc_kernel<TKt, NBt, GEPI, POLY><<<grid, NTHREADS, smem, e->stream>>>(p);
  return MFA_OK;
}

int launch_tc(mfa_engine *e, const TcParams &p_in, int tk) {
  TcParams p = p_in;
  const int grid = std::min(p.n_items, e->sm_count);
  MFA_TRY(e->getT<int>(DB_TC_CTR, 1, &p.counter));
  CUDA_TRY(cudaMemsetAsync(p.counter, 0, sizeof(int), e->stream));
#if MFA_TC_EXP == 7
  static long long *d_dbg = nullptr;
  if (!d_dbg) cudaMalloc((void **)&d_dbg, 1024 * 16 * sizeof(long long));
  cudaMemsetAsync(d_dbg, 0, 1024 * 16 * sizeof(long long), e->stream);
  p.dbg = d_dbg;
#endif
  const bool poly = e->cfg.tc_poly > 0;   // one in four exponentials on the FMA pipe
  if (tk == 80) MFA_TRY(poly ? (launch_tc_t<80, 3, true, true>(e, p, grid)) : (launch_tc_t<80, 3, true, false>(e, p, grid)));
  else MFA_TRY(poly ? (launch_tc_t<96, 2, false, true>(e, p, grid)) : (launch_tc_t<96, 2, false, false>(e, p, grid)));
  e->launches++;
  CUDA_TRY(cudaGetLastError());
#if MFA_TC_EXP == 7
  {
    cudaStreamSynchronize(e->stream);
    std::vector<long long> h(1024 * 16);
    cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double a[16] = {0};
    for (int b = 0; b < grid; b++) for (int k = 0; k < 16; k++) a[k] += (double)h[b * 16 + k] / grid;
    fprintf(stderr, "[tc-exp7] per CTA (cycles): producer wait empty_a %.0f empty_b %.0f | aux producer wait gempty %.0f | issuer wait full_a %.0f full_b %.0f "
                    "tempty %.0f, issue+commit %.0f, loop %.0f, accumulators %.0f | epilogue warp 4: wait tfull %.0f gfull %.0f body %.0f tiles %.0f\n",
            a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[12], a[8], a[9], a[10], a[11]);
  }
#endif
  return MFA_OK;
}

}  // namespace

namespace mfa {

bool gmm_tc_supported(mfa_model *m) {
  if (!m->tc_ready) { if (m->tc_unsupported || build_tc_device(m, true) != MFA_OK) return false; }
  return true;
}

// dense: every pdf for every row of `d_feats`; output pdf-major llT[pdf][ld]
int launch_gmm_tc(mfa_engine *e, mfa_model *m, const float *d_feats, int64_t n_rows, float *d_llT, int64_t ld) {
  if (n_rows == 0) return MFA_OK;
  if (ld < n_rows) return set_error(MFA_ERR_INVALID, "ld < n_rows");
  if (!m->tc_ready) {
    int r = m->tc_unsupported ? MFA_ERR_UNSUPPORTED : build_tc_device(m, true);
    if (r == MFA_ERR_UNSUPPORTED) return launch_gmm_ffma(e, m, d_feats, n_rows, d_llT, ld);  // shapes the tcgen05 kernel does not cover
    if (r) return r;
  }
  const int TK = m->tc_k, KC = TK / 8;
  const uint32_t TILE_BYTES = tile_bytes(TK);
  const int nt = m->tc_n_tiles;
  const int64_t n_ftiles = (n_rows + TM - 1) / TM, n_pairs = (n_ftiles + 1) / 2;
  uint8_t *d_a;
  MFA_TRY(e->getT<uint8_t>(DB_XSPLIT, (size_t)n_pairs * 2 * TILE_BYTES, &d_a));
  const int64_t total = n_pairs * 2 * KC * TM;
  xsplit_kernel<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(d_feats, m->dim, m->d_tc_colscale, d_a, n_pairs * 2, nullptr, nullptr, n_rows, KC,
                                                                         TK == 96);
  e->launches++;
  int splits = 1;
  if (n_pairs < 2 * (int64_t)e->sm_count) splits = (int)std::min<int64_t>(nt, (2 * (int64_t)e->sm_count + n_pairs - 1) / n_pairs);
  const int tps = (nt + splits - 1) / splits;
  splits = (nt + tps - 1) / tps;
  std::vector<TcItem> items;
  items.reserve((size_t)n_pairs * splits);
  for (int64_t pr = 0; pr < n_pairs; pr++)
    for (int sp = 0; sp < splits; sp++) {
      TcItem I{};
      I.a_tile = (uint32_t)(2 * pr); I.b_tile0 = (uint32_t)(sp * tps); I.n_b = (uint32_t)std::min(tps, nt - sp * tps);
      I.rows_valid = (uint32_t)std::min<int64_t>(2 * TM, ld - pr * 2 * TM); I.out_off = (uint64_t)(pr * 2 * TM); I.ld = (uint32_t)ld;
      items.push_back(I);
    }
  if ((int64_t)m->num_pdfs * ld > 0xFFFFFFFFLL) return set_error(MFA_ERR_UNSUPPORTED, "log-likelihood block exceeds 2^32 floats: chunk the frames");
  TcItem *d_items;
  MFA_TRY(e->upload(DB_TC_ITEMS, items.data(), items.size(), &d_items));
  TcParams p;
  p.a_img = d_a; p.b_img = (const uint8_t *)m->d_tc_w; p.aux = (const TcAux *)((const uint8_t *)m->d_tc_w + m->tc_w_bytes);
  p.items = d_items; p.n_items = (int)items.size(); p.out = d_llT;
  e->gmm_flops += 2.0 * (2 * m->dim + 1) * (double)m->num_gauss * (double)n_rows;
  e->gmm_issued += 3.0 * 2.0 * TK * (double)TN * (double)nt * (double)(2 * TM) * (double)n_pairs;
  return launch_tc(e, p, TK);
}

// ---- plan (cached per (graphs, model layout)), built on the device: per utterance, its pdfs by width class into 128-column tiles
static int ensure_rag_plan(mfa_engine *e, mfa_model *m, mfa_graphs *g) {
  if (!m->tc_ready) MFA_TRY(build_tc_device(m, true));
  if (g->rag_version == m->tc_version) return MFA_OK;
  if (g->lp_off[g->n_utts] > 0x7fffffffLL) return set_error(MFA_ERR_UNSUPPORTED, "more than 2^31 (utterance, pdf) pairs in one graph batch");
  for (int64_t k = 0; k < g->lp_off[g->n_utts]; k++)
    if (g->lp2pdf[k] < 0 || g->lp2pdf[k] >= m->num_pdfs) return set_error(MFA_ERR_INVALID, "graph references a pdf outside the model");
  const int nu = g->n_utts;
  int32_t *d_cnt; int32_t *h_cnt;
  MFA_TRY(e->getT<int32_t>(DB_RAG_CNT, (size_t)nu + 1, &d_cnt));
  { void *pp; MFA_TRY(e->get_pinned(PB_D, ((size_t)nu + 1) * sizeof(int32_t), &pp)); h_cnt = (int32_t *)pp; }
  rag_plan_kernel<<<(unsigned)((nu + 127) / 128), 128, 0, e->stream>>>(nu, g->d_lp_off, g->d_lp2pdf, m->d_pdf_off, 0, d_cnt, nullptr, nullptr);
  CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, (size_t)nu * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  g->rag_tile_off.assign((size_t)nu + 1, 0);
  for (int u = 0; u < nu; u++) g->rag_tile_off[u + 1] = g->rag_tile_off[u] + h_cnt[u];
  const size_t n_aux = (size_t)g->rag_tile_off[nu];
  if (g->d_rag && n_aux * sizeof(TcAux) > g->rag_meta_bytes) { dev_cache_give(g->device >= 0 ? g->device : e->device, g->d_rag, g->rag_meta_bytes); g->d_rag = nullptr; }
  if (!g->d_rag) {
    const size_t want = (n_aux + n_aux / 8 + 64) * sizeof(TcAux);
    g->d_rag = dev_cache_take(e->device, want, &g->rag_meta_bytes);
    if (g->d_rag) CUDA_TRY(cudaDeviceSynchronize());   // (a recycled block: its previous user may still have kernels in flight)
    else { g->rag_meta_bytes = want; CUDA_TRY(cudaMalloc(&g->d_rag, g->rag_meta_bytes)); }
  }
  int64_t *d_toff;
  MFA_TRY(e->upload(DB_TILE_ROW0, g->rag_tile_off.data(), g->rag_tile_off.size(), &d_toff));
  rag_plan_kernel<<<(unsigned)((nu + 127) / 128), 128, 0, e->stream>>>(nu, g->d_lp_off, g->d_lp2pdf, m->d_pdf_off, 1, nullptr, d_toff, (TcAux *)g->d_rag);
  e->launches += 2;
  CUDA_TRY(cudaGetLastError());
  g->rag_version = m->tc_version;
  e->pf_g = nullptr;
  return MFA_OK;
}

// B images (and the gconst column of the tiles' side data) of utterances [utt0, utt0 + n_utts) into DB_BIMG on stream `st`
static int gather_range(mfa_engine *e, mfa_model *m, mfa_graphs *g, int utt0, int n_utts, cudaStream_t st, uint8_t **d_b_out) {
  const int TK = m->tc_k, KC = TK / 8;
  const uint32_t TILE_BYTES = tile_bytes(TK);
  const int64_t bt0 = g->rag_tile_off[utt0], n_bt = g->rag_tile_off[utt0 + n_utts] - bt0;
  uint8_t *d_b;
  MFA_TRY(e->getT<uint8_t>(DB_BIMG, (size_t)std::max<int64_t>(n_bt, 1) * TILE_BYTES, &d_b));
  *d_b_out = d_b;
  if (n_bt == 0) return MFA_OK;
  gather_b_kernel<<<(unsigned)(2 * n_bt), 256, 0, st>>>((const __half *)m->d_tc_rows, m->num_gauss, d_b, n_bt, TK == 80 ? -1 : 2 * m->dim, KC, m->d_tc_g,
                                                       (TcAux *)g->d_rag + bt0, g->d_lp2pdf, m->d_pdf_off);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}

// The gather depends on the model and the graphs only, not on the audio: the fused pipeline starts it on a side stream before K1, so the
// ~1.4 ms (10 h) of HBM-bound image building hide behind the issue-bound MFCC kernel instead of sitting in front of the tensor-core kernel.
int prefetch_b_images(mfa_engine *e, mfa_model *m, mfa_graphs *g, int utt0, int n_utts) {
  if (n_utts <= 0) return MFA_OK;
  MFA_TRY(ensure_rag_plan(e, m, g));
  cudaStream_t st = e->sg;
  CUDA_TRY(cudaEventRecord(e->ev_fork, e->stream));      // everything queued so far (the previous call's readers of DB_BIMG) comes first
  CUDA_TRY(cudaStreamWaitEvent(st, e->ev_fork, 0));
  uint8_t *d_b;
  MFA_TRY(gather_range(e, m, g, utt0, n_utts, st, &d_b));
  CUDA_TRY(cudaEventRecord(e->ev_bimg, st));
  e->pf_g = g; e->pf_m = m; e->pf_u0 = utt0; e->pf_u1 = utt0 + n_utts;
  return MFA_OK;
}

// ragged: for each utterance only the pdfs its graph references (g->lp2pdf), output block per utterance [P_u][ld_u]
// at out + ll_off[u].  This is what Kaldi's decodable computes lazily; here it is ~13x less work than the dense matrix.
int launch_gmm_tc_ragged(mfa_engine *e, mfa_model *m, mfa_graphs *g, int utt0, int n_utts, const float *d_feats, const int64_t *h_row_off,
                         const int64_t *h_frame_off, float *d_out, const int64_t *h_ll_off, const int64_t *h_ld) {
  if (n_utts == 0) return MFA_OK;
  MFA_TRY(ensure_rag_plan(e, m, g));
  const int TK = m->tc_k, KC = TK / 8;
  const uint32_t TILE_BYTES = tile_bytes(TK);
  const int64_t bt0 = g->rag_tile_off[utt0], n_bt = g->rag_tile_off[utt0 + n_utts] - bt0;
  // ---- frame tiles follow utterance boundaries (pairs of 128 frames)
  std::vector<int64_t> tile_row0; std::vector<int32_t> tile_rows; std::vector<TcItem> items;
  for (int u = 0; u < n_utts; u++) {
    const int64_t T = h_frame_off[u + 1] - h_frame_off[u];
    const int64_t nb = g->rag_tile_off[utt0 + u + 1] - g->rag_tile_off[utt0 + u];
    const int64_t P_u = g->lp_off[utt0 + u + 1] - g->lp_off[utt0 + u];
    if (P_u * h_ld[u] > 0xFFFFFFFFLL) return set_error(MFA_ERR_UNSUPPORTED, "an utterance's log-likelihood block exceeds 2^32 floats");
    for (int64_t r0 = 0; r0 < T; r0 += 2 * TM) {
      TcItem I{};
      I.a_tile = (uint32_t)tile_row0.size();
      for (int h = 0; h < 2; h++) { tile_row0.push_back(h_row_off[u] + r0 + h * TM); tile_rows.push_back((int32_t)std::max<int64_t>(0, std::min<int64_t>(TM, T - r0 - h * TM))); }
      I.b_tile0 = (uint32_t)(g->rag_tile_off[utt0 + u] - bt0); I.n_b = (uint32_t)nb;
      I.rows_valid = (uint32_t)std::min<int64_t>(2 * TM, T - r0); I.out_off = (uint64_t)(h_ll_off[u] + r0); I.ld = (uint32_t)h_ld[u];
      if (nb > 0) items.push_back(I);
    }
    int64_t ng = 0;
    for (int64_t k = g->lp_off[utt0 + u]; k < g->lp_off[utt0 + u + 1]; k++) ng += m->h_pdf_off[g->lp2pdf[k] + 1] - m->h_pdf_off[g->lp2pdf[k]];
    e->gmm_flops += 2.0 * (2 * m->dim + 1) * (double)ng * (double)T;
    e->gmm_issued += 3.0 * 2.0 * TK * (double)TN * (double)nb * (double)(2 * TM) * (double)((T + 2 * TM - 1) / (2 * TM));
  }
  if (items.empty()) return MFA_OK;
  // longest items first (static round-robin over CTAs then balances well)
  std::stable_sort(items.begin(), items.end(), [](const TcItem &a, const TcItem &b) { return a.n_b > b.n_b; });
  const int64_t n_at = (int64_t)tile_row0.size();
  uint8_t *d_a, *d_b; int64_t *d_row0; int32_t *d_rows; TcItem *d_items;
  MFA_TRY(e->getT<uint8_t>(DB_XSPLIT, (size_t)n_at * TILE_BYTES, &d_a));
  MFA_TRY(e->upload(DB_TILE_ROW0, tile_row0.data(), tile_row0.size(), &d_row0));
  MFA_TRY(e->upload(DB_TILE_ROWS, tile_rows.data(), tile_rows.size(), &d_rows));
  MFA_TRY(e->upload(DB_TC_ITEMS, items.data(), items.size(), &d_items));
  const int64_t tot_a = n_at * KC * TM;
  xsplit_kernel<<<(unsigned)((tot_a + 255) / 256), 256, 0, e->stream>>>(d_feats, m->dim, m->d_tc_colscale, d_a, n_at, d_row0, d_rows, 0, KC, TK == 96);
  e->launches++;
  if (e->pf_g == g && e->pf_m == m && utt0 >= e->pf_u0 && utt0 + n_utts <= e->pf_u1) {
    // images gathered ahead on the side stream: this launch covers a sub-range of them
    CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_bimg, 0));
    d_b = (uint8_t *)e->dev[DB_BIMG].p + (size_t)(bt0 - g->rag_tile_off[e->pf_u0]) * TILE_BYTES;
    if (utt0 + n_utts == e->pf_u1) e->pf_g = nullptr;   // consumed
  } else {
    e->pf_g = nullptr;
    MFA_TRY(gather_range(e, m, g, utt0, n_utts, e->stream, &d_b));
  }
  TcAux *d_aux = (TcAux *)g->d_rag + bt0;
  TcParams p;
  p.a_img = d_a; p.b_img = d_b; p.aux = d_aux; p.items = d_items; p.n_items = (int)items.size(); p.out = d_out;
  return launch_tc(e, p, TK);
}

}  // namespace mfa
