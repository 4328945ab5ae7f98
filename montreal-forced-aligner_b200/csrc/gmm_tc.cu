// gmm_tc.cu -- K2 on the 5th-generation tensor cores: all-pdf diagonal-GMM log-likelihoods as the dense contraction
//   C[t, m] = [x s, (x s)^2, 1, 1, 1] . [mu/(sigma^2 s), -1/(2 sigma^2 s^2), g1, g2, g3]^T * log2(e)       (m = Gaussian)
// issued as tcgen05.mma (kind::f16, fp32 accumulators in TMEM), followed by a per-pdf log-sum-exp computed by the
// epilogue warps straight out of TMEM (tcgen05.ld), one frame per thread.
//
// Replaces DecodableAmDiagGmmScaled::LogLikelihoodZeroBased / gmm_compute_likes (reference call sites:
// montreal_forced_aligner/alignment/multiprocessing.py:846 (inside GmmAligner), :1415); semantics SURVEY.md A.4.
//
// Split precision: both operands are split into fp16 hi + lo parts (22 significand bits) and three products are
// accumulated (hi*hi, hi*lo, lo*hi); features are pre-scaled per dimension by a power of two s_d (folded into the
// weights) so x s and (x s)^2 sit well inside fp16's range; the gconst enters as three fp16 columns against ones.
// K = 2D + 3 padded to 96 -> 6 k-steps x 3 products = 18 MMAs (M=128, N=128, K=16) per 128x128 output tile.
//
// Data movement: operands live in global memory ALREADY in the UMMA canonical K-major no-swizzle layout
// ([k/8][row/8][8 rows][8 halves], core matrix = 128 contiguous bytes), so a tile is one contiguous 48 KB image and
// is fetched with a single 1-D bulk copy (cp.async.bulk -> UBLKCP) completing on an mbarrier; no tensor maps.
// Each CTA keeps the A images of two frame tiles (256 frames) resident and streams the Gaussian tiles through a
// two-stage ring, so every B image fetched from L2 feeds 2 x 18 MMAs.  TMEM holds 2 stages x 2 accumulators of 128
// columns (all 512 columns): the MMA warp runs one Gaussian tile ahead of the two epilogue warpgroups.
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
// Two geometries.  K = 80 (2 dim <= 80; MFA's 39 / 40-dimensional features): the gconst is added by the epilogue from a per-tile fp32
// array, 5 k-steps x 3 products = 15 MMAs per tile, 40 KB tiles, a THREE-stage B ring.  K = 96 (2 dim + 3 <= 96, or engine option tc_k96): the
// gconst rides as three fp16 columns against ones, 18 MMAs per tile, 48 KB tiles, two-stage ring (the first version of this kernel).
__host__ __device__ constexpr uint32_t img_bytes(int tk) { return (uint32_t)(TM * tk * 2); }   // one fp16 image (hi or lo) of a 128 x tk tile
__host__ __device__ constexpr uint32_t tile_bytes(int tk) { return 2 * img_bytes(tk); }         // hi + lo
constexpr uint32_t LBO_BYTES = (TM / 8) * 128;    // K-adjacent core matrices
constexpr uint32_t SBO_BYTES = 128;               // row-group-adjacent core matrices
constexpr int NTHREADS = 640;   // 4 control warps + 16 epilogue warps (4 per SM sub-partition)
constexpr float kLn2 = 0.69314718055994530942f, kLog2e = 1.44269504088896340736f;

struct TcMeta {  // per Gaussian tile: pdf structure of its 128 columns at 4-column group granularity (32 groups)
  uint32_t gstart, gend;  // bit g: group g starts a pdf / is the last group of a pdf (pdf column ranges are multiples of 4)
  int32_t pdf0;           // first output row of the tile: global pdf id (dense tiling) or utterance-local pdf index (ragged)
  int32_t lp0;            // ragged tiles: index into the graphs' lp2pdf list of the tile's first pdf (its pdfs are lp0 .. lp0 + popc(gend) - 1)
};
static_assert(sizeof(TcMeta) == 16, "TcMeta must be 16 bytes");

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

// One 32-column chunk (8 groups of 4 columns) of a frame's component scores (log2 domain).  pdf boundaries fall on group
// boundaries, so the segmented max / sum scans run over 8 group values: group max (tree) -> forward running max -> backward
// broadcast of each pdf's max -> exp2 of the 32 values against their pdf's max -> group sums -> forward running sum; one
// log2 + store per finished pdf.  (cmx, cs) carry an unfinished pdf into the next chunk.  All predicates are warp-uniform.
template <bool GEPI>
__device__ __forceinline__ void lse_chunk(uint32_t (&vr)[32], const float *gp, uint32_t gs, uint32_t ge, float &cmx, float &cs, float *&out,
                                          int64_t ld, bool row_ok) {
  if (GEPI) {   // gconst of the chunk's 32 columns from shared memory (every thread reads the same words: broadcasts)
#pragma unroll
    for (int g = 0; g < 8; g++) {
      const float4 q = *reinterpret_cast<const float4 *>(gp + 4 * g);
      vr[4 * g] = __float_as_uint(__uint_as_float(vr[4 * g]) + q.x); vr[4 * g + 1] = __float_as_uint(__uint_as_float(vr[4 * g + 1]) + q.y);
      vr[4 * g + 2] = __float_as_uint(__uint_as_float(vr[4 * g + 2]) + q.z); vr[4 * g + 3] = __float_as_uint(__uint_as_float(vr[4 * g + 3]) + q.w);
    }
  }
  float r[8];
  float run = cmx;
#pragma unroll
  for (int g = 0; g < 8; g++) {
    const float gm = fmaxf(fmaxf(__uint_as_float(vr[4 * g]), __uint_as_float(vr[4 * g + 1])),
                           fmaxf(__uint_as_float(vr[4 * g + 2]), __uint_as_float(vr[4 * g + 3])));
    run = ((gs >> g) & 1u) ? gm : fmaxf(run, gm);
    r[g] = run;
  }
  float m = r[7];
#pragma unroll
  for (int g = 7; g >= 0; g--) {
    m = ((ge >> g) & 1u) ? r[g] : m;
    r[g] = m;
  }
  float q = cs * ex2(cmx - r[0]);
#pragma unroll
  for (int g = 0; g < 8; g++) {
    const float e0 = ex2(__uint_as_float(vr[4 * g]) - r[g]), e1 = ex2(__uint_as_float(vr[4 * g + 1]) - r[g]);
    const float e2 = ex2(__uint_as_float(vr[4 * g + 2]) - r[g]), e3 = ex2(__uint_as_float(vr[4 * g + 3]) - r[g]);
    const float s4 = (e0 + e1) + (e2 + e3);
    q = ((gs >> g) & 1u) ? s4 : q + s4;
    if ((ge >> g) & 1u) {
      if (row_ok) *out = (r[g] + lg2(q)) * kLn2;
      out += ld;
    }
  }
  cmx = r[7];
  cs = q;
}

// One work item = one pair of frame tiles (256 frames) against a run of Gaussian tiles.
struct TcItem {
  uint32_t a_tile;       // first of the two A images of the pair
  uint32_t b_tile0, n_b; // Gaussian tiles [b_tile0, b_tile0 + n_b): images in b_img, segment masks in meta
  uint32_t rows_valid;   // frames of the pair that exist (<= 256)
  uint64_t out_off;      // float offset of the pair's first frame in `out`
  uint32_t ld, pad;      // leading dimension (frames) of this item's output block
};
static_assert(sizeof(TcItem) == 32, "TcItem must be 32 bytes");

struct TcParams {
  const uint8_t *a_img;   // [n_frame_tiles (even)][tile_bytes]
  const uint8_t *b_img;   // [n_b_tiles][tile_bytes]
  const TcMeta *meta;     // [n_b_tiles]; pdf0 = first output row of the tile (global pdf id, or utterance-local pdf index)
  const float *g_tiles;   // K = 80 geometry: [n_b_tiles][128] gconst * log2(e) per tile column (padding: -60000)
  const TcItem *items;    // [n_items]
  int n_items;
  float *out;             // pdf-major blocks: out[item.out_off + (meta.pdf0 + k) * item.ld + frame]
};

template <int TKt, int NBt, bool GEPI>
__global__ void __launch_bounds__(NTHREADS, 1)
gmm_tc_kernel(TcParams p) {
  constexpr uint32_t IMG_BYTES = img_bytes(TKt), TILE_BYTES = tile_bytes(TKt);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sA = smem;                       // 2 frame tiles x (hi, lo)
  uint8_t *sB = smem + 2 * TILE_BYTES;      // NBt stages x (hi, lo)
  float *sG = (float *)(smem + (2 + NBt) * TILE_BYTES);   // 2 x 128 gconsts (one row per accumulator stage)
  uint64_t *bars = (uint64_t *)(smem + (2 + NBt) * TILE_BYTES + 1024);
  uint64_t *full_a = bars + 0, *empty_a = bars + 1, *full_b = bars + 2, *empty_b = bars + 2 + NBt, *tfull = bars + 2 + 2 * NBt,
           *tempty = tfull + 4, *gfull = tempty + 4, *gempty = gfull + 2;
  uint32_t *tmem_slot = (uint32_t *)(gempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(full_a, 1); mbar_init(empty_a, 1);
    for (int s = 0; s < NBt; s++) { mbar_init(full_b + s, 1); mbar_init(empty_b + s, 1); }
    for (int i = 0; i < 4; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 128); }
    for (int i = 0; i < 2; i++) { mbar_init(gfull + i, 1); mbar_init(gempty + i, 256); }
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
      uint32_t cnt = 0, it = 0, sb = 0, phb = 0;   // sb / phb: slot of the B ring and the phase of its barriers
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
        const TcItem I = p.items[item];
        mbar_wait(empty_a, (it & 1) ^ 1);
        mbar_expect_tx(full_a, 2 * TILE_BYTES);
        bulk_g2s(sA, p.a_img + (size_t)I.a_tile * TILE_BYTES, 2 * TILE_BYTES, full_a);
        for (uint32_t n = I.b_tile0; n < I.b_tile0 + I.n_b; n++, cnt++) {
          mbar_wait(empty_b + sb, phb ^ 1);
          mbar_expect_tx(full_b + sb, TILE_BYTES);
          bulk_g2s(sB + sb * TILE_BYTES, p.b_img + (size_t)n * TILE_BYTES, TILE_BYTES, full_b + sb);
          if (++sb == NBt) { sb = 0; phb ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ===== gconst producer (K = 80 geometry): the tile's 128 gconsts travel with the accumulator stage (cnt & 1) and are released by
    // the epilogue warps -- its own thread, so a slow epilogue never delays the B ring =====
    if (GEPI && lane == 0) {
      uint32_t cnt = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const TcItem I = p.items[item];
        for (uint32_t n = I.b_tile0; n < I.b_tile0 + I.n_b; n++, cnt++) {
          const uint32_t sg = cnt & 1;
          mbar_wait(gempty + sg, ((cnt >> 1) & 1) ^ 1);
          mbar_expect_tx(gfull + sg, TN * 4);
          bulk_g2s(sG + sg * TN, p.g_tiles + (size_t)n * TN, TN * 4, gfull + sg);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      // InstrDescriptor: c_format=F32 (1<<4), a/b format F16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      uint32_t cnt = 0, it = 0, sb = 0, phb = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
        const uint32_t n_b = p.items[item].n_b, rows_valid = p.items[item].rows_valid;
        mbar_wait(full_a, it & 1);
        tc_fence_after();
        for (uint32_t n = 0; n < n_b; n++, cnt++) {
          const uint32_t s = cnt & 1, ph = (cnt >> 1) & 1;
          mbar_wait(full_b + sb, phb);
          tc_fence_after();
#pragma unroll
          for (int f = 0; f < 2; f++) {
            mbar_wait(tempty + s * 2 + f, ph ^ 1);
            tc_fence_after();
            if (f == 0 || rows_valid > TM) {   // a pair whose second tile holds no frames skips its 18 MMAs
              const uint32_t d = tmem_base + s * 256 + f * 128;
              const uint32_t a0 = a_base + f * TILE_BYTES, b0 = b_base + sb * TILE_BYTES;
#pragma unroll
              for (int prod = 0; prod < 3; prod++) {
                const uint32_t ao = a0 + (prod == 2 ? IMG_BYTES : 0), bo = b0 + (prod == 1 ? IMG_BYTES : 0);
#pragma unroll
                for (int k = 0; k < TKt / 16; k++)
                  umma_f16(d, make_desc(ao + k * 2 * LBO_BYTES), make_desc(bo + k * 2 * LBO_BYTES), idesc, (prod | k) != 0);
              }
            }
            umma_commit(tfull + s * 2 + f);
          }
          umma_commit(empty_b + sb);
          if (++sb == NBt) { sb = 0; phb ^= 1; }
        }
        umma_commit(empty_a);
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: 16 warps = 2 accumulator stages x 2 frame tiles x 4 lane quarters; thread = one frame (TMEM lane).
    // The warps of stage s take every other Gaussian tile, so four epilogue warps share each SM sub-partition and the
    // TMEM-load / MUFU latency of one is covered by the others. =====
    const int e = warp - 4, wq = e & 3, f = (e >> 2) & 1, s = e >> 3;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    uint32_t cnt = 0;   // tiles seen by the CTA so far (all roles count alike); this warp serves those with (cnt & 1) == s
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const TcItem I = p.items[item];
      const uint32_t row = f * TM + wq * 32 + lane;
      const bool row_ok = row < I.rows_valid;
      const bool tile_live = (uint32_t)(f * TM) < I.rows_valid;
      float *out_base = p.out + I.out_off + row;
      for (uint32_t n = I.b_tile0; n < I.b_tile0 + I.n_b; n++, cnt++) {
        if ((cnt & 1u) != (uint32_t)s) continue;
        const uint32_t ph = (cnt >> 1) & 1;
        const TcMeta cur = p.meta[n];
        mbar_wait(tfull + s * 2 + f, ph);
        tc_fence_after();
        if (GEPI) mbar_wait(gfull + s, ph);
        if (!tile_live) { tc_fence_before(); mbar_arrive(tempty + s * 2 + f); if (GEPI) mbar_arrive(gempty + s); continue; }
        const float *gs = sG + s * TN;
        const uint32_t t0 = tmem_base + lane_base + s * 256 + f * 128;
        float cmx = -INFINITY, cs = 0.0f;
        float *out = out_base + (size_t)cur.pdf0 * I.ld;
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          tmem_ld32(t0 + 32 * c, v);
          tmem_ld_wait();
          if (c == 3) { tc_fence_before(); mbar_arrive(tempty + s * 2 + f); }   // accumulator drained: the MMA warp may overwrite it
          lse_chunk<GEPI>(v, gs + 32 * c, (cur.gstart >> (8 * c)) & 0xFF, (cur.gend >> (8 * c)) & 0xFF, cmx, cs, out, I.ld, row_ok);
        }
        if (GEPI) mbar_arrive(gempty + s);   // this thread is done with the stage's gconsts
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

// B images for utterance-specific Gaussian tiles: row r of tile t is Gaussian row_src[t*128 + r] of the model (row-major fp16
// hi/lo weight rows), or padding (zero weights; gconst -60000 -> exp2 -> 0) when row_src < 0.  One CTA per (tile, hi | lo): the 128
// source rows are read as whole rows (consecutive threads = consecutive 16-byte pieces of a row: full sectors), re-ordered into the
// canonical [k/8][row/8][8 rows][8 halves] image in shared memory and written out linearly.
// gcol >= 0: K = 96 geometry (padding rows carry -60000 in the first gconst column); gcol < 0: K = 80 geometry, the per-tile gconst array
// g_out[tile][128] is gathered here as well.
// row_src == nullptr (the per-utterance tiles of the ragged path): the source rows follow from the tile's own pdf list -- pdfs
// lp2pdf[meta.lp0 ...], each occupying its Gaussian count rounded up to 4 columns -- so no per-column table exists in memory at all.
__global__ void __launch_bounds__(256)
gather_b_kernel(const __half *__restrict__ w_rows, int64_t num_gauss, const int32_t *__restrict__ row_src, uint8_t *__restrict__ b_img,
                int64_t n_tiles, int gcol, int KC, const float *__restrict__ g_src, float *__restrict__ g_out,
                const TcMeta *__restrict__ meta, const int32_t *__restrict__ lp2pdf, const int32_t *__restrict__ pdf_off) {
  __shared__ int s_src[TN];
  __shared__ int s_c0[33], s_g0[32];
  __shared__ uint4 s_img[(TN + 1) * 12];   // up to K = 96; one padding unit per k-chunk keeps the transposing stores conflict-free
  const int TK = KC * 8;
  const uint32_t IMG_BYTES = img_bytes(TK), TILE_BYTES = 2 * IMG_BYTES;
  const int64_t tile = blockIdx.x >> 1;
  const int which = blockIdx.x & 1, t = threadIdx.x;
  if (tile >= n_tiles) return;
  if (row_src) {
    if (t < TN) s_src[t] = row_src[tile * TN + t];
  } else {
    const TcMeta mt = meta[tile];
    const int npdf = __popc(mt.gend);
    if (t < 32) {   // warp 0: column start of each pdf of the tile = exclusive prefix of the padded component counts
      int ng = 0, g0 = 0;
      if (t < npdf) { const int pdf = lp2pdf[mt.lp0 + t]; g0 = pdf_off[pdf]; ng = pdf_off[pdf + 1] - g0; }
      int pad = (ng + MFA_SEG_ALIGN - 1) / MFA_SEG_ALIGN * MFA_SEG_ALIGN, inc = pad;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (t >= o) inc += v; }
      s_c0[t] = inc - pad; s_g0[t] = g0 | (ng << 24);   // ng <= 128 fits the top byte; g0 < 2^24 Gaussians
      if (t == 31) s_c0[32] = inc;
    }
    __syncthreads();
    if (t < TN) {
      int g = -1;
      for (int i = 0; i < npdf; i++) {
        const int c = t - s_c0[i], ng = s_g0[i] >> 24;
        if (c >= 0 && c < ng) { g = (s_g0[i] & 0xFFFFFF) + c; break; }
      }
      s_src[t] = g;
    }
  }
  __syncthreads();
  if (t < TN && g_out && which == 0) { const int g = s_src[t]; g_out[tile * TN + t] = g >= 0 ? g_src[g] : -60000.0f; }
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

// Per-utterance tile plan of the ragged path, on the device (it is rebuilt whenever an M-step changed some pdf's component count):
// one thread per utterance walks its pdf list (sorted pdf ids, graphs' lp2pdf) and packs the padded component counts into 128-column
// tiles that never split a pdf.  fill == 0: tile count per utterance; fill == 1: the tiles' TcMeta at tile_off[u].
__global__ void rag_plan_kernel(int n_utts, const int64_t *__restrict__ lp_off, const int32_t *__restrict__ lp2pdf, const int32_t *__restrict__ pdf_off,
                                int fill, int32_t *__restrict__ n_tiles_out, const int64_t *__restrict__ tile_off, TcMeta *__restrict__ meta) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_utts) return;
  const int64_t k0 = lp_off[u], k1 = lp_off[u + 1];
  int col = TN;   // forces a new tile at the first pdf
  int64_t t = (fill ? tile_off[u] : 0) - 1;
  TcMeta cur{};
  for (int64_t k = k0; k < k1; k++) {
    const int pdf = lp2pdf[k];
    const int ng = pdf_off[pdf + 1] - pdf_off[pdf], pad = (ng + MFA_SEG_ALIGN - 1) / MFA_SEG_ALIGN * MFA_SEG_ALIGN;
    if (col + pad > TN) {
      if (fill && k > k0) { if (col < TN) cur.gstart |= 1u << (col / 4); meta[t] = cur; }   // trailing padding: a junk segment that never ends
      t++;
      cur.gstart = 0; cur.gend = 0; cur.pdf0 = (int32_t)(k - k0); cur.lp0 = (int32_t)k;
      col = 0;
    }
    cur.gstart |= 1u << (col / 4);
    cur.gend |= 1u << ((col + pad - 1) / 4);
    col += pad;
  }
  if (fill && k1 > k0) { if (col < TN) cur.gstart |= 1u << (col / 4); meta[t] = cur; }
  if (!fill) n_tiles_out[u] = (int)(t + 1);
}

// the dense per-tile gconst array follows the per-Gaussian one; bulk copies need a 16-byte aligned source
static inline size_t tc_gpad(int64_t G) { return (size_t)((G + 3) & ~(int64_t)3); }

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
  const int nt = m->n_tiles;
  cudaStream_t s = e->stream;
  if (!m->d_tc_colscale) CUDA_TRY(cudaMalloc((void **)&m->d_tc_colscale, 64 * sizeof(float)));
  if (!m->d_tc_flag) CUDA_TRY(cudaMalloc((void **)&m->d_tc_flag, 64 * sizeof(double) + 16));   // flag | m2[64]
  double *d_m2 = (double *)((uint8_t *)m->d_tc_flag + 16);
  CUDA_TRY(cudaMemsetAsync(m->d_tc_flag, 0, 64 * sizeof(double) + 16, s));
  if ((size_t)G > m->tc_cap_gauss || geometry_changed || !m->d_tc_rows || (layout_changed && m->d_tc_w)) {
    CUDA_TRY(cudaStreamSynchronize(s));
    for (void **p : {&m->d_tc_rows, (void **)&m->d_tc_g, &m->d_tc_w}) if (*p) { CUDA_TRY(cudaFree(*p)); *p = nullptr; }
    m->tc_cap_gauss = (size_t)G + (size_t)G / 8 + 64;
  }
  const size_t img_total = (size_t)nt * TILE_BYTES, meta_bytes = (size_t)nt * sizeof(TcMeta);
  const bool new_dense = m->d_tc_w == nullptr;
  if (!m->d_tc_rows) CUDA_TRY(cudaMalloc(&m->d_tc_rows, (size_t)2 * m->tc_cap_gauss * TK * sizeof(__half)));
  // per Gaussian | dense per-tile array: the tile count can only be bounded by the number of pdfs
  if (!m->d_tc_g) CUDA_TRY(cudaMalloc((void **)&m->d_tc_g, (tc_gpad((int64_t)m->tc_cap_gauss) + (size_t)m->num_pdfs * TN + TN) * sizeof(float)));
  if (new_dense) CUDA_TRY(cudaMalloc(&m->d_tc_w, img_total + meta_bytes));
  m->tc_w_bytes = img_total;
  tc_moment_kernel<<<std::min(G, 4 * e->sm_count), 128, 0, s>>>(m->d_miv, m->d_iv, G, D, d_m2);
  tc_colscale_kernel<<<1, 64, 0, s>>>(d_m2, G, D, m->d_tc_colscale);
  const int64_t tot = (int64_t)G * TK;
  tc_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(m->d_miv, m->d_iv, m->d_gconsts, m->d_tc_colscale, G, D, TK, k80 ? 1 : 0,
                                                              (__half *)m->d_tc_rows, m->d_tc_g, m->d_tc_flag);
  e->launches += 3;
  CUDA_TRY(cudaGetLastError());
  // dense tiling (all pdfs): per-tile masks + source rows from the host's copy of the layout; the images are gathered on the device
  std::vector<int32_t> row_src((size_t)nt * TN, -1);
  std::vector<TcMeta> meta(nt);
  for (int tl = 0; tl < nt; tl++) {
    int col = 0;
    memset(&meta[tl], 0, sizeof(TcMeta));
    meta[tl].pdf0 = m->h_tile_pdf0[tl];
    for (int pdf = m->h_tile_pdf0[tl]; pdf < m->h_tile_pdf0[tl + 1]; pdf++) {
      const int ng = m->h_pdf_off[pdf + 1] - m->h_pdf_off[pdf], pad = (ng + MFA_SEG_ALIGN - 1) / MFA_SEG_ALIGN * MFA_SEG_ALIGN;
      meta[tl].gstart |= 1u << (col / 4);
      meta[tl].gend |= 1u << ((col + pad - 1) / 4);
      for (int k = 0; k < ng; k++) row_src[(size_t)tl * TN + col + k] = m->h_pdf_off[pdf] + k;
      col += pad;
    }
    if (col < TN) meta[tl].gstart |= 1u << (col / 4);  // trailing padding: one junk segment that never ends
  }
  int32_t *d_src; TcMeta *d_meta_stage;
  MFA_TRY(e->upload(DB_SCRATCH, row_src.data(), row_src.size(), &d_src));
  MFA_TRY(e->upload(DB_TC_ITEMS, meta.data(), meta.size(), &d_meta_stage));
  CUDA_TRY(cudaMemcpyAsync((uint8_t *)m->d_tc_w + img_total, d_meta_stage, meta_bytes, cudaMemcpyDeviceToDevice, s));
  gather_b_kernel<<<(unsigned)(2 * nt), 256, 0, s>>>((const __half *)m->d_tc_rows, G, d_src, (uint8_t *)m->d_tc_w, nt, k80 ? -1 : 2 * D,
                                                    KC, m->d_tc_g, k80 ? m->d_tc_g + tc_gpad((int64_t)m->tc_cap_gauss) : nullptr, nullptr, nullptr, nullptr);
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

int launch_tc(mfa_engine *e, const TcParams &p, int tk) {
  const int grid = std::min(p.n_items, e->sm_count);
  if (tk == 80) {
    const size_t smem = 5 * (size_t)tile_bytes(80) + 1024 + 256;
    CUDA_TRY(cudaFuncSetAttribute(gmm_tc_kernel<80, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gmm_tc_kernel<80, 3, true><<<grid, NTHREADS, smem, e->stream>>>(p);
  } else {
    const size_t smem = 4 * (size_t)tile_bytes(96) + 1024 + 256;
    CUDA_TRY(cudaFuncSetAttribute(gmm_tc_kernel<96, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gmm_tc_kernel<96, 2, false><<<grid, NTHREADS, smem, e->stream>>>(p);
  }
  e->launches++;
  CUDA_TRY(cudaGetLastError());
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
  const int64_t n_ftiles = (n_rows + TM - 1) / TM, n_pairs = (n_ftiles + 1) / 2;
  uint8_t *d_a;
  MFA_TRY(e->getT<uint8_t>(DB_XSPLIT, (size_t)n_pairs * 2 * TILE_BYTES, &d_a));
  const int64_t total = n_pairs * 2 * KC * TM;
  xsplit_kernel<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(d_feats, m->dim, m->d_tc_colscale, d_a, n_pairs * 2, nullptr, nullptr, n_rows, KC,
                                                                         TK == 96);
  e->launches++;
  int splits = 1;
  if (n_pairs < 2 * (int64_t)e->sm_count) splits = (int)std::min<int64_t>(m->n_tiles, (2 * (int64_t)e->sm_count + n_pairs - 1) / n_pairs);
  const int tps = (m->n_tiles + splits - 1) / splits;
  splits = (m->n_tiles + tps - 1) / tps;
  std::vector<TcItem> items;
  items.reserve((size_t)n_pairs * splits);
  for (int64_t pr = 0; pr < n_pairs; pr++)
    for (int sp = 0; sp < splits; sp++) {
      TcItem I{};
      I.a_tile = (uint32_t)(2 * pr); I.b_tile0 = (uint32_t)(sp * tps); I.n_b = (uint32_t)std::min(tps, m->n_tiles - sp * tps);
      I.rows_valid = (uint32_t)std::min<int64_t>(2 * TM, ld - pr * 2 * TM); I.out_off = (uint64_t)(pr * 2 * TM); I.ld = (uint32_t)ld;
      items.push_back(I);
    }
  if (ld > 0xFFFFFFFFLL) return set_error(MFA_ERR_UNSUPPORTED, "leading dimension exceeds 2^32 frames");
  TcItem *d_items;
  MFA_TRY(e->upload(DB_TC_ITEMS, items.data(), items.size(), &d_items));
  TcParams p;
  p.a_img = d_a; p.b_img = (const uint8_t *)m->d_tc_w; p.meta = (const TcMeta *)((const uint8_t *)m->d_tc_w + m->tc_w_bytes);
  p.items = d_items; p.n_items = (int)items.size(); p.out = d_llT;
  p.g_tiles = m->d_tc_g + tc_gpad((int64_t)m->tc_cap_gauss);
  e->gmm_flops += 2.0 * (2 * m->dim + 1) * (double)m->num_gauss * (double)n_rows;
  return launch_tc(e, p, TK);
}

// ragged: for each utterance only the pdfs its graph references (g->lp2pdf), output block per utterance [P_u][ld_u]
// at out + ll_off[u].  This is what Kaldi's decodable computes lazily; here it is ~13x less work than the dense matrix.
int launch_gmm_tc_ragged(mfa_engine *e, mfa_model *m, mfa_graphs *g, int utt0, int n_utts, const float *d_feats, const int64_t *h_row_off,
                         const int64_t *h_frame_off, float *d_out, const int64_t *h_ll_off, const int64_t *h_ld) {
  if (n_utts == 0) return MFA_OK;
  if (!m->tc_ready) MFA_TRY(build_tc_device(m, true));
  const int TK = m->tc_k, KC = TK / 8;
  const uint32_t TILE_BYTES = tile_bytes(TK);
  // ---- plan (cached per (graphs, model tiling)): per utterance, pack its local pdfs into 128-column tiles
  if (g->rag_version != m->tc_version) {
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
    const size_t n_meta = (size_t)g->rag_tile_off[nu];
    if (g->d_rag && n_meta * sizeof(TcMeta) > g->rag_meta_bytes) { CUDA_TRY(cudaFree(g->d_rag)); g->d_rag = nullptr; }
    if (!g->d_rag) {
      g->rag_meta_bytes = (n_meta + n_meta / 8 + 64) * sizeof(TcMeta);
      CUDA_TRY(cudaMalloc(&g->d_rag, g->rag_meta_bytes));
    }
    int64_t *d_toff;
    MFA_TRY(e->upload(DB_TILE_ROW0, g->rag_tile_off.data(), g->rag_tile_off.size(), &d_toff));
    rag_plan_kernel<<<(unsigned)((nu + 127) / 128), 128, 0, e->stream>>>(nu, g->d_lp_off, g->d_lp2pdf, m->d_pdf_off, 1, nullptr, d_toff, (TcMeta *)g->d_rag);
    e->launches += 2;
    CUDA_TRY(cudaGetLastError());
    g->rag_version = m->tc_version;
  }
  const int64_t bt0 = g->rag_tile_off[utt0], n_bt = g->rag_tile_off[utt0 + n_utts] - bt0;
  // ---- frame tiles follow utterance boundaries (pairs of 128 frames)
  std::vector<int64_t> tile_row0; std::vector<int32_t> tile_rows; std::vector<TcItem> items;
  for (int u = 0; u < n_utts; u++) {
    const int64_t T = h_frame_off[u + 1] - h_frame_off[u];
    const int64_t nb = g->rag_tile_off[utt0 + u + 1] - g->rag_tile_off[utt0 + u];
    if (h_ld[u] > 0xFFFFFFFFLL) return set_error(MFA_ERR_UNSUPPORTED, "utterance longer than 2^32 frames");
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
  }
  if (items.empty()) return MFA_OK;
  // longest items first (static round-robin over CTAs then balances well)
  std::stable_sort(items.begin(), items.end(), [](const TcItem &a, const TcItem &b) { return a.n_b > b.n_b; });
  const int64_t n_at = (int64_t)tile_row0.size();
  uint8_t *d_a, *d_b; int64_t *d_row0; int32_t *d_rows; TcItem *d_items;
  MFA_TRY(e->getT<uint8_t>(DB_XSPLIT, (size_t)n_at * TILE_BYTES, &d_a));
  MFA_TRY(e->getT<uint8_t>(DB_BIMG, (size_t)n_bt * TILE_BYTES + (size_t)n_bt * TN * sizeof(float), &d_b));
  float *d_gt = (float *)(d_b + (size_t)n_bt * TILE_BYTES);   // per-tile gconsts behind the images (TILE_BYTES is a multiple of 16)
  MFA_TRY(e->upload(DB_TILE_ROW0, tile_row0.data(), tile_row0.size(), &d_row0));
  MFA_TRY(e->upload(DB_TILE_ROWS, tile_rows.data(), tile_rows.size(), &d_rows));
  MFA_TRY(e->upload(DB_TC_ITEMS, items.data(), items.size(), &d_items));
  const int64_t tot_a = n_at * KC * TM;
  xsplit_kernel<<<(unsigned)((tot_a + 255) / 256), 256, 0, e->stream>>>(d_feats, m->dim, m->d_tc_colscale, d_a, n_at, d_row0, d_rows, 0, KC, TK == 96);
  e->launches++;
  gather_b_kernel<<<(unsigned)(2 * n_bt), 256, 0, e->stream>>>((const __half *)m->d_tc_rows, m->num_gauss, nullptr, d_b, n_bt,
                                                                           TK == 80 ? -1 : 2 * m->dim, KC, m->d_tc_g, TK == 80 ? d_gt : nullptr,
                                                                           (const TcMeta *)g->d_rag + bt0, g->d_lp2pdf, m->d_pdf_off);
  e->launches++;
  TcParams p;
  p.a_img = d_a; p.b_img = d_b; p.meta = (const TcMeta *)g->d_rag + bt0; p.items = d_items; p.n_items = (int)items.size(); p.out = d_out;
  p.g_tiles = d_gt;
  return launch_tc(e, p, TK);
}

}  // namespace mfa
