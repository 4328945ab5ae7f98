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

#include <cmath>
#include <cstring>

#include "cuda_internal.cuh"

using namespace mfa;

namespace {

constexpr int TM = 128, TN = MFA_TILE_N, TK = 96, KC = TK / 8;
constexpr uint32_t IMG_BYTES = TM * TK * 2;       // one fp16 image (hi or lo) of a 128 x 96 tile
constexpr uint32_t TILE_BYTES = 2 * IMG_BYTES;    // hi + lo
constexpr uint32_t LBO_BYTES = (TM / 8) * 128;    // K-adjacent core matrices
constexpr uint32_t SBO_BYTES = 128;               // row-group-adjacent core matrices
constexpr int NTHREADS = 384;
constexpr float kLn2 = 0.69314718055994530942f, kLog2e = 1.44269504088896340736f;

struct TcMeta {  // per Gaussian tile: pdf structure of its 128 columns at 4-column group granularity (32 groups)
  uint32_t gstart, gend;  // bit g: group g starts a pdf / is the last group of a pdf (pdf column ranges are multiples of 4)
  int32_t pdf0, pad;
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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done = 0;
  long long t0 = clock64();
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) break;
    if (clock64() - t0 > 8000000000LL) __trap();  // ~4 s: a protocol bug must not hang the GPU
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
__device__ __forceinline__ void lse_chunk(const uint32_t (&vr)[32], uint32_t gs, uint32_t ge, float &cmx, float &cs, float *&out, int64_t ld,
                                          bool row_ok) {
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

struct TcParams {
  const uint8_t *a_img;   // [n_frame_tiles (even)][TILE_BYTES]
  const uint8_t *b_img;   // [n_gauss_tiles][TILE_BYTES]
  const TcMeta *meta;     // [n_gauss_tiles]
  int n_pairs, n_tiles, n_splits, tiles_per_split;
  float *llT;
  int64_t ld;
};

__global__ void __launch_bounds__(NTHREADS, 1)
gmm_tc_kernel(TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sA = smem;                       // 2 frame tiles x (hi, lo)
  uint8_t *sB = smem + 2 * TILE_BYTES;      // 2 stages x (hi, lo)
  uint64_t *bars = (uint64_t *)(smem + 4 * TILE_BYTES);
  uint64_t *full_a = bars + 0, *empty_a = bars + 1, *full_b = bars + 2, *empty_b = bars + 4, *tfull = bars + 6, *tempty = bars + 10;
  uint32_t *tmem_slot = (uint32_t *)(bars + 14);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(full_a, 1); mbar_init(empty_a, 1);
    for (int s = 0; s < 2; s++) { mbar_init(full_b + s, 1); mbar_init(empty_b + s, 1); }
    for (int i = 0; i < 4; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 128); }
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
  const int n_items = p.n_pairs * p.n_splits;

  if (warp == 0) {
    // ===== producer: bulk copies of the A pair (once per item) and of each B tile =====
    if (lane == 0) {
      uint32_t cnt = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
        const int pair = item / p.n_splits, sp = item % p.n_splits;
        const int n0 = sp * p.tiles_per_split, n1 = min(p.n_tiles, n0 + p.tiles_per_split);
        mbar_wait(empty_a, (it & 1) ^ 1);
        mbar_expect_tx(full_a, 2 * TILE_BYTES);
        bulk_g2s(sA, p.a_img + (size_t)(2 * pair) * TILE_BYTES, TILE_BYTES, full_a);
        bulk_g2s(sA + TILE_BYTES, p.a_img + (size_t)(2 * pair + 1) * TILE_BYTES, TILE_BYTES, full_a);
        for (int n = n0; n < n1; n++, cnt++) {
          const uint32_t s = cnt & 1;
          mbar_wait(empty_b + s, ((cnt >> 1) & 1) ^ 1);
          mbar_expect_tx(full_b + s, TILE_BYTES);
          bulk_g2s(sB + s * TILE_BYTES, p.b_img + (size_t)n * TILE_BYTES, TILE_BYTES, full_b + s);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      // InstrDescriptor: c_format=F32 (1<<4), a/b format F16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      uint32_t cnt = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
        const int sp = item % p.n_splits;
        const int n0 = sp * p.tiles_per_split, n1 = min(p.n_tiles, n0 + p.tiles_per_split);
        mbar_wait(full_a, it & 1);
        tc_fence_after();
        for (int n = n0; n < n1; n++, cnt++) {
          const uint32_t s = cnt & 1, ph = (cnt >> 1) & 1;
          mbar_wait(full_b + s, ph);
          tc_fence_after();
#pragma unroll
          for (int f = 0; f < 2; f++) {
            mbar_wait(tempty + s * 2 + f, ph ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + s * 256 + f * 128;
            const uint32_t a0 = a_base + f * TILE_BYTES, b0 = b_base + s * TILE_BYTES;
#pragma unroll
            for (int prod = 0; prod < 3; prod++) {
              const uint32_t ao = a0 + (prod == 2 ? IMG_BYTES : 0), bo = b0 + (prod == 1 ? IMG_BYTES : 0);
#pragma unroll
              for (int k = 0; k < TK / 16; k++)
                umma_f16(d, make_desc(ao + k * 2 * LBO_BYTES), make_desc(bo + k * 2 * LBO_BYTES), idesc, (prod | k) != 0);
            }
            umma_commit(tfull + s * 2 + f);
          }
          umma_commit(empty_b + s);
        }
        umma_commit(empty_a);
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: warpgroup f handles frame tile f; thread = one frame (TMEM lane) =====
    const int f = (warp - 4) >> 2, wq = warp & 3;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    uint32_t cnt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int pair = item / p.n_splits, sp = item % p.n_splits;
      const int n0 = sp * p.tiles_per_split, n1 = min(p.n_tiles, n0 + p.tiles_per_split);
      const int64_t row = (int64_t)(2 * pair + f) * TM + wq * 32 + lane;
      const bool row_ok = row < p.ld;
      TcMeta mt = p.meta[n0];
      for (int n = n0; n < n1; n++, cnt++) {
        const uint32_t s = cnt & 1, ph = (cnt >> 1) & 1;
        const TcMeta cur = mt;
        if (n + 1 < n1) mt = p.meta[n + 1];
        mbar_wait(tfull + s * 2 + f, ph);
        tc_fence_after();
        const uint32_t t0 = tmem_base + lane_base + s * 256 + f * 128;
        float cmx = -INFINITY, cs = 0.0f;
        float *out = p.llT + (size_t)cur.pdf0 * p.ld + row;
        uint32_t va[32], vb[32];
        tmem_ld32(t0, va);
        tmem_ld_wait();
        tmem_ld32(t0 + 32, vb);                    // chunk c+1 is in flight while chunk c is reduced
        lse_chunk(va, cur.gstart & 0xFF, cur.gend & 0xFF, cmx, cs, out, p.ld, row_ok);
        tmem_ld_wait();
        tmem_ld32(t0 + 64, va);
        lse_chunk(vb, (cur.gstart >> 8) & 0xFF, (cur.gend >> 8) & 0xFF, cmx, cs, out, p.ld, row_ok);
        tmem_ld_wait();
        tmem_ld32(t0 + 96, vb);
        lse_chunk(va, (cur.gstart >> 16) & 0xFF, (cur.gend >> 16) & 0xFF, cmx, cs, out, p.ld, row_ok);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(tempty + s * 2 + f);           // accumulator drained: the MMA warp may overwrite it
        lse_chunk(vb, (cur.gstart >> 24) & 0xFF, (cur.gend >> 24) & 0xFF, cmx, cs, out, p.ld, row_ok);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// features fp32 [n_rows][dim] -> A images (fp16 hi / lo, canonical layout), one 48 KB image per 128-row tile
__global__ void xsplit_kernel(const float *__restrict__ feats, int64_t n_rows, int dim, const float *__restrict__ colscale, uint8_t *__restrict__ a_img,
                              int64_t n_tiles) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (tile, kc, row)
  if (idx >= n_tiles * KC * TM) return;
  const int r = (int)(idx % TM), kc = (int)((idx / TM) % KC);
  const int64_t tile = idx / (TM * KC), row = tile * TM + r;
  __half hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const int k = kc * 8 + e;
    float a = 0.0f;
    if (row < n_rows) {
      if (k < dim) a = feats[row * dim + k] * colscale[k];
      else if (k < 2 * dim) { float x = feats[row * dim + (k - dim)] * colscale[k - dim]; x = fminf(fmaxf(x, -240.0f), 240.0f); a = x * x; }
      else if (k < 2 * dim + 3) a = 1.0f;
      if (k < dim) a = fminf(fmaxf(a, -240.0f), 240.0f);
    }
    hi[e] = __float2half_rn(a);
    lo[e] = __float2half_rn(a - __half2float(hi[e]));
  }
  const size_t off = (size_t)tile * TILE_BYTES + ((size_t)(kc * (TM / 8) + r / 8) * 64 + (size_t)(r % 8) * 8) * 2;
  *(uint4 *)(a_img + off) = *(const uint4 *)hi;
  *(uint4 *)(a_img + off + IMG_BYTES) = *(const uint4 *)lo;
}

// host: fp16 hi/lo images of the weights + per-tile segment masks
int build_tc(mfa_model *m) {
  const int D = m->dim;
  if (2 * D + 3 > TK) return set_error(MFA_ERR_UNSUPPORTED, "tensor-core GMM kernel needs 2*dim+3 <= 96");
  // per-dimension power-of-two scale: x*s has roughly unit spread under the model
  std::vector<double> m1(D, 0.0), m2(D, 0.0);
  for (int g = 0; g < m->num_gauss; g++)
    for (int d = 0; d < D; d++) {
      double iv = m->h_iv[(size_t)g * D + d], mu = m->h_miv[(size_t)g * D + d] / iv;
      m1[d] += mu; m2[d] += mu * mu + 1.0 / iv;
    }
  m->h_tc_colscale.assign(D, 1.0f);
  for (int d = 0; d < D; d++) {
    double mean = m1[d] / m->num_gauss, var = m2[d] / m->num_gauss - mean * mean;
    double rms = std::sqrt(std::max(m2[d] / m->num_gauss, 1e-30));  // second moment about 0: features are not re-centred
    (void)var;
    m->h_tc_colscale[d] = (float)std::exp2(-std::round(std::log2(rms)));
  }
  const int nt = m->n_tiles;
  std::vector<uint8_t> img((size_t)nt * TILE_BYTES, 0);
  std::vector<TcMeta> meta(nt);
  double wmax = 0.0;
  auto put = [&](int tile, int which, int row, int k, float val) {
    size_t off = (size_t)tile * TILE_BYTES + (size_t)which * IMG_BYTES + ((size_t)((k / 8) * (TN / 8) + row / 8) * 64 + (size_t)(row % 8) * 8 + (k % 8)) * 2;
    __half h = __float2half_rn(val);
    memcpy(&img[off], &h, 2);
  };
  auto split2 = [&](int tile, int row, int k, double w) {
    float hi = __half2float(__float2half_rn((float)w));
    put(tile, 0, row, k, hi);
    put(tile, 1, row, k, (float)(w - (double)hi));
    wmax = std::max(wmax, std::fabs(w));
  };
  // walk the tiling produced by mfa_model::rebuild_tiles (whole pdfs per tile, in pdf order, column ranges padded to 4)
  for (int tl = 0; tl < nt; tl++) {
    int col = 0;
    memset(&meta[tl], 0, sizeof(TcMeta));
    meta[tl].pdf0 = m->h_tile_pdf0[tl];
    for (int pdf = m->h_tile_pdf0[tl]; pdf < m->h_tile_pdf0[tl + 1]; pdf++) {
      const int ng = m->h_pdf_off[pdf + 1] - m->h_pdf_off[pdf], pad = (ng + MFA_SEG_ALIGN - 1) / MFA_SEG_ALIGN * MFA_SEG_ALIGN;
      meta[tl].gstart |= 1u << (col / 4);
      meta[tl].gend |= 1u << ((col + pad - 1) / 4);
      int c = col;
      for (int g = m->h_pdf_off[pdf]; g < m->h_pdf_off[pdf + 1]; g++, c++) {
        for (int d = 0; d < D; d++) {
          double s = m->h_tc_colscale[d];
          split2(tl, c, d, (double)m->h_miv[(size_t)g * D + d] / s * kLog2e);
          split2(tl, c, D + d, -0.5 * (double)m->h_iv[(size_t)g * D + d] / (s * s) * kLog2e);
        }
        double gc = (double)m->h_gconsts[g] * kLog2e;
        if (!(gc > -60000.0)) gc = -60000.0;
        float g1 = __half2float(__float2half_rn((float)gc));
        float g2 = __half2float(__float2half_rn((float)(gc - g1)));
        float g3 = (float)(gc - g1 - g2);
        put(tl, 0, c, 2 * D, g1); put(tl, 0, c, 2 * D + 1, g2); put(tl, 0, c, 2 * D + 2, g3);
      }
      for (; c < col + pad; c++) put(tl, 0, c, 2 * D, -60000.0f);  // padding columns inside the pdf's range: exp2 -> 0
      col += pad;
    }
    if (col < TN) meta[tl].gstart |= 1u << (col / 4);  // trailing padding: one junk segment that never ends
    for (; col < TN; col++) put(tl, 0, col, 2 * D, -60000.0f);
  }
  if (wmax > 60000.0) return set_error(MFA_ERR_UNSUPPORTED, "model weights exceed the fp16 range of the tensor-core kernel");
  cudaStream_t s = m->eng->stream;
  if (m->d_tc_w) { CUDA_TRY(cudaStreamSynchronize(s)); CUDA_TRY(cudaFree(m->d_tc_w)); m->d_tc_w = nullptr; }
  if (m->d_tc_colscale) { CUDA_TRY(cudaFree(m->d_tc_colscale)); m->d_tc_colscale = nullptr; }
  size_t meta_bytes = (size_t)nt * sizeof(TcMeta);
  m->tc_w_bytes = img.size();
  CUDA_TRY(cudaMalloc(&m->d_tc_w, img.size() + meta_bytes));
  CUDA_TRY(cudaMemcpyAsync(m->d_tc_w, img.data(), img.size(), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync((uint8_t *)m->d_tc_w + img.size(), meta.data(), meta_bytes, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMalloc((void **)&m->d_tc_colscale, D * sizeof(float)));
  CUDA_TRY(cudaMemcpyAsync(m->d_tc_colscale, m->h_tc_colscale.data(), D * sizeof(float), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  m->tc_ready = true;
  return MFA_OK;
}

}  // namespace

namespace mfa {

int launch_gmm_tc(mfa_engine *e, mfa_model *m, const float *d_feats, int64_t n_rows, float *d_llT, int64_t ld) {
  if (n_rows == 0) return MFA_OK;
  if (ld < n_rows) return set_error(MFA_ERR_INVALID, "ld < n_rows");
  if (!m->tc_ready) {
    int r = build_tc(m);
    if (r == MFA_ERR_UNSUPPORTED) return launch_gmm_ffma(e, m, d_feats, n_rows, d_llT, ld);  // shapes the tcgen05 kernel does not cover
    if (r) return r;
  }
  const int64_t n_ftiles = (n_rows + TM - 1) / TM, n_pairs = (n_ftiles + 1) / 2;
  uint8_t *d_a;
  MFA_TRY(e->getT<uint8_t>(DB_XSPLIT, (size_t)n_pairs * 2 * TILE_BYTES, &d_a));
  const int64_t total = n_pairs * 2 * KC * TM;
  xsplit_kernel<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(d_feats, n_rows, m->dim, m->d_tc_colscale, d_a, n_pairs * 2);
  e->launches++;
  TcParams p;
  p.a_img = d_a; p.b_img = (const uint8_t *)m->d_tc_w; p.meta = (const TcMeta *)((const uint8_t *)m->d_tc_w + m->tc_w_bytes);
  p.n_pairs = (int)n_pairs; p.n_tiles = m->n_tiles; p.llT = d_llT; p.ld = ld;
  int splits = 1;
  if (n_pairs < 2 * (int64_t)e->sm_count) splits = (int)std::min<int64_t>(m->n_tiles, (2 * (int64_t)e->sm_count + n_pairs - 1) / n_pairs);
  p.tiles_per_split = (m->n_tiles + splits - 1) / splits;
  p.n_splits = (m->n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  const size_t smem = 4 * (size_t)TILE_BYTES + 256;
  CUDA_TRY(cudaFuncSetAttribute(gmm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t items = n_pairs * p.n_splits;
  const int grid = (int)std::min<int64_t>(items, e->sm_count);
  gmm_tc_kernel<<<grid, NTHREADS, smem, e->stream>>>(p);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}

}  // namespace mfa
