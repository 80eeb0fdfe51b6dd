// K8a: residual-VQ nearest-codebook search on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Replaces the EuclideanCodebook distance + argmax of vector_quantize_pytorch.ResidualVQ (eval) called
// at meshgpt_pytorch.py:842 - a (R x K x D) = (28288 x 16384 x 192) contraction per level at the
// MeshGPT shape, which the reference does as an fp32 cuBLAS GEMM into a materialised (R, K) tensor.
//
// Screening GEMM:   s~[r, j] = |e_j|^2 - 2 <fp16(r), fp16(e_j)>          (fp32 accumulate in TMEM)
//   * operands are fp16, not bf16: tcgen05 kind::f16 runs both at the same rate, and fp16's 11-bit
//     significand makes the screening error bound 8x tighter, which is what keeps the number of rows
//     that need the exact re-scan negligible.  Rows are pre-scaled by a per-row power of two (max |r_d|
//     -> [0.5, 1)) and the codebook by one global power of two, so nothing overflows and the scaling is
//     undone exactly in the epilogue;
//   * work unit = (256-row tile of residuals) x (one split of the codebook, <= 4096 codes);
//   * the row tile lives in shared memory as fp16 in the canonical K-major UMMA layout (two M=128
//     halves -> two accumulators), the codebook streams through a 6-stage ring of 16 KB stages
//     (128 codes x 64 dims) filled by cp.async.bulk from a pre-tiled fp16 image of the codebook;
//   * accumulators are double buffered in TMEM (2 buffers x 2 halves x 128 columns = all 512 columns) so
//     the epilogue of tile t overlaps the MMAs of tile t+1;
//   * epilogue warps (one TMEM lane = one row per thread) reduce every 8-code chunk to its minimum s~ (one FFMA
//     and one FMNMX per value - the epilogue, not the MMA, was the limiter with per-code bookkeeping) and keep the
//     two best chunks per (row, split).
// The exact fp32 decision is taken by rvq_rescore_kernel on every code of the screened chunks, with the
// reference's own formula; rows whose candidate list might be incomplete go to the exact fp32 scan.
//
// Warp roles (384 threads): w0 bulk-copy producer, w1 MMA issuer, w2 TMEM allocator, w3 idle,
// w4..w11 epilogue + row-tile loader (w4-7: rows 0-127, w8-11: rows 128-255; TMEM lane quarter = warp%4).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"
#include "rvq_screen.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace meshtok {

namespace {

using namespace tc;

constexpr int M_TILE = 256;        // rows per unit (two UMMA M=128 halves)
constexpr int N_TILE = 128;        // codes per accumulator tile (UMMA N)
constexpr int KC = 64;             // K elements per pipeline stage
constexpr int STAGES = 6;
constexpr int STAGE_BYTES = N_TILE * KC * 2;   // 16 KB
constexpr int CHUNK = RVQ_CHUNK;               // screening granularity: codes per candidate chunk
constexpr int MAX_SPLIT_CODES = 4096;          // e2 slice kept in shared memory per unit
constexpr int NUM_THREADS = 384;
constexpr int EPI_THREADS = 256;
constexpr uint32_t SBO_B = (KC / 8) * 128;     // 1024: next 8 codes inside a stage

template <int D>
struct Smem {
  static constexpr uint32_t SBO_A = (D / 8) * 128;
  static constexpr uint32_t HALF_BYTES = 128 * D * 2;
  static constexpr uint32_t A_BYTES = 2 * HALF_BYTES;
  static constexpr uint32_t B_OFF = A_BYTES;
  static constexpr uint32_t E2_OFF = B_OFF + STAGES * STAGE_BYTES;
  static constexpr uint32_t BAR_OFF = E2_OFF + MAX_SPLIT_CODES * 4;
  static constexpr uint32_t TOTAL = BAR_OFF + 256;
};

struct SearchArgs {
  const float* rows;       // (R, D) fp32 residuals
  const uint8_t* image;    // fp16 codebook image: [K/128 tiles][D/64 chunks][16 KB canonical stage]
  const float* e2;         // (K) |e_j|^2
  const float* rscale;     // (R) power-of-two row scales
  float escale;            // power-of-two codebook scale baked into the image
  const int* n_rows_dev;
  int max_rows, K, n_splits, tiles_per_split;
  float4* partial;         // (R, n_splits) entries of RVQ_SLOT_F4 float4 (rvq_screen.cuh)
  int debug;               // timing experiments only (MESHTOK_RVQ_DEBUG): 1 = epilogue skips the TMEM loads, 2 = loads but no
                           // math, 3 = every CTA walks its code tiles from a different start (no lockstep reads of one tile)
};

template <int D, int DBG>
__global__ void __launch_bounds__(NUM_THREADS, 1) rvq_tc_search_kernel(SearchArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using L = Smem<D>;
  constexpr int KCHUNKS = D / KC;
  uint8_t* sA = smem;
  uint8_t* sB = smem + L::B_OFF;
  float* sE2 = reinterpret_cast<float*>(smem + L::E2_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* b_full = bars;                 // [STAGES]
  uint64_t* b_empty = bars + STAGES;       // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;  // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint64_t* a_full = acc_empty + 2;
  uint64_t* a_empty = a_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 1);

  const int warp = warp_id(), lane = lane_id();

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI_THREADS / 32); }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above is independent of earlier kernels; from here on their results are read
  pdl_wait();
  const int n_rows = min(*a.n_rows_dev, a.max_rows);
  const int m_tiles = (n_rows + M_TILE - 1) / M_TILE;
  const int units = m_tiles * a.n_splits;
  // contiguous unit ranges so that consecutive units of one CTA share the row tile
  const int u0 = (int)((long long)units * blockIdx.x / gridDim.x);
  const int u1 = (int)((long long)units * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0) {
    // ===== producer (warp-converged; one elected lane issues): stream codebook stages =====
    uint32_t it = 0;
    for (int u = u0; u < u1; ++u) {
      const int tile0 = (u % a.n_splits) * a.tiles_per_split;
      for (int t0 = 0; t0 < a.tiles_per_split; ++t0) {
        const int t = DBG == 3 ? (t0 + (int)blockIdx.x * 5) % a.tiles_per_split : t0;
        const uint8_t* src = a.image + (size_t)(tile0 + t) * rvq_image_tile_bytes(D);   // (the 4 KB extra block is not used here)
        for (int kc = 0; kc < KCHUNKS; ++kc, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&b_empty[s], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&b_full[s], STAGE_BYTES);
            bulk_g2s(sB + s * STAGE_BYTES, src + (size_t)kc * STAGE_BYTES, STAGE_BYTES, &b_full[s]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-converged; one elected lane issues) =====
    constexpr uint32_t idesc = make_idesc_f16_f32(128, N_TILE, /*bf16=*/false);
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    uint32_t it = 0, acc_it = 0, a_phase = 0;
    int prev_m = -1;
    for (int u = u0; u < u1; ++u) {
      const int m_tile = u / a.n_splits;
      if (m_tile != prev_m) {
        mbar_wait(a_full, a_phase);
        a_phase ^= 1;
        prev_m = m_tile;
        tcgen05_fence_after();
      }
      for (int t = 0; t < a.tiles_per_split; ++t, ++acc_it) {
        const uint32_t buf = acc_it & 1, aph = (acc_it >> 1) & 1;
        mbar_wait(&acc_empty[buf], aph ^ 1);
        tcgen05_fence_after();
        for (int kc = 0; kc < KCHUNKS; ++kc, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&b_full[s], ph);
          tcgen05_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int k16 = 0; k16 < KC / 16; ++k16) {
                const uint64_t da = make_smem_desc_sw128(sA_addr + (kc * 2 + h) * STAGE_BYTES + k16 * 32);
                const uint64_t db = make_smem_desc_sw128(sB_addr + s * STAGE_BYTES + k16 * 32);
                umma_f16(tmem_base + (buf * 2 + h) * N_TILE, da, db, idesc, (kc | k16) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&b_empty[s]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&acc_full[buf]);
        __syncwarp();
      }
      const int next_m = (u + 1 < u1) ? (u + 1) / a.n_splits : -1;
      if (next_m != m_tile) {
        if (elect_one()) umma_commit(a_empty);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== row-tile loader + epilogue =====
    const int ew = warp - 4;                 // 0..7
    const int et = threadIdx.x - 128;        // 0..255
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int h = ew >> 2;                   // row half
    const int row_in_tile = h * 128 + q * 32 + lane;
    uint32_t acc_it = 0, a_empty_phase = 0;
    int prev_m = -1;
    for (int u = u0; u < u1; ++u) {
      const int m_tile = u / a.n_splits, split = u % a.n_splits;
      const int tile0 = split * a.tiles_per_split;
      named_bar_sync(1, EPI_THREADS);  // every epilogue thread is done with the previous unit's sE2
      const bool new_tile = m_tile != prev_m;
      if (new_tile) {
        if (prev_m != -1) { mbar_wait(a_empty, a_empty_phase); a_empty_phase ^= 1; }
        // fp32 rows -> bf16 canonical layout.  One warp-iteration = 8 rows x 4 k-chunks (16 B each):
        // global reads are 8 x 128 contiguous bytes, shared writes 512 contiguous bytes.
        constexpr int KQ = D / 32;
        for (int item = ew; item < (M_TILE / 8) * KQ; item += EPI_THREADS / 32) {
          const int g = item / KQ, kq = item % KQ;
          const int r = g * 8 + (lane & 7);
          const int k8 = kq * 4 + (lane >> 3);
          const long long grow = (long long)m_tile * M_TILE + r;
          float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
          float sc = 1.f;
          if (grow < n_rows) {
            const float4* src = reinterpret_cast<const float4*>(a.rows + (size_t)grow * D + k8 * 8);
            v0 = src[0];
            v1 = src[1];
            sc = __ldg(a.rscale + grow);
          }
          __half2 p0 = __floats2half2_rn(v0.x * sc, v0.y * sc), p1 = __floats2half2_rn(v0.z * sc, v0.w * sc);
          __half2 p2 = __floats2half2_rn(v1.x * sc, v1.y * sc), p3 = __floats2half2_rn(v1.z * sc, v1.w * sc);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
          pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
          // [k chunk of 64][row half][128 rows x 128 B, SWIZZLE_128B]
          const uint32_t off = (uint32_t)(((k8 >> 3) * 2 + (r >> 7)) * STAGE_BYTES) + sw128_offset(r & 127, (k8 & 7) * 8);
          *reinterpret_cast<uint4*>(sA + off) = pk;
        }
        fence_proxy_async_smem();
      }
      for (int i = et; i < a.tiles_per_split * N_TILE; i += EPI_THREADS) sE2[i] = __ldg(a.e2 + (size_t)tile0 * N_TILE + i);
      named_bar_sync(1, EPI_THREADS);
      if (new_tile) {
        if (et == 0) mbar_arrive(a_full);
        prev_m = m_tile;
      }

      RvqTop top;
      top.reset();
      const long long my_row = (long long)m_tile * M_TILE + row_in_tile;
      // acc = <r * rscale, e * escale>  ->  -2 <r, e> = acc * cmul
      const float cmul = -2.f / ((my_row < n_rows ? __ldg(a.rscale + my_row) : 1.f) * a.escale);
      for (int t0 = 0; t0 < a.tiles_per_split; ++t0, ++acc_it) {
        const int t = DBG == 3 ? (t0 + (int)blockIdx.x * 5) % a.tiles_per_split : t0;
        const uint32_t buf = acc_it & 1, aph = (acc_it >> 1) & 1;
        mbar_wait(&acc_full[buf], aph);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (buf * 2 + h) * N_TILE;
        const int chunk0 = (tile0 + t) * (N_TILE / CHUNK);
        const float* e2t = sE2 + t * N_TILE;
        // Per CHUNK (8) codes only the MINIMUM screened score is kept (one FFMA + one FMNMX per value, one
        // compare + rarely-taken branch per chunk); the re-score kernel evaluates every code of the winning
        // chunks exactly.  TMEM loads are double buffered: 32 columns are in flight while 32 are reduced.
        uint32_t va[32], vb[32];
        if (DBG == 1) {   // experiment: how fast is the MMA pipeline when nobody reads the accumulators
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
          continue;
        }
        tmem_ld32(taddr, va);
#pragma unroll
        for (int c = 0; c < N_TILE / 32; ++c) {
          tmem_ld_wait();
          uint32_t(&v)[32] = (c & 1) ? vb : va;
          if (c + 1 < N_TILE / 32) tmem_ld32(taddr + (c + 1) * 32, (c & 1) ? va : vb);
          if (DBG == 2) {   // experiment: TMEM loads without the reduction
            if (__uint_as_float(v[lane & 31]) == 123.456f) top.insert(0.f, c);
            continue;
          }
          rvq_screen32(v, e2t + c * 32, cmul, chunk0 + (c * 32) / CHUNK, top);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
      const long long grow = (long long)m_tile * M_TILE + row_in_tile;
      if (grow < n_rows)
        top.store(a.partial + ((size_t)grow * a.n_splits + split) * RVQ_SLOT_F4);
    }
  }

  pdl_launch_dependents();   // this CTA's work is done: let the next kernel's CTAs start their prologue
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- re-score (+ exact scan of the rare ambiguous row + residual update), ONE launch per level ----------------------------
// One warp per row.  Screening error bound: both operands are rounded to fp16 (relative 2^-11 each, the
// power-of-two scalings are exact) so |s~ - s| <= 2 ((1+2^-11)^2 - 1) sum_d |r_d e_jd| < 2^-9 * 1.01 * |r| max|e|,
// plus 2^-25 absolute per element that falls into the fp16 subnormal range after scaling  =: delta  (fp32
// accumulation error is orders of magnitude below).  Every code whose true score ties the minimum therefore has
// s~ <= min s~ + 2 delta; if some slot's LAST kept chunk (RVQ_TOPK per slot) is inside that window a further one may hide
// behind it: the row is "ambiguous" and gets the exact fp32 scan of the whole codebook.
//
// Round 1 ran this as four launches per level (re-score, exact scan of the flagged rows, residual update / codes-only); the
// scan launch was empty at the bench workload (0 of 56 448 row-levels flagged) and still cost its launch latency twice per
// step.  Now the warp that decided a row also finishes it: code id -> vcodes, r <- r - e[idx], the next level's row scale and
// fp16 row image (the row is in its registers already).  An ambiguous row is scanned by ITS CTA, all warps together, one
// code per thread at a time with the row broadcast from shared memory (sequential fmaf over the dimensions like
// rvq_exact_kernel; the screened candidates are then ignored, so the result is the plain fp32 arg-min with the lowest
// index on ties): ~20 us for the CTA that holds such a row, nothing for the others, no launch for the common case.
template <int PER2>   // float2 pairs per lane: D <= 64 * PER2
__global__ void __launch_bounds__(256) rvq_rescore_kernel(float* __restrict__ rows, int D, const float* __restrict__ cb,
                                                          const float* __restrict__ e2, int K, float max_norm, float escale, float c_scale,
                                                          float* __restrict__ rscale,
                                                          const float4* __restrict__ partial, RvqPartialLayout lay,
                                                          const int* __restrict__ n_dev, int max_rows,
                                                          int32_t* __restrict__ codes, int Q, int level, int update,
                                                          uint8_t* __restrict__ row_image, int* __restrict__ stats) {
  __shared__ float s_x[64 * PER2];
  __shared__ float s_x2;
  __shared__ int s_row[8];
  __shared__ unsigned long long s_best[8];
  pdl_wait();
  const int lane = lane_id(), warp = warp_id();
  const int wpb = blockDim.x >> 5;   // 8
  const int n = min(*n_dev, max_rows);
  const int d2 = D >> 1;
  // loop invariants, hoisted by hand (the 64-bit divisions of the work split were a third of this kernel's instructions
  // when they sat in the row loop): W = steps per cluster of the CTA-pair search, see rvq_tc2.cu
  unsigned W = 1u;
  if (lay.mode == 2) {
    const long long S = (long long)((n + 255) >> 8) * lay.T;
    long long w = (S + lay.n_clusters - 1) / lay.n_clusters;
    if (w < lay.w_min) w = lay.w_min;
    W = (unsigned)w;
  }
  const float sqrt_d = sqrtf((float)D);
  const float mn2 = max_norm * max_norm;
  const float inv_escale = 1.f / escale;   // powers of two: the reciprocals are exact
  const int stride = gridDim.x * wpb;
  const int iters = (n + stride - 1) / stride;   // the same trip count for every warp of the grid: the loop holds CTA barriers
  for (int it = 0; it < iters; ++it) {
    const int row = it * stride + blockIdx.x * wpb + warp;
    const bool active = row < n;
    float2 x[PER2];
    float x2 = 0.f;
    float bd = INFINITY;
    int bi = 0x7fffffff;
    bool ambiguous = false;
    if (active) {
      const float2* xr = reinterpret_cast<const float2*>(rows + (size_t)row * D);
#pragma unroll
      for (int i = 0; i < PER2; ++i) {
        const int c = lane + 32 * i;
        x[i] = c < d2 ? xr[c] : make_float2(0.f, 0.f);
      }
      // this row's candidate entries (see RvqPartialLayout): the slots of the clusters whose step range meets the row tile's.  Lane s
      // holds entry s (n_splits <= 32 is checked by the launcher).  Issued together with the row and its scale, BEFORE the first
      // reduction: the kernel is bound by the dependent round trips of a row, and these three do not depend on each other.
      int n_splits = lay.stride;
      if (lay.mode == 2) {
        const unsigned a = (unsigned)(row >> 8) * (unsigned)lay.T;      // first step of the 256-row tile (the launcher checks the range)
        n_splits = lay.parts * (int)((a + (unsigned)lay.T - 1u) / W - a / W + 1u);
      }
      const float4* prow = partial + (size_t)row * lay.stride * RVQ_SLOT_F4;
      float4 p0 = make_float4(INFINITY, 0.f, INFINITY, 0.f), p1 = p0;
      if (lane < n_splits) { p0 = prow[lane * RVQ_SLOT_F4]; p1 = prow[lane * RVQ_SLOT_F4 + 1]; }
      const float rs = rscale[row];
#pragma unroll
      for (int i = 0; i < PER2; ++i) {
        x2 = fmaf(x[i].x, x[i].x, x2);
        x2 = fmaf(x[i].y, x[i].y, x2);
      }
      x2 = warp_sum(x2);
      const float xn = sqrtf(x2);
      float delta = 0.001953125f * 1.01f * xn * max_norm +
                    5.96e-8f * sqrt_d * (max_norm * __frcp_rn(rs) + xn * inv_escale) + 1e-7f * (x2 + mn2);
      if (c_scale > 0.f) {
        // CTA-pair kernel: |e|^2 enters the MMA chain as an fp16 (hi, lo) pair (relative error 2^-21) times the row constant
        // rscale * c_scale; should that constant underflow fp16 (astronomically large rows) the |e|^2 term is lost altogether
        delta += 1e-6f * mn2;
        if (rs * c_scale < 5.9604645e-8f) delta += mn2;
      }
      float smin = p0.x;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
      const float window = smin + 2.f * delta;
      // chunks of this lane's entry inside the window (ascending list -> a prefix), and whether the list may be incomplete
      const int mine = (p0.x <= window) + (p0.z <= window) + (p1.x <= window) + (p1.z <= window);
      // (a NaN row has no chunk inside the window and is not ambiguous: it falls through to code 0 below)
      ambiguous = __any_sync(0xffffffffu, p1.z <= window);
      // walk the candidates entry by entry (ascending entry = ascending code range within a cluster's slots is not
      // guaranteed, so ties are resolved explicitly on the code index below)
      unsigned live = ambiguous ? 0u : __ballot_sync(0xffffffffu, mine > 0);   // (an ambiguous row is decided by the full scan alone)
      while (live) {
        const int owner = __ffs(live) - 1;
        live &= live - 1;
        const int cnt = __shfl_sync(0xffffffffu, mine, owner);
        const float y0 = __shfl_sync(0xffffffffu, p0.y, owner), y1 = __shfl_sync(0xffffffffu, p0.w, owner);
        const float y2 = __shfl_sync(0xffffffffu, p1.y, owner), y3 = __shfl_sync(0xffffffffu, p1.w, owner);
#pragma unroll 1
        for (int which = 0; which < cnt; ++which) {
          const int code0 = __float_as_int(which == 0 ? y0 : which == 1 ? y1 : which == 2 ? y2 : y3) * CHUNK;
          // exact fp32 distance of all CHUNK codes at once (coalesced codebook rows, CHUNK independent dot products)
          float d[CHUNK], en[CHUNK];
#pragma unroll
          for (int j = 0; j < CHUNK; ++j) { d[j] = 0.f; en[j] = __ldg(e2 + code0 + j); }   // |e|^2 travels with the codebook rows
#pragma unroll
          for (int i = 0; i < PER2; ++i) {
            const int c = lane + 32 * i;
            if (c < d2) {
#pragma unroll
              for (int j = 0; j < CHUNK; ++j) {
                const float2 e = __ldg(reinterpret_cast<const float2*>(cb + (size_t)(code0 + j) * D) + c);
                d[j] = fmaf(x[i].x, e.x, d[j]);
                d[j] = fmaf(x[i].y, e.y, d[j]);
              }
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) d[j] += __shfl_xor_sync(0xffffffffu, d[j], o);
#pragma unroll
          for (int j = 0; j < CHUNK; ++j) {
            const float q = sqrtf(fmaxf(__fadd_rn(__fadd_rn(x2, en[j]), __fmul_rn(d[j], -2.f)), 0.f));
            if (q < bd || (q == bd && code0 + j < bi)) { bd = q; bi = code0 + j; }
          }
        }
      }
    }
    // ---- ambiguous rows of this CTA: exact fp32 scan of the whole codebook by all warps (rare; the barrier is the common-case cost) ----
    if (__syncthreads_or(ambiguous ? 1 : 0)) {
      if (lane == 0) { s_row[warp] = ambiguous ? row : -1; s_best[warp] = ~0ull; }
      __syncthreads();
      for (int w = 0; w < wpb; ++w) {
        const int r = s_row[w];
        if (r < 0) continue;   // (uniform: shared memory)
        for (int k = threadIdx.x; k < D; k += blockDim.x) s_x[k] = rows[(size_t)r * D + k];
        if (warp == w && lane == 0) s_x2 = x2;   // the row's owner has |x|^2 (the same value for every code, whatever its rounding)
        __syncthreads();
        const float xx = s_x2;
        float sd = INFINITY;
        int si = 0x7fffffff;
        for (int code = threadIdx.x; code < K; code += blockDim.x) {   // ascending per thread: strict < keeps the lowest index
          const float4* er = reinterpret_cast<const float4*>(cb + (size_t)code * D);
          float acc = 0.f;
          for (int k4 = 0; k4 < (D >> 2); ++k4) {
            const float4 e = __ldg(er + k4);
            const float4 xv = *reinterpret_cast<const float4*>(s_x + 4 * k4);
            acc = fmaf(xv.x, e.x, acc); acc = fmaf(xv.y, e.y, acc); acc = fmaf(xv.z, e.z, acc); acc = fmaf(xv.w, e.w, acc);
          }
          const float q = sqrtf(fmaxf(__fadd_rn(__fadd_rn(xx, __ldg(e2 + code)), __fmul_rn(acc, -2.f)), 0.f));
          if (q < sd) { sd = q; si = code; }
        }
        // (distance bits << 32 | index): non-negative floats order like unsigned ints, equal distances resolve to the lowest index
        unsigned long long key = si == 0x7fffffff ? ~0ull : (((unsigned long long)__float_as_uint(sd) << 32) | (unsigned int)si);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
          key = other < key ? other : key;
        }
        if (lane == 0) atomicMin(&s_best[w], key);
        __syncthreads();   // s_best[w] complete; s_x free for the next row
      }
      if (ambiguous) {
        const unsigned long long key = s_best[warp];
        bd = __uint_as_float((unsigned int)(key >> 32));
        bi = key == ~0ull ? 0x7fffffff : (int)(unsigned int)(key & 0xffffffffull);
        if (lane == 0) atomicAdd(stats + 1, 1);
      }
    }
    if (!active) continue;
    // ---- decided: code id, residual update, the next level's row scale and image ----
    if (bi == 0x7fffffff) bi = 0;   // NaN row: no comparison succeeded (torch.argmax of an all-NaN row is 0 as well)
    if (lane == 0) codes[(size_t)row * Q + level] = bi;
    if (update) {
      const float2* er = reinterpret_cast<const float2*>(cb + (size_t)bi * D);
      float2* dst = reinterpret_cast<float2*>(rows + (size_t)row * D);
      float vmax = 0.f;
#pragma unroll
      for (int i = 0; i < PER2; ++i) {
        const int c = lane + 32 * i;
        if (c < d2) {
          const float2 e = __ldg(er + c);
          x[i].x = __fsub_rn(x[i].x, e.x);
          x[i].y = __fsub_rn(x[i].y, e.y);
          dst[c] = x[i];
          vmax = fmaxf(vmax, fmaxf(fabsf(x[i].x), fabsf(x[i].y)));
        }
      }
      rvq_emit_row_scale_and_image<PER2>(x, vmax, row, D, rscale, row_image, c_scale);
    }
  }
  pdl_launch_dependents();
}

// Round 2: the same re-score with FOUR rows per warp (eight lanes per row).  The kernel above executes ~730 warp instructions per row,
// each once, and most of them are scalar per-row work (error window, step ranges of the partial layout, comparisons, addresses) that
// all 32 lanes repeat; at 1 024 meshes per launch it is issue bound (0.6 instructions per cycle and scheduler).  Here a lane group of 8
// owns a row: lane `sub` holds the float4 columns sub, sub + 8, ... (128-byte row pieces per step, coalesced), reductions take three
// shuffle steps instead of five, and the scalar work of four rows shares one instruction stream.  Same arithmetic per code
// (fp32 dot product, sqrt(max(|x|^2 + |e|^2 - 2 <x, e>, 0)), lowest index on ties); only the summation order inside a dot product
// differs (eight partial sums instead of 32), which the parity tests cover like any other fp32 evaluation order.
template <int NI, int MINB>   // float4 per lane: D <= 32 * NI; MINB CTAs per SM the register allocation is held to
__global__ void __launch_bounds__(256, MINB) rvq_rescore4_kernel(float* __restrict__ rows, int D, const float* __restrict__ cb,
                                                              const float* __restrict__ e2, int K, float max_norm, float escale, float c_scale,
                                                              float* __restrict__ rscale, const float4* __restrict__ partial, RvqPartialLayout lay,
                                                              const int* __restrict__ n_dev, int max_rows, int32_t* __restrict__ codes, int Q,
                                                              int level, int update, uint8_t* __restrict__ row_image, int* __restrict__ stats) {
  __shared__ float s_x[32 * 4 * NI];
  __shared__ float s_x2;
  __shared__ int s_row[32];
  __shared__ unsigned long long s_best[32];
  pdl_wait();
  const int lane = lane_id(), warp = warp_id();
  const int sub = lane & 7, grp = lane >> 3;
  const unsigned gmask = 0xffu << (grp * 8);
  const int rpb = (blockDim.x >> 5) * 4;   // rows per CTA and iteration: 32
  const int n = min(*n_dev, max_rows);
  const int d4 = D >> 2;
  unsigned W = 1u;
  if (lay.mode == 2) {
    const long long S = (long long)((n + 255) >> 8) * lay.T;
    long long w = (S + lay.n_clusters - 1) / lay.n_clusters;
    if (w < lay.w_min) w = lay.w_min;
    W = (unsigned)w;
  }
  const float sqrt_d = sqrtf((float)D);
  const float mn2 = max_norm * max_norm;
  const float inv_escale = 1.f / escale;
  const int stride = gridDim.x * rpb;
  const int iters = (n + stride - 1) / stride;   // the same trip count for every warp of the grid: the loop holds CTA barriers
  for (int it = 0; it < iters; ++it) {
    const int slot = warp * 4 + grp;             // this row's slot in the CTA's lists
    const int row = it * stride + blockIdx.x * rpb + slot;
    const bool active = row < n;
    float4 x[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float x2 = 0.f;
    float bd = INFINITY;
    int bi = 0x7fffffff;
    bool ambiguous = false;
    int n_splits = 0;
    const float4* prow = partial;
    float4 p0 = make_float4(INFINITY, 0.f, INFINITY, 0.f), p1 = p0;
    float window = -INFINITY;
    if (active) {
      const float4* xr = reinterpret_cast<const float4*>(rows + (size_t)row * D);
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int c = sub + 8 * i;
        x[i] = c < d4 ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      n_splits = lay.stride;
      if (lay.mode == 2) {
        const unsigned a = (unsigned)(row >> 8) * (unsigned)lay.T;
        n_splits = lay.parts * (int)((a + (unsigned)lay.T - 1u) / W - a / W + 1u);
      }
      prow = partial + (size_t)row * lay.stride * RVQ_SLOT_F4;
      if (sub < n_splits) { p0 = prow[sub * RVQ_SLOT_F4]; p1 = prow[sub * RVQ_SLOT_F4 + 1]; }   // entries sub, (sub + 8, ... below)
    }
    const float rs = active ? rscale[row] : 1.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      x2 = fmaf(x[i].x, x[i].x, x2); x2 = fmaf(x[i].y, x[i].y, x2);
      x2 = fmaf(x[i].z, x[i].z, x2); x2 = fmaf(x[i].w, x[i].w, x2);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) x2 += __shfl_xor_sync(0xffffffffu, x2, o);
    // entries 8 .. 31 of a row with more than eight (wide partial layouts only): their best scores join the minimum
    float smin = p0.x;
    for (int e = sub + 8; e < n_splits; e += 8) smin = fminf(smin, prow[e * RVQ_SLOT_F4].x);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
    if (active) {
      const float xn = sqrtf(x2);
      float delta = 0.001953125f * 1.01f * xn * max_norm +
                    5.96e-8f * sqrt_d * (max_norm * __frcp_rn(rs) + xn * inv_escale) + 1e-7f * (x2 + mn2);
      if (c_scale > 0.f) {
        delta += 1e-6f * mn2;
        if (rs * c_scale < 5.9604645e-8f) delta += mn2;
      }
      window = smin + 2.f * delta;
    }
    // passes over the row's entries, eight per pass (lane sub holds entry pass * 8 + sub); one pass unless the layout is wide
    const int n_pass = __reduce_max_sync(0xffffffffu, (n_splits + 7) >> 3);
    for (int pass = 0; pass < n_pass; ++pass) {
      if (pass > 0) {
        p0 = make_float4(INFINITY, 0.f, INFINITY, 0.f); p1 = p0;
        const int e = pass * 8 + sub;
        if (e < n_splits) { p0 = prow[e * RVQ_SLOT_F4]; p1 = prow[e * RVQ_SLOT_F4 + 1]; }
      }
      // chunks of this lane's entry inside the window (ascending list -> a prefix); a list whose LAST kept chunk is inside may be incomplete
      const int mine = (p0.x <= window) + (p0.z <= window) + (p1.x <= window) + (p1.z <= window);
      ambiguous = ambiguous || ((__ballot_sync(0xffffffffu, p1.z <= window) & gmask) != 0u);
    }
    // (an ambiguous row is decided by the full scan alone; a NaN row has no chunk inside the window and falls through to code 0)
    for (int pass = 0; pass < n_pass; ++pass) {
      if (n_pass > 1) {
        p0 = make_float4(INFINITY, 0.f, INFINITY, 0.f); p1 = p0;
        const int e = pass * 8 + sub;
        if (e < n_splits) { p0 = prow[e * RVQ_SLOT_F4]; p1 = prow[e * RVQ_SLOT_F4 + 1]; }
      }
      const int mine = ambiguous ? 0 : (p0.x <= window) + (p0.z <= window) + (p1.x <= window) + (p1.z <= window);
      unsigned live = (__ballot_sync(0xffffffffu, mine > 0) & gmask) >> (grp * 8);   // this row's entries with candidates
      int which = 0;   // next chunk of the current owner entry
      while (__any_sync(0xffffffffu, live != 0u)) {
        const bool busy = live != 0u;
        const int owner = grp * 8 + (busy ? __ffs(live) - 1 : 0);
        const int cnt = __shfl_sync(0xffffffffu, mine, owner);
        const float y0 = __shfl_sync(0xffffffffu, p0.y, owner), y1 = __shfl_sync(0xffffffffu, p0.w, owner);
        const float y2 = __shfl_sync(0xffffffffu, p1.y, owner), y3 = __shfl_sync(0xffffffffu, p1.w, owner);
        const int code0 = __float_as_int(which == 0 ? y0 : which == 1 ? y1 : which == 2 ? y2 : y3) * CHUNK;
        float d[CHUNK], en[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) { d[j] = 0.f; en[j] = 0.f; }
        if (busy) {
#pragma unroll
          for (int j = 0; j < CHUNK; ++j) en[j] = __ldg(e2 + code0 + j);
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const int c = sub + 8 * i;
            if (c < d4) {
              float4 e[CHUNK];
#pragma unroll
              for (int j = 0; j < CHUNK; ++j) e[j] = __ldg(reinterpret_cast<const float4*>(cb + (size_t)(code0 + j) * D) + c);
#pragma unroll
              for (int j = 0; j < CHUNK; ++j) {
                d[j] = fmaf(x[i].x, e[j].x, d[j]); d[j] = fmaf(x[i].y, e[j].y, d[j]);
                d[j] = fmaf(x[i].z, e[j].z, d[j]); d[j] = fmaf(x[i].w, e[j].w, d[j]);
              }
            }
          }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1)
#pragma unroll
          for (int j = 0; j < CHUNK; ++j) d[j] += __shfl_xor_sync(0xffffffffu, d[j], o);
        if (busy) {
#pragma unroll
          for (int j = 0; j < CHUNK; ++j) {
            const float q = sqrtf(fmaxf(__fadd_rn(__fadd_rn(x2, en[j]), __fmul_rn(d[j], -2.f)), 0.f));
            if (q < bd || (q == bd && code0 + j < bi)) { bd = q; bi = code0 + j; }
          }
          if (++which >= cnt) { which = 0; live &= live - 1; }
        }
      }
    }
    // ---- ambiguous rows of this CTA: exact fp32 scan of the whole codebook by all warps (rare; the barrier is the common-case cost) ----
    if (__syncthreads_or(ambiguous ? 1 : 0)) {
      if (sub == 0) { s_row[slot] = ambiguous ? row : -1; s_best[slot] = ~0ull; }
      __syncthreads();
      for (int w = 0; w < rpb; ++w) {
        const int r = s_row[w];
        if (r < 0) continue;   // (uniform: shared memory)
        for (int k = threadIdx.x; k < D; k += blockDim.x) s_x[k] = rows[(size_t)r * D + k];
        if (slot == w && sub == 0) s_x2 = x2;   // the row's owner has |x|^2 (the same value for every code, whatever its rounding)
        __syncthreads();
        const float xx = s_x2;
        float sd = INFINITY;
        int si = 0x7fffffff;
        for (int code = threadIdx.x; code < K; code += blockDim.x) {   // ascending per thread: strict < keeps the lowest index
          const float4* er = reinterpret_cast<const float4*>(cb + (size_t)code * D);
          float acc = 0.f;
          for (int k4 = 0; k4 < d4; ++k4) {
            const float4 e = __ldg(er + k4);
            const float4 xv = *reinterpret_cast<const float4*>(s_x + 4 * k4);
            acc = fmaf(xv.x, e.x, acc); acc = fmaf(xv.y, e.y, acc); acc = fmaf(xv.z, e.z, acc); acc = fmaf(xv.w, e.w, acc);
          }
          const float q = sqrtf(fmaxf(__fadd_rn(__fadd_rn(xx, __ldg(e2 + code)), __fmul_rn(acc, -2.f)), 0.f));
          if (q < sd) { sd = q; si = code; }
        }
        unsigned long long key = si == 0x7fffffff ? ~0ull : (((unsigned long long)__float_as_uint(sd) << 32) | (unsigned int)si);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
          key = other < key ? other : key;
        }
        if (lane == 0) atomicMin(&s_best[w], key);
        __syncthreads();   // s_best[w] complete; s_x free for the next row
      }
      if (ambiguous) {
        const unsigned long long key = s_best[slot];
        bd = __uint_as_float((unsigned int)(key >> 32));
        bi = key == ~0ull ? 0x7fffffff : (int)(unsigned int)(key & 0xffffffffull);
        if (sub == 0) atomicAdd(stats + 1, 1);
      }
    }
    // ---- decided: code id, residual update, the next level's row scale and image (groups without a row take part in the shuffles) ----
    if (bi == 0x7fffffff) bi = 0;   // NaN row: no comparison succeeded (torch.argmax of an all-NaN row is 0 as well)
    if (active && sub == 0) codes[(size_t)row * Q + level] = bi;
    if (update) {
      float vmax = 0.f;
      if (active) {
        const float4* er = reinterpret_cast<const float4*>(cb + (size_t)bi * D);
        float4* dst = reinterpret_cast<float4*>(rows + (size_t)row * D);
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int c = sub + 8 * i;
          if (c < d4) {
            const float4 e = __ldg(er + c);
            x[i].x = __fsub_rn(x[i].x, e.x); x[i].y = __fsub_rn(x[i].y, e.y);
            x[i].z = __fsub_rn(x[i].z, e.z); x[i].w = __fsub_rn(x[i].w, e.w);
            dst[c] = x[i];
            vmax = fmaxf(vmax, fmaxf(fmaxf(fabsf(x[i].x), fabsf(x[i].y)), fmaxf(fabsf(x[i].z), fabsf(x[i].w))));
          }
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      if (active) {
        float sc = rvq_pow2_scale(vmax);
        if (row_image) sc = fminf(sc, 32768.f / c_scale);   // (see rvq_emit_row_scale_and_image)
        if (sub == 0 && rscale) rscale[row] = sc;
        if (row_image) {
          uint8_t* tile = row_image + (size_t)(row >> 7) * rvq_image_tile_bytes(D);
          const int r = row & 127;
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const int c = sub + 8 * i;
            if (c < d4) {
              const int k = 4 * c;
              const __half2 h0 = __floats2half2_rn(x[i].x * sc, x[i].y * sc), h1 = __floats2half2_rn(x[i].z * sc, x[i].w * sc);
              const uint32_t off = (uint32_t)(k >> 6) * (128u * 128u) + (uint32_t)r * 128u + ((uint32_t)(((k >> 3) ^ r) & 7) << 4) + (uint32_t)(k & 7) * 2u;
              *reinterpret_cast<uint2*>(tile + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            }
          }
          if (sub < 2) {
            const __half ch = __float2half_rn(sc * c_scale);
            const __half2 cc = __halves2half2(ch, ch);
            uint8_t* ex = tile + (size_t)128 * D * 2 + (size_t)(r >> 3) * RVQ_EXTRA_SBO + (size_t)(r & 7) * 16 + (size_t)sub * 128;
            *reinterpret_cast<uint4*>(ex) = sub == 0 ? make_uint4(*reinterpret_cast<const uint32_t*>(&cc), 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
    }
  }
  pdl_launch_dependents();
}

// fp32 codebook * escale -> fp16 image in stage order: [128-code tile][64-dim chunk][128 rows x 128 B, SWIZZLE_128B].
__global__ void __launch_bounds__(256) build_image_kernel(const float* __restrict__ cb, const float* __restrict__ e2, int K, int D, float escale,
                                                          float kappa, uint8_t* __restrict__ image) {
  const size_t tile_bytes = rvq_image_tile_bytes(D);
  const long long total = (long long)K * (D / 8);  // 16-byte chunks
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int code = (int)(i / (D / 8));
    const int k8 = (int)(i % (D / 8));
    const float4* src = reinterpret_cast<const float4*>(cb + (size_t)code * D + k8 * 8);
    const float4 v0 = src[0], v1 = src[1];
    __half2 p0 = __floats2half2_rn(v0.x * escale, v0.y * escale), p1 = __floats2half2_rn(v0.z * escale, v0.w * escale);
    __half2 p2 = __floats2half2_rn(v1.x * escale, v1.y * escale), p3 = __floats2half2_rn(v1.z * escale, v1.w * escale);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
    pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
    const int tile = code / N_TILE, r = code % N_TILE;
    const int kc = (k8 * 8) / KC, kin = (k8 * 8) % KC;
    const size_t off = (size_t)tile * tile_bytes + (size_t)kc * STAGE_BYTES + tc::sw128_offset(r, kin);
    *reinterpret_cast<uint4*>(image + off) = pk;
    if (k8 == 0) {
      // extra K block: columns 0, 1 = fp16 pair (hi, lo) of -kappa * |e|^2, columns 2..15 zero
      const float t = -kappa * e2[code];
      const __half hi = __float2half_rn(t);
      const __half lo = __float2half_rn(t - __half2float(hi));
      const __half2 hl = __halves2half2(hi, lo);
      uint4 first = make_uint4(*reinterpret_cast<const uint32_t*>(&hl), 0u, 0u, 0u);
      uint8_t* ex = image + (size_t)tile * tile_bytes + (size_t)(D / KC) * STAGE_BYTES + (size_t)(r >> 3) * RVQ_EXTRA_SBO + (size_t)(r & 7) * 16;
      *reinterpret_cast<uint4*>(ex) = first;                         // K core matrix 0 (columns 0..7)
      *reinterpret_cast<uint4*>(ex + 128) = make_uint4(0u, 0u, 0u, 0u);  // K core matrix 1 (columns 8..15)
    }
  }
}

template <int D, int DBG>
int launch_search(const SearchArgs& a, int grid, cudaStream_t stream) {
  MT_ENSURE_SMEM((rvq_tc_search_kernel<D, DBG>), Smem<D>::TOTAL);
  MT_LAUNCH_PDL((rvq_tc_search_kernel<D, DBG>), grid, NUM_THREADS, Smem<D>::TOTAL, stream, a);
  return MESHTOK_OK;
}

}  // namespace

int rvq_tc_plan(int max_rows, int K, int D, RvqTcPlan* plan) {
  memset(plan, 0, sizeof(*plan));
  if (K % N_TILE != 0 || (D != 64 && D != 128 && D != 192)) {
    set_error("tcgen05 RVQ search supports codebook_size %% 128 == 0 and dim in {64,128,192} (got K=%d, D=%d); "
              "use the exact path (rvq_exact) for other shapes", K, D);
    return MESHTOK_ERR_UNSUPPORTED;
  }
  const int tiles = K / N_TILE;
  const int m_tiles = max(1, ceil_div(max_rows, M_TILE));
  int best_ns = 0;
  double best_eff = -1.0;
  for (int ns = 1; ns <= tiles && ns <= 32; ns *= 2) {   // the re-score kernel reads one entry per lane
    if (tiles % ns != 0) break;
    if ((tiles / ns) * N_TILE > MAX_SPLIT_CODES) continue;
    const long long units = (long long)m_tiles * ns;
    const long long waves = (units + 147) / 148;
    const double eff = (double)units / (double)(waves * 148);
    if (eff > best_eff + 0.02) { best_eff = eff; best_ns = ns; }
    if (units >= 148 * 4) break;
  }
  if (best_ns == 0) {
    set_error("tcgen05 RVQ search: cannot split %d code tiles", tiles);
    return MESHTOK_ERR_UNSUPPORTED;
  }
  plan->m_tiles = m_tiles;
  plan->n_splits = best_ns;
  plan->tiles_per_split = tiles / best_ns;
  plan->units = m_tiles * best_ns;
  plan->grid = min(plan->units, 148);
  plan->partial_bytes = align_up((size_t)max(max_rows, 1) * best_ns * RVQ_SLOT_F4 * sizeof(float4), 256);
  return MESHTOK_OK;
}

int launch_rvq_tc_search(const float* rows, const float* rscale, int D, const void* codebook_tiles, float escale, const float* e2,
                         int K, const int* n_rows_dev, int max_rows, const RvqTcPlan& plan, float4* partial, cudaStream_t stream) {
  if (max_rows == 0) return MESHTOK_OK;
  SearchArgs a;
  a.rows = rows; a.rscale = rscale; a.escale = escale; a.image = static_cast<const uint8_t*>(codebook_tiles); a.e2 = e2; a.n_rows_dev = n_rows_dev;
  a.debug = 0;
  a.max_rows = max_rows; a.K = K; a.n_splits = plan.n_splits; a.tiles_per_split = plan.tiles_per_split; a.partial = partial;
  switch (D) {
    case 64: return launch_search<64, 0>(a, plan.grid, stream);
    case 128: return launch_search<128, 0>(a, plan.grid, stream);
    case 192: {
#ifdef MESHTOK_EXPERIMENTS   // timing experiments that produce WRONG codes; never in a product build
      static int debug = -1;
      if (debug < 0) { const char* e = getenv("MESHTOK_RVQ_DEBUG"); debug = e ? atoi(e) : 0; }
      a.debug = debug;
      switch (a.debug) {
        case 1: return launch_search<192, 1>(a, plan.grid, stream);
        case 2: return launch_search<192, 2>(a, plan.grid, stream);
        case 3: return launch_search<192, 3>(a, plan.grid, stream);
      }
#endif
      return launch_search<192, 0>(a, plan.grid, stream);
    }
  }
  set_error("tcgen05 RVQ search: unsupported dim %d", D);
  return MESHTOK_ERR_UNSUPPORTED;
}

int launch_rvq_rescore(float* rows, float* rscale, int D, const float* codebook, const float* e2, int K, float max_norm,
                       float escale, float c_scale, const float4* partial, const RvqPartialLayout& layout, const int* n_rows_dev, int max_rows,
                       int32_t* codes, int Q, int level, bool update, void* row_image, int* stats, cudaStream_t stream) {
  if (max_rows == 0) return MESHTOK_OK;
  MT_CHECK_ARG(layout.stride <= 32, "rvq re-score: at most 32 candidate entries per row (got %d)", layout.stride);
  MT_CHECK_ARG(layout.mode != 2 || (long long)ceil_div(max_rows, 256) * layout.T < (1ll << 31),
               "rvq re-score: %d rows x %d code tiles exceed the 32-bit step range", max_rows, layout.T);
  MT_CHECK_ARG(D % 4 == 0 && D <= 192, "rvq re-score: dim %d must be a multiple of 4 and <= 192", D);
  MT_CHECK_ARG(row_image == nullptr || (D % 64 == 0 && c_scale > 0.f), "rvq: a row image needs dim %% 64 == 0 and c_scale > 0");
  // four rows per warp (rvq_rescore4_kernel) unless MESHTOK_RESCORE4=0 asks for the one-row-per-warp kernel (A/B runs)
  static int four = -1;
  if (four < 0) { const char* e = getenv("MESHTOK_RESCORE4"); four = (e && atoi(e) == 0) ? 0 : 1; }
  if (four == 1 && D % 4 == 0) {
    const int blocks = max(1, min(ceil_div(max_rows, 32), 148 * 8));
    // (held to 80 registers for three CTAs per SM - 40 bytes of spills - the kernel measured 58.5 against 43.8 us per step at C2: it stays at 128)
    MT_LAUNCH_PDL((rvq_rescore4_kernel<6, 2>), blocks, 256, 0, stream, rows, D, codebook, e2, K, max_norm, escale, c_scale, rscale, partial, layout,
                  n_rows_dev, max_rows, codes, Q, level, update ? 1 : 0, static_cast<uint8_t*>(row_image), stats);
    return MESHTOK_OK;
  }
  const int blocks = max(1, min(ceil_div(max_rows, 8), 148 * 16));
  MT_LAUNCH_PDL(rvq_rescore_kernel<3>, blocks, 256, 0, stream, rows, D, codebook, e2, K, max_norm, escale, c_scale, rscale, partial, layout,
                n_rows_dev, max_rows, codes, Q, level, update ? 1 : 0, static_cast<uint8_t*>(row_image), stats);
  return MESHTOK_OK;
}

size_t rvq_codebook_image_bytes(int K, int D) { return (size_t)ceil_div(K, N_TILE) * rvq_image_tile_bytes(D); }

int launch_rvq_build_image(const float* codebook, const float* e2, int K, int D, float escale, float max_norm, void* image, cudaStream_t stream) {
  MT_CHECK_ARG(K % N_TILE == 0 && D % KC == 0, "codebook image needs codebook_size %% 128 == 0 and dim %% 64 == 0 (got %d, %d)", K, D);
  const long long total = (long long)K * (D / 8);
  const long long want = (total + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148 * 16 ? want : 148 * 16);
  build_image_kernel<<<blocks, 256, 0, stream>>>(codebook, e2, K, D, escale, rvq_e2_kappa(max_norm), static_cast<uint8_t*>(image));
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

}  // namespace meshtok
