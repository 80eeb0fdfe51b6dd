// K8a, CTA-pair version: residual-VQ nearest-codebook screening with tcgen05.mma.cta_group::2.
//
// Same contract as rvq_tc.cu (fp16 screening GEMM  s~[r, j] = |e_j|^2 - 2 <fp16(r), fp16(e_j)>  with fp32 accumulation
// in TMEM, two best 8-code chunks kept per (row, partial slot), exact fp32 decision by rvq_rescore_kernel), replacing
// the EuclideanCodebook distance + argmax of vector_quantize_pytorch.ResidualVQ called at meshgpt_pytorch.py:842.
//
// Why a second version: ncu on the single-CTA kernel (profiles/r01d_*) showed the tensor pipe active only 48 % of the
// time with neither the epilogue nor L2 the limiter - an M=128 x N=128 MMA needs 8 KB of shared-memory operands per 64
// cycles, the full 128 B/cycle of one SM, before the bulk copies and the epilogue touch shared memory at all.  Here two
// CTAs of a cluster (one TPC) issue ONE M=256 x N=256 MMA per K step: every CTA stages its own 128 residual rows and
// its own 128 codes of the 256-code tile, i.e. 8 KB per 128 cycles = half the operand traffic per SM, while the
// codebook bytes pulled from L2 per row stay the same as before.
//   * row tile: 128 rows per CTA, fp16 pre-scaled image written by the residual prep / update kernels, one 48 KB bulk
//     copy (no register staging, no conversion here);
//   * codebook: 8-stage ring of 16 KB stages (128 codes x 64 dims) per CTA from the same image rvq_tc.cu uses
//     (CTA rank r takes image tile 2t + r);
//   * accumulators: 2 buffers x 256 columns of TMEM in each CTA (lanes = the CTA's 128 rows);
//   * work split: the (row-pair-tile x code-pair-tile) steps of the whole problem are cut into equal contiguous ranges,
//     one per cluster ("stream-K"), so the load is balanced to one step whatever the row count; a row tile cut by a
//     range boundary simply produces one more partial slot;
//   * synchronisation: the leader CTA (cluster rank 0) issues all MMAs.  Its full barriers count two arrivals: its own
//     producer's expect_tx and a remote arrive forwarded by the peer CTA once the peer's copy has landed; stage /
//     row-tile / accumulator releases are tcgen05.commit multicasts to both CTAs; the peer's epilogue warps release
//     the accumulator with remote arrives.
// Warp roles (384 threads per CTA): w0 producer, w1 MMA issuer (leader only), w2 TMEM allocator, w3 completion
// forwarder (peer only), w4..w11 epilogue (TMEM lane quarter = warp % 4, column half = (warp - 4) / 4).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"
#include "rvq_screen.cuh"
#include "tc_common.cuh"

namespace meshtok {

namespace {

using namespace tc;

constexpr int M_CTA = 128;         // residual rows per CTA
constexpr int M_PAIR = 256;        // rows per cluster step
constexpr int N_CTA = 128;         // codes staged per CTA and stage (one tile of the codebook image)
constexpr int N_PAIR = 256;        // codes per MMA
constexpr int KC = 64;             // K elements per pipeline stage
constexpr int STAGES = 9;
constexpr int STAGE_BYTES = N_CTA * KC * 2;    // 16 KB
constexpr int CHUNK = RVQ_CHUNK;
constexpr int EPI_WARPS = 16;       // 4 per TMEM lane quarter, each takes COLS_PER_WARP of the 256 columns
constexpr int NUM_THREADS = (4 + EPI_WARPS) * 32;
constexpr int CSPLIT = EPI_WARPS / 4;  // column parts = partial entries per (row, cluster segment)
constexpr int COLS_PER_WARP = N_PAIR / CSPLIT;
constexpr uint32_t SBO_B = (KC / 8) * 128;     // 1024
constexpr int W_MIN = 16;          // fewest steps a cluster takes; max(W_MIN, T / 4) bounds the partial slots per row to 6

template <int D>
struct Smem2 {
  static constexpr uint32_t A_MAIN = M_CTA * D * 2;                      // D/64 SWIZZLE_128B blocks of 16 KB
  static constexpr uint32_t A_BYTES = A_MAIN + RVQ_EXTRA_BYTES;          // + the extra K block (per-row constant)
  static constexpr uint32_t B_OFF = A_BYTES;
  // extra K blocks (|e|^2 columns) in flight, one per code tile, released with the tile's last stage: the producer runs at
  // most ceil(STAGES / stages per tile) tiles ahead of the MMAs, so that many + 1 slots are live; one more for margin
  static constexpr int EX_RING = (STAGES + D / KC - 1) / (D / KC) + 2;
  static constexpr uint32_t X_OFF = B_OFF + STAGES * STAGE_BYTES;         // ring of EX_RING extra K blocks of the codebook
  static constexpr uint32_t BAR_OFF = X_OFF + EX_RING * RVQ_EXTRA_BYTES;
  static constexpr uint32_t TOTAL = BAR_OFF + 512;
};
static_assert(Smem2<192>::TOTAL <= 227 * 1024 && Smem2<128>::TOTAL <= 227 * 1024 && Smem2<64>::TOTAL <= 227 * 1024, "shared memory budget");

struct Search2Args {
  const uint8_t* rimg;     // fp16 image of the scaled residuals per 128-row tile (rvq_update_kernel): D/64 blocks + extra K block
  const uint8_t* image;    // fp16 codebook image per 128-code tile: D/64 stages of 16 KB + extra K block (rvq_screen.cuh)
  const float* rscale;     // (R) power-of-two row scales baked into rimg
  float escale;            // power-of-two codebook scale baked into the image
  const int* n_rows_dev;
  int max_rows, T, max_slots, w_min;   // T = K / 256 code-pair-tiles
  float4* partial;         // (R, CSPLIT * max_slots) entries of RVQ_SLOT_F4 float4 (rvq_screen.cuh)
  int pdl;                 // launched with programmatic stream serialization: barriers / TMEM are set up before the wait
};

__host__ __device__ __forceinline__ int w_min_of(int T) { return T / 4 > W_MIN ? T / 4 : W_MIN; }
__device__ __forceinline__ int steps_per_cluster(long long S, int n_clusters, int w_min) {
  const long long w = (S + n_clusters - 1) / n_clusters;
  return (int)(w < w_min ? w_min : w);
}

template <int D, int DBG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) rvq_tc2_search_kernel(Search2Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using L = Smem2<D>;
  constexpr int KCH = D / KC;
  uint8_t* sA = smem;
  uint8_t* sB = smem + L::B_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* b_full = bars;                   // [STAGES]
  uint64_t* b_empty = bars + STAGES;         // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;    // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]  (the leader's are the ones waited on)
  uint64_t* a_full = acc_empty + 2;
  uint64_t* a_empty = a_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 1);
  uint8_t* sX = smem + L::X_OFF;

  const int warp = warp_id(), lane = lane_id();
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  // earlier kernels' results: without PDL the row count is read first, so that its load travels during the set-up; with PDL the
  // set-up overlaps the previous kernel's tail and everything an earlier kernel wrote is read after the wait below
  int n_rows = a.max_rows;
  if (!a.pdl) n_rows = min(__ldg(a.n_rows_dev), a.max_rows);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&b_full[i], leader ? 2 : 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * EPI_WARPS); }
    mbar_init(a_full, leader ? 2 : 1);
    mbar_init(a_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers and TMEM exist before anything crosses the pair
  tcgen05_fence_after();
  if (a.pdl) {
    pdl_wait();
    n_rows = min(*reinterpret_cast<const volatile int*>(a.n_rows_dev), a.max_rows);
  }
  const uint32_t tmem_base = *tmem_slot;
  const int G = (n_rows + M_PAIR - 1) / M_PAIR;
  const int T = a.T;
  const long long S = (long long)G * T;
  const int n_clusters = gridDim.x >> 1;
  const int cid = blockIdx.x >> 1;
  const int W = steps_per_cluster(S, n_clusters, a.w_min);
  const long long lo = min(S, (long long)cid * W), hi = min(S, lo + W);

  if (warp == 0) {
    // ===== producer (both CTAs; warp-converged, one elected lane issues): own row tile + own half of every stage =====
    uint32_t it = 0, a_cnt = 0, tile_it = 0;
    int prev_g = -1;
    for (long long step = lo; step < hi; ++step, ++tile_it) {
      const int g = (int)(step / T), t = (int)(step % T);
      if (g != prev_g) {
        if (a_cnt > 0) mbar_wait(a_empty, (a_cnt - 1) & 1);   // MMAs on the previous row tile are complete
        if (elect_one()) {
          mbar_arrive_expect_tx(a_full, L::A_BYTES);
          bulk_g2s(sA, a.rimg + (size_t)(2 * g + (int)rank) * rvq_image_tile_bytes(D), L::A_BYTES, a_full);
        }
        __syncwarp();
        ++a_cnt;
        prev_g = g;
      }
      const uint8_t* src = a.image + (size_t)(2 * t + (int)rank) * rvq_image_tile_bytes(D);
      for (int kc = 0; kc < KCH; ++kc, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&b_empty[s], ph ^ 1);
        if (elect_one()) {
          // the tile's last stage also brings its extra K block (|e|^2 columns).  Its ring slot was used EX_RING tiles ago (see Smem2),
          // and this stage slot only became free after that tile's MMAs - the extra one included - had completed.
          const bool last = kc == KCH - 1;
          if (DBG == 2 && it >= (uint32_t)STAGES) {   // experiment: no codebook traffic after the ring's first fill (the MMAs re-read stale stages)
            mbar_arrive(&b_full[s]);
          } else {
          mbar_arrive_expect_tx(&b_full[s], STAGE_BYTES + (last ? RVQ_EXTRA_BYTES : 0));
          bulk_g2s(sB + s * STAGE_BYTES, src + (size_t)kc * STAGE_BYTES, STAGE_BYTES, &b_full[s]);
          if (last) bulk_g2s(sX + (tile_it % L::EX_RING) * RVQ_EXTRA_BYTES, src + (size_t)KCH * STAGE_BYTES, RVQ_EXTRA_BYTES, &b_full[s]);
          }
        }
        __syncwarp();
      }
    }
    // drain: every multicast commit aimed at this CTA has landed before the CTA may exit
    const uint32_t pending = it < (uint32_t)STAGES ? it : (uint32_t)STAGES;
    for (uint32_t i = 0; i < pending; ++i) {
      const uint32_t j = it - 1 - i;
      mbar_wait(&b_empty[j % STAGES], (j / STAGES) & 1);
    }
    if (a_cnt > 0) mbar_wait(a_empty, (a_cnt - 1) & 1);
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only; warp-converged, one elected lane issues): M=256 x N=256 x K=16 per MMA =====
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16_f32(M_PAIR, N_PAIR, /*bf16=*/false);
      const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
      uint32_t it = 0, acc_it = 0, a_phase = 0;
      int prev_g = -1;
      for (long long step = lo; step < hi; ++step, ++acc_it) {
        const int g = (int)(step / T);
        if (g != prev_g) {
          mbar_wait(a_full, a_phase);
          a_phase ^= 1;
          prev_g = g;
          tcgen05_fence_after();
        }
        const uint32_t buf = acc_it & 1, aph = (acc_it >> 1) & 1;
        mbar_wait(&acc_empty[buf], aph ^ 1);
        tcgen05_fence_after();
        for (int kc = 0; kc < KCH; ++kc, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&b_full[s], ph);
          tcgen05_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k16 = 0; k16 < KC / 16; ++k16) {
              const uint64_t da = make_smem_desc_sw128(sA_addr + kc * STAGE_BYTES + k16 * 32);
              const uint64_t db = make_smem_desc_sw128(sB_addr + s * STAGE_BYTES + k16 * 32);
              umma_f16_pair(tmem_base + buf * N_PAIR, da, db, idesc, (kc | k16) != 0 ? 1u : 0u);
            }
            if (kc == KCH - 1 && DBG != 3) {   // extra K block: + c_row * (hi, lo)(-kappa |e|^2)  ->  acc = -(rscale * escale / 2) * s~   (DBG 3: experiment without it)
              const uint64_t da = make_smem_desc(sA_addr + L::A_MAIN, 128, RVQ_EXTRA_SBO);
              const uint64_t db = make_smem_desc(smem_u32(sX) + (acc_it % L::EX_RING) * RVQ_EXTRA_BYTES, 128, RVQ_EXTRA_SBO);
              umma_f16_pair(tmem_base + buf * N_PAIR, da, db, idesc, 1u);
            }
            umma_commit_pair_mc(&b_empty[s], 3);
          }
          __syncwarp();
        }
        const int next_g = (step + 1 < hi) ? (int)((step + 1) / T) : -1;
        if (elect_one()) {
          umma_commit_pair_mc(&acc_full[buf], 3);
          if (next_g != g) umma_commit_pair_mc(a_empty, 3);
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ===== completion forwarder (peer CTA only): tells the leader that this CTA's copies have landed.  Lane s follows
    // stage s, lane STAGES follows the row tile, so the remote arrives of different stages overlap. =====
    if (!leader) {
      const uint32_t total_it = (uint32_t)(hi - lo) * KCH;
      const uint32_t n_a = hi > lo ? (uint32_t)((hi - 1) / T - lo / T + 1) : 0u;
      if (lane < STAGES) {
        const uint32_t remote = mapa_shared(smem_u32(&b_full[lane]), 0);
        uint32_t ph = 0;
        for (uint32_t it = lane; it < total_it; it += STAGES, ph ^= 1) {
          mbar_wait(&b_full[lane], ph);
          mbar_arrive_cluster_relaxed(remote);   // see linear_tc.cu: no data travels with this arrival, a release fence per stage is wasted time
        }
      } else if (lane == STAGES) {
        const uint32_t remote = mapa_shared(smem_u32(a_full), 0);
        for (uint32_t i = 0; i < n_a; ++i) {
          mbar_wait(a_full, i & 1);
          mbar_arrive_cluster_relaxed(remote);
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): running two best 8-code chunks per row and column half =====
    const int ew = warp - 4;
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int chalf = ew >> 2;               // column part: columns [chalf * COLS_PER_WARP, +COLS_PER_WARP) of the 256-code tile
    const int row_in_cta = q * 32 + lane;
    const uint32_t acc_empty_remote0 = mapa_shared(smem_u32(&acc_empty[0]), 0);
    const uint32_t acc_empty_remote1 = mapa_shared(smem_u32(&acc_empty[1]), 0);
    uint32_t acc_it = 0;
    int prev_g = -1;
    float cmul = 0.f, icm = 0.f, thr = -INFINITY;
    RvqTop top;
    top.reset();
    long long my_row = 0;
    auto flush = [&](int g) {
      if (my_row < n_rows) {
        const int slot = CSPLIT * (cid - (int)(((long long)g * T) / W)) + chalf;
        top.store(a.partial + ((size_t)my_row * (CSPLIT * a.max_slots) + slot) * RVQ_SLOT_F4);
      }
    };
    for (long long step = lo; step < hi; ++step, ++acc_it) {
      const int g = (int)(step / T), t = (int)(step % T);
      if (g != prev_g) {
        if (prev_g >= 0) flush(prev_g);
        prev_g = g;
        my_row = (long long)g * M_PAIR + (long long)rank * M_CTA + row_in_cta;
        // acc = rscale * escale * (<r, e> - |e|^2 / 2)  ->  s~ = |e|^2 - 2 <r, e> = acc * cmul
        cmul = -2.f / ((my_row < n_rows ? __ldg(a.rscale + my_row) : 1.f) * a.escale);
        icm = 1.f / cmul;
        top.reset();
        thr = -INFINITY;
      }
      const uint32_t buf = acc_it & 1, aph = (acc_it >> 1) & 1;
      mbar_wait(&acc_full[buf], aph);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * N_PAIR + chalf * COLS_PER_WARP;
      const int code0 = t * N_PAIR + chalf * COLS_PER_WARP;
      uint32_t va[32], vb[32];
      if (DBG == 1) {   // timing experiment: nobody reads the accumulators
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&acc_empty[buf]);
          else mbar_arrive_cluster_relaxed(buf ? acc_empty_remote1 : acc_empty_remote0);
        }
        continue;
      }
      tmem_ld32(taddr, va);
#pragma unroll
      for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
        tmem_ld_wait();
        uint32_t(&v)[32] = (c & 1) ? vb : va;
        if (c + 1 < COLS_PER_WARP / 32) tmem_ld32(taddr + (c + 1) * 32, (c & 1) ? va : vb);
        rvq_screen32_max(v, cmul, icm, (code0 + c * 32) / CHUNK, top, thr);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {   // the TMEM reads are complete (wait::ld); no data is handed over through memory, hence relaxed
        if (leader) mbar_arrive(&acc_empty[buf]);
        else mbar_arrive_cluster_relaxed(buf ? acc_empty_remote1 : acc_empty_remote0);
      }
    }
    if (prev_g >= 0) flush(prev_g);
  }

  pdl_launch_dependents();   // this CTA's work is done: let the next kernel's CTAs start their prologue
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();   // nothing of the pair is still in flight towards either CTA
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int D, int DBG>
int launch_search2(const Search2Args& a, int grid, cudaStream_t stream) {
  MT_ENSURE_SMEM((rvq_tc2_search_kernel<D, DBG>), Smem2<D>::TOTAL);
  MT_LAUNCH_PDL((rvq_tc2_search_kernel<D, DBG>), grid, NUM_THREADS, Smem2<D>::TOTAL, stream, a);
  return MESHTOK_OK;
}

int sm_pairs() {
  static int pairs = 0;
  if (pairs == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 2) sms = 148;
    pairs = sms / 2;
  }
  return pairs;
}

}  // namespace

int rvq_tc2_plan(int max_rows, int K, int D, RvqTc2Plan* plan) {
  memset(plan, 0, sizeof(*plan));
  if (K % N_PAIR != 0 || (D != 64 && D != 128 && D != 192)) {
    set_error("tcgen05 CTA-pair RVQ search supports codebook_size %% 256 == 0 and dim in {64,128,192} (got K=%d, D=%d)", K, D);
    return MESHTOK_ERR_UNSUPPORTED;
  }
  plan->T = K / N_PAIR;
  plan->n_clusters = sm_pairs();
  plan->max_slots = (plan->T - 1) / w_min_of(plan->T) + 2;
  const size_t rows_padded = (size_t)ceil_div(max(max_rows, 1), M_PAIR) * M_PAIR;
  plan->partial_bytes = align_up(rows_padded * CSPLIT * plan->max_slots * RVQ_SLOT_F4 * sizeof(float4), 256);
  plan->row_image_bytes = align_up(rows_padded / M_CTA * rvq_image_tile_bytes(D), 256);
  return MESHTOK_OK;
}

int launch_rvq_tc2_search(const void* row_image, const float* rscale, int D, const void* codebook_tiles, float escale,
                          const int* n_rows_dev, int max_rows, const RvqTc2Plan& plan, float4* partial, cudaStream_t stream) {
  if (max_rows == 0) return MESHTOK_OK;
  Search2Args a;
  a.rimg = static_cast<const uint8_t*>(row_image); a.image = static_cast<const uint8_t*>(codebook_tiles); a.rscale = rscale;
  a.escale = escale; a.n_rows_dev = n_rows_dev; a.max_rows = max_rows; a.T = plan.T; a.max_slots = plan.max_slots; a.w_min = w_min_of(plan.T); a.partial = partial;
  a.pdl = pdl_enabled() ? 1 : 0;
  // enough clusters for the worst case, never more than the machine has SM pairs
  const long long steps = (long long)ceil_div(max_rows, M_PAIR) * plan.T;
  const int wm = w_min_of(plan.T);
  const int clusters = (int)min((long long)plan.n_clusters, (steps + wm - 1) / wm);
  const int grid = 2 * max(1, clusters);
  switch (D) {
    case 64: return launch_search2<64, 0>(a, grid, stream);
    case 128: return launch_search2<128, 0>(a, grid, stream);
    case 192: {
#ifdef MESHTOK_EXPERIMENTS   // timing experiments that produce WRONG codes (DBG = 1: nobody reads the accumulators); never in a product build
      static int debug = -1;
      if (debug < 0) { const char* e = getenv("MESHTOK_RVQ_DEBUG"); debug = e ? atoi(e) : 0; }
      if (debug == 1) return launch_search2<192, 1>(a, grid, stream);
      if (debug == 2) return launch_search2<192, 2>(a, grid, stream);
      if (debug == 3) return launch_search2<192, 3>(a, grid, stream);
#endif
      return launch_search2<192, 0>(a, grid, stream);
    }
  }
  set_error("tcgen05 CTA-pair RVQ search: unsupported dim %d", D);
  return MESHTOK_ERR_UNSUPPORTED;
}

// partial-slot enumeration shared with the re-score kernel (must mirror the kernel's work split)
RvqPartialLayout rvq_tc2_partial_layout(const RvqTc2Plan& plan, int max_rows) {
  RvqPartialLayout l;
  l.mode = 2;
  l.stride = CSPLIT * plan.max_slots;
  l.parts = CSPLIT;
  l.T = plan.T;
  const long long steps = (long long)ceil_div(max_rows, M_PAIR) * plan.T;
  const int wm = w_min_of(plan.T);
  l.n_clusters = max(1, (int)min((long long)plan.n_clusters, (steps + wm - 1) / wm));
  l.w_min = wm;
  return l;
}

}  // namespace meshtok
