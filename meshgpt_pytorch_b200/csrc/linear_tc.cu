// Dense linears of the encoder on tcgen05 tensor cores with fp32-class accuracy ("split bf16").
//
// Replaces the cuBLAS SGEMMs behind torch.nn.functional.linear in SAGEConv.lin / lin_l / lin_r
// (torch_geometric, called at meshgpt_pytorch.py:764, :769) and project_dim_codebook (:803).
//   C = act([A1 | A2] W^T + bias),  A (M, K1+K2), W (N, K), C fp32 (M, N)
// Every fp32 operand x is split as x = hi + lo with hi = bf16(x), lo = bf16(x - hi) and the product is
// accumulated in fp32 (TMEM) as  lo.hi + hi.lo + hi.hi  (the lo.lo term, relative 2^-16, is dropped), i.e.
// three bf16 UMMAs per K step: relative error ~1e-5 instead of bf16's 4e-3, at 3/8 of the cost of an
// fp32-exact 3xTF32 scheme.
//
// Both operands arrive as ready-made shared-memory images (tc_common.cuh): W is split and tiled once at load
// time, the activations are written in image form by the kernels that produce them (face_embed, gather_mean,
// row_norm), so a pipeline stage = two cp.async.bulk copies issued by one thread - no register staging, no
// per-chunk conversion in this kernel.  (A first version converted fp32 rows in eight loader warps; it ran at a
// fifth of the MMA-bound time because the loader warps, not the tensor pipe, set the pace.)
//   * unit of work = (128-row tile, pass of NT <= 256 output columns), K walked in 32-wide chunks through a
//     4-stage ring (16 KB of A + up to 32 KB of W per stage);
//   * accumulators are double buffered in TMEM so the epilogue of one unit overlaps the MMAs of the next;
//   * plain epilogue: TMEM -> registers -> per-warp shared-memory transpose -> +bias -> ReLU -> fp32 stores
//     (8 rows x 64 B per instruction);
//   * fused row epilogue (single pass, N <= 256): the F.normalize(p=2) that ends every SAGEConv - and for the first
//     layer SiLU + LayerNorm (meshgpt_pytorch.py:545-548) - is done on the accumulator rows (one TMEM lane = one row per
//     thread, the two column halves exchange their partial sums through shared memory) and the result is written
//     straight into the NEXT linear's activation image: 16 B hi + 16 B lo per 8 columns, 8 lanes = 128 contiguous bytes.
// Warp roles (384 threads): w0 producer, w1 MMA issuer, w2 TMEM allocator, w3 idle, w4-11 epilogue
// (TMEM lane quarter = warp % 4, column half = (warp - 4) / 4).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"
#ifndef MESHTOK_SILU_FAST
#define MESHTOK_SILU_FAST 0
#endif

namespace meshtok {

namespace {

using namespace tc;

constexpr int M_TILE = IMG_ROWS;   // 128
constexpr int KC = IMG_KC;         // 32
constexpr int NT_MAX = 256;
constexpr uint32_t SBO = (KC / 8) * 128;                   // 512
constexpr uint32_t A_PART = IMG_PART;                      // 8 KB (hi or lo)
constexpr int EPI_LD = 20;                                 // padded row length (floats) of the 32 x 16 epilogue staging tiles
constexpr int VEC_MAX_N = 1536;   // widest output (the decoder's 9 x 128 coordinate logits = 1152)

// PAIR = two CTAs of a cluster work on one 256-row tile with tcgen05.mma.cta_group::2: every CTA stages its own 128 rows
// of A and HALF of the W chunk, so the bytes pulled from L2 per MMA cycle drop from (128 + NT) to (128 + NT / 2) rows.
// ncu on the single-CTA kernel (capture r01h, in the git history of profiles/) had the two widest linears at 10.5 TB/s of L2 -> SM traffic, the
// measured cap of the L2 slices, with the tensor pipe waiting.
// GW = gather warps (0 or 4): warps behind the epilogue warps that PRODUCE the A1 operand tiles of a SAGE output linear in shared memory
// (mean over the in-neighbours of fp32 rows, see the gather role in linear_tc_body) instead of the producer's bulk copy of an image block.
template <bool PAIR, int CPARTS_, int GW_ = 0>
struct Cfg {
  static constexpr int MAX_STAGES = 10;
  static constexpr int CPARTS = CPARTS_;                   // column parts of a pass = epilogue warps per TMEM lane quarter
  static constexpr int EPI_WARPS = 4 * CPARTS;
  static constexpr int GW = GW_;
  static constexpr int NUM_THREADS = (4 + EPI_WARPS + GW) * 32;
  static constexpr uint32_t BAR_OFF = 0;
  static constexpr uint32_t EPI_OFF = BAR_OFF + (GW ? 384 : 256);
  static constexpr uint32_t XCH_OFF = EPI_OFF + EPI_WARPS * 32 * EPI_LD * 4;      // [2 unit parities][3 sums][CPARTS][128 rows] floats
  static constexpr uint32_t VEC_OFF = XCH_OFF + 2 * 3 * CPARTS * 128 * 4;         // bias (N <= 1024), gamma, beta (N <= 64) staged once per CTA
  static constexpr uint32_t RING_OFF = (VEC_OFF + (VEC_MAX_N + 2 * 64) * 4 + 127) / 128 * 128;
  static constexpr uint32_t SMEM_TOTAL = 227 * 1024;
  // stage = own A block (hi + lo, 16 KB) + the W rows this CTA stages (hi + lo); as many stages as fit, the smaller the
  // pass the deeper the ring (a CTA pair's stage round trip crosses the cluster twice and needs the depth)
  // (a_slot: bytes reserved for the A block; 16 KB for the plain image, 128 * (128 + 2 halo) for a halo image)
  __host__ __device__ static constexpr uint32_t stage_bytes(int NT, uint32_t a_slot = 2 * A_PART) {
    return a_slot + 2 * (uint32_t)(PAIR ? NT / 2 : NT) * KC * 2;
  }
  __host__ __device__ static constexpr int stages(int NT, uint32_t a_slot = 2 * A_PART) {
    const int n = (int)((SMEM_TOTAL - RING_OFF) / stage_bytes(NT, a_slot));
    return n < MAX_STAGES ? n : MAX_STAGES;
  }
};
static_assert(Cfg<false, 2>::stages(NT_MAX) >= 4 && Cfg<true, 2>::stages(NT_MAX) >= 6, "shared memory budget");

struct LinArgs {
  const uint8_t* a1_image;
  const uint8_t* a2_image;
  int K1, K2;
  const uint8_t* w_image;
  const float* bias;
  float* C;
  int N, NT, passes;
  const int* m_dev;
  int m_max;
  int relu;
  int norm_mode;            // 0 plain, 1 row L2-normalise, 2 L2-normalise + SiLU + LayerNorm(gamma, beta), 3 PixelNorm + SiLU -> halo image
  const float* gamma;
  const float* beta;
  uint8_t* out_image;       // activation image of the output rows (K = N) or null
  // plain mode, deferred row normalisation (layers too wide for the fused epilogue):
  float* ssq_out;           // (M, CPARTS * passes): sum of squares of this warp's part of every output row, or null
  const float* row_ssq;     // (M, row_ssq_parts): parts of |x_row|^2 of the INPUT rows; C = acc / max(|x|, 1e-12) + bias
  int row_ssq_parts;
  // Conv1d over the row sequence (decode path): a1_image is a HALO image (tc::halo_*): every (128-row tile, 32-column chunk) block
  // holds rows -halo .. 128 + halo of the tile with consecutive rows 16 bytes apart, so tap t of the kernel is the SAME block read
  // through a descriptor that starts t * 16 bytes further (LBO = rows * 16, SBO = 128).  K = taps * K1, W columns tap-major.
  int conv, taps, halo;
  // norm_mode 3 (decode path, Block tail :345-351): pass 0 of the launch holds whole rows of width NT; its epilogue writes
  // silu(PixelNorm(row)) = silu(x * sqrt(NT) / max(|x|, eps)) ONLY as the halo image `out_image` (width NT, halo out_halo), zero rows
  // where row_valid is 0.  A second pass (residual_conv riding in the same launch) takes the plain epilogue into C.
  const uint8_t* row_valid;
  int out_halo;
  float norm_eps;
  // norm_mode 4 (decode path, to_coor_logits + arg-max :935-944): a pass is CPARTS coordinates of part_cols == nbins logits each; every
  // epilogue thread takes the arg-max of its row's nbins logits of coordinate pass * CPARTS + cpart (first maximum wins) straight from
  // TMEM and writes the bin and the undiscretised coordinate of the UNPADDED face (b * nf + f) - the logits never reach global memory.
  float* am_coords;         // (B, P - lead, ncoord) fp32, NaN where row_valid is 0
  long long* am_bins;       // the same shape, int64 (0 where row_valid is 0), or null
  int am_B, am_P, am_lead, am_ncoord;
  float am_lo, am_hi;
  // fused neighbour aggregation (GW > 0 instantiation, phase 0 of a SAGE output linear): when g_x is set the K1 columns of A are the
  // mean over the in-neighbours of the fp32 rows g_x (M, K1), produced tile by tile by the gather warps; a1_image is not read
  const float* g_x;
  const int* g_indptr;
  const int* g_src;
  long long g_cap;
};
__host__ __device__ inline int lin_chunks(const LinArgs& a) { return (a.conv ? a.taps : 1) * (a.K1 / KC) + a.K2 / KC; }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// One launch runs one linear (nphase == 1) or two back to back (nphase == 2): phase 0 = a layer's output linear with the fused
// normalisation epilogue, phase 1 = `lin` of the NEXT layer, whose input rows are exactly the 128-row tile phase 0 has just
// written (as an activation image, to global memory = L2).  A CTA (pair) interleaves its tiles as A0 A1 A2 B0 A3 B1 ... so the
// tensor pipe works on other tiles while the epilogue of tile i normalises and stores; the producer waits on a_done[] (arrivals
// of the epilogue warps after a gpu-scope fence and a proxy fence) before it bulk-copies a tile back in for phase 1.  Saves a
// launch with its ramp (prologue, first loads, last epilogue, ~10 us at C2) per SAGE layer.
struct LinArgs2 {
  LinArgs p[2];
  int nphase;
  int pdl;            // launched with programmatic stream serialization: set up barriers / TMEM BEFORE waiting for the previous kernel
  long long* trace;   // timing experiments (MESHTOK_LINEAR_TRACE): clock64 stamps of cluster 0, [2 ranks][8 events][16 units]; else null
};

#define LT_STAMP(ev, unit)                                                                          \
  do {                                                                                              \
    if (g.trace != nullptr && blockIdx.x < 2 && (unit) < 16)                                        \
      g.trace[((int)blockIdx.x * 8 + (ev)) * 16 + (unit)] = clock64();                              \
  } while (0)

struct UnitSeq {   // this CTA (pair)'s units in issue order
  int n, u0, ustep, passes, nphase;
  __device__ __forceinline__ int count() const { return nphase == 2 ? 2 * n : n; }
  // j -> phase, row group (tile or tile pair), pass, and for two phases the ordinal i of the row group within this CTA
  __device__ __forceinline__ void get(int j, int& phase, int& grp, int& pass, int& i) const {
    if (nphase == 1) {
      const int u = u0 + j * ustep;
      phase = 0; grp = u / passes; pass = u % passes; i = j;
      return;
    }
    // A0 A1 | A2 B0 | A3 B1 | ... | B(n-2) B(n-1): phase 1 of a tile is issued two phase-0 units after its phase 0, so the
    // epilogue of tile i (normalise, store, fence) and the copy back in have two units of MMA time to hide behind
    if (n == 1) { phase = j; i = 0; }
    else if (j < 2) { phase = 0; i = j; }
    else {
      const int t = j - 2;
      if (t < 2 * (n - 2)) { phase = t & 1; i = (t & 1) ? (t >> 1) : (t >> 1) + 2; }
      else { phase = 1; i = n - 2 + (t - 2 * (n - 2)); }
    }
    grp = u0 + i * ustep; pass = 0;
  }
};

// TAIL: the instantiation that carries the norm_mode 3 epilogue (decode path); the tokenizer's linears run the one without it
// (with the extra epilogue in the same kernel they had 4 registers more and were 1 - 2 % slower)
template <bool PAIR, int CPARTS, bool TAIL = false, int GW = 0>
__device__ __forceinline__ void linear_tc_body(const LinArgs2& g) {
  using G = Cfg<PAIR, CPARTS, GW>;
  constexpr int EPI_WARPS = G::EPI_WARPS;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nphase = g.nphase;
  const int ring_nt = nphase == 2 ? max(g.p[0].NT, g.p[1].NT) : g.p[0].NT;   // the ring is laid out for the wider phase
  const uint32_t a_slot = g.p[0].conv ? tc::halo_block_bytes(g.p[0].halo) : 2 * A_PART;   // conv: single phase only
  const int STAGES = G::stages(ring_nt, a_slot);
  const uint32_t stage_bytes = G::stage_bytes(ring_nt, a_slot);
  uint8_t* ring = smem + G::RING_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::BAR_OFF);
  uint64_t* full = bars;                        // [STAGES]  producer expect_tx (+ in a pair, on the leader: the peer's forwarded arrival)
  uint64_t* empty = bars + G::MAX_STAGES;       // [STAGES]  MMA commit (multicast to both CTAs of a pair)
  uint64_t* acc_full = empty + G::MAX_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]       (in a pair the leader's are the ones waited on)
  uint64_t* a_done = acc_empty + 2;             // [4]       phase-0 tile i of this CTA is in global memory (epilogue warps -> producer)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_done + 4);
  uint64_t* gfull = a_done + 5;                 // [STAGES]  (GW > 0) the gather warps' A tile of the stage is written (both CTAs of a pair -> leader)

  const int warp = warp_id(), lane = lane_id();
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  // Earlier kernels' results are read from here on (the row count first: its load travels while the barriers and TMEM are
  // set up).  With PDL enabled this gives up the overlap of the prologue; PDL is opt-in and measured slower anyway.
  // Without PDL the row count's load travels while the barriers and TMEM are set up (~2.5 us per launch); with it the set-up itself
  // overlaps the previous kernel's tail and the count is read after the wait.
  int M = g.p[0].m_max;
  if (!g.pdl && g.p[0].m_dev) M = min(__ldg(g.p[0].m_dev), g.p[0].m_max);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], PAIR && leader ? 2 : 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], (PAIR ? 2 : 1) * EPI_WARPS);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&a_done[i], 1);
    if (GW > 0)
      for (int i = 0; i < STAGES; ++i) mbar_init(&gfull[i], PAIR ? 2 : 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair(tmem_slot, 512);
    else tmem_alloc(tmem_slot, 512);
  }
  float* sbias = reinterpret_cast<float*>(smem + G::VEC_OFF);   // phase p at sbias + p * 512 (two phases: N <= 256 each)
  float* sgamma = sbias + VEC_MAX_N;
  float* sbeta = sgamma + 64;
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // both CTAs' barriers and TMEM exist before anything crosses the pair
  tcgen05_fence_after();
  if (g.pdl) {
    pdl_wait();
    if (g.p[0].m_dev) M = min(*reinterpret_cast<const volatile int*>(g.p[0].m_dev), g.p[0].m_max);
  }
  const uint32_t tmem_base = *tmem_slot;
  const int m_tiles = (M + M_TILE - 1) / M_TILE;
  const int m_groups = PAIR ? (m_tiles + 1) / 2 : m_tiles;   // a group = the row tile(s) one MMA covers
  UnitSeq seq;
  seq.u0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  seq.ustep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  seq.passes = g.p[0].passes;
  seq.nphase = nphase;
  {
    const int total = nphase == 2 ? m_groups : m_groups * g.p[0].passes;
    seq.n = seq.u0 < total ? (total - seq.u0 + seq.ustep - 1) / seq.ustep : 0;
  }
  const int n_units = seq.count();

  if (warp == 0) {
    // ===== producer (warp-converged; one elected lane issues): own A block (16 KB) + own rows of the W chunk per stage =====
    uint32_t it = 0, s = 0, ph = 0;
    for (int j = 0; j < n_units; ++j) {
      int phase, grp, pass, ti;
      seq.get(j, phase, grp, pass, ti);
      const LinArgs& a = g.p[phase];
      const int kch1 = a.K1 / KC, kch2 = a.K2 / KC;
      const int kchunks = lin_chunks(a);
      const int taps = a.conv ? a.taps : 1;
      const uint32_t a_blk = a.conv ? a_slot : IMG_BLOCK;
      const uint32_t w_part = (uint32_t)a.NT * KC * 2;          // hi (or lo) bytes of one W chunk
      const uint32_t w_mine = PAIR ? w_part / 2 : w_part;       // ... of the rows this CTA stages
      const int m_tile = PAIR ? 2 * grp + (int)rank : grp;
      const bool have_a = m_tile < m_tiles;   // an odd tile count leaves the peer of the last pair without rows
      if (phase == 1) {   // the rows of this unit are the image tile this CTA's epilogue wrote in phase 0
        mbar_wait(&a_done[ti & 3], (uint32_t)(ti >> 2) & 1u);
        fence_proxy_async_all();
      }
      if (lane == 0) LT_STAMP(0, j);   // producer: free to issue this unit's loads
      const uint8_t* wsrc = a.w_image + (size_t)pass * kchunks * (2 * w_part) + (size_t)rank * w_mine;
      const uint8_t* a1src = a.a1_image + (size_t)m_tile * kch1 * a_blk;
      const uint8_t* a2src = a.a2_image ? a.a2_image + (size_t)m_tile * kch2 * IMG_BLOCK : nullptr;
      // conv: the block of column chunk ka is staged once per tap (ka outer, tap inner: the re-reads hit L2), with the W chunk of
      // columns tap * K1 + ka * 32
      for (int kc = 0, ka = 0, tap = 0; kc < kchunks; ++kc, ++it) {
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          const bool copy_a = have_a && !(GW > 0 && a.g_x != nullptr && ka < kch1);   // (else: the gather warps write the A tile)
          mbar_arrive_expect_tx(&full[s], (copy_a ? a_blk : 0u) + 2 * w_mine);
          uint8_t* stage = ring + s * stage_bytes;
          const uint8_t* asrc = ka < kch1 ? a1src + (size_t)ka * a_blk : a2src + (size_t)(ka - kch1) * IMG_BLOCK;
          if (copy_a) bulk_g2s(stage, asrc, a_blk, &full[s]);
          const uint8_t* wc = wsrc + (size_t)(a.conv ? tap * kch1 + ka : kc) * (2 * w_part);
          if (PAIR) {
            bulk_g2s(stage + a_slot, wc, w_mine, &full[s]);
            bulk_g2s(stage + a_slot + w_mine, wc + w_part, w_mine, &full[s]);
          } else {
            bulk_g2s(stage + a_slot, wc, 2 * w_part, &full[s]);
          }
        }
        __syncwarp();
        if (++tap == taps) { tap = 0; ++ka; }
        if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
      }
      if (lane == 0) LT_STAMP(1, j);   // producer: last load of the unit issued
    }
    if (PAIR) {   // drain: every multicast commit aimed at this CTA has landed before the CTA may exit
      const uint32_t pending = it < (uint32_t)STAGES ? it : (uint32_t)STAGES;
      for (uint32_t i = 0; i < pending; ++i) {
        const uint32_t jj = it - 1 - i;
        mbar_wait(&empty[jj % STAGES], (jj / STAGES) & 1);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (the leader CTA of a pair; warp-converged, one elected lane issues) =====
    if (leader) {
      const uint32_t s0 = smem_u32(ring);
      uint32_t s = 0, ph = 0;
      uint32_t gpar = 0;   // bit s: parity of the next completion of gfull[s] (only the gather chunks complete it)
      for (int j = 0; j < n_units; ++j) {
        int phase, grp, pass, ti;
        seq.get(j, phase, grp, pass, ti);
        const LinArgs& a = g.p[phase];
        const int kchunks = lin_chunks(a);
        const int g_chunks = (GW > 0 && a.g_x != nullptr) ? a.K1 / KC : 0;   // the first g_chunks A tiles come from the gather warps
        const int taps = a.conv ? a.taps : 1;
        const uint32_t rows_h = (uint32_t)(M_TILE + 2 * a.halo);
        const uint32_t a_lbo = a.conv ? rows_h * 16 : 128, a_sbo = a.conv ? 128 : SBO;   // halo image: consecutive rows 16 B apart
        const uint32_t a_lo = a.conv ? rows_h * 64 : A_PART;                              // offset of the lo part inside the block
        const uint32_t w_part = (uint32_t)a.NT * KC * 2;
        const uint32_t w_mine = PAIR ? w_part / 2 : w_part;
        const uint32_t idesc = make_idesc_f16_f32(PAIR ? 2 * M_TILE : M_TILE, a.NT, /*bf16=*/true);
        const uint32_t buf = (uint32_t)j & 1, aph = ((uint32_t)j >> 1) & 1;
        if (lane == 0) LT_STAMP(2, j);   // MMA: previous unit committed, waiting for the accumulator buffer
        mbar_wait(&acc_empty[buf], aph ^ 1);
        tcgen05_fence_after();
        if (lane == 0) LT_STAMP(3, j);   // MMA: accumulator buffer free
        const uint32_t d = tmem_base + buf * NT_MAX;
        for (int kc = 0, tap = 0; kc < kchunks; ++kc) {
          mbar_wait(&full[s], ph);
          if (GW > 0 && kc < g_chunks) {
            if (PAIR) mbar_wait_cluster(&gfull[s], (gpar >> s) & 1u);
            else mbar_wait(&gfull[s], (gpar >> s) & 1u);
            gpar ^= 1u << s;
          }
          tcgen05_fence_after();
          if (kc == 0 && lane == 0) LT_STAMP(4, j);   // MMA: first stage of the unit has landed
          if (elect_one()) {
            const uint32_t ah = s0 + s * stage_bytes + (uint32_t)(tap + a.halo - taps / 2) * 16, al = ah + a_lo;   // tap t = the block read t rows further
            const uint32_t wh = s0 + s * stage_bytes + a_slot, wl = wh + w_mine;
#pragma unroll
            for (int k16 = 0; k16 < KC / 16; ++k16) {
              const uint32_t ko = k16 * 256, koa = k16 * 2 * a_lbo;
              const uint64_t dah = make_smem_desc(ah + koa, a_lbo, a_sbo), dal = make_smem_desc(al + koa, a_lbo, a_sbo);
              const uint64_t dwh = make_smem_desc(wh + ko, 128, SBO), dwl = make_smem_desc(wl + ko, 128, SBO);
              const uint32_t acc = (kc | k16) != 0 ? 1u : 0u;
              if (PAIR) {
                umma_f16_pair(d, dal, dwh, idesc, acc);  // small terms first
                umma_f16_pair(d, dah, dwl, idesc, 1u);
                umma_f16_pair(d, dah, dwh, idesc, 1u);
              } else {
                umma_f16(d, dal, dwh, idesc, acc);
                umma_f16(d, dah, dwl, idesc, 1u);
                umma_f16(d, dah, dwh, idesc, 1u);
              }
            }
            if (PAIR) umma_commit_pair_mc(&empty[s], 3);
            else umma_commit(&empty[s]);
          }
          __syncwarp();
          if (++tap == taps) tap = 0;
          if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one()) {
          if (PAIR) umma_commit_pair_mc(&acc_full[buf], 3);
          else umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ===== completion forwarder (peer CTA of a pair): tells the leader that this CTA's copies have landed; lane s follows stage s =====
    if (PAIR && !leader && lane < STAGES) {
      uint32_t total_it = (uint32_t)seq.n * (uint32_t)lin_chunks(g.p[0]);
      if (nphase == 2) total_it += (uint32_t)seq.n * (uint32_t)lin_chunks(g.p[1]);
      const uint32_t remote = mapa_shared(smem_u32(&full[lane]), 0);
      uint32_t ph = 0;
      for (uint32_t it = lane; it < total_it; it += STAGES, ph ^= 1) {
        mbar_wait(&full[lane], ph);
        // relaxed: the bytes are in this CTA's shared memory once the phase flips, and only the tensor core reads them.
        // (A release.cluster arrive costs a full memory barrier per stage; ncu had this warp in membar stalls most of
        // the time and the leader's MMA warp waiting for it.)
        mbar_arrive_cluster_relaxed(remote);
      }
    }
  } else if (warp >= 4 && warp < 4 + EPI_WARPS) {
    // ===== epilogue (both CTAs of a pair, each on its own 128 rows) =====
    const int ew = warp - 4;
    const int q = warp & 3;
    const int cpart = ew >> 2;   // which part of the NT columns this warp handles
    float* stage = reinterpret_cast<float*>(smem + G::EPI_OFF) + ew * (32 * EPI_LD);
    const uint32_t acc_empty_remote0 = PAIR ? mapa_shared(smem_u32(&acc_empty[0]), 0) : 0u;
    const uint32_t acc_empty_remote1 = PAIR ? mapa_shared(smem_u32(&acc_empty[1]), 0) : 0u;
    // bias / gamma / beta -> shared memory, by the epilogue warps themselves while the first stages are in flight (the
    // epilogue reads them per 16-column block; as global loads they sat on its critical path)
    for (int p = 0; p < nphase; ++p) {
      const LinArgs& ap = g.p[p];
      for (int i = threadIdx.x - 128; i < ap.N; i += EPI_WARPS * 32) {
        sbias[p * 512 + i] = ap.bias ? __ldg(ap.bias + i) : 0.f;
        if (ap.norm_mode == 2) { sgamma[i] = __ldg(ap.gamma + i); sbeta[i] = __ldg(ap.beta + i); }
      }
    }
    named_bar_sync(3, EPI_WARPS * 32);
    int tail_units = 0;
    for (int j = 0; j < n_units; ++j) {
      int phase, grp, pass, ti;
      seq.get(j, phase, grp, pass, ti);
      const LinArgs& a = g.p[phase];
      const float* sb = sbias + phase * 512;
      const int part_cols = (((a.NT + CPARTS - 1) / CPARTS + 15) / 16) * 16;
      const int c_lo = min(a.NT, cpart * part_cols), c_hi = min(a.NT, c_lo + part_cols);
      const uint32_t acc_it = (uint32_t)j;
      const int m_tile = PAIR ? 2 * grp + (int)rank : grp;
      const uint32_t buf = acc_it & 1, aph = (acc_it >> 1) & 1;
      if (ew == 0 && lane == 0) LT_STAMP(5, j);   // epilogue: ready for the unit
      mbar_wait(&acc_full[buf], aph);
      tcgen05_fence_after();
      if (ew == 0 && lane == 0) LT_STAMP(6, j);   // epilogue: accumulator complete (= the unit's MMAs are done)
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * NT_MAX;
      const int n0 = pass * a.NT;
      const int row0 = m_tile * M_TILE + q * 32;
      const bool mode3 = TAIL && a.norm_mode == 3;
      if (TAIL && a.norm_mode == 4) {
        // ---- arg-max epilogue: thread = row, this warp's column part = one coordinate's logits ----
        const long long grow = (long long)row0 + lane;
        const int coord = pass * CPARTS + cpart;
        float best = -INFINITY;
        int bi = 0;
        for (int c = c_lo; c < c_hi; c += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = __uint_as_float(v[j]) + sb[n0 + c + j];
            if (x > best) { best = x; bi = c - c_lo + j; }   // ascending bins, strict: the first maximum wins like torch.argmax
          }
        }
        if (grow < M && coord < a.am_ncoord) {
          const int b = (int)(grow / a.am_P);
          const int f = (int)(grow - (long long)b * a.am_P) - a.am_lead;
          if (b < a.am_B && f >= 0) {   // not a separator row
            const long long item = ((long long)b * (a.am_P - a.am_lead) + f) * a.am_ncoord + coord;
            if (a.row_valid[grow] == 0) {
              a.am_coords[item] = __int_as_float(0x7fc00000);
              if (a.am_bins) a.am_bins[item] = 0;
            } else {
              float t = __fadd_rn((float)bi, 0.5f);
              t = __fdiv_rn(t, (float)(c_hi - c_lo));
              a.am_coords[item] = __fadd_rn(__fmul_rn(t, __fsub_rn(a.am_hi, a.am_lo)), a.am_lo);
              if (a.am_bins) a.am_bins[item] = bi;
            }
          }
        }
      } else if (a.norm_mode != 0 && !(mode3 && pass != 0)) {
        // ---- fused row epilogue (passes == 1, so NT == N and n0 == 0; mode 3: pass 0 of at most two) ----
        // alternates per phase-0 unit; mode 3 counts its own units (plain second-pass units in between have no barrier)
        const int xsel = mode3 ? (tail_units++ & 1) : (ti & 1);
        float* xs = reinterpret_cast<float*>(smem + G::XCH_OFF) + xsel * (3 * CPARTS * 128);
        const int rit = q * 32 + lane;
        const long long grow = (long long)row0 + lane;
        const bool live = grow < M;
        auto xsum = [&](int k) {   // row total of exchange slot k over the column parts
          float t = 0.f;
#pragma unroll
          for (int p = 0; p < CPARTS; ++p) t += xs[(k * CPARTS + p) * 128 + rit];
          return t;
        };
        auto put = [&](int c, const float (&o)[16]) {   // 16 consecutive columns of this thread's row
          if (!live) return;
          if (a.C) {
            float4* dst = reinterpret_cast<float4*>(a.C + (size_t)grow * a.N + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          }
          if (a.out_image) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint4 hi, lo;
              split2(o[8 * g + 0], o[8 * g + 1], hi.x, lo.x);
              split2(o[8 * g + 2], o[8 * g + 3], hi.y, lo.y);
              split2(o[8 * g + 4], o[8 * g + 5], hi.z, lo.z);
              split2(o[8 * g + 6], o[8 * g + 7], hi.w, lo.w);
              uint8_t* pimg = a.out_image + image_offset(a.N, grow, c + 8 * g);
              *reinterpret_cast<uint4*>(pimg) = hi;
              *reinterpret_cast<uint4*>(pimg + IMG_PART) = lo;
            }
          }
        };
        auto load16 = [&](int c, float (&o)[16]) {       // accumulator + bias
          uint32_t v[16];
          tmem_ld16(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(sb + c + j);
            o[j] = __uint_as_float(v[j]) + bb.x; o[j + 1] = __uint_as_float(v[j + 1]) + bb.y;
            o[j + 2] = __uint_as_float(v[j + 2]) + bb.z; o[j + 3] = __uint_as_float(v[j + 3]) + bb.w;
          }
        };
        // x / max(|x|, eps) is done as x * (1 / max(|x|, eps)): one rounding more than the reference's division, far
        // below the 1e-5 of the split-bf16 product itself, at a tenth of the instructions
        if (mode3) {
          float ss = 0.f;
          for (int c = c_lo; c < c_hi; c += 16) {
            float o[16];
            load16(c, o);
#pragma unroll
            for (int j = 0; j < 16; ++j) ss = fmaf(o[j], o[j], ss);
          }
          xs[cpart * 128 + rit] = ss;
          named_bar_sync(2, EPI_WARPS * 32);
          const bool okrow = live && a.row_valid[grow] != 0;
          const float sc = sqrtf((float)a.NT) / fmaxf(sqrtf(xsum(0)), a.norm_eps);
          const int halo = a.out_halo;
          const long long tile = grow >> 7, n_tiles = ((long long)M + 127) >> 7;
          const int r = (int)(grow & 127);
          const uint32_t lo_off = 64u * (uint32_t)(IMG_ROWS + 2 * halo);
          for (int c = c_lo; c < c_hi; c += 16) {
            float o[16];
            load16(c, o);
            if (!live) continue;
#pragma unroll
            for (int gg = 0; gg < 2; ++gg) {
              uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
              if (okrow) {
                float t[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float x = o[8 * gg + j] * sc; t[j] = __fdividef(x, 1.f + __expf(-x)); }
                split2(t[0], t[1], hi.x, lo.x);
                split2(t[2], t[3], hi.y, lo.y);
                split2(t[4], t[5], hi.z, lo.z);
                split2(t[6], t[7], hi.w, lo.w);
              }
              const int col = c + 8 * gg;
              uint8_t* pimg = a.out_image + halo_offset(a.NT, halo, tile, r, col);
              *reinterpret_cast<uint4*>(pimg) = hi;
              *reinterpret_cast<uint4*>(pimg + lo_off) = lo;
              if (r < halo && tile > 0) {                       // the copy behind the previous tile
                pimg = a.out_image + halo_offset(a.NT, halo, tile - 1, r + IMG_ROWS, col);
                *reinterpret_cast<uint4*>(pimg) = hi;
                *reinterpret_cast<uint4*>(pimg + lo_off) = lo;
              }
              if (r >= IMG_ROWS - halo && tile + 1 < n_tiles) {  // the copy in front of the next tile
                pimg = a.out_image + halo_offset(a.NT, halo, tile + 1, r - IMG_ROWS, col);
                *reinterpret_cast<uint4*>(pimg) = hi;
                *reinterpret_cast<uint4*>(pimg + lo_off) = lo;
              }
            }
          }
        } else if (a.norm_mode == 1) {
          float ss = 0.f;
          for (int c = c_lo; c < c_hi; c += 16) {
            float o[16];
            load16(c, o);
#pragma unroll
            for (int j = 0; j < 16; ++j) ss = fmaf(o[j], o[j], ss);
          }
          xs[cpart * 128 + rit] = ss;
          named_bar_sync(2, EPI_WARPS * 32);
          const float inv = 1.f / fmaxf(sqrtf(xsum(0)), 1e-12f);
          for (int c = c_lo; c < c_hi; c += 16) {
            float o[16];
            load16(c, o);
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] *= inv;
            put(c, o);
          }
        } else {
          // <= 32 columns per thread (host checks NT <= 64): the row stays in registers
          float o0[16], o1[16];
          const bool h0 = c_lo < c_hi, h1 = c_lo + 16 < c_hi;
#pragma unroll
          for (int j = 0; j < 16; ++j) { o0[j] = 0.f; o1[j] = 0.f; }
          if (h0) load16(c_lo, o0);
          if (h1) load16(c_lo + 16, o1);
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) { ss = fmaf(o0[j], o0[j], ss); ss = fmaf(o1[j], o1[j], ss); }
          xs[cpart * 128 + rit] = ss;
          named_bar_sync(2, EPI_WARPS * 32);
          const float inv = 1.f / fmaxf(sqrtf(xsum(0)), 1e-12f);
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float t0 = o0[j] * inv, t1 = o1[j] * inv;
#if MESHTOK_SILU_FAST
            // |t| <= 1 after the normalisation: ex2.approx and the approximate divide are each within ~2 ulp here, far below the 1e-5 of
            // the split-bf16 product feeding them; the IEEE divide + expf were two thirds of this epilogue's instructions
            o0[j] = h0 ? __fdividef(t0, 1.f + __expf(-t0)) : 0.f;
            o1[j] = h1 ? __fdividef(t1, 1.f + __expf(-t1)) : 0.f;
#else
            o0[j] = h0 ? __fdiv_rn(t0, 1.f + expf(-t0)) : 0.f;
            o1[j] = h1 ? __fdiv_rn(t1, 1.f + expf(-t1)) : 0.f;
#endif
            sum += o0[j] + o1[j];
          }
          xs[(CPARTS + cpart) * 128 + rit] = sum;
          named_bar_sync(2, EPI_WARPS * 32);
          const float mean = xsum(1) / (float)a.N;
          float var = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float d0 = h0 ? o0[j] - mean : 0.f, d1 = h1 ? o1[j] - mean : 0.f;
            var = fmaf(d0, d0, var);
            var = fmaf(d1, d1, var);
          }
          xs[(2 * CPARTS + cpart) * 128 + rit] = var;
          named_bar_sync(2, EPI_WARPS * 32);
          const float rstd = 1.f / sqrtf(xsum(2) / (float)a.N + 1e-5f);
          if (h0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o0[j] = (o0[j] - mean) * rstd * sgamma[c_lo + j] + sbeta[c_lo + j];
            put(c_lo, o0);
          }
          if (h1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o1[j] = (o1[j] - mean) * rstd * sgamma[c_lo + 16 + j] + sbeta[c_lo + 16 + j];
            put(c_lo + 16, o1);
          }
        }
      } else {
      // columns [c_lo, c_hi) of this pass in blocks of 16 (NT is a multiple of 16); TMEM loads are double buffered.
      // Row phase (thread = row): scale, bias, ReLU, image chunks, sum of squares; then a per-warp shared-memory
      // transpose for the fp32 row-major stores.
      const int c4 = (lane & 3) * 4, rsub = lane >> 2;
      const long long grow = (long long)row0 + lane;
      const bool live = grow < M;
      float scale = 1.f;
      if (a.row_ssq != nullptr && live) {   // the producer left its rows un-normalised: x / |x| commutes with the linear map
        float t = 0.f;
        for (int i = 0; i < a.row_ssq_parts; ++i) t += a.row_ssq[(size_t)grow * a.row_ssq_parts + i];
        scale = 1.f / fmaxf(sqrtf(t), 1e-12f);
      }
      float ss = 0.f;
      // row-major destinations of the four 8-row groups this lane stores after the transpose (null: row beyond M)
      float* cdst[4];
#pragma unroll
      for (int it2 = 0; it2 < 4; ++it2) {
        const int r = it2 * 8 + rsub;
        cdst[it2] = (a.C != nullptr && row0 + r < M) ? a.C + (size_t)(row0 + r) * a.N + n0 + c4 : nullptr;
      }
      const float* stage_rd = stage + rsub * EPI_LD + c4;
      float* stage_wr = stage + lane * EPI_LD;
      auto process = [&](const int c, const uint32_t (&v)[16]) {   // one 16-column block of this thread's row
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(sb + n0 + c + j);
          o[j] = fmaf(__uint_as_float(v[j]), scale, bb.x); o[j + 1] = fmaf(__uint_as_float(v[j + 1]), scale, bb.y);
          o[j + 2] = fmaf(__uint_as_float(v[j + 2]), scale, bb.z); o[j + 3] = fmaf(__uint_as_float(v[j + 3]), scale, bb.w);
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
        }
        if (a.ssq_out) {
#pragma unroll
          for (int j = 0; j < 16; ++j) ss = fmaf(o[j], o[j], ss);
        }
        if (a.out_image && live && !mode3) {   // (mode 3: out_image is the halo image of pass 0, not of these columns)
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            uint4 hi, lo;
            split2(o[8 * gg + 0], o[8 * gg + 1], hi.x, lo.x);
            split2(o[8 * gg + 2], o[8 * gg + 3], hi.y, lo.y);
            split2(o[8 * gg + 4], o[8 * gg + 5], hi.z, lo.z);
            split2(o[8 * gg + 6], o[8 * gg + 7], hi.w, lo.w);
            uint8_t* pimg = a.out_image + image_offset(a.N, grow, n0 + c + 8 * gg);
            *reinterpret_cast<uint4*>(pimg) = hi;
            *reinterpret_cast<uint4*>(pimg + IMG_PART) = lo;
          }
        }
        if (a.C) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(stage_wr + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
          __syncwarp();
#pragma unroll
          for (int it2 = 0; it2 < 4; ++it2) {
            const float4 t = *reinterpret_cast<const float4*>(stage_rd + it2 * 8 * EPI_LD);
            if (cdst[it2] != nullptr) *reinterpret_cast<float4*>(cdst[it2] + c) = t;
          }
          __syncwarp();
        }
      };
      // two blocks per iteration with fixed register sets (va, vb): written with `parity ? vb : va` the compiler kept the
      // loop rolled and paid 60 SEL / MOV per block to pick the set - a third of the epilogue's instructions, on a path that
      // is bound by the latency of one warp's dependent instruction stream (clock64 trace: ~800 cycles per block)
      uint32_t va[16], vb[16];
      if (c_lo < c_hi) tmem_ld16(taddr + c_lo, va);
      for (int c = c_lo; c < c_hi; c += 32) {
        tmem_ld_wait();
        if (c + 16 < c_hi) tmem_ld16(taddr + c + 16, vb);
        process(c, va);
        if (c + 16 < c_hi) {
          tmem_ld_wait();
          if (c + 32 < c_hi) tmem_ld16(taddr + c + 32, va);
          process(c + 16, vb);
        }
      }
      if (a.ssq_out && live) a.ssq_out[(size_t)grow * (CPARTS * a.passes) + CPARTS * pass + cpart] = ss;
      }
      tcgen05_fence_before();
      __syncwarp();
      if (ew == 0 && lane == 0) LT_STAMP(7, j);   // epilogue: unit stored
      if (lane == 0) {   // the TMEM reads are complete (wait::ld); nothing is handed over through memory, hence relaxed
        if (leader) mbar_arrive(&acc_empty[buf]);
        else mbar_arrive_cluster_relaxed(buf ? acc_empty_remote1 : acc_empty_remote0);
      }
      if (nphase == 2 && phase == 0) {
        // The image tile must be in L2 and ordered before the producer's bulk copy (async proxy) reads it back.  One thread
        // fences for all: the barrier makes the other epilogue threads' stores happen-before its gpu-scope fence (a fence per
        // warp cost 8 L1 invalidations per unit and 12 % of the kernel's stall samples).
        named_bar_sync(4, EPI_WARPS * 32);
        if (ew == 0 && lane == 0) {
          __threadfence();
          fence_proxy_async_all();
          mbar_arrive(&a_done[ti & 3]);
        }
      }
    }
  }

  else if (GW > 0 && warp >= 4 + EPI_WARPS) {
    // ===== gather warps: the aggregated-neighbour half of A, produced in shared memory (SAGEConv mean aggregation) =====
    // Tile = this CTA's 128 rows; per 32-column chunk a warp takes every GW-th 8-row group: lane = (row & 7) + 8 * (8-column group), so a
    // warp's 16-byte stores fill 512 contiguous bytes of the canonical K-major tile (no bank conflicts) and a row's four lanes read one
    // 128-byte line per neighbour.  The CSR entries of a lane's rows (first four sources, count, 1 / count) are read once per tile and
    // kept in registers for all chunks; neighbours are added in list order, then sum * (1 / count): the same values, bit for bit, as
    // gather_mean_smem_wide_kernel writes into the aggregate image.  x rows come through L1 (neighbouring rows share their sources).
    constexpr int STEPS = 16 / (GW > 0 ? GW : 1);
    const int gw = warp - (4 + EPI_WARPS);
    const int r8 = lane >> 2, cg = lane & 3;   // a row's four lanes side by side: a quarter warp's 16-byte loads touch two lines, not eight
    const uint32_t gfull_remote = PAIR ? mapa_shared(smem_u32(&gfull[0]), 0) : 0u;
    uint32_t s = 0, ph = 0;
    for (int j = 0; j < n_units; ++j) {
      int phase, grp, pass, ti;
      seq.get(j, phase, grp, pass, ti);
      const LinArgs& a = g.p[phase];
      const int kchunks = lin_chunks(a);
      if (a.g_x == nullptr) {
        // Nothing to gather in this unit, but the ring is still followed stage by stage: a parity wait only tells two consecutive
        // phases apart, so a warp that skipped a use of stage s could be a whole ring turn ahead of the MMA warp at its next wait on
        // empty[s] and take the stale phase for the one it needs (seen as wrong tokens at 512 meshes per launch).
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&empty[s], ph ^ 1);
          if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
        }
        continue;
      }
      const int kch1 = a.K1 / KC;
      const int m_tile = PAIR ? 2 * grp + (int)rank : grp;
      const long long row0 = (long long)m_tile * M_TILE;
      int id[STEPS][4], cnt[STEPS], beg[STEPS];
      float rcp[STEPS];
#pragma unroll
      for (int t = 0; t < STEPS; ++t) {
        const long long row = row0 + (t * GW + gw) * 8 + r8;
        cnt[t] = -1;   // row beyond M: nothing to read, nothing to write
        beg[t] = 0; rcp[t] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) id[t][k] = -1;
        if (row < M) {
          const int b0 = __ldg(a.g_indptr + row), e_true = __ldg(a.g_indptr + row + 1);
          const int e = (int)min((long long)e_true, a.g_cap);
          beg[t] = b0;
          cnt[t] = max(e - b0, 0);
          rcp[t] = __frcp_rn(fmaxf((float)(e_true - b0), 1.f));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < cnt[t]) {
              const int v = __ldg(a.g_src + b0 + k);
              id[t][k] = (unsigned)v < (unsigned)M ? v : -1;   // (a corrupt list contributes nothing)
            }
        }
      }
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&empty[s], ph ^ 1);   // every stage, also the bulk-copied ones (see above)
        if (kc < kch1) {
          uint8_t* stage = ring + s * stage_bytes;
          const float* xcol = a.g_x + kc * KC + cg * 8;
#pragma unroll
          for (int t0 = 0; t0 < STEPS; t0 += 2) {
            float4 v[2][4][2];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                v[u][k][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                v[u][k][1] = v[u][k][0];
                if (id[t0 + u][k] >= 0) {
                  const float4* p = reinterpret_cast<const float4*>(xcol + (size_t)id[t0 + u][k] * a.K1);
                  v[u][k][0] = __ldg(p);
                  v[u][k][1] = __ldg(p + 1);
                }
              }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int t = t0 + u;
              if (cnt[t] < 0) continue;
              float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < cnt[t]) {   // (an invalid source adds the zeros loaded above: x + 0 = x)
                  a0.x += v[u][k][0].x; a0.y += v[u][k][0].y; a0.z += v[u][k][0].z; a0.w += v[u][k][0].w;
                  a1.x += v[u][k][1].x; a1.y += v[u][k][1].y; a1.z += v[u][k][1].z; a1.w += v[u][k][1].w;
                }
              for (int k = 4; k < cnt[t]; ++k) {   // faces with more than four in-neighbours
                const int sv = __ldg(a.g_src + beg[t] + k);
                if ((unsigned)sv < (unsigned)M) {
                  const float4* p = reinterpret_cast<const float4*>(xcol + (size_t)sv * a.K1);
                  const float4 w0 = __ldg(p), w1 = __ldg(p + 1);
                  a0.x += w0.x; a0.y += w0.y; a0.z += w0.z; a0.w += w0.w;
                  a1.x += w1.x; a1.y += w1.y; a1.z += w1.z; a1.w += w1.w;
                }
              }
              const float rc = rcp[t];
              uint4 hi, lo;
              split2(a0.x * rc, a0.y * rc, hi.x, lo.x);
              split2(a0.z * rc, a0.w * rc, hi.y, lo.y);
              split2(a1.x * rc, a1.y * rc, hi.z, lo.z);
              split2(a1.z * rc, a1.w * rc, hi.w, lo.w);
              uint8_t* dst = stage + (uint32_t)(t * GW + gw) * 512u + (uint32_t)cg * 128u + (uint32_t)r8 * 16u;
              *reinterpret_cast<uint4*>(dst) = hi;
              *reinterpret_cast<uint4*>(dst + A_PART) = lo;
            }
          }
          fence_proxy_async_smem();               // the tile is read by the tensor core (async proxy)
          named_bar_sync(5, GW * 32);
          if (gw == 0 && lane == 0) {
            if (leader) mbar_arrive(&gfull[s]);
            // relaxed, like the completion forwarder's arrival: the tile is in THIS CTA's shared memory (every writer fenced it for the
            // async proxy before the barrier above) and only this SM's tensor core reads it; a release.cluster arrival is a full
            // memory barrier per stage
            else mbar_arrive_cluster_relaxed(gfull_remote + s * 8u);
          }
        }
        if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
      }
    }
  }

  pdl_launch_dependents();   // this CTA's work is done: let the next kernel's CTAs start their prologue
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // nothing of the pair is still in flight towards either CTA
  if (warp == 2) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(Cfg<false, 2>::NUM_THREADS, 1) linear_tc_kernel(const __grid_constant__ LinArgs2 a) { linear_tc_body<false, 2>(a); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg<true, 2>::NUM_THREADS, 1) linear_tc_pair_kernel(const __grid_constant__ LinArgs2 a) { linear_tc_body<true, 2>(a); }
__global__ void __launch_bounds__(Cfg<false, 2>::NUM_THREADS, 1) linear_tc_tail_kernel(const __grid_constant__ LinArgs2 a) { linear_tc_body<false, 2, true>(a); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg<true, 2>::NUM_THREADS, 1) linear_tc_pair_tail_kernel(const __grid_constant__ LinArgs2 a) { linear_tc_body<true, 2, true>(a); }
// the SAGE output linears with the neighbour aggregation fused in: four gather warps behind the epilogue warps (512 threads, 128 registers)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg<true, 2, 4>::NUM_THREADS, 1) linear_tc_pair_gather_kernel(const __grid_constant__ LinArgs2 a) { linear_tc_body<true, 2, false, 4>(a); }

// W (N, K) fp32 -> image [pass][kchunk]{hi (NT x 32), lo (NT x 32)} canonical K-major, SBO 512.
__global__ void __launch_bounds__(256) build_w_image_kernel(const float* __restrict__ W, int N, int K, int NT, uint8_t* __restrict__ image) {
  const long long total = (long long)N * (K / 8);
  const int kchunks = K / KC;
  const uint32_t w_part = (uint32_t)NT * KC * 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / (K / 8));
    const int k8 = (int)(i % (K / 8));
    const float4* src = reinterpret_cast<const float4*>(W + (size_t)n * K + k8 * 8);
    const float4 v0 = src[0], v1 = src[1];
    uint4 hi, lo;
    split2(v0.x, v0.y, hi.x, lo.x);
    split2(v0.z, v0.w, hi.y, lo.y);
    split2(v1.x, v1.y, hi.z, lo.z);
    split2(v1.z, v1.w, hi.w, lo.w);
    const int pass = n / NT, r = n % NT;
    const int kc = (k8 * 8) / KC, kin = (k8 * 8) % KC;
    const size_t base = ((size_t)pass * kchunks + kc) * (2 * (size_t)w_part);
    const uint32_t off = canon_offset(r, kin, SBO);
    *reinterpret_cast<uint4*>(image + base + off) = hi;
    *reinterpret_cast<uint4*>(image + base + w_part + off) = lo;
  }
}

// fp32 rows (M, K) -> activation image (used for tensors that do not come from an image-writing kernel)
__global__ void __launch_bounds__(256) rows_to_image_kernel(const float* __restrict__ A, int K, const int* __restrict__ m_dev, int m_max,
                                                            uint8_t* __restrict__ image) {
  const int M = m_dev ? min(*m_dev, m_max) : m_max;
  const long long total = (long long)M * (K / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / (K / 4);
    const int col = (int)(i % (K / 4)) * 4;
    image_store4(image, K, row, col, *reinterpret_cast<const float4*>(A + (size_t)row * K + col));
  }
}

}  // namespace

int linear_tc_plan(int N, int K1, int K2, int* passes, int* NT) {
  if (N > VEC_MAX_N) {
    set_error("tcgen05 linear: output width %d is more than %d; use the SIMT path", N, VEC_MAX_N);
    return MESHTOK_ERR_UNSUPPORTED;
  }
  if (N <= 0 || K1 <= 0 || K1 % KC != 0 || K2 % KC != 0) {
    set_error("tcgen05 linear: input widths must be multiples of %d (got %d, %d); use the SIMT path", KC, K1, K2);
    return MESHTOK_ERR_UNSUPPORTED;
  }
  // fewest passes of equal width (a multiple of 16, at most NT_MAX columns)
  int p = ceil_div(N, NT_MAX);
  while (p <= N / 16 && (N % p != 0 || (N / p) % 16 != 0)) ++p;
  if (p > N / 16 || N % p != 0 || (N / p) % 16 != 0) {
    set_error("tcgen05 linear: output width %d does not split into passes of a multiple of 16 columns; use the SIMT path", N);
    return MESHTOK_ERR_UNSUPPORTED;
  }
  *passes = p;
  *NT = N / p;
  return MESHTOK_OK;
}

int linear_tc_fused_norm_ok(int N, int norm_mode) {
  if (N > NT_MAX || N % 16 != 0 || (norm_mode == 2 && N > 64)) {
    set_error("tcgen05 linear: fused row epilogue needs N %% 16 == 0 and N <= %d (<= 64 with LayerNorm), got %d", NT_MAX, N);
    return MESHTOK_ERR_UNSUPPORTED;
  }
  return MESHTOK_OK;
}

// CTA pairs unless MESHTOK_LINEAR_TC=1 asks for the single-CTA kernel (kept for A/B runs).  (A pair kernel with 16
// epilogue warps was tried too: 96 registers per thread -> spills, and a shallower ring; slower than 8 warps.)
static bool linear_tc_pair() {
  static int pair = -1;
  if (pair < 0) { const char* e = getenv("MESHTOK_LINEAR_TC"); pair = (e && atoi(e) == 1) ? 0 : 1; }
  return pair == 1;
}

int linear_tc_ssq_parts(int N) {
  int passes = 0, NT = 0;
  if (linear_tc_plan(N, KC, 0, &passes, &NT) != MESHTOK_OK) return 1 << 20;   // not a tcgen05 shape: no deferred normalisation
  return 2 * passes;
}

size_t linear_tc_image_bytes(int N, int K) { return (size_t)N * K * 4; }

size_t rows_image_bytes(long long rows, int K) { return tc::image_bytes(rows, K); }

int launch_linear_tc_build_image(const float* W, int N, int K, void* image, cudaStream_t stream) {
  int passes, NT;
  int rc = linear_tc_plan(N, K, 0, &passes, &NT);
  if (rc != MESHTOK_OK) return rc;
  const long long total = (long long)N * (K / 8);
  const long long want = (total + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148 * 8 ? want : 148 * 8);
  build_w_image_kernel<<<blocks, 256, 0, stream>>>(W, N, K, NT, static_cast<uint8_t*>(image));
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int launch_rows_to_image(const float* A, int K, const int* m_dev, int m_max, void* image, cudaStream_t stream) {
  MT_CHECK_ARG(K % KC == 0, "rows_to_image: width %d must be a multiple of %d", K, KC);
  if (m_max == 0) return MESHTOK_OK;
  const long long want = ((long long)m_max * (K / 4) + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148 * 16 ? want : 148 * 16);
  rows_to_image_kernel<<<blocks, 256, 0, stream>>>(A, K, m_dev, m_max, static_cast<uint8_t*>(image));
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

static int fill_lin_args(LinArgs& a, const void* a1_image, int K1, const void* a2_image, int K2, const void* w_image, const float* bias,
                         float* C, int N, const int* m_dev, int m_max, bool relu, const LinearRowEpilogue* epi) {
  MT_CHECK_ARG(w_image != nullptr && a1_image != nullptr && (K2 == 0 || a2_image != nullptr), "tcgen05 linear: operand image is null");
  a.norm_mode = 0; a.gamma = a.beta = nullptr; a.out_image = nullptr;
  a.conv = 0; a.taps = 1; a.halo = 0;
  a.row_valid = nullptr; a.out_halo = 0; a.norm_eps = 0.f;
  a.am_coords = nullptr; a.am_bins = nullptr; a.am_B = a.am_P = a.am_lead = a.am_ncoord = 0; a.am_lo = a.am_hi = 0.f;
  a.ssq_out = nullptr; a.row_ssq = nullptr; a.row_ssq_parts = 0;
  a.g_x = nullptr; a.g_indptr = nullptr; a.g_src = nullptr; a.g_cap = 0;
  if (epi != nullptr && epi->norm_mode == 0) {
    // plain epilogue with the deferred-normalisation extras
    MT_CHECK_ARG(epi->out_image == nullptr || N % KC == 0, "tcgen05 linear: an output image needs N %% %d == 0 (got %d)", KC, N);
    MT_CHECK_ARG(C != nullptr || epi->out_image != nullptr, "tcgen05 linear: no output given");
    a.out_image = static_cast<uint8_t*>(epi->out_image);
    a.ssq_out = epi->ssq_out; a.row_ssq = epi->row_ssq; a.row_ssq_parts = epi->row_ssq_parts;
  } else if (epi != nullptr && epi->norm_mode != 0) {
    int rc0 = linear_tc_fused_norm_ok(N, epi->norm_mode);
    if (rc0 != MESHTOK_OK) return rc0;
    MT_CHECK_ARG(!relu, "tcgen05 linear: the fused row epilogue does not combine with ReLU");
    MT_CHECK_ARG(epi->norm_mode == 1 || (epi->gamma && epi->beta), "tcgen05 linear: LayerNorm needs gamma and beta");
    MT_CHECK_ARG(epi->out_image == nullptr || N % KC == 0, "tcgen05 linear: an output image needs N %% %d == 0 (got %d)", KC, N);
    MT_CHECK_ARG(C != nullptr || epi->out_image != nullptr, "tcgen05 linear: no output given");
    a.norm_mode = epi->norm_mode; a.gamma = epi->gamma; a.beta = epi->beta; a.out_image = static_cast<uint8_t*>(epi->out_image);
  } else {
    MT_CHECK_ARG(C != nullptr, "tcgen05 linear: C is null");
  }
  a.a1_image = static_cast<const uint8_t*>(a1_image); a.a2_image = static_cast<const uint8_t*>(a2_image);
  a.K1 = K1; a.K2 = K2; a.w_image = static_cast<const uint8_t*>(w_image); a.bias = bias; a.C = C;
  a.N = N; a.m_dev = m_dev; a.m_max = m_max; a.relu = relu ? 1 : 0;
  return linear_tc_plan(N, K1, K2, &a.passes, &a.NT);
}

// MESHTOK_LINEAR_TRACE=k: the k-th (mod 7) tcgen05 linear launch of every step records clock64 stamps of its cluster 0
static long long* g_lin_trace = nullptr;
static int g_lin_trace_k = -2;
static long long g_lin_launches = 0;

static int launch_lin2(const LinArgs2& g_in, int m_max, cudaStream_t stream) {
  LinArgs2 g = g_in;
  if (g_lin_trace_k == -2) {
    const char* e = getenv("MESHTOK_LINEAR_TRACE");
    g_lin_trace_k = e ? atoi(e) : -1;
    if (g_lin_trace_k >= 0) { MT_CHECK_CUDA(cudaMalloc(&g_lin_trace, 2 * 8 * 16 * sizeof(long long))); MT_CHECK_CUDA(cudaMemset(g_lin_trace, 0, 2 * 8 * 16 * sizeof(long long))); }
  }
  g.trace = (g_lin_trace_k >= 0 && (g_lin_launches++ % 7) == g_lin_trace_k) ? g_lin_trace : nullptr;
  g.pdl = pdl_enabled() ? 1 : 0;
  MT_ENSURE_SMEM(linear_tc_kernel, (Cfg<false, 2>::SMEM_TOTAL));
  MT_ENSURE_SMEM(linear_tc_pair_kernel, (Cfg<true, 2>::SMEM_TOTAL));
  MT_ENSURE_SMEM(linear_tc_tail_kernel, (Cfg<false, 2>::SMEM_TOTAL));
  MT_ENSURE_SMEM(linear_tc_pair_tail_kernel, (Cfg<true, 2>::SMEM_TOTAL));
  const int passes = g.nphase == 2 ? 1 : g.p[0].passes;
  const bool tail = g.p[0].norm_mode == 3 || g.p[0].norm_mode == 4;   // decode path: Block tail / arg-max in the epilogue
  if (g.p[0].g_x != nullptr) {
    using PG = Cfg<true, 2, 4>;
    MT_CHECK_ARG(linear_tc_pair() && !tail, "tcgen05 linear: the fused aggregation exists for the CTA-pair kernel only");
    MT_ENSURE_SMEM(linear_tc_pair_gather_kernel, (PG::SMEM_TOTAL));
    const int units = ceil_div(ceil_div(m_max, M_TILE), 2) * passes;
    const int grid = 2 * min(units, 148 / 2);
    MT_LAUNCH_PDL(linear_tc_pair_gather_kernel, grid, PG::NUM_THREADS, PG::SMEM_TOTAL, stream, g);
  } else if (linear_tc_pair()) {
    using P2 = Cfg<true, 2>;
    const int units = ceil_div(ceil_div(m_max, M_TILE), 2) * passes;
    const int grid = 2 * min(units, 148 / 2);
    if (tail) { MT_LAUNCH_PDL(linear_tc_pair_tail_kernel, grid, P2::NUM_THREADS, P2::SMEM_TOTAL, stream, g); }
    else { MT_LAUNCH_PDL(linear_tc_pair_kernel, grid, P2::NUM_THREADS, P2::SMEM_TOTAL, stream, g); }
  } else {
    using S2 = Cfg<false, 2>;
    const int units = ceil_div(m_max, M_TILE) * passes;
    const int grid = min(units, 148);
    if (tail) { MT_LAUNCH_PDL(linear_tc_tail_kernel, grid, S2::NUM_THREADS, S2::SMEM_TOTAL, stream, g); }
    else { MT_LAUNCH_PDL(linear_tc_kernel, grid, S2::NUM_THREADS, S2::SMEM_TOTAL, stream, g); }
  }
  return MESHTOK_OK;
}

int launch_linear_tc(const void* a1_image, int K1, const void* a2_image, int K2, const void* w_image, const float* bias, float* C,
                     int N, const int* m_dev, int m_max, bool relu, const LinearRowEpilogue* epi, cudaStream_t stream) {
  if (m_max == 0) return MESHTOK_OK;
  LinArgs2 g;
  memset(&g, 0, sizeof(g));
  g.nphase = 1;
  int rc = fill_lin_args(g.p[0], a1_image, K1, a2_image, K2, w_image, bias, C, N, m_dev, m_max, relu, epi);
  if (rc != MESHTOK_OK) return rc;
  return launch_lin2(g, m_max, stream);
}

// Conv1d(C_in -> N, kernel `taps`) over the rows of a halo image (padded row space: at least taps / 2 zero rows between two meshes);
// w_image = image of the (N, taps * C_in) matrix with tap-major columns
int launch_conv_tc(const void* a_halo_image, int C_in, int taps, int halo, const void* w_image, const float* bias, float* C, int N, int m_max,
                   cudaStream_t stream, const ConvBlockTail* tail) {
  if (m_max == 0) return MESHTOK_OK;
  MT_CHECK_ARG(taps >= 1 && (taps & 1) == 1 && halo >= taps / 2 && halo <= 8 && C_in % KC == 0, "tcgen05 conv: kernel %d must be odd, halo %d >= kernel / 2 and <= 8, C_in %d a multiple of %d", taps, halo, C_in, KC);
  LinArgs2 g;
  memset(&g, 0, sizeof(g));
  g.nphase = 1;
  // (a fused tail without a second pass has no fp32 output: fill_lin_args only needs to see SOME output pointer)
  float* c_seen = (C == nullptr && tail != nullptr) ? static_cast<float*>(tail->out_image) : C;
  int rc = fill_lin_args(g.p[0], a_halo_image, C_in, nullptr, 0, w_image, bias, c_seen, N, nullptr, m_max, false, nullptr);
  if (rc != MESHTOK_OK) return rc;
  g.p[0].C = C;
  g.p[0].conv = 1; g.p[0].taps = taps; g.p[0].halo = halo;
  rc = linear_tc_plan(N, taps * C_in, 0, &g.p[0].passes, &g.p[0].NT);
  if (rc != MESHTOK_OK) return rc;
  if (tail != nullptr) {
    // the first tail->width output columns are whole rows of a Block: they must be exactly pass 0
    if (g.p[0].NT != tail->width || g.p[0].passes > 2 || tail->width % KC != 0) {
      set_error("tcgen05 conv: the fused Block tail needs the %d block columns to be one pass (N %d splits into %d passes of %d)", tail->width, N,
                g.p[0].passes, g.p[0].NT);
      return MESHTOK_ERR_UNSUPPORTED;
    }
    MT_CHECK_ARG(tail->valid && tail->out_image && tail->out_halo >= 0 && tail->out_halo <= 8 && (g.p[0].passes == 1 || C != nullptr),
                 "tcgen05 conv: fused Block tail needs valid, an output image and (two passes) the fp32 output for the second pass");
    g.p[0].norm_mode = 3;
    g.p[0].row_valid = tail->valid;
    g.p[0].out_image = static_cast<uint8_t*>(tail->out_image);
    g.p[0].out_halo = tail->out_halo;
    g.p[0].norm_eps = tail->eps;
  }
  return launch_lin2(g, m_max, stream);
}

// to_coor_logits + arg-max (decode path): Conv1d / Linear whose passes are pairs of coordinates (N = passes * 2 * nbins, the weight rows of
// coordinates >= ncoord zero padding); the epilogue keeps only the arg-max of every coordinate's nbins logits, see LinArgs::am_*
int launch_conv_argmax_tc(const void* a_halo_image, int C_in, int taps, int halo, const void* w_image, const float* bias, int N, int m_max,
                          const ConvArgmax& am, cudaStream_t stream) {
  if (m_max == 0) return MESHTOK_OK;
  MT_CHECK_ARG(taps >= 1 && (taps & 1) == 1 && halo >= taps / 2 && halo <= 8 && C_in % KC == 0, "tcgen05 conv: kernel %d must be odd, halo %d >= kernel / 2 and <= 8, C_in %d a multiple of %d", taps, halo, C_in, KC);
  MT_CHECK_ARG(am.valid && am.coords && am.B >= 0 && am.lead >= 0 && am.P > am.lead && am.ncoord >= 1 && am.nbins >= 16 && (long long)am.B * am.P <= m_max,
               "tcgen05 conv + arg-max: bad row space / outputs");
  LinArgs2 g;
  memset(&g, 0, sizeof(g));
  g.nphase = 1;
  int rc = fill_lin_args(g.p[0], a_halo_image, C_in, nullptr, 0, w_image, bias, am.coords, N, nullptr, m_max, false, nullptr);
  if (rc != MESHTOK_OK) return rc;
  g.p[0].C = nullptr;
  g.p[0].conv = 1; g.p[0].taps = taps; g.p[0].halo = halo;
  rc = linear_tc_plan(N, taps * C_in, 0, &g.p[0].passes, &g.p[0].NT);
  if (rc != MESHTOK_OK) return rc;
  constexpr int CP = Cfg<true, 2>::CPARTS;
  if (g.p[0].NT != CP * am.nbins || am.nbins % 16 != 0 || g.p[0].passes * CP < am.ncoord) {
    set_error("tcgen05 conv + arg-max: a pass must be %d coordinates of %d logits (N %d splits into %d passes of %d for %d coordinates)", CP, am.nbins, N,
              g.p[0].passes, g.p[0].NT, am.ncoord);
    return MESHTOK_ERR_UNSUPPORTED;
  }
  g.p[0].norm_mode = 4;
  g.p[0].row_valid = am.valid;
  g.p[0].am_coords = am.coords; g.p[0].am_bins = reinterpret_cast<long long*>(am.bins);
  g.p[0].am_B = am.B; g.p[0].am_P = am.P; g.p[0].am_lead = am.lead; g.p[0].am_ncoord = am.ncoord;
  g.p[0].am_lo = am.lo; g.p[0].am_hi = am.hi;
  return launch_lin2(g, m_max, stream);
}

// MESHTOK_LINEAR_B2B=0 keeps every linear in its own launch (A/B runs)
bool linear_tc_b2b_ok(int N, int norm_mode) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MESHTOK_LINEAR_B2B"); on = (e && atoi(e) == 0) ? 0 : 1; }
  // both phases single pass; phase 1 reads the image phase 0 writes: K = N, a multiple of the 32-column image chunk
  return on == 1 && norm_mode != 0 && N <= NT_MAX && N % KC == 0 && (norm_mode != 2 || N <= 64);
}

// MESHTOK_GATHER_FUSED=1 turns the fused aggregation on (needs the CTA-pair kernel and K1 a multiple of 32).  It is OFF by default:
// bit-identical to the separate launches (scripts/gpu_gather_fused_check.py) but measured SLOWER at the bench workload - 4 aggregation
// launches (95 us) disappear and the four output linears grow from 160 to 350 us, because four gather warps at 128 registers keep only
// 32 KB of neighbour rows in flight per SM and every 32-column chunk then costs four exposed load round trips (ncu: 54 % of the gather
// warps' samples in long-scoreboard stalls, tensor pipe active 24 %).
bool linear_tc_gather_ok(int K1, int N, int norm_mode) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MESHTOK_GATHER_FUSED"); on = (e && atoi(e) == 1) ? 1 : 0; }
  return on == 1 && linear_tc_pair() && K1 % KC == 0 && linear_tc_b2b_ok(N, norm_mode);
}

// phase 0: C/out_image = norm([A1 | A2] W^T + bias) (epi: fused row epilogue, out_image required);  phase 1: C2 = relu(out W2^T + bias2)
int launch_linear_tc_b2b(const void* a1_image, int K1, const void* a2_image, int K2, const void* w_image, const float* bias, float* C,
                         int N, const int* m_dev, int m_max, const LinearRowEpilogue* epi, const void* w2_image, const float* bias2,
                         float* C2, cudaStream_t stream, const LinearGather* gather) {
  if (m_max == 0) return MESHTOK_OK;
  MT_CHECK_ARG(epi != nullptr && epi->norm_mode != 0 && epi->out_image != nullptr, "tcgen05 linear pair of phases: phase 0 needs the fused row epilogue and an output image");
  if (gather != nullptr) {
    MT_CHECK_ARG(linear_tc_gather_ok(K1, N, epi->norm_mode), "tcgen05 linear: fused aggregation not available for K1 %d, N %d", K1, N);
    MT_CHECK_ARG(gather->x && gather->indptr && gather->src && gather->cap > 0 && gather->x != C2,
                 "tcgen05 linear: fused aggregation needs x / indptr / src, and x must not be the buffer phase 1 writes");
    if (a1_image == nullptr) a1_image = a2_image;   // (never read; fill_lin_args wants a pointer)
  }
  MT_CHECK_ARG(linear_tc_b2b_ok(N, epi->norm_mode), "tcgen05 linear pair of phases: width %d not supported", N);
  LinArgs2 g;
  memset(&g, 0, sizeof(g));
  g.nphase = 2;
  int rc = fill_lin_args(g.p[0], a1_image, K1, a2_image, K2, w_image, bias, C, N, m_dev, m_max, false, epi);
  if (rc != MESHTOK_OK) return rc;
  rc = fill_lin_args(g.p[1], epi->out_image, N, nullptr, 0, w2_image, bias2, C2, N, m_dev, m_max, true, nullptr);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(g.p[0].passes == 1 && g.p[1].passes == 1, "tcgen05 linear pair of phases: both must be single pass");
  if (gather != nullptr) { g.p[0].g_x = gather->x; g.p[0].g_indptr = gather->indptr; g.p[0].g_src = gather->src; g.p[0].g_cap = gather->cap; }
  return launch_lin2(g, m_max, stream);
}

int linear_tc_read_trace(long long* out, int n) {
  if (g_lin_trace == nullptr) return 0;
  cudaDeviceSynchronize();
  cudaMemcpy(out, g_lin_trace, sizeof(long long) * (size_t)min(n, 2 * 8 * 16), cudaMemcpyDeviceToHost);
  return min(n, 2 * 8 * 16);
}

}  // namespace meshtok
