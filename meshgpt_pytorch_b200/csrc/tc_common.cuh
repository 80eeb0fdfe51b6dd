// sm_100a building blocks shared by the tensor-core kernels: mbarrier, 1-D bulk async copy (UBLKCP),
// tcgen05 TMEM allocation / MMA / commit / load, and the UMMA shared-memory + instruction descriptors.
//
// Operand layout used everywhere here: K-major, SWIZZLE_NONE ("interleaved") canonical UMMA layout.
//   core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes;
//   LBO (leading byte offset) = distance between core matrices adjacent along K,
//   SBO (stride  byte offset) = distance between core matrices adjacent along M/N (next 8 rows).
// Weights / the codebook are re-tiled into exactly this image once at load time, so a stage is one
// contiguous cp.async.bulk (no tensor map, no swizzle); activations are converted fp32 -> bf16 by the
// loader warps straight into the same layout.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace meshtok {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (reported as a CUDA error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("meshtok: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x, (int)threadIdx.x,
             smem_u32(bar), parity);
      __trap();
    }
  }
}

// One lane of a fully converged warp (the lowest).  Role loops run warp-converged and wrap only the single-thread
// instructions (bulk copies, MMAs, commits) in `if (elect_one())`: issued from a diverged `if (lane == 0)` region the
// compiler has to move every descriptor into uniform registers with an ELECT / R2UR.BROADCAST / BRA.U.ANY sequence per
// instruction, which made the MMA issue thread - not the tensor pipe - the limiter (profiles/r01d_stalls_rvq_tc_search.txt).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- proxies / bulk copy -----------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// orders this thread's generic-proxy accesses (all state spaces) against async-proxy accesses (bulk copies, tcgen05)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 2-D tiled TMA load (cp.async.bulk.tensor, SASS UTMALDG): box (c0 .., r0 ..) of the tensor described by `tmap` -> dst_smem (row major
// box, 128-byte aligned), completion counted in bytes on `bar`.  Elements outside the tensor arrive as zeros and still count.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int c0, int r0, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst_smem)),
               "l"(tmap), "r"(c0), "r"(r0), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- tcgen05 ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]^T, fp16 or bf16 inputs (per the instruction descriptor), fp32 accumulate;
// issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all tcgen05 ops issued so far by this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA pairs (cta_group::2): one MMA spans two SMs of a cluster ------------------------------------------
// Each CTA stages its own 128 rows of A and its own half of B; the leader CTA (cluster rank 0) issues the MMA for the
// pair, which halves the shared-memory operand traffic per SM compared with two independent M=128 MMAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` (a shared::cta address) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded like mbar_wait; acquire at cluster scope (the barrier is also arrived on by the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("meshtok: cluster mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x, (int)threadIdx.x,
             smem_u32(bar), parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {  // the same warp in BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A . B^T over the CTA pair (M = 256: 128 rows per CTA; N codes split across the CTAs' smem)
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at this shared-memory offset in every CTA of `cta_mask` once all tcgen05 ops issued so far
// by this thread have completed
__device__ __forceinline__ void umma_commit_pair_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- descriptors -----------------------------------------------------------------------------------
// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_NONE, K-major.
//   [0,14) start address >> 4   [16,30) LBO >> 4   [32,46) SBO >> 4   [46,48) version = 1
//   [49,52) base offset = 0     [52] lbo mode = 0  [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
// Instruction descriptor, kind::f16: [4,6) D format (1 = f32)  [7,10) A format (0 = f16, 1 = bf16)  [10,13) B format
//   [15] A major (0 = K)  [16] B major (0 = K)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f16_f32(int M, int N, bool bf16) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// ---- SWIZZLE_128B K-major tiles (16-bit elements) ---------------------------------------------------------
// A tile is [rows][64 elements] = rows of 128 bytes; the 16-byte chunk c of row r is stored at chunk c ^ (r & 7)
// (Swizzle<3,4,3> on the byte address), 8-row groups are 1024 B apart, and the tile base must be 1024-byte aligned.
// One MMA K step (16 elements) advances the descriptor start address by 32 bytes inside the 128-byte row.
constexpr int SW128_K = 64;
__host__ __device__ constexpr uint32_t sw128_offset(int r, int k) {
  return static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(((k >> 3) ^ r) & 7) << 4) + static_cast<uint32_t>(k & 7) * 2u;
}
//   [0,14) start address >> 4   [16,30) LBO >> 4 = 1 (unused)   [32,46) SBO >> 4 = 64   [46,48) version = 1   [61,64) = 2
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// byte offset of element (row r, column k) of a K-major canonical tile whose 8-row groups are `sbo` apart
__host__ __device__ constexpr uint32_t canon_offset(int r, int k, uint32_t sbo) {
  return static_cast<uint32_t>(r >> 3) * sbo + static_cast<uint32_t>(k >> 3) * 128u + static_cast<uint32_t>(r & 7) * 16u +
         static_cast<uint32_t>(k & 7) * 2u;
}

}  // namespace tc
}  // namespace meshtok

// ---- split-bf16 activation images ---------------------------------------------------------------------
// An (rows x K) fp32 activation is kept in global memory as the exact byte image the tcgen05 linear kernel
// wants in shared memory: for every (128-row tile, 32-column chunk) a 16 KB block = 8 KB of hi = bf16(x) followed
// by 8 KB of lo = bf16(x - hi), each in the canonical K-major layout (8-row groups 512 B apart).  A GEMM stage is
// then one contiguous cp.async.bulk.  Rows are padded to a multiple of 128 (pad rows are never stored by the GEMM).
namespace meshtok {
namespace tc {

constexpr int IMG_ROWS = 128;
constexpr int IMG_KC = 32;
constexpr uint32_t IMG_PART = IMG_ROWS * IMG_KC * 2;   // 8 KB
constexpr uint32_t IMG_BLOCK = 2 * IMG_PART;           // 16 KB

__host__ __device__ inline size_t image_bytes(long long rows, int K) {
  return (size_t)((rows + IMG_ROWS - 1) / IMG_ROWS) * (size_t)(K / IMG_KC) * IMG_BLOCK;
}
// byte offset of the hi element (row, col); the lo element is IMG_PART bytes further
__device__ __forceinline__ size_t image_offset(int K, long long row, int col) {
  const long long m_tile = row >> 7;
  const int r = (int)(row & 127), kc = col >> 5, kin = col & 31;
  return ((size_t)m_tile * (K >> 5) + kc) * IMG_BLOCK + (size_t)((r >> 3) * 512 + (kin >> 3) * 128 + (r & 7) * 16 + (kin & 7) * 2);
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// two adjacent columns (col even)
__device__ __forceinline__ void image_store2(uint8_t* img, int K, long long row, int col, float x, float y) {
  uint32_t hi, lo;
  split2(x, y, hi, lo);
  uint8_t* p = img + image_offset(K, row, col);
  *reinterpret_cast<uint32_t*>(p) = hi;
  *reinterpret_cast<uint32_t*>(p + IMG_PART) = lo;
}
// four adjacent columns (col % 4 == 0)
__device__ __forceinline__ void image_store4(uint8_t* img, int K, long long row, int col, float4 v) {
  uint2 hi, lo;
  split2(v.x, v.y, hi.x, lo.x);
  split2(v.z, v.w, hi.y, lo.y);
  uint8_t* p = img + image_offset(K, row, col);
  *reinterpret_cast<uint2*>(p) = hi;
  *reinterpret_cast<uint2*>(p + IMG_PART) = lo;
}


// ---- halo images (Conv1d over the row sequence) ------------------------------------------------------------
// Same split-bf16 idea, different order inside a (128-row tile, 32-column chunk) block: [hi | lo][4 column groups of 8][128 + 2 halo
// rows][16 bytes], i.e. CONSECUTIVE ROWS 16 BYTES APART, the tile's rows preceded by the last `halo` rows of the previous tile and
// followed by the first `halo` rows of the next one (stored twice).  As a UMMA operand that is LBO = rows * 16, SBO = 128, and the
// tile shifted down by t rows is the same block read from 16 t bytes further - tap t of a convolution needs no im2col copy.
__host__ __device__ constexpr uint32_t halo_block_bytes(int halo) { return 128u * (uint32_t)(IMG_ROWS + 2 * halo); }
__host__ __device__ inline size_t halo_image_bytes(long long rows, int C, int halo) {
  return (size_t)((rows + IMG_ROWS - 1) / IMG_ROWS) * (size_t)(C / IMG_KC) * halo_block_bytes(halo);
}
// byte offset of the hi 16-byte group (tile, tile-relative row r in [-halo, 128 + halo), columns col .. col + 7); lo is 64 * rows further
__device__ __forceinline__ size_t halo_offset(int C, int halo, long long tile, int r, int col) {
  const uint32_t rows = (uint32_t)(IMG_ROWS + 2 * halo);
  return ((size_t)tile * (C >> 5) + (col >> 5)) * (128u * rows) + (size_t)(((col >> 3) & 3) * rows + (uint32_t)(r + halo)) * 16u;
}

}  // namespace tc
}  // namespace meshtok
