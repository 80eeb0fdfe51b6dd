// tcgen05.mma issue-rate microbenchmark (sm_100a): how many cycles does one kind::f16 MMA take when nothing else is in
// the way?  Operands sit in shared memory (zeros, SWIZZLE_128B K-major), one elected thread issues ITERS x 4 MMAs back
// to back into alternating accumulators and commits once.  Variants: N in {64,128,192,256}, cta_group 1 (M=128) and 2
// (M=256 over a CTA pair), with and without other warps reading the accumulators (tcgen05.ld) or bulk-copying into
// shared memory at the same time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I meshgpt_pytorch_b200/csrc scripts/mma_microbench.cu -o gpurun_out/mma_microbench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace meshtok {
void set_error(const char*, ...) {}
long long g_launch_count = 0;
}  // namespace meshtok

using namespace meshtok;
using namespace meshtok::tc;

constexpr int ITERS = 2000;

// mode bit 0: 4 extra warps read the accumulators with tcgen05.ld all the time
// mode bit 1: one warp bulk-copies 16 KB blocks from global into a scratch stage all the time
template <int N, int CG>
__global__ void __launch_bounds__(384, 1) mma_bench(long long* cycles, const uint8_t* gsrc, int mode, unsigned long long* ld_count) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                       // 128 rows x 64 k fp16 = 16 KB
  uint8_t* sB = smem + 16384;               // up to 256 rows x 64 k = 32 KB
  uint8_t* sScratch = smem + 16384 + 32768; // 16 KB bulk-copy target
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + 32768 + 16384);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  volatile int* stop = reinterpret_cast<volatile int*>(tmem_slot + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    *stop = 0;
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (warp == 2) {
    if (CG == 1) tmem_alloc(tmem_slot, 512);
    else tmem_alloc_pair(tmem_slot, 512);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool leader = CG == 1 || cluster_ctarank() == 0;

  if (warp == 1) {
    if (leader) {
      const uint32_t idesc = make_idesc_f16_f32(CG == 1 ? 128 : 256, N, false);
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      long long t0 = 0, t1 = 0;
      if (mode & 4) {
        // latency test: 4 MMAs, commit, wait for the commit to land, repeat -> cycles per round trip
        t0 = clock64();
        for (int it = 0; it < ITERS / 4; ++it) {
          if (elect_one()) {
            const uint32_t d = tmem_base + (it & 1) * 256;
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
              const uint64_t da = make_smem_desc_sw128(a0 + k16 * 32), db = make_smem_desc_sw128(b0 + k16 * 32);
              if (CG == 1) umma_f16(d, da, db, idesc, k16 ? 1u : 0u);
              else umma_f16_pair(d, da, db, idesc, k16 ? 1u : 0u);
            }
            if (CG == 1) umma_commit(&bars[0]);
            else umma_commit(&bars[0]);
          }
          __syncwarp();
          mbar_wait(&bars[0], it & 1);
          tcgen05_fence_after();
        }
        t1 = clock64();
        if (lane == 0) { cycles[blockIdx.x] = (t1 - t0) * 4; *stop = 1; }   // x4: report per "ITERS x 4 MMAs" like the other modes
        if (CG == 2 && elect_one()) umma_commit_pair_mc(&bars[2], 2);      // release the peer
        __syncwarp();
      } else {
      if (elect_one()) {
        t0 = clock64();
        for (int it = 0; it < ITERS; ++it) {
          const uint32_t d = tmem_base + (it & 1) * 256;
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            // mode bit 3: SWIZZLE_NONE canonical K-major operands (32-wide K chunks: LBO 128, SBO 512) as the split-bf16 linear uses
            const uint32_t ko = (mode & 8) ? (k16 & 1) * 256 + (k16 >> 1) * 8192 : k16 * 32;
            const uint64_t da = (mode & 8) ? make_smem_desc(a0 + (k16 & 1) * 256 + (k16 >> 1) * 4096, 128, 512) : make_smem_desc_sw128(a0 + ko);
            const uint64_t db = (mode & 8) ? make_smem_desc(b0 + ko, 128, 512) : make_smem_desc_sw128(b0 + ko);
            if (CG == 1) umma_f16(d, da, db, idesc, k16 ? 1u : 0u);
            else umma_f16_pair(d, da, db, idesc, k16 ? 1u : 0u);
          }
        }
        if (CG == 1) umma_commit(&bars[0]);
        else umma_commit_pair_mc(&bars[0], 3);
      }
      __syncwarp();
      mbar_wait(&bars[0], 0);
      t1 = clock64();
      if (lane == 0) { cycles[blockIdx.x] = t1 - t0; *stop = 1; }
      }
    } else {
      mbar_wait((mode & 4) ? &bars[2] : &bars[0], 0);
      if (lane == 0) *stop = 1;
    }
  } else if (warp >= 4 && (mode & 1)) {
    // 8 warps read like the screening epilogue: lane quarter = warp % 4, column half = (warp - 4) / 4, 4 x 32 columns,
    // double-buffered loads; counts the loads so that the TMEM read bandwidth can be reported
    const int q = warp & 3, chalf = (warp - 4) >> 2;
    uint32_t acc = 0;
    unsigned long long n = 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + chalf * 128;
    while (!*stop) {
      uint32_t va[32], vb[32];
      tmem_ld32(taddr, va);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        uint32_t(&v)[32] = (c & 1) ? vb : va;
        if (c + 1 < 4) tmem_ld32(taddr + (c + 1) * 32, (c & 1) ? va : vb);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += v[j];
      }
      n += 4;
    }
    if (acc == 0x12345678u) cycles[0] = 0;
    if (lane == 0) atomicAdd(ld_count, n);
  } else if (warp == 0 && (mode & 2)) {
    uint32_t ph = 0;
    while (!*stop) {
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars[1], 16384);
        bulk_g2s(sScratch, gsrc + (size_t)(blockIdx.x % 64) * 16384, 16384, &bars[1]);
      }
      __syncwarp();
      mbar_wait(&bars[1], ph);
      ph ^= 1;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    if (CG == 1) tmem_dealloc(tmem_base, 512);
    else tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int N, int CG>
void run(int mode, long long* d_cycles, const uint8_t* gsrc, unsigned long long* d_ld) {
  const int smem = 16384 + 32768 + 16384 + 256;
  cudaFuncSetAttribute(mma_bench<N, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d_cycles, 0, 148 * sizeof(long long));
    cudaMemset(d_ld, 0, sizeof(unsigned long long));
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma_bench<N, CG>, d_cycles, gsrc, mode, d_ld);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d CG=%d mode=%d: %s\n", N, CG, mode, cudaGetErrorString(e)); exit(1); }
  }
  long long h[148];
  cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  int n = 0;
  double sum = 0;
  for (int i = 0; i < 148; ++i) if (h[i] > 0) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; ++n; }
  const double per = sum / n / (ITERS * 4.0);
  const double ideal = (CG == 1 ? 128.0 : 256.0) * N / (256.0 * CG);
  unsigned long long lds = 0;
  cudaMemcpy(&lds, d_ld, sizeof(lds), cudaMemcpyDeviceToHost);
  // every counted load moves 32 lanes x 32 columns x 4 B; all CTAs ran for about sum/n cycles
  const double ld_bytes_per_clk_sm = (double)lds * 4096.0 / 148.0 / (sum / n);
  printf("N=%3d cta_group=%d mode=%d (ld=%d copy=%d): %.1f cycles/MMA (max CTA %.1f), ideal %.0f -> %.0f %% of the tensor peak; tcgen05.ld %.0f B/clk/SM\n", N, CG, mode,
         mode & 1, (mode >> 1) & 1, per, mx / (ITERS * 4.0), ideal, 100.0 * ideal / per, ld_bytes_per_clk_sm);
}

int main() {
  long long* d_cycles;
  uint8_t* gsrc;
  cudaMalloc(&d_cycles, 148 * sizeof(long long));
  cudaMalloc(&gsrc, 64 * 16384);
  cudaMemset(gsrc, 0, 64 * 16384);
  unsigned long long* d_ld;
  cudaMalloc(&d_ld, sizeof(unsigned long long));
  // latency: cycles per (4 MMAs + commit + wait) = reported cycles/MMA x 4
  run<128, 1>(4, d_cycles, gsrc, d_ld);
  run<256, 1>(4, d_cycles, gsrc, d_ld);
  run<256, 2>(4, d_cycles, gsrc, d_ld);
  for (int mode = 8; mode < 12; ++mode) {   // SWIZZLE_NONE operands: alone, + tcgen05.ld, + bulk copies, + both
    run<128, 1>(mode, d_cycles, gsrc, d_ld);
    run<192, 1>(mode, d_cycles, gsrc, d_ld);
    run<256, 1>(mode, d_cycles, gsrc, d_ld);
    run<192, 2>(mode, d_cycles, gsrc, d_ld);
    run<256, 2>(mode, d_cycles, gsrc, d_ld);
  }
  for (int mode = 0; mode < 4; ++mode) {
    run<192, 2>(mode, d_cycles, gsrc, d_ld);
    run<64, 1>(mode, d_cycles, gsrc, d_ld);
    run<128, 1>(mode, d_cycles, gsrc, d_ld);
    run<192, 1>(mode, d_cycles, gsrc, d_ld);
    run<256, 1>(mode, d_cycles, gsrc, d_ld);
    run<128, 2>(mode, d_cycles, gsrc, d_ld);
    run<256, 2>(mode, d_cycles, gsrc, d_ld);
  }
  return 0;
}
