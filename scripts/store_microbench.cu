// Store-path microbenchmark (sm_100a): how fast can the 8 epilogue warps of one CTA per SM push 128 KB units into an L2-resident buffer
//   (a) with st.global.v4 from registers, a warp writing 4 x 128 contiguous bytes per instruction (the image stores of linear_tc.cu),
//   (b) staged: st.shared.v4 into a per-warp 4 KB buffer, then two cp.async.bulk shared -> global copies of 2 KB issued by one lane,
// optionally (mode bit 1) while a ninth warp streams bulk loads global -> shared as the producer of the linear kernel does?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/store_microbench.cu -o gpurun_out/store_microbench
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int UNITS = 64;            // 128 KB units per CTA per launch
constexpr int UNIT_BYTES = 131072;
constexpr int WINDOW = 4;            // units of a CTA's output window

// every warp owns a 16 KB share of a unit: 4 chunks of (hi 2 KB | lo 2 KB)
__global__ void __launch_bounds__(288, 1) store_bench(uint8_t* out, const uint8_t* src, int mode, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8 * 8192 + 65536);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint8_t* base = out + (size_t)blockIdx.x * WINDOW * UNIT_BYTES;   // a window that stays in L2 (148 x 512 KB = 74 MB)
  const long long t0 = clock64();
  if (warp < 8) {
    uint8_t* stage = smem + warp * 8192;   // two 4 KB buffers
    uint4 v = make_uint4(lane, warp, blockIdx.x, 7);
    int n_issued = 0;
    for (int u = 0; u < UNITS; ++u) {
      uint8_t* ubase = base + (size_t)(u % WINDOW) * UNIT_BYTES + warp * 16384;
      for (int ch = 0; ch < 4; ++ch) {
        uint8_t* dst = ubase + ch * 4096;
        v.x += ch;
        if ((mode & 1) == 0) {
#pragma unroll
          for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int cg = 0; cg < 4; ++cg)   // lane = row: (row >> 3) * 512 + cg * 128 + (row & 7) * 16
              *reinterpret_cast<uint4*>(dst + part * 2048 + (lane >> 3) * 512 + cg * 128 + (lane & 7) * 16) = v;
        } else {
          uint8_t* sb = stage + (n_issued & 1) * 4096;
          if (n_issued >= 2) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
          }
#pragma unroll
          for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int cg = 0; cg < 4; ++cg)
              *reinterpret_cast<uint4*>(sb + part * 2048 + (lane >> 3) * 512 + cg * 128 + (lane & 7) * 16) = v;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(sb)), "r"(2048) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + 2048), "r"(smem_u32(sb + 2048)), "r"(2048) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++n_issued;
        }
      }
    }
    if ((mode & 1) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if ((mode & 2) && lane == 0) {
    // a producer: 2 x 128 KB of bulk loads per unit into a 64 KB scratch ring (what the MMA operands cost)
    uint8_t* scratch = smem + 8 * 8192;
    uint32_t phase = 0;
    const uint8_t* s = src + (size_t)blockIdx.x * 262144;
    for (int u = 0; u < UNITS; ++u)
      for (int k = 0; k < 4; ++k) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(65536) : "memory");
        for (int j = 0; j < 4; ++j)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(scratch + j * 16384)),
                       "l"(s + ((u * 4 + k) % 4) * 65536 + j * 16384), "r"(16384), "r"(smem_u32(bar))
                       : "memory");
        uint32_t done = 0;
        while (!done)
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
        phase ^= 1;
      }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  int n_sm = 0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  const size_t out_bytes = (size_t)n_sm * UNITS * UNIT_BYTES;   // bytes stored per launch
  uint8_t *out, *src;
  long long* cyc;
  cudaMalloc(&out, (size_t)n_sm * WINDOW * UNIT_BYTES);
  cudaMalloc(&src, (size_t)n_sm * 262144);
  cudaMalloc(&cyc, n_sm * sizeof(long long));
  cudaMemset(src, 1, (size_t)n_sm * 262144);
  const size_t smem = 8 * 8192 + 65536 + 64;
  cudaFuncSetAttribute(store_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int mode = 0; mode < 4; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(a);
      store_bench<<<n_sm, 288, smem>>>(out, src, mode, cyc);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    const double gb = (double)out_bytes / 1e9;
    printf("mode %d (%s stores%s): %.3f ms, %.0f GB/s of stores = %.1f B/clk/SM at 1.965 GHz  [%s]\n", mode, (mode & 1) ? "bulk shared->global" : "st.global.v4",
           (mode & 2) ? " + 2x bulk loads" : "", best, gb / (best * 1e-3), gb * 1e9 / (best * 1e-3) / n_sm / 1.965e9, cudaGetErrorString(e));
  }
  return 0;
}
