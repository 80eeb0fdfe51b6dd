// Shared helpers for libmeshtok (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/meshtok.h"

namespace meshtok {

void set_error(const char* fmt, ...);

#define MT_CHECK_ARG(cond, ...)                       \
  do {                                                \
    if (!(cond)) {                                    \
      ::meshtok::set_error(__VA_ARGS__);              \
      return MESHTOK_ERR_INVALID_ARGUMENT;            \
    }                                                 \
  } while (0)

#define MT_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::meshtok::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MESHTOK_ERR_CUDA;                                                                \
    }                                                                                         \
  } while (0)

// opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device (remembered per kernel and device, profile.cu)
int ensure_dynamic_smem(const void* kernel, int bytes);
#define MT_ENSURE_SMEM(kernel, bytes)                                                                         \
  do {                                                                                                        \
    const int _rc = ::meshtok::ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), (int)(bytes));      \
    if (_rc != MESHTOK_OK) return _rc;                                                                        \
  } while (0)

extern long long g_launch_count;   // kernels launched by this library (bench.py reports it as gpu_launches)
#define MT_CHECK_LAUNCH()                  \
  do {                                     \
    ++::meshtok::g_launch_count;           \
    MT_CHECK_CUDA(cudaGetLastError());     \
  } while (0)

// Optional per-section device timing (cudaEvent pairs on the caller's stream); off by default.
enum ProfSection {
  PROF_TOPOLOGY = 0, PROF_FACE_EMBED, PROF_SAGE_LINEAR, PROF_SAGE_GATHER, PROF_SAGE_NORM, PROF_PDC_LINEAR,
  PROF_SCATTER_MEAN, PROF_LFQ, PROF_RVQ_PREP_UPDATE, PROF_RVQ_SEARCH, PROF_RVQ_RESCORE, PROF_RVQ_EXACT,
  PROF_GATHER_CODES, PROF_OTHER, PROF_NUM_SECTIONS
};
void prof_begin(int section, cudaStream_t s);
void prof_end(int section, cudaStream_t s);
struct ProfScope {
  int sec; cudaStream_t s;
  ProfScope(int section, cudaStream_t st) : sec(section), s(st) { prof_begin(sec, s); }
  ~ProfScope() { prof_end(sec, s); }
};

// ---- programmatic dependent launch -------------------------------------------------------------------------------------
// The pipeline is ~22 short kernels in one stream / graph.  Launched with programmatic stream serialization, kernel n+1
// may start its CTAs (launch latency, barrier init, TMEM allocation) once kernel n's CTAs have signalled
// pdl_launch_dependents(), and blocks in pdl_wait() until kernel n has completed and its writes are visible; every kernel
// launched through launch_pdl() calls pdl_wait() before it touches anything an earlier kernel wrote (device-side counters
// included).  Round 1 measured this SLOWER (1.113 against 1.062 ms per C2 step: ~34 launches then, every wait at the top of
// its kernel, so only the launch latency could overlap).  Round 2, with 22 launches and the two tcgen05 kernels setting up
// their barriers / TMEM before the wait (linear_tc.cu, rvq_tc2.cu): 0.829 against 0.837 ms at C2 (e2e 66.8 against 65.5 M
// faces/s), +1.4 % at C3, +0.7 % at C4, device time unchanged and e2e +2-3 % at 1 024 meshes per launch (A/B on one box each,
// scripts/gpu_ab.sh "MESHTOK_PDL=0" "MESHTOK_PDL=1"); all GPU tests green with it.  On by default; MESHTOK_PDL=0 turns the
// attribute off (pdl_wait / pdl_launch_dependents are then no-ops).
bool pdl_enabled();
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define MT_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                                              \
  do {                                                                                                     \
    ++::meshtok::g_launch_count;                                                                           \
    MT_CHECK_CUDA(::meshtok::launch_pdl(kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__));      \
  } while (0)
#ifndef MESHTOK_PDL_EARLY
#define MESHTOK_PDL_EARLY 0
#endif
// -DMESHTOK_PDL_EARLY=1 (python build.py --variant early -DMESHTOK_PDL_EARLY=1): pdl_wait() also SIGNALS - once every CTA of a kernel has
// passed its wait the next kernel's CTAs may become resident in this kernel's tail and run their launch latency and prologue there.
// (Signalling only AFTER the own wait keeps the invariant every prologue relies on: when kernel n+1 runs anything at all, kernels
// <= n-1 are complete and visible.)  MEASURED at C2 on one box, two runs each: 0.805 / 0.806 ms against 0.806 / 0.805 ms with the signal
// at the end of every CTA's work - the ramps of these kernels are pipeline fill and wave tails, not launch latency.  Off.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#if MESHTOK_PDL_EARLY
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// -DMESHTOK_BOUNDS=1 (python build.py --variant bounds -DMESHTOK_BOUNDS=1, loaded with MESHTOK_LIBRARY=): device-side index checks that trap
// with a message.  compute-sanitizer is closed on the GPU pool this was developed on; the GPU test suite run against this build is the
// memory-safety evidence for the index arithmetic of the kernels that carry the checks.
#if defined(MESHTOK_BOUNDS) && MESHTOK_BOUNDS
#define MT_DCHECK(c)                                                                        \
  do {                                                                                      \
    if (!(c)) {                                                                             \
      printf("MT_DCHECK failed: %s (%s:%d, block %d thread %d)\n", #c, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
      __trap();                                                                             \
    }                                                                                       \
  } while (0)
#else
#define MT_DCHECK(c) do { } while (0)
#endif
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

constexpr int kWarp = 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// Packed fp32 add (sm_100: SASS FADD2): acc.{x,y} += v.{x,y} and acc.{z,w} += v.{z,w}, each element rounded like a scalar add.rn -
// half the issue slots of four FADDs.  The summing kernels (folded embedding tables, neighbour aggregation) are bound by instruction
// issue and by dependent add chains, not by the fp32 pipe.
__device__ __forceinline__ void add4(float4& acc, const float4& v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(&acc);
  const unsigned long long* b = reinterpret_cast<const unsigned long long*>(&v);
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a[0]) : "l"(b[0]));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a[1]) : "l"(b[1]));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_min_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// inclusive prefix sum over the warp
__device__ __forceinline__ int warp_inclusive_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= o) v += n;
  }
  return v;
}

// In-place exclusive scan of data[0..n) held in shared memory by the whole block.
// scratch: >= 33 ints of shared memory.  Returns the total in every thread.
__device__ __forceinline__ int block_exclusive_scan(int* data, int n, int* scratch) {
  const int nt = blockDim.x;
  const int per = (n + nt - 1) / nt;
  const int beg = min(n, (int)threadIdx.x * per);
  const int end = min(n, beg + per);
  int local = 0;
  for (int i = beg; i < end; ++i) local += data[i];
  int inc = warp_inclusive_scan(local);
  if (lane_id() == 31) scratch[warp_id()] = inc;
  __syncthreads();
  if (warp_id() == 0) {
    int nw = (nt + 31) >> 5;
    int w = lane_id() < nw ? scratch[lane_id()] : 0;
    int winc = warp_inclusive_scan(w);
    scratch[lane_id()] = winc - w;
    if (lane_id() == 31) scratch[32] = winc;
  }
  __syncthreads();
  int run = scratch[warp_id()] + inc - local;
  const int total = scratch[32];
  for (int i = beg; i < end; ++i) {
    int v = data[i];
    data[i] = run;
    run += v;
  }
  __syncthreads();
  return total;
}

// workspace layout shared by the pipeline kernels -------------------------------------------------
struct Counts {  // device resident, int32[8]
  int n_faces;   // valid faces in the batch  (packed rows)
  int n_vrows;   // referenced (mesh, vertex) pairs = quantizer rows
  int n_edges;   // directed edges
  int pad[5];
};

}  // namespace meshtok
