// Candidate bookkeeping shared by the tcgen05 RVQ screening kernels (rvq_tc.cu, rvq_tc2.cu) and the re-score kernel.
//
// The screening epilogue reduces every CHUNK consecutive codes to the minimum screened score and keeps the TOPK best
// chunks per (row, partial slot).  The re-score kernel evaluates every code of the kept chunks that lie inside the error
// window exactly; a slot whose LAST kept chunk is still inside the window may hide a further one, and only then does the
// row go to the exact fp32 scan.  CHUNK = 4 keeps the exact re-evaluation at 4 codebook rows (3 KB) per candidate - the
// re-score kernel is bound by those L2 reads, not by arithmetic - and TOPK = 4 makes the exact scan a rare event.
#pragma once
#include <cuda_runtime.h>

namespace meshtok {

constexpr int RVQ_CHUNK = 4;     // rvq_screen32 below is written for 4
constexpr int RVQ_TOPK = 4;
constexpr int RVQ_SLOT_F4 = 2;   // float4 per partial entry: {b0, i0, b1, i1}, {b2, i2, b3, i3}

struct RvqTop {   // scalars (not arrays) so that everything stays in registers
  float b0, b1, b2, b3;
  int i0, i1, i2, i3;
  __device__ __forceinline__ void reset() {
    b0 = b1 = b2 = b3 = INFINITY;
    i0 = i1 = i2 = i3 = 0;
  }
  __device__ __forceinline__ float worst() const { return b3; }
  // called only when m < worst() (rare): replace the last entry and bubble it up.  Strict comparisons: chunks are visited
  // in ascending order, so among equal scores the lower chunk id stays first.
  __device__ __forceinline__ void insert(float m, int idx) {
    b3 = m; i3 = idx;
    if (b3 < b2) { const float t = b2; b2 = b3; b3 = t; const int u = i2; i2 = i3; i3 = u; }
    if (b2 < b1) { const float t = b1; b1 = b2; b2 = t; const int u = i1; i1 = i2; i2 = u; }
    if (b1 < b0) { const float t = b0; b0 = b1; b1 = t; const int u = i0; i0 = i1; i1 = u; }
  }
  __device__ __forceinline__ void store(float4* dst) const {
    dst[0] = make_float4(b0, __int_as_float(i0), b1, __int_as_float(i1));
    dst[1] = make_float4(b2, __int_as_float(i2), b3, __int_as_float(i3));
  }
};

// 32 accumulator columns of one row (acc = <r * rscale, e * escale>; s~ = e2 + cmul * acc) -> 8 chunk minima -> top list.
// One rarely-taken branch per 32 columns guards the list update, so the hot path is FFMA + FMNMX only.
__device__ __forceinline__ void rvq_screen32(const uint32_t (&v)[32], const float* __restrict__ e2, float cmul, int chunk_base,
                                             RvqTop& top) {
  float m[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 e = *reinterpret_cast<const float4*>(e2 + 4 * q);
    m[q] = fminf(fminf(fmaf(__uint_as_float(v[4 * q + 0]), cmul, e.x), fmaf(__uint_as_float(v[4 * q + 1]), cmul, e.y)),
                 fminf(fmaf(__uint_as_float(v[4 * q + 2]), cmul, e.z), fmaf(__uint_as_float(v[4 * q + 3]), cmul, e.w)));
  }
  const float mm = fminf(fminf(fminf(m[0], m[1]), fminf(m[2], m[3])), fminf(fminf(m[4], m[5]), fminf(m[6], m[7])));
  if (mm < top.worst()) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (m[q] < top.worst()) top.insert(m[q], chunk_base + q);
  }
}

}  // namespace meshtok
