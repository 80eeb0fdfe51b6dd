// Candidate bookkeeping shared by the tcgen05 RVQ screening kernels (rvq_tc.cu, rvq_tc2.cu) and the re-score kernel.
//
// The screening epilogue reduces every CHUNK consecutive codes to the minimum screened score and keeps the TOPK best
// chunks per (row, partial slot).  The re-score kernel evaluates every code of the kept chunks that lie inside the error
// window exactly; a slot whose LAST kept chunk is still inside the window may hide a further one, and only then does the
// row go to the exact fp32 scan.  CHUNK = 4 keeps the exact re-evaluation at 4 codebook rows (3 KB) per candidate - the
// re-score kernel is bound by those L2 reads, not by arithmetic - and TOPK = 4 makes the exact scan a rare event.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>

namespace meshtok {

constexpr int RVQ_CHUNK = 4;     // rvq_screen32 below is written for 4
constexpr int RVQ_TOPK = 4;
constexpr int RVQ_SLOT_F4 = 2;
// Codebook image: per 128-code tile, D/64 stages of 16 KB (128 codes x 64 dims fp16, SWIZZLE_128B) followed by one 4 KB
// "extra" K block (128 codes x 16 columns, canonical no-swizzle layout, 8-code groups 256 B apart) that carries
// -kappa * |e_j|^2 as an fp16 pair (hi, lo) in columns 0 and 1.  The CTA-pair kernel multiplies it with a per-row constant
// in the same MMA chain, so the accumulator already holds a multiple of the screened score and the epilogue needs no
// per-value FMA and no |e|^2 loads; the single-CTA kernel skips the block.
constexpr int RVQ_TILE_CODES = 128;
constexpr int RVQ_EXTRA_BYTES = RVQ_TILE_CODES * 16 * 2;   // 4 KB
constexpr unsigned RVQ_EXTRA_SBO = 256;
__host__ __device__ inline size_t rvq_image_tile_bytes(int D) { return (size_t)RVQ_TILE_CODES * D * 2 + RVQ_EXTRA_BYTES; }
// power of two kappa with kappa * max_norm^2 in [0.5, 1)
inline float rvq_e2_kappa(float max_norm) {
  const float m2 = max_norm * max_norm;
  if (!(m2 > 0.f) || !(m2 < 3.0e38f)) return 1.f;
  int e;
  frexpf(m2, &e);
  e = e > 100 ? 100 : (e < -100 ? -100 : e);
  return ldexpf(1.f, -e);
}   // float4 per partial entry: {b0, i0, b1, i1}, {b2, i2, b3, i3}

// power of two s such that max|v| * s lies in [0.5, 1)  (1 for an all-zero row)
__device__ __forceinline__ float rvq_pow2_scale(float vmax) {
  if (!(vmax > 0.f) || !(vmax < INFINITY)) return 1.f;
  int e;
  frexpf(vmax, &e);  // vmax = m * 2^e, m in [0.5, 1)
  e = max(-100, min(100, e));
  return ldexpf(1.f, -e);
}

// One warp leaves one residual row ready for the next screening search: lane l holds columns 2 * (l + 32 i), + 1 in v[i].
// rscale[row] = the row's power-of-two scale; row_image (optional, CTA-pair kernel): fp16(r * scale) at
// [128-row tile]{[64-dim chunk][128 rows x 128 B, SWIZZLE_128B], 4 KB extra K block}, the extra block holding
// c = scale * c_scale in columns 0 and 1 (it multiplies the codebook's (hi, lo) of -kappa |e|^2), zeros elsewhere.
template <int PER2>
__device__ __forceinline__ void rvq_emit_row_scale_and_image(const float2 (&v)[PER2], float vmax_lane, int row, int D, float* __restrict__ rscale,
                                                             uint8_t* __restrict__ row_image, float c_scale) {
  const int lane = threadIdx.x & 31;
  const int d2 = D >> 1;
  float vmax = vmax_lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  float sc = rvq_pow2_scale(vmax);
  // the pair kernel multiplies the folded |e|^2 column by c = sc * c_scale, which must stay a finite fp16 power of two
  if (row_image) sc = fminf(sc, 32768.f / c_scale);
  if (lane == 0 && rscale) rscale[row] = sc;
  if (row_image) {
    uint8_t* tile = row_image + (size_t)(row >> 7) * rvq_image_tile_bytes(D);
    const int r = row & 127;
#pragma unroll
    for (int i = 0; i < PER2; ++i) {
      const int c = lane + 32 * i;
      if (c < d2) {
        const int k = 2 * c;
        const __half2 h = __floats2half2_rn(v[i].x * sc, v[i].y * sc);
        const uint32_t off = (uint32_t)(k >> 6) * (128u * 128u) + (uint32_t)r * 128u + ((uint32_t)(((k >> 3) ^ r) & 7) << 4) + (uint32_t)(k & 7) * 2u;
        *reinterpret_cast<__half2*>(tile + off) = h;
      }
    }
    if (lane < 2) {
      const __half ch = __float2half_rn(sc * c_scale);
      const __half2 cc = __halves2half2(ch, ch);
      uint8_t* ex = tile + (size_t)128 * D * 2 + (size_t)(r >> 3) * RVQ_EXTRA_SBO + (size_t)(r & 7) * 16 + (size_t)lane * 128;
      *reinterpret_cast<uint4*>(ex) = lane == 0 ? make_uint4(*reinterpret_cast<const uint32_t*>(&cc), 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

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

// The same for accumulators that already hold a NEGATIVE multiple of the screened score (|e|^2 folded into the MMA chain):
// s~ = cmul * acc with cmul < 0, so the best codes have the LARGEST acc.  thr = top.worst() / cmul is the accumulator value
// a chunk maximum has to exceed to enter the list; the hot path is 31 FMNMX and one compare per 32 columns.
__device__ __forceinline__ void rvq_screen32_max(const uint32_t (&v)[32], float cmul, float icm, int chunk_base, RvqTop& top, float& thr) {
  float m[8];
#pragma unroll
  for (int q = 0; q < 8; ++q)
    m[q] = fmaxf(fmaxf(__uint_as_float(v[4 * q + 0]), __uint_as_float(v[4 * q + 1])),
                 fmaxf(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])));
  const float mm = fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
  if (mm > thr) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (m[q] > thr) {
        top.insert(cmul * m[q], chunk_base + q);
        thr = top.worst() * icm;
      }
  }
}

}  // namespace meshtok
