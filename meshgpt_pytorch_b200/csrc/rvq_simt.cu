// K8a (exact path): fp32 nearest-codebook search and the residual update of ResidualVQ (eval).
//
// Replaces vector_quantize_pytorch's EuclideanCodebook forward as called through ResidualVQ at
// meshgpt_pytorch.py:842:   dist = -sqrt(clamp(|x|^2 + |e|^2 - 2 x.e, 0)); idx = argmax(dist)  (lowest
// index on ties); r <- r - e[idx].  The reference materialises a (1, R, K) fp32 distance tensor and a
// one-hot of the same size per level; here a 64x64 register-tiled sweep keeps a running (dist, idx) per
// row.  This kernel is (a) the re-check of the rows the tcgen05 screening flags as ambiguous, and
// (b) the exact checker the GPU tests compare the tensor-core path against.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"
#include "rvq_screen.cuh"

namespace meshtok {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

// grid = (row blocks of 64, code splits): every CTA scans `codes_per_split` codes for its 64 rows and merges its
// (distance, index) into packed[row] with a 64-bit atomicMin - distance bits in the high word (non-negative floats
// order like unsigned ints), index in the low word, so equal distances resolve to the lowest index.
__global__ void __launch_bounds__(256) rvq_exact_kernel(const float* __restrict__ rows, int D, const float* __restrict__ cb,
                                                        const float* __restrict__ e2, int K, int codes_per_split,
                                                        const int* __restrict__ row_list, const int* __restrict__ n_dev, int max_rows,
                                                        unsigned long long* __restrict__ packed) {
  pdl_wait();
  __shared__ float As[BK][BM + 4];
  __shared__ float Es[BK][BN + 4];
  __shared__ float x2s[BM];
  __shared__ int rid[BM];
  __shared__ float red_d[BM][17];
  __shared__ int red_i[BM][17];
  const int n = n_dev ? min(*n_dev, max_rows) : max_rows;
  const int tid = threadIdx.x;
  for (int m0 = blockIdx.x * BM; m0 < n; m0 += gridDim.x * BM) {   // row blocks (grid.x may be capped in list mode)
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 2, lk = (tid & 3) * 4;

  if (tid < BM) {
    const int m = m0 + tid;
    rid[tid] = m < n ? (row_list ? row_list[m] : m) : -1;
  }
  __syncthreads();
  {  // |x|^2, 4 threads per row
    const int r = rid[lr];
    float s = 0.f;
    if (r >= 0) {   // independent 16-byte loads (D is a multiple of 16), 4 threads interleaved over the row
      const float4* xr = reinterpret_cast<const float4*>(rows + (size_t)r * D);
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int k4 = (tid & 3); k4 < D / 4; k4 += 4) {
        const float4 v = xr[k4];
        s4[0] = fmaf(v.x, v.x, s4[0]); s4[1] = fmaf(v.y, v.y, s4[1]); s4[2] = fmaf(v.z, v.z, s4[2]); s4[3] = fmaf(v.w, v.w, s4[3]);
      }
      s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if ((tid & 3) == 0) x2s[lr] = s;
  }
  __syncthreads();

  float bd[4];
  int bi[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { bd[i] = INFINITY; bi[i] = 0x7fffffff; }

  const int c_begin = blockIdx.y * codes_per_split;
  const int c_end = min(K, c_begin + codes_per_split);
  for (int n0 = c_begin; n0 < c_end; n0 += BN) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // software pipelined: the slice k0+BK is in flight (registers) while slice k0 is multiplied
    const int r_ld = rid[lr];
    const bool a_ok = r_ld >= 0, e_ok = n0 + lr < c_end;
    const float4* a_src = reinterpret_cast<const float4*>(rows + (size_t)(a_ok ? r_ld : 0) * D + lk);
    const float4* e_src = reinterpret_cast<const float4*>(cb + (size_t)(e_ok ? n0 + lr : 0) * D + lk);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 av_next = a_ok ? a_src[0] : zero4;
    float4 ev_next = e_ok ? __ldg(e_src) : zero4;
    for (int k0 = 0; k0 < D; k0 += BK) {
      const float4 av = av_next, ev = ev_next;
      if (k0 + BK < D) {
        av_next = a_ok ? a_src[(k0 + BK) / 4] : zero4;
        ev_next = e_ok ? __ldg(e_src + (k0 + BK) / 4) : zero4;
      }
      As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
      Es[lk + 0][lr] = ev.x; Es[lk + 1][lr] = ev.y; Es[lk + 2][lr] = ev.z; Es[lk + 3][lr] = ev.w;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 w = *reinterpret_cast<const float4*>(&Es[kk][tx * 4]);
        const float ar[4] = {a.x, a.y, a.z, a.w};
        const float wr[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int code = n0 + tx * 4 + j;
      if (code >= c_end) continue;
      const float y2 = __ldg(e2 + code);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d2 = __fadd_rn(__fadd_rn(x2s[ty * 4 + i], y2), __fmul_rn(acc[i][j], -2.f));
        const float dist = sqrtf(fmaxf(d2, 0.f));
        if (dist < bd[i] || (dist == bd[i] && code < bi[i])) { bd[i] = dist; bi[i] = code; }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { red_d[ty * 4 + i][tx] = bd[i]; red_i[ty * 4 + i][tx] = bi[i]; }
  __syncthreads();
  if (tid < BM && rid[tid] >= 0) {
    float d = red_d[tid][0];
    int ix = red_i[tid][0];
    for (int t = 1; t < 16; ++t) {
      const float dt = red_d[tid][t];
      const int it = red_i[tid][t];
      if (dt < d || (dt == d && it < ix)) { d = dt; ix = it; }
    }
    if (ix == 0x7fffffff) { ix = 0; d = INFINITY; }   // a NaN row: no comparison succeeded; code 0, never an out-of-range index
    const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)ix;
    atomicMin(packed + rid[tid], key);
  }
  }
}

__global__ void rvq_exact_init_kernel(const int* __restrict__ row_list, const int* __restrict__ n_dev, int max_rows,
                                      unsigned long long* __restrict__ packed) {
  const int n = n_dev ? min(*n_dev, max_rows) : max_rows;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) packed[row_list ? row_list[i] : i] = ~0ull;
}

template <int PER2, bool UPDATE>   // float2 pairs per lane: D <= 64 * PER2
__global__ void __launch_bounds__(256) rvq_update_kernel(const float* __restrict__ rows_in, float* __restrict__ rows, int D,
                                                         const float* __restrict__ cb, const unsigned long long* __restrict__ packed,
                                                         int32_t* __restrict__ codes, int Q, int level, float* __restrict__ rscale,
                                                         uint8_t* __restrict__ row_image, float c_scale, const int* __restrict__ n_dev,
                                                         int max_rows, int K) {
  pdl_wait();
  const int wpb = blockDim.x >> 5;
  const int lane = lane_id();
  const int n = n_dev ? min(*n_dev, max_rows) : max_rows;
  const int d2 = D >> 1;
  for (int row = blockIdx.x * wpb + warp_id(); row < n; row += gridDim.x * wpb) {
    int idx = 0;
    if (UPDATE) {
      idx = (int)(unsigned int)(packed[row] & 0xffffffffull);   // low word of (distance bits, index)
      idx = min(max(idx, 0), K - 1);                             // (defensive: the producers never leave an index outside the codebook)
      if (lane == 0) codes[(size_t)row * Q + level] = idx;
    }
    float2 v[PER2];
    float vmax = 0.f;
    const float2* src = reinterpret_cast<const float2*>(rows_in + (size_t)row * D);
    const float2* code = reinterpret_cast<const float2*>(cb + (size_t)idx * D);
    float2* dst = reinterpret_cast<float2*>(rows + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < PER2; ++i) {
      const int c = lane + 32 * i;
      v[i] = make_float2(0.f, 0.f);
      if (c < d2) {
        v[i] = src[c];
        if (UPDATE) {
          const float2 e = __ldg(code + c);
          v[i].x = __fsub_rn(v[i].x, e.x);
          v[i].y = __fsub_rn(v[i].y, e.y);
        }
        dst[c] = v[i];
        vmax = fmaxf(vmax, fmaxf(fabsf(v[i].x), fabsf(v[i].y)));
      }
    }
    rvq_emit_row_scale_and_image<PER2>(v, vmax, row, D, rscale, row_image, c_scale);
  }
  pdl_launch_dependents();
}

// codes[row, level] = winner of packed[row]; the last level of a call that does not return the residual needs no more
__global__ void __launch_bounds__(256) rvq_codes_only_kernel(const unsigned long long* __restrict__ packed, int32_t* __restrict__ codes, int Q,
                                                             int level, const int* __restrict__ n_dev, int max_rows, int K) {
  pdl_wait();
  const int n = n_dev ? min(*n_dev, max_rows) : max_rows;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x)
    codes[(size_t)row * Q + level] = min(max((int)(unsigned int)(packed[row] & 0xffffffffull), 0), K - 1);
  pdl_launch_dependents();
}

}  // namespace

int launch_rvq_codes_only(const unsigned long long* packed, int32_t* codes, int Q, int level, int K, const int* n_rows_dev, int max_rows,
                          cudaStream_t stream) {
  if (max_rows == 0) return MESHTOK_OK;
  const int blocks = max(1, min(ceil_div(max_rows, 256), 148 * 4));
  MT_LAUNCH_PDL(rvq_codes_only_kernel, blocks, 256, 0, stream, packed, codes, Q, level, n_rows_dev, max_rows, K);
  return MESHTOK_OK;
}

int launch_rvq_exact(const float* rows, int D, const float* codebook, const float* e2, int K, const int* row_list,
                     const int* n_dev, int max_rows, unsigned long long* packed, cudaStream_t stream) {
  MT_CHECK_ARG(D % BK == 0, "rvq: dim %d must be a multiple of %d", D, BK);
  if (max_rows == 0) return MESHTOK_OK;
  // few rows (the re-check of ambiguous rows): many code splits so the sweep is short; all rows: fewer, larger splits
  // list mode: the number of rows is a device-side count that is small in practice; cap the grid and loop
  const int row_blocks = row_list ? min(ceil_div(max_rows, BM), 2) : ceil_div(max_rows, BM);
  int splits = row_list ? 1024 : 8;   // list mode is latency bound: one 64-code tile per CTA
  while (splits > 1 && K / splits < BN) splits >>= 1;
  const int codes_per_split = ceil_div(ceil_div(K, splits), BN) * BN;
  splits = ceil_div(K, codes_per_split);
  if (row_list == nullptr) {   // full scan: start from +inf; list mode merges into the re-score kernel's (distance, index)
    rvq_exact_init_kernel<<<ceil_div(max_rows, 256), 256, 0, stream>>>(row_list, n_dev, max_rows, packed);
    MT_CHECK_LAUNCH();
  }
  MT_LAUNCH_PDL(rvq_exact_kernel, dim3(row_blocks, splits), 256, 0, stream, rows, D, codebook, e2, K, codes_per_split, row_list, n_dev, max_rows,
                packed);
  return MESHTOK_OK;
}

int launch_rvq_update(float* rows, int D, const float* codebook, int K, const unsigned long long* packed, int32_t* codes, int Q, int level,
                      float* rscale, void* row_image, float c_scale, const int* n_rows_dev, int max_rows, cudaStream_t stream) {
  MT_CHECK_ARG(D <= 512 && D % 2 == 0, "rvq: dim %d must be even and <= 512", D);
  MT_CHECK_ARG(row_image == nullptr || (D % 64 == 0 && c_scale > 0.f), "rvq: a row image needs dim %% 64 == 0 and c_scale > 0");
  if (max_rows == 0) return MESHTOK_OK;
  const int blocks = max(1, min(ceil_div(max_rows, 8), 148 * 16));
  const int per2 = ceil_div(D / 2, 32);
#define RU(P) MT_LAUNCH_PDL((rvq_update_kernel<P, true>), blocks, 256, 0, stream, rows, rows, D, codebook, packed, codes, Q, level, rscale, static_cast<uint8_t*>(row_image), c_scale, n_rows_dev, max_rows, K)
  if (per2 <= 3) RU(3);
  else RU(8);
#undef RU
  return MESHTOK_OK;
}

int launch_rvq_prep(const float* rows_in, int D, float* resid, float* rscale, void* row_image, float c_scale, const int* n_rows_dev,
                    int max_rows, cudaStream_t stream) {
  MT_CHECK_ARG(D <= 512 && D % 2 == 0, "rvq: dim %d must be even and <= 512", D);
  MT_CHECK_ARG(row_image == nullptr || (D % 64 == 0 && c_scale > 0.f), "rvq: a row image needs dim %% 64 == 0 and c_scale > 0");
  if (max_rows == 0) return MESHTOK_OK;
  const int blocks = max(1, min(ceil_div(max_rows, 8), 148 * 16));
  const int per2 = ceil_div(D / 2, 32);
#define RP(P) MT_LAUNCH_PDL((rvq_update_kernel<P, false>), blocks, 256, 0, stream, rows_in, resid, D, nullptr, nullptr, nullptr, 0, 0, rscale, static_cast<uint8_t*>(row_image), c_scale, n_rows_dev, max_rows, 1)
  if (per2 <= 3) RP(3);
  else RP(8);
#undef RP
  return MESHTOK_OK;
}

}  // namespace meshtok
