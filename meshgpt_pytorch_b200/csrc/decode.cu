// Decode path: the memory-bound stages between the tcgen05 linears of
// MeshAutoencoder.decode_from_codes_to_faces (meshgpt_pytorch.py:898-947) and decode (:872-896).
//
// The decoder is a stack of 1-d convolutions over the face sequence of every mesh.  Rows are UNPACKED here: row = b * nf + n
// for every face slot of the batch, with a per-row validity flag (a padded face is a zero input to every convolution, exactly
// like a position beyond the end of the mesh: meshgpt_pytorch.py:341-349 masks before and after each Conv1d).  A Conv1d with
// kernel k is run as ONE split-bf16 tcgen05 linear (linear_tc.cu) on an im2col activation image: column t * C + c of row n
// holds x[n + t - k/2][c] (zero outside the mesh / on padded faces), the weight (C_out, C_in, k) is re-tiled at load time to
// (C_out, k * C_in) in the same tap-major order.  Everything else is here:
//   dequantize_*      codes -> summed codebook vectors (get_output_from_indices of vector_quantize_pytorch, called at :917)
//   im2col_image      fp32 rows -> split-bf16 im2col image
//   pixelnorm_silu    Block: mask, PixelNorm (F.normalize eps 1e-4, * sqrt(C)), SiLU          (:286-294, :345-351)
//   silu_layernorm    init_decoder_conv tail: SiLU, LayerNorm(eps 1e-5)                       (:621-627)
//   segment_mean      SqueezeExcite: masked mean over the faces of a mesh                      (:313-319)
//   se_gate           Linear -> SiLU -> Linear -> Sigmoid on the (B, C) means                  (:304-310)
//   gate_residual     h * gate[b] + res                                                        (:324, :378-379)
//   argmax_coords     logits -> arg-max bin per coordinate -> undiscretize, NaN on padded faces (:928-942)
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace meshtok {

namespace {

// codes (R * nvf, Q) int64 -> rows (R, nvf * D): sum over the levels of codebook[code]; a negative code contributes zeros
__global__ void __launch_bounds__(256) dequantize_rvq_kernel(const int64_t* __restrict__ codes, const float* __restrict__ codebook, int n_tok, int Q,
                                                             int D, float* __restrict__ out) {
  const int d4 = D >> 2;
  const long long total = (long long)n_tok * d4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long tok = i / d4;
    const int c = (int)(i % d4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < Q; ++q) {
      const long long code = codes[tok * Q + q];
      if (code >= 0) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(codebook + (size_t)code * D) + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    reinterpret_cast<float4*>(out + (size_t)tok * D)[c] = acc;
  }
}

// LFQ: bits of every level -> +-2^-level, summed over the levels, then project_out (D x bits) + bias; one warp per token
__global__ void __launch_bounds__(256) dequantize_lfq_kernel(const int64_t* __restrict__ codes, const float* __restrict__ w_out,
                                                             const float* __restrict__ b_out, int n_tok, int Q, int bits, int D,
                                                             float* __restrict__ out) {
  const int lane = lane_id();
  const int wpb = blockDim.x >> 5;
  for (int tok = blockIdx.x * wpb + warp_id(); tok < n_tok; tok += gridDim.x * wpb) {
    float mine = 0.f;   // lane d keeps dimension d of the summed +-scale vector
    for (int q = 0; q < Q; ++q) {
      const long long code = codes[(size_t)tok * Q + q];
      if (code >= 0 && lane < bits) {
        const float scale = exp2f(-(float)q);
        mine += ((code >> (bits - 1 - lane)) & 1) ? scale : -scale;
      }
    }
    for (int c = lane; c < D; c += 32) {
      float o = b_out[c];
      for (int dd = 0; dd < bits; ++dd) o = fmaf(__shfl_sync(0xffffffffu, mine, dd), __ldg(w_out + (size_t)c * bits + dd), o);
      out[(size_t)tok * D + c] = o;
    }
  }
}

// im2col activation image: (R, k * C) with column t * C + c of row (b, n) = x[(b, n + t - k/2)][c], zero when that face is outside
// the mesh or not valid.  One thread per 4 columns.
__global__ void __launch_bounds__(256) im2col_image_kernel(const float* __restrict__ x, const uint8_t* __restrict__ valid, int R, int nf, int C,
                                                           int k, uint8_t* __restrict__ image) {
  const int c4 = C >> 2, KC = k * C;
  const long long total = (long long)R * k * c4;
  const int half = k >> 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / (k * c4);
    const int rem = (int)(i % (k * c4));
    const int t = rem / c4, c = (rem % c4) * 4;
    const int n = (int)(row % nf);
    const int nn = n + t - half;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nn >= 0 && nn < nf) {
      const long long src = row + (t - half);
      if (valid[src]) v = *reinterpret_cast<const float4*>(x + (size_t)src * C + c);
    }
    tc::image_store4(image, KC, row, t * C + c, v);
  }
}

// y = silu(F.normalize(x, eps) * sqrt(C)) on valid rows, 0 on the others (in place); one warp per row
__global__ void __launch_bounds__(256) pixelnorm_silu_kernel(float* __restrict__ x, const uint8_t* __restrict__ valid, int R, int C, float eps) {
  const int lane = lane_id();
  const int wpb = blockDim.x >> 5;
  const float root_c = sqrtf((float)C);
  for (int row = blockIdx.x * wpb + warp_id(); row < R; row += gridDim.x * wpb) {
    float* xr = x + (size_t)row * C;
    if (!valid[row]) {
      for (int c = lane; c < C; c += 32) xr[c] = 0.f;
      continue;
    }
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) ss = fmaf(xr[c], xr[c], ss);
    ss = warp_sum(ss);
    const float den = fmaxf(sqrtf(ss), eps);
    for (int c = lane; c < C; c += 32) {
      const float t = __fdiv_rn(xr[c], den) * root_c;
      xr[c] = __fdiv_rn(t, 1.f + expf(-t));
    }
  }
}

// y = LayerNorm(silu(x)) (biased variance, eps) in place; one warp per row
__global__ void __launch_bounds__(256) silu_layernorm_kernel(float* __restrict__ x, int R, int C, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps) {
  const int lane = lane_id();
  const int wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + warp_id(); row < R; row += gridDim.x * wpb) {
    float* xr = x + (size_t)row * C;
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float t = xr[c];
      const float s = __fdiv_rn(t, 1.f + expf(-t));
      xr[c] = s;
      sum += s;
    }
    const float mean = warp_sum(sum) / (float)C;
    float var = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; var = fmaf(d, d, var); }
    var = warp_sum(var) / (float)C;
    const float rstd = 1.f / sqrtf(var + eps);
    for (int c = lane; c < C; c += 32) xr[c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
  }
}

// mean[b][c] = sum over the valid faces of mesh b of x[(b, n)][c] / max(#valid, 1e-5); one CTA per (mesh, 32-column group),
// rows summed in ascending order by every thread group, combined in a fixed order (deterministic)
__global__ void __launch_bounds__(256) segment_mean_kernel(const float* __restrict__ x, const uint8_t* __restrict__ valid, int nf, int C,
                                                           float* __restrict__ mean) {
  __shared__ float part[8][32];
  __shared__ int cnt[8];
  const int b = blockIdx.x, c = blockIdx.y * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5;
  float acc = 0.f;
  int n_valid = 0;
  for (int n = g; n < nf; n += 8) {
    const long long row = (long long)b * nf + n;
    if (valid[row]) {
      ++n_valid;
      if (c < C) acc += x[(size_t)row * C + c];
    }
  }
  part[g][threadIdx.x & 31] = acc;
  if ((threadIdx.x & 31) == 0) cnt[g] = n_valid;
  __syncthreads();
  if (g == 0 && c < C) {
    float t = 0.f;
    int nv = 0;
    for (int i = 0; i < 8; ++i) { t += part[i][threadIdx.x & 31]; nv += cnt[i]; }
    mean[(size_t)b * C + c] = t / fmaxf((float)nv, 1e-5f);
  }
}

// gate[b] = sigmoid(W2 silu(W1 mean[b] + b1) + b2); one CTA per mesh (C <= 1024, inner <= 256)
__global__ void __launch_bounds__(256) se_gate_kernel(const float* __restrict__ mean, int C, int inner, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                                      float* __restrict__ gate) {
  __shared__ float sm[1024];
  __shared__ float sh[256];
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sm[c] = mean[(size_t)b * C + c];
  __syncthreads();
  for (int j = threadIdx.x; j < inner; j += blockDim.x) {
    float t = b1[j];
    for (int c = 0; c < C; ++c) t = fmaf(w1[(size_t)j * C + c], sm[c], t);
    sh[j] = __fdiv_rn(t, 1.f + expf(-t));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = b2[c];
    for (int j = 0; j < inner; ++j) t = fmaf(w2[(size_t)c * inner + j], sh[j], t);
    gate[(size_t)b * C + c] = 1.f / (1.f + expf(-t));
  }
}

// out = h * gate[b] + res on valid rows, 0 elsewhere
__global__ void __launch_bounds__(256) gate_residual_kernel(const float* __restrict__ h, const float* __restrict__ gate, const float* __restrict__ res,
                                                            const uint8_t* __restrict__ valid, int R, int nf, int C, float* __restrict__ out) {
  const int c4 = C >> 2;
  const long long total = (long long)R * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / c4;
    const int c = (int)(i % c4);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid[row]) {
      const float4 hv = reinterpret_cast<const float4*>(h + (size_t)row * C)[c];
      const float4 gv = reinterpret_cast<const float4*>(gate + (size_t)(row / nf) * C)[c];
      const float4 rv = reinterpret_cast<const float4*>(res + (size_t)row * C)[c];
      o = make_float4(fmaf(hv.x, gv.x, rv.x), fmaf(hv.y, gv.y, rv.y), fmaf(hv.z, gv.z, rv.z), fmaf(hv.w, gv.w, rv.w));
    }
    reinterpret_cast<float4*>(out + (size_t)row * C)[c] = o;
  }
}

// logits (R, ncoord * nbins) -> arg-max bin per coordinate (first maximum, like torch.argmax) -> bin centre; padded faces: NaN / -1.
// One warp per (row, coordinate).
__global__ void __launch_bounds__(256) argmax_coords_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ valid, int R, int ncoord,
                                                            int nbins, float lo, float hi, float* __restrict__ coords, int64_t* __restrict__ bins) {
  const int lane = lane_id();
  const long long total = (long long)R * ncoord;
  const int wpb = blockDim.x >> 5;
  for (long long item = (long long)blockIdx.x * wpb + warp_id(); item < total; item += (long long)gridDim.x * wpb) {
    const long long row = item / ncoord;
    if (!valid[row]) {
      if (lane == 0) { coords[item] = __int_as_float(0x7fc00000); if (bins) bins[item] = 0; }
      continue;
    }
    const float* lr = logits + item * nbins;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < nbins; j += 32) {
      const float v = lr[j];
      if (v > best || (v == best && j < bi)) { best = v; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      if (bi == 0x7fffffff) bi = 0;   // a row of NaN logits
      // undiscretize (meshgpt_pytorch.py:203-218): ((bin + 0.5) / nbins) * (hi - lo) + lo, in that order
      float t = (float)bi;
      t = __fadd_rn(t, 0.5f);
      t = __fdiv_rn(t, (float)nbins);
      coords[item] = __fadd_rn(__fmul_rn(t, __fsub_rn(hi, lo)), lo);
      if (bins) bins[item] = bi;
    }
  }
}


// ---- fused stages: no im2col ------------------------------------------------------------------------------------------------
// PADDED ROW SPACE.  The fused schedule keeps every activation as rows b * P + lead + n (P = nf + lead, lead = the widest kernel's
// half width), i.e. `lead` separator rows in front of every mesh and after the last one: R = B * P + lead rows, `valid` = 0 on
// separators and padded faces.  Every stage writes zeros to the image rows of non-valid rows, so a convolution tap that reaches over
// the end of a mesh reads a zero row - the zero padding of Conv1d and the masked_fill of Block.forward (meshgpt_pytorch.py:341-349)
// at once, with no per-tap boundary test anywhere.
// HALO IMAGES (tc_common.cuh).  A stage writes its output once, as the split-bf16 halo image of the convolution that follows; the
// tcgen05 kernel reads tap t of the kernel from the same block 16 t bytes further (linear_tc.cu, conv mode).  The first version's
// im2col image (k copies of every activation) is gone: a k=3 convolution's input is written as 4 bytes per element instead of 12.
// Thread layout ("8 x 4"): a warp owns one aligned group of 8 rows, lane = rowsub * 4 + colsub; it walks the row in steps of 32
// columns, 8 adjacent columns (16 bytes of hi, 16 of lo) per lane.  The four lanes of a row sit next to each other, so a quarter
// warp's 16-byte loads fall into 2 cache lines (the first layout, lane = colsub * 8 + rowsub, made the image stores contiguous but
// spread every quarter warp's loads over 8 lines: ncu showed the L1 data pipe at 93 % of its wavefront peak and DRAM at 34 %).
// A row reduction is two shuffles (xor 1, 2).  Needs C % 32 == 0.
// The float chains here use reciprocal multiplies and __expf (PixelNorm / SiLU are tolerance-checked stages, not bit-exact ones);
// the staged kernels above keep the IEEE divisions.
struct Lane84 {
  long long row;   // padded row; may be >= R in the last group
  int b;           // mesh (row / P; == B on the trailing separators)
  int colsub;
  bool active;     // row < R
};
__device__ __forceinline__ Lane84 lane84(long long group, int R, int P) {
  Lane84 l;
  const int lane = lane_id();
  l.row = group * 8 + (lane >> 2);
  l.colsub = lane & 3;
  l.active = l.row < R;
  l.b = (int)(l.row / P);
  return l;
}
__device__ __forceinline__ float row_sum84(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
// 8 adjacent values of (row, columns col .. col + 7) -> the halo image (R, C): the row's own tile and, for the first / last `halo`
// rows of a tile, the halo copy in the neighbouring tile.  v must already be zero for a non-valid row.
__device__ __forceinline__ void emit_halo(uint8_t* img, int C, int halo, long long n_tiles, long long row, int col, const float (&v)[8]) {
  uint4 hi, lo;
  tc::split2(v[0], v[1], hi.x, lo.x);
  tc::split2(v[2], v[3], hi.y, lo.y);
  tc::split2(v[4], v[5], hi.z, lo.z);
  tc::split2(v[6], v[7], hi.w, lo.w);
  const long long tile = row >> 7;
  const int r = (int)(row & 127);
  const uint32_t lo_off = 64u * (uint32_t)(tc::IMG_ROWS + 2 * halo);
  uint8_t* p = img + tc::halo_offset(C, halo, tile, r, col);
  *reinterpret_cast<uint4*>(p) = hi;
  *reinterpret_cast<uint4*>(p + lo_off) = lo;
  if (r < halo && tile > 0) {
    p = img + tc::halo_offset(C, halo, tile - 1, r + tc::IMG_ROWS, col);
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + lo_off) = lo;
  }
  if (r >= tc::IMG_ROWS - halo && tile + 1 < n_tiles) {
    p = img + tc::halo_offset(C, halo, tile + 1, r - tc::IMG_ROWS, col);
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + lo_off) = lo;
  }
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float silu_fast(float t) { return __fdividef(t, 1.f + __expf(-t)); }

// x (B * nf rows of C floats, UNPADDED, row stride ld) -> halo image in the padded row space; zero rows on separators / masked faces
__global__ void __launch_bounds__(256) rows_halo_kernel(const float* __restrict__ x, int ld, const uint8_t* __restrict__ valid, int R, int P, int lead,
                                                        int C, int halo, uint8_t* __restrict__ image) {
  const long long groups = ((long long)R + 7) >> 3, n_tiles = ((long long)R + 127) >> 7;
  const int nf = P - lead;
  for (long long g = (long long)blockIdx.x * 8 + warp_id(); g < groups; g += (long long)gridDim.x * 8) {
    const Lane84 l = lane84(g, R, P);
    if (!l.active) continue;
    const bool ok = valid[l.row] != 0;
    const long long src = (long long)l.b * nf + (l.row - (long long)l.b * P - lead);
    for (int col = l.colsub * 8; col < C; col += 32) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (ok) load8(x + (size_t)src * ld + col, v);
      emit_halo(image, C, halo, n_tiles, l.row, col, v);
    }
  }
}

// codes (B * nf * nvf, Q; unpadded) -> summed codebook rows (nvf * D per face) -> halo image of the first convolution; never stored as fp32
__global__ void __launch_bounds__(256) dequantize_rvq_halo_kernel(const int64_t* __restrict__ codes, const float* __restrict__ codebook,
                                                                  const uint8_t* __restrict__ valid, int R, int P, int lead, int nvf, int Q, int D, int halo,
                                                                  uint8_t* __restrict__ image) {
  const int C = nvf * D, nf = P - lead;
  const long long groups = ((long long)R + 7) >> 3, n_tiles = ((long long)R + 127) >> 7;
  for (long long g = (long long)blockIdx.x * 8 + warp_id(); g < groups; g += (long long)gridDim.x * 8) {
    const Lane84 l = lane84(g, R, P);
    if (!l.active) continue;
    const bool ok = valid[l.row] != 0;
    const long long face = (long long)l.b * nf + (l.row - (long long)l.b * P - lead);
    for (int col = l.colsub * 8; col < C; col += 32) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (ok) {
        const int slot = col / D, d = col - slot * D;
        const long long tok = face * nvf + slot;
        for (int q = 0; q < Q; ++q) {
          const long long code = codes[tok * Q + q];
          if (code >= 0) {
            float w[8];
            load8(codebook + (size_t)code * D + d, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += w[i];
          }
        }
      }
      emit_halo(image, C, halo, n_tiles, l.row, col, v);
    }
  }
}

// init_decoder_conv tail + the first block's image: y = LayerNorm(silu(x)) -> out (fp32, the residual input) and halo image.
// Non-valid rows: zeros in both.
__global__ void __launch_bounds__(256) silu_layernorm_halo_kernel(const float* __restrict__ x, const uint8_t* __restrict__ valid, int R, int P, int C,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                                  float* __restrict__ out, int halo, uint8_t* __restrict__ image) {
  const long long groups = ((long long)R + 7) >> 3, n_tiles = ((long long)R + 127) >> 7;
  const float inv_c = 1.f / (float)C;
  for (long long g = (long long)blockIdx.x * 8 + warp_id(); g < groups; g += (long long)gridDim.x * 8) {
    const Lane84 l = lane84(g, R, P);
    const bool ok = l.active && valid[l.row] != 0;
    const float* xr = x + (size_t)(l.active ? l.row : 0) * C;
    float sum = 0.f;
    if (ok)
      for (int col = l.colsub * 8; col < C; col += 32) {
        float v[8];
        load8(xr + col, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += silu_fast(v[i]);
      }
    const float mean = row_sum84(sum) * inv_c;
    float var = 0.f;
    if (ok)
      for (int col = l.colsub * 8; col < C; col += 32) {
        float v[8];
        load8(xr + col, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = silu_fast(v[i]) - mean; var = fmaf(d, d, var); }
      }
    var = row_sum84(var) * inv_c;
    const float rstd = rsqrtf(var + eps);
    if (!l.active) continue;
    for (int col = l.colsub * 8; col < C; col += 32) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (ok) {
        float gm[8], bt[8];
        load8(xr + col, v);
        load8(gamma + col, gm);
        load8(beta + col, bt);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf((silu_fast(v[i]) - mean) * rstd, gm[i], bt[i]);
      }
      store8(out + (size_t)l.row * C + col, v);
      if (image) emit_halo(image, C, halo, n_tiles, l.row, col, v);
    }
  }
}

// Block tail feeding the next convolution: silu(PixelNorm(h)) written ONLY as the halo image (h: row stride ld, left untouched)
__global__ void __launch_bounds__(256) pixelnorm_silu_halo_kernel(const float* __restrict__ h, int ld, const uint8_t* __restrict__ valid, int R, int P,
                                                                  int C, float eps, int halo, uint8_t* __restrict__ image) {
  const long long groups = ((long long)R + 7) >> 3, n_tiles = ((long long)R + 127) >> 7;
  const float root_c = sqrtf((float)C);
  for (long long g = (long long)blockIdx.x * 8 + warp_id(); g < groups; g += (long long)gridDim.x * 8) {
    const Lane84 l = lane84(g, R, P);
    const bool ok = l.active && valid[l.row] != 0;
    const float* hr = h + (size_t)(l.active ? l.row : 0) * ld;
    float ss = 0.f;
    if (ok)
      for (int col = l.colsub * 8; col < C; col += 32) {
        float v[8];
        load8(hr + col, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) ss = fmaf(v[i], v[i], ss);
      }
    const float scale = root_c / fmaxf(sqrtf(row_sum84(ss)), eps);
    if (!l.active) continue;
    for (int col = l.colsub * 8; col < C; col += 32) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (ok) {
        load8(hr + col, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = silu_fast(v[i] * scale);
      }
      emit_halo(image, C, halo, n_tiles, l.row, col, v);
    }
  }
}

// The same with the whole row slice of a lane in registers (C <= 32 * NIT): every load of the row is in flight before the first use
// and h is read once.  The loop version above keeps one 32-column step in flight per warp and reads h twice; with all warps of the
// grid resident in a single wave its run time is the sum of a warp's dependent memory round trips (ncu: DRAM at 35 % of peak).
template <int NIT>
__global__ void __launch_bounds__(256) pixelnorm_silu_halo_rows_kernel(const float* __restrict__ h, int ld, const uint8_t* __restrict__ valid, int R,
                                                                       int P, int C, float eps, int halo, uint8_t* __restrict__ image) {
  const long long groups = ((long long)R + 7) >> 3, n_tiles = ((long long)R + 127) >> 7;
  const float root_c = sqrtf((float)C);
  for (long long g = (long long)blockIdx.x * 8 + warp_id(); g < groups; g += (long long)gridDim.x * 8) {
    const Lane84 l = lane84(g, R, P);
    const bool ok = l.active && valid[l.row] != 0;
    const float* hr = h + (size_t)(l.active ? l.row : 0) * ld + l.colsub * 8;
    float v[NIT][8];
#pragma unroll
    for (int u = 0; u < NIT; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[u][i] = 0.f;
      if (ok && u * 32 < C) load8(hr + u * 32, v[u]);
    }
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < NIT; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(v[u][i], v[u][i], ss);
    const float scale = root_c / fmaxf(sqrtf(row_sum84(ss)), eps);
    if (!l.active) continue;
#pragma unroll
    for (int u = 0; u < NIT; ++u) {
      if (u * 32 >= C) break;
      if (ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[u][i] = silu_fast(v[u][i] * scale);
      }
      emit_halo(image, C, halo, n_tiles, l.row, l.colsub * 8 + u * 32, v[u]);
    }
  }
}

// Second Block tail + SqueezeExcite's mean, without storing the activation: scale[row] = sqrt(C) / max(||h[row]||, eps) and, per
// (mesh, slice of its faces), the column sums of silu(h * scale) over the valid faces of the slice.  CTA (mesh, slice); a warp
// per face, lane owns columns lane * 4 + 128 j; rows are added in ascending order per warp and the 8 warps are combined in a fixed
// order, so the result does not depend on scheduling.  Rows of mesh b: b * P + lead + n.
constexpr int SE_MAX_C = 1024;
__device__ __forceinline__ void se_gate_tail(int b, int T, const float* __restrict__ part, const int* __restrict__ part_cnt, int S, int C, int inner,
                                             const float* __restrict__ w1t, const float* __restrict__ b1, const float* __restrict__ w2t,
                                             const float* __restrict__ b2, float* __restrict__ gate, float* sm, float* sq, float* sh);
template <int NJ>   // NJ = ceil(C / 128) float4 per lane
__global__ void __launch_bounds__(256) pixelnorm_colsum_kernel(const float* __restrict__ h, int ld, const uint8_t* __restrict__ valid, int P, int lead, int C,
                                                               float eps, int S, float* __restrict__ scale_out, float* __restrict__ part,
                                                               int* __restrict__ part_cnt, int* __restrict__ tickets, int inner,
                                                               const float* __restrict__ w1t, const float* __restrict__ b1,
                                                               const float* __restrict__ w2t, const float* __restrict__ b2, float* __restrict__ gate) {
  extern __shared__ float4 sm_dyn[];   // [8 warps][NJ * 32]; the gate tail reuses it (C + 256 + 256 floats; the launcher sizes it for both)
  __shared__ int cnt[8];
  __shared__ int is_last;
  const int b = blockIdx.x, s = blockIdx.y, lane = lane_id(), w = warp_id();
  const int nf = P - lead;
  const int per = (nf + S - 1) / S;
  const int n0 = s * per, n1 = min(nf, n0 + per);
  const long long base = (long long)b * P + lead;
  const float root_c = sqrtf((float)C);
  float4 acc[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  int n_valid = 0;
  // the next row's loads are issued before this row's reduction (one row of latency hidden per warp)
  float4 nxt[NJ];
  bool nxt_ok = false;
  auto fetch = [&](int n) {
    nxt_ok = n < n1 && valid[base + n] != 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = lane * 4 + j * 128;
      nxt[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (nxt_ok && c < C) nxt[j] = *reinterpret_cast<const float4*>(h + (size_t)(base + n) * ld + c);
    }
  };
  fetch(n0 + w);
  for (int n = n0 + w; n < n1; n += 8) {
    float4 v[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) v[j] = nxt[j];
    const bool ok = nxt_ok;
    fetch(n + 8);
    if (!ok) continue;      // scale is only read on valid rows
    ++n_valid;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      ss = fmaf(v[j].x, v[j].x, ss); ss = fmaf(v[j].y, v[j].y, ss); ss = fmaf(v[j].z, v[j].z, ss); ss = fmaf(v[j].w, v[j].w, ss);
    }
    const float scale = root_c / fmaxf(sqrtf(warp_sum(ss)), eps);
    if (lane == 0) scale_out[base + n] = scale;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {   // padding columns hold 0: silu(0) = 0
      acc[j].x += silu_fast(v[j].x * scale);
      acc[j].y += silu_fast(v[j].y * scale);
      acc[j].z += silu_fast(v[j].z * scale);
      acc[j].w += silu_fast(v[j].w * scale);
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) sm_dyn[w * (NJ * 32) + lane + j * 32] = acc[j];
  if (lane == 0) cnt[w] = n_valid;
  __syncthreads();
  float* out = part + ((size_t)b * S + s) * C;
  for (int c4 = threadIdx.x; c4 < (C >> 2); c4 += blockDim.x) {
    float4 t = sm_dyn[c4];
    for (int i = 1; i < 8; ++i) { const float4 o = sm_dyn[i * (NJ * 32) + c4]; t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w; }
    reinterpret_cast<float4*>(out)[c4] = t;
  }
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < 8; ++i) t += cnt[i];
    part_cnt[b * S + s] = t;
  }
  if (tickets == nullptr) return;
  // The CTA that finishes a mesh's last slice computes the mesh's gate (a launch of its own cost 12 us of ramp per ResnetBlock):
  // every thread's slice sums are fenced to the device, then one ticket per CTA; the ticket counter is left at 0 for the next launch.
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(tickets + b, 1);
    is_last = t == S - 1;
    if (is_last) tickets[b] = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float* fs = reinterpret_cast<float*>(sm_dyn);
  se_gate_tail(b, 256, part, part_cnt, S, C, inner, w1t, b1, w2t, b2, gate, fs, fs + C, fs + C + 256);
}

// gate[b] = sigmoid(W2 silu(W1 mean[b] + b1) + b2) with mean[b] = (sum of the S slice sums) / max(#valid faces, 1e-5), by one CTA
// of T threads (T a multiple of 32, inner <= T).  A chain of three dependent global-memory round trips (slice sums, W1, W2), so every
// phase is spread over all threads with all of a thread's loads independent: both weight matrices are read TRANSPOSED (w1t (C,
// inner), w2t (inner, C), built once by the host side: a warp reads consecutive floats); in the first linear a thread owns (part of
// the reduction range, hidden unit) and the parts are added in a fixed order through shared memory.
// sm: C floats, sq: T floats, sh: 256 floats of shared memory.  part / part_cnt were written by other CTAs: read through L2.
__device__ __forceinline__ void se_gate_tail(int b, int T, const float* __restrict__ part, const int* __restrict__ part_cnt, int S, int C, int inner,
                                             const float* __restrict__ w1t, const float* __restrict__ b1, const float* __restrict__ w2t,
                                             const float* __restrict__ b2, float* __restrict__ gate, float* sm, float* sq, float* sh) {
  const int tid = threadIdx.x;
  int nv = 0;
  for (int s = 0; s < S; ++s) nv += __ldcg(part_cnt + b * S + s);
  const float inv = 1.f / fmaxf((float)nv, 1e-5f);
  for (int c = tid; c < C; c += T) {
    float t = 0.f;
    for (int s = 0; s < S; ++s) t += __ldcg(part + ((size_t)b * S + s) * C + c);
    sm[c] = t * inv;
  }
  __syncthreads();
  {  // hidden = silu(W1 mean + b1): thread (part p of the C range, hidden unit j)
    const int jn = (inner + 31) & ~31, parts = T / jn, p = tid / jn, j = tid - p * jn;
    const int span = (C + parts - 1) / parts;
    float t = 0.f;
    if (p < parts && j < inner) {
      const int c1 = min(C, (p + 1) * span);
#pragma unroll 16
      for (int c = p * span; c < c1; ++c) t = fmaf(__ldg(w1t + (size_t)c * inner + j), sm[c], t);
    }
    sq[tid] = t;
    __syncthreads();
    if (tid < inner) {
      float a = b1[tid];
      for (int q = 0; q < parts; ++q) a += sq[q * jn + tid];
      sh[tid] = silu_fast(a);
    }
    __syncthreads();
  }
  // gate = sigmoid(W2 hidden + b2): a thread per column, `inner` (<= 256) independent coalesced loads each
  for (int c = tid; c < C; c += T) {
    float a = b2[c];
#pragma unroll 16
    for (int j = 0; j < inner; ++j) a = fmaf(__ldg(w2t + (size_t)j * C + c), sh[j], a);
    gate[(size_t)b * C + c] = __fdividef(1.f, 1.f + __expf(-a));
  }
}

// stand-alone launch of the gate (one CTA of 1024 threads per mesh); the fused schedule runs se_gate_tail in the last column-sum CTA
__global__ void __launch_bounds__(1024) se_gate_slices_kernel(const float* __restrict__ part, const int* __restrict__ part_cnt, int S, int C, int inner,
                                                              const float* __restrict__ w1t, const float* __restrict__ b1, const float* __restrict__ w2t,
                                                              const float* __restrict__ b2, float* __restrict__ gate) {
  __shared__ float sm[SE_MAX_C];
  __shared__ float sq[1024];
  __shared__ float sh[256];
  se_gate_tail(blockIdx.x, 1024, part, part_cnt, S, C, inner, w1t, b1, w2t, b2, gate, sm, sq, sh);
}

// ResnetBlock tail + the next convolution's image: out = silu(h * scale[row]) * gate[b] + res (scale == NULL: h * gate[b] + res)
// on valid rows, zeros elsewhere; out as fp32 (the next residual input) and, when image != NULL, as a halo image.
__global__ void __launch_bounds__(256) gate_residual_halo_kernel(const float* __restrict__ h, int ldh, const float* __restrict__ scale,
                                                                 const float* __restrict__ gate, const float* __restrict__ res, int ldres,
                                                                 const uint8_t* __restrict__ valid, int R, int P, int C, float* __restrict__ out, int halo,
                                                                 uint8_t* __restrict__ image) {
  const long long groups = ((long long)R + 7) >> 3, n_tiles = ((long long)R + 127) >> 7;
  for (long long g = (long long)blockIdx.x * 8 + warp_id(); g < groups; g += (long long)gridDim.x * 8) {
    const Lane84 l = lane84(g, R, P);
    if (!l.active) continue;
    const bool ok = valid[l.row] != 0;
    const float sc = (ok && scale) ? scale[l.row] : 1.f;
    // U = 4 steps of 32 columns per round: the 16 loads of a round are issued before the first use (one step per round left a warp
    // with 2 KB in flight and the kernel, a single wave of warps, bound by the sum of its memory round trips)
    constexpr int U = 4;
    for (int col0 = l.colsub * 8; col0 < C; col0 += 32 * U) {
      float v[U][8], rs[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[u][i] = 0.f; rs[u][i] = 0.f; }
        if (ok && col0 + 32 * u < C) {
          load8(h + (size_t)l.row * ldh + col0 + 32 * u, v[u]);
          load8(res + (size_t)l.row * ldres + col0 + 32 * u, rs[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int col = col0 + 32 * u;
        if (col >= C) break;
        if (ok) {
          float gt[8];
          load8(gate + (size_t)l.b * C + col, gt);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[u][i] = fmaf(scale ? silu_fast(v[u][i] * sc) : v[u][i], gt[i], rs[u][i]);
        }
        store8(out + (size_t)l.row * C + col, v[u]);
        if (image) emit_halo(image, C, halo, n_tiles, l.row, col, v[u]);
      }
    }
  }
}

// logits in the padded row space -> coordinates / bins of the UNPADDED faces (b * nf + n); same arithmetic as argmax_coords_kernel
__global__ void __launch_bounds__(256) argmax_coords_padded_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ valid, int B, int P, int lead,
                                                                   int ncoord, int nbins, float lo, float hi, float* __restrict__ coords,
                                                                   int64_t* __restrict__ bins) {
  const int lane = lane_id(), nf = P - lead;
  const long long total = (long long)B * nf * ncoord;
  const int wpb = blockDim.x >> 5;
  for (long long item = (long long)blockIdx.x * wpb + warp_id(); item < total; item += (long long)gridDim.x * wpb) {
    const long long face = item / ncoord;
    const int coord = (int)(item - face * ncoord);
    const int b = (int)(face / nf);
    const long long row = (long long)b * P + lead + (face - (long long)b * nf);
    if (!valid[row]) {
      if (lane == 0) { coords[item] = __int_as_float(0x7fc00000); if (bins) bins[item] = 0; }
      continue;
    }
    const float* lr = logits + ((size_t)row * ncoord + coord) * nbins;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < nbins; j += 32) {
      const float v = lr[j];
      if (v > best || (v == best && j < bi)) { best = v; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      if (bi == 0x7fffffff) bi = 0;
      float t = (float)bi;
      t = __fadd_rn(t, 0.5f);
      t = __fdiv_rn(t, (float)nbins);
      coords[item] = __fadd_rn(__fmul_rn(t, __fsub_rn(hi, lo)), lo);
      if (bins) bins[item] = bi;
    }
  }
}

// nbins == 128: a warp per face, lane owns 4 adjacent bins of every coordinate (one 16-byte load), U coordinates' loads in flight at
// a time.  (The kernel above, a warp per (face, coordinate) with scalar loads, kept 512 bytes in flight per warp: 1.8 TB/s.)
__global__ void __launch_bounds__(256) argmax_coords_padded128_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ valid, int B, int P,
                                                                      int lead, int ncoord, float lo, float hi, float* __restrict__ coords,
                                                                      int64_t* __restrict__ bins) {
  constexpr int U = 3;
  const int lane = lane_id(), nf = P - lead;
  const long long total = (long long)B * nf;
  for (long long face = (long long)blockIdx.x * 8 + warp_id(); face < total; face += (long long)gridDim.x * 8) {
    const int b = (int)(face / nf);
    const long long row = (long long)b * P + lead + (face - (long long)b * nf);
    if (!valid[row]) {
      for (int k = lane; k < ncoord; k += 32) { coords[face * ncoord + k] = __int_as_float(0x7fc00000); if (bins) bins[face * ncoord + k] = 0; }
      continue;
    }
    const float4* lr = reinterpret_cast<const float4*>(logits + (size_t)row * ncoord * 128) + lane;
    for (int k0 = 0; k0 < ncoord; k0 += U) {
      float4 x[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (k0 + u < ncoord) x[u] = __ldcs(lr + (size_t)(k0 + u) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k0 + u >= ncoord) break;
        float best = -INFINITY;                    // first maximum wins, like torch.argmax; same comparisons as the scalar kernel
        int bi = 0x7fffffff;
        const float xv[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (xv[i] > best || (xv[i] == best && lane * 4 + i < bi)) { best = xv[i]; bi = lane * 4 + i; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) {
          if (bi == 0x7fffffff) bi = 0;
          float t = (float)bi;
          t = __fadd_rn(t, 0.5f);
          t = __fdiv_rn(t, 128.f);
          coords[face * ncoord + k0 + u] = __fadd_rn(__fmul_rn(t, __fsub_rn(hi, lo)), lo);
          if (bins) bins[face * ncoord + k0 + u] = bi;
        }
      }
    }
  }
}

inline unsigned grid_groups(long long R) {
  const long long want = (((R + 7) >> 3) + 7) >> 3;
  return (unsigned)(want < 1 ? 1 : (want > 148ll * 8 ? 148ll * 8 : want));
}

inline unsigned grid_for(long long items, int per_block) {
  const long long want = (items + per_block - 1) / per_block;
  return (unsigned)(want < 1 ? 1 : (want > 148ll * 16 ? 148ll * 16 : want));
}

}  // namespace
}  // namespace meshtok

using namespace meshtok;

extern "C" {

int meshtok_dec_dequantize_rvq(const int64_t* codes, const float* codebook, int n_tokens, int Q, int D, float* out, meshtok_stream_t stream) {
  MT_CHECK_ARG(codes && codebook && out && n_tokens >= 0 && Q >= 1 && D % 4 == 0, "dequantize: bad arguments");
  if (n_tokens == 0) return MESHTOK_OK;
  dequantize_rvq_kernel<<<grid_for((long long)n_tokens * (D / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(codes, codebook, n_tokens, Q, D, out);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_dec_dequantize_lfq(const int64_t* codes, const float* w_out, const float* b_out, int n_tokens, int Q, int bits, int D, float* out,
                               meshtok_stream_t stream) {
  MT_CHECK_ARG(codes && w_out && b_out && out && n_tokens >= 0 && Q >= 1 && bits >= 1 && bits <= 32, "dequantize: bad arguments");
  if (n_tokens == 0) return MESHTOK_OK;
  dequantize_lfq_kernel<<<grid_for(n_tokens, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(codes, w_out, b_out, n_tokens, Q, bits, D, out);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_dec_im2col_image(const float* x, const uint8_t* valid, int R, int nf, int C, int k, void* image, meshtok_stream_t stream) {
  MT_CHECK_ARG(x && valid && image && R >= 0 && nf >= 1 && C % 4 == 0 && (k & 1) == 1 && (k * C) % tc::IMG_KC == 0,
               "im2col: width %d x kernel %d must be a multiple of %d, kernel odd", C, k, tc::IMG_KC);
  if (R == 0) return MESHTOK_OK;
  im2col_image_kernel<<<grid_for((long long)R * k * (C / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, valid, R, nf, C, k, static_cast<uint8_t*>(image));
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_dec_pixelnorm_silu(float* x, const uint8_t* valid, int R, int C, float eps, meshtok_stream_t stream) {
  MT_CHECK_ARG(x && valid && R >= 0 && C >= 1, "pixelnorm: bad arguments");
  if (R == 0) return MESHTOK_OK;
  pixelnorm_silu_kernel<<<grid_for(R, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, valid, R, C, eps);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_dec_silu_layernorm(float* x, int R, int C, const float* gamma, const float* beta, float eps, meshtok_stream_t stream) {
  MT_CHECK_ARG(x && gamma && beta && R >= 0 && C >= 1, "layernorm: bad arguments");
  if (R == 0) return MESHTOK_OK;
  silu_layernorm_kernel<<<grid_for(R, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, R, C, gamma, beta, eps);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_dec_squeeze_excite_gate(const float* x, const uint8_t* valid, int B, int nf, int C, int inner, const float* w1, const float* b1,
                                    const float* w2, const float* b2, float* mean_ws, float* gate, meshtok_stream_t stream) {
  MT_CHECK_ARG(x && valid && w1 && b1 && w2 && b2 && mean_ws && gate && C <= 1024 && inner <= 256, "squeeze-excite: bad arguments (C <= 1024, inner <= 256)");
  if (B == 0) return MESHTOK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  segment_mean_kernel<<<dim3(B, (C + 31) / 32), 256, 0, s>>>(x, valid, nf, C, mean_ws);
  MT_CHECK_LAUNCH();
  se_gate_kernel<<<B, 256, 0, s>>>(mean_ws, C, inner, w1, b1, w2, b2, gate);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_dec_gate_residual(const float* h, const float* gate, const float* res, const uint8_t* valid, int R, int nf, int C, float* out,
                              meshtok_stream_t stream) {
  MT_CHECK_ARG(h && gate && res && valid && out && C % 4 == 0, "gate_residual: bad arguments");
  if (R == 0) return MESHTOK_OK;
  gate_residual_kernel<<<grid_for((long long)R * (C / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(h, gate, res, valid, R, nf, C, out);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_dec_argmax_coords(const float* logits, const uint8_t* valid, int R, int ncoord, int nbins, float lo, float hi, float* coords,
                              int64_t* bins, meshtok_stream_t stream) {
  MT_CHECK_ARG(logits && valid && coords && ncoord >= 1 && nbins >= 1, "argmax_coords: bad arguments");
  if (R == 0) return MESHTOK_OK;
  argmax_coords_kernel<<<grid_for((long long)R * ncoord, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, valid, R, ncoord, nbins, lo, hi, coords, bins);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

/* ---- fused stages: padded row space, halo images (see the comment above the kernels) ---- */
int meshtok_dec_se_slices(int B, int nf) {
  if (B < 1 || nf < 1) return 1;
  int S = (148 * 4) / B;               /* one wave of four 8-warp CTAs per SM */
  const int most = (nf + 15) / 16;     /* at least 16 faces per slice */
  if (S > most) S = most;
  if (S > 64) S = 64;
  return S < 1 ? 1 : S;
}

int meshtok_halo_image_bytes(int R, int C, int halo, size_t* bytes) {
  MT_CHECK_ARG(bytes && R >= 0 && C > 0 && C % 32 == 0 && halo >= 0 && halo <= 8, "halo image: width %d must be a multiple of 32, halo %d in [0, 8]", C, halo);
  *bytes = tc::halo_image_bytes(R, C, halo);
  return MESHTOK_OK;
}

int meshtok_conv1d(const void* a_halo_image, const void* w_image, const float* bias, float* out, int R, int N, int C_in, int k, int halo,
                   meshtok_stream_t stream) {
  MT_CHECK_ARG(a_halo_image && w_image && out && R >= 0 && N > 0 && C_in > 0, "conv1d: bad arguments");
  return launch_conv_tc(a_halo_image, C_in, k, halo, w_image, bias, out, N, R, static_cast<cudaStream_t>(stream));
}

#define MT_ROWSPACE_OK(R, P, lead) ((R) >= 0 && (lead) >= 0 && (P) > (lead) && ((R) - (lead)) % (P) == 0)

/* Conv1d + the Block tail in the convolution's epilogue: the first `width` output columns (block1.proj / block2.proj) leave the kernel only
 * as the halo image of silu(PixelNorm(row)); columns width .. N (residual_conv riding in the same launch, N == 2 * width) go to `out`. */
int meshtok_conv1d_pixelnorm_silu(const void* a_halo_image, const void* w_image, const float* bias, const uint8_t* valid, float* out, int R, int N,
                                  int C_in, int k, int halo, int width, float eps, int out_halo, void* out_image, meshtok_stream_t stream) {
  MT_CHECK_ARG(a_halo_image && w_image && valid && out_image && R >= 0 && N > 0 && C_in > 0 && (N == width || (N == 2 * width && out)),
               "conv1d_pixelnorm_silu: bad arguments (N must be width, or 2 * width with an fp32 output for the second half)");
  ConvBlockTail tail{width, valid, out_image, out_halo, eps};
  return launch_conv_tc(a_halo_image, C_in, k, halo, w_image, bias, out, N, R, static_cast<cudaStream_t>(stream), &tail);
}

/* to_coor_logits + arg-max + undiscretize in one launch: the logits stay in TMEM */
int meshtok_conv1d_argmax_coords(const void* a_halo_image, const void* w_image, const float* bias, const uint8_t* valid, int R, int N, int C_in, int k,
                                 int halo, int B, int P, int lead, int ncoord, int nbins, float lo, float hi, float* coords, int64_t* bins,
                                 meshtok_stream_t stream) {
  MT_CHECK_ARG(a_halo_image && w_image && valid && coords && MT_ROWSPACE_OK(R, P, lead) && (long long)B * P + lead == R && N > 0 && C_in > 0,
               "conv1d_argmax_coords: bad arguments (R must be B * P + lead)");
  ConvArgmax am{valid, B, P, lead, ncoord, nbins, lo, hi, coords, bins};
  return launch_conv_argmax_tc(a_halo_image, C_in, k, halo, w_image, bias, N, R, am, static_cast<cudaStream_t>(stream));
}

int meshtok_dec_rows_halo(const float* x, int ld, const uint8_t* valid, int R, int P, int lead, int C, int halo, void* image, meshtok_stream_t stream) {
  MT_CHECK_ARG(x && valid && image && MT_ROWSPACE_OK(R, P, lead) && C % 32 == 0 && ld >= C && ld % 4 == 0 && halo >= 0 && halo <= lead,
               "rows_halo: width %d must be a multiple of 32, R = B * P + lead, halo <= lead", C);
  if (R == 0) return MESHTOK_OK;
  rows_halo_kernel<<<grid_groups(R), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ld, valid, R, P, lead, C, halo, static_cast<uint8_t*>(image));
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_dec_dequantize_rvq_halo(const int64_t* codes, const float* codebook, const uint8_t* valid, int R, int P, int lead, int nvf, int Q, int D,
                                    int halo, void* image, meshtok_stream_t stream) {
  MT_CHECK_ARG(codes && codebook && valid && image && MT_ROWSPACE_OK(R, P, lead) && nvf >= 1 && Q >= 1 && D % 8 == 0 && (nvf * D) % 32 == 0 && halo >= 0 &&
                   halo <= lead,
               "dequantize_rvq_halo: D %d must be a multiple of 8, nvf * D of 32, R = B * P + lead, halo <= lead", D);
  if (R == 0) return MESHTOK_OK;
  dequantize_rvq_halo_kernel<<<grid_groups(R), 256, 0, static_cast<cudaStream_t>(stream)>>>(codes, codebook, valid, R, P, lead, nvf, Q, D, halo,
                                                                                           static_cast<uint8_t*>(image));
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_dec_silu_layernorm_halo(const float* x, const uint8_t* valid, int R, int P, int C, const float* gamma, const float* beta, float eps,
                                    float* out, int halo, void* image, meshtok_stream_t stream) {
  MT_CHECK_ARG(x && valid && gamma && beta && out && R >= 0 && P >= 1 && C % 32 == 0 && halo >= 0 && halo <= 8, "silu_layernorm_halo: width %d must be a multiple of 32", C);
  if (R == 0) return MESHTOK_OK;
  silu_layernorm_halo_kernel<<<grid_groups(R), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, valid, R, P, C, gamma, beta, eps, out, halo,
                                                                                           static_cast<uint8_t*>(image));
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_dec_pixelnorm_silu_halo(const float* h, int ld, const uint8_t* valid, int R, int P, int C, float eps, int halo, void* image,
                                    meshtok_stream_t stream) {
  MT_CHECK_ARG(h && valid && image && R >= 0 && P >= 1 && C % 32 == 0 && ld >= C && ld % 4 == 0 && halo >= 0 && halo <= 8, "pixelnorm_silu_halo: width %d must be a multiple of 32", C);
  if (R == 0) return MESHTOK_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* img = static_cast<uint8_t*>(image);
  if (C <= 128) pixelnorm_silu_halo_rows_kernel<4><<<grid_groups(R), 256, 0, st>>>(h, ld, valid, R, P, C, eps, halo, img);
  else if (C <= 256) pixelnorm_silu_halo_rows_kernel<8><<<grid_groups(R), 256, 0, st>>>(h, ld, valid, R, P, C, eps, halo, img);
  else if (C <= 384) pixelnorm_silu_halo_rows_kernel<12><<<grid_groups(R), 256, 0, st>>>(h, ld, valid, R, P, C, eps, halo, img);
  else pixelnorm_silu_halo_kernel<<<grid_groups(R), 256, 0, st>>>(h, ld, valid, R, P, C, eps, halo, img);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_dec_pixelnorm_squeeze_excite_gate(const float* h, int ld, const uint8_t* valid, int B, int P, int lead, int C, float eps, int inner,
                                              const float* w1t, const float* b1, const float* w2t, const float* b2, float* scale, float* slice_ws,
                                              int32_t* tickets, float* gate, meshtok_stream_t stream) {
  MT_CHECK_ARG(h && valid && w1t && b1 && w2t && b2 && scale && slice_ws && gate && C % 4 == 0 && C <= SE_MAX_C && inner >= 1 && inner <= 256 && ld >= C &&
                   ld % 4 == 0 && lead >= 0 && P > lead,
               "pixelnorm_squeeze_excite_gate: bad arguments (C %% 4 == 0, C <= %d, inner <= 256)", SE_MAX_C);
  if (B == 0) return MESHTOK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int S = meshtok_dec_se_slices(B, P - lead);
  int* cnt = reinterpret_cast<int*>(slice_ws + (size_t)B * S * C);
  const int nj = (C + 127) / 128;
  size_t sm_bytes = (size_t)8 * nj * 32 * sizeof(float4);
  if (tickets != nullptr && sm_bytes < (size_t)(C + 512) * sizeof(float)) sm_bytes = (size_t)(C + 512) * sizeof(float);
#define MT_COLSUM(NJ) \
  pixelnorm_colsum_kernel<NJ><<<dim3(B, S), 256, sm_bytes, s>>>(h, ld, valid, P, lead, C, eps, S, scale, slice_ws, cnt, tickets, inner, w1t, b1, w2t, b2, gate)
  switch (nj) {
    case 1: MT_COLSUM(1); break;
    case 2: MT_COLSUM(2); break;
    case 3: MT_COLSUM(3); break;
    case 4: MT_COLSUM(4); break;
    case 5: MT_COLSUM(5); break;
    case 6: MT_COLSUM(6); break;
    case 7: MT_COLSUM(7); break;
    default: MT_COLSUM(8); break;
  }
#undef MT_COLSUM
  MT_CHECK_LAUNCH();
  if (tickets == nullptr) {   // no ticket counters: the gate as a launch of its own
    se_gate_slices_kernel<<<B, 1024, 0, s>>>(slice_ws, cnt, S, C, inner, w1t, b1, w2t, b2, gate);
    MT_CHECK_LAUNCH();
  }
  return MESHTOK_OK;
}

int meshtok_dec_gate_residual_halo(const float* h, int ldh, const float* scale, const float* gate, const float* res, int ldres, const uint8_t* valid,
                                   int R, int P, int C, float* out, int halo, void* image, meshtok_stream_t stream) {
  MT_CHECK_ARG(h && gate && res && valid && out && R >= 0 && P >= 1 && C % 32 == 0 && ldh >= C && ldh % 4 == 0 && ldres >= C && ldres % 4 == 0 && halo >= 0 &&
                   halo <= 8,
               "gate_residual_halo: width %d must be a multiple of 32", C);
  if (R == 0) return MESHTOK_OK;
  gate_residual_halo_kernel<<<grid_groups(R), 256, 0, static_cast<cudaStream_t>(stream)>>>(h, ldh, scale, gate, res, ldres, valid, R, P, C, out, halo,
                                                                                          static_cast<uint8_t*>(image));
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int meshtok_dec_argmax_coords_padded(const float* logits, const uint8_t* valid, int B, int P, int lead, int ncoord, int nbins, float lo, float hi,
                                     float* coords, int64_t* bins, meshtok_stream_t stream) {
  MT_CHECK_ARG(logits && valid && coords && B >= 0 && lead >= 0 && P > lead && ncoord >= 1 && nbins >= 1, "argmax_coords_padded: bad arguments");
  if (B == 0) return MESHTOK_OK;
  if (nbins == 128)
    argmax_coords_padded128_kernel<<<grid_for((long long)B * (P - lead), 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, valid, B, P, lead, ncoord, lo, hi,
                                                                                                                         coords, bins);
  else
    argmax_coords_padded_kernel<<<grid_for((long long)B * (P - lead) * ncoord, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, valid, B, P, lead, ncoord,
                                                                                                                               nbins, lo, hi, coords, bins);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

}  // extern "C"
