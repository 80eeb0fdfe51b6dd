// K7 scatter-mean faces -> vertices, K8b residual LFQ, K9 gather codes back per face-vertex.
//
// Replaces MeshAutoencoder.quantize (meshgpt_pytorch.py:787-870): the padding-aware scatter_mean
// (:243-257, :807-824) becomes a segmented sum over the vertex incidence CSR built by the topology
// kernel (one warp per referenced vertex, entries in the reference's (face, slot) order, no atomics,
// no (b, nf*nvf, d) int64 index tensor); ResidualLFQ (vector_quantize_pytorch, called at :842) becomes
// log2(K) dot products + sign tests per vertex; the two get_at gathers and masked_fills (:856-866,
// :1018) become one kernel that writes the final int64 codes and the codebook-usage histogram.
#include <string.h>

#include "common.cuh"
#include "kernels.cuh"
#include "rvq_screen.cuh"

namespace meshtok {

namespace {

template <int PER2>  // float2 per lane: D <= 64 * PER2
__global__ void __launch_bounds__(256) scatter_mean_kernel(const float* __restrict__ y, int D, const int* __restrict__ vptr,
                                                           const int* __restrict__ vinc, float* __restrict__ avg,
                                                           const Counts* counts, RvqPrepOut prep) {
  pdl_wait();
  const int lane = lane_id();
  const int wpb = blockDim.x >> 5;
  const int n = counts->n_vrows;
  const int d2 = D >> 1;
  for (int row = blockIdx.x * wpb + warp_id(); row < n; row += gridDim.x * wpb) {
    const int s = vptr[row], e = vptr[row + 1];
    MT_DCHECK(s >= 0 && e >= s && e <= counts->n_faces * 4);
    float2 acc[PER2];
#pragma unroll
    for (int i = 0; i < PER2; ++i) acc[i] = make_float2(0.f, 0.f);
    for (int q = s; q < e; ++q) {
      MT_DCHECK(vinc[q] >= 0 && vinc[q] < counts->n_faces * 4);
      const float2* yr = reinterpret_cast<const float2*>(y + (size_t)vinc[q] * D);
#pragma unroll
      for (int i = 0; i < PER2; ++i) {
        const int c = lane + 32 * i;
        if (c < d2) {
          const float2 v = __ldg(yr + c);
          acc[i].x += v.x;
          acc[i].y += v.y;
        }
      }
    }
    const float cnt = fmaxf((float)(e - s), 1e-5f);  // den.clamp(min = eps)
    float2* out = reinterpret_cast<float2*>(avg + (size_t)row * D);
    float2 v[PER2];
    float vmax = 0.f;
#pragma unroll
    for (int i = 0; i < PER2; ++i) {
      const int c = lane + 32 * i;
      v[i] = make_float2(0.f, 0.f);
      if (c < d2) {
        v[i] = make_float2(__fdiv_rn(acc[i].x, cnt), __fdiv_rn(acc[i].y, cnt));
        out[c] = v[i];
        vmax = fmaxf(vmax, fmaxf(fabsf(v[i].x), fabsf(v[i].y)));
      }
    }
    if (prep.resid != nullptr) {
      // the averaged vertex row is level 0's residual of the residual VQ: leave it ready for the screening search
      // (copy, power-of-two row scale, fp16 row image) instead of running a separate preparation pass
      float2* res = reinterpret_cast<float2*>(prep.resid + (size_t)row * D);
#pragma unroll
      for (int i = 0; i < PER2; ++i) {
        const int c = lane + 32 * i;
        if (c < d2) res[c] = v[i];
      }
      rvq_emit_row_scale_and_image<PER2>(v, vmax, row, D, prep.rscale, prep.row_image, prep.c_scale);
    }
  }
  pdl_launch_dependents();
}

// sum over the warp of 16 per-lane values at once: a halving butterfly (8 + 4 + 2 + 1 + 1 shuffles instead of 16 x 5);
// afterwards every lane holds the total of value number (lane >> 1) & 15
__device__ __forceinline__ float warp_sum16(float (&p)[16]) {
  const int lane = lane_id();
#pragma unroll
  for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = up ? p[j] : p[j + half];
      const float keep = up ? p[j + half] : p[j];
      p[j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  return p[0] + __shfl_xor_sync(0xffffffffu, p[0], 1);
}

// Residual LFQ (eval).  x = W_in v + b; level l: bit_d = r_d > 0, q_d = +-2^-l, r <- r - q;
// code = sum_d bit_d << (bits-1-d)  (dimension 0 is the most significant bit).
// The soft clamp tanh(r/c)*c of the upstream module preserves the sign and is therefore skipped.
// Two rows per warp pass: every weight is loaded once for both, and the `bits` (<= 16) dot products of a row are reduced
// together (warp_sum16).  The first version reduced one dot product at a time: 14 x 5 shuffles per row, 57 us for the
// 55 168 rows of the quad workload.
// PER = elements per lane; VEC2: D is a multiple of 64 and a lane owns PER / 2 pairs of adjacent columns (8-byte loads, no
// tail predicates: half the load instructions of the scalar layout, which is kept for other widths)
template <int PER, bool VEC2>
__global__ void __launch_bounds__(256) lfq_kernel(const float* __restrict__ rows, int D, const float* __restrict__ w_in,
                                                  const float* __restrict__ b_in, int bits, int Q,
                                                  const float* __restrict__ w_out, const float* __restrict__ b_out,
                                                  int32_t* __restrict__ codes, float* __restrict__ quantized_out,
                                                  const int* __restrict__ n_rows_dev, int max_rows) {
  const int lane = lane_id();
  const int wpb = blockDim.x >> 5;
  const int n = n_rows_dev ? min(*n_rows_dev, max_rows) : max_rows;
  const int d_mine = (lane >> 1) & 15;   // the projected dimension whose total this lane ends up with
  const float bias_mine = d_mine < bits ? b_in[d_mine] : 0.f;
  for (int row0 = 2 * (blockIdx.x * wpb + warp_id()); row0 < n; row0 += 2 * gridDim.x * wpb) {
    const bool two = row0 + 1 < n;
    float v0[PER], v1[PER];
    float p0[16], p1[16];
    if (VEC2) {
      const int np = D >> 6;   // pairs per lane (D / 64); PER / 2 at most
#pragma unroll
      for (int i = 0; i < PER / 2; ++i) {
        float2 a = make_float2(0.f, 0.f), b = a;
        if (i < np) {
          a = *reinterpret_cast<const float2*>(rows + (size_t)row0 * D + 2 * (lane + 32 * i));
          if (two) b = *reinterpret_cast<const float2*>(rows + (size_t)(row0 + 1) * D + 2 * (lane + 32 * i));
        }
        v0[2 * i] = a.x; v0[2 * i + 1] = a.y; v1[2 * i] = b.x; v1[2 * i + 1] = b.y;
      }
#pragma unroll
      for (int dd = 0; dd < 16; ++dd) {
        p0[dd] = 0.f; p1[dd] = 0.f;
        if (dd < bits) {
#pragma unroll
          for (int i = 0; i < PER / 2; ++i) {
            if (i < np) {
              const float2 w = __ldg(reinterpret_cast<const float2*>(w_in + (size_t)dd * D) + lane + 32 * i);
              p0[dd] = fmaf(v0[2 * i + 1], w.y, fmaf(v0[2 * i], w.x, p0[dd]));
              p1[dd] = fmaf(v1[2 * i + 1], w.y, fmaf(v1[2 * i], w.x, p1[dd]));
            }
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        v0[i] = c < D ? rows[(size_t)row0 * D + c] : 0.f;
        v1[i] = (c < D && two) ? rows[(size_t)(row0 + 1) * D + c] : 0.f;
      }
#pragma unroll
      for (int dd = 0; dd < 16; ++dd) {
        p0[dd] = 0.f; p1[dd] = 0.f;
        if (dd < bits) {
#pragma unroll
          for (int i = 0; i < PER; ++i) {
            const int c = lane + 32 * i;
            if (c < D) {
              const float w = __ldg(w_in + (size_t)dd * D + c);
              p0[dd] = fmaf(v0[i], w, p0[dd]);
              p1[dd] = fmaf(v1[i], w, p1[dd]);
            }
          }
        }
      }
    }
    const float x0 = warp_sum16(p0) + bias_mine, x1 = warp_sum16(p1) + bias_mine;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      if (rr == 1 && !two) break;
      const int row = row0 + rr;
      float r = rr ? x1 : x0, qsum = 0.f;
      for (int l = 0; l < Q; ++l) {
        const float scale = exp2f(-(float)l);
        const bool bit = (d_mine < bits) && (r > 0.f);
        // dimension 0 is the most significant bit; the two lanes that hold a dimension contribute the same bit
        const unsigned code = __reduce_or_sync(0xffffffffu, bit ? (1u << (bits - 1 - d_mine)) : 0u);
        if (lane == 0) codes[(size_t)row * Q + l] = (int32_t)code;
        const float q = bit ? scale : -scale;
        r -= q;
        qsum += q;
      }
      if (quantized_out != nullptr) {
        // project_out: (D x bits) . qsum + b_out   (lane 2 * dd holds dimension dd)
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          const int c = lane + 32 * i;
          float o = 0.f;
          for (int dd = 0; dd < bits; ++dd) {
            const float qd = __shfl_sync(0xffffffffu, qsum, 2 * dd);
            if (c < D) o = fmaf(qd, __ldg(w_out + (size_t)c * bits + dd), o);
          }
          if (c < D) quantized_out[(size_t)row * D + c] = o + b_out[c];
        }
      }
    }
  }
}

// hist[q][code] += number of output tokens that carry the code = number of (face, slot) incidences of the vertex.
// One weighted atomic per DISTINCT code per warp and level: the 32 rows of a warp are combined first (codebook usage
// is skewed, so many rows of a warp share a code and un-aggregated atomics serialise on a few hot addresses).
// Rides in the first ceil(rows / 256) CTAs of the code-gather kernels (it used to be a launch of its own: ~5 us of work
// behind ~2 us of launch latency); every thread of the CTA must call it.
struct HistArgs {
  const int* vptr;
  const Counts* counts;
  int K, blocks;                 // blocks = CTAs that carry histogram rows (0: no histogram)
  unsigned long long* hist;      // (Q, K) or null
};
__device__ __forceinline__ void histogram_rows(const HistArgs& h, const int32_t* __restrict__ vcodes, int Q) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = row < h.counts->n_vrows;
  const unsigned w = live ? (unsigned)(h.vptr[row + 1] - h.vptr[row]) : 0u;
  const int lane = lane_id();
  for (int q = 0; q < Q; ++q) {
    int c = live ? vcodes[(size_t)row * Q + q] : -1;
    if (c >= h.K) c = -1;
    unsigned sum = 0;
    int leader = 32;
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const int cl = __shfl_sync(0xffffffffu, c, l);
      const unsigned wl = __shfl_sync(0xffffffffu, w, l);
      if (cl == c) { sum += wl; leader = min(leader, l); }
    }
    if (c >= 0 && leader == lane && sum) atomicAdd(h.hist + (size_t)q * h.K + c, (unsigned long long)sum);
  }
}

__global__ void __launch_bounds__(256) gather_codes_kernel(const int64_t* __restrict__ faces, const int* __restrict__ face_row,
                                                           const int* __restrict__ vert_row, const int32_t* __restrict__ vcodes,
                                                           int BF, int F, int nvf, int V, int Q, int64_t pad_id,
                                                           int64_t* __restrict__ codes_out, HistArgs h) {
  pdl_wait();
  if ((int)blockIdx.x < h.blocks) histogram_rows(h, vcodes, Q);
  const long long tok = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= (long long)BF * nvf) return;
  const int pf = (int)(tok / nvf);
  int64_t* dst = codes_out + tok * Q;
  if (face_row[pf] < 0) {
    for (int q = 0; q < Q; ++q) dst[q] = pad_id;
    return;
  }
  const int b = pf / F;
  const long long v = faces[tok];
  MT_DCHECK(v >= 0 && v < V);
  const int vr = vert_row[(size_t)b * V + v];
  MT_DCHECK(vr >= 0 && (long long)vr < (long long)BF / F * V);
  for (int q = 0; q < Q; ++q) {
    dst[q] = vcodes[(size_t)vr * Q + q];
  }
}

// The same for the ragged input layout: one thread per token of the flat (total_faces * nvf) sequence; the mesh of a face is found
// by bisection in the (B + 1) input offsets (L1 resident), which gives its slot b * F + f in the workspace's maps.
__global__ void __launch_bounds__(256) gather_codes_ragged_kernel(const int64_t* __restrict__ faces, const int32_t* __restrict__ in_face_off,
                                                                  const int* __restrict__ face_row, const int* __restrict__ vert_row,
                                                                  const int32_t* __restrict__ vcodes, long long total_faces, int B, int F,
                                                                  int nvf, int V, int Q, int64_t pad_id, int64_t* __restrict__ codes_out, HistArgs h) {
  pdl_wait();
  if ((int)blockIdx.x < h.blocks) histogram_rows(h, vcodes, Q);
  const long long tok = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // total_faces (host) is the capacity of the buffers, the live count is the last offset: a graph captured on upper bounds serves every batch
  if (tok >= min(total_faces, (long long)__ldg(in_face_off + B)) * nvf) return;
  const int g = (int)(tok / nvf);
  int lo = 0, hi = B;            // the last b with in_face_off[b] <= g (empty meshes repeat an offset: the last one owns the face)
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(in_face_off + mid) <= g) lo = mid; else hi = mid;
  }
  const int b = lo, f = g - __ldg(in_face_off + b);
  int64_t* dst = codes_out + tok * Q;
  const int row = (f < F && g < __ldg(in_face_off + b + 1)) ? face_row[(size_t)b * F + f] : -1;
  if (row < 0) {
    for (int q = 0; q < Q; ++q) dst[q] = pad_id;
    return;
  }
  const int vr = vert_row[(size_t)b * V + faces[tok]];
  for (int q = 0; q < Q; ++q) dst[q] = vcodes[(size_t)vr * Q + q];
}

// ragged input: flat face g -> (mesh b, face f inside it); the last b with in_face_off[b] <= g (empty meshes repeat an offset)
__device__ __forceinline__ int ragged_mesh_of_face(const int32_t* __restrict__ in_face_off, int B, long long g) {
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((long long)__ldg(in_face_off + mid) <= g) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) gather_quantized_kernel(const int64_t* __restrict__ faces, const int* __restrict__ face_row,
                                                               const int* __restrict__ vert_row, const float* __restrict__ qrows,
                                                               int BF, int F, int nvf, int V, int D,
                                                               const float* __restrict__ pad_row, float* __restrict__ out,
                                                               const int32_t* __restrict__ in_face_off, int B, long long total_faces) {
  const int wpb = blockDim.x >> 5;
  const long long tok = (long long)blockIdx.x * wpb + warp_id();
  int pf, b;
  if (in_face_off != nullptr) {   // ragged layout: tok indexes the flat (total_faces * nvf) sequence
    if (tok >= min(total_faces, (long long)__ldg(in_face_off + B)) * nvf) return;
    const long long g = tok / nvf;
    b = ragged_mesh_of_face(in_face_off, B, g);
    const int f = (int)(g - __ldg(in_face_off + b));
    pf = f < F ? b * F + f : -1;
  } else {
    if (tok >= (long long)BF * nvf) return;
    pf = (int)(tok / nvf);
    b = pf / F;
  }
  const float* srcp;
  if (pf < 0 || face_row[pf] < 0) srcp = pad_row;
  else srcp = qrows + (size_t)vert_row[(size_t)b * V + faces[tok]] * D;
  for (int c = lane_id(); c < D; c += 32) out[tok * D + c] = srcp ? srcp[c] : 0.f;
}

__global__ void __launch_bounds__(256) unpack_rows_kernel(const float* __restrict__ packed, const int* __restrict__ face_row,
                                                          int BF, int d4, float4* __restrict__ out, const int32_t* __restrict__ in_face_off, int B, int F,
                                                          long long total_faces) {
  const int wpb = blockDim.x >> 5;
  if (in_face_off != nullptr) {   // ragged layout: out is (total_faces, d), one row per flat face
    const long long n = min(total_faces, (long long)__ldg(in_face_off + B));
    for (long long g = (long long)blockIdx.x * wpb + warp_id(); g < n; g += (long long)gridDim.x * wpb) {
      const int b = ragged_mesh_of_face(in_face_off, B, g);
      const int f = (int)(g - __ldg(in_face_off + b));
      const int row = f < F ? face_row[(size_t)b * F + f] : -1;
      const float4* s = reinterpret_cast<const float4*>(packed) + (size_t)(row < 0 ? 0 : row) * d4;
      for (int c = lane_id(); c < d4; c += 32) out[g * d4 + c] = row < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : s[c];
    }
    return;
  }
  for (long long pf = (long long)blockIdx.x * wpb + warp_id(); pf < BF; pf += (long long)gridDim.x * wpb) {
    const int row = face_row[pf];
    const float4* s = reinterpret_cast<const float4*>(packed) + (size_t)(row < 0 ? 0 : row) * d4;
    for (int c = lane_id(); c < d4; c += 32) out[pf * d4 + c] = row < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : s[c];
  }
}

__global__ void __launch_bounds__(256) pack_rows_kernel(const float4* __restrict__ padded, const int* __restrict__ row_face,
                                                        int d4, float4* __restrict__ packed, const Counts* counts, const int32_t* __restrict__ in_face_off,
                                                        int F) {
  const int wpb = blockDim.x >> 5;
  const int n = counts->n_faces;
  for (int row = blockIdx.x * wpb + warp_id(); row < n; row += gridDim.x * wpb) {
    size_t from = (size_t)row_face[row];   // slot b*F+f of the padded layout ...
    if (in_face_off != nullptr) {          // ... or flat face in_face_off[b] + f of the ragged one
      const int b = (int)(from / F);
      from = (size_t)__ldg(in_face_off + b) + (from - (size_t)b * F);
    }
    const float4* s = padded + from * d4;
    for (int c = lane_id(); c < d4; c += 32) packed[(size_t)row * d4 + c] = s[c];
  }
}


// Residual LFQ, round 2: THREAD = ROW.  The warp-per-two-rows kernel above spends its instructions on weight loads (every lane re-reads
// its 14 x 6 weights from L1 for every pair of rows) and on the butterfly that joins 32 partial sums per dot product: ~210 instructions
// per row, 35 us for the 55 168 rows of the quad workload.  Here a thread keeps all `bits` dot products of ITS row in registers and walks
// the row's columns: per 4 columns one 16-byte read of the row and, per projected dimension, one 16-byte BROADCAST read of the weights
// (shared memory, staged once per CTA before griddepcontrol.wait - they are model constants) + 4 FMAs; no shuffles, no reductions, the
// sign bits and the residual levels are per-thread scalar code: ~90 instructions per row.  Rows arrive as 128-row x 32-column chunks by
// cp.async (coalesced 128-byte pieces), double buffered, in rows of 36 floats so that the eight 16-byte reads of a quarter warp cover all
// 32 banks.  Dot products accumulate in column order (one FMA chain per projected dimension).
constexpr int LFQ_TR = 128;   // rows per tile = threads per CTA
constexpr int LFQ_CH = 32;    // columns per staged chunk
constexpr int LFQ_LD = 36;    // floats per staged row

// DS = 2: two adjacent lanes share a row, each owns every other projected dimension (twice the warps for the same rows: the kernel is
// bound by the latency of its shared-memory reads at ~3 warps per scheduler, ncu); the weight rows are padded by 4 floats so that the two
// 16-byte broadcast reads of an instruction fall into different banks.
template <int NB, bool EXACT, int RPT, int DS>   // NB accumulators per row; EXACT: bits == NB (no guards); RPT rows per thread
__global__ void __launch_bounds__(LFQ_TR * DS / RPT) lfq_rows_kernel(const float* __restrict__ rows, int D, const float* __restrict__ w_in,
                                                                     const float* __restrict__ b_in, int bits, int Q,
                                                                     const float* __restrict__ w_out, const float* __restrict__ b_out,
                                                                     int32_t* __restrict__ codes, float* __restrict__ quantized_out,
                                                                     const int* __restrict__ n_rows_dev, int max_rows) {
  constexpr int NT = LFQ_TR * DS / RPT;   // threads
  constexpr int RS = LFQ_TR / RPT;        // a thread's rows are RS apart in the tile
  constexpr int NBT = NB / DS;            // projected dimensions per thread: h, h + DS, ...
  static_assert(NB % DS == 0, "NB must split evenly");
  extern __shared__ __align__(16) float lfq_sm[];
  const int WLD = D + 4;                  // floats per staged weight row
  float* wsm = lfq_sm;                    // [bits][WLD]
  float* tile = lfq_sm + bits * WLD;      // [2][LFQ_TR][LFQ_LD]
  const int tid = threadIdx.x;
  const int trow = tid / DS, h = tid % DS;
  {
    const int d4 = D >> 2;
    for (int i = tid; i < bits * d4; i += NT) {
      const int dd = i / d4, c4 = i - dd * d4;
      *reinterpret_cast<float4*>(wsm + dd * WLD + 4 * c4) = __ldg(reinterpret_cast<const float4*>(w_in + (size_t)dd * D) + c4);
    }
  }
  pdl_wait();
  const int n = n_rows_dev ? min(*n_rows_dev, max_rows) : max_rows;
  const int nchunks = D / LFQ_CH;
  const uint32_t tile_addr = (uint32_t)__cvta_generic_to_shared(tile);
  for (int tile0 = blockIdx.x * LFQ_TR; tile0 < n; tile0 += gridDim.x * LFQ_TR) {
    auto issue = [&](int k, int buf) {   // chunk k of this tile's rows -> buffer buf (rows beyond n are left as they are: never used)
#pragma unroll
      for (int j = 0; j < LFQ_TR * (LFQ_CH / 4) / NT; ++j) {
        const int idx = tid + j * NT, r = idx >> 3, c4 = idx & 7;
        MT_DCHECK(r < LFQ_TR && buf >= 0 && buf < 2 && k >= 0 && (k + 1) * LFQ_CH <= D);
        if (tile0 + r < n) {
          const float* src = rows + (size_t)(tile0 + r) * D + k * LFQ_CH + c4 * 4;
          const uint32_t dst = tile_addr + (uint32_t)((buf * LFQ_TR + r) * LFQ_LD + c4 * 4) * 4u;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float p[RPT][NBT];
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr)
#pragma unroll
      for (int i = 0; i < NBT; ++i) p[rr][i] = 0.f;
    issue(0, 0);
    for (int k = 0; k < nchunks; ++k) {
      const int buf = k & 1;
      if (k + 1 < nchunks) {
        issue(k + 1, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();   // chunk k (and, the first time, the weights) visible to every thread
      const float* tr = tile + (buf * LFQ_TR + trow) * LFQ_LD;
      const float* wk = wsm + h * WLD + k * LFQ_CH;
#pragma unroll
      for (int c4 = 0; c4 < LFQ_CH / 4; ++c4) {
        float4 r[RPT];
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) r[rr] = *reinterpret_cast<const float4*>(tr + rr * RS * LFQ_LD + 4 * c4);
#pragma unroll
        for (int i = 0; i < NBT; ++i) {
          if (EXACT || h + DS * i < bits) {
            const float4 w = *reinterpret_cast<const float4*>(wk + i * DS * WLD + 4 * c4);   // one broadcast read serves RPT rows
#pragma unroll
            for (int rr = 0; rr < RPT; ++rr)
              p[rr][i] = fmaf(r[rr].w, w.w, fmaf(r[rr].z, w.z, fmaf(r[rr].y, w.y, fmaf(r[rr].x, w.x, p[rr][i]))));
          }
        }
      }
      __syncthreads();   // every read of this buffer is done before the iteration after next refills it
    }
    float bias[NBT];
#pragma unroll
    for (int i = 0; i < NBT; ++i) bias[i] = (EXACT || h + DS * i < bits) ? __ldg(b_in + h + DS * i) : 0.f;
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr) {
      const int row = tile0 + trow + rr * RS;
      const bool live = row < n;
#pragma unroll
      for (int i = 0; i < NBT; ++i) p[rr][i] += bias[i];
      float qsum[NBT];
#pragma unroll
      for (int i = 0; i < NBT; ++i) qsum[i] = 0.f;
      for (int l = 0; l < Q; ++l) {
        const float scale = exp2f(-(float)l);
        unsigned code = 0u;
#pragma unroll
        for (int i = 0; i < NBT; ++i) {
          if (EXACT || h + DS * i < bits) {
            const bool bit = p[rr][i] > 0.f;   // dimension 0 is the most significant bit
            code |= bit ? (1u << (bits - 1 - (h + DS * i))) : 0u;
            const float q = bit ? scale : -scale;
            p[rr][i] -= q;
            qsum[i] += q;
          }
        }
        if (DS == 2) code |= __shfl_xor_sync(0xffffffffu, code, 1);
        if (live && h == 0) codes[(size_t)row * Q + l] = (int32_t)code;
      }
#pragma unroll
      for (int i = 0; i < NBT; ++i) p[rr][i] = qsum[i];   // (kept for project_out below)
    }
    if (quantized_out != nullptr) {   // project_out: (D x bits) . qsum + b_out, through the staging tile for coalesced stores
      for (int k = 0; k < nchunks; ++k) {
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) {
          float* tw = tile + (trow + rr * RS) * LFQ_LD;
#pragma unroll
          for (int c4 = 0; c4 < LFQ_CH / 4; ++c4) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = k * LFQ_CH + c4 * 4 + j;
              float t = 0.f;
#pragma unroll
              for (int i = 0; i < NBT; ++i)
                if (EXACT || h + DS * i < bits) t = fmaf(p[rr][i], __ldg(w_out + (size_t)c * bits + h + DS * i), t);
              if (DS == 2) t += __shfl_xor_sync(0xffffffffu, t, 1);
              o[j] = t + __ldg(b_out + c);
            }
            if (h == 0) *reinterpret_cast<float4*>(tw + 4 * c4) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < LFQ_TR * (LFQ_CH / 4) / NT; ++j) {
          const int idx = tid + j * NT, r = idx >> 3, c4 = idx & 7;
          if (tile0 + r < n)
            *reinterpret_cast<float4*>(quantized_out + (size_t)(tile0 + r) * D + k * LFQ_CH + c4 * 4) = *reinterpret_cast<const float4*>(tile + r * LFQ_LD + c4 * 4);
        }
        __syncthreads();
      }
    }
  }
  pdl_launch_dependents();
}

}  // namespace

int launch_scatter_mean(const float* y, int D, const int* vptr, const int* vinc, float* avg, const Counts* counts,
                        int max_vrows, const RvqPrepOut* prep, cudaStream_t stream) {
  MT_CHECK_ARG(D % 2 == 0 && D <= 512, "scatter_mean: dim %d must be even and <= 512", D);
  const int blocks = max(1, min(ceil_div(max_vrows, 8), 148 * 16));
  const int per2 = ceil_div(D / 2, 32);
  RvqPrepOut po;
  memset(&po, 0, sizeof(po));
  if (prep) po = *prep;
  MT_CHECK_ARG(po.row_image == nullptr || (D % 64 == 0 && po.c_scale > 0.f), "scatter_mean: a row image needs dim %% 64 == 0 and c_scale > 0");
#define SM(P) MT_LAUNCH_PDL(scatter_mean_kernel<P>, blocks, 256, 0, stream, y, D, vptr, vinc, avg, counts, po)
  if (per2 <= 1) SM(1);
  else if (per2 <= 2) SM(2);
  else if (per2 <= 3) SM(3);
  else if (per2 <= 4) SM(4);
  else SM(8);
#undef SM
  return MESHTOK_OK;
}

int launch_lfq(const float* rows, int D, const float* w_in, const float* b_in, int bits, int Q, const float* w_out,
               const float* b_out, int32_t* codes, float* quantized_out, const int* n_rows_dev, int max_rows,
               cudaStream_t stream) {
  MT_CHECK_ARG(bits >= 1 && bits <= 30, "lfq: codebook_size must be 2^1..2^30 (bits=%d)", bits);
  MT_CHECK_ARG(D >= 1 && D <= 512, "lfq: dim %d must be <= 512", D);
  if (max_rows == 0) return MESHTOK_OK;
  MT_CHECK_ARG(bits >= 1 && bits <= 16, "lfq: %d bits per level (codebook_size up to 65536 is supported)", bits);
  // thread = row (lfq_rows_kernel): whole 32-column chunks, 16-byte aligned rows and weights; MESHTOK_LFQ_ROWS=0 keeps the warp kernel (A/B)
  static int use_rows = -1;
  if (use_rows < 0) { const char* e = getenv("MESHTOK_LFQ_ROWS"); use_rows = (e && atoi(e) == 0) ? 0 : 1; }
  const size_t smem_rows = ((size_t)bits * (D + 4) + 2 * (size_t)LFQ_TR * LFQ_LD) * sizeof(float);
  if (use_rows && D % LFQ_CH == 0 && (reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(w_in) | reinterpret_cast<uintptr_t>(quantized_out)) % 16 == 0 &&
      smem_rows <= 200 * 1024) {
    const int blocks_r = max(1, min(ceil_div(max_rows, LFQ_TR), 148 * 4));
#define LR1(NB, EX, R, S)                                                                                                                 \
  do {                                                                                                                                    \
    MT_ENSURE_SMEM((lfq_rows_kernel<NB, EX, R, S>), 200 * 1024);                                                                          \
    MT_LAUNCH_PDL((lfq_rows_kernel<NB, EX, R, S>), blocks_r, LFQ_TR * S / R, smem_rows, stream, rows, D, w_in, b_in, bits, Q, w_out,      \
                  b_out, codes, quantized_out, n_rows_dev, max_rows);                                                                     \
  } while (0)
#define LR(NB, EX)                                                                                                                        \
  do {                                                                                                                                    \
    if (variant == 12) LR1(NB, EX, 1, 2);                                                                                                 \
    else if (variant == 22) LR1(NB, EX, 2, 2);                                                                                            \
    else if (variant == 21) LR1(NB, EX, 2, 1);                                                                                            \
    else LR1(NB, EX, 1, 1);                                                                                                               \
  } while (0)
    // variant = 10 * rows per thread + lanes per row (MESHTOK_LFQ_VARIANT=11|21|12|22; measured on C3, B200: see DESIGN)
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("MESHTOK_LFQ_VARIANT"); variant = e ? atoi(e) : 12; }
    if (bits == 14) LR(14, true);
    else if (bits == 16) LR(16, true);
    else if (bits == 12) LR(12, true);
    else if (bits == 10) LR(10, true);
    else if (bits == 8) LR(8, true);
    else if (bits < 8) LR(8, false);
    else LR(16, false);
#undef LR1
#undef LR
    return MESHTOK_OK;
  }
  const int blocks = max(1, min(ceil_div(max_rows, 16), 148 * 16));
  const int per = ceil_div(D, 32);
#define LQ(P, V) lfq_kernel<P, V><<<blocks, 256, 0, stream>>>(rows, D, w_in, b_in, bits, Q, w_out, b_out, codes, quantized_out, n_rows_dev, max_rows)
  const bool vec2 = D % 64 == 0 && (reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(w_in)) % 8 == 0;
  if (vec2 && per <= 2) LQ(2, true);
  else if (vec2 && per <= 6) LQ(6, true);
  else if (vec2 && per <= 8) LQ(8, true);
  else if (per <= 2) LQ(2, false);
  else if (per <= 6) LQ(6, false);
  else if (per <= 8) LQ(8, false);
  else LQ(16, false);
#undef LQ
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

static HistArgs hist_args(const GatherHist* gh, int K) {
  HistArgs h;
  memset(&h, 0, sizeof(h));
  if (gh != nullptr && gh->hist != nullptr && gh->max_vrows > 0) {
    h.vptr = gh->vptr; h.counts = gh->counts; h.K = K; h.hist = gh->hist;
    h.blocks = ceil_div(gh->max_vrows, 256);
  }
  return h;
}

int launch_gather_codes(const int64_t* faces, const int* face_row, const int* vert_row, const int32_t* vcodes, int B,
                        int F, int nvf, int V, int Q, int64_t pad_id, int64_t* codes_out, const GatherHist* gh, int K, cudaStream_t stream) {
  const long long n = (long long)B * F * nvf;
  const HistArgs h = hist_args(gh, K);
  if (n == 0 && h.blocks == 0) return MESHTOK_OK;
  const unsigned blocks = (unsigned)max((long long)h.blocks, (n + 255) / 256);
  MT_LAUNCH_PDL(gather_codes_kernel, blocks, 256, 0, stream, faces, face_row, vert_row, vcodes, B * F, F, nvf, V, Q, pad_id, codes_out, h);
  return MESHTOK_OK;
}

int launch_gather_codes_ragged(const int64_t* faces, const int32_t* in_face_off, const int* face_row, const int* vert_row, const int32_t* vcodes,
                               long long total_faces, int B, int F, int nvf, int V, int Q, int64_t pad_id, int64_t* codes_out,
                               const GatherHist* gh, int K, cudaStream_t stream) {
  const long long n = total_faces * nvf;
  const HistArgs h = hist_args(gh, K);
  if (n == 0 && h.blocks == 0) return MESHTOK_OK;
  const unsigned blocks = (unsigned)max((long long)h.blocks, (n + 255) / 256);
  MT_LAUNCH_PDL(gather_codes_ragged_kernel, blocks, 256, 0, stream, faces, in_face_off, face_row, vert_row, vcodes, total_faces,
                B, F, nvf, V, Q, pad_id, codes_out, h);
  return MESHTOK_OK;
}

int launch_gather_quantized(const int64_t* faces, const int* face_row, const int* vert_row, const float* qrows, int B,
                            int F, int nvf, int V, int D, const float* pad_row, float* out, const int32_t* in_face_off, long long total_faces,
                            cudaStream_t stream) {
  const long long n = (in_face_off ? total_faces : (long long)B * F) * nvf;
  if (n == 0) return MESHTOK_OK;
  gather_quantized_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(faces, face_row, vert_row, qrows, B * F, F, nvf, V, D,
                                                                      pad_row, out, in_face_off, B, total_faces);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int launch_unpack_rows(const float* packed, const int* face_row, int BF, int d, float* out, const int32_t* in_face_off, int B, int F,
                       long long total_faces, cudaStream_t stream) {
  MT_CHECK_ARG(d % 4 == 0, "unpack_rows: width %d must be a multiple of 4", d);
  const long long n = in_face_off ? total_faces : (long long)BF;
  if (n == 0) return MESHTOK_OK;
  const int blocks = (int)max(1ll, min((n + 7) / 8, 148ll * 16));
  unpack_rows_kernel<<<blocks, 256, 0, stream>>>(packed, face_row, BF, d / 4, reinterpret_cast<float4*>(out), in_face_off, B, F, total_faces);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int launch_pack_rows(const float* padded, const int* row_face, int d, float* packed, const Counts* counts, int max_rows,
                     const int32_t* in_face_off, int F, cudaStream_t stream) {
  MT_CHECK_ARG(d % 4 == 0, "pack_rows: width %d must be a multiple of 4", d);
  if (max_rows == 0) return MESHTOK_OK;
  const int blocks = max(1, min(ceil_div(max_rows, 8), 148 * 16));
  pack_rows_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(padded), row_face, d / 4,
                                               reinterpret_cast<float4*>(packed), counts, in_face_off, F);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

}  // namespace meshtok
