// K4 (memory-bound half): SAGEConv mean aggregation over the face CSR and the row epilogues.
//
// Replaces the message passing of torch_geometric.nn.SAGEConv (scatter-add of x[edge_index[0]] onto
// edge_index[1], divided by clamp(count, 1)) called at meshgpt_pytorch.py:764 and :769, the
// F.normalize(p=2) that ends every SAGEConv (normalize=True, :471-474) and init_encoder_act_and_norm
// (SiLU + LayerNorm, :545-548, applied at :766).  Neighbour rows are summed in ascending source order, the
// order the reference's CPU scatter_add visits them in, so the result is deterministic (the reference's GPU
// path uses atomics and is not).  Two aggregation kernels: per-mesh column slices staged in shared memory
// (default; MEASURED against the other with scripts/gpu_gather_ab.sh: 128 vs 137 us at C2, 1.66 vs 2.09 ms at 1 024
// meshes per launch) and one warp per face row reading neighbour rows from L2 (meshes too large for a slice,
// batches too small to fill the SMs).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace meshtok {

namespace {

// Tensor map of a row-major fp32 matrix (rows x cols) for boxes of (box_rows x 32 columns): what the aggregation kernel's TMA loads walk.
// cuTensorMapEncodeTiled is a driver entry point; it is fetched through the runtime (no link against libcuda).  false: not available.
bool make_tmap_rows_f32(const float* base, int cols, long long rows, int box_rows, CUtensorMap* out) {
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    else
      (void)cudaGetLastError();
  }
  if (encode == nullptr || rows <= 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  return encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// sum / cnt for a small positive integer-valued cnt, as sum * rc with rc = RN(1 / cnt): exact for cnt = 1, 2, 4, 8 (a manifold triangle
// mesh: three neighbours + the self loop), within one rounding of the quotient otherwise - three orders of magnitude below the split-bf16
// error of the linear that consumes it.  (Round 1 corrected the product to the exactly rounded quotient with two FMAs per element: a
// sixth of the aggregation kernels' instructions, on a loop that is issue bound.)  Every aggregation kernel uses this one definition,
// so their outputs stay bit-identical to each other.
__device__ __forceinline__ float div_count(float x, float cnt, float rc) {
  (void)cnt;
  return x * rc;
}

// timing experiments (MESHTOK_GATHER_TRACE=k: the k-th aggregation launch of a step): per CTA [0] globaltimer at start, [1..5] clock64 at
// start / x copies issued / CSR staged / data landed / rows done, [6] globaltimer at the end, [7] SM id
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ int smid() { int s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
#define GM_STAMP(i, v) do { if (trace != nullptr && threadIdx.x == 0 && blockIdx.x < 1024) trace[blockIdx.x * 8 + (i)] = (v); } while (0)

template <int CHUNKS>  // float4 chunks per lane (d = 128 * CHUNKS at most)
__global__ void __launch_bounds__(256) gather_mean_kernel(const float* __restrict__ x, int d, const int* __restrict__ indptr,
                                                          const int* __restrict__ src, long long cap,
                                                          float* __restrict__ agg, uint8_t* __restrict__ agg_image,
                                                          const Counts* counts) {
  pdl_wait();
  const int lane = lane_id();
  const int wpb = blockDim.x >> 5;
  const int n_rows = counts->n_faces;
  const int d4 = d >> 2;
  // image offset of this lane's first float4 (columns 4*lane ..): chunk of 32 columns = lane >> 3, 8-column group inside
  // it = (lane & 7) >> 1, element in the group = (lane & 1) * 4; every further 32 lanes' worth is 4 chunks on
  const uint32_t img_lane = (uint32_t)(lane >> 3) * tc::IMG_BLOCK + (uint32_t)((lane & 7) >> 1) * 128u + (uint32_t)(lane & 1) * 8u;
  const size_t img_tile = (size_t)(d >> 5) * tc::IMG_BLOCK;   // bytes per 128-row tile
  for (int row = blockIdx.x * wpb + warp_id(); row < n_rows; row += gridDim.x * wpb) {
    const int s = indptr[row];
    const int e_true = indptr[row + 1];
    const int e = (long long)e_true < cap ? e_true : (int)cap;
    float4 acc[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = s; q < e; ++q) {
      const float4* xr = reinterpret_cast<const float4*>(x + (size_t)__ldg(src + q) * d);
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        const int i4 = lane + 32 * c;
        if (i4 < d4) {
          const float4 v = __ldg(xr + i4);
          acc[c].x += v.x; acc[c].y += v.y; acc[c].z += v.z; acc[c].w += v.w;
        }
      }
    }
    const float cnt = fmaxf((float)(e_true - s), 1.f);
    const float rc = __frcp_rn(cnt);
    uint8_t* img_row = agg_image ? agg_image + (size_t)(row >> 7) * img_tile + (uint32_t)((row & 127) >> 3) * 512u + (uint32_t)(row & 7) * 16u + img_lane
                                 : nullptr;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const int i4 = lane + 32 * c;
      if (i4 < d4) {
        const float4 o = make_float4(div_count(acc[c].x, cnt, rc), div_count(acc[c].y, cnt, rc), div_count(acc[c].z, cnt, rc),
                                     div_count(acc[c].w, cnt, rc));
        if (agg) reinterpret_cast<float4*>(agg + (size_t)row * d)[i4] = o;
        if (agg_image) {
          uint2 hi, lo;
          tc::split2(o.x, o.y, hi.x, lo.x);
          tc::split2(o.z, o.w, hi.y, lo.y);
          uint8_t* pimg = img_row + (size_t)c * (4 * tc::IMG_BLOCK);
          *reinterpret_cast<uint2*>(pimg) = hi;
          *reinterpret_cast<uint2*>(pimg + tc::IMG_PART) = lo;
        }
      }
    }
  }
  pdl_launch_dependents();
}

// The same aggregation with the neighbour rows read from shared memory.  Edges never leave a mesh, so CTA (mesh, slice)
// copies columns [slice * S, +S) of the mesh's x rows into shared memory once (F x S floats; 100 KB for 800 faces and
// S = 32, two CTAs per SM) and every neighbour read is a 16-byte shared-memory load: S / 8 threads per face, a quarter
// warp reads one 128-byte row slice (no bank conflicts).  L2 -> SM traffic drops from (neighbours + 1) x to 1 x the
// activation - the warp-per-row kernel above ran at the bandwidth cap of the L2 slices (8-9 TB/s into the SMs).  Same
// neighbour order and the same rounding as above, hence the same bits.
template <int S, int NTHR, int CTAS>   // slice width, threads per CTA, CTAs per SM the registers must allow
__global__ void __launch_bounds__(NTHR, CTAS) gather_mean_smem_kernel(const float* __restrict__ x, int d, const int* __restrict__ indptr,
                                                                  const int* __restrict__ src, long long cap, float* __restrict__ agg,
                                                                  uint8_t* __restrict__ agg_image, const int* __restrict__ face_off,
                                                                  int rows_cap, int edges_cap, long long* __restrict__ trace) {
  // [rows_cap + 1][S / 4] x slice (the last row stays zero) | [rows_cap + 1] local row pointers | [edges_cap] source row * S/4 (u16)
  extern __shared__ float4 sx[];
  constexpr int s4 = S >> 2;
  int* sptr = reinterpret_cast<int*>(sx + (size_t)(rows_cap + 1) * s4);
  unsigned short* ssrc = reinterpret_cast<unsigned short*>(sptr + rows_cap + 1);
  pdl_wait();
  GM_STAMP(0, gtimer()); GM_STAMP(1, clock64()); GM_STAMP(7, smid());
  const int nslices = d / S;
  const int b = blockIdx.x / nslices, slice = blockIdx.x % nslices;
  const int r0 = face_off[b];
  const int n_b = min(face_off[b + 1] - r0, rows_cap);
  const int tid = threadIdx.x;
  const int c = tid % s4;
  // x slice: 16-byte cp.async copies, all of a thread's in flight at once (a load-then-store loop paid one memory round
  // trip per iteration: 12 per thread)
  const float* xs = x + (size_t)r0 * d + slice * S;
  const uint32_t sx_addr = tc::smem_u32(sx);
  for (int i = tid; i < n_b * s4; i += NTHR) {
    const int row = i / s4;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sx_addr + (uint32_t)i * 16u), "l"(xs + (size_t)row * d + c * 4) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  GM_STAMP(2, clock64());
  if (tid < s4) sx[n_b * s4 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);   // the row an invalid source id points at
  // the mesh's part of the CSR: row pointers relative to its first edge, sources as shared-memory row offsets
  const int e0 = indptr[r0];
  const long long e1_true = indptr[r0 + n_b];
  const int n_e = (int)((e1_true < cap ? e1_true : cap) - e0);   // edges that exist in memory (capacity overflow is flagged elsewhere)
  const bool fast = n_e <= edges_cap && e1_true <= cap && (rows_cap + 1) * s4 <= 65536;
  for (int i = tid; i <= n_b; i += NTHR) sptr[i] = indptr[r0 + i] - e0;
  if (fast) {
    for (int base = tid; base < n_e; base += NTHR * 8) {
      int t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = base + j * NTHR < n_e ? __ldg(src + e0 + base + j * NTHR) : 0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (base + j * NTHR < n_e) {
          const unsigned nb = (unsigned)(t[j] - r0);   // (always < n_b: edges are intra-mesh; a corrupt list reads the zero row)
          ssrc[base + j * NTHR] = (unsigned short)((nb < (unsigned)n_b ? nb : (unsigned)n_b) * s4);
        }
    }
  }
  GM_STAMP(3, clock64());
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  GM_STAMP(4, clock64());
  // S / 8 threads per face, 8 adjacent columns (two 16-byte chunks) each.  The two faces of a quarter warp read their chunks
  // in opposite order (even chunk first / odd chunk first), so the eight 16-byte reads of one instruction fall on eight
  // different bank groups whatever rows they come from.
  constexpr int TPR = s4 / 2;
  const int g = tid % TPR;
  const int swap = (tid / TPR) & 1;
  const int col = slice * S + g * 8;
  const uint32_t img_col = (uint32_t)(col >> 5) * tc::IMG_BLOCK + (uint32_t)((col & 31) >> 3) * 128u;
  const size_t img_tile = (size_t)(d >> 5) * tc::IMG_BLOCK;   // bytes per 128-row tile
  const float4* sx_first = sx + 2 * g + swap;
  const float4* sx_second = sx + 2 * g + (swap ^ 1);
  for (int lrow = tid / TPR; lrow < n_b; lrow += NTHR / TPR) {
    const int s = sptr[lrow];
    const int e_true = sptr[lrow + 1];
    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
    if (fast) {
      for (int q = s; q < e_true; ++q) {
        const int off = ssrc[q];
        const float4 v = sx_first[off], w = sx_second[off];
        a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
        a2.x += w.x; a2.y += w.y; a2.z += w.z; a2.w += w.w;
      }
    } else {   // a mesh with more edges than the staging area holds, or a list cut by the edge capacity
      const int e = min(e_true, n_e);
      for (int q = s; q < e; ++q) {
        const unsigned nb = (unsigned)(__ldg(src + e0 + q) - r0);
        const int off = (int)(nb < (unsigned)n_b ? nb : (unsigned)n_b) * s4;
        const float4 v = sx_first[off], w = sx_second[off];
        a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
        a2.x += w.x; a2.y += w.y; a2.z += w.z; a2.w += w.w;
      }
    }
    const float4 lo4 = swap ? a2 : a1, hi4 = swap ? a1 : a2;   // columns col .. col+3 and col+4 .. col+7
    const float cnt = fmaxf((float)(e_true - s), 1.f);
    const float rc = __frcp_rn(cnt);
    const float4 o0 = make_float4(div_count(lo4.x, cnt, rc), div_count(lo4.y, cnt, rc), div_count(lo4.z, cnt, rc), div_count(lo4.w, cnt, rc));
    const float4 o1 = make_float4(div_count(hi4.x, cnt, rc), div_count(hi4.y, cnt, rc), div_count(hi4.z, cnt, rc), div_count(hi4.w, cnt, rc));
    const int row = r0 + lrow;
    if (agg) {
      float4* dst = reinterpret_cast<float4*>(agg + (size_t)row * d + col);
      dst[0] = o0;
      dst[1] = o1;
    }
    if (agg_image) {
      uint4 hi, lo;
      tc::split2(o0.x, o0.y, hi.x, lo.x);
      tc::split2(o0.z, o0.w, hi.y, lo.y);
      tc::split2(o1.x, o1.y, hi.z, lo.z);
      tc::split2(o1.z, o1.w, hi.w, lo.w);
      uint8_t* pimg = agg_image + (size_t)(row >> 7) * img_tile + (uint32_t)((row & 127) >> 3) * 512u + (uint32_t)(row & 7) * 16u + img_col;
      *reinterpret_cast<uint4*>(pimg) = hi;
      *reinterpret_cast<uint4*>(pimg + tc::IMG_PART) = lo;
    }
  }
  GM_STAMP(5, clock64()); GM_STAMP(6, gtimer());
  pdl_launch_dependents();
}

// The same aggregation with one CTA per (mesh, GROUP of consecutive slices): the mesh's CSR is staged once for the group and the x slices
// are double buffered - the cp.async copies of slice k + 1 are in flight while slice k is summed and stored, inside one CTA instead of
// between two independent ones.  2 x (F + 1) x 128 bytes + CSR = 216 KB for 800 faces, 1 024 threads, one CTA per SM.  Large batches only
// (a group of >= 2 slices per CTA must still leave >= 2 CTAs per SM's worth of work): see launch_gather_mean for when it is taken.
template <int S, int NTHR>
__global__ void __launch_bounds__(NTHR, 1) gather_mean_smem_group_kernel(const float* __restrict__ x, int d, const int* __restrict__ indptr,
                                                                         const int* __restrict__ src, long long cap, float* __restrict__ agg,
                                                                         uint8_t* __restrict__ agg_image, const int* __restrict__ face_off,
                                                                         int rows_cap, int edges_cap, int group, long long* __restrict__ trace) {
  extern __shared__ float4 sx[];
  constexpr int s4 = S >> 2;
  const int buf_stride = (rows_cap + 1) * s4;   // float4 per buffer
  int* sptr = reinterpret_cast<int*>(sx + 2 * (size_t)buf_stride);
  unsigned short* ssrc = reinterpret_cast<unsigned short*>(sptr + rows_cap + 1);
  pdl_wait();
  GM_STAMP(0, gtimer()); GM_STAMP(1, clock64()); GM_STAMP(7, smid());
  const int nslices = d / S;
  const int ngroups = (nslices + group - 1) / group;
  const int b = blockIdx.x / ngroups, grp = blockIdx.x % ngroups;
  const int slice0 = grp * group, slice1 = min(nslices, slice0 + group);
  const int r0 = face_off[b];
  const int n_b = min(face_off[b + 1] - r0, rows_cap);
  const int tid = threadIdx.x;
  const int c = tid % s4;
  const uint32_t sx_addr = tc::smem_u32(sx);
  auto issue = [&](int slice, int buf) {   // the x slice of this mesh -> buffer buf, one commit group
    const float* xs = x + (size_t)r0 * d + slice * S;
    const uint32_t base = sx_addr + (uint32_t)buf * (uint32_t)buf_stride * 16u;
    for (int i = tid; i < n_b * s4; i += NTHR) {
      const int row = i / s4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + (uint32_t)i * 16u), "l"(xs + (size_t)row * d + c * 4) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(slice0, 0);
  if (tid < s4) {   // the row an invalid source id points at, in both buffers
    sx[n_b * s4 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    sx[buf_stride + n_b * s4 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int e0 = indptr[r0];
  const long long e1_true = indptr[r0 + n_b];
  const int n_e = (int)((e1_true < cap ? e1_true : cap) - e0);
  const bool fast = n_e <= edges_cap && e1_true <= cap && (rows_cap + 1) * s4 <= 65536;
  for (int i = tid; i <= n_b; i += NTHR) sptr[i] = indptr[r0 + i] - e0;
  if (fast) {
    for (int i = tid; i < n_e; i += NTHR) {
      const unsigned nb = (unsigned)(__ldg(src + e0 + i) - r0);
      ssrc[i] = (unsigned short)((nb < (unsigned)n_b ? nb : (unsigned)n_b) * s4);
    }
  }
  constexpr int TPR = s4 / 2;
  const int g = tid % TPR;
  const int swap = (tid / TPR) & 1;
  const size_t img_tile = (size_t)(d >> 5) * tc::IMG_BLOCK;
  GM_STAMP(3, clock64());
  for (int slice = slice0; slice < slice1; ++slice) {
    const int buf = (slice - slice0) & 1;
    if (slice + 1 < slice1) {
      issue(slice + 1, buf ^ 1);   // (its last readers passed the barrier that ends the previous iteration)
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (slice == slice0) GM_STAMP(4, clock64());
    const int col = slice * S + g * 8;
    const uint32_t img_col = (uint32_t)(col >> 5) * tc::IMG_BLOCK + (uint32_t)((col & 31) >> 3) * 128u;
    const float4* sx_first = sx + (size_t)buf * buf_stride + 2 * g + swap;
    const float4* sx_second = sx + (size_t)buf * buf_stride + 2 * g + (swap ^ 1);
    for (int lrow = tid / TPR; lrow < n_b; lrow += NTHR / TPR) {
      const int s = sptr[lrow];
      const int e_true = sptr[lrow + 1];
      float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
      if (fast) {
        for (int q = s; q < e_true; ++q) {
          const int off = ssrc[q];
          const float4 v = sx_first[off], w = sx_second[off];
          a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
          a2.x += w.x; a2.y += w.y; a2.z += w.z; a2.w += w.w;
        }
      } else {
        const int e = min(e_true, n_e);
        for (int q = s; q < e; ++q) {
          const unsigned nb = (unsigned)(__ldg(src + e0 + q) - r0);
          const int off = (int)(nb < (unsigned)n_b ? nb : (unsigned)n_b) * s4;
          const float4 v = sx_first[off], w = sx_second[off];
          a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
          a2.x += w.x; a2.y += w.y; a2.z += w.z; a2.w += w.w;
        }
      }
      const float4 lo4 = swap ? a2 : a1, hi4 = swap ? a1 : a2;
      const float cnt = fmaxf((float)(e_true - s), 1.f);
      const float rc = __frcp_rn(cnt);
      const float4 o0 = make_float4(div_count(lo4.x, cnt, rc), div_count(lo4.y, cnt, rc), div_count(lo4.z, cnt, rc), div_count(lo4.w, cnt, rc));
      const float4 o1 = make_float4(div_count(hi4.x, cnt, rc), div_count(hi4.y, cnt, rc), div_count(hi4.z, cnt, rc), div_count(hi4.w, cnt, rc));
      const int row = r0 + lrow;
      if (agg) {
        float4* dst = reinterpret_cast<float4*>(agg + (size_t)row * d + col);
        dst[0] = o0;
        dst[1] = o1;
      }
      if (agg_image) {
        uint4 hi, lo;
        tc::split2(o0.x, o0.y, hi.x, lo.x);
        tc::split2(o0.z, o0.w, hi.y, lo.y);
        tc::split2(o1.x, o1.y, hi.z, lo.z);
        tc::split2(o1.z, o1.w, hi.w, lo.w);
        uint8_t* pimg = agg_image + (size_t)(row >> 7) * img_tile + (uint32_t)((row & 127) >> 3) * 512u + (uint32_t)(row & 7) * 16u + img_col;
        *reinterpret_cast<uint4*>(pimg) = hi;
        *reinterpret_cast<uint4*>(pimg + tc::IMG_PART) = lo;
      }
    }
    __syncthreads();   // every read of this buffer is done before the iteration after next refills it
  }
  GM_STAMP(5, clock64()); GM_STAMP(6, gtimer());
  pdl_launch_dependents();
}

// Round 2 version of the shared-memory aggregation.  A clock64 trace of the kernels above (MESHTOK_GATHER_TRACE, scripts/gpu_gather_trace.py)
// showed where their time goes at the bench workload: per (mesh, 32-column slice) 2.3 k cycles issuing the x copies, 5.6 k staging the CSR
// (three dependent global round trips: face_off -> indptr -> src) and 15 k in the row loop, i.e. ~2.4 k cycles per 8 rows per warp at
// ~0.7 instructions per cycle and scheduler - the loop is ISSUE bound, not memory bound: four threads per face with 8 columns each spend
// ~190 instructions per row slice, half of them per-row overhead (row pointers, loop control, source ids, reciprocal, image addresses)
// that every one of the four repeats.  Here TWO threads share a face, 16 columns each (four 16-byte shared-memory reads per neighbour),
// the mean is sum * (1 / count) (one rounding more than the division, exact for the 1, 2, 4 neighbours of most faces), and the
// mesh's first edge comes from the per-mesh edge offsets the topology kernel leaves next to the face offsets (one round trip less).
// One CTA per (mesh, group of `group` consecutive slices), slices double buffered when group >= 2.  The eight threads of a quarter warp
// (four faces x two halves) read eight different 16-byte chunks in every instruction whatever rows they come from: a face takes its two
// 8-column pairs in the order (face & 1) and the two chunks of a pair in the order (face >> 1) & 1 - no bank conflicts.
// (Later finding, same tool: with two threads per face the row loop is no longer issue bound - a loop with half the instructions per
//  neighbour measured 2 % SLOWER - but bound by the SM's shared-memory / L1 data pipe: a CTA alone on its SM takes 8.3 k cycles for the
//  loop, two take 16.6 k; 780 16-byte-per-lane reads + bank conflicts + image stores keep that pipe ~58 % busy.  DESIGN section 9.)
// ---- per-mesh local CSR blocks (round 2, last session) -------------------------------------------------------------------------------
// clock64 trace of the wide kernel at C2: of a CTA's ~23 k cycles, ~6 k went into its prologue - face / edge offsets, then row pointers
// and source ids (dependent global round trips), then the conversion into shared memory - repeated by every (mesh, slice) CTA of all
// five layers.  local_csr_kernel (or, on the default path, the one-launch topology kernel itself) does that conversion ONCE per batch and mesh, into a block with the aggregation kernel's shared-memory
// layout, so the CTA's prologue is one bulk copy (cp.async.bulk, completion on an mbarrier, issued ahead of griddepcontrol.wait).
__host__ __device__ inline int lcsr_ptr_bytes(int F) { return (((F + 1) * 4) + 15) & ~15; }
__host__ __device__ inline int lcsr_src_bytes(int F) { return ((5 * F * 2) + 15) & ~15; }
__host__ __device__ inline int lcsr_stride_bytes(int F) { return 16 + lcsr_ptr_bytes(F) + lcsr_src_bytes(F); }
static inline void gather_tma_geometry(int F, int* box_rows, int* n_boxes) {   // boxes of at most 256 rows x 32 columns
  *n_boxes = (F + 255) / 256;
  *box_rows = (F + *n_boxes - 1) / *n_boxes;
}

constexpr int LCSR_PARTS = 4;   // CTAs per mesh
__global__ void __launch_bounds__(256) local_csr_kernel(const int* __restrict__ indptr, const int* __restrict__ src, long long cap,
                                                        const int* __restrict__ face_off, const int* __restrict__ edge_off, int F,
                                                        int zero_row, uint8_t* __restrict__ lcsr) {
  const int b = blockIdx.x / LCSR_PARTS, part = blockIdx.x % LCSR_PARTS;
  const int edges_cap = 5 * F;
  uint8_t* blk = lcsr + (size_t)b * lcsr_stride_bytes(F);
  int* hdr = reinterpret_cast<int*>(blk);
  int* sptr = reinterpret_cast<int*>(blk + 16);
  unsigned short* ssrc = reinterpret_cast<unsigned short*>(blk + 16 + lcsr_ptr_bytes(F));
  const int r0 = face_off[b];
  const int n_b = min(face_off[b + 1] - r0, F);
  const int e0 = edge_off[b];
  const long long e1_true = edge_off[b + 1];
  const int n_e = (int)((e1_true < cap ? e1_true : cap) - e0);
  const bool fast = n_e <= edges_cap && e1_true <= cap && (zero_row + 1) * 8 <= 65536;
  const int t = part * blockDim.x + threadIdx.x, nt = LCSR_PARTS * blockDim.x;
  if (t == 0) { hdr[0] = n_e; hdr[1] = fast ? 1 : 0; hdr[2] = e0; hdr[3] = 0; }
  MT_DCHECK(n_b >= 0 && n_b <= F && n_e >= 0 && r0 >= 0);
  for (int i = t; i <= n_b; i += nt) { sptr[i] = indptr[r0 + i] - e0; MT_DCHECK(sptr[i] >= 0 && (long long)e0 + sptr[i] <= e1_true); }
  if (fast)
    for (int q = t; q < n_e; q += nt) {
      MT_DCHECK(q < edges_cap && (long long)e0 + q < cap);
      const unsigned nb = (unsigned)(src[e0 + q] - r0);   // (always < n_b: edges are intra-mesh; a corrupt list reads the zero row)
      ssrc[q] = (unsigned short)((nb < (unsigned)n_b ? nb : (unsigned)zero_row) * 8u);
    }
}

template <int NTHR, int CTAS>
__global__ void __launch_bounds__(NTHR, CTAS) gather_mean_smem_wide_kernel(const float* __restrict__ x, int d, const int* __restrict__ indptr,
                                                                           const int* __restrict__ src, long long cap, float* __restrict__ agg,
                                                                           uint8_t* __restrict__ agg_image, const int* __restrict__ face_off,
                                                                           const int* __restrict__ edge_off, int rows_cap, int edges_cap, int group,
                                                                           int nbuf, long long* __restrict__ trace, int exp,
                                                                           const __grid_constant__ CUtensorMap tmap, int box_rows, int n_boxes, int early,
                                                                           const uint8_t* __restrict__ lcsr, int stagger_ns, int n_sm) {
  // [nbuf][buf_rows + 1][8] float4 x slices (row buf_rows stays zero: what an invalid source id points at) | [rows_cap + 1] local row
  // pointers | [edges_cap] source row * 8 (u16) | [2] mbarriers.  box_rows > 0: the slices arrive by TMA (n_boxes boxes of box_rows x 32
  // columns, one elected thread, completion on an mbarrier); box_rows == 0: by 16-byte cp.async copies of all threads.
  extern __shared__ __align__(128) float4 sx[];
  constexpr int S = 32, s4 = S >> 2;
#ifndef MESHTOK_EXPERIMENTS
  exp = 0;   // (timing experiments that skip work - and produce wrong rows - exist only in the -DMESHTOK_EXPERIMENTS build)
#endif
  const bool tma = box_rows > 0;
  const int buf_rows = tma ? box_rows * n_boxes : rows_cap;
  const int buf_stride = (buf_rows + 1) * s4 + ((buf_rows + 1) & 1) * s4;   // float4 per buffer, a multiple of 256 bytes: every box lands 128-byte aligned
  // behind the slices: the mesh's local CSR block in the layout of local_csr_kernel (header | row pointers | source ids), the mbarriers
  uint8_t* blk = reinterpret_cast<uint8_t*>(sx + (size_t)nbuf * buf_stride);
  int* hdr = reinterpret_cast<int*>(blk);
  int* sptr = reinterpret_cast<int*>(blk + 16);
  unsigned short* ssrc = reinterpret_cast<unsigned short*>(blk + 16 + lcsr_ptr_bytes(rows_cap));
  uint64_t* full = reinterpret_cast<uint64_t*>(blk + lcsr_stride_bytes(rows_cap));   // [nbuf] slices, [2] the CSR block
  // Programmatic dependent launch: this kernel's CTAs become resident while the linear in front of it is still in its last round.  The
  // CSR of the mesh (face / edge offsets, row pointers, source ids) was written by the topology kernels, which completed before that
  // linear passed ITS wait, so it is staged BEFORE griddepcontrol.wait (L2 loads, ld.global.cg: no line an earlier launch left in this
  // SM's L1 can be taken for it); only the x slices - the linear's output - wait.  With TMA the last warp alone waits and issues the
  // copies (every read of x goes through the mbarrier its copies complete); without, every thread copies, so every thread waits.
  if (!early) pdl_wait();   // MESHTOK_GATHER_EARLY=0: everything behind the wait, as before (A/B)
  GM_STAMP(0, gtimer()); GM_STAMP(1, clock64()); GM_STAMP(7, smid());
  const int nslices = d / S;
  const int ngroups = (nslices + group - 1) / group;
  const int b = blockIdx.x / ngroups, grp = blockIdx.x % ngroups;
  const int slice0 = grp * group, slice1 = min(nslices, slice0 + group);
  const int tid = threadIdx.x;
  constexpr int issuer = NTHR - 32;   // first lane of the last warp
  const bool use_l = tma && lcsr != nullptr;
  if (tma && tid == issuer) {
    for (int k = 0; k < nbuf; ++k) tc::mbar_init(&full[k], 1);
    tc::mbar_init(&full[2], 1);
    tc::fence_barrier_init();
    tc::fence_proxy_async_smem();
    if (use_l) {   // the mesh's CSR block: written once per batch next to the topology kernels - nothing the kernel in front of this one touches
      tc::mbar_arrive_expect_tx(&full[2], (uint32_t)lcsr_stride_bytes(rows_cap));
      tc::bulk_g2s(blk, lcsr + (size_t)b * lcsr_stride_bytes(rows_cap), (uint32_t)lcsr_stride_bytes(rows_cap), &full[2]);
    }
  }
  const int r0 = __ldcg(face_off + b);
  const int n_b = min(__ldcg(face_off + b + 1) - r0, rows_cap);
  const int c = tid % s4;
  const uint32_t sx_addr = tc::smem_u32(sx);
  auto issue = [&](int slice, int buf) {   // the x slice of this mesh -> buffer buf
    if (tma) {
      if (tid == issuer) {   // rows beyond the mesh (the next mesh's, or zeros beyond the tensor) land in the buffer too and are never referenced
        // only the boxes this mesh reaches into: a 100-face mesh of a batch whose largest has 800 loads one box of four
        const int nb = min(n_boxes, (n_b + box_rows - 1) / box_rows);
        tc::mbar_arrive_expect_tx(&full[buf], (uint32_t)nb * (uint32_t)box_rows * 128u);
        for (int i = 0; i < nb; ++i)
          tc::tma_load_2d(sx + (size_t)buf * buf_stride + (size_t)i * box_rows * s4, &tmap, slice * S, r0 + i * box_rows, &full[buf]);
      }
      return;
    }
    const float* xs = x + (size_t)r0 * d + slice * S;
    const uint32_t base = sx_addr + (uint32_t)buf * (uint32_t)buf_stride * 16u;
    for (int i = tid; i < n_b * s4; i += NTHR) {
      const int row = i / s4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + (uint32_t)i * 16u), "l"(xs + (size_t)row * d + c * 4) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (tma && tid >= issuer) {   // the last warp: wait for the producer of x, first slice on its way - then its share of the CSR
    pdl_wait();
    // Two CTAs share an SM; started together they load together (each at half the SM's L2 bandwidth) and then loop together (each at
    // half the shared-memory pipe).  The second CTA of an SM's first wave issues its slice a little later: the first one gets the full
    // load bandwidth and starts its loop while the second one's slice is still in flight.
    if (stagger_ns > 0 && ((blockIdx.x / n_sm) & 1) != 0 && blockIdx.x < 2 * n_sm) __nanosleep(stagger_ns);
    issue(slice0, 0);
    __syncwarp();
  }
  GM_STAMP(2, 0);
  const int zero_row = tma ? buf_rows : n_b;
  if (tid < s4)
    for (int k = 0; k < nbuf; ++k) sx[(size_t)k * buf_stride + zero_row * s4 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);   // the row an invalid source id points at
  int e0, n_e;
  bool fast;
  if (use_l) {
    __syncthreads();                 // the barriers exist
    tc::mbar_wait(&full[2], 0);      // header, row pointers and source ids have landed
    n_e = hdr[0]; fast = hdr[1] != 0; e0 = hdr[2];
    MT_DCHECK(n_e >= 0 && (!fast || n_e <= edges_cap) && n_b >= 0 && n_b <= rows_cap);
    MT_DCHECK(sptr[0] == 0 && (!fast || sptr[n_b] == n_e));
  } else {
    e0 = __ldcg(edge_off + b);                                 // = indptr[r0]
    const long long e1_true = __ldcg(edge_off + b + 1);        // = indptr[r0 + n_b] (n_b is the whole mesh: rows_cap >= every mesh)
    n_e = (int)((e1_true < cap ? e1_true : cap) - e0);
    fast = n_e <= edges_cap && e1_true <= cap && (buf_rows + 1) * s4 <= 65536;
    for (int i = tid; i <= n_b; i += NTHR) sptr[i] = __ldcg(indptr + r0 + i) - e0;
    if (fast) {
      for (int base = tid; base < n_e; base += NTHR * 4) {
        int t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = base + j * NTHR < n_e ? __ldcg(src + e0 + base + j * NTHR) : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (base + j * NTHR < n_e) {
            const unsigned nb = (unsigned)(t[j] - r0);   // (always < n_b: edges are intra-mesh; a corrupt list reads the zero row)
            ssrc[base + j * NTHR] = (unsigned short)((nb < (unsigned)n_b ? nb : (unsigned)zero_row) * s4);
          }
      }
    }
  }
  if (!tma) {
    pdl_wait();
    issue(slice0, 0);
  }
  GM_STAMP(3, clock64());
  const size_t img_tile = (size_t)(d >> 5) * tc::IMG_BLOCK;   // bytes per 128-row tile
  for (int slice = slice0; slice < slice1; ++slice) {
    const int buf = nbuf == 2 ? ((slice - slice0) & 1) : 0;
    if (nbuf == 2 && slice + 1 < slice1) issue(slice + 1, buf ^ 1);   // (its last readers passed the barrier that ends the previous iteration)
    const long long t_wait0 = trace ? clock64() : 0;
    if (tma) {
      __syncthreads();   // the CSR and the zero rows (first slice)
      tc::mbar_wait(&full[buf], (uint32_t)(((slice - slice0) >> (nbuf == 2 ? 1 : 0)) & 1));
    } else {
      if (nbuf == 2 && slice + 1 < slice1) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }
    if (slice == slice0) GM_STAMP(4, clock64());
    if (trace != nullptr && tid == 0 && blockIdx.x < 1024 && slice > slice0) trace[blockIdx.x * 8 + 2] += clock64() - t_wait0;   // waiting for slices 1.. (thread 0)
    const float4* sxb = sx + (size_t)buf * buf_stride;
    for (int item = tid; item < 2 * n_b; item += NTHR) {
      const int lrow = item >> 1, h = item & 1;
      const int rot = lrow & 1, sw = (lrow >> 1) & 1;
      // chunk (16 bytes = 4 columns) ids of this thread's four reads: pair p = (k + rot) & 1 of its half, inside the pair first chunk sw
      const float4* p0a = sxb + 4 * h + 2 * rot + sw;
      const float4* p0b = sxb + 4 * h + 2 * rot + (sw ^ 1);
      const float4* p1a = sxb + 4 * h + 2 * (rot ^ 1) + sw;
      const float4* p1b = sxb + 4 * h + 2 * (rot ^ 1) + (sw ^ 1);
      const int s = sptr[lrow];
      const int e_true = sptr[lrow + 1];
      MT_DCHECK(lrow < n_b && s >= 0 && e_true >= s);
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
      const int e = fast ? e_true : min(e_true, n_e);
      for (int q = s; q < e; ++q) {
        int off;
        MT_DCHECK(q >= 0 && (!fast || q < n_e));
        if (fast) { off = ssrc[q]; MT_DCHECK(off >= 0 && off <= zero_row * s4 && (off & (s4 - 1)) == 0); }
        else {   // a mesh with more edges than the staging area holds, or a list cut by the edge capacity
          const unsigned nb = (unsigned)(__ldg(src + e0 + q) - r0);
          off = (int)(nb < (unsigned)n_b ? nb : (unsigned)zero_row) * s4;
        }
        if (exp == 2) { a0.x += __int_as_float(off); continue; }   // experiment: no neighbour row reads
        const float4 v0 = p0a[off], v1 = p0b[off], v2 = p1a[off], v3 = p1b[off];
        add4(a0, v0);   // packed fp32 adds (FADD2): the same roundings in half the issue slots
        add4(a1, v1);
        add4(a2, v2);
        add4(a3, v3);
      }
      const float rc = __frcp_rn(fmaxf((float)(e_true - s), 1.f));
      const int row = r0 + lrow;
      if (exp == 1 && a0.x + a1.y + a2.z + a3.w != 123.456f) continue;   // experiment: no epilogue, no stores
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4 fa = k == 0 ? a0 : a2, fb = k == 0 ? a1 : a3;   // first-read / second-read chunk of the pair
        const float4 lo4 = sw ? fb : fa, hi4 = sw ? fa : fb;           // columns col .. col+3 and col+4 .. col+7
        const float4 o0 = make_float4(lo4.x * rc, lo4.y * rc, lo4.z * rc, lo4.w * rc);
        const float4 o1 = make_float4(hi4.x * rc, hi4.y * rc, hi4.z * rc, hi4.w * rc);
        const int col = slice * S + h * 16 + ((k + rot) & 1) * 8;
        if (agg) {
          float4* dst = reinterpret_cast<float4*>(agg + (size_t)row * d + col);
          dst[0] = o0;
          dst[1] = o1;
        }
        if (agg_image) {
          uint4 hi, lo;
          tc::split2(o0.x, o0.y, hi.x, lo.x);
          tc::split2(o0.z, o0.w, hi.y, lo.y);
          tc::split2(o1.x, o1.y, hi.z, lo.z);
          tc::split2(o1.z, o1.w, hi.w, lo.w);
          const uint32_t img_col = (uint32_t)(col >> 5) * tc::IMG_BLOCK + (uint32_t)((col & 31) >> 3) * 128u;
          uint8_t* pimg = agg_image + (size_t)(row >> 7) * img_tile + (uint32_t)((row & 127) >> 3) * 512u + (uint32_t)(row & 7) * 16u + img_col;
          if (exp == 3 && (hi.x ^ lo.y) != 0x12345678u) continue;   // experiment: the whole epilogue but no stores
          *reinterpret_cast<uint4*>(pimg) = hi;
          *reinterpret_cast<uint4*>(pimg + tc::IMG_PART) = lo;
        }
      }
    }
    if (nbuf == 2) __syncthreads();   // every read of this buffer is done before the iteration after next refills it
    else if (slice + 1 < slice1) { __syncthreads(); issue(slice + 1, 0); }
  }
  GM_STAMP(5, clock64()); GM_STAMP(6, gtimer());
  pdl_launch_dependents();
}

template <int PER2>  // float2 pairs per lane (d <= 64 * PER2, d even)
__global__ void __launch_bounds__(256) row_norm_kernel(float* __restrict__ x, int d, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, uint8_t* __restrict__ image,
                                                       const Counts* counts) {
  const int lane = lane_id();
  const int wpb = blockDim.x >> 5;
  const int n_rows = counts->n_faces;
  const int d2 = d >> 1;
  for (int row = blockIdx.x * wpb + warp_id(); row < n_rows; row += gridDim.x * wpb) {
    float2* xr = reinterpret_cast<float2*>(x + (size_t)row * d);
    float2 v[PER2];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < PER2; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < d2 ? xr[c] : make_float2(0.f, 0.f);
      ss += v[i].x * v[i].x + v[i].y * v[i].y;
    }
    ss = warp_sum(ss);
    const float den = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int i = 0; i < PER2; ++i) { v[i].x = __fdiv_rn(v[i].x, den); v[i].y = __fdiv_rn(v[i].y, den); }
    if (gamma != nullptr) {
      // SiLU then LayerNorm (biased variance, eps 1e-5)
      float mean = 0.f;
#pragma unroll
      for (int i = 0; i < PER2; ++i) {
        const int c = lane + 32 * i;
        const float2 t = v[i];
        v[i] = c < d2 ? make_float2(__fdiv_rn(t.x, 1.f + expf(-t.x)), __fdiv_rn(t.y, 1.f + expf(-t.y))) : make_float2(0.f, 0.f);
        mean += v[i].x + v[i].y;
      }
      mean = warp_sum(mean) / (float)d;
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < PER2; ++i) {
        const int c = lane + 32 * i;
        const float tx = c < d2 ? v[i].x - mean : 0.f, ty = c < d2 ? v[i].y - mean : 0.f;
        var += tx * tx + ty * ty;
      }
      var = warp_sum(var) / (float)d;
      const float rstd = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
      for (int i = 0; i < PER2; ++i) {
        const int c = lane + 32 * i;
        if (c < d2) {
          const float2 g = *reinterpret_cast<const float2*>(gamma + 2 * c), b = *reinterpret_cast<const float2*>(beta + 2 * c);
          v[i].x = (v[i].x - mean) * rstd * g.x + b.x;
          v[i].y = (v[i].y - mean) * rstd * g.y + b.y;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < PER2; ++i) {
      const int c = lane + 32 * i;
      if (c < d2) {
        xr[c] = v[i];
        if (image) tc::image_store2(image, d, row, 2 * c, v[i].x, v[i].y);
      }
    }
  }
}

}  // namespace

static long long* g_gm_trace = nullptr;
static int g_gm_trace_k = -2;
static long long g_gm_launches = 0;

int gather_mean_read_trace(long long* out, int n) {
  if (g_gm_trace == nullptr) return 0;
  cudaDeviceSynchronize();
  cudaMemcpy(out, g_gm_trace, sizeof(long long) * (size_t)min(n, 1024 * 8), cudaMemcpyDeviceToHost);
  return min(n, 1024 * 8);
}

size_t local_csr_stride(int F) { return (size_t)lcsr_stride_bytes(F); }
void local_csr_layout(int F, int* ptr_bytes, int* zero_row) {
  int box_rows, n_boxes;
  gather_tma_geometry(F, &box_rows, &n_boxes);
  *ptr_bytes = lcsr_ptr_bytes(F);
  *zero_row = box_rows * n_boxes;
}

int launch_local_csr(const int* indptr, const int* src, long long edge_capacity, const int* face_off, int B, int F, void* lcsr, cudaStream_t stream) {
  if (B == 0 || lcsr == nullptr) return MESHTOK_OK;
  int box_rows, n_boxes;
  gather_tma_geometry(F, &box_rows, &n_boxes);
  local_csr_kernel<<<B * LCSR_PARTS, 256, 0, stream>>>(indptr, src, edge_capacity, face_off, face_off + 2 * (size_t)(B + 1), F, box_rows * n_boxes,
                                          static_cast<uint8_t*>(lcsr));
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int launch_gather_mean(const float* x, int d, const int* indptr, const int* src, long long edge_capacity, float* agg,
                       void* agg_image, const Counts* counts, int max_rows, const int* face_off, int B, int F, cudaStream_t stream, const void* lcsr) {
  MT_CHECK_ARG(d % 4 == 0 && d <= 1024, "gather_mean: feature width %d must be a multiple of 4 and <= 1024", d);
  if (max_rows == 0) return MESHTOK_OK;
  if (g_gm_trace_k == -2) {   // timing experiments only: MESHTOK_GATHER_TRACE=k stamps the k-th aggregation launch of every step (5 per step)
    const char* e = getenv("MESHTOK_GATHER_TRACE");
    g_gm_trace_k = e ? atoi(e) : -1;
    if (g_gm_trace_k >= 0) { MT_CHECK_CUDA(cudaMalloc(&g_gm_trace, 1024 * 8 * sizeof(long long))); MT_CHECK_CUDA(cudaMemset(g_gm_trace, 0, 1024 * 8 * sizeof(long long))); }
  }
  long long* trace = (g_gm_trace_k >= 0 && (g_gm_launches++ % 5) == g_gm_trace_k) ? g_gm_trace : nullptr;
  // per-mesh slices in shared memory when a 32-column slice of the largest mesh fits (two CTAs per SM up to 880 faces,
  // one up to 1 760); MESHTOK_GATHER_L2=1 keeps the warp-per-row kernel for A/B runs
  static int use_smem = -1;
  if (use_smem < 0) { const char* e = getenv("MESHTOK_GATHER_L2"); use_smem = (e && atoi(e) == 1) ? 0 : 1; }
  // slice width: 32 columns (2 CTAs per SM at 800 faces), or 16 with 256 threads (MESHTOK_GATHER_S=16: 3 CTAs per SM, A/B switch)
  static int slice_cols = -1;
  if (slice_cols < 0) { const char* e = getenv("MESHTOK_GATHER_S"); slice_cols = (e && atoi(e) == 16) ? 16 : 32; }
  const int S = slice_cols;
  const int edges_cap = 5 * F;   // staged source ids per mesh (u16); a mesh with more edges reads them from global memory
  const size_t smem = align_up((size_t)(F + 1) * S * sizeof(float) + (size_t)(F + 1) * sizeof(int) + (size_t)edges_cap * sizeof(unsigned short), 16);
  // one CTA per (mesh, slice): enough CTAs to fill the machine, or so little work that it does not matter
  const bool enough_ctas = (long long)B * (d / S) >= 64 || max_rows <= 8192;
  // Large batches take the double-buffered kernel, one CTA per (mesh, group of slices): by default from 512 meshes per launch (measured at
  // 1 024: 1.74 -> 1.54 ms over the five layers, bit-identical rows); MESHTOK_GATHER_GROUP=1 takes it whenever a group holds >= 2 slices,
  // =0 never (A/B, scripts/gpu_gather_group_ab.sh)
  // round 2: two threads per face, 16 columns each (gather_mean_smem_wide_kernel); MESHTOK_GATHER_WIDE=0 keeps the kernels below (A/B)
  static int use_wide = -1;
  if (use_wide < 0) { const char* e = getenv("MESHTOK_GATHER_WIDE"); use_wide = (e && atoi(e) == 0) ? 0 : 1; }
  if (use_wide && use_smem && face_off != nullptr && d % 32 == 0) {
    const int nslices = d / 32;
    // x slices by TMA (cp.async.bulk.tensor.2d over a tensor map of x): boxes of box_rows x 32 columns, at most 256 rows each
    static int use_tma = -1;
    if (use_tma < 0) { const char* e = getenv("MESHTOK_GATHER_TMA"); use_tma = (e && atoi(e) == 0) ? 0 : 1; }
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int n_boxes, box_rows;
    gather_tma_geometry(F, &box_rows, &n_boxes);
    const bool tma = use_tma && make_tmap_rows_f32(x, d, (long long)max_rows, box_rows, &tmap);
    if (!tma) { box_rows = 0; n_boxes = 0; }
    const int buf_rows = tma ? box_rows * n_boxes : F;
    const size_t per_buf = (size_t)((buf_rows + 1) + ((buf_rows + 1) & 1)) * 32 * sizeof(float);
    const size_t csr = (size_t)lcsr_stride_bytes(F) + 3 * sizeof(uint64_t);
    static int use_lcsr = -1;   // MESHTOK_GATHER_LCSR=0: every CTA converts its mesh's CSR itself, as before (A/B)
    if (use_lcsr < 0) { const char* e = getenv("MESHTOK_GATHER_LCSR"); use_lcsr = (e && atoi(e) == 0) ? 0 : 1; }
    const uint8_t* lc = (tma && use_lcsr) ? static_cast<const uint8_t*>(lcsr) : nullptr;
    // Slices per CTA.  MEASURED (B200, scripts/gpu_ab.sh; 64 and 1 024 meshes per launch): one slice per CTA with 512 threads and two
    // CTAs per SM - each slice arriving by TMA while the other CTA computes - is as fast as or faster than every grouping with
    // double-buffered slices (C2: 126.7 us over the five layers against 132 - 202 us for groups of 2 / 4 / 8; 1 024 meshes: 1.50 ms
    // against 1.58 - 1.73 ms), so groups are only formed when a single-slice CTA does not fit twice per SM (meshes above ~800 faces).
    static int forced_group = -2;
    if (forced_group == -2) { const char* e = getenv("MESHTOK_GATHER_WIDE_GROUP"); forced_group = e ? atoi(e) : -1; }
    int group = 1;
    double best = 1e30;
    for (int g = 1; g <= nslices; ++g) {
      const bool two_bufs = g >= 2;
      const size_t smem_g = align_up((two_bufs ? 2 : 1) * per_buf + csr, 128);
      if (smem_g > 226 * 1024) continue;
      const int per_sm = (g == 1 && 2 * (smem_g + 1024) <= 226 * 1024) ? 2 : 1;
      if (g == 1 && per_sm == 2) { best = 0.0; group = 1; break; }
      const long long ctas = (long long)B * ceil_div(nslices, g);
      const double rounds = (double)((ctas + 148 * per_sm - 1) / (148 * per_sm));
      const double cost = rounds * (g + 1.5);
      if (cost < best - 1e-9) { best = cost; group = g; }
    }
    if (forced_group >= 1 && best < 1e30) {
      group = min(forced_group, nslices);
      if (align_up((group >= 2 ? 2 : 1) * per_buf + csr, 128) > 226 * 1024) group = 1;
    }
    static int gexp = -1;
    if (gexp < 0) { const char* e = getenv("MESHTOK_GATHER_EXP"); gexp = e ? atoi(e) : 0; }
    static int early = -1;   // CSR staged ahead of griddepcontrol.wait (default); 0 = behind it (A/B)
    if (early < 0) { const char* e = getenv("MESHTOK_GATHER_EARLY"); early = (e && atoi(e) == 0) ? 0 : 1; }
    if (best < 1e30) {
      const int nbuf = group >= 2 ? 2 : 1;
      const size_t smem_w = align_up(nbuf * per_buf + csr, 128);
      const int ngroups = ceil_div(nslices, group);
      const int* edge_off = face_off + 2 * (size_t)(B + 1);
      // ns by which the second CTA of an SM's first wave delays its slice (MESHTOK_GATHER_STAGGER=0: off).  Measured at C2 on one box:
      // 128.9 -> 124.5 us over the five launches for 600, 1 200 and 2 000 ns alike.
      static int stagger = -1;
      if (stagger < 0) { const char* e = getenv("MESHTOK_GATHER_STAGGER"); stagger = e ? atoi(e) : 1000; }
      int n_sm = 148;
      {   // SM count of the current device, looked up once per device
        static int sm_of[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64) {
          if (sm_of[dev] == 0) { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sm_of[dev] = v; }
          if (sm_of[dev] > 0) n_sm = sm_of[dev];
        }
      }
      // (512 threads, two CTAs per SM) where a single-slice CTA fits twice; else 1 024 threads and the SM to itself - also with ONE slice per
      // CTA when two buffers of a large mesh do not fit (meshes of ~900 .. 1 700 faces; found by test_meshes_above_800_faces)
      if (group == 1 && 2 * (smem_w + 1024) <= 226 * 1024) {
        MT_ENSURE_SMEM((gather_mean_smem_wide_kernel<512, 2>), 120 * 1024);
        // No more CTAs than SMs (the 64-wide layer at 64 meshes: 128): the block scheduler would still put them two to an SM - 65 SMs busy,
        // 83 idle, two row loops sharing one shared-memory pipe (clock64 trace: 15.9 k cycles against 9.3 k alone).  A request of more
        // than half an SM's shared memory spreads them one per SM.
        static int spread = -1;
        if (spread < 0) { const char* e = getenv("MESHTOK_GATHER_SPREAD"); spread = (e && atoi(e) == 0) ? 0 : 1; }
        const size_t smem_launch = (spread && (long long)B * ngroups <= n_sm) ? (size_t)max((long long)smem_w, 118ll * 1024) : smem_w;
        MT_LAUNCH_PDL((gather_mean_smem_wide_kernel<512, 2>), B * ngroups, 512, smem_launch, stream, x, d, indptr, src, edge_capacity, agg,
                      static_cast<uint8_t*>(agg_image), face_off, edge_off, F, edges_cap, group, nbuf, trace, gexp, tmap, box_rows, n_boxes, early, lc, stagger, n_sm);
      } else {
        MT_ENSURE_SMEM((gather_mean_smem_wide_kernel<1024, 1>), 226 * 1024);
        MT_LAUNCH_PDL((gather_mean_smem_wide_kernel<1024, 1>), B * ngroups, 1024, smem_w, stream, x, d, indptr, src, edge_capacity, agg,
                      static_cast<uint8_t*>(agg_image), face_off, edge_off, F, edges_cap, group, nbuf, trace, gexp, tmap, box_rows, n_boxes, early, lc, stagger, n_sm);
      }
      return MESHTOK_OK;
    }
  }
  static int use_group = -1;
  if (use_group < 0) { const char* e = getenv("MESHTOK_GATHER_GROUP"); use_group = e ? atoi(e) : 2; }
  if (use_group != 0 && use_smem && face_off != nullptr && d % 32 == 0) {
    const int nslices = d / 32;
    // Slices per CTA.  One CTA (1 024 threads, 216 KB) per SM; a CTA's time ~ (its slices + half a slice for the first, un-overlapped
    // staging); the launch takes ceil(CTAs / 148) rounds of that.  Large batches end up with whole rows per CTA; at 64 meshes the
    // best cut is 128 CTAs (d = 256: 4 slices each), which beats 512 single-slice CTAs that each pay their own staging.
    int group = 1;
    if (use_group == 2 && B < 512) {
      group = 1;   // default below 512 meshes per launch: the per-slice kernel (MESHTOK_GATHER_GROUP=1 applies the cost model everywhere)
    } else {
      double best = 1e30;
      for (int g = 2; g <= nslices; ++g) {
        const long long ctas = (long long)B * ceil_div(nslices, g);
        const double cost = (double)((ctas + 147) / 148) * (g + 0.5);
        if (cost < best - 1e-9) { best = cost; group = g; }
      }
    }
    const size_t smem2 = align_up(2 * (size_t)(F + 1) * 32 * sizeof(float) + (size_t)(F + 1) * sizeof(int) + (size_t)edges_cap * sizeof(unsigned short), 16);
    if (group >= 2 && smem2 <= 226 * 1024) {
      MT_ENSURE_SMEM((gather_mean_smem_group_kernel<32, 1024>), 226 * 1024);
      const int ngroups = ceil_div(nslices, group);
      MT_LAUNCH_PDL((gather_mean_smem_group_kernel<32, 1024>), B * ngroups, 1024, smem2, stream, x, d, indptr, src, edge_capacity, agg,
                    static_cast<uint8_t*>(agg_image), face_off, F, edges_cap, group, trace);
      return MESHTOK_OK;
    }
  }
  if (use_smem && face_off != nullptr && d % S == 0 && smem <= 220 * 1024 && enough_ctas) {
    MT_ENSURE_SMEM((gather_mean_smem_kernel<32, 512, 2>), 220 * 1024);
    MT_ENSURE_SMEM((gather_mean_smem_kernel<16, 256, 3>), 220 * 1024);
    if (S == 32) {
      MT_LAUNCH_PDL((gather_mean_smem_kernel<32, 512, 2>), B * (d / S), 512, smem, stream, x, d, indptr, src, edge_capacity, agg,
                    static_cast<uint8_t*>(agg_image), face_off, F, edges_cap, trace);
    } else {
      MT_LAUNCH_PDL((gather_mean_smem_kernel<16, 256, 3>), B * (d / S), 256, smem, stream, x, d, indptr, src, edge_capacity, agg,
                    static_cast<uint8_t*>(agg_image), face_off, F, edges_cap, trace);
    }
    return MESHTOK_OK;
  }
  const int blocks = max(1, min(ceil_div(max_rows, 8), 148 * 16));
  const int chunks = ceil_div(d / 4, 32);
#define GM(C) MT_LAUNCH_PDL(gather_mean_kernel<C>, blocks, 256, 0, stream, x, d, indptr, src, edge_capacity, agg, static_cast<uint8_t*>(agg_image), counts)
  switch (chunks) {
    case 1: GM(1); break;
    case 2: GM(2); break;
    case 3: GM(3); break;
    case 4: GM(4); break;
    default: GM(8); break;
  }
#undef GM
  return MESHTOK_OK;
}

int launch_row_norm(float* x, int d, const float* gamma, const float* beta, void* image_out, const Counts* counts, int max_rows,
                    cudaStream_t stream) {
  MT_CHECK_ARG(d >= 2 && d % 2 == 0 && d <= 1024, "row_norm: feature width %d must be even and in [2, 1024]", d);
  MT_CHECK_ARG(image_out == nullptr || d % tc::IMG_KC == 0, "row_norm: image output needs a width that is a multiple of %d", tc::IMG_KC);
  const int blocks = max(1, min(ceil_div(max_rows, 8), 148 * 16));
  const int per2 = ceil_div(d / 2, 32);
#define RN(P) row_norm_kernel<P><<<blocks, 256, 0, stream>>>(x, d, gamma, beta, static_cast<uint8_t*>(image_out), counts)
  if (per2 <= 1) RN(1);
  else if (per2 <= 2) RN(2);
  else if (per2 <= 4) RN(4);
  else if (per2 <= 9) RN(9);
  else RN(16);
#undef RN
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

}  // namespace meshtok
