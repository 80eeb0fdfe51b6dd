// K2: gather face coordinates -> derived features -> discretise -> (embedding o project_in) tables.
//
// Replaces MeshAutoencoder.encode lines meshgpt_pytorch.py:711-741 (get_at gather, get_derived_face_features
// :155-183, discretize :187-201, the four nn.Embedding lookups, the 832/1040-wide concat and project_in).
// The embedding lookups followed by the linear are folded at load time into one (bins x dim) table per
// slot of the concat, so a face costs `num_slots` table-row reads (L2 resident, 1.5 MB) instead of a
// materialised (N, 832) activation and a GEMM.
//
// This file is compiled with -fmad=false: the discretisation chain must round exactly like the
// reference's separate fp32 ops ((t-lo)/(hi-lo), *n, -0.5, round-half-even; meshgpt_pytorch.py:197-201).
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace meshtok {

namespace {

// meshgpt_pytorch.py:197-201 with the CPU's float->int64 conversion behaviour for NaN / huge values
// (the x86 "indefinite integer" is INT64_MIN, which the clamp turns into bin 0).
__device__ __forceinline__ int discretize_rn(float t, float lo, float den, float nbins_f, int nbins) {
  float u = __fdiv_rn(__fsub_rn(t, lo), den);
  u = __fmul_rn(u, nbins_f);
  u = __fsub_rn(u, 0.5f);
  u = rintf(u);
  if (!(fabsf(u) < 9.2233720368547758e18f)) return 0;
  int q = (u <= 0.f) ? 0 : ((u >= nbins_f) ? nbins - 1 : (int)u);
  return min(q, nbins - 1);
}

__device__ __forceinline__ void unit3(const float v[3], float out[3]) {
  // F.normalize(p=2, eps=1e-12): v / max(|v|, eps)
  const float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
  const float d = fmaxf(n, 1e-12f);
  out[0] = __fdiv_rn(v[0], d);
  out[1] = __fdiv_rn(v[1], d);
  out[2] = __fdiv_rn(v[2], d);
}

// Two phases per chunk of ROWS face rows (small chunks so that the grid-stride loop balances over the SMs):
//   1. one THREAD per face: gather the corner positions, derive the features, discretise -> NSLOT bin ids in shared memory
//      (no lane repeats another lane's acosf / sqrt / divide);
//   2. one WARP per face, two faces in flight: x0[row] = bias + sum_s table_s[bin_s] in fixed slot order, 16 B per lane and
//      table row (the tables are L2 resident; the kernel is bound by the latency of these reads, so a lane keeps
//      2 faces x NSLOT independent 16-byte loads in flight).
template <int NVF>
__global__ void __launch_bounds__(256) face_embed_kernel(FeatParams p) {
  pdl_wait();
  constexpr int NSLOT = NVF * 3 + NVF + 1 + 3;
  constexpr int ROWS = 64;
  constexpr int RPW = ROWS / 8;   // rows per warp in phase 2
  __shared__ unsigned short sbins[ROWS][NSLOT];
  __shared__ int spf[ROWS];
  const int lane = lane_id(), warp = warp_id();
  const int n_rows = p.counts->n_faces;
  const int D = p.D, d4 = D >> 2;
  for (int base = blockIdx.x * ROWS; base < n_rows; base += gridDim.x * ROWS) {
    // ---- phase 1 ----
    const int row = base + threadIdx.x;
    if (threadIdx.x < ROWS && row < n_rows) {
      const int pf = p.row_face[row];
      const int b = pf / p.F;
      const int64_t* fptr = p.faces + (size_t)pf * NVF;
      float pos[NVF][3];
#pragma unroll
      for (int k = 0; k < NVF; ++k) {
        const long long v = fptr[k];
        const float* src = p.vertices + ((size_t)b * p.V + (size_t)v) * 3;
        pos[k][0] = src[0]; pos[k][1] = src[1]; pos[k][2] = src[2];
      }
      unsigned short* bins = sbins[threadIdx.x];
      // coordinates: slot = vertex*3 + xyz (meshgpt_pytorch.py:732-733)
#pragma unroll
      for (int k = 0; k < NVF; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          bins[k * 3 + c] = (unsigned short)discretize_rn(pos[k][c], p.coor_lo, p.coor_den, (float)p.n_coor, p.n_coor);
      // angles between position vectors k and k-1 (meshgpt_pytorch.py:151-153, 164-166)
      float un[NVF][3];
#pragma unroll
      for (int k = 0; k < NVF; ++k) unit3(pos[k], un[k]);
#pragma unroll
      for (int k = 0; k < NVF; ++k) {
        const int km = (k + NVF - 1) % NVF;
        float z = __fadd_rn(__fadd_rn(__fmul_rn(un[k][0], un[km][0]), __fmul_rn(un[k][1], un[km][1])), __fmul_rn(un[k][2], un[km][2]));
        z = (z != z) ? z : fminf(fmaxf(z, p.clip_lo), p.clip_hi);  // torch.clamp propagates NaN
        bins[NVF * 3 + k] = (unsigned short)discretize_rn(acosf(z), 0.f, p.angle_den, (float)p.n_angle, p.n_angle);
      }
      // edges: rows 0 and 1 of (p - prev); prev = one step back (tri) / two steps back (quad) (:168-172)
      constexpr int SH = (NVF == 3) ? 1 : 2;
      float e1[3], e2[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        e1[c] = __fsub_rn(pos[0][c], pos[(0 + NVF - SH) % NVF][c]);
        e2[c] = __fsub_rn(pos[1][c], pos[(1 + NVF - SH) % NVF][c]);
      }
      float cr[3];
      cr[0] = __fsub_rn(__fmul_rn(e1[1], e2[2]), __fmul_rn(e1[2], e2[1]));
      cr[1] = __fsub_rn(__fmul_rn(e1[2], e2[0]), __fmul_rn(e1[0], e2[2]));
      cr[2] = __fsub_rn(__fmul_rn(e1[0], e2[1]), __fmul_rn(e1[1], e2[0]));
      const float cn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(cr[0], cr[0]), __fmul_rn(cr[1], cr[1])), __fmul_rn(cr[2], cr[2])));
      bins[NVF * 4] = (unsigned short)discretize_rn(__fmul_rn(cn, 0.5f), 0.f, p.area_den, (float)p.n_area, p.n_area);
      const float cd = fmaxf(cn, 1e-12f);
#pragma unroll
      for (int c = 0; c < 3; ++c)
        bins[NVF * 4 + 1 + c] = (unsigned short)discretize_rn(__fdiv_rn(cr[c], cd), p.coor_lo, p.coor_den, (float)p.n_normal, p.n_normal);
      spf[threadIdx.x] = pf;
      if (p.face_coords_out) {
#pragma unroll
        for (int s = 0; s < NVF * 3; ++s) p.face_coords_out[(size_t)pf * (NVF * 3) + s] = bins[s];
      }
    }
    __syncthreads();
    // ---- phase 2 ----
    const int rows_here = min(ROWS, n_rows - base);
    for (int r0 = warp * RPW; r0 < min(rows_here, warp * RPW + RPW); r0 += 2) {
      const bool two = r0 + 1 < rows_here;
      const unsigned short* ba = sbins[r0];
      const unsigned short* bb = sbins[two ? r0 + 1 : r0];
      for (int i4 = lane; i4 < d4; i4 += 32) {
        const float4 bias = __ldg(reinterpret_cast<const float4*>(p.bias) + i4);
        float4 acc_a = bias, acc_b = bias;
        float4 ta[NSLOT], tb[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          ta[s] = __ldg(reinterpret_cast<const float4*>(p.tables + (size_t)(p.slot_off[s] + ba[s]) * D) + i4);
          tb[s] = __ldg(reinterpret_cast<const float4*>(p.tables + (size_t)(p.slot_off[s] + bb[s]) * D) + i4);
        }
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          acc_a.x += ta[s].x; acc_a.y += ta[s].y; acc_a.z += ta[s].z; acc_a.w += ta[s].w;
          acc_b.x += tb[s].x; acc_b.y += tb[s].y; acc_b.z += tb[s].z; acc_b.w += tb[s].w;
        }
        const long long ra = base + r0;
        if (p.x0) *reinterpret_cast<float4*>(p.x0 + (size_t)ra * D + i4 * 4) = acc_a;
        if (p.x0_image) tc::image_store4(p.x0_image, D, ra, i4 * 4, acc_a);
        if (two) {
          if (p.x0) *reinterpret_cast<float4*>(p.x0 + (size_t)(ra + 1) * D + i4 * 4) = acc_b;
          if (p.x0_image) tc::image_store4(p.x0_image, D, ra + 1, i4 * 4, acc_b);
        }
      }
    }
    __syncthreads();
  }
  pdl_launch_dependents();
}

}  // namespace

int launch_face_embed(const FeatParams& p, int max_rows, cudaStream_t stream) {
  MT_CHECK_ARG(p.D % 4 == 0, "face_embed: dim_codebook must be a multiple of 4 (got %d)", p.D);
  MT_CHECK_ARG(max(max(p.n_coor, p.n_angle), max(p.n_area, p.n_normal)) <= 65535, "face_embed: at most 65535 bins per feature");
  const int blocks = max(1, min(ceil_div(max_rows, 64), 148 * 2));
  if (p.nvf == 3) MT_LAUNCH_PDL(face_embed_kernel<3>, blocks, 256, 0, stream, p);
  else MT_LAUNCH_PDL(face_embed_kernel<4>, blocks, 256, 0, stream, p);
  return MESHTOK_OK;
}

}  // namespace meshtok
