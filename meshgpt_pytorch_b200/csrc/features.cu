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
#include <stdlib.h>

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

// Two kernels:
//   face_bins_kernel: one THREAD per face - gather the corner positions, derive the features, discretise -> NSLOT bin ids
//     (uint16) in global memory (no lane repeats another lane's acosf / sqrt / divide);
//   face_sum_kernel: one WARP per pair of faces - x0[row] = bias + sum_s table_s[bin_s] in fixed slot order.  The 2 * D/4
//     float4 columns of the pair are spread over the lanes so that all of them work; a lane keeps half of the table rows of
//     its column in flight at a time.  The tables are L2 resident and the kernel is bound by the latency of those reads, so
//     it is written for occupancy (<= 64 registers, no shared memory, no barriers).
template <int NVF>
__global__ void __launch_bounds__(128) face_bins_kernel(FeatParams p) {
  pdl_wait();
  constexpr int NSLOT = NVF * 3 + NVF + 1 + 3;
  // One thread per INPUT face (slot b*F+f of the padded layout / row g of the ragged one), not per packed row: the bins depend on
  // nothing the topology kernel computes, so this launch runs next to it on the side stream.  A face is skipped exactly when the
  // topology kernel treats it as padding: any slot == pad_id, or a vertex id outside the mesh (it flags the latter in the status word).
  const long long n_in = p.in_face_off ? min(p.total_faces, (long long)__ldg(p.in_face_off + p.B)) : (long long)p.B * p.F;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n_in; idx += (long long)gridDim.x * blockDim.x) {
    int b, f;
    if (p.in_face_off) {   // the last b with in_face_off[b] <= idx (empty meshes repeat an offset: the last one owns the face)
      int lo = 0, hi = p.B;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((long long)__ldg(p.in_face_off + mid) <= idx) lo = mid; else hi = mid;
      }
      b = lo;
      f = (int)(idx - __ldg(p.in_face_off + b));
      if (f >= p.F || idx >= __ldg(p.in_face_off + b + 1)) continue;   // (a mesh above the stated bound: flagged by the topology kernel)
    } else {
      b = (int)(idx / p.F);
      f = (int)(idx - (long long)b * p.F);
    }
    const size_t in_face = (size_t)idx;
    const size_t in_vert0 = p.in_vert_off ? (size_t)__ldg(p.in_vert_off + b) : (size_t)b * p.V;
    const long long v_b = p.in_vert_off ? (long long)(__ldg(p.in_vert_off + b + 1) - __ldg(p.in_vert_off + b)) : (long long)p.V;
    const int64_t* fptr = p.faces + in_face * NVF;
    long long vid[NVF];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < NVF; ++k) {
      vid[k] = fptr[k];
      ok &= vid[k] != p.pad_id && vid[k] >= 0 && vid[k] < v_b;
    }
    if (!ok) continue;
    const size_t row = (size_t)b * p.F + f;   // where face_sum looks the ids up: row_face[packed row] = b*F+f
    float pos[NVF][3];
#pragma unroll
    for (int k = 0; k < NVF; ++k) {
      const float* src = p.vertices + (in_vert0 + (size_t)vid[k]) * 3;
      pos[k][0] = src[0]; pos[k][1] = src[1]; pos[k][2] = src[2];
    }
    unsigned short bins[NSLOT];
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
    // table row ids (slot offset + bin), 2 per 32-bit store
    uint32_t* dst = reinterpret_cast<uint32_t*>(p.bins_out + (size_t)row * NSLOT);
    if (p.pair_layout) {   // [even slots | odd slots], ids = shared-memory line numbers (face_sum_pair_kernel)
      constexpr int H = NSLOT / 2;
#pragma unroll
      for (int j = 0; j < H; j += 2) {
        dst[j / 2] = (uint32_t)(p.pair_off[2 * j] + bins[2 * j]) | ((uint32_t)(p.pair_off[2 * j + 2] + bins[2 * j + 2]) << 16);
        dst[H / 2 + j / 2] = (uint32_t)(p.pair_off[2 * j + 1] + bins[2 * j + 1]) | ((uint32_t)(p.pair_off[2 * j + 3] + bins[2 * j + 3]) << 16);
      }
    } else {
#pragma unroll
      for (int s = 0; s < NSLOT; s += 2)
        dst[s / 2] = (uint32_t)(p.slot_off[s] + bins[s]) | ((uint32_t)(p.slot_off[s + 1] + bins[s + 1]) << 16);
    }
    if (p.face_coords_out) {
#pragma unroll
      for (int s = 0; s < NVF * 3; ++s) p.face_coords_out[in_face * (NVF * 3) + s] = bins[s];
    }
  }
  pdl_launch_dependents();
}

template <int NVF>
__global__ void __launch_bounds__(256, 4) face_sum_kernel(FeatParams p) {
  pdl_wait();
  constexpr int NSLOT = NVF * 3 + NVF + 1 + 3;
  constexpr int H = NSLOT / 2;   // table rows in flight per lane at a time
  const int lane = lane_id();
  const int n_rows = p.counts->n_faces;
  const int D = p.D, d4 = D >> 2;
  const int items = 2 * d4;     // float4 columns of a pair of rows
  const int n_warps = gridDim.x * (blockDim.x >> 5);
  for (int pair = blockIdx.x * (blockDim.x >> 5) + warp_id(); 2 * pair < n_rows; pair += n_warps) {
    const int r0 = 2 * pair;
    const bool two = r0 + 1 < n_rows;
    // lane s holds table row id s of row r0 (low half) and of row r0 + 1 (high half)
    unsigned ids = 0;
    if (lane < NSLOT) {   // (the bins are stored per input face slot: row_face maps the packed row to it)
      ids = p.bins_out[(size_t)p.row_face[r0] * NSLOT + lane];
      if (two) ids |= (unsigned)p.bins_out[(size_t)p.row_face[r0 + 1] * NSLOT + lane] << 16;
    }
    for (int item = lane; item < items; item += 32) {
      const int second = item >= d4;
      const int i4 = item - second * d4;
      float4 acc = __ldg(reinterpret_cast<const float4*>(p.bias) + i4);
      float4 t[H];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int s = 0; s < H; ++s) {
          const unsigned both = __shfl_sync(0xffffffffu, ids, half * H + s);
          const unsigned id = second ? (both >> 16) : (both & 0xffffu);
          t[s] = __ldg(reinterpret_cast<const float4*>(p.tables + (size_t)id * D) + i4);
        }
#pragma unroll
        for (int s = 0; s < H; ++s) { acc.x += t[s].x; acc.y += t[s].y; acc.z += t[s].z; acc.w += t[s].w; }
      }
      const long long row = r0 + second;
      if (!second || two) {
        if (p.x0) *reinterpret_cast<float4*>(p.x0 + (size_t)row * D + i4 * 4) = acc;
        if (p.x0_image) tc::image_store4(p.x0_image, D, row, i4 * 4, acc);
      }
    }
  }
  pdl_launch_dependents();
}

// The same sum with the tables staged in shared memory.  The folded tables (2048 rows x 64 floats = 512 KB for the default
// model) do not fit one SM, a slice of S of the D columns does: CTA (g, slice) copies columns [slice * S, +S) of every
// table row into shared memory once (rows_total x S floats, 128 KB) and sums them for face group g.  S / 4 consecutive
// threads take one face (a float4 column each), so a warp reads S * 4-byte pieces of 32 / (S / 4) table rows per slot from
// shared memory instead of L2, and writes whole 128-byte segments of the activation image.  Same slot order of the
// additions as face_sum_kernel -> the same bits.
template <int NVF, int S, int NT>
__global__ void __launch_bounds__(NT, 1) face_sum_smem_kernel(FeatParams p, int rows_total, int groups) {
  extern __shared__ float4 stab[];   // [rows_total][S / 4]
  constexpr int NSLOT = NVF * 3 + NVF + 1 + 3;
  constexpr int s4 = S >> 2, ROWS_PER_IT = NT / s4;
  const int D = p.D, nslices = D / S;
  const int slice = blockIdx.x % nslices, g = blockIdx.x / nslices;
  const int tid = threadIdx.x;
  const int c = tid % s4;
  // the tables are model constants: staged before the wait on the preceding kernels
  for (int i = tid; i < rows_total * s4; i += NT) {
    const int row = i / s4, cc = i % s4;
    stab[i] = __ldg(reinterpret_cast<const float4*>(p.tables + (size_t)row * D + slice * S) + cc);
  }
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + slice * S) + c);
  pdl_wait();
  __syncthreads();
  const int n_rows = p.counts->n_faces;
  const int per = (((n_rows + groups - 1) / groups) + 7) & ~7;   // whole 8-row groups of the image per CTA
  const int r_lo = g * per, r_hi = min(n_rows, r_lo + per);
  int row = r_lo + tid / s4;
  uint32_t w[NSLOT / 2], wn[NSLOT / 2];
  auto load_ids = [&](int r, uint32_t (&dst)[NSLOT / 2]) {
    const uint32_t* bp = reinterpret_cast<const uint32_t*>(p.bins_out + (size_t)__ldg(p.row_face + r) * NSLOT);
#pragma unroll
    for (int s2 = 0; s2 < NSLOT / 2; ++s2) dst[s2] = __ldg(bp + s2);
  };
  if (row < r_hi) load_ids(row, w);
  for (; row < r_hi; row += ROWS_PER_IT) {
    if (row + ROWS_PER_IT < r_hi) load_ids(row + ROWS_PER_IT, wn);   // the next row's ids travel while this row is summed
    float4 acc = b4;
#pragma unroll
    for (int s2 = 0; s2 < NSLOT / 2; ++s2) {
      const float4 t0 = stab[(w[s2] & 0xffffu) * s4 + c];
      const float4 t1 = stab[(w[s2] >> 16) * s4 + c];
      acc.x += t0.x; acc.y += t0.y; acc.z += t0.z; acc.w += t0.w;
      acc.x += t1.x; acc.y += t1.y; acc.z += t1.z; acc.w += t1.w;
    }
    const int col = slice * S + c * 4;
    if (p.x0) *reinterpret_cast<float4*>(p.x0 + (size_t)row * D + col) = acc;
    if (p.x0_image) tc::image_store4(p.x0_image, D, row, col, acc);
#pragma unroll
    for (int s2 = 0; s2 < NSLOT / 2; ++s2) w[s2] = wn[s2];
  }
  pdl_launch_dependents();
}

template <int NVF, int S>
int launch_face_sum_smem(const FeatParams& p, int rows_total, int max_rows, cudaStream_t stream) {
  const size_t smem = (size_t)rows_total * S * sizeof(float);
  const int nslices = p.D / S;
  const int groups = max(1, min(148 / nslices, ceil_div(max_rows, 64)));
  // MESHTOK_FACE_SUM_T=1024: 1 024 threads (64 registers) per CTA - one CTA per SM (the table slice is 128 KB), so the thread count IS
  // the occupancy; it changed nothing, the kernel runs at the pace of the shared-memory pipe (see face_sum_pair_kernel)
  static int nt = -1;
  if (nt < 0) { const char* e = getenv("MESHTOK_FACE_SUM_T"); nt = (e && atoi(e) == 1024) ? 1024 : 512; }   // (1 024 threads measured no faster: 57.4 vs 57.6 us at C2)
  if (nt == 1024) {
    MT_ENSURE_SMEM((face_sum_smem_kernel<NVF, S, 1024>), 200 * 1024);
    MT_LAUNCH_PDL((face_sum_smem_kernel<NVF, S, 1024>), groups * nslices, 1024, smem, stream, p, rows_total, groups);
  } else {
    MT_ENSURE_SMEM((face_sum_smem_kernel<NVF, S, 512>), 200 * 1024);
    MT_LAUNCH_PDL((face_sum_smem_kernel<NVF, S, 512>), groups * nslices, 512, smem, stream, p, rows_total, groups);
  }
  return MESHTOK_OK;
}

// Round 2: the same sum without shared-memory bank conflicts.  ncu on the kernel above: 1.4 M bank conflicts per launch - with 16-column
// slices a table row is 64 bytes, so the two rows a quarter warp reads per instruction (two faces x one slot) share their banks half of
// the time, and the kernel runs at the shared-memory pipe's pace (1 024 threads instead of 512 changed nothing: 57.4 vs 57.6 us).  Here
// the slots are split into the even and the odd ones and LINE i of the table (128 bytes) holds row i of the even slots in its first 64
// bytes and row i of the odd slots in its second 64.  The two faces of a quarter warp walk their slots in opposite order - lanes 0-3
// even, odd, even, ... and lanes 4-7 odd, even, odd, ... - so every instruction reads one even and one odd row: disjoint banks.  No
// select is needed for that: the ids are stored [even slots | odd slots] per face, a thread only swaps its two id pointers and its two
// table-half pointers by lane parity, keeps one accumulator per half (the even slots are summed in the same order whichever half a lane
// calls "first") and adds the two at the end (a + b = b + a bit for bit), then the bias.  Same instruction count as the kernel above
// plus four adds.  (A version with EIGHT lanes per face - one half each, joined by shuffles - was conflict free too but executed 2.6 x
// the instructions and took 65 us.)  The table slice is staged with cp.async, 16 copies per thread in flight.
template <int NVF, int NT>
__global__ void __launch_bounds__(NT, 1) face_sum_pair_kernel(FeatParams p, int lines, int groups) {
  extern __shared__ float4 stab[];   // [lines][8]: even row (4 float4) | odd row (4 float4)
  constexpr int NSLOT = NVF * 3 + NVF + 1 + 3, H = NSLOT / 2, HW = H / 2;   // slots, slots per half, id words per half
  constexpr int S = 16, FACES_PER_IT = NT / 4;
  const int D = p.D, nslices = D / S;
  const int slice = blockIdx.x % nslices, g = blockIdx.x / nslices;
  const int tid = threadIdx.x;
  const int c = tid & 3, par = (tid >> 2) & 1;
  // the tables are model constants: staged before the wait on the preceding kernels
  {
    const uint32_t sbase = tc::smem_u32(stab);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int n_s = (s + 1 < NSLOT ? p.slot_off[s + 1] : p.slot_off[NSLOT - 1] + p.n_normal) - p.slot_off[s];
      const float* src = p.tables + (size_t)p.slot_off[s] * D + slice * S;
      const uint32_t dst = sbase + (uint32_t)p.pair_off[s] * 128u + (uint32_t)(s & 1) * 64u;
      for (int i = tid; i < n_s * 4; i += NT) {
        const int r = i >> 2, cc = i & 3;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)r * 128u + (uint32_t)cc * 16u), "l"(src + (size_t)r * D + cc * 4) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + slice * S) + c);
  pdl_wait();
  const int n_rows = p.counts->n_faces;
  const int per = (((n_rows + groups - 1) / groups) + 7) & ~7;   // whole 8-row groups of the image per CTA (r_lo is even)
  const int r_lo = g * per, r_hi = min(n_rows, r_lo + per);
  int row = r_lo + (tid >> 2);
  // ids of the half this lane reads first / second: the current face's, the next one's and the one after.  The ids sit behind TWO
  // dependent global loads (row_face -> bins); ncu on the round-1 kernel, which fetched both one iteration ahead, had 40 % of its stall
  // samples waiting for them (an iteration is ~300 cycles of work).  Here row_face runs three faces ahead and the ids two.
  uint32_t wf[HW], ws[HW], nf[HW], ns[HW], mf[HW], ms[HW];
  auto load_ids = [&](int rf, uint32_t (&d0)[HW], uint32_t (&d1)[HW]) {
    const uint32_t* bp = reinterpret_cast<const uint32_t*>(p.bins_out + (size_t)rf * NSLOT);
#pragma unroll
    for (int j = 0; j < HW; ++j) { d0[j] = __ldg(bp + par * HW + j); d1[j] = __ldg(bp + (par ^ 1) * HW + j); }
  };
  auto rf_at = [&](int r) { return r < r_hi ? __ldg(p.row_face + r) : 0; };
  {
    const int rf0 = rf_at(row), rf1 = rf_at(row + FACES_PER_IT);
    if (row < r_hi) load_ids(rf0, wf, ws);
    if (row + FACES_PER_IT < r_hi) load_ids(rf1, nf, ns);
  }
  int rf2 = rf_at(row + 2 * FACES_PER_IT);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const float4* tf = stab + par * 4 + c;          // the half read first: even rows for lanes 0-3 of a quarter warp, odd rows for lanes 4-7
  const float4* ts = stab + (par ^ 1) * 4 + c;
  for (; row < r_hi; row += FACES_PER_IT) {
    if (row + 2 * FACES_PER_IT < r_hi) load_ids(rf2, mf, ms);
    rf2 = rf_at(row + 3 * FACES_PER_IT);
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
    for (int j = 0; j < HW; ++j) {
      MT_DCHECK((int)(wf[j] & 0xffffu) < lines && (int)(wf[j] >> 16) < lines && (int)(ws[j] & 0xffffu) < lines && (int)(ws[j] >> 16) < lines);
      const float4 f0 = tf[(wf[j] & 0xffffu) * 8];
      const float4 s0 = ts[(ws[j] & 0xffffu) * 8];
      const float4 f1 = tf[(wf[j] >> 16) * 8];
      const float4 s1 = ts[(ws[j] >> 16) * 8];
      add4(a0, f0);
      add4(a1, s0);
      add4(a0, f1);
      add4(a1, s1);
    }
    add4(a0, a1);
    add4(a0, b4);
    const float4 sum = a0;
    const int col = slice * S + c * 4;
    if (p.x0) *reinterpret_cast<float4*>(p.x0 + (size_t)row * D + col) = sum;
    if (p.x0_image) tc::image_store4(p.x0_image, D, row, col, sum);
#pragma unroll
    for (int j = 0; j < HW; ++j) { wf[j] = nf[j]; ws[j] = ns[j]; nf[j] = mf[j]; ns[j] = ms[j]; }
  }
  pdl_launch_dependents();
}

// Shared-memory lines of the pair layout and every slot's first line; 0 when the layout does not apply (slice not 16 columns, odd slot
// count per half, table too large for one SM) or MESHTOK_FACE_SUM_PAIR=0 asks for the round-1 kernel (A/B runs).
static int face_pair_plan(const FeatParams& p, int* pair_off) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MESHTOK_FACE_SUM_PAIR"); on = (e && atoi(e) == 0) ? 0 : 1; }
  const int nslot = p.nvf * 4 + 4;
  if (on == 0 || p.D % 16 != 0 || p.D / 16 > 148 || (nslot / 2) % 2 != 0) return 0;   // (only fields that both launchers see alike)
  int lines[2] = {0, 0};
  for (int s = 0; s < nslot; ++s) {
    const int n = (s + 1 < nslot ? p.slot_off[s + 1] : p.slot_off[nslot - 1] + p.n_normal) - p.slot_off[s];
    pair_off[s] = lines[s & 1];
    lines[s & 1] += n;
  }
  const int L = max(lines[0], lines[1]);
  if (L > 65535 || (size_t)L * 128 > (size_t)200 * 1024) return 0;
  return L;
}

constexpr int FACE_SUM_SMEM_MAX = 200 * 1024;

}  // namespace

static int face_embed_check(const FeatParams& p) {
  MT_CHECK_ARG(p.D % 4 == 0, "face_embed: dim_codebook must be a multiple of 4 (got %d)", p.D);
  MT_CHECK_ARG(p.bins_out != nullptr, "face_embed: no workspace for the bin ids");
  const int rows_total = p.n_coor * p.nvf * 3 + p.n_angle * p.nvf + p.n_area + p.n_normal * 3;
  MT_CHECK_ARG(rows_total <= 65535, "face_embed: at most 65535 folded table rows (got %d)", rows_total);
  return MESHTOK_OK;
}

// features + discretisation -> table row ids per input face; independent of the topology kernel (may run next to it)
int launch_face_bins(const FeatParams& p, long long max_in_faces, cudaStream_t stream) {
  int rc = face_embed_check(p);
  if (rc != MESHTOK_OK) return rc;
  if (max_in_faces == 0) return MESHTOK_OK;
  const int blocks1 = (int)max(1ll, min((max_in_faces + 127) / 128, 148ll * 16));
  FeatParams q = p;
  q.pair_layout = face_pair_plan(p, q.pair_off) > 0 ? 1 : 0;   // the id format follows the kernel launch_face_sum will take
  if (p.nvf == 3) MT_LAUNCH_PDL(face_bins_kernel<3>, blocks1, 128, 0, stream, q);
  else MT_LAUNCH_PDL(face_bins_kernel<4>, blocks1, 128, 0, stream, q);
  return MESHTOK_OK;
}

// x0[row] = bias + sum of the table rows of the packed row's face (after launch_face_bins AND the topology maps)
int launch_face_sum(const FeatParams& p, int max_rows, cudaStream_t stream) {
  int rc = face_embed_check(p);
  if (rc != MESHTOK_OK) return rc;
  if (max_rows == 0) return MESHTOK_OK;
  {
    FeatParams q = p;
    const int L = face_pair_plan(p, q.pair_off);
    if (L > 0) {
      q.pair_layout = 1;
      const int nslices = p.D / 16;
      const int groups = max(1, min(148 / nslices, ceil_div(max_rows, 64)));
      const size_t smem = (size_t)L * 128;
      if (p.nvf == 3) {
        MT_ENSURE_SMEM((face_sum_pair_kernel<3, 512>), 200 * 1024);
        MT_LAUNCH_PDL((face_sum_pair_kernel<3, 512>), groups * nslices, 512, smem, stream, q, L, groups);
      } else {
        MT_ENSURE_SMEM((face_sum_pair_kernel<4, 512>), 200 * 1024);
        MT_LAUNCH_PDL((face_sum_pair_kernel<4, 512>), groups * nslices, 512, smem, stream, q, L, groups);
      }
      return MESHTOK_OK;
    }
  }
  // table slices in shared memory when a slice of >= 8 columns fits; else table rows straight from L2
  const int rows_total = p.n_coor * p.nvf * 3 + p.n_angle * p.nvf + p.n_area + p.n_normal * 3;
  int S = 0;
  for (int cand = 32; cand >= 8 && S == 0; cand >>= 1)
    if (p.D % cand == 0 && (size_t)rows_total * cand * sizeof(float) <= (size_t)FACE_SUM_SMEM_MAX) S = cand;
  static int use_smem = -1;
  if (use_smem < 0) { const char* e = getenv("MESHTOK_FACE_SUM_L2"); use_smem = (e && atoi(e) == 1) ? 0 : 1; }
  if (S != 0 && use_smem && p.D / S <= 148) {
    if (p.nvf == 3) return S == 32 ? launch_face_sum_smem<3, 32>(p, rows_total, max_rows, stream)
                         : S == 16 ? launch_face_sum_smem<3, 16>(p, rows_total, max_rows, stream)
                                   : launch_face_sum_smem<3, 8>(p, rows_total, max_rows, stream);
    return S == 32 ? launch_face_sum_smem<4, 32>(p, rows_total, max_rows, stream)
         : S == 16 ? launch_face_sum_smem<4, 16>(p, rows_total, max_rows, stream)
                   : launch_face_sum_smem<4, 8>(p, rows_total, max_rows, stream);
  }
  const int blocks2 = max(1, min(ceil_div(ceil_div(max_rows, 2), 8), 148 * 4));
  if (p.nvf == 3) MT_LAUNCH_PDL(face_sum_kernel<3>, blocks2, 256, 0, stream, p);
  else MT_LAUNCH_PDL(face_sum_kernel<4>, blocks2, 256, 0, stream, p);
  return MESHTOK_OK;
}

}  // namespace meshtok
