// K1: face adjacency + batch topology (packed face rows, vertex incidence CSR, face CSR by target).
//
// Replaces derive_face_edges_from_faces (reference meshgpt_pytorch/data.py:297-340) and the
// offset / boolean-index packing of MeshAutoencoder.encode (meshgpt_pytorch.py:748-760) and the
// scatter index construction of quantize (meshgpt_pytorch.py:807-824).
//
// One CTA per mesh, everything in shared memory:
//   1. faces -> int32, validity (a face is padding if ANY slot == pad_id, data.py:317), rank of
//      each valid face among the valid faces (= its packed row inside the mesh);
//   2. counting sort of (face,slot) by vertex id -> incidence lists, each list ascending in
//      (face, slot) = the order in which the reference's scatter_add visits them on the CPU
//      (short lists: insertion sort by one thread; high-valence vertices: a per-warp bitmap sort);
//   3. per face i (ONE THREAD): a k-way merge over the sorted incidence lists of the face's distinct
//      vertices visits every candidate neighbour k exactly once, in ascending order (= the reference's
//      row-major order); for each candidate the exact slot counts n(i,k), n(k,i) of data.py:324-325 are
//      recomputed from the two faces.  A first pass counts the degrees, a second one emits the
//      neighbours at the prefix of the degrees - no bitmap, no sort, no warp reductions on this path.
// The relation is NOT symmetric for degenerate faces (SURVEY hard part 1), hence separate out- and in-degrees.
//
// The pipeline uses ONE launch, MODE_CSR (MODE_MAPS + MODE_EDGES on two streams is the A/B partner and the form quantize() without an
// encoder pass and user-supplied edges take): each CTA takes a mesh by ticket, counts its in-degrees, obtains the
// exclusive prefix of (valid faces, referenced vertices, edges) over the preceding meshes with a warp-wide decoupled
// look-back over per-mesh descriptors in global memory (aggregate / inclusive-prefix flags, as in single-pass scans),
// and then emits its slice of the batch-global CSR - no separate count launch, offsets kernel or host sync.  The
// dense (B, emax, 2) list of the stand-alone derive_face_edges_from_faces drop-in keeps the count / fill pair
// because its output length has to reach the host in between.
#include <limits.h>

#include "common.cuh"
#include "topology.cuh"

namespace meshtok {

namespace {

constexpr int MODE_COUNT = 0;
constexpr int MODE_FILL_DENSE = 1;
constexpr int MODE_CSR = 2;     // maps + edges in one launch
constexpr int MODE_MAPS = 3;    // packed face rows + vertex incidence CSR only (everything the feature kernel needs)
constexpr int MODE_EDGES = 4;   // face CSR by target only; runs after MODE_MAPS, typically on a second stream

// Walks the set bits of words[wlo..whi] in ascending order; fn(rank, bit_index).  Returns the count.
template <class Fn>
__device__ __forceinline__ int warp_emit_bits(const uint32_t* words, int wlo, int whi, Fn fn) {
  int base = 0;
  const int lane = lane_id();
  for (int w0 = wlo; w0 <= whi; w0 += 32) {
    const int w = w0 + lane;
    uint32_t bits = (w <= whi) ? words[w] : 0u;
    const int c = __popc(bits);
    const int incl = warp_inclusive_scan(c);
    int off = base + incl - c;
    while (bits) {
      const int bpos = __ffs(bits) - 1;
      bits &= bits - 1;
      fn(off++, w * 32 + bpos);
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
  return base;
}

template <int MODE, int NVF>   // NVF compile-time: the (face, slot) index arithmetic divides by it everywhere
__global__ void __launch_bounds__(1024, 1) topology_kernel(TopoParams p) {
  extern __shared__ int smem[];
  __shared__ int s_ticket;
  const int tid = threadIdx.x, nt = blockDim.x;
  constexpr bool kLookBack = MODE == MODE_CSR || MODE == MODE_MAPS || MODE == MODE_EDGES;
  constexpr bool kMaps = MODE == MODE_CSR || MODE == MODE_MAPS;
  constexpr bool kEdges = MODE == MODE_CSR || MODE == MODE_EDGES;
  if (kLookBack) {   // meshes are taken in ticket order so that every predecessor of the look-back is running or done
    if (tid == 0) s_ticket = atomicAdd(p.ticket, 1);
    __syncthreads();
  }
  const int b = kLookBack ? s_ticket : (int)blockIdx.x;
  const int F = p.F, V = p.V;
  constexpr int nvf = NVF;
  const int lane = lane_id(), warp = warp_id(), nwarps = nt >> 5;
  const int FN = F * nvf;

  int* fv = smem;                 // [FN]   vertex id of (face,slot) or -1
  int* frank = fv + FN;           // [F]    exclusive rank among valid faces
  int* fptr = frank + F;          // [F]    exclusive prefix of the degrees (fill modes)
  int* vstart = fptr + F;         // [V+1]  incidence list offsets
  int* vcur = vstart + (V + 1);   // [V+1]  fill cursors, later vertex ranks
  int* inc = vcur + (V + 1);      // [FN]   incidence entries face*nvf+slot
  int* scratch = inc + FN;        // [48]
  uint32_t* bits = reinterpret_cast<uint32_t*>(scratch + 48);  // [nwarps][p.words_per_warp]
  // CSR modes: the first p.ncache in-neighbours of every face, kept by the counting pass so that the emitting pass does not walk the
  // incidence lists a second time (the k-way merge is most of this kernel's instructions; ncu: issue bound at ~0.55 IPC per scheduler)
  unsigned short* ncache = reinterpret_cast<unsigned short*>(bits + (size_t)(blockDim.x >> 5) * p.words_per_warp);   // [F][p.ncache]
  const int Wf = (F + 31) >> 5;
  const int Ww = p.words_per_warp;

  // ---- 1. faces -------------------------------------------------------------------------------
  int bad = 0;
  const int64_t* fsrc = p.faces + (size_t)b * FN;
  int fn_b = FN, v_b = V;   // this mesh's own (face, slot) and vertex counts: the padded sizes unless the input is ragged
  if (p.in_face_off != nullptr) {
    const int f0 = p.in_face_off[b], f1 = p.in_face_off[b + 1], v0 = p.in_vert_off[b], v1 = p.in_vert_off[b + 1];
    fsrc = p.faces + (size_t)f0 * nvf;
    fn_b = (f1 - f0) * nvf;
    v_b = v1 - v0;
    if (f0 < 0 || f1 < f0 || f1 - f0 > F || v1 < v0 || v_b > V) {   // does not fit the workspace: an empty mesh + a status bit
      fn_b = 0;
      if (tid == 0) atomicOr(p.status, MESHTOK_WS_STATUS_RAGGED_BOUNDS);
    }
  }
  for (int idx = tid; idx < FN; idx += nt) {
    const long long v = idx < fn_b ? fsrc[idx] : p.pad_id;
    int iv;
    if (v == p.pad_id) iv = -1;
    else if (v < 0 || v >= v_b) { iv = -1; bad = 1; }
    else iv = (int)v;
    fv[idx] = iv;
  }
  if (bad) atomicOr(p.status, MESHTOK_WS_STATUS_VERTEX_RANGE);
  for (int i = tid; i < nwarps * Ww; i += nt) bits[i] = 0u;
  for (int v = tid; v <= V; v += nt) vstart[v] = 0;
  if (tid < 48) scratch[tid] = 0;
  __syncthreads();

  auto face_valid = [&](int f) {
    bool ok = true;
    for (int c = 0; c < nvf; ++c) ok &= fv[f * nvf + c] >= 0;
    return ok;
  };

  for (int f = tid; f < F; f += nt) frank[f] = face_valid(f) ? 1 : 0;
  __syncthreads();
  const int nvalid = block_exclusive_scan(frank, F, scratch);

  // ---- 2. incidence lists ---------------------------------------------------------------------
  for (int idx = tid; idx < FN; idx += nt) {
    const int f = idx / nvf;
    if (face_valid(f)) atomicAdd(&vstart[fv[idx]], 1);
  }
  __syncthreads();
  {
    int local = 0;
    for (int v = tid; v < V; v += nt) local += vstart[v] > 0;
    local = warp_sum_int(local);
    if (lane == 0 && local) atomicAdd(&scratch[40], local);
  }
  __syncthreads();
  const int nref = scratch[40];
  block_exclusive_scan(vstart, V + 1, scratch);
  for (int v = tid; v <= V; v += nt) vcur[v] = vstart[v];
  __syncthreads();
  for (int idx = tid; idx < FN; idx += nt) {
    const int f = idx / nvf;
    if (face_valid(f)) {
      const int pos = atomicAdd(&vcur[fv[idx]], 1);
      inc[pos] = idx;
    }
  }
  __syncthreads();
  // short lists: one thread each, insertion sort
  for (int v = tid; v < V; v += nt) {
    const int s = vstart[v], e = vstart[v + 1];
    if (e - s > 1 && e - s <= 32) {
      for (int i = s + 1; i < e; ++i) {
        const int key = inc[i];
        int j = i - 1;
        while (j >= s && inc[j] > key) { inc[j + 1] = inc[j]; --j; }
        inc[j + 1] = key;
      }
    }
  }
  __syncthreads();
  // long lists (high-valence vertices): one warp each, bitmap sort over the FN key space
  {
    uint32_t* wb = bits + warp * Ww;
    const int Wk = (FN + 31) >> 5;
    for (int v = warp; v < V; v += nwarps) {
      const int s = vstart[v], e = vstart[v + 1];
      if (e - s <= 32) continue;
      for (int i = s + lane; i < e; i += 32) {
        const int key = inc[i];
        atomicOr(&wb[key >> 5], 1u << (key & 31));
      }
      __syncwarp();
      warp_emit_bits(wb, 0, Wk - 1, [&](int r, int key) { inc[s + r] = key; });
      __syncwarp();
      for (int w = lane; w < Wk; w += 32) wb[w] = 0u;
      __syncwarp();
    }
  }
  __syncthreads();

  // ---- 3. adjacency ---------------------------------------------------------------------------
  // One thread per face.  The incidence lists are sorted, so a k-way merge over the lists of the face's distinct
  // vertices visits every candidate face k exactly once and in ascending order - no bitmap, no sort, no warp
  // reductions.  (The first version gave a warp to every face and collected the neighbours in per-warp bitmaps: ~6 of
  // 32 lanes had a list entry to look at and every face paid a dozen warp-wide reductions; 100 us for 64 x 800 faces.)
  // visit(k, n_ik, n_ki): n_ik = vertex slots of i found in k, n_ki = vertex slots of k found in i.
  auto for_each_candidate = [&](const int i, auto&& visit) {
    int mine[4], pos[4], end[4], head[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) mine[c] = c < nvf ? fv[i * nvf + c] : -2 - c;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      pos[c] = 0; end[c] = 0;
      if (c < nvf) {
        bool dup = false;
#pragma unroll
        for (int c0 = 0; c0 < 4; ++c0) dup |= c0 < c && mine[c0] == mine[c];
        if (!dup) { pos[c] = vstart[mine[c]]; end[c] = vstart[mine[c] + 1]; }   // a repeated vertex: same list again
      }
      head[c] = pos[c] < end[c] ? inc[pos[c]] / nvf : INT_MAX;
    }
    for (;;) {
      int k = INT_MAX;
#pragma unroll
      for (int c = 0; c < nvf; ++c) k = min(k, head[c]);
      if (k == INT_MAX) break;
#pragma unroll
      for (int c = 0; c < nvf; ++c) {
        while (head[c] == k) {   // a face with a repeated vertex sits in the list once per slot
          ++pos[c];
          head[c] = pos[c] < end[c] ? inc[pos[c]] / nvf : INT_MAX;
        }
      }
      int other[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) other[d] = d < nvf ? fv[k * nvf + d] : -7 - d;
      int n_ik = 0, n_ki = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        bool in_k = false, in_i = false;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          in_k |= mine[a] == other[d];
          in_i |= other[a] == mine[d];
        }
        n_ik += (a < nvf) && in_k;
        n_ki += (a < nvf) && in_i;
      }
      visit(k, n_ik, n_ki);
    }
  };
  // counting: degrees only (MODE_COUNT: both, to global; CSR modes: in-degree into fptr[]);  else: emit the
  // neighbours in ascending order at fptr[i] (dense list by source / CSR by target).
  // (CSR modes with p.lcsr: the mesh's block of the aggregation kernel's local CSR, see topology.cuh)
  uint8_t* lblk = (kLookBack && p.lcsr != nullptr) ? p.lcsr + (size_t)b * p.lcsr_stride : nullptr;
  unsigned short* lsrc = lblk ? reinterpret_cast<unsigned short*>(lblk + 16 + p.lcsr_ptr_bytes) : nullptr;
  auto adjacency = [&](const bool counting, const int face_off, const int edge_off, const int mesh_edges = 0) {
    const int thr = p.threshold;
    const bool keep3 = p.include_self != 0;
    const int nc = kLookBack ? p.ncache : 0;
    int total_out = 0;
    for (int i = tid; i < F; i += nt) {
      int dout = 0, din = 0, r = 0;
      if (face_valid(i)) {
        const int mybase = counting ? 0 : fptr[i];
        if (!counting && nc > 0) {   // the whole list is in the cache: copy it out
          const int deg = (i + 1 < F ? fptr[i + 1] : mesh_edges) - mybase;
          MT_DCHECK(deg >= 0 && mybase >= 0 && mybase + deg <= mesh_edges);
          if (deg <= nc) {
            for (int q = 0; q < deg; ++q) {
              const long long at = (long long)edge_off + mybase + q;
              MT_DCHECK(ncache[i * nc + q] < F);
              const int fr = frank[ncache[i * nc + q]];
              MT_DCHECK(fr >= 0 && fr < nvalid);
              if (at < p.edge_capacity) p.src[at] = face_off + fr;
              if (lsrc != nullptr && mybase + q < 5 * F) lsrc[mybase + q] = (unsigned short)(fr * 8);
            }
            continue;
          }
        }
        for_each_candidate(i, [&](const int k, const int n_ik, const int n_ki) {
          const bool eo = !kLookBack && n_ik >= thr && (keep3 || n_ik != 3);
          const bool ei = (MODE != MODE_FILL_DENSE) && n_ki >= thr && (keep3 || n_ki != 3);
          if (counting) {
            dout += eo;
            if (ei && din < nc) ncache[i * nc + din] = (unsigned short)k;
            din += ei;
          } else if (MODE == MODE_FILL_DENSE) {
            if (eo) {
              int64_t* dst = p.dense_out + ((size_t)b * p.emax + mybase + r) * 2;
              dst[0] = i;
              dst[1] = k;
              ++r;
            }
          } else if (ei) {
            const long long at = (long long)edge_off + mybase + r;
            if (at < p.edge_capacity) p.src[at] = face_off + frank[k];
            if (lsrc != nullptr && mybase + r < 5 * F) lsrc[mybase + r] = (unsigned short)(frank[k] * 8);
            ++r;
          }
        });
      }
      if (counting) {
        if (MODE == MODE_COUNT) {
          p.deg_out[(size_t)b * F + i] = dout;
          p.deg_in[(size_t)b * F + i] = din;
          total_out += dout;
        } else {
          fptr[i] = din;
        }
      }
    }
    if (counting && MODE == MODE_COUNT) {
      total_out = warp_sum_int(total_out);
      if (lane == 0 && total_out) atomicAdd(&scratch[41], total_out);
    }
  };

  if (MODE == MODE_COUNT) {
    if (p.do_adjacency) adjacency(true, 0, 0);
    __syncthreads();
    if (tid == 0) {
      int* mc = p.mesh_counts + 4 * b;
      mc[0] = nvalid;
      mc[1] = nref;
      mc[2] = scratch[41];
      mc[3] = 0;
      if (p.edge_counts_out) p.edge_counts_out[b] = scratch[41];
    }
    return;
  }
  if (MODE == MODE_FILL_DENSE) {
    // local prefix of the out-degrees computed by the COUNT launch
    for (int f = tid; f < F; f += nt) fptr[f] = p.deg_out[(size_t)b * F + f];
    __syncthreads();
    const int mesh_total = block_exclusive_scan(fptr, F, scratch);
    adjacency(false, 0, 0);
    int64_t* dst = p.dense_out + (size_t)b * p.emax * 2;
    for (int t = mesh_total * 2 + tid; t < p.emax * 2; t += nt) dst[t] = p.pad_id;
    return;
  }

  // ---- CSR modes: degrees -> look-back over the preceding meshes -> emit ------------------------
  int mesh_total = 0;
  if (kEdges && p.do_adjacency) {
    adjacency(true, 0, 0);
    __syncthreads();
    mesh_total = block_exclusive_scan(fptr, F, scratch);
  }
  if (warp == 0) {
    volatile int* agg = p.mesh_counts;   // [B][4]: nvalid, nref, nedges, ready flag
    volatile int* pre = p.mesh_prefix;   // [B][4]: inclusive prefixes of the same, ready flag
    if (lane == 0) {
      agg[4 * b + 0] = nvalid; agg[4 * b + 1] = nref; agg[4 * b + 2] = mesh_total;
      __threadfence();
      agg[4 * b + 3] = 1;
    }
    int ex0 = 0, ex1 = 0, ex2 = 0;
    for (int j = b - 1;; j -= 32) {          // 32 predecessors per step; lanes past mesh 0 act as a zero prefix
      const int idx = j - lane;
      bool is_prefix = true;
      int v0 = 0, v1 = 0, v2 = 0;
      if (idx >= 0) {
        for (;;) {
          if (pre[4 * idx + 3] != 0) { is_prefix = true; break; }
          if (agg[4 * idx + 3] != 0) { is_prefix = false; break; }
        }
        __threadfence();
        volatile int* src_d = is_prefix ? pre : agg;
        v0 = src_d[4 * idx + 0]; v1 = src_d[4 * idx + 1]; v2 = src_d[4 * idx + 2];
      }
      const unsigned m = __ballot_sync(0xffffffffu, is_prefix);
      const int stop = m ? __ffs(m) - 1 : 32;   // lanes < stop: aggregates of nearer meshes; lane stop: a full prefix
      if (lane > stop) { v0 = 0; v1 = 0; v2 = 0; }
      ex0 += warp_sum_int(v0); ex1 += warp_sum_int(v1); ex2 += warp_sum_int(v2);
      if (m) break;
    }
    if (lane == 0) {
      pre[4 * b + 0] = ex0 + nvalid; pre[4 * b + 1] = ex1 + nref; pre[4 * b + 2] = ex2 + mesh_total;
      __threadfence();
      pre[4 * b + 3] = 1;
      scratch[42] = ex0; scratch[43] = ex1; scratch[44] = ex2;
    }
  }
  __syncthreads();
  const int face_off = scratch[42], vert_off = scratch[43], edge_off = scratch[44];
  if (kEdges && p.do_adjacency) adjacency(false, face_off, edge_off, mesh_total);
  __syncthreads();

  // ---- 4. outputs: maps + vertex CSR + face CSR row pointers ------------------------------------
  const bool last_mesh = b == (int)gridDim.x - 1;   // the last mesh knows the batch totals
  if (kEdges && p.do_adjacency) {
    if (tid == 0) p.edge_off[b] = edge_off;
    for (int f = tid; f < F; f += nt)
      if (face_valid(f)) p.indptr[face_off + frank[f]] = edge_off + fptr[f];
    if (lblk != nullptr) {   // header and row pointers of the local CSR block (the source ids went out with the edges above)
      int* lptr = reinterpret_cast<int*>(lblk + 16);
      for (int f = tid; f < F; f += nt)
        if (face_valid(f)) { MT_DCHECK(frank[f] >= 0 && frank[f] < nvalid && 16 + 4 * (nvalid + 1) <= 16 + p.lcsr_ptr_bytes); lptr[frank[f]] = fptr[f]; }
      if (tid == 0) {
        lptr[nvalid] = mesh_total;
        const long long e1 = (long long)edge_off + mesh_total;
        const int n_e = (int)((e1 < p.edge_capacity ? e1 : p.edge_capacity) - edge_off);
        int* hdr = reinterpret_cast<int*>(lblk);
        hdr[0] = n_e;
        hdr[1] = (n_e <= 5 * F && e1 <= p.edge_capacity && (p.lcsr_zero_row + 1) * 8 <= 65536) ? 1 : 0;
        hdr[2] = edge_off;
        hdr[3] = 0;
      }
    }
    if ((long long)edge_off + mesh_total > p.edge_capacity && tid == 0) atomicOr(p.status, MESHTOK_WS_STATUS_EDGE_OVERFLOW);
    if (last_mesh && tid == 0) {
      const int nf = face_off + nvalid, ne = edge_off + mesh_total;
      p.edge_off[gridDim.x] = ne;
      p.counts->n_edges = ne;
      p.indptr[nf] = ne;
    }
  }
  if (!kMaps) return;
  if (tid == 0) { p.face_off[b] = face_off; p.vert_off[b] = vert_off; }
  for (int f = tid; f < F; f += nt) {
    const bool ok = face_valid(f);
    const int row = face_off + frank[f];
    p.face_row[(size_t)b * F + f] = ok ? row : -1;
    if (ok) p.row_face[row] = b * F + f;
  }
  for (int v = tid; v < V; v += nt) vcur[v] = (vstart[v + 1] > vstart[v]) ? 1 : 0;
  __syncthreads();
  block_exclusive_scan(vcur, V, scratch);
  for (int v = tid; v < V; v += nt) {
    const bool ref = vstart[v + 1] > vstart[v];
    const int vr = vert_off + vcur[v];
    p.vert_row[(size_t)b * V + v] = ref ? vr : -1;
    if (ref) p.vptr[vr] = face_off * nvf + vstart[v];
  }
  const int nent = vstart[V];
  for (int e = tid; e < nent; e += nt) {
    const int idx = inc[e];
    const int f = idx / nvf, c = idx - f * nvf;
    p.vinc[(size_t)face_off * nvf + e] = (face_off + frank[f]) * nvf + c;
  }
  if (last_mesh && tid == 0) {
    const int nf = face_off + nvalid, nv = vert_off + nref;
    p.face_off[gridDim.x] = nf; p.vert_off[gridDim.x] = nv;
    p.counts->n_faces = nf; p.counts->n_vrows = nv;
    p.vptr[nv] = nf * nvf;
    if (!p.do_adjacency) {   // user-supplied edges follow: keep the totals consistent until they are known.  (With do_adjacency the
                             // edges launch writes the total - possibly BEFORE this launch ends: the two may run side by side.)
      p.edge_off[gridDim.x] = 0;
    }
  }
}

// exclusive prefix sums of the per-mesh counts -> face_off / vert_off / edge_off, and the totals.
__global__ void __launch_bounds__(1024) mesh_offsets_kernel(const int* __restrict__ mesh_counts, int* __restrict__ offs,
                                                            int B, Counts* counts) {
  __shared__ int buf[4096];
  __shared__ int scratch[40];
  for (int q = 0; q < 3; ++q) {
    int running = 0;
    int* out = offs + (size_t)q * (B + 1);
    for (int base = 0; base < B; base += 4096) {
      const int n = min(4096, B - base);
      for (int i = threadIdx.x; i < n; i += blockDim.x) buf[i] = mesh_counts[4 * (base + i) + q];
      __syncthreads();
      const int total = block_exclusive_scan(buf, n, scratch);
      for (int i = threadIdx.x; i < n; i += blockDim.x) out[base + i] = running + buf[i];
      running += total;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      out[B] = running;
      if (q == 0) counts->n_faces = running;
      if (q == 1) counts->n_vrows = running;
      if (q == 2) counts->n_edges = running;
    }
    __syncthreads();
  }
}

// ---- user supplied face_edges -> CSR by target (meshgpt_pytorch.py:752-754) -----------------------
// edges (B,E,2): row0 = first column = source i, row1 = second column = target j, both offset by the
// number of valid faces of the preceding meshes; an edge is kept iff neither entry equals pad_id.
// Ragged layout (in_edge_off != null): edges (in_edge_off[B], 2), mesh b owns rows in_edge_off[b] .. in_edge_off[b + 1]; the mesh of an
// edge is the last b with in_edge_off[b] <= idx (bisection, B + 1 ints from L1).
__device__ __forceinline__ int edge_mesh(long long idx, int E, const int32_t* __restrict__ in_edge_off, int B) {
  if (in_edge_off == nullptr) return (int)(idx / E);
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(in_edge_off + mid) <= idx) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void user_edges_count_kernel(const int64_t* __restrict__ edges, int B, int E, const int32_t* __restrict__ in_edge_off, long long ne,
                                        int64_t pad_id, const int* __restrict__ face_off, int* __restrict__ ecount, int* status) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ne) return;
  const long long i = edges[2 * idx], j = edges[2 * idx + 1];
  if (i == pad_id || j == pad_id) return;
  const int b = edge_mesh(idx, E, in_edge_off, B);
  // both ends must be valid faces of the edge's OWN mesh.  (The reference adds the mesh's row offset and gathers whatever row that is,
  // a face of a neighbouring mesh included; here that is an error, reported through the status word, whichever aggregation kernel runs.)
  const long long nb = face_off[b + 1] - face_off[b];
  if (i < 0 || i >= nb || j < 0 || j >= nb) { atomicOr(status, MESHTOK_WS_STATUS_EDGE_RANGE); return; }
  const long long t = face_off[b] + j;
  atomicAdd(&ecount[t], 1);
}

__global__ void __launch_bounds__(256) user_edges_mesh_sum_kernel(const int* __restrict__ ecount, const int* __restrict__ face_off,
                                                                  int* __restrict__ mesh_counts) {
  __shared__ int acc;
  if (threadIdx.x == 0) acc = 0;
  __syncthreads();
  const int b = blockIdx.x;
  int local = 0;
  for (int r = face_off[b] + threadIdx.x; r < face_off[b + 1]; r += blockDim.x) local += ecount[r];
  local = warp_sum_int(local);
  if (lane_id() == 0 && local) atomicAdd(&acc, local);
  __syncthreads();
  if (threadIdx.x == 0) mesh_counts[4 * b + 2] = acc;
}

__global__ void __launch_bounds__(256) user_edges_indptr_kernel(const int* __restrict__ ecount, const int* __restrict__ face_off,
                                                                const int* __restrict__ edge_off, int* __restrict__ indptr,
                                                                int* __restrict__ cursor, int B) {
  __shared__ int buf[2048];
  __shared__ int scratch[40];
  const int b = blockIdx.x;
  int running = edge_off[b];
  const int r0 = face_off[b], r1 = face_off[b + 1];
  for (int base = r0; base < r1; base += 2048) {
    const int n = min(2048, r1 - base);
    for (int i = threadIdx.x; i < n; i += blockDim.x) buf[i] = ecount[base + i];
    __syncthreads();
    const int total = block_exclusive_scan(buf, n, scratch);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      indptr[base + i] = running + buf[i];
      cursor[base + i] = running + buf[i];
    }
    running += total;
    __syncthreads();
  }
  if (b == B - 1 && threadIdx.x == 0) indptr[face_off[B]] = edge_off[B];
}

__global__ void user_edges_fill_kernel(const int64_t* __restrict__ edges, int B, int E, const int32_t* __restrict__ in_edge_off, long long ne,
                                       int64_t pad_id, const int* __restrict__ face_off, int* __restrict__ cursor, int* __restrict__ etmp,
                                       long long capacity, int* status) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ne) return;
  const long long i = edges[2 * idx], j = edges[2 * idx + 1];
  if (i == pad_id || j == pad_id) return;
  const int b = edge_mesh(idx, E, in_edge_off, B);
  const long long nb = face_off[b + 1] - face_off[b];
  if (i < 0 || i >= nb || j < 0 || j >= nb) return;   // (flagged by the count kernel)
  const long long t = face_off[b] + j;
  const int pos = atomicAdd(&cursor[t], 1);
  if (pos < capacity) etmp[pos] = (int)idx;
  else atomicOr(status, MESHTOK_WS_STATUS_EDGE_OVERFLOW);
}

// per target row: order the incoming edges by their position in the user's list (the order in which
// the reference's scatter_add visits them on the CPU), then replace edge ids by source rows.
__global__ void user_edges_sort_kernel(const int64_t* __restrict__ edges, int B, int E, const int32_t* __restrict__ in_edge_off,
                                       const int* __restrict__ face_off, const int* __restrict__ indptr, const int* __restrict__ etmp, int* __restrict__ src,
                                       const Counts* counts, long long capacity) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= counts->n_faces) return;
  const long long s = indptr[row], e = min((long long)indptr[row + 1], capacity);
  for (long long i = s; i < e; ++i) {  // selection by rank: lists are short
    const int key = etmp[i];
    int rank = 0;
    for (long long j = s; j < e; ++j) rank += etmp[j] < key;
    const int b = edge_mesh(key, E, in_edge_off, B);
    src[s + rank] = face_off[b] + (int)edges[2 * (long long)key];
  }
}

size_t topo_smem_bytes(int F, int nvf, int V, int nthreads, int* words_per_warp) {
  const int Wf = (F + 31) / 32;
  const int Wk = (F * nvf + 31) / 32;
  const int Ww = (2 * Wf > Wk ? 2 * Wf : Wk);
  if (words_per_warp) *words_per_warp = Ww;
  size_t ints = (size_t)F * nvf * 2 + (size_t)F * 2 + (size_t)(V + 1) * 2 + 48 + (size_t)(nthreads / 32) * Ww;
  return ints * 4;
}

}  // namespace

int topo_threads(int F) { return F >= 384 ? 1024 : (F >= 96 ? 512 : 256); }

int topo_check_shape(int B, int F, int nvf, int V) {
  if (B <= 0 || F <= 0 || V <= 0 || (nvf != 3 && nvf != 4)) {
    set_error("topology: need B,F,V > 0 and nvf in {3,4} (got B=%d F=%d nvf=%d V=%d)", B, F, nvf, V);
    return MESHTOK_ERR_INVALID_ARGUMENT;
  }
  if ((long long)B * F >= (1ll << 30) || (long long)B * V >= (1ll << 30) || (long long)B * F * nvf >= (1ll << 31)) {
    set_error("topology: batch too large for int32 row ids (B*F=%lld, B*V=%lld)", (long long)B * F, (long long)B * V);
    return MESHTOK_ERR_UNSUPPORTED;
  }
  int ww;
  const size_t smem = topo_smem_bytes(F, nvf, V, topo_threads(F), &ww);
  if (smem > 220 * 1024) {
    set_error("topology: one mesh (F=%d, nvf=%d, V=%d) needs %zu B of shared memory, more than one SM has; "
              "meshes this large are not supported by the per-mesh kernel", F, nvf, V, smem);
    return MESHTOK_ERR_UNSUPPORTED;
  }
  return MESHTOK_OK;
}

template <int MODE, int NVF>
static int launch_topology_nvf(TopoParams p, cudaStream_t stream) {
  const int nt = topo_threads(p.F);
  int ww = 0;
  size_t smem = topo_smem_bytes(p.F, p.nvf, p.V, nt, &ww);
  p.words_per_warp = ww;
  p.ncache = 0;
  if ((MODE == MODE_CSR || MODE == MODE_EDGES) && p.do_adjacency && p.F <= 65535) {
    // neighbour cache of the counting pass (8 per face: the derived adjacency has at most nvf + 1 entries on a manifold mesh), where it fits
    static int cache_on = -1;   // MESHTOK_TOPO_CACHE=0: the emitting pass walks the incidence lists again, as before (A/B)
    if (cache_on < 0) { const char* e = getenv("MESHTOK_TOPO_CACHE"); cache_on = (e && atoi(e) == 0) ? 0 : 1; }
    const size_t extra = ((size_t)p.F * 8 * sizeof(unsigned short) + 15) & ~(size_t)15;
    if (cache_on && smem + extra <= 220 * 1024) { p.ncache = 8; smem += extra; }
  }
  if (smem > 48 * 1024) MT_ENSURE_SMEM((topology_kernel<MODE, NVF>), 220 * 1024);
  topology_kernel<MODE, NVF><<<p.B, nt, smem, stream>>>(p);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

template <int MODE>
static int launch_topology(const TopoParams& p, cudaStream_t stream) {
  return p.nvf == 3 ? launch_topology_nvf<MODE, 3>(p, stream) : launch_topology_nvf<MODE, 4>(p, stream);
}

int topo_count(const TopoParams& p, cudaStream_t stream) { return launch_topology<MODE_COUNT>(p, stream); }
int topo_fill_dense(const TopoParams& p, cudaStream_t stream) { return launch_topology<MODE_FILL_DENSE>(p, stream); }
int topo_csr(const TopoParams& p, cudaStream_t stream) { return launch_topology<MODE_CSR>(p, stream); }
int topo_maps(const TopoParams& p, cudaStream_t stream) { return launch_topology<MODE_MAPS>(p, stream); }
int topo_edges(const TopoParams& p, cudaStream_t stream) { return launch_topology<MODE_EDGES>(p, stream); }

int topo_offsets(const int* mesh_counts, int* offs, int B, Counts* counts, cudaStream_t stream) {
  mesh_offsets_kernel<<<1, 1024, 0, stream>>>(mesh_counts, offs, B, counts);
  MT_CHECK_LAUNCH();
  return MESHTOK_OK;
}

int topo_user_edges(const int64_t* edges, int B, int E, const int32_t* in_edge_off, long long total_edges, int64_t pad_id, int* mesh_counts, int* offs, Counts* counts,
                    int* ecount, int* cursor, int* etmp, int* indptr, int* src, long long capacity, int max_rows,
                    int* status, cudaStream_t stream) {
  const int* face_off = offs;
  const int* edge_off = offs + 2 * (size_t)(B + 1);
  MT_CHECK_CUDA(cudaMemsetAsync(ecount, 0, sizeof(int) * (size_t)max_rows, stream));
  const long long ne = in_edge_off ? total_edges : (long long)B * E;   // ragged layout: a flat list cut by in_edge_off
  if (ne > 0) {
    const int blocks = (int)((ne + 255) / 256);
    user_edges_count_kernel<<<blocks, 256, 0, stream>>>(edges, B, E, in_edge_off, ne, pad_id, face_off, ecount, status);
    MT_CHECK_LAUNCH();
  }
  user_edges_mesh_sum_kernel<<<B, 256, 0, stream>>>(ecount, face_off, mesh_counts);
  MT_CHECK_LAUNCH();
  mesh_offsets_kernel<<<1, 1024, 0, stream>>>(mesh_counts, offs, B, counts);
  MT_CHECK_LAUNCH();
  user_edges_indptr_kernel<<<B, 256, 0, stream>>>(ecount, face_off, edge_off, indptr, cursor, B);
  MT_CHECK_LAUNCH();
  if (ne > 0) {
    const int blocks = (int)((ne + 255) / 256);
    user_edges_fill_kernel<<<blocks, 256, 0, stream>>>(edges, B, E, in_edge_off, ne, pad_id, face_off, cursor, etmp, capacity, status);
    MT_CHECK_LAUNCH();
    user_edges_sort_kernel<<<ceil_div(max_rows, 256), 256, 0, stream>>>(edges, B, E, in_edge_off, face_off, indptr, etmp, src, counts, capacity);
    MT_CHECK_LAUNCH();
  }
  return MESHTOK_OK;
}

}  // namespace meshtok
