#pragma once
#include "common.cuh"

namespace meshtok {

struct TopoParams {
  const int64_t* faces;   // (B,F,nvf); ragged: (in_face_off[B], nvf)
  const int32_t* in_face_off;   // ragged input: [B+1] first face of every mesh in `faces` (null: padded layout)
  const int32_t* in_vert_off;   // ragged input: [B+1] first vertex of every mesh (only the per-mesh vertex COUNT is used here)
  int B, F, nvf, V;
  int64_t pad_id;
  int threshold;          // shared-slot threshold (2, or 1 for neighbor_if_share_one_vertex)
  int include_self;
  int do_adjacency;       // 0 when the caller supplies face_edges
  int words_per_warp;     // filled by the launcher
  int ncache;             // filled by the launcher: neighbours per face the counting pass keeps for the emitting pass (0: none)
  int* status;
  int* mesh_counts;       // [B][4] nvalid, nref, nedges, ready flag (MODE_CSR look-back descriptor)
  int* mesh_prefix;       // [B][4] inclusive prefixes of the same + ready flag   (MODE_CSR)
  int* ticket;            // mesh ticket counter                                  (MODE_CSR)
  Counts* counts;         // batch totals, written by the CTA of the last mesh    (MODE_CSR)
  int* deg_out;           // [B*F]
  int* deg_in;            // [B*F]
  int32_t* edge_counts_out;  // [B] or null
  // fill dense
  int emax;
  int64_t* dense_out;     // (B, emax, 2)
  // csr (single launch): exclusive offsets per mesh, written by the kernel
  int* face_off;          // [B+1]
  int* vert_off;          // [B+1]
  int* edge_off;          // [B+1]
  int* face_row;          // [B*F]
  int* row_face;          // [B*F]
  int* vert_row;          // [B*V]
  int* vptr;              // [B*V+1]
  int* vinc;              // [B*F*nvf]
  int* indptr;            // [B*F+1]
  int* src;               // [edge_capacity]
  long long edge_capacity;
  // CSR mode, optional: the aggregation kernel's per-mesh local CSR blocks (sage.cu, local_csr_kernel's format), written here while the
  // mesh's row pointers and neighbour lists are still in shared memory: {n_e, fast, e0, 0} | row pointers (F + 1) | source row * 8 (u16)
  uint8_t* lcsr;
  int lcsr_stride, lcsr_ptr_bytes, lcsr_zero_row;
};

int topo_threads(int F);
int topo_check_shape(int B, int F, int nvf, int V);
int topo_count(const TopoParams& p, cudaStream_t stream);
int topo_fill_dense(const TopoParams& p, cudaStream_t stream);
// single-launch batch topology; the caller zeroes mesh_counts, mesh_prefix and ticket first
int topo_csr(const TopoParams& p, cudaStream_t stream);
// the same in two launches: maps first (packed rows, vertex CSR, face / vertex totals), then the face CSR (edges), which
// needs its own zeroed mesh_counts / mesh_prefix / ticket block and may run on another stream once the maps are done
int topo_maps(const TopoParams& p, cudaStream_t stream);
int topo_edges(const TopoParams& p, cudaStream_t stream);
int topo_offsets(const int* mesh_counts, int* offs, int B, Counts* counts, cudaStream_t stream);
int topo_user_edges(const int64_t* edges, int B, int E, const int32_t* in_edge_off, long long total_edges, int64_t pad_id, int* mesh_counts, int* offs, Counts* counts,
                    int* ecount, int* cursor, int* etmp, int* indptr, int* src, long long capacity, int max_rows,
                    int* status, cudaStream_t stream);

}  // namespace meshtok
