// Launch wrappers of the pipeline kernels (host side declarations).
#pragma once
#include "common.cuh"

namespace meshtok {

// ---- K2 ------------------------------------------------------------------------------------------
struct FeatParams {
  const float* vertices;    // (B,V,3)
  const int64_t* faces;     // (B,F,nvf)
  const int* row_face;      // packed row -> b*F+f
  const int32_t* in_face_off;   // ragged input ([B+1] each, null: padded layout): faces (in_face_off[B], nvf), vertices (in_vert_off[B], 3),
  const int32_t* in_vert_off;   // face_coords_out (in_face_off[B], nvf*3)
  const Counts* counts;
  int64_t pad_id;
  long long total_faces;    // ragged input: rows the faces buffer holds (the live count is in_face_off[B])
  int B, F, V, nvf, D;
  float coor_lo, coor_den, angle_den, area_den, clip_lo, clip_hi;
  int n_coor, n_angle, n_area, n_normal;
  const float* tables;      // (sum bins, D)
  const float* bias;        // (D)
  int slot_off[20];
  float* x0;                // (rows, D) fp32 or null
  uint8_t* x0_image;        // split-bf16 activation image of x0 (tc_common.cuh) or null
  int64_t* face_coords_out; // (B,F,nvf*3) or null
  uint16_t* bins_out;       // (B*F, nvf*4+4) workspace: folded-table row id (slot offset + bin) of every slot of every input face b*F+f
  // Set by the launchers (features.cu), not by the caller: pair_layout = 1 -> the ids are written for face_sum_pair_kernel: per face
  // [ids of the even slots | ids of the odd slots], each id = pair_off[slot] + bin = the 128-byte LINE of the shared-memory table that holds
  // the even slot's row in its first half and the odd slot's row in its second half
  int pair_layout;
  int pair_off[20];
};
// features + discretisation -> table row ids per INPUT face: needs nothing from the topology kernel (runs next to it on the side stream)
int launch_face_bins(const FeatParams& p, long long max_in_faces, cudaStream_t stream);
// x0[row] = bias + sum of the table rows of the packed row's face; after launch_face_bins and the topology maps
int launch_face_sum(const FeatParams& p, int max_rows, cudaStream_t stream);

// ---- K4 ------------------------------------------------------------------------------------------
// agg[t] = mean over in-edges (s -> t) of x[s]; sum in ascending CSR order / max(count, 1).
// agg (fp32 rows) and agg_image (split-bf16 activation image) are both optional outputs.
// lcsr (optional): the per-mesh local CSR blocks of launch_local_csr - a CTA then takes its mesh's row pointers and source ids with ONE
// bulk copy instead of three dependent global round trips.
int launch_gather_mean(const float* x, int d, const int* indptr, const int* src, long long edge_capacity,
                       float* agg, void* agg_image, const Counts* counts, int max_rows, const int* face_off, int B, int F,
                       cudaStream_t stream, const void* lcsr = nullptr);
// Per mesh one block of local_csr_stride(F) bytes: {n_e, fast, e0, 0} | row pointers relative to the mesh's first edge (F + 1 ints) |
// source rows relative to the mesh's first row, pre-multiplied for the aggregation kernel's shared-memory layout (5 F u16).
// Runs once per batch after the face CSR exists (next to the topology kernels, off the critical path); all five layers read it.
size_t local_csr_stride(int F);
// layout constants of a block for a producer other than local_csr_kernel (the one-launch topology kernel): bytes of the row-pointer part,
// the local row an invalid source id points at
void local_csr_layout(int F, int* ptr_bytes, int* zero_row);
int launch_local_csr(const int* indptr, const int* src, long long edge_capacity, const int* face_off, int B, int F, void* lcsr, cudaStream_t stream);
int gather_mean_read_trace(long long* out, int n);   // timing experiments, see sage.cu
// in place: x <- x / max(|x|_2, 1e-12); when gamma != null also SiLU then LayerNorm(gamma, beta, eps 1e-5).
// image_out (optional): the same rows as a split-bf16 activation image for the next tcgen05 linear.
int launch_row_norm(float* x, int d, const float* gamma, const float* beta, void* image_out, const Counts* counts,
                    int max_rows, cudaStream_t stream);

// C[m, :] = act([A1 | A2][m, :] . W^T + bias), fp32 SIMT.  m < *m_dev (<= m_max).
int launch_linear_simt(const float* A1, int K1, const float* A2, int K2, const float* W, const float* bias, float* C,
                       int N, const int* m_dev, int m_max, bool relu, cudaStream_t stream);

// ---- K6/K7/K8b/K9 --------------------------------------------------------------------------------
// avg[vrow] = (sum over incidence entries e of y[e*D .. ]) / count      (scatter_mean, meshgpt_pytorch.py:243-257)
// prep (optional): also leave every averaged row ready as level 0's residual of the residual VQ (what launch_rvq_prep does)
struct RvqPrepOut {
  float* resid;        // (rows, D) copy of avg, or null
  float* rscale;       // (rows) power-of-two row scales
  uint8_t* row_image;  // fp16 row image for the CTA-pair search, or null
  float c_scale;
};
int launch_scatter_mean(const float* y, int D, const int* vptr, const int* vinc, float* avg, const Counts* counts,
                        int max_vrows, const RvqPrepOut* prep, cudaStream_t stream);
// residual LFQ codes of rows (R x D): codes (R x Q) int32; optional quantized_out (R x D) = project_out(sum q).
int launch_lfq(const float* rows, int D, const float* w_in, const float* b_in, int bits, int Q, const float* w_out,
               const float* b_out, int32_t* codes, float* quantized_out, const int* n_rows_dev, int max_rows,
               cudaStream_t stream);
// codes_out[b, f*nvf+c, q] = vcodes[vert_row[b, faces[b,f,c]], q] (pad_id on padded faces).
// gh (optional): also hist[q][code] += incidence count of every quantizer row (= non-pad output tokens carrying the code), in the
// same launch (the first CTAs of the grid take the vertex rows).
struct GatherHist {
  const int* vptr;
  const Counts* counts;
  int max_vrows;
  unsigned long long* hist;   // (Q, K)
};
int launch_gather_codes(const int64_t* faces, const int* face_row, const int* vert_row, const int32_t* vcodes, int B,
                        int F, int nvf, int V, int Q, int64_t pad_id, int64_t* codes_out, const GatherHist* gh, int K, cudaStream_t stream);
// quantized_out[b, f, c*D:(c+1)*D] = qrows[vert_row[b, faces[b,f,c]]]; padded faces get the pad-vertex row,
// which the reference leaves at zero for LFQ / at the (zero) input for RVQ.
int launch_gather_codes_ragged(const int64_t* faces, const int32_t* in_face_off, const int* face_row, const int* vert_row, const int32_t* vcodes,
                               long long total_faces, int B, int F, int nvf, int V, int Q, int64_t pad_id, int64_t* codes_out,
                               const GatherHist* gh, int K, cudaStream_t stream);
int launch_gather_quantized(const int64_t* faces, const int* face_row, const int* vert_row, const float* qrows, int B,
                            int F, int nvf, int V, int D, const float* pad_row, float* out, const int32_t* in_face_off, long long total_faces,
                            cudaStream_t stream);
// encoded_out <- packed rows, zeros on padded faces: (B*F, d) in the padded layout, (total_faces, d) in the ragged one (in_face_off != null)
int launch_unpack_rows(const float* packed, const int* face_row, int BF, int d, float* out, const int32_t* in_face_off, int B, int F,
                       long long total_faces, cudaStream_t stream);
// packed rows <- (B*F, d) padded layout / (total_faces, d) ragged layout.
int launch_pack_rows(const float* padded, const int* row_face, int d, float* packed, const Counts* counts, int max_rows,
                     const int32_t* in_face_off, int F, cudaStream_t stream);

// ---- K8a -----------------------------------------------------------------------------------------
// exact fp32 nearest-code search: best[r] = argmin_j sqrt(max(|x|^2 + |e_j|^2 - 2 x.e_j, 0)), lowest index on ties.
// row_list == null: rows 0..*n_rows_dev-1; else rows row_list[0..*n_list_dev-1].
// packed[row] = (distance bits << 32 | index), merged with a 64-bit atomicMin.  Full scan: initialised here; list mode:
// merged into the values the re-score kernel left (its candidate is a valid code, so the minimum is the exact answer).
int launch_rvq_exact(const float* rows, int D, const float* codebook, const float* e2, int K, const int* row_list,
                     const int* n_dev, int max_rows, unsigned long long* packed, cudaStream_t stream);
// r <- r - e[idx]; codes[r*Q + level] = idx; rscale[r] = power of two that brings max|r_d| into [0.5, 1).
// row_image (optional): fp16(r * rscale) in the layout the CTA-pair search bulk-copies, with the per-row constant
// rscale * c_scale of the folded |e|^2 column (c_scale = codebook scale / (2 kappa), rvq_screen.cuh); rscale is then capped
// so that this constant stays a finite fp16 power of two.
int launch_rvq_update(float* rows, int D, const float* codebook, int K, const unsigned long long* packed, int32_t* codes, int Q, int level,
                      float* rscale, void* row_image, float c_scale, const int* n_rows_dev, int max_rows, cudaStream_t stream);
int launch_rvq_codes_only(const unsigned long long* packed, int32_t* codes, int Q, int level, int K, const int* n_rows_dev, int max_rows,
                          cudaStream_t stream);
// resid <- rows_in (copy), rscale and row_image as above.
int launch_rvq_prep(const float* rows_in, int D, float* resid, float* rscale, void* row_image, float c_scale, const int* n_rows_dev,
                    int max_rows, cudaStream_t stream);

// tcgen05 screening search (rvq_tc.cu).  Produces per (row, split) the two best fp16-screened scores + indices.
struct RvqTcPlan {
  int m_tiles, n_splits, tiles_per_split, units, grid;
  size_t partial_bytes;
};
int rvq_tc_plan(int max_rows, int K, int D, RvqTcPlan* plan);
int launch_rvq_tc_search(const float* rows, const float* rscale, int D, const void* codebook_tiles, float escale, const float* e2,
                         int K, const int* n_rows_dev, int max_rows, const RvqTcPlan& plan, float4* partial, cudaStream_t stream);
// Where the screening kernels leave the candidates of a row: mode 1 (rvq_tc.cu): n_splits entries at
// partial[row * n_splits ..]; mode 2 (rvq_tc2.cu): `parts` entries (column parts of the 256-code tile) per cluster whose
// step range touched the row's 256-row tile, at partial[row * stride + parts * (cluster - first cluster of the tile) + part].
struct RvqPartialLayout {
  int mode, stride, T, n_clusters, w_min, parts;
};
// re-score: exact fp32 distance of the screened candidates with the reference formula -> the row's code; a row whose candidate
// list may be incomplete is scanned exactly by its CTA in the same launch (stats[1] counts them); then codes[row, level] = idx and,
// with `update`, r <- r - e[idx], rscale[row] and (row_image != null) the fp16 row image of the next level's search.
// c_scale: 0 for the single-CTA screening kernel, the folded-|e|^2 row constant factor for the CTA-pair kernel
int launch_rvq_rescore(float* rows, float* rscale, int D, const float* codebook, const float* e2, int K, float max_norm,
                       float escale, float c_scale, const float4* partial, const RvqPartialLayout& layout, const int* n_rows_dev, int max_rows,
                       int32_t* codes, int Q, int level, bool update, void* row_image, int* stats, cudaStream_t stream);

// CTA-pair (cta_group::2) screening search (rvq_tc2.cu)
struct RvqTc2Plan {
  int T, n_clusters, max_slots;
  size_t partial_bytes, row_image_bytes;
};
int rvq_tc2_plan(int max_rows, int K, int D, RvqTc2Plan* plan);
RvqPartialLayout rvq_tc2_partial_layout(const RvqTc2Plan& plan, int max_rows);
int launch_rvq_tc2_search(const void* row_image, const float* rscale, int D, const void* codebook_tiles, float escale,
                          const int* n_rows_dev, int max_rows, const RvqTc2Plan& plan, float4* partial, cudaStream_t stream);
size_t rvq_codebook_image_bytes(int K, int D);
int launch_rvq_build_image(const float* codebook, const float* e2, int K, int D, float escale, float max_norm, void* image,
                           cudaStream_t stream);

// tcgen05 split-bf16 linear (linear_tc.cu); w_image from launch_linear_tc_build_image
int linear_tc_plan(int N, int K1, int K2, int* passes, int* NT);
size_t linear_tc_image_bytes(int N, int K);
int launch_linear_tc_build_image(const float* W, int N, int K, void* image, cudaStream_t stream);
// a1_image / a2_image: split-bf16 activation images (tc_common.cuh) written by the producing kernels.
// epi (optional): fused row epilogue - 1: x / max(|x|_2, 1e-12); 2: that + SiLU + LayerNorm(gamma, beta, eps 1e-5);
// the result goes to C (fp32 rows, may be null) and / or out_image (activation image with K = N).
// norm_mode 0 (plain epilogue) with extras, for output rows too wide for the fused normalisation: out_image = image of
// the UN-normalised rows, ssq_out (M, 2 * passes) = parts of their squared norms; the consumer passes them as row_ssq
// and gets C = acc / max(|x_row|, 1e-12) + bias, which equals the linear of the normalised rows.
struct LinearRowEpilogue {
  int norm_mode;
  const float* gamma;
  const float* beta;
  void* out_image;
  float* ssq_out;
  const float* row_ssq;
  int row_ssq_parts;
};
bool linear_tc_b2b_ok(int N, int norm_mode);   // layer output linear + the next layer's `lin` in one launch
// Neighbour aggregation fused into the output linear (SAGEConv mean aggregation, call sites meshgpt_pytorch.py:764, :769): the A1 operand
// (the K1 "agg" columns) is not read from an image but produced inside the launch by gather warps - mean over the in-neighbours
// src[indptr[r] .. indptr[r + 1]) of the fp32 rows x (M, K1), neighbours added in list order, sum * (1 / count) - written as the
// split-bf16 shared-memory tile the MMA reads.  The aggregate never goes to global memory.
struct LinearGather {
  const float* x;        // (M, K1) fp32 rows, NOT written by this launch
  const int* indptr;     // (M + 1) face CSR by target
  const int* src;        // source rows
  long long cap;         // edges stored in src (lists are cut there)
};
bool linear_tc_gather_ok(int K1, int N, int norm_mode);   // the fused aggregation is available for this layer shape (MESHTOK_GATHER_FUSED=0: never)
int launch_linear_tc_b2b(const void* a1_image, int K1, const void* a2_image, int K2, const void* w_image, const float* bias, float* C,
                         int N, const int* m_dev, int m_max, const LinearRowEpilogue* epi, const void* w2_image, const float* bias2,
                         float* C2, cudaStream_t stream, const LinearGather* gather = nullptr);
int linear_tc_read_trace(long long* out, int n);   // timing experiments, see linear_tc.cu
int linear_tc_ssq_parts(int N);   // (column parts per pass) * passes of an N-wide tcgen05 linear, at most 16
int linear_tc_fused_norm_ok(int N, int norm_mode);   // MESHTOK_OK when the fused epilogue covers this width
int launch_linear_tc(const void* a1_image, int K1, const void* a2_image, int K2, const void* w_image, const float* bias, float* C,
                     int N, const int* m_dev, int m_max, bool relu, const LinearRowEpilogue* epi, cudaStream_t stream);
// Block tail fused into the convolution's epilogue (decode path): the first `width` output columns -> silu(PixelNorm(row)) as a halo image
struct ConvBlockTail {
  int width;
  const uint8_t* valid;   // (rows) 0: zero image row
  void* out_image;        // halo image (rows, width) with halo out_halo
  int out_halo;
  float eps;
};
int launch_conv_tc(const void* a_halo_image, int C_in, int taps, int halo, const void* w_image, const float* bias, float* C, int N, int m_max,
                   cudaStream_t stream, const ConvBlockTail* tail = nullptr);   // Conv1d over the rows of a halo image, see linear_tc.cu
// to_coor_logits + arg-max in the convolution's epilogue (decode path)
struct ConvArgmax {
  const uint8_t* valid;   // (rows of the padded row space) 0: NaN coordinate, bin 0
  int B, P, lead;         // rows = B * P + lead, the faces of mesh b are rows b * P + lead .. (b + 1) * P
  int ncoord, nbins;
  float lo, hi;
  float* coords;          // (B, P - lead, ncoord)
  int64_t* bins;          // same shape or null
};
int launch_conv_argmax_tc(const void* a_halo_image, int C_in, int taps, int halo, const void* w_image, const float* bias, int N, int m_max,
                          const ConvArgmax& am, cudaStream_t stream);
size_t rows_image_bytes(long long rows, int K);
int launch_rows_to_image(const float* A, int K, const int* m_dev, int m_max, void* image, cudaStream_t stream);

}  // namespace meshtok
