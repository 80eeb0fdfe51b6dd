/*
 * meshtok.h - C ABI of libmeshtok.so: the B200 (sm_100a) mesh-tokenization hot path of
 * lucidrains/meshgpt-pytorch (MeshAutoencoder.tokenize / encode / quantize and
 * derive_face_edges_from_faces).
 *
 * The reference has no FFI for this path (it is pure Python on torch eager ops); every entry point
 * below names the reference code it replaces (paths relative to the reference repo root) and
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - All data pointers are DEVICE pointers borrowed for the duration of the call; the caller owns
 *    every byte, including the workspace (size from the matching *_workspace_bytes query).
 *  - Nothing allocates, nothing synchronises: work is enqueued on `stream` (a cudaStream_t) and
 *    the call returns immediately.  Data-dependent sizes (number of valid faces, edges, referenced
 *    vertices) stay on the device; grids are sized for the worst case and trimmed on the device.
 *  - Return value: 0 on success, a negative meshtok_status otherwise; meshtok_last_error_string()
 *    gives the thread-local reason.  Data errors that can only be seen on the device (vertex id out
 *    of range, edge buffer too small) set bits in the int32 status word at the start of the
 *    workspace (MESHTOK_WS_STATUS_*); the host wrapper reads it when it next synchronises.
 *  - Integer tensors are int64 at this surface (what torch.randint / the reference produce) and
 *    int32 inside; activations are fp32 row-major; there is no CPU implementation.
 */
#ifndef MESHTOK_H_
#define MESHTOK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MESHTOK_ABI_VERSION 3

typedef void* meshtok_stream_t; /* cudaStream_t */

enum meshtok_status {
  MESHTOK_OK = 0,
  MESHTOK_ERR_INVALID_ARGUMENT = -1,
  MESHTOK_ERR_UNSUPPORTED = -2, /* shape outside what the kernels were built for */
  MESHTOK_ERR_CUDA = -3,        /* a CUDA runtime call failed (see last_error_string) */
  MESHTOK_ERR_WORKSPACE = -4    /* workspace too small / misaligned */
};

/* bits of the device-side status word (int32 at workspace offset 0) */
#define MESHTOK_WS_STATUS_VERTEX_RANGE 1 /* a non-pad face references a vertex id outside [0, V) */
#define MESHTOK_WS_STATUS_EDGE_OVERFLOW 2 /* more edges than edge_capacity */
#define MESHTOK_WS_STATUS_EDGE_RANGE 4    /* a user-supplied edge references a face outside the valid faces of its own mesh */
#define MESHTOK_WS_STATUS_RAGGED_BOUNDS 8 /* ragged input: a mesh has more faces than F or more vertices than V, or decreasing offsets */

int meshtok_version(void);
const char* meshtok_last_error_string(void);

/* Number of kernels this library has launched since it was loaded (process wide). */
int64_t meshtok_kernel_launch_count(void);
/* Timing experiments (MESHTOK_LINEAR_TRACE=k): clock64 stamps of the k-th tcgen05 linear launch of a step, [2 CTAs][8 events][16 units]. */
int meshtok_debug_linear_trace(long long* out, int n);
/* Timing experiments (MESHTOK_GATHER_TRACE=k): per CTA of the k-th neighbour-aggregation launch of a step, 8 stamps (csrc/sage.cu). */
int meshtok_debug_gather_trace(long long* out, int n);

/* Optional device-side section timing: when enabled, the pipeline brackets each group of kernels with
 * cudaEvents on the caller's stream.  meshtok_profile_collect synchronises on those events, ADDS the elapsed
 * milliseconds and launch counts of each section into the caller's arrays (length >= meshtok_profile_sections())
 * and forgets them.  Section names: meshtok_profile_section_name(i). */
int meshtok_profile_enable(int on);
int meshtok_profile_sections(void);
const char* meshtok_profile_section_name(int section);
int meshtok_profile_collect(float* ms_per_section, int32_t* events_per_section, int n);

/* ------------------------------------------------------------------------------------------
 * Model description: dimensions + device pointers to the weights, laid out as the reference's
 * state_dict holds them (row-major fp32, nn.Linear weight = (out, in)) plus the tensors the host
 * wrapper derives once at load time (folded embedding tables, fp16 codebook image).
 * Replaces: the nn.Module attribute tree built in meshgpt_pytorch/meshgpt_pytorch.py:431-602.
 * ------------------------------------------------------------------------------------------ */
#define MESHTOK_MAX_SAGE_LAYERS 8

typedef struct meshtok_sage_layer {
  int32_t dim_in, dim_out;
  const float* lin_w;   /* (dim_in, dim_in)  SAGEConv.lin.weight   */
  const float* lin_b;   /* (dim_in)          SAGEConv.lin.bias     */
  const float* cat_w;   /* (dim_out, 2*dim_in) = [lin_l.weight | lin_r.weight] concatenated along K */
  const float* lin_l_b; /* (dim_out)         SAGEConv.lin_l.bias   */
  const void* lin_w_img; /* split-bf16 UMMA images of lin_w / cat_w (meshtok_linear_build_image); may be NULL */
  const void* cat_w_img; /* when the caller only uses gemm_simt */
} meshtok_sage_layer;

typedef struct meshtok_model {
  int32_t abi_version;   /* MESHTOK_ABI_VERSION */
  int32_t nvf;           /* 3 triangles, 4 quads            (meshgpt_pytorch.py:495-497) */
  int32_t dim_codebook;  /* 192                              (:452) */
  int32_t num_quantizers;/* 2                                (:453) */
  int32_t codebook_size; /* 16384                            (:454) */
  int32_t use_lfq;       /* 1 ResidualLFQ, 0 ResidualVQ      (:455) */
  int32_t num_sage_layers;        /* 5 = init_sage_conv + 4 encoders (:442-444, 543-560) */
  int32_t num_discrete_coors, num_discrete_angle, num_discrete_area, num_discrete_normals; /* :433-440 */
  float coor_lo, coor_hi;         /* coor_continuous_range (:434) */
  /* folded tables: slot s of the concatenated embedding (coords nvf*3, angles nvf, area 1, normals 3)
   * -> table[s][bin][dim_codebook] = Embedding.weight[bin] @ project_in.weight[:, cols(s)]^T.
   * slot_bins[s] = number of bins of slot s; slot_offset[s] = first table row of slot s. */
  int32_t num_slots;
  const float* folded_tables;     /* (sum_s bins_s, dim_codebook) */
  const float* project_in_bias;   /* (dim_codebook) */
  meshtok_sage_layer sage[MESHTOK_MAX_SAGE_LAYERS];
  const float* ln_gamma;          /* init_encoder_act_and_norm.1.weight (dims[0]) */
  const float* ln_beta;           /* init_encoder_act_and_norm.1.bias */
  const float* pdc_w;             /* project_dim_codebook.weight (nvf*dim_codebook, dims[-1]) */
  const float* pdc_b;             /* project_dim_codebook.bias */
  const void* pdc_w_img;          /* split-bf16 UMMA image of pdc_w */
  /* ResidualVQ (use_lfq == 0) */
  const float* codebook;          /* (codebook_size, dim_codebook) fp32 */
  const float* codebook_sqnorm;   /* (codebook_size) fp32  sum_d e^2 */
  const void* codebook_f16_tiles; /* UMMA-canonical fp16 image of codebook * codebook_image_scale,
                                     see meshtok_rvq_codebook_image_bytes / _build_codebook_image */
  float codebook_image_scale;     /* power of two bringing max |e_jd| into [0.5, 1) */
  float codebook_max_norm;        /* max_j |e_j|_2 (bounds the fp16 screening error) */
  /* ResidualLFQ (use_lfq == 1) */
  const float* lfq_in_w;          /* quantizer.project_in.weight (log2 K, dim_codebook) */
  const float* lfq_in_b;          /* quantizer.project_in.bias */
  const float* lfq_out_w;         /* quantizer.project_out.weight (dim_codebook, log2 K) (quantize() only) */
  const float* lfq_out_b;
} meshtok_model;

/* ------------------------------------------------------------------------------------------
 * K1  face adjacency.        Replaces derive_face_edges_from_faces, meshgpt_pytorch/data.py:297-340
 * One CTA per mesh: vertex->(face,slot) incidence by counting sort in shared memory, per-face
 * neighbour bitmaps, popc/prefix compaction.  Vertex ids must lie in [0, V); pad faces are faces
 * with ANY slot == pad_id.  threshold = 2 (1 when neighbor_if_share_one_vertex), include_self as in
 * the reference (pairs with exactly 3 shared slots dropped when 0).
 *
 * Two-phase dense API (the output length is data dependent):
 *   count: edge_counts[b] = number of directed edges of mesh b     (int32, device)
 *   fill : out (B, emax, 2) int64, rows sorted (i, then j), tail padded with pad_id
 * ------------------------------------------------------------------------------------------ */
int meshtok_topology_workspace_bytes(int B, int F, int nvf, int V, int E_user, size_t* bytes);

int meshtok_face_edges_count(const int64_t* faces, int B, int F, int nvf, int V, int64_t pad_id,
                             int threshold, int include_self, void* workspace, size_t workspace_bytes,
                             int32_t* edge_counts, meshtok_stream_t stream);

int meshtok_face_edges_fill(const int64_t* faces, int B, int F, int nvf, int V, int64_t pad_id,
                            int threshold, int include_self, void* workspace, size_t workspace_bytes,
                            int emax, int64_t* out_edges, meshtok_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Whole pipeline.  Replaces MeshAutoencoder.forward(return_codes=True) / tokenize,
 * meshgpt_pytorch/meshgpt_pytorch.py:949-1019 (encode :686-785, quantize :787-870).
 *
 *  vertices (B,V,3) fp32; faces (B,F,nvf) int64 padded with pad_id; face_edges (B,E,2) int64 or NULL
 *  (NULL => derived on the device, never materialised as a dense list).
 *  codes_out (B, F*nvf, Q) int64: codes, pad_id on padded tokens.
 *  Optional outputs (NULL to skip):
 *    encoded_out      (B,F,dims[-1]) fp32, zero on padded faces         (encode()'s first result)
 *    face_coords_out  (B,F,nvf*3) int64 discretised coordinates          (encode()'s second result)
 *    quantized_out    (B,F,nvf*dim_codebook) fp32                        (quantize()'s first result)
 *    histogram        (Q, codebook_size) int64, ACCUMULATED (+=) over non-pad tokens
 *  encoded_in (B,F,dims[-1]) fp32 or NULL: when given the encoder is skipped and quantize() runs on
 *  it (this is MeshAutoencoder.quantize as a stand-alone call).
 *  edge_capacity: number of int32 edge slots reserved in the workspace (derived or user edges).
 * ------------------------------------------------------------------------------------------ */
typedef struct meshtok_tokenize_args {
  const float* vertices;
  const int64_t* faces;
  const int64_t* face_edges;
  const float* encoded_in;
  int32_t B, F, V, E;
  int64_t pad_id;
  int64_t edge_capacity;
  int64_t* codes_out;
  float* encoded_out;
  int64_t* face_coords_out;
  float* quantized_out;
  int64_t* histogram;
  int32_t stop_after_encode; /* 1: encode() only (codes_out may be NULL) */
  int32_t rvq_exact;         /* 1: fp32 SIMT search instead of the tcgen05 screening + re-score */
  int32_t gemm_simt;         /* 1: fp32 SIMT linears instead of the tcgen05 split-bf16 GEMM */
  int32_t keep_taps;         /* 1: also store the fp32 form of every intermediate activation (meshtok_tokenize_taps);
                                0: activations that only feed tcgen05 linears exist as split-bf16 images only */
  /* Ragged input (CSR of meshes, SURVEY 8f rank 3: what custom_collate, data.py:347-369, would have padded with pad_id).
   * Both NULL: the padded layout above.  Both given (device, int32, B + 1 entries, non-decreasing, [0] == 0):
   *   faces (total_faces, nvf), mesh b owns rows face_offsets[b] .. face_offsets[b + 1]; vertex ids are mesh-local;
   *   vertices (vertex_offsets[B], 3), mesh b owns rows vertex_offsets[b] .. vertex_offsets[b + 1];
   *   F and V are upper bounds of the per-mesh face / vertex counts (they size the workspace; a mesh above them sets
   *   MESHTOK_WS_STATUS_RAGGED_BOUNDS and is tokenized as empty);
   *   total_faces = rows the faces / codes_out buffers hold (>= face_offsets[B], the live count the kernels use; a capacity is fine);
   *   codes_out (total_faces * nvf, Q): the tokens of the meshes back to back, no pad tokens between them (a face that
   *   contains pad_id still yields pad_id tokens); face_coords_out (total_faces, nvf * 3).
   *   face_edges (optional) (total_edges, 2) with mesh-local face indices, mesh b owns rows edge_offsets[b] .. edge_offsets[b + 1]
   *   (E > 0 is then only the "edges are given" flag; edge_capacity >= total_edges).
   *   encoded_in / encoded_out (total_faces, dims[-1]) and quantized_out (total_faces, nvf * dim_codebook): one row per flat face
   *   (zeros / the pad row on faces that contain pad_id); with encoded_in only the offsets, faces and V are used (vertices may be NULL). */
  const int32_t* face_offsets;
  const int32_t* vertex_offsets;
  int64_t total_faces;
  const int32_t* edge_offsets;
  int64_t total_edges;
  int64_t total_vertices;   /* ragged input: rows the vertices buffer holds (>= vertex_offsets[B]); 0 = unknown (B * V vertex rows are reserved) */
} meshtok_tokenize_args;

int meshtok_tokenize_workspace_bytes(const meshtok_model* model, int B, int F, int V, int E,
                                     int64_t edge_capacity, size_t* bytes);
/* The same for the ragged layout: the activation buffers are sized by total_faces face rows and total_vertices vertex rows (the values
 * the call will carry in args.total_faces / args.total_vertices) instead of B * F and B * V - the workspace grows with the SUM of the
 * mesh sizes, not with B x the largest mesh (C4: 116 k faces in 204 k padded slots).  0 = B * F / B * V. */
int meshtok_tokenize_workspace_bytes_ragged(const meshtok_model* model, int B, int F, int V, int E, int64_t edge_capacity,
                                            int64_t total_faces, int64_t total_vertices, size_t* bytes);

int meshtok_tokenize(const meshtok_model* model, const meshtok_tokenize_args* args, void* workspace,
                     size_t workspace_bytes, meshtok_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Host-to-host batch pipeline (csrc/pipeline.cu).  Replaces the per-batch host work of a CPU-side caller of
 * MeshAutoencoder.tokenize (meshgpt_pytorch.py:949-976 as called from MeshAutoencoderTrainer.tokenize, trainer.py:176-177, on
 * batches that custom_collate, data.py:347-369, built on the host): a ring of `depth` slots, each with its own stream, CUDA graph of
 * the whole tokenize pipeline and result buffer.  submit() enqueues H2D(vertices, faces) -> graph -> D2H(codes, status word) on the
 * next slot's stream and returns at once, handing back the result of the batch that used that slot `depth` submits ago (waiting for
 * it if need be); drain() waits for everything in flight, oldest first.  The caller owns all memory (passed per slot); the
 * pipeline owns only its streams, events and graph executables and is the one part of the library that synchronises.
 * vertices_host / faces_host should be pinned for the copies to be asynchronous.  Before _create, every slot's device staging
 * buffers must hold a VALID example batch (they are tokenized once eagerly and once under capture).  histogram (device, (Q, K)
 * int64, or NULL) is accumulated by every replay.  One thread at a time per pipeline.
 * ------------------------------------------------------------------------------------------ */
typedef struct meshtok_pipeline meshtok_pipeline;
typedef struct meshtok_pipeline_slot {
  float* vertices;         /* device (B, V, 3) staging */
  int64_t* faces;          /* device (B, F, nvf) staging */
  int64_t* codes;          /* device (B, F*nvf, Q) */
  void* workspace;         /* device, >= meshtok_tokenize_workspace_bytes(model, B, F, V, 0, edge_capacity), 256-byte aligned */
  size_t workspace_bytes;
  int64_t* codes_host;     /* pinned host (B, F*nvf, Q): where the slot's result lands */
  int32_t* status_host;    /* pinned host int32[4]: [0] = the batch's MESHTOK_WS_STATUS_* word */
} meshtok_pipeline_slot;
typedef struct meshtok_pipeline_result {
  int64_t ticket;          /* ordinal of the finished batch (0 for the first submit), -1: nothing finished */
  int32_t slot;
  int32_t status;          /* MESHTOK_WS_STATUS_* bits of that batch; non-zero: its codes are not valid */
  const int64_t* codes_host;
  int64_t tokens;          /* ragged pipelines: code rows (faces * nvf) of that batch in codes_host; 0 for padded pipelines */
} meshtok_pipeline_result;
int meshtok_pipeline_create(const meshtok_model* model, int B, int F, int V, int64_t pad_id, int64_t edge_capacity, int64_t* histogram,
                            const meshtok_pipeline_slot* slots, int depth, meshtok_pipeline** out);
int meshtok_pipeline_submit(meshtok_pipeline* p, const float* vertices_host, const int64_t* faces_host, meshtok_pipeline_result* finished);
int meshtok_pipeline_drain(meshtok_pipeline* p, meshtok_pipeline_result* results /* depth entries */, int* n);
/* The same ring for the RAGGED layout (a CSR of meshes, see meshtok_tokenize_args): every slot holds staging for up to max_meshes meshes,
 * cap_faces face rows and cap_vertices vertex rows, and ONE graph captured on those capacities that serves batches of any composition.
 * Before _create_ragged each slot's device staging (vertices, faces, offsets = [face_offsets (max_meshes + 1) | vertex_offsets
 * (max_meshes + 1)] int32) must hold a valid example batch.  submit_ragged copies only the live part of a batch (offsets_host is
 * pinned scratch of 2 * (max_meshes + 1) int32 per slot) and hands back flat codes (result.tokens rows of Q int64). */
typedef struct meshtok_pipeline_ragged_slot {
  float* vertices;         /* device (cap_vertices, 3) */
  int64_t* faces;          /* device (cap_faces, nvf), mesh-local vertex ids */
  int32_t* offsets;        /* device 2 * (max_meshes + 1) */
  int64_t* codes;          /* device (cap_faces * nvf, Q) */
  void* workspace;         /* device, >= meshtok_tokenize_workspace_bytes_ragged(model, max_meshes, F, V, 0, edge_capacity, cap_faces, cap_vertices) */
  size_t workspace_bytes;
  int64_t* codes_host;     /* pinned host (cap_faces * nvf, Q) */
  int32_t* status_host;    /* pinned host int32[4] */
  int32_t* offsets_host;   /* pinned host 2 * (max_meshes + 1) */
} meshtok_pipeline_ragged_slot;
int meshtok_pipeline_create_ragged(const meshtok_model* model, int max_meshes, int F, int V, int64_t cap_faces, int64_t cap_vertices, int64_t pad_id,
                                   int64_t edge_capacity, int64_t* histogram, const meshtok_pipeline_ragged_slot* slots, int depth,
                                   meshtok_pipeline** out);
int meshtok_pipeline_submit_ragged(meshtok_pipeline* p, const float* vertices_host, const int64_t* faces_host, const int32_t* face_offsets_host,
                                   const int32_t* vertex_offsets_host, int n_meshes, meshtok_pipeline_result* finished);
meshtok_stream_t meshtok_pipeline_stream(meshtok_pipeline* p, int slot);   /* the slot's cudaStream_t (to order caller events against it) */
int meshtok_pipeline_destroy(meshtok_pipeline* p);

/* Debug / parity taps: after meshtok_tokenize with args.keep_taps = 1, describes where the intermediate tensors
 * live inside the workspace (byte offsets) so tests can compare them stage by stage with the oracle. */
typedef struct meshtok_taps {
  int64_t status;          /* int32 status word */
  int64_t counts;          /* int32[4]: n_faces (valid), n_vertex_rows, n_edges, reserved */
  int64_t face_row;        /* int32 (B*F): packed row of each face or -1 */
  int64_t vert_row;        /* int32 (B*V): quantizer row of each vertex or -1 */
  int64_t indptr;          /* int32 (B*F+1): CSR by target face row */
  int64_t src;             /* int32 (edge_capacity): source face rows */
  int64_t x0;              /* fp32 (B*F, dim_codebook): project_in output, packed rows */
  int64_t layer_out[MESHTOK_MAX_SAGE_LAYERS]; /* fp32 (B*F, dim_out): output of each SAGE layer (layer 0: after act+norm) */
  int64_t averaged;        /* fp32 (B*V, dim_codebook): scatter-mean rows (vert_row order) */
  int64_t vertex_codes;    /* int32 (B*V, Q) */
  int64_t rvq_stats;       /* int32[4]: [1] = rows re-checked by the exact scan (summed over levels) in the last call */
} meshtok_taps;

int meshtok_tokenize_taps(const meshtok_model* model, int B, int F, int V, int E, int64_t edge_capacity,
                          meshtok_taps* taps);
int meshtok_tokenize_taps_ragged(const meshtok_model* model, int B, int F, int V, int E, int64_t edge_capacity, int64_t total_faces,
                                 int64_t total_vertices, meshtok_taps* taps);

/* ------------------------------------------------------------------------------------------
 * Per-stage entry points: the stages of meshtok_tokenize one by one, so that a single stage of the reference can be
 * swapped.  All of them take the SAME (model, args, workspace) as meshtok_tokenize - same workspace size and layout, see
 * meshtok_tokenize_taps for where the maps and counts live - and exchange fp32 ROWS: `*_rows` tensors are packed face rows
 * (n_faces valid faces of the batch back to back, in (mesh, face) order; capacity B*F rows) or packed vertex rows (n_vrows
 * referenced (mesh, vertex) pairs; capacity B*V rows).  Order of use: topology first (it fills the maps every other stage reads),
 * then any of the others.  args fields used: faces, vertices, face_edges/E, B, F, V, pad_id, edge_capacity, the ragged offsets,
 * gemm_simt, face_coords_out (features stage), codes_out / histogram (code gather).
 *   meshtok_stage_topology              derive_face_edges_from_faces (data.py:297-340) or the user's edge list (meshgpt_pytorch.py:752-754)
 *                                       -> face CSR by target + packed-row maps + vertex incidence CSR ("face_adjacency_csr" /
 *                                       "csr_from_edge_list")
 *   meshtok_stage_face_features_embed   meshgpt_pytorch.py:711-741 -> x0_rows (n_faces, dim_codebook)
 *   meshtok_stage_sage_project          SAGEConv.lin + ReLU of layer `layer` (call sites :764, :769) -> xp_rows (n_faces, dim_in)
 *   meshtok_stage_sage_aggregate_linear mean aggregation + lin_l/lin_r + F.normalize (+ SiLU + LayerNorm after layer 0, :766)
 *                                       -> out_rows (n_faces, dim_out)
 *   meshtok_stage_project_scatter_mean  project_dim_codebook + scatter_mean (:803-824) -> averaged_rows (n_vrows, dim_codebook);
 *                                       quantise them with meshtok_lfq_codes / meshtok_rvq_codes
 *   meshtok_stage_gather_codes          :856-866, :1018: vertex codes (n_vrows, Q) int32 -> args.codes_out (+ histogram)
 * ------------------------------------------------------------------------------------------ */
int meshtok_stage_topology(const meshtok_model* model, const meshtok_tokenize_args* args, void* workspace, size_t workspace_bytes, meshtok_stream_t stream);
int meshtok_stage_face_features_embed(const meshtok_model* model, const meshtok_tokenize_args* args, float* x0_rows, void* workspace,
                                      size_t workspace_bytes, meshtok_stream_t stream);
int meshtok_stage_sage_project(const meshtok_model* model, const meshtok_tokenize_args* args, int layer, const float* x_rows, float* xp_rows,
                               void* workspace, size_t workspace_bytes, meshtok_stream_t stream);
int meshtok_stage_sage_aggregate_linear(const meshtok_model* model, const meshtok_tokenize_args* args, int layer, const float* x_rows,
                                        const float* xp_rows, float* out_rows, void* workspace, size_t workspace_bytes, meshtok_stream_t stream);
int meshtok_stage_project_scatter_mean(const meshtok_model* model, const meshtok_tokenize_args* args, const float* enc_rows, float* averaged_rows,
                                       void* workspace, size_t workspace_bytes, meshtok_stream_t stream);
int meshtok_stage_gather_codes(const meshtok_model* model, const meshtok_tokenize_args* args, const int32_t* vertex_codes, void* workspace,
                               size_t workspace_bytes, meshtok_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Quantizer entry points on given rows (also used by the parity tests with oracle-injected inputs).
 * ------------------------------------------------------------------------------------------ */

/* K8b residual LFQ on given rows.  Replaces vector_quantize_pytorch.ResidualLFQ.forward (eval),
 * called at meshgpt_pytorch.py:842.  rows (R, dim) fp32 -> codes (R, Q) int32. */
int meshtok_lfq_codes(const meshtok_model* model, const float* rows, int R, int32_t* codes,
                      meshtok_stream_t stream);

/* K8a residual VQ on given rows.  Replaces vector_quantize_pytorch.ResidualVQ.forward (eval, shared
 * codebook), called at meshgpt_pytorch.py:842.  rows (R, dim) fp32 -> codes (R, Q) int32.
 * exact != 0: fp32 SIMT search; 0: tcgen05 fp16 screening + fp32 re-score. */
int meshtok_rvq_workspace_bytes(const meshtok_model* model, int R, size_t* bytes);
int meshtok_rvq_codes(const meshtok_model* model, const float* rows, int R, int exact, int32_t* codes,
                      float* residual_out /* (R, dim) or NULL */, void* workspace, size_t workspace_bytes,
                      meshtok_stream_t stream);

/* Size / builder of the fp16 codebook image consumed by the tcgen05 search: codebook * scale (scale a power of two) in
 * stage order, plus per 128-code tile a 4 KB K block carrying -kappa * |e_j|^2 as an fp16 pair (kappa = the power of two
 * that brings max_norm^2 into [0.5, 1)), which the CTA-pair kernel folds into the MMA chain.  codebook_sqnorm and max_norm
 * are the model's codebook_sqnorm / codebook_max_norm. */
int meshtok_rvq_codebook_image_bytes(int codebook_size, int dim, size_t* bytes);
int meshtok_rvq_build_codebook_image(const float* codebook, const float* codebook_sqnorm, int codebook_size, int dim,
                                     float scale, float max_norm, void* image, meshtok_stream_t stream);

/* Dense linear C = act(A W^T + b).  Replaces torch.nn.functional.linear at meshgpt_pytorch.py:741, 803 and
 * inside SAGEConv.  A (M, K) fp32; W (N, K) fp32; C (M, N) fp32.  relu: 0/1.
 * simt != 0: fp32 SIMT kernel reading A and W.
 * simt == 0: tcgen05 split-bf16 kernel reading two operand images: w_image, the (N*K*4 byte) hi/lo bf16 UMMA
 * image of W built once by meshtok_linear_build_image, and a_image, the hi/lo bf16 image of the activation rows
 * (inside the pipeline every producer kernel writes its output in this form; for a stand-alone call build it
 * with meshtok_rows_build_image). */
int meshtok_linear_image_bytes(int N, int K, size_t* bytes);
int meshtok_linear_build_image(const float* W, int N, int K, void* image, meshtok_stream_t stream);
int meshtok_rows_image_bytes(int M, int K, size_t* bytes);
int meshtok_rows_build_image(const float* A, int M, int K, void* image, meshtok_stream_t stream);
int meshtok_linear(const float* A, const void* a_image, const float* W, const void* w_image, const float* bias,
                   float* C, int M, int N, int K, int relu, int simt, meshtok_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Decode path: the stages of MeshAutoencoder.decode_from_codes_to_faces (meshgpt_pytorch.py:898-947) and
 * decode (:872-896) that are not linears.  Rows are (b, n) for every face slot of the batch, R = B * nf; `valid` (R bytes) is the
 * face mask.  A Conv1d(k) is meshtok_dec_im2col_image followed by meshtok_linear with the weight re-tiled to (C_out, k * C_in),
 * tap-major.  The host side (meshgpt_pytorch_b200/decoder.py) strings them together like the reference's decode().
 * ------------------------------------------------------------------------------------------ */
/* vector_quantize_pytorch get_output_from_indices as called at :917.  codes (n_tokens, Q) int64, negative = masked. */
int meshtok_dec_dequantize_rvq(const int64_t* codes, const float* codebook, int n_tokens, int Q, int D, float* out,
                               meshtok_stream_t stream);
int meshtok_dec_dequantize_lfq(const int64_t* codes, const float* w_out, const float* b_out, int n_tokens, int Q, int bits,
                               int D, float* out, meshtok_stream_t stream);
/* x (R, C) fp32 -> split-bf16 activation image (R, k * C): column t * C + c of row (b, n) = x[(b, n + t - k/2)][c], zero outside the
 * mesh and on masked faces (the masked_fill + zero padding of Block.forward, :341-349).  Size: meshtok_rows_image_bytes(R, k * C). */
int meshtok_dec_im2col_image(const float* x, const uint8_t* valid, int R, int nf, int C, int k, void* image,
                             meshtok_stream_t stream);
/* Block tail (:345-351): PixelNorm (F.normalize eps, * sqrt(C)) then SiLU on valid rows, zeros elsewhere; in place. */
int meshtok_dec_pixelnorm_silu(float* x, const uint8_t* valid, int R, int C, float eps, meshtok_stream_t stream);
/* init_decoder_conv tail (:621-627): SiLU then LayerNorm(gamma, beta, eps); in place. */
int meshtok_dec_silu_layernorm(float* x, int R, int C, const float* gamma, const float* beta, float eps,
                               meshtok_stream_t stream);
/* SqueezeExcite (:296-324): masked mean over the faces of every mesh -> Linear(w1, b1) -> SiLU -> Linear(w2, b2) -> Sigmoid.
 * mean_ws, gate: (B, C) fp32. */
int meshtok_dec_squeeze_excite_gate(const float* x, const uint8_t* valid, int B, int nf, int C, int inner, const float* w1,
                                    const float* b1, const float* w2, const float* b2, float* mean_ws, float* gate,
                                    meshtok_stream_t stream);
/* ResnetBlock tail (:378-379): out = h * gate[b] + res on valid rows, zeros elsewhere. */
int meshtok_dec_gate_residual(const float* h, const float* gate, const float* res, const uint8_t* valid, int R, int nf, int C,
                              float* out, meshtok_stream_t stream);
/* :928-942: arg-max over the nbins logits of every coordinate (first maximum) -> undiscretize; masked faces: NaN (bins: 0).
 * logits (R, ncoord * nbins); coords (R, ncoord) fp32; bins (R, ncoord) int64 or NULL. */
int meshtok_dec_argmax_coords(const float* logits, const uint8_t* valid, int R, int ncoord, int nbins, float lo, float hi,
                              float* coords, int64_t* bins, meshtok_stream_t stream);

/* ---- fused decode stages: no im2col.
 * PADDED ROW SPACE: activations are rows b * P + lead + n (P = nf + lead, lead = half width of the widest kernel), `lead` separator
 * rows in front of every mesh and after the last one: R = B * P + lead; `valid` (R bytes) is 0 on separators and padded faces.  Every
 * stage writes zero image rows for non-valid rows, so a tap that reaches over the end of a mesh reads zeros (Conv1d's zero padding and
 * the masked_fill of Block.forward, :341-349, at once).
 * HALO IMAGE: the split-bf16 image of an (R, C) activation in which consecutive rows are 16 bytes apart and every 128-row tile carries
 * the `halo` rows before and after it; the tcgen05 kernel reads tap t of a convolution from the same block, 16 t bytes further
 * (meshtok_conv1d).  Widths must be multiples of 32.  `ld` / `ldh` / `ldres` are row strides in floats (the linear that computes
 * block1.proj and residual_conv in one launch leaves both side by side in one (R, 2 C) buffer). ---- */
int meshtok_halo_image_bytes(int R, int C, int halo, size_t* bytes);
/* out (R, N) fp32 = Conv1d(C_in -> N, kernel k, zero padding k / 2) over the rows of a halo image (halo >= k / 2) + bias;
 * w_image = meshtok_linear_build_image of the (N, k * C_in) matrix weight.permute(0, 2, 1) (tap-major columns).  k = 1: a Linear. */
int meshtok_conv1d(const void* a_halo_image, const void* w_image, const float* bias, float* out, int R, int N, int C_in, int k,
                   int halo, meshtok_stream_t stream);
/* meshtok_conv1d with the Block tail (:345-351) in the epilogue: the first `width` output columns are written ONLY as the halo image
 * (halo out_halo) of silu(PixelNorm(row, eps)), zero rows where valid is 0 - no fp32 copy, no separate pixelnorm_silu launch.  N == width,
 * or N == 2 * width with residual_conv's rows in the second half: those go to out (R, N) fp32, columns width .. N (the first half of
 * `out` is not written).  Needs width <= 256, a multiple of 32, so that a row is one pass of the kernel; otherwise MESHTOK_ERR_UNSUPPORTED
 * (use meshtok_conv1d + meshtok_dec_pixelnorm_silu_halo). */
int meshtok_conv1d_pixelnorm_silu(const void* a_halo_image, const void* w_image, const float* bias, const uint8_t* valid, float* out, int R, int N,
                                  int C_in, int k, int halo, int width, float eps, int out_halo, void* out_image, meshtok_stream_t stream);
/* meshtok_conv1d + meshtok_dec_argmax_coords_padded in one launch (to_coor_logits :639 + argmax / undiscretize :935-944): w_image is the
 * image of the (N, k * C_in) weight with N = passes * 2 * nbins rows - coordinate c owns rows c * nbins .. (c + 1) * nbins, rows of coordinates
 * >= ncoord are zero padding - so that every pass of the tcgen05 kernel holds two whole coordinates; each epilogue thread takes the arg-max of
 * its row's nbins logits from TMEM (first maximum wins) and writes coords / bins (B, P - lead, ncoord) of the unpadded faces, NaN / 0 where
 * valid is 0.  The logits never reach global memory.  MESHTOK_ERR_UNSUPPORTED when a pass is not 2 * nbins wide (nbins = 128). */
int meshtok_conv1d_argmax_coords(const void* a_halo_image, const void* w_image, const float* bias, const uint8_t* valid, int R, int N, int C_in, int k,
                                 int halo, int B, int P, int lead, int ncoord, int nbins, float lo, float hi, float* coords, int64_t* bins,
                                 meshtok_stream_t stream);
/* number of face slices per mesh the SqueezeExcite partial sums use; slice_ws holds B * S * C floats + B * S int32 */
int meshtok_dec_se_slices(int B, int nf);
/* x (B * nf, C) fp32, UNPADDED rows -> halo image in the padded row space */
int meshtok_dec_rows_halo(const float* x, int ld, const uint8_t* valid, int R, int P, int lead, int C, int halo, void* image,
                          meshtok_stream_t stream);
/* get_output_from_indices (:917) + masked_fill (:924) -> the halo image init_decoder_conv.0 (:621) reads: codes (B * nf * nvf, Q),
 * unpadded -> image (R, nvf * D); the summed codebook vectors are never stored as fp32 */
int meshtok_dec_dequantize_rvq_halo(const int64_t* codes, const float* codebook, const uint8_t* valid, int R, int P, int lead, int nvf,
                                    int Q, int D, int halo, void* image, meshtok_stream_t stream);
/* init_decoder_conv tail (:621-627) -> out (R, C) fp32 (the first residual input) + halo image (NULL: none); non-valid rows 0 */
int meshtok_dec_silu_layernorm_halo(const float* x, const uint8_t* valid, int R, int P, int C, const float* gamma, const float* beta,
                                    float eps, float* out, int halo, void* image, meshtok_stream_t stream);
/* Block tail (:345-351) written only as the halo image of the next convolution; h is not modified */
int meshtok_dec_pixelnorm_silu_halo(const float* h, int ld, const uint8_t* valid, int R, int P, int C, float eps, int halo,
                                    void* image, meshtok_stream_t stream);
/* second Block tail + SqueezeExcite (:296-324) without storing the activation: scale (R) = sqrt(C) / max(||h[row]||, eps) on valid rows
 * and gate (B, C) from the masked mean of silu(h * scale); slice sums are combined in a fixed order (deterministic).  The two
 * SqueezeExcite weights are passed TRANSPOSED: w1t (C, inner) = excite.net.0.weight^T, w2t (inner, C) = excite.net.2.weight^T.
 * tickets: B int32, ZERO on entry and zero again when the kernel has finished (one buffer serves every call on a stream): the CTA
 * that finishes a mesh's last slice computes its gate, so the whole stage is one launch.  NULL: the gate runs as a second launch. */
int meshtok_dec_pixelnorm_squeeze_excite_gate(const float* h, int ld, const uint8_t* valid, int B, int P, int lead, int C, float eps,
                                              int inner, const float* w1t, const float* b1, const float* w2t, const float* b2,
                                              float* scale, float* slice_ws, int32_t* tickets, float* gate, meshtok_stream_t stream);
/* ResnetBlock tail (:378-379): out = silu(h * scale[row]) * gate[b] + res (scale NULL: h * gate[b] + res) -> out (R, C) fp32 and,
 * image != NULL, the halo image of the next convolution; non-valid rows 0 */
int meshtok_dec_gate_residual_halo(const float* h, int ldh, const float* scale, const float* gate, const float* res, int ldres,
                                   const uint8_t* valid, int R, int P, int C, float* out, int halo, void* image,
                                   meshtok_stream_t stream);
/* meshtok_dec_argmax_coords for logits in the padded row space: coords / bins are written for the UNPADDED faces (B * nf, ncoord) */
int meshtok_dec_argmax_coords_padded(const float* logits, const uint8_t* valid, int B, int P, int lead, int ncoord, int nbins,
                                     float lo, float hi, float* coords, int64_t* bins, meshtok_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Token-embedding front end of MeshTransformer.forward_on_codes, meshgpt_pytorch/meshgpt_pytorch.py:1526-1570
 * (the step that follows tokenization).  n = nvf * Q codes per face.
 * ------------------------------------------------------------------------------------------ */
/* Where the codes come from.  Either `codes` (B, L) int64 - the tensor tokenize returns, flattened '(b n q -> b (n q))' - or, skipping
 * that tensor, a meshtok_vertex_code_source: the tokenizer's per-vertex codes and the maps of its workspace (meshtok_tokenize_taps:
 * vertex_codes, face_row, vert_row) plus the faces tensor; then L = F * nvf * Q and token t = (face, slot, level) is looked up as
 * vertex_codes[vert_row[b, faces[b, face, slot]], level], pad_id on padded faces.  Exactly one of the two must be given.
 * eos_id >= 0: append_eos (:1506-1518) - the sequence is one token longer, eos_id at the first pad position of every row (lens_ws: B
 * int32 of scratch); eos_id < 0: none. */
typedef struct meshtok_vertex_code_source {
  const int64_t* faces;         /* (B, F, nvf) */
  const int32_t* face_row;      /* (B*F) packed row of each face or -1 */
  const int32_t* vert_row;      /* (B*V) quantizer row of each vertex or -1 */
  const int32_t* vertex_codes;  /* (n_vrows, Q) */
  int32_t F, nvf, V, Q;
} meshtok_vertex_code_source;
/* :1506-1518 as a tensor: out (B, L + 1) int64 (pad positions stay pad_id, the extra last column is 0 unless the eos landed there). */
int meshtok_frontend_append_eos(const int64_t* codes, const meshtok_vertex_code_source* vsrc, int B, int L, int64_t pad_id, int64_t eos_id,
                                int32_t* lens_ws, int64_t* out, meshtok_stream_t stream);
/* :1528-1547 + :1560: out (B, ceil(Le / n) * n, D) fp32 = ((token_embed[code, pad_id -> 0] + abs_pos_emb[t]) + level_embed[t % Q]) +
 * vertex_embed[(t / Q) % nvf], added in the reference's order (bit-exact); rows t >= Le (the incomplete last face) are zero; Le = L (+ 1
 * with an eos).  token_embed (n_embed, D) with n_embed = codebook_size + 1 (eos), abs_pos_emb (>= Le, D), level_embed (Q, D),
 * vertex_embed (nvf, D). */
int meshtok_frontend_fine_tokens(const int64_t* codes, const meshtok_vertex_code_source* vsrc, int B, int L, int D, int64_t pad_id, int64_t eos_id,
                                 int32_t* lens_ws, int Q, int nvf, int n_embed, const float* token_embed, const float* abs_pos_emb,
                                 const float* level_embed, const float* vertex_embed, float* out, meshtok_stream_t stream);
/* :1564-1570 (+ the sos rows of :1591-1594): out (B, n_sos + Le / n, D) fp32 = [sos (n_sos, D) ; LayerNorm(Linear(n * D -> D)(the n fine
 * tokens of a face)) for the complete faces], computed from tables folded at load time: folded_tables (n, n_embed, D) with
 * folded_tables[j] = token_embed @ W[:, j * D : (j + 1) * D]^T, and face_pos (>= Le / n, D) with
 * face_pos[f] = sum_j W_j (abs_pos_emb[n f + j] + level_embed[j % Q] + vertex_embed[j / Q]) + bias.
 * D must be a multiple of 128, at most 1024; n <= 32; n_sos = 0: no sos rows. */
int meshtok_frontend_face_tokens(const int64_t* codes, const meshtok_vertex_code_source* vsrc, int B, int L, int n, int D, int64_t pad_id, int64_t eos_id,
                                 int32_t* lens_ws, int n_embed, const float* folded_tables, const float* face_pos, const float* ln_gamma,
                                 const float* ln_beta, float eps, const float* sos, int n_sos, float* out, meshtok_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MESHTOK_H_ */
