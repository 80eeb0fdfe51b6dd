// extern "C" surface of libmeshtok.so (see include/meshtok.h) and the per-batch pipeline driver.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.cuh"
#include "rvq_screen.cuh"
#include "topology.cuh"

namespace meshtok {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

namespace {

// Bump allocator over the caller's workspace; with base == nullptr it only measures.
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  template <class T>
  T* take(size_t n, int64_t* tap = nullptr) {
    off = align_up(off, 256);
    if (tap) *tap = (int64_t)off;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct Workspace {
  int* status;
  Counts* counts;
  int* mesh_counts;
  int* offs;
  int *deg_out, *deg_in, *face_row, *row_face, *vert_row, *vptr, *vinc, *indptr, *src;
  uint8_t* lcsr;     // per-mesh local CSR blocks for the aggregation kernel (launch_local_csr)
  int *ecount, *cursor, *etmp;
  float* x0;
  float *xp, *agg;
  uint8_t *img_x[2], *img_agg;   // split-bf16 activation images (tcgen05 linears; layer input ping-pong + aggregate);
                                 // null when a width is not a multiple of 32
  float* layer_out[MESHTOK_MAX_SAGE_LAYERS];
  float* y;
  float* avg;
  float* resid;
  float* rscale;
  float* qrows;
  float* pad_qrow;
  float* zero_row;
  int32_t* vcodes;
  int32_t* best;
  unsigned long long* packed;
  int* rvq_stats;    // [4] inside the status block: [1] = rows of this call that needed the exact codebook scan
  float4* partial;
  uint8_t* rimg;     // fp16 image of the scaled residual rows (CTA-pair RVQ search)
  float* ssq;        // (B*F, 16): parts of the squared row norms of the last SAGE layer (deferred normalisation)
  uint16_t* bins;    // (B*F, nvf*4+4): folded-table row ids of every input face slot
  int rows_cap, vrows_cap;   // packed face rows / quantizer rows the activation buffers hold (<= B*F, B*V)
  size_t total;
};

// Per host thread: a non-blocking side stream and the two events that fork it from / join it to the caller's stream.
// Created on first use (the only allocation this library ever makes) and kept for the life of the thread.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  cudaEvent_t begin = nullptr, bins_done = nullptr;   // the feature-bin launch that runs next to the topology maps
};
SideStream* side_stream() {
  static thread_local SideStream ss;
  static thread_local int device = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  if (ss.stream != nullptr && dev == device) return &ss;
  if (ss.stream != nullptr) {   // the thread moved to another device: events and streams are per device
    cudaStreamDestroy(ss.stream); cudaEventDestroy(ss.fork); cudaEventDestroy(ss.join);
    cudaEventDestroy(ss.begin); cudaEventDestroy(ss.bins_done);
    ss = SideStream();
  }
  if (cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&ss.begin, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&ss.bins_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  device = dev;
  return &ss;
}

// tcgen05 RVQ screening kernel: 2 = CTA-pair (cta_group::2, rvq_tc2.cu) whenever the codebook size allows it, 1 = single-CTA
// (rvq_tc.cu).  MESHTOK_RVQ_TC=1 forces the single-CTA kernel (A/B measurements; at the MeshGPT shape the pair kernel
// takes 171 us per level against 182 us).
int rvq_tc_variant(int K) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("MESHTOK_RVQ_TC");
    forced = (e && e[0] == '1') ? 1 : 0;
  }
  return (forced == 1 || K % 256 != 0) ? 1 : 2;
}

// workspace needs of the screening search for up to R rows: partial candidates (+ the row image of the pair kernel)
int rvq_tc_sizes(int R, int K, int D, size_t* partial_bytes, size_t* rimg_bytes) {
  *rimg_bytes = 0;
  RvqTcPlan p1;
  int rc = rvq_tc_plan(R, K, D, &p1);
  if (rc != MESHTOK_OK) return rc;
  *partial_bytes = p1.partial_bytes;
  if (rvq_tc_variant(K) == 2) {
    RvqTc2Plan p2;
    if ((rc = rvq_tc2_plan(R, K, D, &p2)) != MESHTOK_OK) return rc;
    *partial_bytes = p2.partial_bytes;
    *rimg_bytes = p2.row_image_bytes;
  }
  return MESHTOK_OK;
}

int check_model(const meshtok_model* m) {
  MT_CHECK_ARG(m != nullptr, "model is null");
  MT_CHECK_ARG(m->abi_version == MESHTOK_ABI_VERSION, "model.abi_version %d != library %d", m->abi_version, MESHTOK_ABI_VERSION);
  MT_CHECK_ARG(m->nvf == 3 || m->nvf == 4, "model.nvf must be 3 or 4");
  MT_CHECK_ARG(m->num_sage_layers >= 1 && m->num_sage_layers <= MESHTOK_MAX_SAGE_LAYERS, "model.num_sage_layers out of range");
  MT_CHECK_ARG(m->num_quantizers >= 1 && m->num_quantizers <= 8, "model.num_quantizers out of range");
  MT_CHECK_ARG(m->dim_codebook >= 16 && m->dim_codebook % 16 == 0 && m->dim_codebook <= 512,
               "model.dim_codebook must be a multiple of 16 in [16, 512]");
  MT_CHECK_ARG(m->num_slots == m->nvf * 4 + 4, "model.num_slots must be nvf*4+4");
  MT_CHECK_ARG(m->sage[0].dim_in == m->dim_codebook, "first SAGE layer must take dim_codebook inputs");
  for (int l = 0; l < m->num_sage_layers; ++l) {
    MT_CHECK_ARG(m->sage[l].dim_in % 16 == 0 && m->sage[l].dim_out % 4 == 0 && m->sage[l].dim_in <= 1024 && m->sage[l].dim_out <= 1024,
                 "SAGE layer %d: dims (%d -> %d) must be multiples of 16 / 4 and <= 1024", l, m->sage[l].dim_in, m->sage[l].dim_out);
    if (l) MT_CHECK_ARG(m->sage[l].dim_in == m->sage[l - 1].dim_out, "SAGE layer %d input width mismatch", l);
  }
  return MESHTOK_OK;
}

int enc_dim(const meshtok_model* m) { return m->sage[m->num_sage_layers - 1].dim_out; }

// rows_cap / vrows_cap: how many packed face rows / quantizer (vertex) rows the batch can have at most - B*F and B*V for the padded
// layout; for the ragged layout the rows the caller's faces / vertices buffers hold, so that the workspace grows with the sum of the
// mesh sizes, not with B x the largest mesh.  Arrays indexed by input slot (b*F+f, b*V+v) keep their B*F / B*V size.
void carve_topology(Carver& c, Workspace* w, int B, int F, int nvf, int V, int E, int64_t cap, size_t rows_cap, size_t vrows_cap, meshtok_taps* taps) {
  meshtok_taps dummy;
  meshtok_taps* t = taps ? taps : &dummy;
  const size_t BF = (size_t)B * F, BV = (size_t)B * V;
  const size_t RF = rows_cap, RV = vrows_cap;
  // one block, cleared by ONE memset together with mesh_counts behind it: [0..8) status word, [8..16) Counts, [16..20) RVQ statistics
  w->status = c.take<int>(64, &t->status);
  w->counts = reinterpret_cast<Counts*>(w->status ? w->status + 8 : nullptr);
  w->rvq_stats = w->status ? w->status + 16 : nullptr;
  t->counts = t->status + 32;
  t->rvq_stats = t->status + 64;
  w->mesh_counts = c.take<int>(2 * ((size_t)B * 8 + 4));   // two look-back blocks (maps launch, edges launch), each
                                                            // [B][4] aggregates, [B][4] inclusive prefixes, ticket; one memset
  w->offs = c.take<int>(3 * (size_t)(B + 1));
  w->deg_out = c.take<int>(BF);
  w->deg_in = c.take<int>(BF);
  w->face_row = c.take<int>(BF, &t->face_row);
  w->row_face = c.take<int>(RF);
  w->vert_row = c.take<int>(BV, &t->vert_row);
  w->vptr = c.take<int>(RV + 1);
  w->vinc = c.take<int>(RF * nvf);
  w->indptr = c.take<int>(RF + 1, &t->indptr);
  w->src = c.take<int>((size_t)cap, &t->src);
  if (E > 0) {
    w->ecount = c.take<int>(RF);
    w->cursor = c.take<int>(RF);
    w->etmp = c.take<int>((size_t)cap);
  } else {
    w->ecount = w->cursor = w->etmp = nullptr;
  }
}

int carve_all(const meshtok_model* m, int B, int F, int V, int E, int64_t cap, long long total_faces, long long total_vertices, void* base, Workspace* w,
              meshtok_taps* taps) {
  meshtok_taps dummy;
  meshtok_taps* t = taps ? taps : &dummy;
  memset(t, 0, sizeof(*t));
  Carver c(base);
  const size_t slots = (size_t)B * F;
  // (inside this function BF / BV are ROW capacities; slot-indexed arrays use `slots`)
  const size_t BF = total_faces > 0 ? (size_t)min((long long)slots, total_faces) : slots;
  const size_t BV = total_vertices > 0 ? (size_t)min((long long)B * V, total_vertices) : (size_t)B * V;
  carve_topology(c, w, B, F, m->nvf, V, E, cap, BF, BV, t);
  w->lcsr = c.take<uint8_t>((size_t)B * local_csr_stride(F));
  w->rows_cap = (int)BF;
  w->vrows_cap = (int)BV;
  const int D = m->dim_codebook, Q = m->num_quantizers;
  int max_in = 0;
  for (int l = 0; l < m->num_sage_layers; ++l) max_in = max(max_in, m->sage[l].dim_in);
  bool images = D % 32 == 0;
  int max_w = D;
  for (int l = 0; l < m->num_sage_layers; ++l) {
    images = images && m->sage[l].dim_in % 32 == 0 && m->sage[l].dim_out % 32 == 0;
    max_w = max(max_w, max(m->sage[l].dim_in, m->sage[l].dim_out));
  }
  w->x0 = c.take<float>(BF * D, &t->x0);
  w->xp = c.take<float>(BF * max_in);
  w->agg = c.take<float>(BF * max_in);
  w->img_x[0] = images ? c.take<uint8_t>(rows_image_bytes((long long)BF, max_w)) : nullptr;
  w->img_x[1] = images ? c.take<uint8_t>(rows_image_bytes((long long)BF, max_w)) : nullptr;
  w->img_agg = images ? c.take<uint8_t>(rows_image_bytes((long long)BF, max_in)) : nullptr;
  w->ssq = c.take<float>(BF * 16);
  w->bins = c.take<uint16_t>(slots * (size_t)(m->nvf * 4 + 4));
  for (int l = 0; l < m->num_sage_layers; ++l) w->layer_out[l] = c.take<float>(BF * m->sage[l].dim_out, &t->layer_out[l]);
  w->y = c.take<float>(BF * m->nvf * D);
  w->avg = c.take<float>(BV * D, &t->averaged);
  w->resid = c.take<float>(BV * D);
  w->rscale = c.take<float>(BV);
  w->qrows = c.take<float>(BV * D);
  w->pad_qrow = c.take<float>(D);
  w->zero_row = c.take<float>(D);
  w->vcodes = c.take<int32_t>(BV * Q, &t->vertex_codes);
  w->best = c.take<int32_t>(BV);
  w->packed = c.take<unsigned long long>(BV);
  w->partial = nullptr;
  w->rimg = nullptr;
  if (!m->use_lfq && m->codebook_f16_tiles != nullptr) {
    size_t pb = 0, rb = 0;
    int rc = rvq_tc_sizes((int)BV, m->codebook_size, D, &pb, &rb);
    if (rc != MESHTOK_OK) return rc;
    w->partial = c.take<float4>(pb / sizeof(float4));
    if (rb) w->rimg = c.take<uint8_t>(rb);
  }
  w->total = align_up(c.off, 256);
  return MESHTOK_OK;
}

__global__ void sub_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int D,
                                const int* n_dev, long long max_elems) {
  const long long n = min((long long)(*n_dev) * D, max_elems);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a[i] - b[i];
}

__global__ void copy_rows_kernel(const float4* __restrict__ a, float4* __restrict__ out, int d4, const int* n_dev,
                                 long long max_elems4) {
  const long long n = min((long long)(*n_dev) * d4, max_elems4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a[i];
}

// fp32 SIMT linear on fp32 rows, or the tcgen05 split-bf16 linear on activation images
int linear(bool simt, const float* A1, const void* A1_img, int K1, const float* A2, const void* A2_img, int K2, const float* W,
           const void* w_img, const float* bias, float* C, int N, const int* m_dev, int m_max, bool relu, cudaStream_t s,
           const LinearRowEpilogue* epi = nullptr) {
  if (simt) return launch_linear_simt(A1, K1, A2, K2, W, bias, C, N, m_dev, m_max, relu, s);
  MT_CHECK_ARG(w_img != nullptr, "tcgen05 linear requested but the model carries no weight image (set gemm_simt or build images)");
  MT_CHECK_ARG(A1_img != nullptr, "tcgen05 linear requested but the layer widths do not allow activation images (set gemm_simt)");
  return launch_linear_tc(A1_img, K1, A2_img, K2, w_img, bias, C, N, m_dev, m_max, relu, epi, s);
}

// One residual-VQ pass over `n` rows (device count) starting from rows_in; leaves the codes in vcodes (R x Q) and, with
// need_resid, the final residual in `resid` (else the residual of the last-but-one level: the last subtraction is skipped).
// tcgen05 path: TWO launches per level - the screening search and the re-score kernel, which also scans the rare ambiguous row,
// writes the code and updates the residual / row scale / row image.  stats[1] counts the rows that needed the exact scan
// (clear_stats: zero it here; the pipeline has cleared it with its status block already).
int run_rvq(const meshtok_model* m, const float* rows_in, float* resid, float* rscale, int32_t* vcodes,
            unsigned long long* packed, int* stats, float4* partial, uint8_t* rimg, const int* n_dev, int max_rows,
            bool exact, bool prepped, bool need_resid, bool clear_stats, cudaStream_t s) {
  const int D = m->dim_codebook, K = m->codebook_size, Q = m->num_quantizers;
  if (max_rows == 0) return MESHTOK_OK;
  int rc;
  const int variant = exact ? 0 : rvq_tc_variant(K);
  RvqTcPlan plan;
  RvqTc2Plan plan2;
  RvqPartialLayout layout;
  memset(&layout, 0, sizeof(layout));
  if (!exact) {
    MT_CHECK_ARG(m->codebook_f16_tiles != nullptr && m->codebook_image_scale > 0.f,
                 "model.codebook_f16_tiles / codebook_image_scale not set (needed by the tcgen05 search)");
    MT_CHECK_ARG(partial != nullptr, "no workspace for the tcgen05 search candidates");
    if (variant == 2) {
      MT_CHECK_ARG(rimg != nullptr, "no workspace for the residual row image");
      if ((rc = rvq_tc2_plan(max_rows, K, D, &plan2)) != MESHTOK_OK) return rc;
      layout = rvq_tc2_partial_layout(plan2, max_rows);
    } else {
      if ((rc = rvq_tc_plan(max_rows, K, D, &plan)) != MESHTOK_OK) return rc;
      layout.mode = 1;
      layout.stride = plan.n_splits;
    }
  }
  uint8_t* row_image = variant == 2 ? rimg : nullptr;
  const float c_scale = m->codebook_image_scale / (2.f * rvq_e2_kappa(m->codebook_max_norm));   // powers of two: exact
  if (!prepped) {   // (the pipeline's scatter_mean already left resid / rscale / the row image)
    ProfScope ps(PROF_RVQ_PREP_UPDATE, s);
    if ((rc = launch_rvq_prep(rows_in, D, resid, rscale, row_image, c_scale, n_dev, max_rows, s)) != MESHTOK_OK) return rc;
  }
  if (clear_stats) MT_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(int) * 4, s));
  for (int level = 0; level < Q; ++level) {
    const bool last = level + 1 == Q;
    if (exact) {
      { ProfScope ps(PROF_RVQ_EXACT, s); rc = launch_rvq_exact(resid, D, m->codebook, m->codebook_sqnorm, K, nullptr, n_dev, max_rows, packed, s); }
      if (rc != MESHTOK_OK) return rc;
      { ProfScope ps(PROF_RVQ_PREP_UPDATE, s);
        if (last && !need_resid) rc = launch_rvq_codes_only(packed, vcodes, Q, level, K, n_dev, max_rows, s);
        else rc = launch_rvq_update(resid, D, m->codebook, K, packed, vcodes, Q, level, rscale, nullptr, c_scale, n_dev, max_rows, s); }
      if (rc != MESHTOK_OK) return rc;
      continue;
    }
    { ProfScope ps(PROF_RVQ_SEARCH, s);
      if (variant == 2)
        rc = launch_rvq_tc2_search(rimg, rscale, D, m->codebook_f16_tiles, m->codebook_image_scale, n_dev,
                                   max_rows, plan2, partial, s);
      else
        rc = launch_rvq_tc_search(resid, rscale, D, m->codebook_f16_tiles, m->codebook_image_scale, m->codebook_sqnorm, K,
                                  n_dev, max_rows, plan, partial, s); }
    if (rc != MESHTOK_OK) return rc;
    { ProfScope ps(PROF_RVQ_RESCORE, s);
      rc = launch_rvq_rescore(resid, rscale, D, m->codebook, m->codebook_sqnorm, K, m->codebook_max_norm,
                              m->codebook_image_scale, variant == 2 ? c_scale : 0.f, partial, layout, n_dev, max_rows,
                              vcodes, Q, level, !last || need_resid, last ? nullptr : row_image, stats, s); }
    if (rc != MESHTOK_OK) return rc;
  }
  return MESHTOK_OK;
}

}  // namespace
}  // namespace meshtok

using namespace meshtok;

extern "C" {

int meshtok_version(void) { return MESHTOK_ABI_VERSION; }

int meshtok_debug_linear_trace(long long* out, int n) { return linear_tc_read_trace(out, n); }
int meshtok_debug_gather_trace(long long* out, int n) { return gather_mean_read_trace(out, n); }

const char* meshtok_last_error_string(void) { return g_error; }

int meshtok_topology_workspace_bytes(int B, int F, int nvf, int V, int E_user, size_t* bytes) {
  MT_CHECK_ARG(bytes != nullptr, "bytes is null");
  int rc = topo_check_shape(B, F, nvf, V);
  if (rc != MESHTOK_OK) return rc;
  Carver c(nullptr);
  Workspace w;
  carve_topology(c, &w, B, F, nvf, V, E_user, 16, (size_t)B * F, (size_t)B * V, nullptr);
  *bytes = align_up(c.off, 256);
  return MESHTOK_OK;
}

static int face_edges_common(const int64_t* faces, int B, int F, int nvf, int V, int64_t pad_id, int threshold,
                             int include_self, void* workspace, size_t workspace_bytes, TopoParams* p, Workspace* w) {
  MT_CHECK_ARG(faces != nullptr && workspace != nullptr, "faces / workspace is null");
  MT_CHECK_ARG(threshold == 1 || threshold == 2, "threshold must be 1 or 2");
  int rc = topo_check_shape(B, F, nvf, V);
  if (rc != MESHTOK_OK) return rc;
  Carver c(workspace);
  carve_topology(c, w, B, F, nvf, V, 0, 16, (size_t)B * F, (size_t)B * V, nullptr);
  if (c.off > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", c.off, workspace_bytes);
    return MESHTOK_ERR_WORKSPACE;
  }
  memset(p, 0, sizeof(*p));
  p->faces = faces; p->B = B; p->F = F; p->nvf = nvf; p->V = V; p->pad_id = pad_id;
  p->threshold = threshold; p->include_self = include_self; p->do_adjacency = 1;
  p->status = w->status; p->mesh_counts = w->mesh_counts; p->deg_out = w->deg_out; p->deg_in = w->deg_in;
  return MESHTOK_OK;
}

int meshtok_face_edges_count(const int64_t* faces, int B, int F, int nvf, int V, int64_t pad_id, int threshold,
                             int include_self, void* workspace, size_t workspace_bytes, int32_t* edge_counts,
                             meshtok_stream_t stream) {
  TopoParams p;
  Workspace w;
  int rc = face_edges_common(faces, B, F, nvf, V, pad_id, threshold, include_self, workspace, workspace_bytes, &p, &w);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(edge_counts != nullptr, "edge_counts is null");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MT_CHECK_CUDA(cudaMemsetAsync(w.status, 0, 64, s));
  p.edge_counts_out = edge_counts;
  return topo_count(p, s);
}

int meshtok_face_edges_fill(const int64_t* faces, int B, int F, int nvf, int V, int64_t pad_id, int threshold,
                            int include_self, void* workspace, size_t workspace_bytes, int emax, int64_t* out_edges,
                            meshtok_stream_t stream) {
  TopoParams p;
  Workspace w;
  int rc = face_edges_common(faces, B, F, nvf, V, pad_id, threshold, include_self, workspace, workspace_bytes, &p, &w);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(emax >= 0 && (emax == 0 || out_edges != nullptr), "bad emax / out_edges");
  p.emax = emax;
  p.dense_out = out_edges;
  return topo_fill_dense(p, static_cast<cudaStream_t>(stream));
}

int meshtok_tokenize_workspace_bytes_ragged(const meshtok_model* model, int B, int F, int V, int E, int64_t edge_capacity, int64_t total_faces,
                                            int64_t total_vertices, size_t* bytes) {
  MT_CHECK_ARG(bytes != nullptr, "bytes is null");
  int rc = check_model(model);
  if (rc != MESHTOK_OK) return rc;
  rc = topo_check_shape(B, F, model->nvf, V);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(edge_capacity > 0 && edge_capacity < (1ll << 31), "edge_capacity must be in (0, 2^31)");
  MT_CHECK_ARG(total_faces >= 0 && total_vertices >= 0, "total_faces / total_vertices must be >= 0 (0: B*F / B*V)");
  Workspace w;
  rc = carve_all(model, B, F, V, E, edge_capacity, total_faces, total_vertices, nullptr, &w, nullptr);
  if (rc != MESHTOK_OK) return rc;
  *bytes = w.total;
  return MESHTOK_OK;
}

int meshtok_tokenize_workspace_bytes(const meshtok_model* model, int B, int F, int V, int E, int64_t edge_capacity,
                                     size_t* bytes) {
  return meshtok_tokenize_workspace_bytes_ragged(model, B, F, V, E, edge_capacity, 0, 0, bytes);
}

int meshtok_tokenize_taps_ragged(const meshtok_model* model, int B, int F, int V, int E, int64_t edge_capacity, int64_t total_faces,
                                 int64_t total_vertices, meshtok_taps* taps) {
  MT_CHECK_ARG(taps != nullptr, "taps is null");
  int rc = check_model(model);
  if (rc != MESHTOK_OK) return rc;
  Workspace w;
  return carve_all(model, B, F, V, E, edge_capacity, total_faces, total_vertices, nullptr, &w, taps);
}

int meshtok_tokenize_taps(const meshtok_model* model, int B, int F, int V, int E, int64_t edge_capacity,
                          meshtok_taps* taps) {
  return meshtok_tokenize_taps_ragged(model, B, F, V, E, edge_capacity, 0, 0, taps);
}

int meshtok_tokenize(const meshtok_model* m, const meshtok_tokenize_args* a, void* workspace, size_t workspace_bytes,
                     meshtok_stream_t stream) {
  int rc = check_model(m);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(a != nullptr && workspace != nullptr, "args / workspace is null");
  MT_CHECK_ARG(a->faces != nullptr, "faces is null");
  MT_CHECK_ARG(a->vertices != nullptr || a->encoded_in != nullptr, "vertices is null");
  MT_CHECK_ARG(a->vertices != nullptr || a->vertex_offsets == nullptr || a->encoded_in != nullptr, "vertices is null");
  MT_CHECK_ARG(a->stop_after_encode || a->codes_out != nullptr, "codes_out is null");
  MT_CHECK_ARG(a->face_edges == nullptr || a->E > 0, "face_edges given but E <= 0 (pass one all-pad edge for an empty list)");
  const int B = a->B, F = a->F, V = a->V, nvf = m->nvf, D = m->dim_codebook, Q = m->num_quantizers;
  const int E = a->face_edges ? a->E : 0;
  const bool ragged = a->face_offsets != nullptr || a->vertex_offsets != nullptr;
  if (ragged) {
    MT_CHECK_ARG(a->face_offsets != nullptr && a->vertex_offsets != nullptr && a->total_faces >= 0 && a->total_faces <= (long long)B * F,
                 "ragged input needs face_offsets and vertex_offsets (B + 1 int32 each) and 0 <= total_faces <= B * F");
    MT_CHECK_ARG(a->total_vertices >= 0, "total_vertices must be >= 0 (0: unknown, the workspace is sized for B * V vertex rows)");
    MT_CHECK_ARG(a->face_edges == nullptr || (a->edge_offsets != nullptr && a->total_edges >= 0 && a->total_edges <= a->edge_capacity),
                 "ragged input with face_edges needs edge_offsets (B + 1 int32) and 0 <= total_edges <= edge_capacity");
  }
  rc = topo_check_shape(B, F, nvf, V);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(a->edge_capacity > 0 && a->edge_capacity < (1ll << 31), "edge_capacity must be in (0, 2^31)");
  Workspace w;
  rc = carve_all(m, B, F, V, E, a->edge_capacity, ragged ? a->total_faces : 0, ragged ? a->total_vertices : 0, workspace, &w, nullptr);
  if (rc != MESHTOK_OK) return rc;
  if (w.total > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.total, workspace_bytes);
    return MESHTOK_ERR_WORKSPACE;
  }
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) {
    set_error("workspace must be 256-byte aligned");
    return MESHTOK_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int slots = B * F;                                            // input face slots of the padded layout
  const int max_rows = w.rows_cap, max_vrows = w.vrows_cap;           // packed rows the activation buffers hold
  const bool simt = a->gemm_simt != 0;

  // ---- topology -------------------------------------------------------------------------------
  // status word, counts, RVQ statistics and the look-back descriptors behind them: ONE memset node
  MT_CHECK_CUDA(cudaMemsetAsync(w.status, 0, (size_t)(reinterpret_cast<char*>(w.mesh_counts) - reinterpret_cast<char*>(w.status)) +
                                                 sizeof(int) * 2 * ((size_t)B * 8 + 4), s));
  TopoParams p;
  memset(&p, 0, sizeof(p));
  p.faces = a->faces; p.B = B; p.F = F; p.nvf = nvf; p.V = V; p.pad_id = a->pad_id;
  p.in_face_off = a->face_offsets; p.in_vert_off = a->vertex_offsets;
  p.threshold = 2; p.include_self = 1; p.do_adjacency = a->face_edges ? 0 : 1;
  p.status = w.status; p.mesh_counts = w.mesh_counts; p.deg_out = w.deg_out; p.deg_in = w.deg_in;
  p.mesh_prefix = w.mesh_counts + 4 * (size_t)B; p.ticket = w.mesh_counts + 8 * (size_t)B; p.counts = w.counts;
  p.face_off = w.offs; p.vert_off = w.offs + (B + 1); p.edge_off = w.offs + 2 * (size_t)(B + 1);
  p.face_row = w.face_row; p.row_face = w.row_face; p.vert_row = w.vert_row; p.vptr = w.vptr; p.vinc = w.vinc;
  p.indptr = w.indptr; p.src = w.src; p.edge_capacity = a->edge_capacity;
  // ---- feature bins: gather, derived features, discretisation -> table row ids per INPUT face.  Nothing of it depends on the topology
  // kernel, so it runs on the side stream next to the maps launch (round 1 ran it after the maps, on the critical path).
  const bool want_features = a->encoded_in == nullptr;
  FeatParams fp;
  memset(&fp, 0, sizeof(fp));
  SideStream* side = side_stream();
  MT_CHECK_ARG(side != nullptr, "could not create the side stream");
  bool bins_pending = false;
  if (want_features) {
    fp.vertices = a->vertices; fp.faces = a->faces; fp.row_face = w.row_face; fp.counts = w.counts;
    fp.in_face_off = a->face_offsets; fp.in_vert_off = a->vertex_offsets;
    fp.pad_id = a->pad_id; fp.total_faces = a->total_faces;
    fp.B = B; fp.F = F; fp.V = V; fp.nvf = nvf; fp.D = D;
    fp.coor_lo = m->coor_lo;
    fp.coor_den = (float)((double)m->coor_hi - (double)m->coor_lo);
    fp.angle_den = (float)3.141592653589793;
    fp.area_den = (float)(((double)m->coor_hi - (double)m->coor_lo) * ((double)m->coor_hi - (double)m->coor_lo));
    fp.clip_lo = (float)(-1.0 + 1e-5); fp.clip_hi = (float)(1.0 - 1e-5);
    fp.n_coor = m->num_discrete_coors; fp.n_angle = m->num_discrete_angle; fp.n_area = m->num_discrete_area;
    fp.n_normal = m->num_discrete_normals;
    fp.tables = m->folded_tables; fp.bias = m->project_in_bias;
    {
      int off = 0, s_i = 0;
      for (int k = 0; k < nvf * 3; ++k) { fp.slot_off[s_i++] = off; off += m->num_discrete_coors; }
      for (int k = 0; k < nvf; ++k) { fp.slot_off[s_i++] = off; off += m->num_discrete_angle; }
      fp.slot_off[s_i++] = off; off += m->num_discrete_area;
      for (int k = 0; k < 3; ++k) { fp.slot_off[s_i++] = off; off += m->num_discrete_normals; }
    }
    fp.x0 = (simt || a->keep_taps) ? w.x0 : nullptr; fp.x0_image = simt ? nullptr : w.img_x[0]; fp.face_coords_out = a->face_coords_out;
    fp.bins_out = w.bins;
    if (a->face_coords_out)
      MT_CHECK_CUDA(cudaMemsetAsync(a->face_coords_out, 0, sizeof(int64_t) * (size_t)(ragged ? a->total_faces : slots) * nvf * 3, s));
    MT_CHECK_CUDA(cudaEventRecord(side->begin, s));
    MT_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->begin, 0));
    if ((rc = launch_face_bins(fp, ragged ? a->total_faces : (long long)slots, side->stream)) != MESHTOK_OK) return rc;
    MT_CHECK_CUDA(cudaEventRecord(side->bins_done, side->stream));
    bins_pending = true;
  }
  // The face adjacency reads nothing the maps launch writes (it recomputes the packed rows from the faces and has its own look-back
  // block), so it does not have to wait for it.  MESHTOK_EDGES_EARLY=1 launches it behind the feature bins on the side stream, BEFORE
  // the maps launch: it is then out of the folded-table sum's way sooner (that kernel's 144 CTAs otherwise share the SMs with the 64
  // one-per-SM adjacency CTAs: 54 -> 37 us for the section) and a serial step takes 0.795 instead of 0.804 ms - but host to host, with
  // two batches in flight, the same change costs 1.4 % (68.4 -> 67.4 M faces/s at C2: three one-CTA-per-SM launches at the head of every
  // graph get in the way of the other batch's tensor-core kernels).  The headline is the host-to-host number: off by default.
  static int edges_early_on = -1;
  if (edges_early_on < 0) { const char* e = getenv("MESHTOK_EDGES_EARLY"); edges_early_on = (e && atoi(e) == 1) ? 1 : 0; }
  const bool edges_early = edges_early_on && want_features && !a->face_edges;
  TopoParams pe = p;
  pe.mesh_counts = w.mesh_counts + (8 * (size_t)B + 4);
  pe.mesh_prefix = pe.mesh_counts + 4 * (size_t)B;
  pe.ticket = pe.mesh_counts + 8 * (size_t)B;
  if (edges_early && (rc = topo_edges(pe, side->stream)) != MESHTOK_OK) return rc;
  prof_begin(PROF_TOPOLOGY, s);
  // Two launches: the maps (packed face rows, vertex incidence CSR) are all the feature kernel and the first linear
  // need; the face adjacency (the expensive part, one CTA per mesh = fewer CTAs than SMs for small batches) runs on a
  // side stream next to them and is joined before the first neighbour aggregation.  Inside a stream capture the side
  // stream becomes a parallel branch of the graph.
  // Maps and adjacency in ONE launch on the caller's stream (topology_kernel<CSR>: the faces are staged and sorted once instead of
  // twice).  The folded-table sum then waits for the adjacency too (topology section 27 -> 40 us at C2), but it no longer shares the
  // SMs with the 64 one-per-SM adjacency CTAs of the side stream (its section: 54 -> 38 us), and the step has one launch less:
  // measured on one box each, two runs per setting: C2 0.7985 -> 0.795 ms (host to host 68.8 -> 69.0 M faces/s), C3 0.5725 -> 0.568 ms,
  // C4 1.776 -> 1.748 ms, 1 024 meshes per launch 12.50 -> 12.39 ms.  MESHTOK_TOPO_ONE=0: the two launches of before (A/B).
  static int topo_one_on = -1;
  if (topo_one_on < 0) { const char* e = getenv("MESHTOK_TOPO_ONE"); topo_one_on = (e && atoi(e) == 0) ? 0 : 1; }
  const bool topo_one = topo_one_on && want_features && !a->face_edges && !edges_early;   // (quantize() without an encoder pass keeps the pair)
  if (topo_one) {   // the one launch also writes the aggregation kernels' per-mesh CSR blocks (no local_csr_kernel, no side-stream join)
    p.lcsr = w.lcsr;
    p.lcsr_stride = (int)local_csr_stride(F);
    local_csr_layout(F, &p.lcsr_ptr_bytes, &p.lcsr_zero_row);
  }
  if ((rc = topo_one ? topo_csr(p, s) : topo_maps(p, s)) != MESHTOK_OK) return rc;
  bool edges_pending = false;
  if (a->face_edges) {
    rc = topo_user_edges(a->face_edges, B, E, ragged ? a->edge_offsets : nullptr, a->total_edges, a->pad_id, w.mesh_counts, w.offs, w.counts, w.ecount, w.cursor, w.etmp,
                         w.indptr, w.src, a->edge_capacity, max_rows, w.status, s);
    if (rc != MESHTOK_OK) return rc;
    if (a->encoded_in == nullptr) {   // the per-mesh CSR blocks: on the side stream, next to the folded-table sum and the first linear
      MT_CHECK_CUDA(cudaEventRecord(side->fork, s));
      MT_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
      if ((rc = launch_local_csr(w.indptr, w.src, a->edge_capacity, w.offs, B, F, w.lcsr, side->stream)) != MESHTOK_OK) return rc;
      MT_CHECK_CUDA(cudaEventRecord(side->join, side->stream));
      edges_pending = true;
    }
  } else if (topo_one) {
    // (everything the aggregation needs came out of the one topology launch)
  } else {
    MT_CHECK_CUDA(cudaEventRecord(side->fork, s));
    MT_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));   // (the per-mesh CSR blocks below need the maps launch's face offsets)
    if (!edges_early && !topo_one && (rc = topo_edges(pe, side->stream)) != MESHTOK_OK) return rc;
    // the aggregation kernels' per-mesh CSR blocks: once per batch, behind the adjacency on the side stream
    if (a->encoded_in == nullptr && (rc = launch_local_csr(w.indptr, w.src, a->edge_capacity, w.offs, B, F, w.lcsr, side->stream)) != MESHTOK_OK) return rc;
    MT_CHECK_CUDA(cudaEventRecord(side->join, side->stream));
    edges_pending = true;
  }
  auto join_edges = [&]() -> int {
    if (edges_pending) {
      MT_CHECK_CUDA(cudaStreamWaitEvent(s, side->join, 0));
      edges_pending = false;
    }
    if (bins_pending) {   // (a path that never ran face_sum: the side stream must still rejoin before the call returns)
      MT_CHECK_CUDA(cudaStreamWaitEvent(s, side->bins_done, 0));
      bins_pending = false;
    }
    return MESHTOK_OK;
  };
  prof_end(PROF_TOPOLOGY, s);
  const int* n_faces_dev = &w.counts->n_faces;
  const int* n_vrows_dev = &w.counts->n_vrows;

  // ---- encoder --------------------------------------------------------------------------------
  const int L = m->num_sage_layers;
  const float* enc = nullptr;
  const uint8_t* enc_img = w.img_x[0];
  const float* enc_ssq = nullptr;   // set when the encoder left its output rows un-normalised (deferred to project_dim_codebook)
  int enc_ssq_parts = 0;
  if (a->encoded_in == nullptr) {
    MT_CHECK_CUDA(cudaStreamWaitEvent(s, side->bins_done, 0));
    bins_pending = false;
    { ProfScope ps(PROF_FACE_EMBED, s); if ((rc = launch_face_sum(fp, max_rows, s)) != MESHTOK_OK) return rc; }

    // tcgen05 path: every activation that feeds a linear exists as a split-bf16 image, written by the kernel that
    // produces it (img_x[cur]: the layer input, img_agg: the aggregated neighbours).  Where the layer is narrow enough
    // for a single pass, the L2 normalisation (+ SiLU + LayerNorm after the first layer) runs in the epilogue of the
    // second linear, which writes the next layer's image directly; fp32 rows are then only kept for the parity taps.
    const float* x = w.x0;
    int cur = 0;
    bool xp_ready = false;   // the previous layer's launch already computed this layer's relu(lin(x)) (two linears back to back)
    // relu(lin(x)) of the current layer.  Where the aggregation is fused into the output linear, that launch GATHERS from these rows while
    // its second phase writes the next layer's - so the two alternate between w.xp and w.agg (free on the tcgen05 path)
    float* xp_cur = w.xp;
    float* xp_other = w.agg;
    for (int l = 0; l < L; ++l) {
      const meshtok_sage_layer& ly = m->sage[l];
      const bool first = l == 0, last = l == L - 1;
      const int mode = first ? 2 : 1;
      const bool fused = !simt && linear_tc_fused_norm_ok(ly.dim_out, mode) == MESHTOK_OK;
      // a last layer too wide for the fused epilogue leaves its rows un-normalised (image + squared norms) and
      // project_dim_codebook divides by the norm in ITS epilogue: W (x / |x|) + b = (W x) / |x| + b
      const bool deferred = !simt && !fused && last && !first && linear_tc_ssq_parts(ly.dim_out) <= 16;
      const bool need_fp32 = a->keep_taps || (last && (a->encoded_out != nullptr || a->stop_after_encode));
      // xp = relu(lin(x))
      if (!xp_ready) {
        ProfScope ps(PROF_SAGE_LINEAR, s);
        if ((rc = linear(simt, x, w.img_x[cur], ly.dim_in, nullptr, nullptr, 0, ly.lin_w, ly.lin_w_img, ly.lin_b, xp_cur, ly.dim_in, n_faces_dev, max_rows, true, s)) != MESHTOK_OK) return rc;
      }
      xp_ready = false;
      if ((rc = join_edges()) != MESHTOK_OK) return rc;
      // the next layer's lin reads exactly the rows the output linear writes: both run in one launch, tile by tile
      const bool b2b = fused && !last && ly.cat_w_img != nullptr && m->sage[l + 1].lin_w_img != nullptr && m->sage[l + 1].dim_in == ly.dim_out &&
                       linear_tc_b2b_ok(ly.dim_out, mode);
      // ... and the neighbour aggregation in front of it becomes a role of that launch (gather warps write the A tiles): no aggregate image
      const bool gather_fused = b2b && w.img_agg != nullptr && linear_tc_gather_ok(ly.dim_in, ly.dim_out, mode);
      if (!gather_fused) {
        ProfScope ps(PROF_SAGE_GATHER, s);
        if ((rc = launch_gather_mean(xp_cur, ly.dim_in, w.indptr, w.src, a->edge_capacity, simt ? w.agg : nullptr, simt ? nullptr : w.img_agg, w.counts, max_rows, w.offs, B, F, s, w.lcsr)) != MESHTOK_OK) return rc;
      }
      // out = lin_l(agg) + lin_r(x)  [-> normalise (-> SiLU -> LayerNorm)]
      if (fused) {
        LinearRowEpilogue epi{mode, first ? m->ln_gamma : nullptr, first ? m->ln_beta : nullptr, w.img_x[cur ^ 1], nullptr, nullptr, 0};
        ProfScope ps(PROF_SAGE_LINEAR, s);
        if (b2b) {
          const LinearGather lg{xp_cur, w.indptr, w.src, a->edge_capacity};
          float* xp_next = gather_fused ? xp_other : xp_cur;
          if ((rc = launch_linear_tc_b2b(w.img_agg, ly.dim_in, w.img_x[cur], ly.dim_in, ly.cat_w_img, ly.lin_l_b, need_fp32 ? w.layer_out[l] : nullptr,
                                         ly.dim_out, n_faces_dev, max_rows, &epi, m->sage[l + 1].lin_w_img, m->sage[l + 1].lin_b, xp_next, s,
                                         gather_fused ? &lg : nullptr)) != MESHTOK_OK) return rc;
          if (gather_fused) { xp_other = xp_cur; xp_cur = xp_next; }
          xp_ready = true;
        } else if ((rc = linear(false, nullptr, w.img_agg, ly.dim_in, nullptr, w.img_x[cur], ly.dim_in, ly.cat_w, ly.cat_w_img, ly.lin_l_b,
                                need_fp32 ? w.layer_out[l] : nullptr, ly.dim_out, n_faces_dev, max_rows, false, s, &epi)) != MESHTOK_OK) return rc;
      } else if (deferred) {
        LinearRowEpilogue epi{0, nullptr, nullptr, w.img_x[cur ^ 1], w.ssq, nullptr, 0};
        { ProfScope ps(PROF_SAGE_LINEAR, s);
          if ((rc = linear(false, nullptr, w.img_agg, ly.dim_in, nullptr, w.img_x[cur], ly.dim_in, ly.cat_w, ly.cat_w_img, ly.lin_l_b,
                           need_fp32 ? w.layer_out[l] : nullptr, ly.dim_out, n_faces_dev, max_rows, false, s, &epi)) != MESHTOK_OK) return rc; }
        enc_ssq = w.ssq;
        enc_ssq_parts = linear_tc_ssq_parts(ly.dim_out);
        if (need_fp32) {   // encode() / the parity taps want the normalised fp32 rows
          ProfScope ps(PROF_SAGE_NORM, s);
          if ((rc = launch_row_norm(w.layer_out[l], ly.dim_out, nullptr, nullptr, nullptr, w.counts, max_rows, s)) != MESHTOK_OK) return rc;
        }
      } else {
        { ProfScope ps(PROF_SAGE_LINEAR, s); if ((rc = linear(simt, w.agg, w.img_agg, ly.dim_in, x, w.img_x[cur], ly.dim_in, ly.cat_w, ly.cat_w_img, ly.lin_l_b, w.layer_out[l], ly.dim_out, n_faces_dev, max_rows, false, s)) != MESHTOK_OK) return rc; }
        { ProfScope ps(PROF_SAGE_NORM, s);
          if ((rc = launch_row_norm(w.layer_out[l], ly.dim_out, first ? m->ln_gamma : nullptr, first ? m->ln_beta : nullptr,
                                    simt ? nullptr : w.img_x[cur ^ 1], w.counts, max_rows, s)) != MESHTOK_OK) return rc; }
      }
      cur ^= 1;
      x = w.layer_out[l];
    }
    enc_img = w.img_x[cur];
    enc = x;
  } else {
    if ((rc = join_edges()) != MESHTOK_OK) return rc;   // quantize(): no encoder, but the side stream must rejoin
    if ((rc = launch_pack_rows(a->encoded_in, w.row_face, enc_dim(m), w.layer_out[L - 1], w.counts, max_rows, a->face_offsets, F, s)) != MESHTOK_OK) return rc;
    enc = w.layer_out[L - 1];
    if (!simt && !a->stop_after_encode) {
      MT_CHECK_ARG(w.img_x[0] != nullptr, "tcgen05 linear requested but the layer widths do not allow activation images (set gemm_simt)");
      if ((rc = launch_rows_to_image(enc, enc_dim(m), n_faces_dev, max_rows, w.img_x[0], s)) != MESHTOK_OK) return rc;
    }
  }
  if (a->encoded_out)
    if ((rc = launch_unpack_rows(enc, w.face_row, slots, enc_dim(m), a->encoded_out, a->face_offsets, B, F, a->total_faces, s)) != MESHTOK_OK) return rc;
  if (a->stop_after_encode) return MESHTOK_OK;

  // ---- quantize -------------------------------------------------------------------------------
  { ProfScope ps(PROF_PDC_LINEAR, s);
    LinearRowEpilogue epi{0, nullptr, nullptr, nullptr, nullptr, enc_ssq, enc_ssq_parts};
    if ((rc = linear(simt, enc, enc_img, enc_dim(m), nullptr, nullptr, 0, m->pdc_w, m->pdc_w_img, m->pdc_b, w.y, nvf * D, n_faces_dev, max_rows, false, s,
                     enc_ssq ? &epi : nullptr)) != MESHTOK_OK) return rc; }
  const bool rvq_tc_path = !m->use_lfq && !a->rvq_exact && m->codebook_f16_tiles != nullptr;
  { ProfScope ps(PROF_SCATTER_MEAN, s);
    RvqPrepOut prep;
    memset(&prep, 0, sizeof(prep));
    if (rvq_tc_path) {
      prep.resid = w.resid; prep.rscale = w.rscale;
      if (rvq_tc_variant(m->codebook_size) == 2) {
        prep.row_image = w.rimg;
        prep.c_scale = m->codebook_image_scale / (2.f * rvq_e2_kappa(m->codebook_max_norm));
      }
    }
    if ((rc = launch_scatter_mean(w.y, D, w.vptr, w.vinc, w.avg, w.counts, max_vrows, rvq_tc_path ? &prep : nullptr, s)) != MESHTOK_OK) return rc; }
  const float* pad_row = nullptr;
  if (m->use_lfq) {
    int bits = 0;
    while ((1 << bits) < m->codebook_size) ++bits;
    MT_CHECK_ARG((1 << bits) == m->codebook_size, "LFQ needs a power-of-two codebook_size");
    MT_CHECK_ARG(m->lfq_in_w && m->lfq_in_b, "model.lfq_in_* is null");
    float* qrows = a->quantized_out ? w.qrows : nullptr;
    if (qrows) MT_CHECK_ARG(m->lfq_out_w && m->lfq_out_b, "model.lfq_out_* is null (needed for quantized_out)");
    { ProfScope ps(PROF_LFQ, s);
      if ((rc = launch_lfq(w.avg, D, m->lfq_in_w, m->lfq_in_b, bits, Q, m->lfq_out_w, m->lfq_out_b, w.vcodes, qrows,
                           n_vrows_dev, max_vrows, s)) != MESHTOK_OK) return rc; }
    if (qrows) {  // the reference also quantises its all-zero pad-vertex row; padded faces gather that
      MT_CHECK_CUDA(cudaMemsetAsync(w.zero_row, 0, sizeof(float) * D, s));
      if ((rc = launch_lfq(w.zero_row, D, m->lfq_in_w, m->lfq_in_b, bits, Q, m->lfq_out_w, m->lfq_out_b, w.best, w.pad_qrow,
                           nullptr, 1, s)) != MESHTOK_OK) return rc;
      pad_row = w.pad_qrow;
    }
  } else {
    MT_CHECK_ARG(m->codebook && m->codebook_sqnorm, "model.codebook / codebook_sqnorm is null");
    if ((rc = run_rvq(m, w.avg, w.resid, w.rscale, w.vcodes, w.packed, w.rvq_stats, w.partial, w.rimg, n_vrows_dev, max_vrows,
                      a->rvq_exact != 0, rvq_tc_path, a->quantized_out != nullptr, /*clear_stats=*/false, s)) != MESHTOK_OK) return rc;
    if (a->quantized_out) {
      const int blocks = max(1, min((int)(((long long)max_vrows * D + 255) / 256), 148 * 8));
      sub_rows_kernel<<<blocks, 256, 0, s>>>(w.avg, w.resid, w.qrows, D, n_vrows_dev, (long long)max_vrows * D);
      MT_CHECK_LAUNCH();
    }
  }
  { ProfScope ps(PROF_GATHER_CODES, s);
    GatherHist gh{w.vptr, w.counts, max_vrows, reinterpret_cast<unsigned long long*>(a->histogram)};
    if (ragged) rc = launch_gather_codes_ragged(a->faces, a->face_offsets, w.face_row, w.vert_row, w.vcodes, a->total_faces, B, F, nvf, V, Q, a->pad_id,
                                                a->codes_out, a->histogram ? &gh : nullptr, m->codebook_size, s);
    else rc = launch_gather_codes(a->faces, w.face_row, w.vert_row, w.vcodes, B, F, nvf, V, Q, a->pad_id, a->codes_out,
                                  a->histogram ? &gh : nullptr, m->codebook_size, s);
    if (rc != MESHTOK_OK) return rc; }
  if (a->quantized_out)
    if ((rc = launch_gather_quantized(a->faces, w.face_row, w.vert_row, w.qrows, B, F, nvf, V, D, pad_row, a->quantized_out, a->face_offsets,
                                      a->total_faces, s)) != MESHTOK_OK) return rc;
  return MESHTOK_OK;
}

// ---- per-stage entry points: the stages of meshtok_tokenize one by one, on the SAME workspace layout, with fp32 rows as
// the interchange format so that a reference maintainer can swap a single stage.  Every stage validates the arguments it uses, carves
// the workspace exactly like meshtok_tokenize and launches the kernels the pipeline launches for that stage.
namespace {
struct StageCtx {
  Workspace w;
  int B, F, V, E, nvf, D, Q, slots, max_rows, max_vrows;
  bool ragged, simt;
  cudaStream_t s;
};
int stage_begin(const meshtok_model* m, const meshtok_tokenize_args* a, void* workspace, size_t workspace_bytes, meshtok_stream_t stream, StageCtx* c) {
  int rc = check_model(m);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(a != nullptr && workspace != nullptr, "args / workspace is null");
  c->B = a->B; c->F = a->F; c->V = a->V; c->nvf = m->nvf; c->D = m->dim_codebook; c->Q = m->num_quantizers;
  c->E = a->face_edges ? a->E : 0;
  c->ragged = a->face_offsets != nullptr || a->vertex_offsets != nullptr;
  if (c->ragged)
    MT_CHECK_ARG(a->face_offsets != nullptr && a->vertex_offsets != nullptr && a->total_faces >= 0 && a->total_faces <= (long long)c->B * c->F,
                 "ragged input needs face_offsets and vertex_offsets (B + 1 int32 each) and 0 <= total_faces <= B * F");
  rc = topo_check_shape(c->B, c->F, c->nvf, c->V);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(a->edge_capacity > 0 && a->edge_capacity < (1ll << 31), "edge_capacity must be in (0, 2^31)");
  rc = carve_all(m, c->B, c->F, c->V, c->E, a->edge_capacity, c->ragged ? a->total_faces : 0, c->ragged ? a->total_vertices : 0, workspace, &c->w, nullptr);
  if (rc != MESHTOK_OK) return rc;
  if (c->w.total > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", c->w.total, workspace_bytes);
    return MESHTOK_ERR_WORKSPACE;
  }
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) {
    set_error("workspace must be 256-byte aligned");
    return MESHTOK_ERR_WORKSPACE;
  }
  c->slots = c->B * c->F;
  c->max_rows = c->w.rows_cap;
  c->max_vrows = c->w.vrows_cap;
  c->simt = a->gemm_simt != 0;
  c->s = static_cast<cudaStream_t>(stream);
  return MESHTOK_OK;
}
}  // namespace

int meshtok_stage_topology(const meshtok_model* m, const meshtok_tokenize_args* a, void* workspace, size_t workspace_bytes, meshtok_stream_t stream) {
  StageCtx c;
  int rc = stage_begin(m, a, workspace, workspace_bytes, stream, &c);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(a->faces != nullptr, "faces is null");
  MT_CHECK_ARG(a->face_edges == nullptr || a->E > 0, "face_edges given but E <= 0 (pass one all-pad edge for an empty list)");
  Workspace& w = c.w;
  const int B = c.B;
  MT_CHECK_CUDA(cudaMemsetAsync(w.status, 0, (size_t)(reinterpret_cast<char*>(w.mesh_counts) - reinterpret_cast<char*>(w.status)) +
                                                 sizeof(int) * 2 * ((size_t)B * 8 + 4), c.s));
  TopoParams p;
  memset(&p, 0, sizeof(p));
  p.faces = a->faces; p.B = B; p.F = c.F; p.nvf = c.nvf; p.V = c.V; p.pad_id = a->pad_id;
  p.in_face_off = a->face_offsets; p.in_vert_off = a->vertex_offsets;
  p.threshold = 2; p.include_self = 1; p.do_adjacency = a->face_edges ? 0 : 1;
  p.status = w.status; p.mesh_counts = w.mesh_counts; p.deg_out = w.deg_out; p.deg_in = w.deg_in;
  p.mesh_prefix = w.mesh_counts + 4 * (size_t)B; p.ticket = w.mesh_counts + 8 * (size_t)B; p.counts = w.counts;
  p.face_off = w.offs; p.vert_off = w.offs + (B + 1); p.edge_off = w.offs + 2 * (size_t)(B + 1);
  p.face_row = w.face_row; p.row_face = w.row_face; p.vert_row = w.vert_row; p.vptr = w.vptr; p.vinc = w.vinc;
  p.indptr = w.indptr; p.src = w.src; p.edge_capacity = a->edge_capacity;
  if ((rc = topo_maps(p, c.s)) != MESHTOK_OK) return rc;
  if (a->face_edges)
    return topo_user_edges(a->face_edges, B, c.E, c.ragged ? a->edge_offsets : nullptr, a->total_edges, a->pad_id, w.mesh_counts, w.offs, w.counts, w.ecount,
                           w.cursor, w.etmp, w.indptr, w.src, a->edge_capacity, c.max_rows, w.status, c.s);
  TopoParams pe = p;
  pe.mesh_counts = w.mesh_counts + (8 * (size_t)B + 4);
  pe.mesh_prefix = pe.mesh_counts + 4 * (size_t)B;
  pe.ticket = pe.mesh_counts + 8 * (size_t)B;
  return topo_edges(pe, c.s);
}

int meshtok_stage_face_features_embed(const meshtok_model* m, const meshtok_tokenize_args* a, float* x0_rows, void* workspace, size_t workspace_bytes,
                                      meshtok_stream_t stream) {
  StageCtx c;
  int rc = stage_begin(m, a, workspace, workspace_bytes, stream, &c);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(a->faces != nullptr && a->vertices != nullptr && x0_rows != nullptr, "faces / vertices / x0_rows is null");
  Workspace& w = c.w;
  FeatParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.vertices = a->vertices; fp.faces = a->faces; fp.row_face = w.row_face; fp.counts = w.counts;
  fp.in_face_off = a->face_offsets; fp.in_vert_off = a->vertex_offsets;
  fp.pad_id = a->pad_id; fp.total_faces = a->total_faces;
  fp.B = c.B; fp.F = c.F; fp.V = c.V; fp.nvf = c.nvf; fp.D = c.D;
  fp.coor_lo = m->coor_lo;
  fp.coor_den = (float)((double)m->coor_hi - (double)m->coor_lo);
  fp.angle_den = (float)3.141592653589793;
  fp.area_den = (float)(((double)m->coor_hi - (double)m->coor_lo) * ((double)m->coor_hi - (double)m->coor_lo));
  fp.clip_lo = (float)(-1.0 + 1e-5); fp.clip_hi = (float)(1.0 - 1e-5);
  fp.n_coor = m->num_discrete_coors; fp.n_angle = m->num_discrete_angle; fp.n_area = m->num_discrete_area; fp.n_normal = m->num_discrete_normals;
  fp.tables = m->folded_tables; fp.bias = m->project_in_bias;
  {
    int off = 0, s_i = 0;
    for (int k = 0; k < c.nvf * 3; ++k) { fp.slot_off[s_i++] = off; off += m->num_discrete_coors; }
    for (int k = 0; k < c.nvf; ++k) { fp.slot_off[s_i++] = off; off += m->num_discrete_angle; }
    fp.slot_off[s_i++] = off; off += m->num_discrete_area;
    for (int k = 0; k < 3; ++k) { fp.slot_off[s_i++] = off; off += m->num_discrete_normals; }
  }
  fp.x0 = x0_rows; fp.x0_image = nullptr; fp.face_coords_out = a->face_coords_out; fp.bins_out = w.bins;
  if (a->face_coords_out)
    MT_CHECK_CUDA(cudaMemsetAsync(a->face_coords_out, 0, sizeof(int64_t) * (size_t)(c.ragged ? a->total_faces : c.slots) * c.nvf * 3, c.s));
  if ((rc = launch_face_bins(fp, c.ragged ? a->total_faces : (long long)c.slots, c.s)) != MESHTOK_OK) return rc;
  return launch_face_sum(fp, c.max_rows, c.s);
}

// xp = relu(lin(x)) of SAGE layer `layer`: the `project=True` half of SAGEConv
int meshtok_stage_sage_project(const meshtok_model* m, const meshtok_tokenize_args* a, int layer, const float* x_rows, float* xp_rows, void* workspace,
                               size_t workspace_bytes, meshtok_stream_t stream) {
  StageCtx c;
  int rc = stage_begin(m, a, workspace, workspace_bytes, stream, &c);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(layer >= 0 && layer < m->num_sage_layers && x_rows && xp_rows, "bad layer / x_rows / xp_rows is null");
  const meshtok_sage_layer& ly = m->sage[layer];
  Workspace& w = c.w;
  const bool simt = c.simt || ly.lin_w_img == nullptr || w.img_x[0] == nullptr;
  if (!simt && (rc = launch_rows_to_image(x_rows, ly.dim_in, &w.counts->n_faces, c.max_rows, w.img_x[0], c.s)) != MESHTOK_OK) return rc;
  return linear(simt, x_rows, w.img_x[0], ly.dim_in, nullptr, nullptr, 0, ly.lin_w, ly.lin_w_img, ly.lin_b, xp_rows, ly.dim_in, &w.counts->n_faces,
                c.max_rows, true, c.s);
}

// out = normalize(lin_l(mean over in-edges of xp) + lin_r(x)) [layer 0: -> SiLU -> LayerNorm]: the rest of SAGEConv
int meshtok_stage_sage_aggregate_linear(const meshtok_model* m, const meshtok_tokenize_args* a, int layer, const float* x_rows, const float* xp_rows,
                                        float* out_rows, void* workspace, size_t workspace_bytes, meshtok_stream_t stream) {
  StageCtx c;
  int rc = stage_begin(m, a, workspace, workspace_bytes, stream, &c);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(layer >= 0 && layer < m->num_sage_layers && x_rows && xp_rows && out_rows, "bad layer / row pointers are null");
  const meshtok_sage_layer& ly = m->sage[layer];
  Workspace& w = c.w;
  const bool first = layer == 0;
  const bool simt = c.simt || ly.cat_w_img == nullptr || w.img_x[0] == nullptr;
  const int* n_dev = &w.counts->n_faces;
  if ((rc = launch_gather_mean(xp_rows, ly.dim_in, w.indptr, w.src, a->edge_capacity, simt ? w.agg : nullptr, simt ? nullptr : w.img_agg, w.counts,
                               c.max_rows, w.offs, c.B, c.F, c.s)) != MESHTOK_OK) return rc;
  if (!simt) {
    if ((rc = launch_rows_to_image(x_rows, ly.dim_in, n_dev, c.max_rows, w.img_x[0], c.s)) != MESHTOK_OK) return rc;
    const int mode = first ? 2 : 1;
    if (linear_tc_fused_norm_ok(ly.dim_out, mode) == MESHTOK_OK) {
      LinearRowEpilogue epi{mode, first ? m->ln_gamma : nullptr, first ? m->ln_beta : nullptr, nullptr, nullptr, nullptr, 0};
      return linear(false, nullptr, w.img_agg, ly.dim_in, nullptr, w.img_x[0], ly.dim_in, ly.cat_w, ly.cat_w_img, ly.lin_l_b, out_rows, ly.dim_out, n_dev,
                    c.max_rows, false, c.s, &epi);
    }
  }
  if ((rc = linear(simt, w.agg, w.img_agg, ly.dim_in, x_rows, w.img_x[0], ly.dim_in, ly.cat_w, ly.cat_w_img, ly.lin_l_b, out_rows, ly.dim_out, n_dev,
                   c.max_rows, false, c.s)) != MESHTOK_OK) return rc;
  return launch_row_norm(out_rows, ly.dim_out, first ? m->ln_gamma : nullptr, first ? m->ln_beta : nullptr, nullptr, w.counts, c.max_rows, c.s);
}

// averaged[vertex row] = scatter_mean of project_dim_codebook(enc) over the faces of every referenced vertex (quantize()'s first half)
int meshtok_stage_project_scatter_mean(const meshtok_model* m, const meshtok_tokenize_args* a, const float* enc_rows, float* averaged_rows, void* workspace,
                                       size_t workspace_bytes, meshtok_stream_t stream) {
  StageCtx c;
  int rc = stage_begin(m, a, workspace, workspace_bytes, stream, &c);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(enc_rows && averaged_rows, "enc_rows / averaged_rows is null");
  Workspace& w = c.w;
  const int ed = enc_dim(m);
  const bool simt = c.simt || m->pdc_w_img == nullptr || w.img_x[0] == nullptr;
  if (!simt && (rc = launch_rows_to_image(enc_rows, ed, &w.counts->n_faces, c.max_rows, w.img_x[0], c.s)) != MESHTOK_OK) return rc;
  if ((rc = linear(simt, enc_rows, w.img_x[0], ed, nullptr, nullptr, 0, m->pdc_w, m->pdc_w_img, m->pdc_b, w.y, c.nvf * c.D, &w.counts->n_faces, c.max_rows,
                   false, c.s)) != MESHTOK_OK) return rc;
  return launch_scatter_mean(w.y, c.D, w.vptr, w.vinc, averaged_rows, w.counts, c.max_vrows, nullptr, c.s);
}

// codes_out[token] = vertex_codes[vertex row of the token's vertex] (pad_id on padded faces) (+ histogram): quantize()'s last step
int meshtok_stage_gather_codes(const meshtok_model* m, const meshtok_tokenize_args* a, const int32_t* vertex_codes, void* workspace, size_t workspace_bytes,
                               meshtok_stream_t stream) {
  StageCtx c;
  int rc = stage_begin(m, a, workspace, workspace_bytes, stream, &c);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(a->faces && a->codes_out && vertex_codes, "faces / codes_out / vertex_codes is null");
  Workspace& w = c.w;
  GatherHist gh{w.vptr, w.counts, c.max_vrows, reinterpret_cast<unsigned long long*>(a->histogram)};
  if (c.ragged)
    return launch_gather_codes_ragged(a->faces, a->face_offsets, w.face_row, w.vert_row, vertex_codes, a->total_faces, c.B, c.F, c.nvf, c.V, c.Q, a->pad_id,
                                      a->codes_out, a->histogram ? &gh : nullptr, m->codebook_size, c.s);
  return launch_gather_codes(a->faces, w.face_row, w.vert_row, vertex_codes, c.B, c.F, c.nvf, c.V, c.Q, a->pad_id, a->codes_out,
                             a->histogram ? &gh : nullptr, m->codebook_size, c.s);
}

int meshtok_lfq_codes(const meshtok_model* m, const float* rows, int R, int32_t* codes, meshtok_stream_t stream) {
  int rc = check_model(m);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(rows && codes && R >= 0, "rows / codes is null");
  int bits = 0;
  while ((1 << bits) < m->codebook_size) ++bits;
  MT_CHECK_ARG((1 << bits) == m->codebook_size, "LFQ needs a power-of-two codebook_size");
  MT_CHECK_ARG(m->lfq_in_w && m->lfq_in_b, "model.lfq_in_* is null");
  return launch_lfq(rows, m->dim_codebook, m->lfq_in_w, m->lfq_in_b, bits, m->num_quantizers, nullptr, nullptr, codes,
                    nullptr, nullptr, R, static_cast<cudaStream_t>(stream));
}

// carves the stand-alone RVQ workspace; base == nullptr only measures
static int carve_rvq(const meshtok_model* m, int R, void* base, float** resid, float** rscale, unsigned long long** packed,
                     int** stats, int** n_dev, float4** partial, uint8_t** rimg, size_t* total) {
  size_t pb = 0, rb = 0;
  const bool tc = m->codebook_f16_tiles != nullptr || base == nullptr;
  if (tc) {
    int rc = rvq_tc_sizes(R, m->codebook_size, m->dim_codebook, &pb, &rb);
    if (rc != MESHTOK_OK) { if (base != nullptr) return rc; pb = rb = 0; }   // shapes only the exact path covers
  }
  Carver c(base);
  *resid = c.take<float>((size_t)R * m->dim_codebook);
  *rscale = c.take<float>(R);
  *packed = c.take<unsigned long long>(R);
  *stats = c.take<int>(4);
  *n_dev = c.take<int>(4);
  *partial = pb ? c.take<float4>(pb / sizeof(float4)) : nullptr;
  *rimg = rb ? c.take<uint8_t>(rb) : nullptr;
  *total = align_up(c.off, 256);
  return MESHTOK_OK;
}

int meshtok_rvq_workspace_bytes(const meshtok_model* m, int R, size_t* bytes) {
  int rc = check_model(m);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(bytes && R >= 0, "bad arguments");
  float *resid, *rscale; unsigned long long* packed; int *stats, *n_dev; float4* partial; uint8_t* rimg;
  return carve_rvq(m, R, nullptr, &resid, &rscale, &packed, &stats, &n_dev, &partial, &rimg, bytes);
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

int meshtok_rvq_codes(const meshtok_model* m, const float* rows, int R, int exact, int32_t* codes, float* residual_out,
                      void* workspace, size_t workspace_bytes, meshtok_stream_t stream) {
  int rc = check_model(m);
  if (rc != MESHTOK_OK) return rc;
  MT_CHECK_ARG(rows && codes && workspace && R >= 0, "rows / codes / workspace is null");
  MT_CHECK_ARG(m->codebook && m->codebook_sqnorm, "model.codebook / codebook_sqnorm is null");
  float *resid, *rscale; unsigned long long* packed; int *stats, *n_dev; float4* partial; uint8_t* rimg;
  size_t total = 0;
  if ((rc = carve_rvq(m, R, workspace, &resid, &rscale, &packed, &stats, &n_dev, &partial, &rimg, &total)) != MESHTOK_OK) return rc;
  if (total > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", total, workspace_bytes);
    return MESHTOK_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  set_int_kernel<<<1, 1, 0, s>>>(n_dev, R);
  MT_CHECK_LAUNCH();
  rc = run_rvq(m, rows, resid, rscale, codes, packed, stats, partial, rimg, n_dev, R, exact != 0, false, residual_out != nullptr, /*clear_stats=*/true, s);
  if (rc != MESHTOK_OK) return rc;
  if (residual_out && R > 0)
    MT_CHECK_CUDA(cudaMemcpyAsync(residual_out, resid, sizeof(float) * (size_t)R * m->dim_codebook, cudaMemcpyDeviceToDevice, s));
  return MESHTOK_OK;
}

int meshtok_rvq_codebook_image_bytes(int codebook_size, int dim, size_t* bytes) {
  MT_CHECK_ARG(bytes && codebook_size > 0 && dim > 0, "bad arguments");
  *bytes = rvq_codebook_image_bytes(codebook_size, dim);
  return MESHTOK_OK;
}

int meshtok_rvq_build_codebook_image(const float* codebook, const float* codebook_sqnorm, int codebook_size, int dim, float scale,
                                     float max_norm, void* image, meshtok_stream_t stream) {
  MT_CHECK_ARG(codebook && codebook_sqnorm && image, "codebook / codebook_sqnorm / image is null");
  MT_CHECK_ARG(scale > 0.f, "scale must be a positive power of two");
  MT_CHECK_ARG(max_norm >= 0.f, "max_norm must be max_j |e_j|");
  return launch_rvq_build_image(codebook, codebook_sqnorm, codebook_size, dim, scale, max_norm, image, static_cast<cudaStream_t>(stream));
}

int meshtok_linear_image_bytes(int N, int K, size_t* bytes) {
  MT_CHECK_ARG(bytes && N > 0 && K > 0, "bad arguments");
  int passes, NT;
  int rc = linear_tc_plan(N, K, 0, &passes, &NT);
  if (rc != MESHTOK_OK) return rc;
  *bytes = linear_tc_image_bytes(N, K);
  return MESHTOK_OK;
}

int meshtok_linear_build_image(const float* W, int N, int K, void* image, meshtok_stream_t stream) {
  MT_CHECK_ARG(W && image, "W / image is null");
  return launch_linear_tc_build_image(W, N, K, image, static_cast<cudaStream_t>(stream));
}

int meshtok_rows_image_bytes(int M, int K, size_t* bytes) {
  MT_CHECK_ARG(bytes && M >= 0 && K > 0 && K % 32 == 0, "rows image: K must be a positive multiple of 32");
  *bytes = rows_image_bytes(M, K);
  return MESHTOK_OK;
}

int meshtok_rows_build_image(const float* A, int M, int K, void* image, meshtok_stream_t stream) {
  MT_CHECK_ARG(A && image && M >= 0, "A / image is null");
  return launch_rows_to_image(A, K, nullptr, M, image, static_cast<cudaStream_t>(stream));
}

int meshtok_linear(const float* A, const void* a_image, const float* W, const void* w_image, const float* bias, float* C, int M,
                   int N, int K, int relu, int simt, meshtok_stream_t stream) {
  MT_CHECK_ARG(C && M >= 0 && N > 0 && K > 0, "bad arguments");
  MT_CHECK_ARG(simt ? (A != nullptr && W != nullptr) : (a_image != nullptr && w_image != nullptr),
               "A, W (simt) / a_image, w_image (tcgen05) is null");
  return linear(simt != 0, A, a_image, K, nullptr, nullptr, 0, W, w_image, bias, C, N, nullptr, M, relu != 0, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
