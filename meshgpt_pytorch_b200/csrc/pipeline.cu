// Host-to-host batch pipeline in native code: meshtok_pipeline_* of include/meshtok.h.
//
// What a CPU-side caller of the reference does per batch - MeshAutoencoder.tokenize(vertices, faces) on host tensors
// (meshgpt_pytorch.py:949-976 behind a DataLoader, trainer.py:176-177) - is here a ring of `depth` slots, each with its own
// stream, device staging buffers, workspace, CUDA graph of the whole tokenize pipeline and pinned result buffer:
//     submit(batch i):  [wait until slot i % depth is free]  H2D vertices, H2D faces -> graph launch -> D2H codes, D2H status word
//                       -> event, all on the slot's stream; returns immediately with the result of batch i - depth, if any.
// The copies of one batch run on the copy engines while the other slots' kernels are on the SMs.  Round 1 did this in Python
// (HostPipeline.submit: five torch calls and a stream context per batch, plus a blocking .item() for the status word); measured
// end to end it was host-bound and box dependent (52 vs 65 M faces/s on two driver boxes with identical device times).  Here a
// submit is six driver calls and the status word arrives in pinned memory with the codes - no synchronising read.
//
// Ownership: the caller (PyTorch) still owns every byte - device staging, workspaces, pinned buffers are passed in per slot.
// The pipeline object owns only its streams, events and graph executables (created in _create, released in _destroy); it is
// the one component of this library that is allowed to synchronise (on its own events, when a slot is reused or drained).
#include <string.h>

#include <new>
#include <vector>

#include "common.cuh"

using namespace meshtok;

struct meshtok_pipeline {
  meshtok_model model;
  meshtok_tokenize_args args;                 // template: per-slot pointers are patched in
  int depth = 0, device = 0;
  size_t vert_bytes = 0, face_bytes = 0, code_bytes = 0;
  bool ragged = false;                        // CSR-of-meshes layout: one graph captured on the capacities serves every batch
  int max_meshes = 0;
  long long cap_faces = 0, cap_vertices = 0;
  std::vector<meshtok_pipeline_ragged_slot> rslots;
  std::vector<long long> tokens;              // ragged: int64 code rows (faces * nvf) of the batch in flight in the slot
  std::vector<meshtok_pipeline_slot> slots;
  std::vector<cudaStream_t> streams;
  std::vector<cudaEvent_t> done;
  std::vector<cudaGraphExec_t> graphs;
  std::vector<long long> ticket;              // ticket of the batch in flight in the slot, -1: free
  long long next = 0;
};

namespace {

int collect(meshtok_pipeline* p, int s, meshtok_pipeline_result* out) {
  MT_CHECK_CUDA(cudaEventSynchronize(p->done[s]));
  if (out) {
    out->ticket = p->ticket[s];
    out->slot = s;
    out->status = p->ragged ? p->rslots[s].status_host[0] : p->slots[s].status_host[0];
    out->codes_host = p->ragged ? p->rslots[s].codes_host : p->slots[s].codes_host;
  }
  p->ticket[s] = -1;
  return MESHTOK_OK;
}

void release(meshtok_pipeline* p) {
  for (cudaGraphExec_t g : p->graphs) if (g) cudaGraphExecDestroy(g);
  for (cudaEvent_t e : p->done) if (e) cudaEventDestroy(e);
  for (cudaStream_t s : p->streams) if (s) cudaStreamDestroy(s);
  delete p;
}

}  // namespace

extern "C" {

int meshtok_pipeline_create(const meshtok_model* model, int B, int F, int V, int64_t pad_id, int64_t edge_capacity, int64_t* histogram,
                            const meshtok_pipeline_slot* slots, int depth, meshtok_pipeline** out) {
  MT_CHECK_ARG(model && slots && out, "model / slots / out is null");
  MT_CHECK_ARG(depth >= 1 && depth <= 16, "depth must be in [1, 16]");
  MT_CHECK_ARG(model->abi_version == MESHTOK_ABI_VERSION, "model.abi_version %d != library %d", model->abi_version, MESHTOK_ABI_VERSION);
  size_t need = 0;
  int rc = meshtok_tokenize_workspace_bytes(model, B, F, V, 0, edge_capacity, &need);
  if (rc != MESHTOK_OK) return rc;
  meshtok_pipeline* p = new (std::nothrow) meshtok_pipeline();
  MT_CHECK_ARG(p != nullptr, "out of host memory");
  p->model = *model;
  p->depth = depth;
  const int nvf = model->nvf, Q = model->num_quantizers;
  p->vert_bytes = sizeof(float) * (size_t)B * V * 3;
  p->face_bytes = sizeof(int64_t) * (size_t)B * F * nvf;
  p->code_bytes = sizeof(int64_t) * (size_t)B * F * nvf * Q;
  memset(&p->args, 0, sizeof(p->args));
  p->args.B = B; p->args.F = F; p->args.V = V; p->args.E = 0;
  p->args.pad_id = pad_id; p->args.edge_capacity = edge_capacity;
  p->args.histogram = histogram;
  p->slots.assign(slots, slots + depth);
  p->streams.assign(depth, nullptr);
  p->done.assign(depth, nullptr);
  p->graphs.assign(depth, nullptr);
  p->ticket.assign(depth, -1);
  auto fail = [&](int code) { release(p); return code; };
#define PL_CUDA(expr)                                                                                          \
  do {                                                                                                         \
    cudaError_t _e = (expr);                                                                                   \
    if (_e != cudaSuccess) {                                                                                   \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                   \
      return fail(MESHTOK_ERR_CUDA);                                                                           \
    }                                                                                                          \
  } while (0)
  PL_CUDA(cudaGetDevice(&p->device));
  for (int s = 0; s < depth; ++s) {
    const meshtok_pipeline_slot& sl = p->slots[s];
    if (!(sl.vertices && sl.faces && sl.codes && sl.workspace && sl.codes_host && sl.status_host) || sl.workspace_bytes < need) {
      set_error("pipeline slot %d: null buffer or workspace smaller than %zu bytes", s, need);
      return fail(MESHTOK_ERR_INVALID_ARGUMENT);
    }
    PL_CUDA(cudaStreamCreateWithFlags(&p->streams[s], cudaStreamNonBlocking));
    PL_CUDA(cudaEventCreateWithFlags(&p->done[s], cudaEventDisableTiming));
    meshtok_tokenize_args a = p->args;
    a.vertices = sl.vertices; a.faces = sl.faces; a.codes_out = sl.codes;
    a.histogram = nullptr;   // the eager warm-up must not count tokens
    // One eager run on whatever the staging buffers hold (the caller has put a valid example batch there): opts the kernels in
    // to their shared memory on this device, creates the thread's side stream - nothing of that may happen inside a capture.
    if ((rc = meshtok_tokenize(&p->model, &a, sl.workspace, sl.workspace_bytes, p->streams[s])) != MESHTOK_OK) return fail(rc);
    PL_CUDA(cudaStreamSynchronize(p->streams[s]));
    a.histogram = histogram;
    cudaGraph_t graph = nullptr;
    PL_CUDA(cudaStreamBeginCapture(p->streams[s], cudaStreamCaptureModeThreadLocal));
    rc = meshtok_tokenize(&p->model, &a, sl.workspace, sl.workspace_bytes, p->streams[s]);
    cudaError_t ce = cudaStreamEndCapture(p->streams[s], &graph);
    if (rc != MESHTOK_OK) { if (graph) cudaGraphDestroy(graph); return fail(rc); }
    if (ce != cudaSuccess) { set_error("stream capture of the tokenize pipeline failed: %s", cudaGetErrorString(ce)); return fail(MESHTOK_ERR_CUDA); }
    ce = cudaGraphInstantiate(&p->graphs[s], graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); return fail(MESHTOK_ERR_CUDA); }
  }
#undef PL_CUDA
  *out = p;
  return MESHTOK_OK;
}

int meshtok_pipeline_submit(meshtok_pipeline* p, const float* vertices_host, const int64_t* faces_host, meshtok_pipeline_result* finished) {
  MT_CHECK_ARG(p && !p->ragged && vertices_host && faces_host, "pipeline / host buffers are null (or a ragged pipeline: use meshtok_pipeline_submit_ragged)");
  const int s = (int)(p->next % p->depth);
  if (finished) { finished->ticket = -1; finished->slot = -1; finished->status = 0; finished->codes_host = nullptr; finished->tokens = 0; }
  if (p->ticket[s] >= 0) {
    int rc = collect(p, s, finished);
    if (rc != MESHTOK_OK) return rc;
  }
  const meshtok_pipeline_slot& sl = p->slots[s];
  cudaStream_t st = p->streams[s];
  MT_CHECK_CUDA(cudaMemcpyAsync(sl.vertices, vertices_host, p->vert_bytes, cudaMemcpyHostToDevice, st));
  MT_CHECK_CUDA(cudaMemcpyAsync(sl.faces, faces_host, p->face_bytes, cudaMemcpyHostToDevice, st));
  MT_CHECK_CUDA(cudaGraphLaunch(p->graphs[s], st));
  MT_CHECK_CUDA(cudaMemcpyAsync(sl.codes_host, sl.codes, p->code_bytes, cudaMemcpyDeviceToHost, st));
  MT_CHECK_CUDA(cudaMemcpyAsync(sl.status_host, sl.workspace, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MT_CHECK_CUDA(cudaEventRecord(p->done[s], st));
  p->ticket[s] = p->next++;
  return MESHTOK_OK;
}

int meshtok_pipeline_drain(meshtok_pipeline* p, meshtok_pipeline_result* results, int* n) {
  MT_CHECK_ARG(p && results && n, "pipeline / results / n is null");
  *n = 0;
  // oldest first: the slots after the one the next submit would use
  for (int k = 0; k < p->depth; ++k) {
    const int s = (int)((p->next + k) % p->depth);
    if (p->ticket[s] < 0) continue;
    int rc = collect(p, s, &results[*n]);
    if (rc != MESHTOK_OK) return rc;
    results[*n].tokens = p->ragged ? p->tokens[s] : 0;
    ++*n;
  }
  return MESHTOK_OK;
}

// ---- the same ring for the ragged layout: ONE graph per slot, captured on the capacities, serves batches of any
// composition (fewer meshes = repeated last offset = empty meshes; the kernels take the live counts from the offsets on the device)
int meshtok_pipeline_create_ragged(const meshtok_model* model, int max_meshes, int F, int V, int64_t cap_faces, int64_t cap_vertices, int64_t pad_id,
                                   int64_t edge_capacity, int64_t* histogram, const meshtok_pipeline_ragged_slot* slots, int depth,
                                   meshtok_pipeline** out) {
  MT_CHECK_ARG(model && slots && out, "model / slots / out is null");
  MT_CHECK_ARG(depth >= 1 && depth <= 16, "depth must be in [1, 16]");
  MT_CHECK_ARG(model->abi_version == MESHTOK_ABI_VERSION, "model.abi_version %d != library %d", model->abi_version, MESHTOK_ABI_VERSION);
  MT_CHECK_ARG(max_meshes >= 1 && cap_faces >= 1 && cap_vertices >= 1 && cap_faces <= (int64_t)max_meshes * F && cap_vertices <= (int64_t)max_meshes * V,
               "ragged pipeline: need 1 <= cap_faces <= max_meshes * F and 1 <= cap_vertices <= max_meshes * V");
  size_t need = 0;
  int rc = meshtok_tokenize_workspace_bytes_ragged(model, max_meshes, F, V, 0, edge_capacity, cap_faces, cap_vertices, &need);
  if (rc != MESHTOK_OK) return rc;
  meshtok_pipeline* p = new (std::nothrow) meshtok_pipeline();
  MT_CHECK_ARG(p != nullptr, "out of host memory");
  p->model = *model;
  p->depth = depth;
  p->ragged = true;
  p->max_meshes = max_meshes;
  p->cap_faces = cap_faces;
  p->cap_vertices = cap_vertices;
  memset(&p->args, 0, sizeof(p->args));
  p->args.B = max_meshes; p->args.F = F; p->args.V = V; p->args.E = 0;
  p->args.pad_id = pad_id; p->args.edge_capacity = edge_capacity;
  p->args.total_faces = cap_faces; p->args.total_vertices = cap_vertices;
  p->rslots.assign(slots, slots + depth);
  p->streams.assign(depth, nullptr);
  p->done.assign(depth, nullptr);
  p->graphs.assign(depth, nullptr);
  p->ticket.assign(depth, -1);
  p->tokens.assign(depth, 0);
  auto fail = [&](int code) { release(p); return code; };
#define PL_CUDA(expr)                                                                                          \
  do {                                                                                                         \
    cudaError_t _e = (expr);                                                                                   \
    if (_e != cudaSuccess) {                                                                                   \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                   \
      return fail(MESHTOK_ERR_CUDA);                                                                           \
    }                                                                                                          \
  } while (0)
  PL_CUDA(cudaGetDevice(&p->device));
  for (int s = 0; s < depth; ++s) {
    const meshtok_pipeline_ragged_slot& sl = p->rslots[s];
    if (!(sl.vertices && sl.faces && sl.offsets && sl.codes && sl.workspace && sl.codes_host && sl.status_host && sl.offsets_host) ||
        sl.workspace_bytes < need) {
      set_error("ragged pipeline slot %d: null buffer or workspace smaller than %zu bytes", s, need);
      return fail(MESHTOK_ERR_INVALID_ARGUMENT);
    }
    PL_CUDA(cudaStreamCreateWithFlags(&p->streams[s], cudaStreamNonBlocking));
    PL_CUDA(cudaEventCreateWithFlags(&p->done[s], cudaEventDisableTiming));
    meshtok_tokenize_args a = p->args;
    a.vertices = sl.vertices; a.faces = sl.faces; a.codes_out = sl.codes;
    a.face_offsets = sl.offsets; a.vertex_offsets = sl.offsets + (max_meshes + 1);
    a.histogram = nullptr;
    // the staging buffers hold a valid example (the caller put one there, e.g. a single triangle in mesh 0 and empty meshes behind it)
    if ((rc = meshtok_tokenize(&p->model, &a, sl.workspace, sl.workspace_bytes, p->streams[s])) != MESHTOK_OK) return fail(rc);
    PL_CUDA(cudaStreamSynchronize(p->streams[s]));
    a.histogram = histogram;
    cudaGraph_t graph = nullptr;
    PL_CUDA(cudaStreamBeginCapture(p->streams[s], cudaStreamCaptureModeThreadLocal));
    rc = meshtok_tokenize(&p->model, &a, sl.workspace, sl.workspace_bytes, p->streams[s]);
    cudaError_t ce = cudaStreamEndCapture(p->streams[s], &graph);
    if (rc != MESHTOK_OK) { if (graph) cudaGraphDestroy(graph); return fail(rc); }
    if (ce != cudaSuccess) { set_error("stream capture of the tokenize pipeline failed: %s", cudaGetErrorString(ce)); return fail(MESHTOK_ERR_CUDA); }
    ce = cudaGraphInstantiate(&p->graphs[s], graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); return fail(MESHTOK_ERR_CUDA); }
  }
#undef PL_CUDA
  *out = p;
  return MESHTOK_OK;
}

int meshtok_pipeline_submit_ragged(meshtok_pipeline* p, const float* vertices_host, const int64_t* faces_host, const int32_t* face_offsets_host,
                                   const int32_t* vertex_offsets_host, int n_meshes, meshtok_pipeline_result* finished) {
  MT_CHECK_ARG(p && p->ragged && face_offsets_host && vertex_offsets_host, "ragged pipeline / offset arrays are null (or a padded pipeline)");
  MT_CHECK_ARG(n_meshes >= 0 && n_meshes <= p->max_meshes, "ragged pipeline: %d meshes, at most %d per batch", n_meshes, p->max_meshes);
  const long long nf = face_offsets_host[n_meshes], nv = vertex_offsets_host[n_meshes];
  MT_CHECK_ARG((vertices_host || nv <= 0) && (faces_host || nf <= 0), "ragged pipeline: host buffers are null");
  for (int i = 0; i < n_meshes; ++i)
    MT_CHECK_ARG(face_offsets_host[i] <= face_offsets_host[i + 1] && vertex_offsets_host[i] <= vertex_offsets_host[i + 1],
                 "ragged pipeline: offsets must not decrease (mesh %d)", i);
  MT_CHECK_ARG(face_offsets_host[0] == 0 && vertex_offsets_host[0] == 0 && nf >= 0 && nf <= p->cap_faces && nv >= 0 && nv <= p->cap_vertices,
               "ragged pipeline: offsets must start at 0 and the batch must fit the capacities (%lld faces, %lld vertices)", p->cap_faces, p->cap_vertices);
  const int s = (int)(p->next % p->depth);
  if (finished) { finished->ticket = -1; finished->slot = -1; finished->status = 0; finished->codes_host = nullptr; finished->tokens = 0; }
  if (p->ticket[s] >= 0) {
    int rc = collect(p, s, finished);
    if (rc != MESHTOK_OK) return rc;
    if (finished) finished->tokens = p->tokens[s];
  }
  const meshtok_pipeline_ragged_slot& sl = p->rslots[s];
  cudaStream_t st = p->streams[s];
  const int B1 = p->max_meshes + 1, nvf = p->model.nvf, Q = p->model.num_quantizers;
  // offsets of the live meshes, the last one repeated for the empty meshes behind them (the slot's previous copy out of this pinned
  // staging array finished before its event did)
  int32_t* oh = sl.offsets_host;
  for (int i = 0; i <= n_meshes; ++i) { oh[i] = face_offsets_host[i]; oh[B1 + i] = vertex_offsets_host[i]; }
  for (int i = n_meshes + 1; i < B1; ++i) { oh[i] = (int32_t)nf; oh[B1 + i] = (int32_t)nv; }
  MT_CHECK_CUDA(cudaMemcpyAsync(sl.offsets, oh, sizeof(int32_t) * 2 * (size_t)B1, cudaMemcpyHostToDevice, st));
  if (nv > 0) MT_CHECK_CUDA(cudaMemcpyAsync(sl.vertices, vertices_host, sizeof(float) * 3 * (size_t)nv, cudaMemcpyHostToDevice, st));
  if (nf > 0) MT_CHECK_CUDA(cudaMemcpyAsync(sl.faces, faces_host, sizeof(int64_t) * (size_t)nvf * (size_t)nf, cudaMemcpyHostToDevice, st));
  MT_CHECK_CUDA(cudaGraphLaunch(p->graphs[s], st));
  if (nf > 0) MT_CHECK_CUDA(cudaMemcpyAsync(sl.codes_host, sl.codes, sizeof(int64_t) * (size_t)nvf * (size_t)Q * (size_t)nf, cudaMemcpyDeviceToHost, st));
  MT_CHECK_CUDA(cudaMemcpyAsync(sl.status_host, sl.workspace, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MT_CHECK_CUDA(cudaEventRecord(p->done[s], st));
  p->ticket[s] = p->next++;
  p->tokens[s] = nf * nvf;
  return MESHTOK_OK;
}

meshtok_stream_t meshtok_pipeline_stream(meshtok_pipeline* p, int slot) {
  return (p && slot >= 0 && slot < p->depth) ? static_cast<meshtok_stream_t>(p->streams[slot]) : nullptr;
}

int meshtok_pipeline_destroy(meshtok_pipeline* p) {
  if (p == nullptr) return MESHTOK_OK;
  for (int s = 0; s < p->depth; ++s)
    if (p->done[s] && p->ticket[s] >= 0) cudaEventSynchronize(p->done[s]);
  release(p);
  return MESHTOK_OK;
}

}  // extern "C"
