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
    out->status = p->slots[s].status_host[0];
    out->codes_host = p->slots[s].codes_host;
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
  MT_CHECK_ARG(p && vertices_host && faces_host, "pipeline / host buffers are null");
  const int s = (int)(p->next % p->depth);
  if (finished) { finished->ticket = -1; finished->slot = -1; finished->status = 0; finished->codes_host = nullptr; }
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
    ++*n;
  }
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
