// Launch counter + optional cudaEvent section timing (see meshtok_profile_* in include/meshtok.h).
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace meshtok {

long long g_launch_count = 0;

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MESHTOK_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: a process that moves to a second GPU must opt
// in again there.  One bit per (kernel, device); devices beyond 63 simply set the attribute before every launch.
int ensure_dynamic_smem(const void* kernel, int bytes) {
  struct Entry { const void* fn; int bytes; unsigned long long devices; };
  static std::mutex mu;
  static std::vector<Entry> table;
  int dev = 0;
  MT_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  Entry* e = nullptr;
  for (Entry& t : table)
    if (t.fn == kernel && t.bytes == bytes) { e = &t; break; }
  if (e == nullptr) { table.push_back(Entry{kernel, bytes, 0ull}); e = &table.back(); }
  if (dev >= 0 && dev < 64 && ((e->devices >> dev) & 1ull)) return MESHTOK_OK;
  MT_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (dev >= 0 && dev < 64) e->devices |= 1ull << dev;
  return MESHTOK_OK;
}

namespace {
struct Rec { int section; cudaEvent_t a, b; };
std::mutex g_mu;
bool g_enabled = false;
std::vector<Rec> g_open;      // begun, not yet ended (LIFO)
std::vector<Rec> g_done;
std::vector<cudaEvent_t> g_pool;

cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
const char* kNames[PROF_NUM_SECTIONS] = {"topology", "face_embed", "sage_linear", "sage_gather_mean", "sage_row_norm",
                                         "pdc_linear", "scatter_mean", "lfq", "rvq_prep_update", "rvq_search_tc",
                                         "rvq_rescore", "rvq_exact_scan", "gather_codes", "other"};
}  // namespace

void prof_begin(int section, cudaStream_t s) {
  if (!g_enabled) return;
  std::lock_guard<std::mutex> lk(g_mu);
  Rec r{section, get_event(), get_event()};
  if (r.a) cudaEventRecord(r.a, s);
  g_open.push_back(r);
}

void prof_end(int section, cudaStream_t s) {
  if (!g_enabled) return;
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = (int)g_open.size() - 1; i >= 0; --i) {
    if (g_open[i].section == section) {
      Rec r = g_open[i];
      g_open.erase(g_open.begin() + i);
      if (r.b) cudaEventRecord(r.b, s);
      g_done.push_back(r);
      return;
    }
  }
}

}  // namespace meshtok

using namespace meshtok;

extern "C" {

int64_t meshtok_kernel_launch_count(void) { return g_launch_count; }

int meshtok_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_enabled = on != 0;
  return MESHTOK_OK;
}

int meshtok_profile_sections(void) { return PROF_NUM_SECTIONS; }

const char* meshtok_profile_section_name(int section) {
  return (section >= 0 && section < PROF_NUM_SECTIONS) ? kNames[section] : "";
}

int meshtok_profile_collect(float* ms, int32_t* events, int n) {
  MT_CHECK_ARG(ms && events && n >= PROF_NUM_SECTIONS, "profile_collect: arrays must hold %d entries", (int)PROF_NUM_SECTIONS);
  std::lock_guard<std::mutex> lk(g_mu);
  for (const Rec& r : g_done) {
    float t = 0.f;
    if (r.a && r.b) {
      MT_CHECK_CUDA(cudaEventSynchronize(r.b));
      MT_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    }
    ms[r.section] += t;
    events[r.section] += 1;
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_done.clear();
  return MESHTOK_OK;
}

}  // extern "C"
