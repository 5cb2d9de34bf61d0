// tsm_api.cu - the extern "C" boundary of libtosemscan.so (include/tosemscan.h): context, device
// memory, H2D/D2H staging and kernel launches.  No torch types, no CPU fallback.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "tsm_device.cuh"

#include "tsm_scan_kernels.cuh"
#include "tsm_scan_walk.cuh"
#include "tsm_reduce_kernels.cuh"
#include "tsm_diff_kernels.cuh"
#include "tsm_stmt_kernels.cuh"
#include "tsm_lines_kernels.cuh"

using namespace tsm;

static const char* const kNames[TSM_K] = TSM_CAT_NAMES_INIT;

// Scratch memory of the diff / statements calls: grow-only, kept by the ctx between calls (a cudaMalloc +
// cudaFree pair per buffer and call costs more than the kernels of a C5-sized batch).
struct ScratchPool {
  struct Slot { void* p; size_t cap; bool used; };
  std::vector<Slot> slots;
  void* take(size_t bytes) {
    int best = -1;
    for (int i = 0; i < (int)slots.size(); ++i)
      if (!slots[(size_t)i].used && slots[(size_t)i].cap >= bytes && (best < 0 || slots[(size_t)i].cap < slots[(size_t)best].cap)) best = i;
    if (best < 0) {
      for (size_t i = 0; i < slots.size(); ++i)           // replace a free slot that is too small rather than pile up
        if (!slots[i].used) { cudaFree(slots[i].p); slots.erase(slots.begin() + (long)i); break; }
      void* q = nullptr;
      const size_t cap = bytes + bytes / 8 + 256;
      if (cudaMalloc(&q, cap) != cudaSuccess) return nullptr;
      slots.push_back(Slot{q, cap, true});
      return q;
    }
    slots[(size_t)best].used = true;
    return slots[(size_t)best].p;
  }
  void give(void* p) { for (Slot& s : slots) if (s.p == p) s.used = false; }
  void clear() { for (Slot& s : slots) cudaFree(s.p); slots.clear(); }
};
static thread_local ScratchPool* t_pool = nullptr;        // pool of the ctx whose call runs on this thread
struct PoolScope {
  explicit PoolScope(ScratchPool* p) { t_pool = p; }
  ~PoolScope() { t_pool = nullptr; }
};

struct DevBuf {                                           // device scratch from the ctx's pool, returned on scope exit
  void* p = nullptr;
  ~DevBuf() { reset(); }
  void reset() { if (p && t_pool) t_pool->give(p); p = nullptr; }
  bool alloc(size_t bytes) { reset(); p = t_pool ? t_pool->take(bytes ? bytes : 16) : nullptr; return p != nullptr; }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct SyncGuard {                                        // error paths: wait for the work queued on st before the DevBufs of the
  cudaStream_t st;                                        // scope hand their slots back to the pool (the next call may reuse or free them)
  explicit SyncGuard(cudaStream_t s) : st(s) {}
  ~SyncGuard() { cudaStreamSynchronize(st); }
};

struct tsm_ctx {
  int device = 0;
  ScratchPool pool;
  int cls_grid = 0; size_t cls_smem = (size_t)-1;        // launch shape of k_classify for the current histogram size
  int sms = 0;
  int64_t max_arena = 0;
  int32_t max_files = 0, max_groups = 0;
  int64_t max_events = 0;
  // device buffers
  uint8_t* d_arena = nullptr;
  int32_t* d_off = nullptr;
  int32_t* d_len = nullptr;
  uint8_t* d_ext = nullptr;
  uint16_t* d_grp = nullptr;
  uint32_t* d_unit_file = nullptr;
  uint32_t* d_unit_begin = nullptr;
  uint32_t unit_cap = 0;
  uint8_t* d_zero = nullptr;                // one allocation, zeroed by one memset per scan: ctrl | slab | counts
  Ctrl* d_ctrl = nullptr;
  SlabCtl* d_slab = nullptr;                // [kMaxSlabs]
  cudaStream_t copy_stream = nullptr;       // H2D of arena slabs, overlapped with the scan of earlier slabs
  cudaEvent_t slab_ev[64] = {};
  cudaEvent_t ready_ev = nullptr;
  cudaEvent_t diff_ev[8] = {};             // around the kernels of the diff path (tsm_diff_last_ms)
  uint8_t* h_diff = nullptr;               // 256 B pinned: what the diff path reads back between its kernels (Ctrl x 2, line totals, todo count)
  float diff_ms[3] = {0, 0, 0};            // k_scan over both sides, k_myers, k_myers_trace of the last diff
  struct HostSidePair* res_pair = nullptr; // sides kept in HBM by tsm_diff_upload
  static constexpr int kMaxSlabs = 64;
  tsm_file_stat* d_stats = nullptr;
  unsigned long long* d_cand = nullptr;
  tsm_header_event* d_hev = nullptr;
  tsm_assert_event* d_aev = nullptr;
  unsigned long long* d_counts = nullptr;   // [(max_groups + 1) * K + 4]
  Ctrl* h_ctrl = nullptr;                   // pinned
  // resident corpus
  bool resident = false, scanned = false;
  int32_t n_files = 0, n_groups = 1;
  int64_t arena_bytes = 0;
  uint32_t last_flags = 0;
  int launches = 0;
  // CUDA events around the 4 scan kernels: a ring of sets so that back-to-back scans can be timed
  // per kernel without a host sync inside the timed region.
  static constexpr int kRing = 32;
  cudaEvent_t ev[kRing][5] = {};
  bool ev_used[kRing] = {};
  int ev_next = 0, ev_last = -1;
  double ms_sum[4] = {0, 0, 0, 0};
  long long ms_n = 0;
};

// Fold the elapsed times of event set `i` into the running sums (waits for it if still in flight).
static void fold_events(tsm_ctx* c, int i) {
  if (!c->ev_used[i]) return;
  if (cudaEventSynchronize(c->ev[i][4]) == cudaSuccess) {
    float ms;
    bool ok = true;
    float t[4];
    for (int k = 0; k < 4; ++k) { ok &= cudaEventElapsedTime(&ms, c->ev[i][k], c->ev[i][k + 1]) == cudaSuccess; t[k] = ms; }
    if (ok) { for (int k = 0; k < 4; ++k) c->ms_sum[k] += t[k]; c->ms_n++; }
  }
  c->ev_used[i] = false;
}

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "tosemscan: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return TSM_E_CUDA; } } while (0)

extern "C" int tsm_abi_version(void) { return TSM_ABI_VERSION; }

extern "C" const char* tsm_strerror(int s) {
  switch (s) {
    case TSM_OK: return "ok";
    case TSM_E_ARG: return "bad argument";
    case TSM_E_LAYOUT: return "corpus violates the arena layout (docs/SPEC.md section 1)";
    case TSM_E_CAPACITY: return "corpus or event list exceeds the context capacity";
    case TSM_E_CUDA: return "CUDA error (no device, allocation or launch failure)";
    case TSM_E_NOMEM: return "out of host memory";
    case TSM_E_STATE: return "call out of order";
    default: return "unknown status";
  }
}

extern "C" const char* tsm_category_name(int id) {
  if (id == TSM_CAT_OTHER) return "<other>";
  if (id < 0 || id >= TSM_CAT_NAMED) return "";
  return kNames[id];
}

static void build_lut(uint32_t* lut) {                   // the automaton table of tsm_device.cuh (bit 30 = 'F', bit 31 = newline)
  struct Pat { const char* s; int first; bool ci; };
  static const Pat pats[] = {{"assert", 0, true}, {"EXPECT_", 6, false}, {"class", 13, false}, {"def", 18, false},
                             {"test", 21, true}, {"void", 25, false}, {"{", 29, false}, {"F", 30, false}, {"\n", 31, false}};
  memset(lut, 0, 256 * sizeof(uint32_t));
  for (const Pat& p : pats)
    for (int k = 0; p.s[k]; ++k) {
      const unsigned char c = (unsigned char)p.s[k];
      lut[c] |= 1u << (p.first + k);
      if (p.ci && c >= 'a' && c <= 'z') lut[c - 32] |= 1u << (p.first + k);
    }
}

static void build_lut_b(uint32_t* lut) {                 // Rev-B triggers (tsm_scan_walk.cuh, docs/SPEC.md section 4b)
  struct Pat { const char* s; int first; };
  static const Pat pats[] = {{"_CHECK", 0}, {"TESTEQUAL", 6}, {"FAIL", 15}};
  memset(lut, 0, 256 * sizeof(uint32_t));
  for (const Pat& p : pats)
    for (int k = 0; p.s[k]; ++k) lut[(unsigned char)p.s[k]] |= 1u << (p.first + k);
}

static void build_elut(uint32_t* lut) {                  // operator patterns of SPEC section 6 rule 2
  struct Pat { const char* s; int first; };
  static const Pat pats[] = {{" not ", 0}, {" in ", 5}, {" is not ", 9}, {"True", 17}, {"==", 21}, {"!=", 23},
                             {"<=", 25}, {">=", 27}, {"<", 29}, {">", 30}};
  memset(lut, 0, 256 * sizeof(uint32_t));
  for (const Pat& p : pats)
    for (int k = 0; p.s[k]; ++k) lut[(unsigned char)p.s[k]] |= 1u << (p.first + k);
}

static void free_res_pair(tsm_ctx* c);

extern "C" void tsm_destroy(tsm_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaFree(c->d_arena); cudaFree(c->d_off); cudaFree(c->d_len); cudaFree(c->d_ext); cudaFree(c->d_grp);
  cudaFree(c->d_unit_file); cudaFree(c->d_unit_begin); cudaFree(c->d_zero); cudaFree(c->d_stats);
  c->pool.clear();
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (cudaEvent_t e : c->slab_ev) if (e) cudaEventDestroy(e);
  if (c->ready_ev) cudaEventDestroy(c->ready_ev);
  for (cudaEvent_t e : c->diff_ev) if (e) cudaEventDestroy(e);
  free_res_pair(c);
  cudaFree(c->d_cand); cudaFree(c->d_hev); cudaFree(c->d_aev);
  if (c->h_ctrl) cudaFreeHost(c->h_ctrl);
  if (c->h_diff) cudaFreeHost(c->h_diff);
  for (auto& set : c->ev) for (cudaEvent_t e : set) if (e) cudaEventDestroy(e);
  delete c;
}

extern "C" int tsm_create(tsm_ctx** out, int device, int64_t max_arena_bytes, int32_t max_files,
                          int32_t max_groups, int64_t max_events) {
  if (!out || max_arena_bytes <= 0 || max_arena_bytes >= (1ll << 31) || max_files <= 0 || max_groups <= 0)
    return TSM_E_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    fprintf(stderr, "tosemscan: no usable CUDA device %d (there is no CPU fallback)\n", device);
    return TSM_E_CUDA;
  }
  CU(cudaSetDevice(device));
  tsm_ctx* c = new (std::nothrow) tsm_ctx;
  if (!c) return TSM_E_NOMEM;
  c->device = device;
  {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) { tsm_destroy(c); return TSM_E_CUDA; }
    c->sms = sms;
  }
  c->max_arena = (max_arena_bytes + 127) / 128 * 128;
  c->max_files = max_files;
  c->max_groups = max_groups;
  c->max_events = max_events > 0 ? max_events : c->max_arena / 32 + max_files;
  if (c->max_events > 0xFFFFFFF0ll) c->max_events = 0xFFFFFFF0ll;
  c->unit_cap = (uint32_t)(c->max_arena / CH + max_files);
  int rc = TSM_OK;
  auto A = [&](void** p, size_t bytes) { if (rc == TSM_OK && cudaMalloc(p, bytes ? bytes : 16) != cudaSuccess) rc = TSM_E_CUDA; };
  A((void**)&c->d_arena, (size_t)c->max_arena + 4096);   // slack: bulk copies round sizes up to 16 B
  A((void**)&c->d_off, sizeof(int32_t) * ((size_t)max_files + 1));
  A((void**)&c->d_len, sizeof(int32_t) * (size_t)max_files);
  A((void**)&c->d_ext, (size_t)max_files);
  A((void**)&c->d_grp, sizeof(uint16_t) * (size_t)max_files);
  A((void**)&c->d_unit_file, sizeof(uint32_t) * (size_t)c->unit_cap);
  A((void**)&c->d_unit_begin, sizeof(uint32_t) * (size_t)c->unit_cap);
  {
    const size_t slab_off = 256, counts_off = slab_off + (sizeof(SlabCtl) * tsm_ctx::kMaxSlabs + 255) / 256 * 256;
    A((void**)&c->d_zero, counts_off + sizeof(unsigned long long) * ((size_t)(max_groups + 1) * TSM_K + 4));
    c->d_ctrl = reinterpret_cast<Ctrl*>(c->d_zero);
    c->d_slab = reinterpret_cast<SlabCtl*>(c->d_zero + slab_off);
    c->d_counts = reinterpret_cast<unsigned long long*>(c->d_zero + counts_off);
  }
  if (rc == TSM_OK && cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) rc = TSM_E_CUDA;
  for (cudaEvent_t& e : c->slab_ev) if (rc == TSM_OK && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) rc = TSM_E_CUDA;
  if (rc == TSM_OK && cudaEventCreateWithFlags(&c->ready_ev, cudaEventDisableTiming) != cudaSuccess) rc = TSM_E_CUDA;
  for (cudaEvent_t& e : c->diff_ev) if (rc == TSM_OK && cudaEventCreate(&e) != cudaSuccess) rc = TSM_E_CUDA;
  A((void**)&c->d_stats, sizeof(tsm_file_stat) * (size_t)max_files);
  A((void**)&c->d_cand, sizeof(unsigned long long) * (size_t)c->max_events);
  if (rc == TSM_OK && cudaHostAlloc((void**)&c->h_ctrl, sizeof(Ctrl) + 64, cudaHostAllocDefault) != cudaSuccess) rc = TSM_E_CUDA;
  if (rc == TSM_OK && cudaHostAlloc((void**)&c->h_diff, 256, cudaHostAllocDefault) != cudaSuccess) rc = TSM_E_CUDA;
  for (auto& set : c->ev) for (cudaEvent_t& e : set) if (rc == TSM_OK && cudaEventCreate(&e) != cudaSuccess) rc = TSM_E_CUDA;
  if (rc == TSM_OK) {
    uint32_t lut[256];
    build_lut(lut);
    static const uint8_t slot[TSM_CAT_SLOTS] = TSM_CAT_SLOT_INIT;
    static const uint16_t offs[TSM_CAT_NAMED + 1] = TSM_CAT_OFF_INIT;
    static const char blob[] = TSM_CAT_BLOB_INIT;
    uint32_t elut[256], lutb[256];
    build_elut(elut);
    build_lut_b(lutb);
    if (cudaMemcpyToSymbol(c_lut, lut, sizeof lut) != cudaSuccess ||
        cudaMemcpyToSymbol(c_elut, elut, sizeof elut) != cudaSuccess ||
        cudaMemcpyToSymbol(c_cat_slot, slot, sizeof slot) != cudaSuccess ||
        cudaMemcpyToSymbol(c_cat_off, offs, sizeof offs) != cudaSuccess ||
        cudaMemcpyToSymbol(c_cat_blob, blob, TSM_CAT_BLOB_LEN + 1) != cudaSuccess ||
        cudaMemset(c->d_arena, 0, (size_t)c->max_arena + 4096) != cudaSuccess ||
        cudaMemcpyToSymbol(c_lut_b, lutb, sizeof lutb) != cudaSuccess ||
        cudaFuncSetAttribute(k_scan_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN2_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_scan_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN2_SMEM_B) != cudaSuccess ||
        cudaFuncSetAttribute(k_diff_small<DS1_HCAP, DS1_DCAP, DS1_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(DS1_WARPS * ds_warp_bytes(DS1_HCAP, DS1_DCAP))) != cudaSuccess ||
        cudaFuncSetAttribute(k_diff_small<DS2_HCAP, DS2_DCAP, DS2_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(DS2_WARPS * ds_warp_bytes(DS2_HCAP, DS2_DCAP))) != cudaSuccess ||
        cudaFuncSetAttribute(k_diff_small<DS3_HCAP, DS3_DCAP, DS3_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(DS3_WARPS * ds_warp_bytes(DS3_HCAP, DS3_DCAP))) != cudaSuccess ||
        cudaFuncSetAttribute(k_diff_small<DS4_HCAP, DS4_DCAP, DS4_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(DS4_WARPS * ds_warp_bytes(DS4_HCAP, DS4_DCAP))) != cudaSuccess)
      rc = TSM_E_CUDA;
  }
  if (rc != TSM_OK) {
    fprintf(stderr, "tosemscan: tsm_create failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    tsm_destroy(c);
    return rc;
  }
  *out = c;
  return TSM_OK;
}

// Layout rules of docs/SPEC.md section 1.  The O(1) part, and the per-file part over [f0, f1) (prev_end carries the end of
// the file in front): tsm_scan checks a slab's files while the slab in front of it is on the wire.
static int check_corpus_head(const tsm_ctx* c, const tsm_corpus* k) {
  if (!k || k->n_files < 0 || k->n_groups < 1 || (k->n_files > 0 && (!k->arena || !k->off || !k->len || !k->ext)))
    return TSM_E_ARG;
  if (k->n_files > c->max_files || k->n_groups > c->max_groups) return TSM_E_CAPACITY;
  if (k->n_files == 0) return TSM_OK;
  const int64_t total = k->off[k->n_files];
  if (total < 0 || (total & (TSM_ALIGN - 1))) return TSM_E_LAYOUT;
  if (total > c->max_arena) return TSM_E_CAPACITY;
  return TSM_OK;
}
static int check_files(const tsm_corpus* k, int32_t f0, int32_t f1, int64_t& prev_end) {
  const int64_t total = k->off[k->n_files];
  for (int32_t i = f0; i < f1; ++i) {
    const int64_t o = k->off[i], l = k->len[i];
    if (o < prev_end || (o & (TSM_ALIGN - 1)) || l < 0 || o + l > (int64_t)k->off[i + 1] || (int64_t)k->off[i + 1] > total) return TSM_E_LAYOUT;
    if (k->grp && k->grp[i] >= k->n_groups) return TSM_E_LAYOUT;
    if (k->ext[i] > TSM_EXT_H) return TSM_E_LAYOUT;
    prev_end = o + l;
  }
  return TSM_OK;
}
static int check_corpus(const tsm_ctx* c, const tsm_corpus* k) {
  int rc = check_corpus_head(c, k);
  if (rc != TSM_OK || k->n_files == 0) return rc;
  int64_t prev_end = 0;
  rc = check_files(k, 0, k->n_files, prev_end);
  if (rc == TSM_OK && prev_end > (int64_t)k->off[k->n_files]) rc = TSM_E_LAYOUT;
  return rc;
}

extern "C" int tsm_upload(tsm_ctx* c, const tsm_corpus* k, void* stream) {
  if (!c) return TSM_E_ARG;
  int rc = check_corpus(c, k);
  if (rc != TSM_OK) return rc;
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t n = k->n_files;
  c->n_files = n;
  c->n_groups = k->n_groups;
  c->arena_bytes = n ? k->off[n] : 0;
  if (n) {
    CU(cudaMemcpyAsync(c->d_arena, k->arena, (size_t)c->arena_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->d_off, k->off, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->d_len, k->len, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->d_ext, k->ext, (size_t)n, cudaMemcpyHostToDevice, st));
    if (k->grp) CU(cudaMemcpyAsync(c->d_grp, k->grp, sizeof(uint16_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    else CU(cudaMemsetAsync(c->d_grp, 0, sizeof(uint16_t) * (size_t)n, st));
  }
  c->resident = true;
  c->scanned = false;
  return TSM_OK;
}

static int ensure_event_buffers(tsm_ctx* c, uint32_t flags) {
  if ((flags & TSM_SCAN_ASSERT_EVENTS) && !c->d_aev)
    CU(cudaMalloc((void**)&c->d_aev, sizeof(tsm_assert_event) * (size_t)c->max_events));
  if ((flags & TSM_SCAN_HEADER_EVENTS) && !c->d_hev)
    CU(cudaMalloc((void**)&c->d_hev, sizeof(tsm_header_event) * (size_t)c->max_events));
  return TSM_OK;
}

static ScanParams make_params(const tsm_ctx* c, uint32_t flags) {
  ScanParams p;
  p.arena = c->d_arena; p.off = c->d_off; p.len = c->d_len; p.ext = c->d_ext; p.grp = c->d_grp;
  p.n_files = c->n_files; p.n_groups = c->n_groups;
  p.unit_file = c->d_unit_file; p.unit_begin = c->d_unit_begin; p.unit_cap = c->unit_cap;
  p.ctrl = c->d_ctrl; p.stats = c->d_stats;
  p.cand = c->d_cand; p.cand_cap = (uint32_t)c->max_events;
  p.hev = c->d_hev; p.hev_cap = (uint32_t)c->max_events;
  p.aev = c->d_aev; p.aev_cap = (uint32_t)c->max_events;
  p.counts = c->d_counts; p.flags = flags; p.four = 4; p.cls_last = 1;
  return p;
}

// One scan = [memsets] + per slab (k_plan, k_scan) + k_classify on `st`.  With host != NULL
// the arena slabs are copied on the ctx's copy stream and each slab's kernels wait for its copy only,
// so the H2D of slab s+1 overlaps the scan of slab s (the e2e path); with host == NULL the arena is
// already resident and there is a single slab.
static int launch_scan(tsm_ctx* c, uint32_t flags, cudaStream_t st, const tsm_corpus* host, bool check_files_here = false) {
  int rc = ensure_event_buffers(c, flags);
  if (rc != TSM_OK) return rc;
  ScanParams p = make_params(c, flags);
  const int n = c->n_files;
  c->launches = 0;
  CU(cudaMemsetAsync(c->d_zero, 0, (size_t)(reinterpret_cast<uint8_t*>(c->d_counts) - c->d_zero) +
                     sizeof(unsigned long long) * ((size_t)(c->n_groups + 1) * TSM_K + 4), st));   // ctrl | slab | counts
  if (n) {                                               // (k_plan zeroes the per-file records that k_scan adds into)
    // slab boundaries (file indices): one slab when resident, ~32 MiB of arena each when streaming
    std::vector<int32_t> cut{0};
    if (host) {
      int64_t slab_bytes = 32ll << 20;
      while (c->arena_bytes / slab_bytes + 1 > tsm_ctx::kMaxSlabs) slab_bytes *= 2;
      int64_t next = slab_bytes;
      for (int32_t i = 1; i < n; ++i)
        if ((int64_t)host->off[i] >= next) { cut.push_back(i); next = (int64_t)host->off[i] + slab_bytes; }
      CU(cudaEventRecord(c->ready_ev, st));               // the arena may still be read by earlier work on st
      CU(cudaStreamWaitEvent(c->copy_stream, c->ready_ev, 0));
    }
    cut.push_back(n);
    if ((int)cut.size() - 1 > tsm_ctx::kMaxSlabs) return TSM_E_LAYOUT;   // (only offsets that break the layout rules can cut this often)
    const int es = c->ev_next;
    c->ev_next = (es + 1) % tsm_ctx::kRing;
    fold_events(c, es);                                  // only blocks when 32 scans are in flight
    cudaEvent_t* ev = c->ev[es];
    CU(cudaEventRecord(ev[0], st));
    const int n_slabs = (int)cut.size() - 1;
    const size_t hist = CLS_SMEM_BASE + sizeof(uint32_t) * (c->n_groups <= 16 ? (size_t)c->n_groups * TSM_K : 0);
    if (c->cls_smem != hist) {                           // one resident wave of k_classify (grid-stride inside): measured
      int per_sm = 0;                                    // -7 % on C2 against 8 blocks per SM, equal on C4
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_classify_t<false>, 256, hist) != cudaSuccess || per_sm < 1) per_sm = 4;
      c->cls_grid = c->sms * per_sm;
      c->cls_smem = hist;
    }
    int64_t prev_end = 0;
    for (int s = 0; s < n_slabs; ++s) {
      const int32_t f0 = cut[(size_t)s], f1 = cut[(size_t)s + 1];
      if (host) {
        if (check_files_here && s == 0) {                  // the first slab's files before anything of them is used ...
          const int rc0 = check_files(host, f0, f1, prev_end);
          if (rc0 != TSM_OK) return rc0;
        }
        const size_t b0 = (size_t)host->off[f0], b1 = (size_t)host->off[f1];
        CU(cudaMemcpyAsync(c->d_arena + b0, host->arena + b0, b1 - b0, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(c->slab_ev[s], c->copy_stream));
        CU(cudaStreamWaitEvent(st, c->slab_ev[s], 0));
        if (check_files_here && s + 1 < n_slabs) {         // ... the next slab's while this one is on the wire
          const int rc1 = check_files(host, f1, cut[(size_t)s + 2], prev_end);
          if (rc1 != TSM_OK) { cudaStreamSynchronize(c->copy_stream); cudaStreamSynchronize(st); return rc1; }
        }
      }
      p.slab = c->d_slab + s; p.f_begin = f0; p.f_end = f1;
      // units of earlier slabs are bounded by (arena bytes before f0) / CH + f0
      p.unit_base = (uint32_t)((host ? (int64_t)host->off[f0] : 0) / CH + f0);
      k_plan<<<(f1 - f0 + 255) / 256, 256, 0, st>>>(p);
      CU(cudaGetLastError());
      if (s == 0 && n_slabs == 1) CU(cudaEventRecord(ev[1], st));
      if (flags & TSM_SCAN_REV_B) k_scan_t<true><<<c->sms * SCAN2_CTAS_PER_SM, SCAN2_WARPS * 32, SCAN2_SMEM_B, st>>>(p);
      else k_scan_t<false><<<c->sms * SCAN2_CTAS_PER_SM, SCAN2_WARPS * 32, SCAN2_SMEM, st>>>(p);
      CU(cudaGetLastError());
      if (s + 1 < n_slabs) {                               // streamed scan: this slab's candidates are classified under the next
        p.cls_last = 0;                                    // slab's copy, so that only the last slab's are left behind the last copy
        if (flags & TSM_SCAN_REV_B) k_classify_t<true><<<c->cls_grid, 256, hist, st>>>(p);
        else k_classify_t<false><<<c->cls_grid, 256, hist, st>>>(p);
        CU(cudaGetLastError());
        p.cls_last = 1;
      }
    }
    if (n_slabs > 1) CU(cudaEventRecord(ev[1], st));      // per-kernel split is only meaningful for one slab
    CU(cudaEventRecord(ev[2], st));
    if (flags & TSM_SCAN_REV_B) k_classify_t<true><<<c->cls_grid, 256, hist, st>>>(p);
    else k_classify_t<false><<<c->cls_grid, 256, hist, st>>>(p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ev[3], st));
    CU(cudaEventRecord(ev[4], st));                           // (slot of the former k_totals, now fused into k_classify)
    c->ev_used[es] = (n_slabs == 1);
    c->ev_last = es;
    c->launches = 3 * n_slabs;                             // k_plan + k_scan + k_classify per slab
    CU(cudaGetLastError());
  }
  c->last_flags = flags;
  c->scanned = true;
  return TSM_OK;
}

extern "C" int tsm_scan_resident(tsm_ctx* c, uint32_t flags, void* stream) {
  if (!c) return TSM_E_ARG;
  flags &= TSM_SCAN_ASSERT_EVENTS | TSM_SCAN_HEADER_EVENTS | TSM_SCAN_REV_B;
  if (!c->resident) return TSM_E_STATE;
  CU(cudaSetDevice(c->device));
  return launch_scan(c, flags, (cudaStream_t)stream, nullptr);
}

extern "C" int tsm_device_counts(tsm_ctx* c, void** dptr, int64_t* n_int64) {
  if (!c || !dptr || !n_int64) return TSM_E_ARG;
  if (!c->scanned) return TSM_E_STATE;
  *dptr = c->d_counts;
  *n_int64 = (int64_t)(c->n_groups + 1) * TSM_K + 4;
  return TSM_OK;
}

extern "C" int tsm_last_launch_count(tsm_ctx* c) { return c ? c->launches : 0; }

extern "C" int tsm_last_kernel_ms(tsm_ctx* c, float* ms4) {
  if (!c || !ms4) return TSM_E_ARG;
  if (!c->scanned || c->n_files == 0 || c->ev_last < 0) return TSM_E_STATE;
  CU(cudaSetDevice(c->device));
  cudaEvent_t* ev = c->ev[c->ev_last];
  CU(cudaEventSynchronize(ev[4]));
  for (int i = 0; i < 4; ++i) CU(cudaEventElapsedTime(&ms4[i], ev[i], ev[i + 1]));
  return TSM_OK;
}

extern "C" int tsm_kernel_ms_stats(tsm_ctx* c, double* sum_ms4, int64_t* n_scans, int reset) {
  if (!c) return TSM_E_ARG;
  CU(cudaSetDevice(c->device));
  for (int i = 0; i < tsm_ctx::kRing; ++i) fold_events(c, i);
  if (sum_ms4) for (int k = 0; k < 4; ++k) sum_ms4[k] = c->ms_sum[k];
  if (n_scans) *n_scans = c->ms_n;
  if (reset) { for (double& v : c->ms_sum) v = 0; c->ms_n = 0; }
  return TSM_OK;
}

extern "C" int tsm_download(tsm_ctx* c, tsm_result* r, void* stream) {
  if (!c || !r) return TSM_E_ARG;
  if (!c->scanned) return TSM_E_STATE;
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int n = c->n_files, G = c->n_groups;
  unsigned long long* h_tot = reinterpret_cast<unsigned long long*>(c->h_ctrl + 1);   // pinned tail
  CU(cudaMemcpyAsync(c->h_ctrl, c->d_ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h_tot, c->d_counts + (size_t)(G + 1) * TSM_K, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  if (r->stats && n) CU(cudaMemcpyAsync(r->stats, c->d_stats, sizeof(tsm_file_stat) * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (r->group_counts) CU(cudaMemcpyAsync(r->group_counts, c->d_counts, sizeof(int64_t) * (size_t)G * TSM_K, cudaMemcpyDeviceToHost, st));
  if (r->global_counts) CU(cudaMemcpyAsync(r->global_counts, c->d_counts + (size_t)G * TSM_K, sizeof(int64_t) * TSM_K, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int i = 0; i < 4; ++i) r->totals[i] = (int64_t)h_tot[i];
  r->n_aev = 0; r->n_hev = 0;
  if (c->h_ctrl->overflow) return TSM_E_CAPACITY;
  if ((c->last_flags & TSM_SCAN_ASSERT_EVENTS) && r->aev) {
    const int64_t m = c->h_ctrl->n_aev;
    if (m > r->aev_cap) { r->n_aev = m; return TSM_E_CAPACITY; }
    CU(cudaMemcpyAsync(r->aev, c->d_aev, sizeof(tsm_assert_event) * (size_t)m, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::sort(r->aev, r->aev + m, [](const tsm_assert_event& a, const tsm_assert_event& b) {
      return a.file != b.file ? a.file < b.file : a.line_off < b.line_off; });
    r->n_aev = m;
  } else if (c->last_flags & TSM_SCAN_ASSERT_EVENTS) r->n_aev = c->h_ctrl->n_aev;
  if ((c->last_flags & TSM_SCAN_HEADER_EVENTS) && r->hev) {
    const int64_t m = c->h_ctrl->n_hev;
    if (m > r->hev_cap) { r->n_hev = m; return TSM_E_CAPACITY; }
    CU(cudaMemcpyAsync(r->hev, c->d_hev, sizeof(tsm_header_event) * (size_t)m, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::sort(r->hev, r->hev + m, [](const tsm_header_event& a, const tsm_header_event& b) {
      return a.file != b.file ? a.file < b.file : a.line_off < b.line_off; });
    r->n_hev = m;
  } else if (c->last_flags & TSM_SCAN_HEADER_EVENTS) r->n_hev = c->h_ctrl->n_hev;
  return TSM_OK;
}

extern "C" int tsm_scan(tsm_ctx* c, const tsm_corpus* k, tsm_result* r, uint32_t flags, void* stream) {
  if (!c || !r) return TSM_E_ARG;
  int rc = check_corpus_head(c, k);                       // (the per-file rules are checked slab by slab, under the copies)
  if (rc != TSM_OK) return rc;
  flags &= TSM_SCAN_ASSERT_EVENTS | TSM_SCAN_HEADER_EVENTS | TSM_SCAN_REV_B;
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t n = k->n_files;
  c->n_files = n;
  c->n_groups = k->n_groups;
  c->arena_bytes = n ? k->off[n] : 0;
  if (n) {                                                // the index first (small), the arena slab by slab
    CU(cudaMemcpyAsync(c->d_off, k->off, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->d_len, k->len, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->d_ext, k->ext, (size_t)n, cudaMemcpyHostToDevice, st));
    if (k->grp) CU(cudaMemcpyAsync(c->d_grp, k->grp, sizeof(uint16_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    else CU(cudaMemsetAsync(c->d_grp, 0, sizeof(uint16_t) * (size_t)n, st));
  }
  c->resident = true;
  rc = launch_scan(c, flags, st, k, true);
  if (rc != TSM_OK) { c->resident = false; c->scanned = false; return rc; }
  return tsm_download(c, r, stream);
}

// ------------------------------------------------------------------------------------- S10 reduce
extern "C" int tsm_reduce(tsm_ctx* c, const uint8_t* flags, const int32_t* repo, const int32_t* case_id,
                          int32_t n_rows, int32_t n_flags, int32_t n_repos, int32_t n_cases,
                          int64_t* out, int64_t* cases_per_repo, void* stream) {
  if (!c || n_rows < 0 || n_flags < 0 || n_repos <= 0 || n_cases <= 0 || !out || (n_rows && (!flags || !repo || !case_id)))
    return TSM_E_ARG;
  for (int32_t i = 0; i < n_rows; ++i)
    if (repo[i] < 0 || repo[i] >= n_repos || case_id[i] < 0 || case_id[i] >= n_cases) return TSM_E_ARG;
  CU(cudaSetDevice(c->device));
  PoolScope pool_scope(&c->pool);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t words = ((size_t)n_cases + 31) / 32;
  const size_t nbits = (size_t)(n_flags + 1) * n_repos * words;
  DevBuf b_flags, b_repo, b_case, b_bits, b_out;          // scratch from the ctx's pool (kept between calls)
  SyncGuard guard(st);
  if (!b_flags.alloc((size_t)n_rows * n_flags) || !b_repo.alloc(sizeof(int32_t) * (size_t)n_rows) ||
      !b_case.alloc(sizeof(int32_t) * (size_t)n_rows) || !b_bits.alloc(sizeof(uint32_t) * nbits) ||
      !b_out.alloc(sizeof(unsigned long long) * (size_t)(n_flags + 1) * n_repos))
    return TSM_E_CUDA;
  uint8_t* d_flags = b_flags.as<uint8_t>(); int32_t *d_repo = b_repo.as<int32_t>(), *d_case = b_case.as<int32_t>();
  uint32_t* d_bits = b_bits.as<uint32_t>();
  unsigned long long* d_out = b_out.as<unsigned long long>();
  int rc = TSM_OK;
  std::vector<unsigned long long> h((size_t)(n_flags + 1) * n_repos);
  {
    bool ok = true;
    if (n_rows) {
      ok &= cudaMemcpyAsync(d_flags, flags, (size_t)n_rows * n_flags, cudaMemcpyHostToDevice, st) == cudaSuccess;
      ok &= cudaMemcpyAsync(d_repo, repo, sizeof(int32_t) * (size_t)n_rows, cudaMemcpyHostToDevice, st) == cudaSuccess;
      ok &= cudaMemcpyAsync(d_case, case_id, sizeof(int32_t) * (size_t)n_rows, cudaMemcpyHostToDevice, st) == cudaSuccess;
    }
    ok &= cudaMemsetAsync(d_bits, 0, sizeof(uint32_t) * nbits, st) == cudaSuccess;
    ok &= cudaMemsetAsync(d_out, 0, sizeof(unsigned long long) * h.size(), st) == cudaSuccess;
    if (ok) ok &= launch_reduce(d_flags, d_repo, d_case, n_rows, n_flags, n_repos, n_cases, d_bits, d_out, st) == 0;
    ok &= cudaMemcpyAsync(h.data(), d_out, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaStreamSynchronize(st) == cudaSuccess;
    if (!ok) rc = TSM_E_CUDA;
  }
  if (rc != TSM_OK) return rc;
  // row 0 of the device table = "any row" (cases per repo), rows 1.. = the flags
  for (int32_t r = 0; r < n_repos; ++r) if (cases_per_repo) cases_per_repo[r] = (int64_t)h[r];
  for (int32_t f = 0; f < n_flags; ++f)
    for (int32_t r = 0; r < n_repos; ++r) out[(size_t)f * n_repos + r] = (int64_t)h[(size_t)(f + 1) * n_repos + r];
  c->launches = 2;
  return TSM_OK;
}

// ------------------------------------------------------------------------------------- host helpers
extern "C" void* tsm_host_alloc(int64_t bytes) {
  void* p = nullptr;
  if (bytes <= 0) return nullptr;
  if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
extern "C" void tsm_host_free(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------- S8 diff
namespace {
struct HostSide {                                         // device image of one side of the pairs + its line records
  int32_t n = 0; size_t ab = 0; uint32_t unit_cap = 0;
  DevBuf arena, off, len, ext, line_base, line_end, line_hash, line_flag;
  DevBuf unit_file, unit_begin, cnt, unit_first, bsum, zero, stats, unit_lines, unit_out, unit_line_base, s_hash, s_end, s_flag;
  std::vector<unsigned long long> base;                   // host copy of line_base (only when asked for)
  uint32_t n_units = 0;                                   // (file, chunk) work units: sum of ceil(len / 4 KiB)
  Ctrl hc{};                                              // read back behind the scan: capacity flags, lines written
  unsigned long long total = 0;                           // lines of the side
  Ctrl* pin_hc = nullptr; unsigned long long* pin_total = nullptr;   // where they land (pinned, in the ctx)
  DiffSide d{};
  int launches = 0;
  void drop_staging() { s_hash.reset(); s_end.reset(); s_flag.reset(); }
};

// Exclusive scan of n u32 counts into n + 1 u64 (tsm_lines_kernels.cuh); bsum holds n / 1024 + 2 u64.
static void xscan(const uint32_t* in, uint32_t n, unsigned long long* bsum, unsigned long long* out, cudaStream_t st) {
  const uint32_t nb = (n + XS_TILE - 1) / XS_TILE;
  if (nb == 0) { cudaMemsetAsync(out, 0, sizeof(unsigned long long), st); return; }
  k_xscan_sums<<<nb, 256, 0, st>>>(in, n, bsum);
  k_xscan_top<<<1, 256, 0, st>>>(bsum, nb);
  k_xscan_apply<<<nb, 256, 0, st>>>(in, n, bsum, out);
}

int side_upload(const tsm_corpus* k, HostSide& h, cudaStream_t st) {
  const int32_t n = k->n_files;
  const size_t ab = (size_t)k->off[n];
  h.n = n; h.ab = ab;
  h.unit_cap = (uint32_t)(ab / CH + (size_t)n + 1);
  const uint32_t unit_cap = h.unit_cap;
  unsigned long long nu = 0;
  for (int32_t i = 0; i < n; ++i) nu += ((unsigned long long)(uint32_t)k->len[i] + CH - 1) / CH;
  if (nu > unit_cap) return TSM_E_LAYOUT;
  h.n_units = (uint32_t)nu;
  if (!h.arena.alloc(ab + 4096) || !h.off.alloc(sizeof(int32_t) * ((size_t)n + 1)) || !h.len.alloc(sizeof(int32_t) * (size_t)n) ||
      !h.ext.alloc((size_t)n) || !h.unit_file.alloc(sizeof(uint32_t) * unit_cap) || !h.unit_begin.alloc(sizeof(uint32_t) * unit_cap) ||
      !h.cnt.alloc(sizeof(uint32_t) * (size_t)n) || !h.unit_first.alloc(sizeof(unsigned long long) * ((size_t)n + 1)) ||
      !h.bsum.alloc(sizeof(unsigned long long) * (unit_cap / XS_TILE + 4)) || !h.zero.alloc(256 + sizeof(SlabCtl)) ||
      !h.stats.alloc(sizeof(tsm_file_stat) * (size_t)n) || !h.unit_lines.alloc(sizeof(uint32_t) * unit_cap) ||
      !h.unit_out.alloc(sizeof(uint32_t) * unit_cap) || !h.unit_line_base.alloc(sizeof(unsigned long long) * ((size_t)unit_cap + 1)) ||
      !h.line_base.alloc(sizeof(unsigned long long) * ((size_t)n + 1)))
    return TSM_E_CUDA;
  CU(cudaMemsetAsync(h.arena.as<uint8_t>() + ab, 0, 4096, st));
  CU(cudaMemcpyAsync(h.arena.p, k->arena, ab, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h.off.p, k->off, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h.len.p, k->len, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
  if (k->ext) CU(cudaMemcpyAsync(h.ext.p, k->ext, (size_t)n, cudaMemcpyHostToDevice, st));
  else CU(cudaMemsetAsync(h.ext.p, 0, (size_t)n, st));
  return TSM_OK;
}

// Line records in file order (docs/SPEC.md sections 2-4) of `ns` uploaded sides (the two sides of the revision pairs, or one
// corpus): per side line_base[n+1], and per line its hash, its end and whether it is an assertion line.  One pass of
// k_scan over the source (TSM_SCAN_LINE_HASHES: every chunk writes the records of its own lines into a region of the
// staging arrays), one exclusive scan of the lines per unit, one gather.  The kernels of all sides are queued before the
// host looks at anything: ONE synchronisation (capacity flags + line totals) per call instead of four per side.  The
// staging arrays are sized for 8-byte lines; a side with more lines than that is scanned a second time with the exact
// size (the first pass counted them).  scan_ms adds the device time of the k_scan launches (CUDA events on st).
// host_base: also copy line_base to the host (HostSide::base).
static int side_scan_pass(tsm_ctx* c, HostSide& h, ScanParams& p, size_t cap, int side, cudaStream_t st) {
  const int ev = side ? 6 : 0;
  h.pin_hc = reinterpret_cast<Ctrl*>(c->h_diff + 32 * side);            // pinned: the copies below do not stall the host
  h.pin_total = reinterpret_cast<unsigned long long*>(c->h_diff + 64 + 8 * side);
  const int32_t n = h.n;
  if (cap > 0xFFFFFFF0ull) return TSM_E_CAPACITY;
  if (!h.s_hash.alloc(sizeof(unsigned long long) * cap) || !h.s_end.alloc(sizeof(uint32_t) * cap) || !h.s_flag.alloc(cap)) return TSM_E_CUDA;
  p.lh_hash = h.s_hash.as<unsigned long long>(); p.lh_end = h.s_end.as<uint32_t>(); p.lh_flag = h.s_flag.as<uint8_t>();
  p.lh_cap = (uint32_t)cap;
  CU(cudaMemsetAsync(h.zero.p, 0, 256 + sizeof(SlabCtl), st));
  k_plan_det<<<(n + 1 + 255) / 256, 256, 0, st>>>(p, h.unit_first.as<unsigned long long>());
  CU(cudaEventRecord(c->diff_ev[ev], st));
  k_scan_t<false><<<c->sms * SCAN2_CTAS_PER_SM, SCAN2_WARPS * 32, SCAN2_SMEM, st>>>(p);
  CU(cudaEventRecord(c->diff_ev[ev + 1], st));
  CU(cudaGetLastError());
  // lines per unit -> first line of every unit (the unit count is known on the host: units are (file, chunk) in order)
  xscan(p.unit_lines, h.n_units, h.bsum.as<unsigned long long>(), h.unit_line_base.as<unsigned long long>(), st);
  CU(cudaMemcpyAsync(h.pin_hc, p.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h.pin_total, h.unit_line_base.as<unsigned long long>() + h.n_units, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  h.launches += 6;
  return TSM_OK;
}

int sides_records(tsm_ctx* c, HostSide* const* sides, int ns, cudaStream_t st, float* scan_ms, bool host_base) {
  if (ns < 1 || ns > 2) return TSM_E_ARG;
  ScanParams ps[2];
  size_t caps[2];
  for (int i = 0; i < ns; ++i) {
    HostSide& h = *sides[i];
    const int32_t n = h.n;
    ScanParams& p = ps[i];
    p = ScanParams{};
    p.arena = h.arena.as<uint8_t>(); p.off = h.off.as<int32_t>(); p.len = h.len.as<int32_t>(); p.ext = h.ext.as<uint8_t>();
    p.grp = nullptr; p.n_files = n; p.n_groups = 1;
    p.unit_file = h.unit_file.as<uint32_t>(); p.unit_begin = h.unit_begin.as<uint32_t>(); p.unit_cap = h.unit_cap;
    p.ctrl = reinterpret_cast<Ctrl*>(h.zero.as<uint8_t>()); p.slab = reinterpret_cast<SlabCtl*>(h.zero.as<uint8_t>() + 256);
    p.f_begin = 0; p.f_end = n; p.unit_base = 0;
    p.stats = h.stats.as<tsm_file_stat>();
    p.cand = nullptr; p.cand_cap = 0; p.hev = nullptr; p.hev_cap = 0; p.aev = nullptr; p.aev_cap = 0; p.counts = nullptr;
    p.flags = TSM_SCAN_LINE_HASHES; p.four = 4;
    p.unit_lines = h.unit_lines.as<uint32_t>(); p.unit_out = h.unit_out.as<uint32_t>();
    // units in (file, chunk) order
    k_file_units<<<(n + 255) / 256, 256, 0, st>>>(p.len, (uint32_t)n, h.cnt.as<uint32_t>());
    xscan(h.cnt.as<uint32_t>(), (uint32_t)n, h.bsum.as<unsigned long long>(), h.unit_first.as<unsigned long long>(), st);
    caps[i] = h.ab / 8 + 2 * (size_t)h.unit_cap + 64;
    const int rc = side_scan_pass(c, h, p, caps[i], i, st);
    if (rc != TSM_OK) return rc;
  }
  CU(cudaStreamSynchronize(st));
  for (int i = 0; i < ns; ++i) {
    HostSide& h = *sides[i];
    const int ev = i ? 6 : 0;
    h.hc = *h.pin_hc; h.total = *h.pin_total;
    if (scan_ms) { float ms = 0; if (cudaEventElapsedTime(&ms, c->diff_ev[ev], c->diff_ev[ev + 1]) == cudaSuccess) *scan_ms += ms; }
    if (h.hc.overflow) return TSM_E_CAPACITY;
    if (h.hc.lh_overflow) {                                // more lines than the staging arrays hold: once more, exact size
      const int rc = side_scan_pass(c, h, ps[i], (size_t)h.hc.n_lh + 64, i, st);
      if (rc != TSM_OK) return rc;
      CU(cudaStreamSynchronize(st));
      h.hc = *h.pin_hc; h.total = *h.pin_total;
      if (scan_ms) { float ms = 0; if (cudaEventElapsedTime(&ms, c->diff_ev[ev], c->diff_ev[ev + 1]) == cudaSuccess) *scan_ms += ms; }
      if (h.hc.overflow || h.hc.lh_overflow) return TSM_E_CAPACITY;
    }
  }
  for (int i = 0; i < ns; ++i) {
    HostSide& h = *sides[i];
    const int32_t n = h.n;
    const ScanParams& p = ps[i];
    const unsigned long long total = h.total;
    if (!h.line_end.alloc(sizeof(uint32_t) * (size_t)total) || !h.line_hash.alloc(sizeof(unsigned long long) * (size_t)total) ||
        !h.line_flag.alloc((size_t)total))
      return TSM_E_CUDA;
    if (h.n_units)
      k_gather_lines<<<(h.n_units * 32 + 255) / 256, 256, 0, st>>>(p.unit_lines, p.unit_out, h.unit_line_base.as<unsigned long long>(), h.n_units,
                                                                     p.lh_hash, p.lh_end, p.lh_flag, h.line_hash.as<unsigned long long>(),
                                                                     h.line_end.as<uint32_t>(), h.line_flag.as<uint8_t>());
    k_line_base<<<(n + 1 + 255) / 256, 256, 0, st>>>(h.unit_first.as<unsigned long long>(), h.unit_line_base.as<unsigned long long>(),
                                                       (uint32_t)n, h.line_base.as<unsigned long long>());
    CU(cudaGetLastError());
    h.launches += 5;
    h.base.clear();
    if (host_base) {
      h.base.assign((size_t)n + 1, 0);
      CU(cudaMemcpyAsync(h.base.data(), h.line_base.p, sizeof(unsigned long long) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st));
    }
    h.d.arena = h.arena.as<uint8_t>(); h.d.off = h.off.as<int32_t>(); h.d.len = h.len.as<int32_t>();
    h.d.n_lines = nullptr;
    h.d.line_base = h.line_base.as<unsigned long long>();
    h.d.line_end = h.line_end.as<uint32_t>();
    h.d.line_hash = h.line_hash.as<unsigned long long>();
    h.d.ext = h.ext.as<uint8_t>();
    h.d.line_flag = h.line_flag.as<uint8_t>();
  }
  if (host_base) {
    CU(cudaStreamSynchronize(st));
    for (int i = 0; i < ns; ++i) sides[i]->drop_staging();
  }                                                        // (else the caller drops them behind its next synchronisation:
  return TSM_OK;                                           //  the gather queued above still reads them)
}

int side_records(tsm_ctx* c, HostSide& h, cudaStream_t st, float* scan_ms) {
  HostSide* one[1] = {&h};
  return sides_records(c, one, 1, st, scan_ms, true);
}
}  // namespace

struct HostSidePair { HostSide A, B; int32_t n = 0; bool have_records = false; };

static void free_res_pair(tsm_ctx* c) {
  if (!c->res_pair) return;
  PoolScope pool_scope(&c->pool);                          // the buffers go back to the ctx's pool
  delete c->res_pair;
  c->res_pair = nullptr;
}

static int check_pair_layout(const tsm_corpus* olds, const tsm_corpus* news, bool with_ext) {
  const int32_t n = olds->n_files;
  for (const tsm_corpus* k : {olds, news}) {              // same layout rules as the scan (SPEC section 1)
    if (!k->arena || !k->off || !k->len) return TSM_E_ARG;
    for (int32_t i = 0; i < n; ++i)
      if (k->off[i] < 0 || (k->off[i] & (TSM_ALIGN - 1)) || k->len[i] < 0 || (int64_t)k->off[i] + k->len[i] > k->off[i + 1] ||
          (with_ext && k->ext && k->ext[i] > TSM_EXT_H))
        return TSM_E_LAYOUT;
  }
  return TSM_OK;
}

// The diff proper over two sides whose line records exist.  k_diff_small finishes the common pairs (distance at most
// 127 lines, middle of at most 4 096 lines) start to finish - search in registers, rows of V and backtrack in shared
// memory - in four sizes (512 lines / D <= 31 at 32 pairs per SM, 1 024 / 63 at 12, 4 096 / 63 at 5, 4 096 / 127 at 3), each
// fed on the device by the list the size before it leaves.  What all of them leave over (the `todo` list, normally empty) goes through k_myers (edit distance, V in global scratch) and, for `detail`,
// k_myers_trace (rows of V in global memory sized from those distances, then the canonical script: hunks, changed
// assertion lines).  A pair whose distance D needs more than TSM_DIFF_TRACE_MAX_INTS trace entries ((D+1)(D+2)/2) is
// not traced: it is reported as ONE hunk (add / del / mod by its counts) with added_assert = removed_assert = -1
// (tosemscan.h).  diff_ms[1] = k_diff_small, diff_ms[2] = the two kernels of the left-over pairs.
static int diff_core(tsm_ctx* c, HostSide& A, HostSide& B, int32_t n, int64_t* added, int64_t* removed,
                     tsm_diff_detail* detail, cudaStream_t st) {
  static_assert(sizeof(long long) == sizeof(int64_t), "int64");
  DevBuf d_add, d_rem, d_detail, d_todo1, d_todo2, d_todo3, d_todo, d_ntodo;
  if (!d_add.alloc(sizeof(long long) * (size_t)n) || !d_rem.alloc(sizeof(long long) * (size_t)n) ||
      !d_todo1.alloc(sizeof(int32_t) * (size_t)n) || !d_todo2.alloc(sizeof(int32_t) * (size_t)n) || !d_todo3.alloc(sizeof(int32_t) * (size_t)n) || !d_todo.alloc(sizeof(int32_t) * (size_t)n) ||
      !d_ntodo.alloc(64) || (detail && !d_detail.alloc(sizeof(tsm_diff_detail) * (size_t)n)))
    return TSM_E_CUDA;
  constexpr uint32_t kSmem1 = DS1_WARPS * ds_warp_bytes(DS1_HCAP, DS1_DCAP), kSmem2 = DS2_WARPS * ds_warp_bytes(DS2_HCAP, DS2_DCAP),
                     kSmem3 = DS3_WARPS * ds_warp_bytes(DS3_HCAP, DS3_DCAP), kSmem4 = DS4_WARPS * ds_warp_bytes(DS4_HCAP, DS4_DCAP);
  static_assert(kSmem1 * 4 + 4 * 1024 <= 233472 && kSmem2 * 6 + 6 * 1024 <= 233472 && kSmem3 * 5 + 5 * 1024 <= 233472 &&
                kSmem4 * 3 + 3 * 1024 <= 233472, "pairs per SM");
  // (the kernels' dynamic shared memory limits are raised per device in tsm_create)
  CU(cudaMemsetAsync(d_ntodo.p, 0, 64, st));
  CU(cudaMemsetAsync(d_add.p, 0, sizeof(long long) * (size_t)n, st));     // (the first copy back covers every pair, also the ones
  CU(cudaMemsetAsync(d_rem.p, 0, sizeof(long long) * (size_t)n, st));     //  the four sizes leave to k_myers / k_myers_trace)
  if (detail) CU(cudaMemsetAsync(d_detail.p, 0, sizeof(tsm_diff_detail) * (size_t)n, st));
  uint32_t* cnt = d_ntodo.as<uint32_t>();                  // [0..3] pairs each size left over, [4..7] the sizes' work counters
  const uint8_t* fa = detail ? A.d.line_flag : nullptr;
  const uint8_t* fb = detail ? B.d.line_flag : nullptr;
  tsm_diff_detail* d_det = detail ? d_detail.as<tsm_diff_detail>() : nullptr;
  long long* da = d_add.as<long long>();
  long long* dr = d_rem.as<long long>();
  CU(cudaEventRecord(c->diff_ev[2], st));
  k_diff_small<DS1_HCAP, DS1_DCAP, DS1_WARPS><<<std::min((n + DS1_WARPS - 1) / DS1_WARPS, c->sms * 4), DS1_WARPS * 32, kSmem1, st>>>(
      A.d.line_hash, A.d.line_base, fa, B.d.line_hash, B.d.line_base, fb, nullptr, nullptr, n, cnt + 4, da, dr, d_det, d_todo1.as<int32_t>(), cnt + 0);
  k_diff_small<DS2_HCAP, DS2_DCAP, DS2_WARPS><<<std::min((n + DS2_WARPS - 1) / DS2_WARPS, c->sms * 6), DS2_WARPS * 32, kSmem2, st>>>(
      A.d.line_hash, A.d.line_base, fa, B.d.line_hash, B.d.line_base, fb, d_todo1.as<int32_t>(), cnt + 0, n, cnt + 5, da, dr, d_det, d_todo2.as<int32_t>(), cnt + 1);
  k_diff_small<DS3_HCAP, DS3_DCAP, DS3_WARPS><<<std::min(n, c->sms * 5), DS3_WARPS * 32, kSmem3, st>>>(
      A.d.line_hash, A.d.line_base, fa, B.d.line_hash, B.d.line_base, fb, d_todo2.as<int32_t>(), cnt + 1, n, cnt + 6, da, dr, d_det, d_todo3.as<int32_t>(), cnt + 2);
  k_diff_small<DS4_HCAP, DS4_DCAP, DS4_WARPS><<<std::min(n, c->sms * 3), DS4_WARPS * 32, kSmem4, st>>>(
      A.d.line_hash, A.d.line_base, fa, B.d.line_hash, B.d.line_base, fb, d_todo3.as<int32_t>(), cnt + 2, n, cnt + 7, da, dr, d_det, d_todo.as<int32_t>(), cnt + 3);
  CU(cudaEventRecord(c->diff_ev[3], st));
  CU(cudaGetLastError());
  uint32_t* pin_nt = reinterpret_cast<uint32_t*>(c->h_diff + 80);
  CU(cudaMemcpyAsync(pin_nt, cnt + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(added, d_add.p, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(removed, d_rem.p, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (detail) CU(cudaMemcpyAsync(detail, d_detail.p, sizeof(tsm_diff_detail) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  A.drop_staging(); B.drop_staging();
  const uint32_t nt = *pin_nt;
  { float ms = 0; if (cudaEventElapsedTime(&ms, c->diff_ev[2], c->diff_ev[3]) == cudaSuccess) c->diff_ms[1] = ms; }
  c->launches = A.launches + B.launches + 4;
  c->diff_ms[2] = 0;
  if (nt == 0) return TSM_OK;
  // ---- the left-over pairs: long middles, far-apart revisions
  std::vector<int32_t> todo(nt);
  CU(cudaMemcpyAsync(todo.data(), d_todo.p, sizeof(int32_t) * (size_t)nt, cudaMemcpyDeviceToHost, st));
  for (HostSide* h : {&A, &B})
    if (h->base.empty()) {
      h->base.assign((size_t)n + 1, 0);
      CU(cudaMemcpyAsync(h->base.data(), h->line_base.p, sizeof(unsigned long long) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st));
    }
  CU(cudaStreamSynchronize(st));
  std::vector<unsigned long long> vbase((size_t)nt + 1, 0);
  for (uint32_t s = 0; s < nt; ++s) {
    const size_t i = (size_t)todo[s];
    vbase[(size_t)s + 1] = vbase[s] + 2 * ((A.base[i + 1] - A.base[i]) + (B.base[i + 1] - B.base[i])) + 3;
  }
  DevBuf d_vbase, d_v;
  if (!d_vbase.alloc(sizeof(unsigned long long) * ((size_t)nt + 1)) || !d_v.alloc(sizeof(int32_t) * (size_t)vbase[nt])) return TSM_E_CUDA;
  CU(cudaMemcpyAsync(d_vbase.p, vbase.data(), sizeof(unsigned long long) * ((size_t)nt + 1), cudaMemcpyHostToDevice, st));
  CU(cudaEventRecord(c->diff_ev[4], st));
  k_myers<<<(nt * 32 + 127) / 128, 128, 0, st>>>(A.d.line_hash, A.d.line_base, B.d.line_hash, B.d.line_base, (int32_t)nt,
                                                d_v.as<int32_t>(), d_vbase.as<unsigned long long>(),
                                                d_add.as<long long>(), d_rem.as<long long>(), d_todo.as<int32_t>());
  CU(cudaEventRecord(c->diff_ev[5], st));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(added, d_add.p, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(removed, d_rem.p, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  { float ms = 0; if (cudaEventElapsedTime(&ms, c->diff_ev[4], c->diff_ev[5]) == cudaSuccess) c->diff_ms[2] += ms; }
  c->launches++;
  if (!detail) return TSM_OK;
  // ---- their hunks: second search with the rows of V kept; rows sized from the distances just computed,
  //      pairs processed in batches of at most 2^28 trace ints (1 GiB)
  d_v.reset();                                             // the first search's scratch is no longer needed
  DevBuf d_tbase;
  if (!d_tbase.alloc(sizeof(unsigned long long) * ((size_t)nt + 1))) return TSM_E_CUDA;
  std::vector<unsigned long long> tbase((size_t)nt + 1, 0);
  std::vector<int32_t> untraced;
  const unsigned long long kBatch = TSM_DIFF_TRACE_MAX_INTS;
  uint32_t p0 = 0;
  while (p0 < nt) {
    uint32_t p1 = p0;
    unsigned long long tot = 0;
    while (p1 < nt) {
      const int32_t i = todo[p1];
      const unsigned long long D = (unsigned long long)(added[i] + removed[i]);
      unsigned long long need = (D + 1) * (D + 2) / 2;
      if (need > kBatch) { untraced.push_back(i); need = 1; }       // k_myers_trace skips it (row table of one int)
      if (p1 > p0 && tot + need > kBatch) break;
      tbase[p1] = tot;
      tot += need;
      ++p1;
    }
    DevBuf d_trace;
    if (!d_trace.alloc(sizeof(int32_t) * (size_t)tot)) return TSM_E_CUDA;
    CU(cudaMemcpyAsync(d_tbase.as<unsigned long long>() + p0, tbase.data() + p0, sizeof(unsigned long long) * (size_t)(p1 - p0),
                       cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(c->diff_ev[4], st));
    k_myers_trace<<<((p1 - p0) * 32 + 127) / 128, 128, 0, st>>>(
        A.d.line_hash, A.d.line_base, A.d.line_flag, B.d.line_hash, B.d.line_base, B.d.line_flag, (int32_t)p0, (int32_t)(p1 - p0),
        d_trace.as<int32_t>(), d_tbase.as<unsigned long long>(), d_add.as<long long>(), d_rem.as<long long>(),
        (long long)TSM_DIFF_TRACE_MAX_D, d_detail.as<tsm_diff_detail>(), d_todo.as<int32_t>());
    CU(cudaEventRecord(c->diff_ev[5], st));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    { float ms = 0; if (cudaEventElapsedTime(&ms, c->diff_ev[4], c->diff_ev[5]) == cudaSuccess) c->diff_ms[2] += ms; }
    c->launches++;
    p0 = p1;
  }
  CU(cudaMemcpyAsync(detail, d_detail.p, sizeof(tsm_diff_detail) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int32_t i : untraced) {                              // too far apart to trace: one hunk, assertion counts unknown
    tsm_diff_detail d{0, 0, 0, -1, -1};
    if (added[i] && removed[i]) d.hunks_mod = 1; else if (added[i]) d.hunks_add = 1; else d.hunks_del = 1;
    detail[i] = d;
  }
  return TSM_OK;
}

extern "C" int tsm_diff_pairs_detail(tsm_ctx* c, const tsm_corpus* olds, const tsm_corpus* news,
                                     int64_t* added, int64_t* removed, tsm_diff_detail* detail, void* stream) {
  if (!c || !olds || !news || !added || !removed || olds->n_files != news->n_files) return TSM_E_ARG;
  const int32_t n = olds->n_files;
  if (n == 0) return TSM_OK;
  int rc = check_pair_layout(olds, news, detail != nullptr);
  if (rc != TSM_OK) return rc;
  CU(cudaSetDevice(c->device));
  PoolScope pool_scope(&c->pool);
  cudaStream_t st = (cudaStream_t)stream;
  HostSide A, B;
  SyncGuard guard(st);                                     // no buffer goes back to the pool while work on st may still use it
  c->diff_ms[0] = 0;
  rc = side_upload(olds, A, st);
  if (rc == TSM_OK) rc = side_upload(news, B, st);
  HostSide* both[2] = {&A, &B};
  if (rc == TSM_OK) rc = sides_records(c, both, 2, st, &c->diff_ms[0], false);
  if (rc != TSM_OK) return rc;
  return diff_core(c, A, B, n, added, removed, detail, st);
}

// Resident variant (what bench.py's `value` times for config C5): the two sides go to HBM once ...
extern "C" int tsm_diff_upload(tsm_ctx* c, const tsm_corpus* olds, const tsm_corpus* news, void* stream) {
  if (!c || !olds || !news || olds->n_files != news->n_files || olds->n_files <= 0) return TSM_E_ARG;
  int rc = check_pair_layout(olds, news, true);
  if (rc != TSM_OK) return rc;
  CU(cudaSetDevice(c->device));
  free_res_pair(c);
  PoolScope pool_scope(&c->pool);
  cudaStream_t st = (cudaStream_t)stream;
  SyncGuard guard(st);
  c->res_pair = new (std::nothrow) HostSidePair;
  if (!c->res_pair) return TSM_E_NOMEM;
  c->res_pair->n = olds->n_files;
  rc = side_upload(olds, c->res_pair->A, st);
  if (rc == TSM_OK) rc = side_upload(news, c->res_pair->B, st);
  if (rc == TSM_OK) rc = cudaStreamSynchronize(st) == cudaSuccess ? TSM_OK : TSM_E_CUDA;
  if (rc != TSM_OK) { delete c->res_pair; c->res_pair = nullptr; }
  return rc;
}

// ... and every call runs the kernels over them: k_scan over both sides (line records), k_myers, k_myers_trace.
extern "C" int tsm_diff_resident(tsm_ctx* c, int64_t* added, int64_t* removed, tsm_diff_detail* detail, void* stream) {
  if (!c || !added || !removed) return TSM_E_ARG;
  if (!c->res_pair) return TSM_E_STATE;
  CU(cudaSetDevice(c->device));
  PoolScope pool_scope(&c->pool);
  cudaStream_t st = (cudaStream_t)stream;
  SyncGuard guard(st);
  HostSidePair& P = *c->res_pair;
  c->diff_ms[0] = 0;
  P.A.launches = P.B.launches = 0;
  HostSide* both[2] = {&P.A, &P.B};
  int rc = sides_records(c, both, 2, st, &c->diff_ms[0], false);
  if (rc != TSM_OK) return rc;
  return diff_core(c, P.A, P.B, P.n, added, removed, detail, st);
}

extern "C" int tsm_diff_last_ms(tsm_ctx* c, float* ms3) {
  if (!c || !ms3) return TSM_E_ARG;
  for (int i = 0; i < 3; ++i) ms3[i] = c->diff_ms[i];
  return TSM_OK;
}

extern "C" int tsm_diff_pairs(tsm_ctx* c, const tsm_corpus* olds, const tsm_corpus* news,
                              int64_t* added, int64_t* removed, void* stream) {
  return tsm_diff_pairs_detail(c, olds, news, added, removed, nullptr, stream);
}

// ------------------------------------------------------------------------------------- S9 line / n-gram hashes
extern "C" int tsm_line_hashes(tsm_ctx* c, const tsm_corpus* k, int64_t* line_base, uint64_t* line_hash, uint32_t* line_end,
                               uint8_t* line_flag, int64_t cap, int64_t* n_lines, int32_t ngram_n, uint64_t* ngram_hash,
                               void* stream) {
  if (!c || !k || !line_base || !n_lines || cap < 0 || k->n_files < 0 || ngram_n < 0 || (ngram_hash && ngram_n < 1)) return TSM_E_ARG;
  const int32_t n = k->n_files;
  *n_lines = 0;
  line_base[0] = 0;
  if (n == 0) return TSM_OK;
  if (!k->arena || !k->off || !k->len) return TSM_E_ARG;
  for (int32_t i = 0; i < n; ++i)
    if (k->off[i] < 0 || (k->off[i] & (TSM_ALIGN - 1)) || k->len[i] < 0 || (int64_t)k->off[i] + k->len[i] > k->off[i + 1] ||
        (k->ext && k->ext[i] > TSM_EXT_H))
      return TSM_E_LAYOUT;
  CU(cudaSetDevice(c->device));
  PoolScope pool_scope(&c->pool);
  cudaStream_t st = (cudaStream_t)stream;
  HostSide S;
  SyncGuard guard(st);
  int rc = side_upload(k, S, st);
  if (rc == TSM_OK) rc = side_records(c, S, st, nullptr);
  if (rc != TSM_OK) return rc;
  const unsigned long long total = S.base[(size_t)n];
  for (int32_t i = 0; i <= n; ++i) line_base[i] = (int64_t)S.base[(size_t)i];
  *n_lines = (int64_t)total;
  c->launches = S.launches;
  if ((unsigned long long)cap < total) return TSM_E_CAPACITY;   // line_base / n_lines are filled: allocate and call again
  if (total == 0) return TSM_OK;
  if (line_hash) CU(cudaMemcpyAsync(line_hash, S.d.line_hash, sizeof(uint64_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
  if (line_end) CU(cudaMemcpyAsync(line_end, S.d.line_end, sizeof(uint32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
  if (line_flag) CU(cudaMemcpyAsync(line_flag, S.d.line_flag, (size_t)total, cudaMemcpyDeviceToHost, st));
  if (ngram_hash) {
    DevBuf d_ng;
    if (!d_ng.alloc(sizeof(unsigned long long) * (size_t)total)) { cudaStreamSynchronize(st); return TSM_E_CUDA; }
    k_ngrams<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(S.d.line_hash, S.d.line_base, (uint32_t)n, total, (uint32_t)ngram_n,
                                                               d_ng.as<unsigned long long>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(ngram_hash, d_ng.p, sizeof(uint64_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    c->launches++;
  }
  CU(cudaStreamSynchronize(st));
  return TSM_OK;
}

// ------------------------------------------------------------------------------------- SPEC section 10 statements
extern "C" int tsm_statements(tsm_ctx* c, const tsm_corpus* k, int64_t* line_base, uint32_t* line_end,
                              uint8_t* line_kind, int64_t cap, int64_t* n_lines, void* stream) {
  if (!c || !k || !line_base || !n_lines || cap < 0 || k->n_files < 0) return TSM_E_ARG;
  const int32_t n = k->n_files;
  *n_lines = 0;
  line_base[0] = 0;
  if (n == 0) return TSM_OK;
  if (!k->arena || !k->off || !k->len) return TSM_E_ARG;
  for (int32_t i = 0; i < n; ++i)
    if (k->off[i] < 0 || (k->off[i] & (TSM_ALIGN - 1)) || k->len[i] < 0 || (int64_t)k->off[i] + k->len[i] > k->off[i + 1])
      return TSM_E_LAYOUT;
  CU(cudaSetDevice(c->device));
  PoolScope pool_scope(&c->pool);
  cudaStream_t st = (cudaStream_t)stream;
  HostSide S;
  SyncGuard guard(st);
  int rc = side_upload(k, S, st);                         // line records from one pass of the scan
  if (rc == TSM_OK) rc = side_records(c, S, st, nullptr);
  if (rc != TSM_OK) return rc;
  const unsigned long long total = S.base[(size_t)n];
  for (int32_t i = 0; i <= n; ++i) line_base[i] = (int64_t)S.base[(size_t)i];
  *n_lines = (int64_t)total;
  if ((unsigned long long)cap < total) return TSM_E_CAPACITY;   // line_base / n_lines are filled: allocate and call again
  if (total == 0) return TSM_OK;
  if (!line_end || !line_kind) return TSM_E_ARG;
  DevBuf d_delta, d_kind;
  if (!d_delta.alloc(sizeof(int32_t) * (size_t)total) || !d_kind.alloc((size_t)total)) return TSM_E_CUDA;
  k_line_parens<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(S.d, n, total, d_delta.as<int32_t>(), d_kind.as<uint8_t>());
  k_stmt_kinds<<<(n * 32 + 127) / 128, 128, 0, st>>>(S.d, n, d_delta.as<int32_t>(), d_kind.as<uint8_t>());
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(line_end, S.d.line_end, sizeof(uint32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(line_kind, d_kind.p, (size_t)total, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  c->launches = S.launches + 2;
  return TSM_OK;
}
