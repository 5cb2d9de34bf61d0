// tosem_scan_cli.cpp - `tosem-scan`, the C++ host driver of the corpus-scan loop.
//
// The reference package ships no entry point to stay compatible with; what it
// ships are the loop's OUTPUT SCHEMAS, and this driver writes exactly those:
//   raw rows      fileName,extension,test_name,method,statement,counts,category
//                 (Important-files/ML-Testing-v1.xlsx!apollo_tests:R1)
//   per-file      Id,FileName,total assert,assertion   ("32:assertEqual, 15:assertIn, ...")
//                 (selection/completed-labels/Release-Meta-tpot.csv:1-2)
//   RQ tables     RQs/RQ3/tests_strategy_rq32.csv, RQs/RQ4/tests_methods_v2.csv from
//                 RQs/taxonomy_test2.csv
//   churn         cloc,added,removed  (Important-files/ML-Testing-v1.xlsx!projects:R1)
// All CSVs are CRLF, UTF-8, RFC-4180 quoted, like every CSV the package ships.
//
// Host work only: tree walk + test-file selection (S0), extension / test_name tags (S1, S2), packing
// into the pinned byte arena, method strings (S3, host side of SPEC section 5), CSV.  Every byte of the scan
// itself goes through libtosemscan.so (sm_100a kernels); there is no CPU fallback.
//
//   tosem-scan scan   <project-root>... [--rows F] [--summary F] [--gpus N] [--all-files] [--batch-bytes N]
//   tosem-scan reduce <taxonomy.csv> [--strategy F] [--methods F] [--properties F] [--correlate F] [--correlate-tex F] [--correlate-counts F] [--correlate-merged F]
//   tosem-scan diff   <old-root> <new-root> [--out F]
//   tosem-scan body   <project-root>... [--out F]
//   tosem-scan releases <snapshot-root>=<tag>... [--out F]   |   releases --git <repository> [<revision>...] [--out F]
//   tosem-scan history <git-repository> [--rev R] [--max-commits N] [--all-files] [--dry-run] [--out F]
#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <filesystem>
#include <fstream>
#include <future>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>

#include "../../include/tosemscan.h"
#include "git_store.hpp"

namespace fs = std::filesystem;

static bool die(const std::string& m) { fprintf(stderr, "tosem-scan: %s\n", m.c_str()); exit(2); return false; }
static void ck(int rc, const char* what) { if (rc != TSM_OK) die(std::string(what) + ": " + tsm_strerror(rc)); }

static std::string lower(std::string s) { for (char& c : s) if (c >= 'A' && c <= 'Z') c = (char)(c + 32); return s; }
static bool is_w(unsigned char c) { return c == 0x20 || c == 0x09 || c == 0x0D || c == 0x0B || c == 0x0C; }

// ---------------------------------------------------------------------------------- CSV (RFC 4180, CRLF)
static std::string csv_cell(const std::string& s) {
  if (s.find_first_of(",\"\r\n") == std::string::npos) return s;
  std::string o = "\"";
  for (char c : s) { if (c == '"') o += '"'; o += c; }
  return o + "\"";
}
static void csv_row(std::ostream& os, const std::vector<std::string>& cells) {
  for (size_t i = 0; i < cells.size(); ++i) { if (i) os << ','; os << csv_cell(cells[i]); }
  os << "\r\n";
}
static std::vector<std::vector<std::string>> csv_read(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) die("cannot open " + path);
  std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  if (data.compare(0, 3, "\xEF\xBB\xBF") == 0) data.erase(0, 3);
  std::vector<std::vector<std::string>> rows;
  std::vector<std::string> row;
  std::string cell;
  bool q = false, any = false;
  for (size_t i = 0; i < data.size(); ++i) {
    const char c = data[i];
    if (q) {
      if (c == '"') { if (i + 1 < data.size() && data[i + 1] == '"') { cell += '"'; ++i; } else q = false; }
      else cell += c;
    } else if (c == '"') { q = true; any = true; }
    else if (c == ',') { row.push_back(cell); cell.clear(); any = true; }
    else if (c == '\n' || c == '\r') {
      if (c == '\r' && i + 1 < data.size() && data[i + 1] == '\n') ++i;
      if (any || !cell.empty()) { row.push_back(cell); rows.push_back(row); }
      row.clear(); cell.clear(); any = false;
    } else { cell += c; any = true; }
  }
  if (any || !cell.empty()) { row.push_back(cell); rows.push_back(row); }
  return rows;
}

// ---------------------------------------------------------------------------------- S0-S2: walk and tag
struct FileEntry {
  std::string rel, abs; int ext; int grp; int64_t size;
  std::shared_ptr<const std::vector<uint8_t>> blob;       // set instead of `abs` when the bytes come from a git object store
};

static int ext_tag(const std::string& rel) {               // S1
  const size_t d = rel.rfind('.');
  if (d == std::string::npos || rel.find('/', d) != std::string::npos) return TSM_EXT_OTHER;
  const std::string e = rel.substr(d + 1);
  if (e == "py") return TSM_EXT_PY;
  if (e == "cc") return TSM_EXT_CC;
  if (e == "cpp") return TSM_EXT_CPP;
  if (e == "java") return TSM_EXT_JAVA;
  if (e == "c") return TSM_EXT_C;
  if (e == "h") return TSM_EXT_H;
  return TSM_EXT_OTHER;
}
static const char* ext_name(int t) { static const char* n[] = {"", "py", "cc", "cpp", "java", "c", "h"}; return n[t]; }

static std::string test_name_tag(const std::string& rel, bool fixture) {   // S2 (mock / Module: parity unpinned, omitted)
  std::string t;
  if (rel.rfind("external/", 0) == 0) t = "external";
  else if (rel.find("integration") != std::string::npos) t = "integration";
  else if (rel.find("regression") != std::string::npos) t = "regression";
  else if (rel.find("swarming") != std::string::npos) t = "swarming";
  else t = "unit_test";
  bool proto = false, smoke = false;
  size_t p = 0;
  while (p < rel.size()) {
    size_t q = rel.find('/', p);
    if (q == std::string::npos) break;                     // last component is the file name
    const std::string comp = rel.substr(p, q - p);
    if (comp.rfind("protobuf-", 0) == 0) proto = true;
    if (comp == "smoke") smoke = true;
    p = q + 1;
  }
  if (proto) t += ", Protocol Buffers";
  if (smoke) t += ", smoke";
  if (fixture) t += ", Fixture";
  return t;
}

static void walk(const std::string& root, int grp, bool all_files, std::vector<FileEntry>& out) {
  std::vector<fs::path> paths;
  for (auto it = fs::recursive_directory_iterator(root, fs::directory_options::skip_permission_denied);
       it != fs::recursive_directory_iterator(); ++it)
    if (it->is_regular_file() && !it->is_symlink()) paths.push_back(it->path());
  std::sort(paths.begin(), paths.end());
  for (const fs::path& p : paths) {
    const std::string rel = fs::relative(p, root).generic_string();
    const int ext = ext_tag(rel);
    if (!all_files) {
      if (lower(rel).find("test") == std::string::npos) continue;          // S0: path contains `test`
      if (ext == TSM_EXT_OTHER) continue;                                   // S1: no rows for other extensions
    }
    out.push_back({rel, p.string(), ext, grp, (int64_t)fs::file_size(p), nullptr});
  }
}

// ---------------------------------------------------------------------------------- S3: method strings
static std::string method_string(int ext, const uint8_t* line, uint32_t len) {   // docs/SPEC.md section 5
  uint32_t b = 0, e = len;
  while (b < e && is_w(line[b])) ++b;
  while (e > b && is_w(line[e - 1])) --e;
  auto sw = [&](uint32_t i, const char* pat) { const size_t m = strlen(pat); return i + m <= e && memcmp(line + i, pat, m) == 0; };
  std::string o;
  if (ext == TSM_EXT_PY) {
    uint32_t i = b;
    if (sw(i, "class")) i += 5;
    while (i < e) {
      if (sw(i, "def")) { i += 3; continue; }
      if (!is_w(line[i])) o += (char)line[i];
      ++i;
    }
    if (!o.empty() && o.back() == ':') o.pop_back();
  } else if (ext == TSM_EXT_JAVA) {
    static const char* const words[] = {"public", "private", "protected", "static", "void", "class"};
    uint32_t i = b;
    while (i < e) {
      bool hit = false;
      for (const char* w : words) if (sw(i, w)) { i += (uint32_t)strlen(w); hit = true; break; }
      if (hit) continue;
      if (!is_w(line[i])) o += (char)line[i];
      ++i;
    }
  } else {
    uint32_t t = b;
    while (t < e && line[t] != ')') ++t;
    uint32_t s = b, u = t;
    while (s < t && (is_w(line[s]) || line[s] == '{')) ++s;
    while (u > s && (is_w(line[u - 1]) || line[u - 1] == '{')) --u;
    for (uint32_t i = s; i < u; ++i) if (line[i] != '{') o += (char)line[i];
  }
  return o;
}

// ---------------------------------------------------------------------------------- scan
struct Batch {                                             // one packed arena (<= ~1 GiB) of files of one GPU's share
  std::vector<uint32_t> idx;                               // indices into the walk's file list, ascending
  std::vector<int32_t> off, len;
  std::vector<uint8_t> ext;
  std::vector<uint16_t> grp;
  uint8_t* arena = nullptr;
  int64_t bytes = 0;
  size_t count() const { return idx.size(); }
};

static void load_batch(const std::vector<FileEntry>& files, Batch& b) {
  const size_t n = b.count();
  b.len.resize(n); b.off.resize(n + 1); b.ext.resize(n); b.grp.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const FileEntry& f = files[b.idx[i]];
    b.len[i] = (int32_t)f.size; b.ext[i] = (uint8_t)f.ext; b.grp[i] = (uint16_t)f.grp;
  }
  b.bytes = tsm_layout(b.len.data(), (int32_t)n, b.off.data());
  if (b.bytes < 0) die("batch does not fit an int32-indexed arena");
  b.arena = (uint8_t*)tsm_host_alloc(std::max<int64_t>(b.bytes, 128));
  if (!b.arena) die("pinned arena allocation failed (no CUDA device? there is no CPU fallback)");
  memset(b.arena, 0, (size_t)std::max<int64_t>(b.bytes, 128));
  // the reads of a batch run on a few host threads (a source tree is many small files: latency-bound)
  const unsigned nt = std::max(1u, std::min({std::thread::hardware_concurrency(), 32u, (unsigned)((n + 63) / 64)}));
  std::atomic<long> bad{-1};
  auto reader = [&](unsigned t) {
    for (size_t i = t; i < n && bad.load(std::memory_order_relaxed) < 0; i += nt) {
      if (files[b.idx[i]].blob) { if (b.len[i]) memcpy(b.arena + b.off[i], files[b.idx[i]].blob->data(), (size_t)b.len[i]); continue; }
      const int fd = open(files[b.idx[i]].abs.c_str(), O_RDONLY);
      int64_t got = 0;
      while (fd >= 0 && got < b.len[i]) {
        const ssize_t r = read(fd, b.arena + b.off[i] + got, (size_t)(b.len[i] - got));
        if (r <= 0) break;
        got += r;
      }
      if (fd >= 0) close(fd);
      if (got != b.len[i]) bad.store((long)i);
    }
  };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; ++t) th.emplace_back(reader, t);
  reader(0);
  for (std::thread& x : th) x.join();
  if (bad.load() >= 0) die("short read: " + files[b.idx[(size_t)bad.load()]].abs);
}

static void cu_ck(cudaError_t e, const char* what) { if (e != cudaSuccess) die(std::string(what) + ": " + cudaGetErrorString(e)); }
static void nccl_ck(ncclResult_t r, const char* what) { if (r != ncclSuccess) die(std::string(what) + ": " + ncclGetErrorString(r)); }

// The raw rows (fileName, extension, test_name, method, statement, counts, category: ML-Testing-v1.xlsx!apollo_tests:R1) and
// the summary row (Id, FileName, total assert, assertion: Release-Meta-tpot.csv:1-2) of one file, from its events.
static void render_file(const FileEntry& f, int64_t id, const uint8_t* base, int32_t size, uint32_t slot, const tsm_file_stat& st,
                        const std::vector<tsm_assert_event>& aev, size_t& ai, const std::vector<tsm_header_event>& hev, size_t& hi,
                        bool want_rows, bool want_sum, std::string& rows_txt, std::string& sum_txt) {
  struct Row { int64_t hdr; std::string stmt; int cat; std::string catname; int64_t count; bool fixture; std::string method; };
  std::vector<Row> rows;
  std::map<std::pair<int64_t, std::string>, size_t> index;
  std::map<std::string, int64_t> hist; std::vector<std::string> hist_order;
  int64_t cur_hdr = -1; bool cur_fix = false; std::string cur_method = "xxxx";
  while (ai < aev.size() && aev[ai].file == slot) {
    const tsm_assert_event& ev = aev[ai++];
    while (hi < hev.size() && hev[hi].file == slot && hev[hi].line_off <= ev.line_off) {   // governing header
      const tsm_header_event& h = hev[hi++];
      cur_hdr = h.line_off; cur_fix = (h.kind & 1u) != 0;
      cur_method = method_string(f.ext, base + h.line_off, h.line_len);
    }
    // the statement may be longer than the 16-bit event field: re-derive its end on the host if saturated
    uint32_t sl = ev.stmt_len;
    if (sl == 65535) { const uint8_t* p = base + ev.stmt_off; uint32_t e = 0, last = 0; while (ev.stmt_off + e < (uint32_t)size && p[e] != '\n' && p[e] != '(') { if (!is_w(p[e])) last = e + 1; ++e; } sl = last; }
    std::string stmt((const char*)base + ev.stmt_off, sl);
    std::string cat = ev.cat == 127 ? std::string((const char*)base + ev.ident_off, ev.ident_len) : std::string(tsm_category_name(ev.cat));
    auto key = std::make_pair(cur_hdr, stmt);
    auto it = index.find(key);
    if (it == index.end()) { index[key] = rows.size(); rows.push_back({cur_hdr, stmt, ev.cat, cat, 1, cur_fix, cur_method}); }
    else rows[it->second].count++;
    if (!hist.count(cat)) hist_order.push_back(cat);
    hist[cat]++;
  }
  while (hi < hev.size() && hev[hi].file == slot) ++hi;
  if (want_rows && !rows.empty()) {
    std::ostringstream os;
    for (const Row& r : rows)
      csv_row(os, {f.rel, ext_name(f.ext), test_name_tag(f.rel, r.fixture), r.method, r.stmt, std::to_string(r.count), r.catname});
    rows_txt = os.str();
  }
  if (want_sum) {
    std::stable_sort(hist_order.begin(), hist_order.end(), [&](const std::string& x, const std::string& y) { return hist[x] > hist[y]; });
    std::string a;
    for (const std::string& c : hist_order) { if (!a.empty()) a += ", "; a += std::to_string(hist[c]) + ":" + c; }
    std::ostringstream os;
    csv_row(os, {std::to_string(id), f.rel, std::to_string(st.n_assert), a});
    sum_txt = os.str();
  }
}

static int cmd_scan(const std::vector<std::string>& roots, const std::string& rows_path, const std::string& summary_path,
                    int gpus, bool all_files, int64_t batch_bytes, bool rev_b) {
  std::vector<FileEntry> files;
  for (size_t g = 0; g < roots.size(); ++g) walk(roots[g], (int)g, all_files, files);
  const int n_groups = (int)std::max<size_t>(roots.size(), 1);
  fprintf(stderr, "tosem-scan: %zu files selected under %zu root(s)\n", files.size(), roots.size());
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) die("no CUDA device (there is no CPU fallback)");
  gpus = std::max(1, std::min(gpus, ndev));
  // ---- shares of the GPUs: files sorted by size, descending, each to the GPU with the least bytes
  //      so far (LPT), so that every GPU gets the same byte total whatever the size law is; then, per GPU, batches of at most
  //      --batch-bytes of arena (and 1M files) in walk order
  std::vector<std::vector<Batch>> share((size_t)gpus);
  {
    std::vector<uint32_t> order(files.size());
    for (size_t i = 0; i < files.size(); ++i) {
      if (files[i].size >= (1ll << 30)) die("file larger than 1 GiB: " + files[i].abs);
      order[i] = (uint32_t)i;
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return files[x].size > files[y].size; });
    std::vector<std::vector<uint32_t>> mine((size_t)gpus);
    std::vector<int64_t> load((size_t)gpus, 0);            // greedy LPT: the next largest file goes to the lightest GPU
    for (uint32_t i : order) {
      size_t g = 0;
      for (size_t k = 1; k < (size_t)gpus; ++k) if (load[k] < load[g]) g = k;
      mine[g].push_back(i);
      load[g] += files[i].size + 128;
    }
    for (int g = 0; g < gpus; ++g) {
      std::sort(mine[(size_t)g].begin(), mine[(size_t)g].end());
      int64_t cur = 0;
      share[(size_t)g].emplace_back();
      for (uint32_t i : mine[(size_t)g]) {
        const int64_t padded = (files[i].size + 127) / 128 * 128;
        Batch* b = &share[(size_t)g].back();
        if (b->count() && (cur + padded > batch_bytes || b->count() >= (1u << 20))) { share[(size_t)g].emplace_back(); b = &share[(size_t)g].back(); cur = 0; }
        b->idx.push_back(i);
        cur += padded;
      }
      if (share[(size_t)g].back().count() == 0) share[(size_t)g].pop_back();
    }
  }
  // one host thread per GPU; one ncclAllReduce of the count table at the end
  std::vector<ncclComm_t> comms(gpus);
  std::vector<int> devs(gpus);
  for (int i = 0; i < gpus; ++i) devs[i] = i;
  if (gpus > 1) {
    // stdout carries the aggregate CSV: whatever NCCL prints while initialising (its version banner
    // under NCCL_DEBUG=VERSION) is routed to stderr
    fflush(stdout);
    const int saved = dup(1);
    dup2(2, 1);
    const ncclResult_t nr = ncclCommInitAll(comms.data(), gpus, devs.data());
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
    nccl_ck(nr, "ncclCommInitAll");
  }
  const size_t table = (size_t)(n_groups + 1) * TSM_NUM_CATEGORIES + 4;
  std::vector<std::vector<int64_t>> totals(gpus, std::vector<int64_t>(table, 0));
  const bool want_rows = !rows_path.empty(), want_sum = !summary_path.empty();
  std::vector<std::string> rows_txt(want_rows ? files.size() : 0), sum_txt(want_sum ? files.size() : 0);   // per file, written in walk order at the end
  auto worker = [&](int g) {
    std::vector<Batch>& batches = share[(size_t)g];
    cu_ck(cudaSetDevice(g), "cudaSetDevice");
    cudaStream_t st;
    cu_ck(cudaStreamCreate(&st), "cudaStreamCreate");
    int64_t max_arena = 1 << 20; int32_t max_files = 16;
    for (const Batch& b : batches) {
      int64_t bytes = 0;
      for (uint32_t i : b.idx) bytes += (files[i].size + 127) / 128 * 128;
      max_arena = std::max(max_arena, bytes + 4096);
      max_files = std::max<int32_t>(max_files, (int32_t)b.count());
    }
    tsm_ctx* ctx = nullptr;
    ck(tsm_create(&ctx, g, max_arena, max_files, std::max(n_groups, 1), 0), "tsm_create");
    int64_t* d_acc = nullptr;                               // this GPU's count table, input and output of the allreduce
    cu_ck(cudaMalloc((void**)&d_acc, table * sizeof(int64_t)), "cudaMalloc");
    // host pipeline: while the GPU scans batch b (tsm_scan overlaps its H2D slabs with the kernels), a background
    // task already reads the files of the next batch into its pinned arena.  At most two arenas per GPU are alive:
    // a batch's rows are rendered and its arena and events are freed as soon as its scan returns.
    std::future<void> next_load;
    if (!batches.empty()) next_load = std::async(std::launch::async, [&files, &batches, g] { cudaSetDevice(g); load_batch(files, batches[0]); });
    std::vector<tsm_file_stat> stats;
    std::vector<tsm_assert_event> aev;
    std::vector<tsm_header_event> hev;
    std::vector<int64_t> group_counts, h;
    for (size_t b = 0; b < batches.size(); ++b) {
      Batch& B = batches[b];
      next_load.get();
      if (b + 1 < batches.size())
        next_load = std::async(std::launch::async, [&files, &batches, b, g] { cudaSetDevice(g); load_batch(files, batches[b + 1]); });
      tsm_corpus c{B.arena, B.off.data(), B.len.data(), B.ext.data(), B.grp.data(), (int32_t)B.count(), n_groups};
      stats.resize(B.count());
      group_counts.assign((size_t)n_groups * TSM_NUM_CATEGORIES, 0);
      const int64_t cap = std::max<int64_t>(B.bytes / 8 + 1024, 1024);
      aev.resize((size_t)cap); hev.resize((size_t)cap);
      tsm_result r{};
      r.stats = stats.data(); r.group_counts = group_counts.data();
      r.aev = aev.data(); r.aev_cap = cap; r.hev = hev.data(); r.hev_cap = cap;
      ck(tsm_scan(ctx, &c, &r, TSM_SCAN_ASSERT_EVENTS | TSM_SCAN_HEADER_EVENTS | (rev_b ? TSM_SCAN_REV_B : 0u), st), "tsm_scan");
      aev.resize((size_t)r.n_aev); hev.resize((size_t)r.n_hev);
      void* dptr = nullptr; int64_t n64 = 0;
      ck(tsm_device_counts(ctx, &dptr, &n64), "tsm_device_counts");
      h.resize((size_t)n64);
      cu_ck(cudaMemcpyAsync(h.data(), dptr, (size_t)n64 * sizeof(int64_t), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
      cu_ck(cudaStreamSynchronize(st), "cudaStreamSynchronize");
      for (size_t i = 0; i < (size_t)n64 && i < table; ++i) totals[g][i] += h[i];
      if (want_rows || want_sum) {
        size_t ai = 0, hi = 0;
        static std::string none;
        for (size_t i = 0; i < B.count(); ++i) {
          const uint32_t fi = B.idx[i];
          render_file(files[fi], (int64_t)fi + 1, B.arena + B.off[i], B.len[i], (uint32_t)i, stats[i], aev, ai, hev, hi, want_rows, want_sum,
                      want_rows ? rows_txt[fi] : none, want_sum ? sum_txt[fi] : none);
        }
      }
      tsm_host_free(B.arena);
      B.arena = nullptr;
      std::vector<int32_t>().swap(B.off); std::vector<int32_t>().swap(B.len);
    }
    cu_ck(cudaMemcpyAsync(d_acc, totals[g].data(), table * sizeof(int64_t), cudaMemcpyHostToDevice, st), "cudaMemcpyAsync");
    if (gpus > 1)                                           // the single collective of the path
      nccl_ck(ncclAllReduce(d_acc, d_acc, table, ncclInt64, ncclSum, comms[g], st), "ncclAllReduce");
    cu_ck(cudaMemcpyAsync(totals[g].data(), d_acc, table * sizeof(int64_t), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
    cu_ck(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    cudaFree(d_acc);
    tsm_destroy(ctx);
    cudaStreamDestroy(st);
  };
  if (gpus == 1) worker(0);
  else {
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; ++g) th.emplace_back(worker, g);
    for (auto& t : th) t.join();
    for (int g = 0; g < gpus; ++g) ncclCommDestroy(comms[g]);
  }
  // ---- rows + summary, in walk order
  if (want_rows) {
    std::ofstream os(rows_path, std::ios::binary);
    csv_row(os, {"fileName", "extension", "test_name", "method", "statement", "counts", "category"});
    for (const std::string& t : rows_txt) os << t;
  }
  if (want_sum) {
    std::ofstream os(summary_path, std::ios::binary);
    csv_row(os, {"Id", "FileName", "total assert", "assertion"});
    for (const std::string& t : sum_txt) os << t;
  }
  // ---- the aggregate table (global counts after the allreduce) to stdout
  const std::vector<int64_t>& T = totals[0];
  printf("category,count\r\n");
  for (int k = 0; k < TSM_NUM_CATEGORIES; ++k) {
    const int64_t v = T[(size_t)n_groups * TSM_NUM_CATEGORIES + k];
    if (v) printf("%s,%lld\r\n", k == 0 ? "" : tsm_category_name(k), (long long)v);
  }
  const size_t tot = (size_t)(n_groups + 1) * TSM_NUM_CATEGORIES;
  int64_t share_min = -1, share_max = 0;
  for (int g = 0; g < gpus; ++g) {
    int64_t bsum = 0;
    for (const Batch& b : share[(size_t)g]) for (uint32_t i : b.idx) bsum += files[i].size;
    share_min = share_min < 0 ? bsum : std::min(share_min, bsum); share_max = std::max(share_max, bsum);
  }
  fprintf(stderr, "tosem-scan: lines=%lld assertion_lines=%lld headers=%lld fixture_headers=%lld on %d GPU(s), shares %lld..%lld bytes\n",
          (long long)T[tot], (long long)T[tot + 1], (long long)T[tot + 2], (long long)T[tot + 3], gpus, (long long)std::max<int64_t>(share_min, 0), (long long)share_max);
  return 0;
}

// ---------------------------------------------------------------------------------- reduce (S10)
struct FlagDef { const char* name; const char* col; const char* val; const char* col2; };
// the column mapping of tools/make_golden.py: 171 / 171 cells of tests_strategy_rq32.csv reproduce.  `val` may list
// several values separated by '|': the error rows merge Error_Type values (the merge sets were recovered by
// exhaustive search over the value subsets against the nine shipped per-repository cells, tools/make_golden.py)
static const FlagDef kStrategy[] = {
    {"status_analysis", "status_test", "1", nullptr}, {"value_error", "Error_Type", "ValueError", nullptr},
    {"runtime_error", "Error_Type", "RuntimeError|Exception|NotImplementedError|StopIteration|TimeOut|Timeout|TimeoutError|Warning|nullptr", nullptr}, {"memory_error", "Error_Type", "MemoryError", nullptr},
    {"type_error", "Error_Type", "TypeError", nullptr}, {"import_error", "Error_Type", "ImportError", nullptr},
    {"key_error", "Error_Type", "KeyError", nullptr}, {"AssertionError", "Error_Type", "AssertionError|SyntaxError", nullptr},
    {"FileError", "Error_Type", "FileError|SchemaError", nullptr}, {"NotImplementedError", "Error_Type", "NotImplementedError", nullptr},
    {"negative_test", "negative_test", "1", nullptr}, {"logical_condition", "logical_statement", "1", "logical_expression"},
    {"Null_pointer", "null_pointer", "1", nullptr}, {"value_range", "value_range", "1", nullptr},
    {"absolute_relative_tolerence", "Approximation_Type", "absolute_relative_tolerence", nullptr},
    {"error_bounding", "Approximation_Type", "error_bounding", nullptr},
    {"rounding_tolence", "Approximation_Type", "rounding_tolence", nullptr},
    {"instance_check", "checks_type", "instance_check", nullptr}, {"sub_set_checks", "checks_type", "sub_set_checks", nullptr}};
static const FlagDef kMethods[] = {
    {"regression", "regression", nullptr, nullptr}, {"integration", "Integration", nullptr, nullptr},
    {"end_to_end", "end_to_end", nullptr, nullptr}, {"sanity", "sanity", nullptr, nullptr},
    {"mock_test", "mock_test", nullptr, nullptr}, {"periodic_validation", "periodic_validation", nullptr, nullptr},
    {"example_test", "example_test", nullptr, nullptr}, {"static_inspection", "static_inspection_test", nullptr, nullptr},
    {"robustness_test", "roboustness", nullptr, nullptr}, {"experimental", "Experimental_benchmark_test", nullptr, nullptr},
    {"api_test", "API", nullptr, nullptr}, {"threat", "ThreadTest", nullptr, nullptr}, {"blob", "blob_performance", nullptr, nullptr}};

// RQ3 property table (RQs/RQ3/tests_prop_rq3.csv): a case has a property when the `Data` or the `Model` label of
// one of its rows is in the property's label set.  The sets are not written down in the package; 17 of the 21 were
// recovered by search against the nine shipped per-repository cells (exact, tools/make_golden.py), the other four
// (Consistency, Features Importance, Concurrency, Anomaly) use the label of the same name.
struct PropDef { const char* name; const char* labels; };
static const PropDef kProperties[] = {
    {"Consistency", "Consistency"}, {"Data Distribution", "Distribution"},
    {"Data Validity", "Validity|Data Error|Data Error and Validity"}, {"Completeness", "Completeness"},
    {"Correctness", "Correctness|Accuracy & Precision|Statistical Evidence/ explainability"}, {"Robustness", "Robustness"},
    {"Efficiency", "Time behaviour|Resource Usage|Training Efficiency"},
    {"Data Relation", "Relation & Association|Closeness|Missing Data|Data Differencing|Data Quality"},
    {"Scalability", "Scalability"}, {"Features Importance", "Feature Importance"},
    {"Data Restoration and Recoverability", "Recoverability|Data Restoration"},
    {"Concurrency and Parallelism", "Parallel Processing|parallel"}, {"Uncertainty", "uncertainty"}, {"Anomaly", "Anomaly"},
    {"Data Migration Loss and Corruption", "Data Loss"}, {"Bias and Fairness", "Model Bias"},
    {"Security and Privacy", "Security|Data Encapsulation"}, {"Data Uniqueness", "Uniqueness"},
    {"Data Timeliness", "Timeliness"}, {"Data Integration Integrity", "Validate data integration and integrity"},
    {"Compatibility and Portability", "Compatibility"}};

// RQ3 strategy x property table (RQs/RQ3/tests_correlate_rq3.csv): 20 rows x 21 columns, one cell = the share of a
// repository's cases that have BOTH the strategy flag and the property, for the nine repositories in the order below.
// Row predicates are single taxonomy values (the strategy table above merges Error_Type values, this one does not);
// `decision` reads logical_statement and `logical_condition` reads logical_expression - recovered against the shipped
// cells, 394 of 420 bit-identical (tools/make_golden.py, tests/golden/ledger.json G3).
struct CorrRow { const char* name; const char* col; const char* val; };
static const CorrRow kCorrRows[] = {
    {"rounding_tolence", "Approximation_Type", "rounding_tolence"}, {"instance_check", "checks_type", "instance_check"},
    {"MemoryError", "Error_Type", "MemoryError"}, {"negative_test", "negative_test", "1"},
    {"status_analysis", "status_test", "1"}, {"value_range_analysis", "value_range", "1"},
    {"sub_set_checks", "checks_type", "sub_set_checks"}, {"ValueError", "Error_Type", "ValueError"},
    {"decision", "logical_statement", "1"}, {"error_bounding", "Approximation_Type", "error_bounding"},
    {"Null_pointer", "null_pointer", "1"}, {"boundary", "boundary", "1"},
    {"absolute_relative_tolerence", "Approximation_Type", "absolute_relative_tolerence"},
    {"ImportError", "Error_Type", "ImportError"}, {"pseaudo_oracle", "Pseaudo_Oracle", "1"},
    {"RuntimeError", "Error_Type", "RuntimeError"}, {"logical_condition", "logical_expression", "1"},
    {"TypeError", "Error_Type", "TypeError"}, {"KeyError", "Error_Type", "KeyError"},
    {"NotImplementedError", "Error_Type", "NotImplementedError"}};
struct CorrCol { const char* name; const char* prop; };        // column header of the shipped table -> name in kProperties
static const CorrCol kCorrCols[] = {
    {"Distribution", "Data Distribution"}, {"Validity", "Data Validity"}, {"Consistency", "Consistency"},
    {"Completeness", "Completeness"}, {"Correctness", "Correctness"}, {"Robustness", "Robustness"},
    {"Efficiency", "Efficiency"}, {"Relation", "Data Relation"}, {"Scalability", "Scalability"},
    {"Feature Importance", "Features Importance"}, {"Restoration", "Data Restoration and Recoverability"},
    {"Concurrency", "Concurrency and Parallelism"}, {"uncertainty", "Uncertainty"}, {"Anomaly", "Anomaly"},
    {"Data Loss", "Data Migration Loss and Corruption"}, {"Bias", "Bias and Fairness"},
    {"Security", "Security and Privacy"}, {"Uniqueness", "Data Uniqueness"}, {"Timeliness", "Data Timeliness"},
    {"integration", "Data Integration Integrity"}, {"Compatibility", "Compatibility and Portability"}};

// Four more one-row tables in the same layout for the MERGED rows of the strategy table
// (RQs/RQ3/tests_correlate_{FileError,RuntimeError,assertion,logical}.csv): row name -> row of kStrategy.
struct MergedRow { const char* name; const char* strategy; };
static const MergedRow kMergedRows[] = {{"FileError", "FileError"}, {"RuntimeError", "runtime_error"},
                                        {"AssertionError", "AssertionError"}, {"logical", "logical_condition"}};

static bool one_of(const std::string& have, const std::string& want) {   // `want` = '|'-separated values
  for (size_t a = 0; a <= want.size();) {
    const size_t b = std::min(want.find('|', a), want.size());
    if (have == want.substr(a, b - a)) return true;
    a = b + 1;
  }
  return false;
}

static std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && (s[a] == ' ' || s[a] == '\t')) ++a;
  while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t')) --b;
  return s.substr(a, b - a);
}
static std::string fmt_num(double v, int dec) {            // shipped cells drop trailing zeros ("0", "43.771")
  char buf[64];
  snprintf(buf, sizeof buf, "%.*f", dec, v);
  std::string s = buf;
  if (s.find('.') != std::string::npos) { while (!s.empty() && s.back() == '0') s.pop_back(); if (!s.empty() && s.back() == '.') s.pop_back(); }
  return s.empty() ? "0" : s;
}
static std::string fmt_pyfloat2(double v) {                 // repr(round(v, 2)) of the shipped correlate cells: "0.0", "1.22", "12.2"
  char buf[64];
  snprintf(buf, sizeof buf, "%.2f", v);
  std::string s = buf;
  while (s.size() > 1 && s.back() == '0' && s[s.size() - 2] != '.') s.pop_back();
  return s;
}
static double round_to(double v, int dec) { const double p = std::pow(10.0, dec); return std::round(v * p) / p; }

static int cmd_reduce(const std::string& path, const std::string& strategy_path, const std::string& methods_path,
                      const std::string& properties_path, const std::string& correlate_path, const std::string& correlate_tex_path,
                      const std::string& correlate_counts_path, const std::string& correlate_merged_path) {
  auto rows = csv_read(path);
  if (rows.size() < 2) die("empty taxonomy");
  std::map<std::string, int> col;
  for (size_t i = 0; i < rows[0].size(); ++i) col[rows[0][i]] = (int)i;
  for (const char* need : {"Cases", "Repo"}) if (!col.count(need)) die(std::string("taxonomy lacks column ") + need);
  // repo ids in the column order of tests_strategy_rq32.csv:1 when all nine are present, else first-seen order
  std::vector<std::string> repos = {"autokeras", "auto_sklearn", "tpot", "Ray", "DeepSpeech2", "google_automl", "nni", "Apollo", "Nupic"};
  std::map<std::string, int> rid, cid;
  for (size_t r = 1; r < rows.size(); ++r) if ((int)rows[r].size() > col["Repo"] && !std::count(repos.begin(), repos.end(), rows[r][col["Repo"]])) repos.push_back(rows[r][col["Repo"]]);
  for (size_t i = 0; i < repos.size(); ++i) rid[repos[i]] = (int)i;
  const int nS = sizeof(kStrategy) / sizeof(kStrategy[0]), nM = sizeof(kMethods) / sizeof(kMethods[0]);
  const int nP = sizeof(kProperties) / sizeof(kProperties[0]);
  const int nCR = sizeof(kCorrRows) / sizeof(kCorrRows[0]), nCC = sizeof(kCorrCols) / sizeof(kCorrCols[0]);
  const bool corr = !correlate_path.empty() || !correlate_tex_path.empty() || !correlate_counts_path.empty();
  const bool merged = !correlate_merged_path.empty();
  const int nMR = sizeof(kMergedRows) / sizeof(kMergedRows[0]);
  const int nF = nS + nM + nP + (corr ? nCR * nCC : 0) + (merged ? nMR * nCC : 0);   // the correlate tables are more flag columns of the same reduction
  int merged_row[sizeof(kMergedRows) / sizeof(kMergedRows[0])];
  for (int j = 0; j < nMR; ++j) {
    merged_row[j] = -1;
    for (int k = 0; k < nS; ++k) if (!strcmp(kMergedRows[j].strategy, kStrategy[k].name)) merged_row[j] = k;
    if (merged_row[j] < 0) die(std::string("no strategy row named ") + kMergedRows[j].strategy);
  }
  int corr_prop[sizeof(kCorrCols) / sizeof(kCorrCols[0])];
  for (int q = 0; q < nCC; ++q) {
    corr_prop[q] = -1;
    for (int j = 0; j < nP; ++j) if (!strcmp(kCorrCols[q].prop, kProperties[j].name)) corr_prop[q] = j;
    if (corr_prop[q] < 0) die(std::string("no property named ") + kCorrCols[q].prop);
  }
  std::vector<uint8_t> flags; std::vector<int32_t> repo, cas;
  auto cell = [&](const std::vector<std::string>& r, const char* c) -> std::string {
    auto it = col.find(c); return (it == col.end() || it->second >= (int)r.size()) ? std::string() : trim(r[it->second]); };
  for (size_t r = 1; r < rows.size(); ++r) {
    const auto& R = rows[r];
    if ((int)R.size() <= std::max(col["Cases"], col["Repo"])) continue;
    const std::string cs = R[col["Cases"]];
    if (!cid.count(cs)) { const int k = (int)cid.size(); cid[cs] = k; }
    repo.push_back(rid[R[col["Repo"]]]); cas.push_back(cid[cs]);
    const size_t s0 = flags.size();
    for (int j = 0; j < nS; ++j) {
      bool v = one_of(cell(R, kStrategy[j].col), kStrategy[j].val);
      if (kStrategy[j].col2) v = v || cell(R, kStrategy[j].col2) == "1";
      flags.push_back(v);
    }
    for (int j = 0; j < nM; ++j) { const std::string v = cell(R, kMethods[j].col); flags.push_back(!(v.empty() || v == "0")); }
    const std::string data = cell(R, "Data"), model = cell(R, "Model");
    const size_t p0 = flags.size();
    for (int j = 0; j < nP; ++j)
      flags.push_back((!data.empty() && one_of(data, kProperties[j].labels)) || (!model.empty() && one_of(model, kProperties[j].labels)));
    if (corr)
      for (int j = 0; j < nCR; ++j) {
        const bool s_on = cell(R, kCorrRows[j].col) == kCorrRows[j].val;
        for (int q = 0; q < nCC; ++q) flags.push_back(s_on && flags[p0 + (size_t)corr_prop[q]]);
      }
    if (merged)
      for (int j = 0; j < nMR; ++j) {
        const bool s_on = flags[s0 + (size_t)merged_row[j]] != 0;
        for (int q = 0; q < nCC; ++q) flags.push_back(s_on && flags[p0 + (size_t)corr_prop[q]]);
      }
  }
  const int n_rows = (int)repo.size(), n_repos = (int)repos.size(), n_cases = (int)cid.size();
  tsm_ctx* ctx = nullptr;
  ck(tsm_create(&ctx, 0, 1 << 20, 16, 1, 0), "tsm_create");
  std::vector<int64_t> out((size_t)nF * n_repos), cpr((size_t)n_repos);
  ck(tsm_reduce(ctx, flags.data(), repo.data(), cas.data(), n_rows, nF, n_repos, n_cases, out.data(), cpr.data(), nullptr), "tsm_reduce");
  tsm_destroy(ctx);
  int64_t all_cases = 0; for (int64_t c : cpr) all_cases += c;
  if (!strategy_path.empty()) {                             // layout of RQs/RQ3/tests_strategy_rq32.csv
    std::ofstream os(strategy_path, std::ios::binary);
    std::vector<std::string> h = {"Tests"};
    for (auto& r : repos) h.push_back(r);
    h.push_back("");
    for (auto& r : repos) h.push_back(r);
    csv_row(os, h);
    std::vector<std::vector<double>> v(nS, std::vector<double>(n_repos));
    std::vector<double> colsum(n_repos, 0.0);
    for (int j = 0; j < nS; ++j) for (int r = 0; r < n_repos; ++r) {
      // rounded twice, like the shipped cells (docs/SPEC.md section 9): 26/142 -> 18.3099 -> /1.1 -> 16.6454
      v[j][r] = cpr[r] ? round_to(round_to(100.0 * out[(size_t)j * n_repos + r] / cpr[r], 4) / 1.1, 4) : 0.0;
      colsum[r] += v[j][r];
    }
    for (int j = 0; j < nS; ++j) {
      std::vector<std::string> row = {kStrategy[j].name};
      for (int r = 0; r < n_repos; ++r) row.push_back(fmt_num(v[j][r], 4));
      row.push_back("");
      for (int r = 0; r < n_repos; ++r) row.push_back(fmt_num(colsum[r] > 0 ? round_to(v[j][r] / colsum[r] * 100.0, 2) : 0.0, 2));
      csv_row(os, row);
    }
    std::vector<std::string> last = {""};
    for (int r = 0; r < n_repos; ++r) last.push_back(fmt_num(colsum[r], 4));
    last.push_back("");
    for (int r = 0; r < n_repos; ++r) last.push_back("100");
    csv_row(os, last);
  }
  if (!methods_path.empty()) {                              // first three columns of RQs/RQ4/tests_methods_v2.csv
    std::ofstream os(methods_path, std::ios::binary);
    csv_row(os, {"Test_methods", "total_cases", "percentage"});
    for (int j = 0; j < nM; ++j) {
      int64_t t = 0;
      for (int r = 0; r < n_repos; ++r) t += out[(size_t)(nS + j) * n_repos + r];
      csv_row(os, {kMethods[j].name, std::to_string(t), fmt_num(all_cases ? round_to(100.0 * t / all_cases, 4) : 0.0, 4)});
    }
  }
  if (!properties_path.empty()) {                           // layout of RQs/RQ3/tests_prop_rq3.csv:1-10
    std::ofstream os(properties_path, std::ios::binary);
    std::vector<std::string> h = {"Repos"};
    for (int j = 0; j < nP; ++j) h.push_back(kProperties[j].name);
    csv_row(os, h);
    int64_t denom = 0;                                      // one denominator for every row: Apollo's case count (the 216 of
    if (rid.count("Apollo")) denom = cpr[(size_t)rid["Apollo"]];   // the shipped table), else the largest repository
    if (denom == 0) for (int64_t c : cpr) denom = std::max(denom, c);
    // shipped row order when all nine repositories are present, else the order of `repos`
    std::vector<std::string> order = {"auto_sklearn", "google_automl", "tpot", "autokeras", "Nupic", "Apollo", "nni", "Ray", "DeepSpeech2"};
    for (auto& r : repos) if (!std::count(order.begin(), order.end(), r)) order.push_back(r);
    for (auto& name : order) {
      if (!rid.count(name) || cpr[(size_t)rid[name]] == 0) continue;
      const int r = rid[name];
      std::vector<std::string> row = {name};
      for (int j = 0; j < nP; ++j)
        row.push_back(fmt_num(denom ? round_to(100.0 * out[(size_t)(nS + nM + j) * n_repos + r] / denom, 4) : 0.0, 4));
      csv_row(os, row);
    }
  }
  if (corr) {
    // three shipped layouts of the same 20 x 21 counts: RQs/RQ3/tests_correlate_rq3.csv ("repo:(p%), " for every repository),
    // tests_correlate_rq4.csv (LaTeX cells "$repo:p\%$, " of the non-zero repositories) and
    // tests_combined_correlate_rq3.csv (the distinct cases of all repositories together)
    for (const CorrRow& cr : kCorrRows) if (!col.count(cr.col)) die(std::string("taxonomy lacks column ") + cr.col);
    std::vector<std::string> order = {"auto_sklearn", "google_automl", "tpot", "autokeras", "Nupic", "Apollo", "nni", "Ray", "DeepSpeech2"};
    for (auto& r : repos) if (!std::count(order.begin(), order.end(), r)) order.push_back(r);
    const size_t c0 = (size_t)(nS + nM + nP);
    for (int layout = 0; layout < 3; ++layout) {
      const std::string& outp = layout == 0 ? correlate_path : layout == 1 ? correlate_tex_path : correlate_counts_path;
      if (outp.empty()) continue;
      std::ofstream os(outp, std::ios::binary);
      std::vector<std::string> h = {"Tests"};
      for (int q = 0; q < nCC; ++q) h.push_back(kCorrCols[q].name);
      csv_row(os, h);
      for (int j = 0; j < nCR; ++j) {
        std::vector<std::string> row = {kCorrRows[j].name};
        for (int q = 0; q < nCC; ++q) {
          const int64_t* d = &out[(c0 + (size_t)j * nCC + q) * n_repos];
          int64_t all = 0;
          for (int r = 0; r < n_repos; ++r) all += d[r];
          std::string cellv = "0";                          // a pairing no case has is the bare string "0" in every layout
          if (layout == 2) cellv = std::to_string(all);
          else if (all) {
            cellv.clear();
            for (auto& name : order) {
              if (!rid.count(name) || cpr[(size_t)rid[name]] == 0) continue;
              const int r = rid[name];
              const std::string pct = fmt_pyfloat2(100.0 * (double)d[r] / (double)cpr[r]);
              if (layout == 0) cellv += name + ":(" + pct + "%), ";
              else if (d[r]) cellv += "$" + name + ":" + pct + "\\%$, ";
            }
          }
          row.push_back(cellv);
        }
        csv_row(os, row);
      }
    }
  }
  if (merged) {                                             // the four one-row tables, as four rows of one file
    std::ofstream os(correlate_merged_path, std::ios::binary);
    std::vector<std::string> h = {"Tests"};
    for (int q = 0; q < nCC; ++q) h.push_back(kCorrCols[q].name);
    csv_row(os, h);
    std::vector<std::string> order = {"auto_sklearn", "google_automl", "tpot", "autokeras", "Nupic", "Apollo", "nni", "Ray", "DeepSpeech2"};
    for (auto& r : repos) if (!std::count(order.begin(), order.end(), r)) order.push_back(r);
    const size_t c0 = (size_t)(nS + nM + nP + (corr ? nCR * nCC : 0));
    for (int j = 0; j < nMR; ++j) {
      std::vector<std::string> row = {kMergedRows[j].name};
      for (int q = 0; q < nCC; ++q) {
        const int64_t* d = &out[(c0 + (size_t)j * nCC + q) * n_repos];
        int64_t all = 0;
        for (int r = 0; r < n_repos; ++r) all += d[r];
        std::string cellv = "0";
        if (all) {
          cellv.clear();
          for (auto& name : order) {
            if (!rid.count(name) || cpr[(size_t)rid[name]] == 0) continue;
            const int r = rid[name];
            cellv += name + ":(" + fmt_pyfloat2(100.0 * (double)d[r] / (double)cpr[r]) + "%), ";
          }
        }
        row.push_back(cellv);
      }
      csv_row(os, row);
    }
  }
  fprintf(stderr, "tosem-scan: reduce %d rows, %d cases, %d repos\n", n_rows, n_cases, n_repos);
  return 0;
}

// ---------------------------------------------------------------------------------- body statements (SPEC section 10)
// Case name of a header line: PY - identifier after `def`; C family - 2nd macro argument of TEST / TEST_F /
// TEST_P, `TEST_CASE(X)` for BOOST_AUTO_TEST_CASE(X); otherwise the SPEC section 5 method string.
static std::string case_name(int ext, const uint8_t* line, uint32_t len) {
  uint32_t b = 0, e = len;
  while (b < e && is_w(line[b])) ++b;
  while (e > b && is_w(line[e - 1])) --e;
  const std::string s((const char*)line + b, e - b);
  if (ext == TSM_EXT_PY) {
    const size_t d = s.find("def");
    if (d != std::string::npos) {
      size_t i = d + 3;
      while (i < s.size() && is_w((unsigned char)s[i])) ++i;
      size_t j = i;
      while (j < s.size() && (isalnum((unsigned char)s[j]) || s[j] == '_')) ++j;
      if (j > i) return s.substr(i, j - i);
    }
  } else {
    auto trim = [](std::string t) { size_t a = 0, z = t.size(); while (a < z && is_w((unsigned char)t[a])) ++a; while (z > a && is_w((unsigned char)t[z - 1])) --z; return t.substr(a, z - a); };
    if (s.rfind("TEST(", 0) == 0 || s.rfind("TEST_F(", 0) == 0 || s.rfind("TEST_P(", 0) == 0) {
      const size_t c = s.find(','), r = s.find(')');
      if (c != std::string::npos && (r == std::string::npos || c < r)) return trim(s.substr(c + 1, (r == std::string::npos ? s.size() : r) - c - 1));
    }
    if (s.rfind("BOOST_AUTO_TEST_CASE(", 0) == 0) {
      const size_t r = s.find(')');
      return "TEST_CASE(" + trim(s.substr(21, (r == std::string::npos ? s.size() : r) - 21)) + ")";
    }
  }
  return method_string(ext, line, len);
}

static int cmd_body(const std::vector<std::string>& roots, const std::string& out_path) {
  std::vector<FileEntry> files;
  for (size_t g = 0; g < roots.size(); ++g) walk(roots[g], (int)g, false, files);
  fprintf(stderr, "tosem-scan: %zu files selected under %zu root(s)\n", files.size(), roots.size());
  std::ofstream os;
  if (!out_path.empty()) { os.open(out_path, std::ios::binary); csv_row(os, {"Index", "text", "Category", "cases", "File_ID", "Component"}); }
  int64_t index = 0, cases = 0, file_id = 0, n_stmt = 0;
  size_t first = 0;
  while (first < files.size()) {
    Batch B;
    int64_t cur = 0;
    while (first < files.size() && (B.count() == 0 || (cur + files[first].size < (1ll << 29) && B.count() < (1u << 19)))) {
      cur += (files[first].size + 127) / 128 * 128; B.idx.push_back((uint32_t)first); ++first;
    }
    load_batch(files, B);
    tsm_ctx* ctx = nullptr;
    ck(tsm_create(&ctx, 0, B.bytes + 4096, (int32_t)B.count(), (int32_t)std::max<size_t>(roots.size(), 1), 0), "tsm_create");
    tsm_corpus c{B.arena, B.off.data(), B.len.data(), B.ext.data(), B.grp.data(), (int32_t)B.count(), (int32_t)std::max<size_t>(roots.size(), 1)};
    const int64_t cap = std::max<int64_t>(B.bytes / 8 + 1024, 1024);
    std::vector<tsm_header_event> hev((size_t)cap);
    tsm_result r{};
    r.hev = hev.data(); r.hev_cap = cap;
    ck(tsm_scan(ctx, &c, &r, TSM_SCAN_HEADER_EVENTS, nullptr), "tsm_scan");
    hev.resize((size_t)r.n_hev);
    std::vector<int64_t> base(B.count() + 1);
    int64_t nl = 0;
    int rc = tsm_statements(ctx, &c, base.data(), nullptr, nullptr, 0, &nl, nullptr);
    if (rc != TSM_OK && rc != TSM_E_CAPACITY) ck(rc, "tsm_statements");
    std::vector<uint32_t> lend((size_t)std::max<int64_t>(nl, 1));
    std::vector<uint8_t> kind((size_t)std::max<int64_t>(nl, 1));
    ck(tsm_statements(ctx, &c, base.data(), lend.data(), kind.data(), nl, &nl, nullptr), "tsm_statements");
    tsm_destroy(ctx);
    size_t hi = 0;
    for (size_t i = 0; i < B.count(); ++i) {
      const FileEntry& f = files[B.idx[i]];
      const uint8_t* p = B.arena + B.off[i];
      ++file_id;
      bool in_case = false;
      std::string cur_stmt; bool have = false, listed = false;
      auto flush = [&]() {
        if (have && listed && cur_stmt.find_first_not_of("{}(); \t\r\x0b\x0c") != std::string::npos) {
          ++n_stmt;
          if (os.is_open()) csv_row(os, {std::to_string(++index), cur_stmt, "", std::to_string(cases), std::to_string(file_id), ""});
        }
        have = false; cur_stmt.clear();
      };
      uint32_t pos = 0;
      for (int64_t l = base[i]; l < base[i + 1]; ++l) {
        const uint32_t e = lend[(size_t)l];
        const bool is_hdr = hi < hev.size() && hev[hi].file == i && hev[hi].line_off == pos;
        if (kind[(size_t)l] == 1) { flush(); listed = in_case && !is_hdr; have = true; }
        if (is_hdr) {                                       // a new test case starts here
          ++hi; in_case = true; ++cases;
          if (os.is_open()) csv_row(os, {std::to_string(++index), case_name(f.ext, p + pos, e - pos), "", std::to_string(cases), std::to_string(file_id), ""});
        }
        if (kind[(size_t)l] != 0 && have) {
          uint32_t b = pos, z = e;
          while (b < z && is_w(p[b])) ++b;
          while (z > b && is_w(p[z - 1])) --z;
          if (!cur_stmt.empty()) cur_stmt += ' ';
          cur_stmt.append((const char*)p + b, z - b);
        }
        pos = e + 1;
      }
      flush();
      while (hi < hev.size() && hev[hi].file == i) ++hi;
    }
    tsm_host_free(B.arena);
  }
  printf("files,cases,statements\r\n%lld,%lld,%lld\r\n", (long long)file_id, (long long)cases, (long long)n_stmt);
  return 0;
}

// ---------------------------------------------------------------------------------- release presence matrix (S7, SPEC section 11)
struct SnapFile { std::string rel; uint64_t digest; int64_t size; uint32_t n_assert; std::string hist; };

// Scan one snapshot: per selected test file its digest, assertion total and "n:category, ..." histogram.
static std::vector<SnapFile> scan_snapshot(const std::vector<FileEntry>& files) {
  std::vector<SnapFile> out;
  size_t first = 0;
  while (first < files.size()) {
    Batch B;
    int64_t cur = 0;
    while (first < files.size() && (B.count() == 0 || (cur + files[first].size < (1ll << 29) && B.count() < (1u << 19)))) {
      cur += (files[first].size + 127) / 128 * 128; B.idx.push_back((uint32_t)first); ++first;
    }
    load_batch(files, B);
    tsm_ctx* ctx = nullptr;
    ck(tsm_create(&ctx, 0, B.bytes + 4096, (int32_t)B.count(), 1, 0), "tsm_create");
    tsm_corpus c{B.arena, B.off.data(), B.len.data(), B.ext.data(), B.grp.data(), (int32_t)B.count(), 1};
    const int64_t cap = std::max<int64_t>(B.bytes / 8 + 1024, 1024);
    std::vector<tsm_file_stat> stats(B.count());
    std::vector<tsm_assert_event> aev((size_t)cap);
    tsm_result r{};
    r.stats = stats.data(); r.aev = aev.data(); r.aev_cap = cap;
    ck(tsm_scan(ctx, &c, &r, TSM_SCAN_ASSERT_EVENTS, nullptr), "tsm_scan");
    tsm_destroy(ctx);
    size_t ai = 0;
    for (size_t i = 0; i < B.count(); ++i) {
      const uint8_t* base = B.arena + B.off[i];
      std::map<std::string, int64_t> hist; std::vector<std::string> order;
      for (; ai < (size_t)r.n_aev && aev[ai].file == i; ++ai) {
        const tsm_assert_event& ev = aev[ai];
        const std::string cat = ev.cat == 127 ? std::string((const char*)base + ev.ident_off, ev.ident_len) : std::string(tsm_category_name(ev.cat));
        if (!hist.count(cat)) order.push_back(cat);
        hist[cat]++;
      }
      std::stable_sort(order.begin(), order.end(), [&](const std::string& x, const std::string& y) { return hist[x] > hist[y]; });
      std::string a;
      for (const std::string& k : order) { if (!a.empty()) a += ", "; a += std::to_string(hist[k]) + ":" + k; }
      out.push_back({files[B.idx[i]].rel, stats[i].digest, files[B.idx[i]].size, stats[i].n_assert, a});
    }
    tsm_host_free(B.arena);
  }
  return out;
}

// The selected test files (S0 + S1) of one tree of a git repository, in the order `walk` gives for a checkout of it
// (names sorted at every level, depth first); their bytes are inflated here and scanned like files read from disk.
static void walk_git(gitstore::Store& gs, const gitstore::Oid& tree, const std::string& prefix, bool all_files, std::vector<FileEntry>& out) {
  std::vector<gitstore::TreeEntry> es;
  if (!gs.tree(tree, es)) die("unreadable tree " + tree.hex());
  std::sort(es.begin(), es.end(), [](const gitstore::TreeEntry& a, const gitstore::TreeEntry& b) { return a.name < b.name; });
  for (const gitstore::TreeEntry& e : es) {
    const std::string rel = prefix + e.name;
    if (e.is_tree()) { walk_git(gs, e.oid, rel + "/", all_files, out); continue; }
    if (!e.is_blob()) continue;
    const int ext = ext_tag(rel);
    if (!all_files && (lower(rel).find("test") == std::string::npos || ext == TSM_EXT_OTHER)) continue;
    gitstore::Object o;
    if (!gs.read(e.oid, o) || o.type != gitstore::OBJ_BLOB) die("unreadable blob " + e.oid.hex());
    if (o.data.size() > 0x7fff0000u) die("blob too large: " + rel);
    FileEntry f{rel, "", ext, 0, (int64_t)o.data.size(), nullptr};
    f.blob = std::make_shared<const std::vector<uint8_t>>(std::move(o.data));
    out.push_back(std::move(f));
  }
}

// specs: <snapshot-root>=<tag>...; or, with --git <repository>, revisions (tags, branches, object names) in release
// order - none = every tag of the repository, oldest commit first.
static int cmd_releases(const std::vector<std::string>& specs_in, const std::string& out_path, const std::string& git_repo) {
  std::vector<std::string> specs = specs_in;
  gitstore::Store gs;
  if (!git_repo.empty()) {
    std::string err;
    if (!gs.open(git_repo, err)) die(err);
    if (specs.empty()) {
      std::vector<std::pair<long long, std::string>> byt;
      for (const std::string& t : gs.tag_names()) {
        gitstore::Oid id; gitstore::Commit c;
        if (gs.resolve("refs/tags/" + t, id) && gs.commit(id, c)) byt.push_back({c.time, t});
      }
      std::sort(byt.begin(), byt.end());
      for (auto& kv : byt) specs.push_back(kv.second);
      if (specs.empty()) die("the repository has no tags; name the revisions");
    }
  }
  struct Identity { std::string name; std::vector<std::string> path; uint64_t digest; int64_t size; uint32_t n_assert; std::string hist; std::string cur; };
  std::vector<std::string> tags;
  std::vector<Identity> ids;
  for (size_t t = 0; t < specs.size(); ++t) {
    std::vector<FileEntry> files;
    if (!git_repo.empty()) {
      gitstore::Oid id; gitstore::Commit c;
      if (!gs.resolve(specs[t], id) || !gs.commit(id, c)) die("cannot resolve revision " + specs[t]);
      tags.push_back(specs[t]);
      walk_git(gs, c.tree, "", false, files);
    } else {
      const size_t eq = specs[t].rfind('=');
      if (eq == std::string::npos) die("releases arguments are <snapshot-root>=<tag>");
      tags.push_back(specs[t].substr(eq + 1));
      walk(specs[t].substr(0, eq), 0, false, files);
    }
    const std::vector<SnapFile> snap = scan_snapshot(files);
    std::vector<char> id_taken(ids.size(), 0), f_done(snap.size(), 0);
    auto base_of = [](const std::string& p) { const size_t s = p.rfind('/'); return s == std::string::npos ? p : p.substr(s + 1); };
    auto bind = [&](size_t fi, size_t id) {
      const SnapFile& f = snap[fi];
      ids[id].path[t] = f.rel; ids[id].cur = f.rel; ids[id].digest = f.digest; ids[id].size = f.size;
      ids[id].n_assert = f.n_assert; ids[id].hist = f.hist; id_taken[id] = 1; f_done[fi] = 1;
    };
    for (Identity& I : ids) I.path.resize(t + 1);
    for (size_t fi = 0; fi < snap.size(); ++fi)                       // (1) same relative path
      for (size_t id = 0; id < id_taken.size(); ++id)
        if (!id_taken[id] && ids[id].cur == snap[fi].rel) { bind(fi, id); break; }
    for (size_t fi = 0; fi < snap.size(); ++fi)                       // (2) same content: a pure move
      if (!f_done[fi])
        for (size_t id = 0; id < id_taken.size(); ++id)
          if (!id_taken[id] && ids[id].size == snap[fi].size && ids[id].digest == snap[fi].digest) { bind(fi, id); break; }
    for (size_t fi = 0; fi < snap.size(); ++fi) {                     // (3) same base name, unambiguous
      if (f_done[fi]) continue;
      int hit = -1, n = 0;
      for (size_t id = 0; id < id_taken.size(); ++id)
        if (!id_taken[id] && base_of(ids[id].cur) == base_of(snap[fi].rel)) { hit = (int)id; ++n; }
      if (n == 1) bind(fi, (size_t)hit);
    }
    for (size_t fi = 0; fi < snap.size(); ++fi)                       // new identities, in walk order
      if (!f_done[fi]) {
        Identity I; I.name = snap[fi].rel; I.path.assign(t + 1, "");
        ids.push_back(I); id_taken.push_back(0);
        bind(fi, ids.size() - 1);
      }
  }
  std::ofstream os;
  if (!out_path.empty()) {
    os.open(out_path, std::ios::binary);
    std::vector<std::string> h = {"Id", "FileName"};
    for (const std::string& g : tags) h.push_back(g);
    h.push_back("total assert"); h.push_back("assertion");
    csv_row(os, h);
    for (size_t i = 0; i < ids.size(); ++i) {
      std::vector<std::string> row = {std::to_string(i + 1), ids[i].name};
      for (size_t t = 0; t < tags.size(); ++t) row.push_back(t < ids[i].path.size() ? ids[i].path[t] : "");
      row.push_back(std::to_string(ids[i].n_assert)); row.push_back(ids[i].hist);
      csv_row(os, row);
    }
  }
  printf("identities,snapshots\r\n%zu,%zu\r\n", ids.size(), tags.size());
  return 0;
}

// ---------------------------------------------------------------------------------- diff (S8)
static int cmd_diff(const std::string& old_root, const std::string& new_root, const std::string& out_path) {
  std::vector<FileEntry> a, b;
  walk(old_root, 0, true, a);
  walk(new_root, 0, true, b);
  std::map<std::string, const FileEntry*> bm;
  for (const FileEntry& f : b) bm[f.rel] = &f;
  // pairs by relative path; a file present on one side only is paired with the empty file
  struct Pair { std::string rel; const FileEntry* o; const FileEntry* n; };
  std::vector<Pair> pairs;
  std::map<std::string, bool> seen;
  for (const FileEntry& f : a) { auto it = bm.find(f.rel); pairs.push_back({f.rel, &f, it == bm.end() ? nullptr : it->second}); seen[f.rel] = true; }
  for (const FileEntry& f : b) if (!seen.count(f.rel)) pairs.push_back({f.rel, nullptr, &f});
  // git's numstat reports no line counts for binary files: a pair is skipped when either side has a NUL byte in its first 8000
  {
    auto binary = [](const FileEntry* f) {
      if (!f || f->abs.empty() || f->size == 0) return false;
      char buf[8000];
      const int fd = open(f->abs.c_str(), O_RDONLY);
      if (fd < 0) return false;
      const ssize_t r = read(fd, buf, sizeof buf);
      close(fd);
      return r > 0 && memchr(buf, 0, (size_t)r) != nullptr;
    };
    std::vector<Pair> text;
    size_t skipped = 0;
    for (const Pair& p : pairs) { if (binary(p.o) || binary(p.n)) ++skipped; else text.push_back(p); }
    if (skipped) fprintf(stderr, "tosem-scan: %zu binary file(s) skipped\n", skipped);
    pairs.swap(text);
  }
  auto pack = [&](bool old_side, Batch& B, std::vector<FileEntry>& tmp) {
    for (const Pair& p : pairs) { const FileEntry* f = old_side ? p.o : p.n; tmp.push_back(f ? *f : FileEntry{p.rel, "", 0, 0, 0, nullptr}); }
    const size_t n = tmp.size();
    B.idx.resize(n);
    for (size_t i = 0; i < n; ++i) B.idx[i] = (uint32_t)i;
    B.len.resize(n); B.off.resize(n + 1); B.ext.assign(n, 0); B.grp.assign(n, 0);
    for (size_t i = 0; i < n; ++i) B.len[i] = (int32_t)tmp[i].size;
    B.bytes = tsm_layout(B.len.data(), (int32_t)n, B.off.data());
    if (B.bytes < 0) die("tree does not fit one int32-indexed arena; diff it per sub-directory");
    B.arena = (uint8_t*)tsm_host_alloc(std::max<int64_t>(B.bytes, 128));
    if (!B.arena) die("pinned arena allocation failed");
    memset(B.arena, 0, (size_t)std::max<int64_t>(B.bytes, 128));
    for (size_t i = 0; i < n; ++i) {
      if (tmp[i].abs.empty() || B.len[i] == 0) continue;
      const int fd = open(tmp[i].abs.c_str(), O_RDONLY);
      int64_t got = 0;
      while (fd >= 0 && got < B.len[i]) {
        const ssize_t r = read(fd, B.arena + B.off[i] + got, (size_t)(B.len[i] - got));
        if (r <= 0) break;
        got += r;
      }
      if (fd >= 0) close(fd);
      if (got != B.len[i]) die("short read: " + tmp[i].abs);
    }
  };
  Batch A, N; std::vector<FileEntry> ta, tn;
  pack(true, A, ta); pack(false, N, tn);
  tsm_ctx* ctx = nullptr;
  ck(tsm_create(&ctx, 0, 1 << 20, 16, 1, 0), "tsm_create");
  tsm_corpus ca{A.arena, A.off.data(), A.len.data(), A.ext.data(), A.grp.data(), (int32_t)A.count(), 1};
  tsm_corpus cn{N.arena, N.off.data(), N.len.data(), N.ext.data(), N.grp.data(), (int32_t)N.count(), 1};
  // ext tags of both sides feed the assertion-line classification of the changed lines
  for (size_t i = 0; i < pairs.size(); ++i) { A.ext[i] = (uint8_t)ext_tag(pairs[i].rel); N.ext[i] = A.ext[i]; }
  std::vector<int64_t> added(pairs.size()), removed(pairs.size());
  std::vector<tsm_diff_detail> det(pairs.size());
  ck(tsm_diff_pairs_detail(ctx, &ca, &cn, added.data(), removed.data(), det.data(), nullptr), "tsm_diff_pairs_detail");
  tsm_destroy(ctx);
  std::ofstream os;
  if (!out_path.empty()) {
    os.open(out_path, std::ios::binary);
    csv_row(os, {"fileName", "cloc", "added", "removed", "hunks_add", "hunks_del", "hunks_mod", "added_assert", "removed_assert"});
  }
  int64_t ta_ = 0, tr_ = 0;
  for (size_t i = 0; i < pairs.size(); ++i) {
    ta_ += added[i]; tr_ += removed[i];
    if (os.is_open() && (added[i] || removed[i]))
      csv_row(os, {pairs[i].rel, std::to_string(added[i] + removed[i]), std::to_string(added[i]), std::to_string(removed[i]),
                   std::to_string(det[i].hunks_add), std::to_string(det[i].hunks_del), std::to_string(det[i].hunks_mod),
                   std::to_string(det[i].added_assert), std::to_string(det[i].removed_assert)});
  }
  printf("cloc,added,removed\r\n%lld,%lld,%lld\r\n", (long long)(ta_ + tr_), (long long)ta_, (long long)tr_);
  tsm_host_free(A.arena); tsm_host_free(N.arena);
  return 0;
}

// ---------------------------------------------------------------------------------- S8 on a real repository
// `tosem-scan history <repo>`: the churn of the test files along the first-parent history of a revision, read straight
// from the git object store (host/git_store.hpp: loose objects, packfiles, refs - no `git` process, no checkout).
// Host: commit chain, tree diff by object name (only entries whose blob changed are opened), blob inflation into the
// two pinned arenas.  GPU: line records of both sides + per-pair Myers + hunks + changed assertion lines
// (tsm_diff_pairs_detail), one call per batch of at most ~512 MiB per side.  Rows: one per (commit, changed file).
struct BlobChange { std::string path; gitstore::Oid o, n; bool has_o, has_n; };

static void tree_diff(gitstore::Store& gs, const gitstore::Oid* a, const gitstore::Oid* b, const std::string& prefix,
                      bool all_files, std::vector<BlobChange>& out) {
  std::vector<gitstore::TreeEntry> ea, eb;
  if (a && !gs.tree(*a, ea)) die("unreadable tree " + a->hex());
  if (b && !gs.tree(*b, eb)) die("unreadable tree " + b->hex());
  std::map<std::string, const gitstore::TreeEntry*> ma, mb;
  for (auto& e : ea) ma[e.name] = &e;
  for (auto& e : eb) mb[e.name] = &e;
  std::vector<std::string> names;
  for (auto& kv : ma) names.push_back(kv.first);
  for (auto& kv : mb) if (!ma.count(kv.first)) names.push_back(kv.first);
  std::sort(names.begin(), names.end());
  for (const std::string& nm : names) {
    const gitstore::TreeEntry* x = ma.count(nm) ? ma[nm] : nullptr;
    const gitstore::TreeEntry* y = mb.count(nm) ? mb[nm] : nullptr;
    if (x && y && x->oid == y->oid && x->is_tree() == y->is_tree()) continue;     // same object: nothing below it changed
    const std::string path = prefix + nm;
    const bool xt = x && x->is_tree(), yt = y && y->is_tree();
    if (xt || yt) tree_diff(gs, xt ? &x->oid : nullptr, yt ? &y->oid : nullptr, path + "/", all_files, out);
    const bool xb = x && x->is_blob(), yb = y && y->is_blob();                     // (symlinks and submodules are not files of the study)
    if (!xb && !yb) continue;
    if (!all_files && (lower(path).find("test") == std::string::npos || ext_tag(path) == TSM_EXT_OTHER)) continue;   // S0, S1
    BlobChange c{path, {}, {}, xb, yb};
    if (xb) c.o = x->oid;
    if (yb) c.n = y->oid;
    out.push_back(c);
  }
}

// --dry-run: no GPU - the rows carry the object names, sizes and an FNV-1a checksum of both blobs instead of the counts
// (what the CPU tests compare with `git diff-tree` / `git cat-file`).
static int cmd_history(const std::string& repo, const std::string& rev, int64_t max_commits, bool all_files, const std::string& out_path,
                       bool dry_run) {
  gitstore::Store gs;
  std::string err;
  if (!gs.open(repo, err)) die(err);
  gitstore::Oid head;
  if (!gs.resolve(rev, head)) die("cannot resolve revision " + rev);
  struct Step { gitstore::Oid id; gitstore::Commit c; };
  std::vector<Step> chain;
  for (gitstore::Oid id = head; max_commits <= 0 || (int64_t)chain.size() < max_commits;) {
    Step st{id, {}};
    if (!gs.commit(id, st.c)) die("unreadable commit " + id.hex());
    chain.push_back(st);
    if (st.c.parents.empty()) break;
    id = st.c.parents[0];
  }
  std::reverse(chain.begin(), chain.end());                 // oldest first
  struct Row { size_t step; BlobChange ch; };
  std::vector<Row> rows;
  for (size_t i = 0; i < chain.size(); ++i) {
    gitstore::Commit parent;
    const bool has_parent = !chain[i].c.parents.empty();
    if (has_parent && !gs.commit(chain[i].c.parents[0], parent)) die("unreadable commit " + chain[i].c.parents[0].hex());
    if (has_parent && parent.tree == chain[i].c.tree) continue;
    std::vector<BlobChange> ch;
    tree_diff(gs, has_parent ? &parent.tree : nullptr, &chain[i].c.tree, "", all_files, ch);
    for (auto& c : ch) rows.push_back({i, c});
  }
  std::ofstream os;
  if (!out_path.empty()) {
    os.open(out_path, std::ios::binary);
    if (dry_run) csv_row(os, {"commit", "parent", "time", "fileName", "old_blob", "new_blob", "old_size", "new_size", "old_fnv", "new_fnv"});
    else csv_row(os, {"commit", "parent", "time", "fileName", "cloc", "added", "removed", "hunks_add", "hunks_del", "hunks_mod", "added_assert", "removed_assert"});
  }
  tsm_ctx* ctx = nullptr;
  if (!dry_run) ck(tsm_create(&ctx, 0, 1 << 20, 16, 1, 0), "tsm_create");
  std::vector<int64_t> per_add(chain.size(), 0), per_rem(chain.size(), 0), per_files(chain.size(), 0);
  int64_t binaries = 0, pairs_done = 0;
  const int64_t kBatch = 512ll << 20;
  size_t r0 = 0;
  while (r0 < rows.size()) {
    // inflate blobs until a side of the batch is full
    std::vector<std::vector<uint8_t>> bo, bn;
    std::vector<size_t> idx;
    int64_t so = 0, sn = 0;
    size_t r1 = r0;
    for (; r1 < rows.size() && so < kBatch && sn < kBatch; ++r1) {
      const BlobChange& c = rows[r1].ch;
      gitstore::Object x, y;
      if (c.has_o && (!gs.read(c.o, x) || x.type != gitstore::OBJ_BLOB)) die("unreadable blob " + c.o.hex());
      if (c.has_n && (!gs.read(c.n, y) || y.type != gitstore::OBJ_BLOB)) die("unreadable blob " + c.n.hex());
      auto binary = [](const std::vector<uint8_t>& v) { return !v.empty() && memchr(v.data(), 0, std::min<size_t>(v.size(), 8000)) != nullptr; };
      if (dry_run) {
        auto fnv = [](const std::vector<uint8_t>& v) { uint64_t h = 0xcbf29ce484222325ull; for (uint8_t b : v) h = (h ^ b) * 0x100000001b3ull; char buf[24]; snprintf(buf, sizeof buf, "%016llx", (unsigned long long)h); return std::string(buf); };
        const Step& st = chain[rows[r1].step];
        if (os.is_open())
          csv_row(os, {st.id.hex(), st.c.parents.empty() ? "" : st.c.parents[0].hex(), std::to_string(st.c.time), c.path, c.has_o ? c.o.hex() : "", c.has_n ? c.n.hex() : "",
                       std::to_string(x.data.size()), std::to_string(y.data.size()), fnv(x.data), fnv(y.data)});
        per_files[rows[r1].step]++;
        continue;
      }
      if (binary(x.data) || binary(y.data)) { ++binaries; continue; }              // like git's numstat: no line counts for binary files
      if (x.data.size() > 0x7fff0000u || y.data.size() > 0x7fff0000u) die("blob too large: " + c.path);
      so += (int64_t)x.data.size() + 256; sn += (int64_t)y.data.size() + 256;
      bo.push_back(std::move(x.data)); bn.push_back(std::move(y.data)); idx.push_back(r1);
    }
    const size_t n = idx.size();
    if (n) {
      Batch A, N;
      auto pack = [&](Batch& B, const std::vector<std::vector<uint8_t>>& blobs) {
        B.len.resize(n); B.off.resize(n + 1); B.ext.assign(n, 0); B.grp.assign(n, 0);
        for (size_t i = 0; i < n; ++i) B.len[i] = (int32_t)blobs[i].size();
        B.bytes = tsm_layout(B.len.data(), (int32_t)n, B.off.data());
        if (B.bytes < 0) die("batch does not fit one int32-indexed arena");
        B.arena = (uint8_t*)tsm_host_alloc(std::max<int64_t>(B.bytes, 128));
        if (!B.arena) die("pinned arena allocation failed");
        memset(B.arena, 0, (size_t)std::max<int64_t>(B.bytes, 128));
        for (size_t i = 0; i < n; ++i) if (B.len[i]) memcpy(B.arena + B.off[i], blobs[i].data(), blobs[i].size());
      };
      pack(A, bo); pack(N, bn);
      for (size_t i = 0; i < n; ++i) { A.ext[i] = (uint8_t)ext_tag(rows[idx[i]].ch.path); N.ext[i] = A.ext[i]; }
      tsm_corpus ca{A.arena, A.off.data(), A.len.data(), A.ext.data(), A.grp.data(), (int32_t)n, 1};
      tsm_corpus cn{N.arena, N.off.data(), N.len.data(), N.ext.data(), N.grp.data(), (int32_t)n, 1};
      std::vector<int64_t> added(n), removed(n);
      std::vector<tsm_diff_detail> det(n);
      ck(tsm_diff_pairs_detail(ctx, &ca, &cn, added.data(), removed.data(), det.data(), nullptr), "tsm_diff_pairs_detail");
      for (size_t i = 0; i < n; ++i) {
        const Row& r = rows[idx[i]];
        per_add[r.step] += added[i]; per_rem[r.step] += removed[i]; per_files[r.step]++;
        if (os.is_open()) {
          const Step& st = chain[r.step];
          csv_row(os, {st.id.hex(), st.c.parents.empty() ? "" : st.c.parents[0].hex(), std::to_string(st.c.time), r.ch.path,
                       std::to_string(added[i] + removed[i]), std::to_string(added[i]), std::to_string(removed[i]),
                       std::to_string(det[i].hunks_add), std::to_string(det[i].hunks_del), std::to_string(det[i].hunks_mod),
                       std::to_string(det[i].added_assert), std::to_string(det[i].removed_assert)});
        }
      }
      pairs_done += (int64_t)n;
      tsm_host_free(A.arena); tsm_host_free(N.arena);
    }
    r0 = r1;
  }
  if (ctx) tsm_destroy(ctx);
  printf("commit,files,cloc,added,removed\r\n");
  int64_t ta = 0, tr = 0;
  for (size_t i = 0; i < chain.size(); ++i) {
    ta += per_add[i]; tr += per_rem[i];
    printf("%s,%lld,%lld,%lld,%lld\r\n", chain[i].id.hex().c_str(), (long long)per_files[i], (long long)(per_add[i] + per_rem[i]),
           (long long)per_add[i], (long long)per_rem[i]);
  }
  fprintf(stderr, "tosem-scan: history of %s: %zu commits, %lld changed files diffed on the GPU, %lld binary skipped, cloc %lld (+%lld -%lld)\n",
          rev.c_str(), chain.size(), (long long)pairs_done, (long long)binaries, (long long)(ta + tr), (long long)ta, (long long)tr);
  return 0;
}

static void usage() {
  fprintf(stderr,
          "usage: tosem-scan scan   <project-root>... [--rows F] [--summary F] [--gpus N] [--all-files] [--batch-bytes N] [--rev-b]\n"
          "       tosem-scan reduce <taxonomy.csv> [--strategy F] [--methods F] [--properties F] [--correlate F] [--correlate-tex F] [--correlate-counts F] [--correlate-merged F]\n"
          "       tosem-scan diff   <old-root> <new-root> [--out F]\n"
          "       tosem-scan body   <project-root>... [--out F]\n"
          "       tosem-scan releases <snapshot-root>=<tag>... [--out F]   |   releases --git <repository> [<revision>...] [--out F]\n"
          "       tosem-scan history <git-repository> [--rev R] [--max-commits N] [--all-files] [--dry-run] [--out F]\n"
          "Scans run on the GPU through libtosemscan.so (sm_100a); there is no CPU fallback.\n");
}

int main(int argc, char** argv) {
  if (argc < 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) { usage(); return argc < 2 ? 2 : 0; }
  const std::string cmd = argv[1];
  std::vector<std::string> pos;
  std::map<std::string, std::string> opt;
  bool all_files = false, rev_b = false, dry_run = false;
  for (int i = 2; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--all-files") all_files = true;
    else if (a == "--rev-b") rev_b = true;
    else if (a == "--dry-run") dry_run = true;
    else if (a.rfind("--", 0) == 0) { if (i + 1 >= argc) die("missing value for " + a); opt[a] = argv[++i]; }
    else pos.push_back(a);
  }
  if (cmd == "scan") { if (pos.empty()) die("scan needs at least one project root"); return cmd_scan(pos, opt["--rows"], opt["--summary"], opt.count("--gpus") ? atoi(opt["--gpus"].c_str()) : 1, all_files,
                                        opt.count("--batch-bytes") ? std::max<int64_t>(4096, atoll(opt["--batch-bytes"].c_str())) : (1ll << 30), rev_b); }
  if (cmd == "reduce") { if (pos.size() != 1) die("reduce needs the taxonomy csv"); return cmd_reduce(pos[0], opt["--strategy"], opt["--methods"], opt["--properties"], opt["--correlate"], opt["--correlate-tex"], opt["--correlate-counts"], opt["--correlate-merged"]); }
  if (cmd == "releases") { if (pos.empty() && !opt.count("--git")) die("releases needs <root>=<tag>... or --git <repository>"); return cmd_releases(pos, opt["--out"], opt["--git"]); }
  if (cmd == "body") { if (pos.empty()) die("body needs at least one project root"); return cmd_body(pos, opt["--out"]); }
  if (cmd == "history") { if (pos.size() != 1) die("history needs the repository"); return cmd_history(pos[0], opt.count("--rev") ? opt["--rev"] : "HEAD",
                                                  opt.count("--max-commits") ? atoll(opt["--max-commits"].c_str()) : 0, all_files, opt["--out"], dry_run); }
  if (cmd == "diff") { if (pos.size() != 2) die("diff needs <old-root> <new-root>"); return cmd_diff(pos[0], pos[1], opt["--out"]); }
  usage();
  return 2;
}
