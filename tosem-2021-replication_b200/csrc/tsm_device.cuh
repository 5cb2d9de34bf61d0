// tsm_device.cuh - shared device-side definitions of the sm_100a corpus-scan kernels.
//
// Data layout in HBM (docs/SPEC.md section 1, DESIGN.md "layout"):
//   arena      u8   files end-to-end, each start 128-B aligned (cp.async.bulk needs 16 B)
//   off/len    i32  per-file start and size          ext u8, grp u16 per-file tags
//   unit_file / unit_begin  u32  work units = (file, 4 KiB chunk), built by k_plan (or k_plan_det, in file order)
//   stats      {u32 lines, asserts, headers, fixtures; u64 digest} per file
//   cand       u64  (file << 32 | line offset) of every assertion line, consumed by k_classify
//   counts     i64  [n_groups + 1][128]   (row n_groups = global)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/tosemscan.h"
#include "../../include/tsm_categories.h"

namespace tsm {

// ---- chunk geometry -------------------------------------------------------------------------
constexpr uint32_t CH = 4096;                 // bytes of a file owned by one work unit
constexpr uint32_t PRE = 16;                  // bytes loaded in front (is the chunk start a line start?)
constexpr uint32_t EXT = 240;                 // bytes loaded behind (terminator of the last owned line)
constexpr uint32_t BUF = PRE + CH + EXT;      // 4352 = 34 * 128 = 32 * 136
constexpr uint32_t STRIPE = BUF / 32;         // 136 = 8 * 17: stride of conflict-free per-lane LDS.64
static_assert(STRIPE * 32 == BUF && STRIPE % 8 == 0 && (STRIPE / 8) % 2 == 1, "stripe geometry");
constexpr uint32_t LUT_BYTES = 1024 + 128;    // the 256-entry automaton table + the per-language pattern-end masks

// ---- multi-pattern Shift-And automaton (SPEC sections 4, 5) ------------------------------------------
// One state bit per pattern byte; D' = ((D << 1) | FIRST) & LUT[c]; a line's OR of all D tells which
// patterns ended somewhere inside it.  ONE table for all languages (its address is a compile-time
// constant); the language only decides which pattern ends count:
//   bits  0.. 5  assert (ci)    bits  6..12  EXPECT_ (cs)   bits 13..17  class     bits 18..20  def
//   bits 21..24  test (ci)      bits 25..28  void           bit  29      {         bit  30      F (gate of the TEST_F check)
//   bit  31      '\n' (the OR of a word's states says whether the word holds a newline)
constexpr uint32_t B_FIRST = (1u << 0) | (1u << 6) | (1u << 13) | (1u << 18) | (1u << 21) | (1u << 25) | (1u << 29) | (1u << 30) | (1u << 31);
constexpr uint32_t AF_ASSERT = 1u << 5, AF_EXPECT = 1u << 12, A_CLASS = 1u << 17, A_DEF = 1u << 20, A_TEST = 1u << 24,
                   A_VOID = 1u << 28, A_BRACE = 1u << 29, B_F = 1u << 30;
// header patterns per language family: group 1, group 2 (SPEC section 5)
constexpr uint32_t PY_G1 = A_DEF, PY_G2 = A_CLASS, CJ_G1 = A_TEST, CJ_G2 = A_BRACE | A_CLASS | A_VOID;

// flags of a finished line
constexpr uint8_t LF_CAND = 1, LF_HDR = 2, LF_FIX = 4;

struct Ctrl {                     // device control block, zeroed before every scan
  uint32_t n_cand;                // candidates appended by k_scan
  uint32_t n_hev;
  uint32_t n_aev;
  uint32_t overflow;              // some list hit its capacity
  uint32_t n_lh;                  // line-record slots reserved by k_scan (TSM_SCAN_LINE_HASHES)
  uint32_t lh_overflow;           // ... and whether the staging arrays were too small for them
  uint32_t cls_done;              // candidates already classified (streamed scans classify slab by slab); set by k_plan
};

// tsm_diff_pairs_detail keeps (D+1)(D+2)/2 ints per pair for the backtrack: at most 2^28 (1 GiB), i.e. D <= 23 168
constexpr unsigned long long TSM_DIFF_TRACE_MAX_INTS = 1ull << 28;
constexpr long long TSM_DIFF_TRACE_MAX_D = 23168;
static_assert((TSM_DIFF_TRACE_MAX_D + 1) * (TSM_DIFF_TRACE_MAX_D + 2) / 2 <= (long long)TSM_DIFF_TRACE_MAX_INTS && (TSM_DIFF_TRACE_MAX_D + 2) * (TSM_DIFF_TRACE_MAX_D + 3) / 2 > (long long)TSM_DIFF_TRACE_MAX_INTS, "cap");

struct SlabCtl { uint32_t n_units, work; };   // per-slab unit count (k_plan) and work cursor (k_scan)

struct ScanParams {
  const uint8_t* arena;
  const int32_t* off;
  const int32_t* len;
  const uint8_t* ext;
  const uint16_t* grp;
  int32_t n_files;
  int32_t n_groups;
  uint32_t* unit_file;
  uint32_t* unit_begin;
  uint32_t unit_cap;
  SlabCtl* slab;                  // this launch's slab
  int32_t f_begin, f_end;         // files of this slab
  uint32_t unit_base;             // first slot of this slab in the unit table
  Ctrl* ctrl;
  tsm_file_stat* stats;
  unsigned long long* cand;
  uint32_t cand_cap;
  tsm_header_event* hev;
  uint32_t hev_cap;
  tsm_assert_event* aev;
  uint32_t aev_cap;
  unsigned long long* counts;     // [n_groups + 1][K]
  uint32_t flags;
  // per-line output (TSM_SCAN_LINE_HASHES; docs/SPEC.md section 3, S9): every chunk reserves one region of the staging
  // arrays and writes the records of its own lines in order; unit_out / unit_lines say where and how many
  unsigned long long* lh_hash;    // line_hash
  uint32_t* lh_end;               // file-relative end of the line (position of its LF, or the file size)
  uint8_t* lh_flag;               // 1 = assertion line (SPEC section 4)
  uint32_t lh_cap;
  uint32_t* unit_lines;           // [unit slots]
  uint32_t* unit_out;             // [unit slots]
  uint32_t four;                  // 4 (a multiplier the compiler must not see: tsm_scan_walk.cuh, lut_at)
  uint32_t cls_last;              // k_classify: 1 = the last launch of the scan (adds the totals of the per-file records)
};

// ---- hashing (SPEC section 3) -------------------------------------------------------------------------
constexpr unsigned long long M61 = 0x1FFFFFFFFFFFFFFFull;

__device__ __forceinline__ unsigned long long fold61(unsigned long long x) { return (x & M61) + (x >> 61); }
__device__ __forceinline__ unsigned long long rotl61(unsigned long long x, uint32_t r) {  // x < 2^61, r < 61
  return ((x << r) & M61) | (x >> (61 - r));   // r == 0: x >> 61 == 0 for x < 2^61
}
__device__ __forceinline__ unsigned long long canon61(unsigned long long acc) {
  acc = fold61(fold61(acc));
  return acc == M61 ? 0ull : acc;
}
__device__ __forceinline__ unsigned long long mix_hash(unsigned long long h61, unsigned long long len) {
  unsigned long long x = h61 ^ (len * 0x9E3779B97F4A7C15ull);
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

__device__ __forceinline__ bool is_w(uint32_t c) { return c == 0x20 || c == 0x09 || c == 0x0D || c == 0x0B || c == 0x0C; }
__device__ __forceinline__ bool is_ident(uint32_t c) {
  return (c - 'a') < 26u || (c - 'A') < 26u || (c - '0') < 10u || c == '_';
}

// ---- mbarrier / bulk-copy (TMA 1-D) PTX ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy accesses of shared memory are ordered in front of later async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

}  // namespace tsm
