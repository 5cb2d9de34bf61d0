// tsm_stmt_kernels.cuh - body statements (docs/SPEC.md section 10, SURVEY.md section 8f item 1, golden G2).
// Pins: Important-files/ML-Analysis-v4.xlsx!Apollo:R2-R26 = src/apollo/v6.0.0/modules/common/math/
// aabox2d_test.cc:27-53 (multi-line statements joined while the parentheses are open).
//
//   k_scan (TSM_SCAN_LINE_HASHES) + tsm_lines_kernels.cuh   line records (ordered line ends) per file
//   k_line_parens   thread per line: SWAR count of '(' minus ')' and blank test            -> delta[], nonblank
//   k_stmt_kinds    warp per file: clamped running depth d' = max(0, d + delta) as a warp scan
//                   over the monoid f(d) = max(a, d + b); kind 1 = first line of a statement,
//                   2 = continuation, 0 = blank
#pragma once
#include "tsm_diff_kernels.cuh"

namespace tsm {

__device__ __forceinline__ uint32_t count_byte64(unsigned long long w, unsigned long long rep) {
  const unsigned long long x = w ^ rep, k7 = 0x7F7F7F7F7F7F7F7Full;
  return (uint32_t)__popcll(~(((x & k7) + k7) | x | k7));    // one 0x80 per matching byte
}

__global__ void k_line_parens(DiffSide d, int32_t n, unsigned long long total, int32_t* delta, uint8_t* kind) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  int lo = 0, hi = n;                                   // file of line i: line_base[lo] <= i < line_base[hi]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (d.line_base[mid] <= i) lo = mid; else hi = mid; }
  const uint8_t* g = d.arena + (uint32_t)d.off[lo];
  const uint32_t e = d.line_end[i];
  const uint32_t s = (i == d.line_base[lo]) ? 0u : d.line_end[i - 1] + 1u;
  bool nonblank = false;
  for (uint32_t q = s; q < e; ++q)
    if (!is_w(__ldg(g + q))) { nonblank = true; break; }
  int32_t dl = 0;
  if (nonblank) {
    for (uint32_t wb = s & ~7u; wb < e; wb += 8) {
      unsigned long long w = __ldg(reinterpret_cast<const unsigned long long*>(g + wb));
      // bytes outside [s, e) become 0 (neither parenthesis)
      if (wb < s) w &= ~0ull << (8u * (s - wb));
      if (wb + 8 > e) w &= ~0ull >> (8u * (wb + 8 - e));
      dl += (int32_t)count_byte64(w, 0x2828282828282828ull) - (int32_t)count_byte64(w, 0x2929292929292929ull);
    }
  }
  delta[i] = dl;
  kind[i] = nonblank ? 1 : 0;
}

__global__ void k_stmt_kinds(DiffSide d, int32_t n, const int32_t* delta, uint8_t* kind) {
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (f >= n) return;
  const unsigned long long b0 = d.line_base[f], b1 = d.line_base[f + 1];
  const long long NEG = -(1ll << 60);
  long long depth = 0;                                   // depth before the current group of 32 lines
  for (unsigned long long base = b0; base < b1; base += 32) {
    const unsigned long long i = base + lane;
    const bool valid = i < b1;
    const bool nb = valid && kind[i] != 0;
    long long a = nb ? 0 : NEG, b = nb ? (long long)delta[i] : 0;   // f(x) = max(a, x + b); identity for blanks
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {                   // inclusive scan, composing left to right
      const long long la = __shfl_up_sync(0xffffffffu, a, k), lb = __shfl_up_sync(0xffffffffu, b, k);
      if (lane >= k) { a = max(a, la + b); b = lb + b; }
    }
    const long long after = max(a, depth + b);
    long long before = __shfl_up_sync(0xffffffffu, after, 1);
    if (lane == 0) before = depth;
    if (valid) kind[i] = nb ? (before == 0 ? 1 : 2) : 0;
    depth = __shfl_sync(0xffffffffu, after, 31);
  }
}

}  // namespace tsm
