// tsm_diff_kernels.cuh - S8 revision-pair churn (docs/SPEC.md section 8): per (old, new) pair the number of
// added / removed lines = |new| - LCS, |old| - LCS over the SPEC section 3 line-hash sequences.
// Pins: Important-files/ML-Testing-v1.xlsx!projects:R1 (`cloc = added + removed`); no revision
// history ships, so parity is against the oracle's O(n*m) DP (a different algorithm: here Myers O(ND)).
//
//   k_count_lines   warp per file: SWAR newline count                     -> n_lines[file]
//   k_mark_lines    warp per file: ordered newline positions (ballot compaction) -> line_end[]
//   k_hash_lines    thread per line: Mersenne-61 line hash from HBM       -> line_hash[]
//   k_myers         warp per pair: common prefix/suffix trim, then the greedy furthest-reaching
//                   D-path search with the diagonals of one D spread over the lanes
#pragma once
#include "tsm_scan_kernels.cuh"

namespace tsm {

struct DiffSide {                   // one corpus (old or new) on the device
  const uint8_t* arena; const int32_t* off; const int32_t* len;
  uint32_t* n_lines;                // [n]
  const unsigned long long* line_base;   // [n+1] exclusive prefix of n_lines
  uint32_t* line_end;               // [total lines] file-relative end of each line (position of its LF or EOF)
  unsigned long long* line_hash;    // [total lines]
  const uint8_t* ext;               // [n] S1 tags (NULL = all 0), only read when line_flag != NULL
  uint8_t* line_flag;               // [total lines] 1 = assertion line (SPEC section 4); NULL = not wanted
};

// Edit distance D (insertions + deletions) of hash sequences a[0..n) and b[0..m); one warp per pair.
// V (furthest x per diagonal) lives in global scratch of 2*(n+m)+3 ints per pair.
__global__ void k_myers(const unsigned long long* ha, const unsigned long long* la, const unsigned long long* hb,
                        const unsigned long long* lb, int32_t n_pairs, int32_t* vbuf, const unsigned long long* vbase,
                        long long* added, long long* removed) {
  const int pr = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (pr >= n_pairs) return;
  const unsigned long long* a = ha + la[pr];
  const unsigned long long* b = hb + lb[pr];
  int n = (int)(la[pr + 1] - la[pr]), m = (int)(lb[pr + 1] - lb[pr]);
  const int n0 = n, m0 = m;
  // common prefix
  int pre = 0;
  while (true) {
    const int i = pre + lane;
    const bool ne = !(i < n && i < m && a[i] == b[i]);
    const uint32_t mk = __ballot_sync(0xffffffffu, ne);
    if (mk) { pre += __ffs(mk) - 1; break; }
    pre += 32;
  }
  a += pre; b += pre; n -= pre; m -= pre;
  int suf = 0;
  while (true) {
    const int i = suf + lane;
    const bool ne = !(i < n && i < m && a[n - 1 - i] == b[m - 1 - i]);
    const uint32_t mk = __ballot_sync(0xffffffffu, ne);
    if (mk) { suf += __ffs(mk) - 1; break; }
    suf += 32;
  }
  n -= suf; m -= suf;
  int D = 0;
  if (n == 0 || m == 0) D = n + m;
  else {
    int32_t* V = vbuf + vbase[pr] + (n + m + 1);          // V[k], k in [-(n+m)-1, n+m+1]
    if (lane == 0) V[1] = 0;
    __syncwarp();
    bool done = false;
    for (D = 0; D <= n + m && !done; ++D) {
      bool hit = false;
      for (int k = -D + 2 * lane; k <= D; k += 64) {
        int x;
        if (k == -D || (k != D && V[k - 1] < V[k + 1])) x = V[k + 1]; else x = V[k - 1] + 1;
        int y = x - k;
        while (x < n && y < m && a[x] == b[y]) { ++x; ++y; }
        V[k] = x;
        if (x >= n && y >= m) hit = true;
      }
      __syncwarp();
      done = __any_sync(0xffffffffu, hit);
    }
    --D;                                                  // the loop increments once more after the hit
  }
  if (lane == 0) {
    const long long lcs = ((long long)(n + m) - D) / 2 + pre + suf;
    removed[pr] = n0 - lcs;
    added[pr] = m0 - lcs;
  }
}

// Hunks and their classification (docs/SPEC.md section 8): the same search with one row of V kept per D
// (row d holds the diagonals -d, -d+2, ..., d), then the canonical backtrack by lane 0.
// trace_base[pr] = first int of pair pr's rows, sized (D+1)(D+2)/2 from the distances of k_myers.
__global__ void k_myers_trace(const unsigned long long* ha, const unsigned long long* la, const uint8_t* fa,
                              const unsigned long long* hb, const unsigned long long* lb, const uint8_t* fb,
                              int32_t pair0, int32_t n_pairs, int32_t* trace, const unsigned long long* trace_base,
                              const long long* added, const long long* removed, long long max_d, tsm_diff_detail* detail) {
  const int pr = pair0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (pr >= pair0 + n_pairs) return;
  if (added[pr] + removed[pr] > max_d) return;             // too far apart to keep the rows of V: the host reports one hunk
  const unsigned long long* a = ha + la[pr];
  const unsigned long long* b = hb + lb[pr];
  const uint8_t* qa = fa + la[pr];
  const uint8_t* qb = fb + lb[pr];
  int n = (int)(la[pr + 1] - la[pr]), m = (int)(lb[pr + 1] - lb[pr]);
  int pre = 0;
  while (true) {
    const int i = pre + lane;
    const uint32_t mk = __ballot_sync(0xffffffffu, !(i < n && i < m && a[i] == b[i]));
    if (mk) { pre += __ffs(mk) - 1; break; }
    pre += 32;
  }
  a += pre; b += pre; qa += pre; qb += pre; n -= pre; m -= pre;
  int suf = 0;
  while (true) {
    const int i = suf + lane;
    const uint32_t mk = __ballot_sync(0xffffffffu, !(i < n && i < m && a[n - 1 - i] == b[m - 1 - i]));
    if (mk) { suf += __ffs(mk) - 1; break; }
    suf += 32;
  }
  n -= suf; m -= suf;
  long long h_add = 0, h_del = 0, h_mod = 0, a_as = 0, r_as = 0;
  if (n == 0 || m == 0) {                                 // one pure hunk (or none)
    int ca = 0, cb = 0;
    for (int i = lane; i < n; i += 32) ca += qa[i] != 0;
    for (int i = lane; i < m; i += 32) cb += qb[i] != 0;
#pragma unroll
    for (int k = 16; k; k >>= 1) { ca += __shfl_xor_sync(0xffffffffu, ca, k); cb += __shfl_xor_sync(0xffffffffu, cb, k); }
    if (n) { h_del = 1; r_as = ca; }
    if (m) { h_add = 1; a_as = cb; }
  } else {
    int32_t* R = trace + trace_base[pr];
    int D = 0;
    bool done = false;
    for (int d = 0; !done; ++d) {
      int32_t* row = R + (size_t)d * (d + 1) / 2;
      const int32_t* P = R + (size_t)(d - 1) * d / 2;     // previous row, entry (kk + d - 1) / 2
      bool hit = false;
      for (int k = -d + 2 * lane; k <= d; k += 64) {
        int x;
        if (d == 0) x = 0;
        else {
          const bool down = (k == -d) || (k != d && P[(k - 1 + d - 1) / 2] < P[(k + 1 + d - 1) / 2]);
          x = down ? P[(k + 1 + d - 1) / 2] : P[(k - 1 + d - 1) / 2] + 1;
        }
        int y = x - k;
        while (x < n && y < m && a[x] == b[y]) { ++x; ++y; }
        row[(k + d) / 2] = x;
        if (x >= n && y >= m) hit = true;
      }
      __syncwarp();
      done = __any_sync(0xffffffffu, hit);
      D = d;
    }
    if (lane == 0) {                                      // canonical backtrack: edits from the last to the first
      int x = n, y = m;
      bool in_hunk = false, has_add = false, has_del = false;
      for (int d = D; d >= 1; --d) {
        const int k = x - y;
        const int32_t* P = R + (size_t)(d - 1) * d / 2;
        const bool down = (k == -d) || (k != d && P[(k - 1 + d - 1) / 2] < P[(k + 1 + d - 1) / 2]);
        const int pk = down ? k + 1 : k - 1;
        const int px = P[(pk + d - 1) / 2], py = px - pk;
        const int midx = down ? px : px + 1;
        if (in_hunk && x - midx > 0) {
          if (has_add && has_del) ++h_mod; else if (has_add) ++h_add; else ++h_del;
          has_add = has_del = false;
        }
        in_hunk = true;
        if (down) { has_add = true; a_as += qb[py] != 0; }
        else { has_del = true; r_as += qa[px] != 0; }
        x = px; y = py;
      }
      if (in_hunk) { if (has_add && has_del) ++h_mod; else if (has_add) ++h_add; else ++h_del; }
    }
  }
  if (lane == 0) detail[pr] = tsm_diff_detail{h_add, h_del, h_mod, a_as, r_as};
}

}  // namespace tsm
