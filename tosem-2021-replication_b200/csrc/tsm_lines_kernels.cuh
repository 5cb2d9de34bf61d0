// tsm_lines_kernels.cuh - S9 line records in file order (docs/SPEC.md section 3): the small kernels around
// k_scan's TSM_SCAN_LINE_HASHES output.
//
// k_scan hands out work units in any order, so a chunk cannot know the index of its first line.  Instead every
// chunk writes the records of its own lines, in order, into one region of the staging arrays and notes where
// (unit_out) and how many (unit_lines).  With the unit table in (file, chunk) order - the deterministic plan below -
// one exclusive scan of unit_lines gives every unit's first line index, and one gather puts the records in place.
// The source bytes are read once; the records (13 B per line) are read and written once more.
#pragma once
#include "tsm_device.cuh"

namespace tsm {

constexpr uint32_t XS_TILE = 1024;                       // items per block of the exclusive scan (256 threads x 4)

__global__ void k_file_units(const int32_t* len, uint32_t n, uint32_t* cnt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt[i] = ((uint32_t)len[i] + CH - 1) / CH;
}

__device__ __forceinline__ unsigned long long block_sum(unsigned long long v, unsigned long long* sh) {   // 256 threads
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long t = 0;
  for (int w = 0; w < 8; ++w) t += sh[w];
  __syncthreads();
  return t;
}

// Exclusive scan of n u32 items into n + 1 u64 (out[n] = total): tile sums, scan of the sums, apply.
__global__ void __launch_bounds__(256) k_xscan_sums(const uint32_t* in, uint32_t n, unsigned long long* bsum) {
  __shared__ unsigned long long sh[8];
  const uint32_t i0 = blockIdx.x * XS_TILE + threadIdx.x * 4u;
  unsigned long long s = 0;
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) if (i0 + k < n) s += in[i0 + k];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) bsum[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_xscan_top(unsigned long long* bsum, uint32_t nb) {   // one block; bsum[nb] = total
  __shared__ unsigned long long sh[8];
  __shared__ unsigned long long wsum[8];
  unsigned long long carry = 0;
  for (uint32_t base = 0; base < nb; base += 256) {
    const uint32_t i = base + threadIdx.x;
    const unsigned long long v = i < nb ? bsum[i] : 0ull;
    unsigned long long incl = v;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    unsigned long long off = carry;
    for (int k = 0; k < w; ++k) off += wsum[k];
    if (i < nb) bsum[i] = off + incl - v;
    const unsigned long long tile = block_sum(v, sh);
    carry += tile;
  }
  if (threadIdx.x == 0) bsum[nb] = carry;
}
__global__ void __launch_bounds__(256) k_xscan_apply(const uint32_t* in, uint32_t n, const unsigned long long* bsum, unsigned long long* out) {
  __shared__ unsigned long long wsum[8];
  const uint32_t i0 = blockIdx.x * XS_TILE + threadIdx.x * 4u;
  uint32_t v[4];
  unsigned long long s = 0;
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) { v[k] = i0 + k < n ? in[i0 + k] : 0u; s += v[k]; }
  unsigned long long incl = s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  unsigned long long off = bsum[blockIdx.x] + incl - s;
  for (int k = 0; k < w; ++k) off += wsum[k];
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) { if (i0 + k < n) out[i0 + k] = off; off += v[k]; }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = bsum[gridDim.x];
}

// Deterministic unit table: units in (file, chunk) order.  unit_first[f] = first unit of file f.
__global__ void k_plan_det(ScanParams p, const unsigned long long* unit_first) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= p.n_files) {
    if (f == p.n_files) p.slab->n_units = (uint32_t)unit_first[p.n_files];
    return;
  }
  const uint32_t nu = ((uint32_t)p.len[f] + CH - 1) / CH;
  if (nu != 1) p.stats[f] = tsm_file_stat{0, 0, 0, 0, 0};
  uint32_t at = (uint32_t)unit_first[f];
  for (uint32_t u = 0; u < nu; ++u, ++at) {
    if (at < p.unit_cap) { p.unit_file[at] = (uint32_t)f; p.unit_begin[at] = u * CH; }
    else p.ctrl->overflow = 1;
  }
}

// Records of every unit from its staging region to their place in file order; one warp per unit.
__global__ void k_gather_lines(const uint32_t* unit_lines, const uint32_t* unit_out, const unsigned long long* unit_line_base,
                               uint32_t n_units, const unsigned long long* s_hash, const uint32_t* s_end, const uint8_t* s_flag,
                               unsigned long long* hash, uint32_t* end, uint8_t* flag) {
  const uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (u >= n_units) return;
  const uint32_t n = unit_lines[u], src = unit_out[u];
  const unsigned long long dst = unit_line_base[u];
  for (uint32_t i = lane; i < n; i += 32) {
    hash[dst + i] = s_hash[src + i];
    end[dst + i] = s_end[src + i];
    flag[dst + i] = s_flag[src + i];
  }
}

__global__ void k_line_base(const unsigned long long* unit_first, const unsigned long long* unit_line_base, uint32_t n_files,
                            unsigned long long* line_base) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f <= n_files) line_base[f] = unit_line_base[unit_first[f]];
}

// n-gram hashes (SPEC section 3): window of up to n consecutive line hashes of one file, starting at every line.
__global__ void k_ngrams(const unsigned long long* hash, const unsigned long long* line_base, uint32_t n_files,
                         unsigned long long total, uint32_t n, unsigned long long* out) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  uint32_t lo = 0, hi = n_files;                          // file of line i: line_base[lo] <= i < line_base[hi]
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (line_base[mid] <= i) lo = mid; else hi = mid; }
  const unsigned long long fend = line_base[lo + 1];
  unsigned long long acc = 0;
  uint32_t k = 0, r = 0;
  for (; k < n && i + k < fend; ++k) {
    acc = fold61(acc + rotl61(canon61(hash[i + k]), r));
    r += 13; if (r >= 61) r -= 61;
  }
  out[i] = mix_hash(canon61(acc), k);
}

}  // namespace tsm
