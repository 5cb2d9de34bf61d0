// tsm_reduce_kernels.cuh - S10 reduce (docs/SPEC.md section 9): distinct case ids per (flag, repo).
// Pins: RQs/taxonomy_test2.csv -> RQs/RQ3/tests_strategy_rq32.csv, RQs/RQ4/tests_methods_v2.csv
// (golden G3).  9 685 rows: a bitmap per (flag, repo) set with atomicOr, then one popcount pass.
#pragma once
#include "tsm_device.cuh"

namespace tsm {

// bits[(f * n_repos + r) * words + case/32]; f == 0 is the implicit "any row" flag.
__global__ void k_reduce_mark(const uint8_t* flags, const int32_t* repo, const int32_t* case_id, int32_t n_rows,
                              int32_t n_flags, int32_t n_repos, uint32_t words, uint32_t* bits) {
  const long long total = (long long)n_rows * (n_flags + 1);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int32_t row = (int32_t)(t / (n_flags + 1)), f = (int32_t)(t % (n_flags + 1));
    if (f && flags[(size_t)row * n_flags + (f - 1)] == 0) continue;
    const uint32_t c = (uint32_t)case_id[row];
    atomicOr(&bits[((size_t)f * n_repos + repo[row]) * words + (c >> 5)], 1u << (c & 31));
  }
}

// one warp per (flag, repo) cell
__global__ void k_reduce_count(const uint32_t* bits, int32_t cells, uint32_t words, unsigned long long* out) {
  const int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (cell >= cells) return;
  uint32_t n = 0;
  for (uint32_t w = lane; w < words; w += 32) n += __popc(bits[(size_t)cell * words + w]);
#pragma unroll
  for (int d = 16; d; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
  if (lane == 0) out[cell] = n;
}

inline int launch_reduce(const uint8_t* d_flags, const int32_t* d_repo, const int32_t* d_case, int32_t n_rows,
                         int32_t n_flags, int32_t n_repos, int32_t n_cases, uint32_t* d_bits,
                         unsigned long long* d_out, cudaStream_t st) {
  const uint32_t words = ((uint32_t)n_cases + 31) / 32;
  const int cells = (n_flags + 1) * n_repos;
  if (n_rows) {
    const long long total = (long long)n_rows * (n_flags + 1);
    const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
    k_reduce_mark<<<blocks, 256, 0, st>>>(d_flags, d_repo, d_case, n_rows, n_flags, n_repos, words, d_bits);
  }
  k_reduce_count<<<(cells * 32 + 255) / 256, 256, 0, st>>>(d_bits, cells, words, d_out);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace tsm
