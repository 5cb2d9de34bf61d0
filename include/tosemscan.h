/* tosemscan.h - C ABI of libtosemscan.so, the B200-native corpus-scan hot path.
 *
 * Drop-in boundary.  The reference package (openjamoses/TOSEM-2021-Replication) ships NO code, hence
 * no plugin / operator / FFI interface to mirror: the only observable contract is
 * its file formats.  Each entry point below therefore cites the reference ARTEFACT whose
 * producing stage it replaces; docs/SPEC.md gives the byte-level rules, INTEGRATION.md the
 * bindings (ctypes / C++ CLI) a maintainer would add.
 *
 * Conventions: plain pointers and sizes, no exceptions, no global state; every call returns
 * TSM_OK (0) or a negative tsm_status; caller owns all host memory; a tsm_ctx owns device memory
 * and is bound to one CUDA device; calls on one ctx must be serialised by the caller, calls on
 * different ctxs are independent (one host thread / process per GPU).  `stream` is a cudaStream_t
 * passed as void* (NULL = the legacy default stream).  There is no CPU fallback: without a usable
 * CUDA device tsm_create() fails with TSM_E_CUDA.
 */
#ifndef TOSEMSCAN_H
#define TOSEMSCAN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TSM_ABI_VERSION 1
#define TSM_NUM_CATEGORIES 128   /* K of docs/SPEC.md section 6 */
#define TSM_ALIGN 128            /* file start alignment inside the arena */

typedef enum {
  TSM_OK = 0,
  TSM_E_ARG = -1,        /* null / negative / inconsistent argument */
  TSM_E_LAYOUT = -2,     /* corpus violates SPEC section 1 (alignment, bounds, grp >= n_groups) */
  TSM_E_CAPACITY = -3,   /* corpus or event count exceeds what the ctx was created for */
  TSM_E_CUDA = -4,       /* CUDA runtime error (no device, launch failure, ...) */
  TSM_E_NOMEM = -5,
  TSM_E_STATE = -6       /* call order (e.g. scan_resident before upload) */
} tsm_status;

/* ext tags (S1: value census of the `extension` column, Important-files/ML-Testing-v1.xlsx) */
enum { TSM_EXT_OTHER = 0, TSM_EXT_PY = 1, TSM_EXT_CC = 2, TSM_EXT_CPP = 3, TSM_EXT_JAVA = 4, TSM_EXT_C = 5, TSM_EXT_H = 6 };

/* scan flags */
#define TSM_SCAN_ASSERT_EVENTS 1u  /* produce assertion events (raw scan rows need them) */
#define TSM_SCAN_HEADER_EVENTS 2u  /* produce header events (method column needs them) */
#define TSM_SCAN_REV_B 8u          /* the later revision of the lost tool (docs/SPEC.md section 4b; golden G1: ML-Testing-v1.xlsx!DeepSpeech):
                                      also triggers on _CHECK / TESTEQUAL / FAIL, full statements for BOOST_CHECK( / NTA_CHECK( / Java */
#define TSM_SCAN_LINE_HASHES 4u    /* internal to tsm_line_hashes / tsm_diff_pairs*: per-line records; ignored by tsm_scan* */

/* Per-file record; replaces the per-file summary stage (S6: `total assert` of
 * selection/completed-labels/Release-Meta-tpot.csv:1-2). */
typedef struct tsm_file_stat {
  uint32_t n_lines, n_assert, n_headers, n_fixture;
  uint64_t digest;
} tsm_file_stat;

/* One assertion line; the raw-row stage (fileName,extension,test_name,method,statement,counts,
 * category: Important-files/ML-Testing-v1.xlsx!apollo_tests:R1) groups these. Offsets are
 * relative to the file start. */
typedef struct tsm_assert_event {
  uint32_t file, line_off, stmt_off;
  uint16_t stmt_len, cat;
  uint32_t ident_off;
  uint16_t ident_len, pad;
  uint64_t stmt_hash;
} tsm_assert_event;

typedef struct tsm_header_event { uint32_t file, line_off, line_len, kind; } tsm_header_event; /* kind bit0 = TEST_F */

/* Packed corpus (SPEC section 1). All pointers host memory for the host-path calls. */
typedef struct tsm_corpus {
  const uint8_t* arena;   /* off[n_files] bytes */
  const int32_t* off;     /* [n_files+1], multiples of TSM_ALIGN, ascending */
  const int32_t* len;     /* [n_files], off[i]+len[i] <= off[i+1] */
  const uint8_t* ext;     /* [n_files] TSM_EXT_* */
  const uint16_t* grp;    /* [n_files] < n_groups; NULL = all 0 */
  int32_t n_files;
  int32_t n_groups;       /* >= 1 */
} tsm_corpus;

/* Host-side result buffers. Any pointer may be NULL (that output is skipped). */
typedef struct tsm_result {
  tsm_file_stat* stats;          /* [n_files] */
  int64_t* group_counts;         /* [n_groups][TSM_NUM_CATEGORIES] */
  int64_t* global_counts;        /* [TSM_NUM_CATEGORIES] */
  tsm_assert_event* aev; int64_t aev_cap; int64_t n_aev;   /* canonical order (file, line_off) */
  tsm_header_event* hev; int64_t hev_cap; int64_t n_hev;
  int64_t totals[4];             /* lines, assertion lines, headers, fixture headers */
} tsm_result;

typedef struct tsm_ctx tsm_ctx;

int tsm_abi_version(void);
const char* tsm_strerror(int status);
const char* tsm_category_name(int id);   /* "" for 0/reserved, "<other>" for 127 */

/* Context sized for corpora up to (max_arena_bytes, max_files, max_groups) and max_events events
 * of each kind (0 = derive from max_arena_bytes). */
int tsm_create(tsm_ctx** out, int device, int64_t max_arena_bytes, int32_t max_files,
               int32_t max_groups, int64_t max_events);
void tsm_destroy(tsm_ctx* ctx);

/* Host path, end to end: H2D of the corpus (overlapped slab by slab with the scan), the scan and
 * classify/aggregate kernels, D2H of the requested results, stream-synchronised on return.
 * Replaces stage (C) "corpus scan" of SURVEY.md section 1 for one packed batch. */
int tsm_scan(tsm_ctx* ctx, const tsm_corpus* corpus, tsm_result* result, uint32_t flags, void* stream);

/* Resident path (what bench.py's `value` times): upload once, scan many times, fetch once. */
int tsm_upload(tsm_ctx* ctx, const tsm_corpus* corpus, void* stream);
int tsm_scan_resident(tsm_ctx* ctx, uint32_t flags, void* stream);       /* kernels only, async */
int tsm_download(tsm_ctx* ctx, tsm_result* result, void* stream);        /* synchronises */
/* Device address of the [n_groups+1][K] int64 count table of the last scan (row n_groups = global),
 * for the single multi-GPU allreduce; valid until the next scan. */
int tsm_device_counts(tsm_ctx* ctx, void** dptr, int64_t* n_int64);
/* Kernel launches issued by the last tsm_scan / tsm_scan_resident call. */
int tsm_last_launch_count(tsm_ctx* ctx);
/* Device time (CUDA events on the launching stream) of the kernels of the last scan, in launch
 * order: k_plan, k_scan, k_classify (+ a 4th slot that is 0: the totals are fused into k_classify).
 * Synchronises on the last of them. */
int tsm_last_kernel_ms(tsm_ctx* ctx, float* ms4);
/* Per-kernel device time summed over every scan since the last reset (event ring, no host sync
 * inside a back-to-back series); waits for scans still in flight. */
int tsm_kernel_ms_stats(tsm_ctx* ctx, double* sum_ms4, int64_t* n_scans, int reset);

/* S8 revision-pair churn (Important-files/ML-Testing-v1.xlsx!projects:R1 `cloc, added, removed`):
 * pair i = (olds file i, news file i); added/removed [n_pairs] host arrays. */
int tsm_diff_pairs(tsm_ctx* ctx, const tsm_corpus* olds, const tsm_corpus* news,
                   int64_t* added, int64_t* removed, void* stream);

/* The same plus, per pair, the hunks of the canonical edit script (docs/SPEC.md section 8) classified as
 * add / del / mod, and how many inserted / deleted lines are assertion lines (BASELINE config C5,
 * "per-hunk diff + classify").  `ext` of both corpora is used for the assertion rule. */
typedef struct tsm_diff_detail { int64_t hunks_add, hunks_del, hunks_mod, added_assert, removed_assert; } tsm_diff_detail;
int tsm_diff_pairs_detail(tsm_ctx* ctx, const tsm_corpus* olds, const tsm_corpus* news,
                          int64_t* added, int64_t* removed, tsm_diff_detail* detail, void* stream);

/* A pair whose edit distance D is larger than 23 168 lines needs more than 2^28 trace entries ((D+1)(D+2)/2 ints, 1 GiB)
 * for its backtrack: tsm_diff_pairs_detail does not trace it - added / removed are exact, the detail reports ONE hunk
 * (add, del or mod by the counts) and added_assert = removed_assert = -1.  Every other pair of the call is unaffected.
 *
 * Resident variant (bench.py's `value` for config C5): tsm_diff_upload copies both sides to HBM once and keeps them in
 * the ctx; every tsm_diff_resident call runs the kernels over them (k_scan over both sides for the line records,
 * k_diff_small - search, rows of V and backtrack of a pair in shared memory, four launches for four sizes of pairs - then
 * k_myers and, when detail != NULL, k_myers_trace for the pairs they leave over: distances above 127 lines, changed
 * regions of more than 4 096 lines) and copies the
 * per-pair results back.  tsm_diff_last_ms: device time (CUDA events on the launching stream) of the last diff call:
 * ms3 = { k_scan over both sides, k_diff_small, k_myers + k_myers_trace of the left-over pairs }. */
int tsm_diff_upload(tsm_ctx* ctx, const tsm_corpus* olds, const tsm_corpus* news, void* stream);
int tsm_diff_resident(tsm_ctx* ctx, int64_t* added, int64_t* removed, tsm_diff_detail* detail, void* stream);
int tsm_diff_last_ms(tsm_ctx* ctx, float* ms3);

/* S9 line / n-gram hashes (docs/SPEC.md section 3; SURVEY.md section 8a S9 - a design choice of the north star, attested by no
 * artefact of the package): the records of every line of every file, files in order, from ONE pass of the scan
 * kernel over the source.  line_base[n_files+1] and *n_lines are always filled; line_hash (SPEC section 3), line_end
 * (file-relative position of the line's LF, or the file size), line_flag (1 = assertion line, SPEC section 4, by the
 * file's ext) and ngram_hash hold `cap` lines each and may be NULL; if cap < *n_lines the call returns TSM_E_CAPACITY
 * so that the caller can size the arrays and call again.  ngram_hash[i] = hash of the window of up to ngram_n
 * consecutive lines of the same file that starts at line i.  tsm_diff_pairs* and tsm_statements use the same pass. */
int tsm_line_hashes(tsm_ctx* ctx, const tsm_corpus* corpus, int64_t* line_base, uint64_t* line_hash, uint32_t* line_end,
                    uint8_t* line_flag, int64_t cap, int64_t* n_lines, int32_t ngram_n, uint64_t* ngram_hash, void* stream);

/* Body statements (docs/SPEC.md section 10; Important-files/ML-Analysis-v4.xlsx!Apollo:R2-R26, golden G2): the
 * kind of every line of every file - 0 blank, 1 first line of a statement, 2 continuation (lines
 * are joined while the parentheses are open).  line_base[n_files+1] and *n_lines are always filled;
 * line_end (file-relative end of each line) and line_kind hold `cap` lines: if cap < *n_lines the call
 * returns TSM_E_CAPACITY so that the caller can size the arrays and call again. */
int tsm_statements(tsm_ctx* ctx, const tsm_corpus* corpus, int64_t* line_base, uint32_t* line_end,
                   uint8_t* line_kind, int64_t cap, int64_t* n_lines, void* stream);

/* S10 reduce (RQs/taxonomy_test2.csv -> RQs/RQ3/tests_strategy_rq32.csv, RQs/RQ4/
 * tests_methods_v2.csv): out[f*n_repos+r] = distinct case ids with flags[row*n_flags+f] != 0 in
 * repo r; cases_per_repo[r] = distinct case ids of repo r. Host arrays; integer work on device. */
int tsm_reduce(tsm_ctx* ctx, const uint8_t* flags, const int32_t* repo, const int32_t* case_id,
               int32_t n_rows, int32_t n_flags, int32_t n_repos, int32_t n_cases,
               int64_t* out, int64_t* cases_per_repo, void* stream);

/* ---- host-only helpers (no CUDA context needed) ------------------------------------------- */
/* Pinned host memory for arenas/results (cudaHostAlloc); returns NULL on failure. */
void* tsm_host_alloc(int64_t bytes);
void tsm_host_free(void* p);
/* Arena size (multiple of TSM_ALIGN) for files of the given sizes; fills off[n+1]. <0 on overflow. */
int64_t tsm_layout(const int32_t* len, int32_t n_files, int32_t* off);
/* Deterministic synthetic corpus (std::mt19937_64, one engine per file).
 * size_law 0: every file exactly fixed_size bytes; 1: truncated power law pdf ~ x^-1.5 on
 * [128 B, 1 MiB] rounded up to whole lines.  Slot i holds logical file first_index + i*index_stride,
 * so ranks generate disjoint round-robin shards of one logical corpus.  Two steps: tsm_gen_sizes
 * -> tsm_layout -> tsm_gen_fill. */
int tsm_gen_sizes(uint64_t seed, int32_t n_files, int size_law, int32_t fixed_size, int32_t first_index,
                  int32_t index_stride, int32_t* len, uint8_t* ext, uint16_t* grp, int32_t n_groups);
int tsm_gen_fill(uint64_t seed, int32_t n_files, int size_law, int32_t first_index, int32_t index_stride,
                 const int32_t* off, const int32_t* len, const uint8_t* ext, uint8_t* arena);
/* new = old with Poisson(lambda) line edits (insert/delete/replace of Geometric(0.4) runs). Returns
 * bytes written to dst (<= cap) or <0. */
int64_t tsm_gen_edit(uint64_t seed, const uint8_t* src, int32_t src_len, double lambda,
                     uint8_t* dst, int64_t cap);

/* BASELINE config C5: n (old, new) revision pairs; old ~ the size law 1 with the target clamped to cap bytes,
 * new = old with Poisson(lambda) line edits.  Slot i holds logical pair index[i], or first_index + i*index_stride
 * when index is NULL (ranks take size-balanced shares of one logical pair set through index lists).
 * tsm_gen_pair_sizes -> tsm_layout (twice) -> tsm_gen_pair_fill. */
int tsm_gen_pair_sizes(uint64_t seed, int32_t n_pairs, const int32_t* index, int32_t first_index, int32_t index_stride,
                       int32_t cap, double lambda, int32_t* len_old, int32_t* len_new, uint8_t* ext);
int tsm_gen_pair_fill(uint64_t seed, int32_t n_pairs, const int32_t* index, int32_t first_index, int32_t index_stride,
                      int32_t cap, double lambda, const uint8_t* ext, const int32_t* off_old, const int32_t* len_old,
                      uint8_t* arena_old, const int32_t* off_new, const int32_t* len_new, uint8_t* arena_new);

#ifdef __cplusplus
}
#endif
#endif
