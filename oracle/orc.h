/* oracle/orc.h - CPU ORACLE of the corpus-scan hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain single-thread C restatement of docs/SPEC.md.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the product path
 * (tosem-2021-replication_b200/) never links, imports or calls it.
 *
 * There is no reference scanner to restate (the package ships the loop's inputs
 * and outputs, not its code), so each function cites the reference ARTEFACT that pins its rule.
 * Pinned against golden vectors: S4 truncation + S5 categories (G4, 11 954 / 11 981 rows),
 * S3 header rule (Apollo ledger), S10 reduce (G3).  PARITY UNPINNED (nothing in the package
 * fixes the result): S9 hashing, S8 churn, S2 mock/Module modifiers.
 */
#ifndef ORC_H
#define ORC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_K 128
#define ORC_CAT_OTHER 127

typedef struct { uint32_t n_lines, n_assert, n_headers, n_fixture; uint64_t digest; } orc_file_stat;
typedef struct {
  uint32_t file, line_off, stmt_off; uint16_t stmt_len, cat;
  uint32_t ident_off; uint16_t ident_len, pad; uint64_t stmt_hash;
} orc_assert_event;
typedef struct { uint32_t file, line_off, line_len, kind; } orc_header_event;

/* SPEC section 3 */
uint64_t orc_bytes_hash(const uint8_t* p, uint64_t len);
uint64_t orc_line_hash(const uint8_t* line, uint64_t len); /* drops one trailing CR */

/* SPEC section 3: n-gram hashes over the line hashes of every file (window of up to n lines starting at every line). */
void orc_ngram_hashes(const uint64_t* line_hash, const int64_t* line_base, int32_t n_files, int32_t n, uint64_t* out);

/* SPEC section 6: category id of statement T; ident_off/ident_len locate L inside T. */
int orc_classify(const uint8_t* t, uint32_t len, uint32_t* ident_off, uint32_t* ident_len);
const char* orc_category_name(int id);

/* SPEC section 4: statement of a line -> offset/length inside the line. */
void orc_statement(const uint8_t* line, uint32_t len, uint32_t* stmt_off, uint32_t* stmt_len);
int orc_is_assert_line(const uint8_t* line, uint32_t len);
/* SPEC section 5: returns 0 (not a header), 1 (header) or 3 (fixture header). */
int orc_header_kind(int ext, const uint8_t* line, uint32_t len);
/* SPEC section 5 method string; returns its length (truncated to cap). */
uint32_t orc_method_string(int ext, const uint8_t* line, uint32_t len, uint8_t* out, uint32_t cap);

/* Full scan (SPEC sections 2-7).  Event arrays may be NULL (then only counted); *n_* return the
 * number of events that exist (may exceed cap; only cap are written).  line_hash may be NULL; if
 * not, it receives every line hash, files in order, and line_base[n_files+1] the per-file starts.
 * Returns 0, or -1 on a malformed corpus. */
int orc_scan(const uint8_t* arena, const int32_t* off, const int32_t* len, const uint8_t* ext,
             const uint16_t* grp, int32_t n_files, int32_t n_groups,
             orc_file_stat* stats, int64_t* group_counts, int64_t* global_counts,
             orc_assert_event* aev, int64_t aev_cap, int64_t* n_aev,
             orc_header_event* hev, int64_t hev_cap, int64_t* n_hev,
             uint64_t* line_hash, int64_t* line_base);

/* SPEC section 8: LCS length of two hash sequences (exact, O(n*m) DP with O(min) memory). */
int64_t orc_lcs(const uint64_t* a, int64_t n, const uint64_t* b, int64_t m);
int orc_diff_pairs(const uint8_t* arena_old, const int32_t* off_old, const int32_t* len_old,
                   const uint8_t* arena_new, const int32_t* off_new, const int32_t* len_new,
                   int32_t n_pairs, int64_t* added, int64_t* removed);

/* SPEC section 8, hunks: canonical edit script of two hash sequences; fa/fb = per-line assertion flags (may be
 * NULL).  Returns the edit distance D (insertions + deletions) or -1. */
typedef struct { int64_t hunks_add, hunks_del, hunks_mod, added_assert, removed_assert; } orc_diff_detail;
int64_t orc_diff_script(const uint64_t* a, int64_t n, const uint64_t* b, int64_t m, const uint8_t* fa,
                        const uint8_t* fb, orc_diff_detail* out);
int orc_diff_pairs_detail(const uint8_t* arena_old, const int32_t* off_old, const int32_t* len_old, const uint8_t* ext_old,
                          const uint8_t* arena_new, const int32_t* off_new, const int32_t* len_new, const uint8_t* ext_new,
                          int32_t n_pairs, int64_t* added, int64_t* removed, orc_diff_detail* detail);

/* SPEC section 4b (Rev B, golden G1): trigger, statement and category of the later revision of the lost tool; orc_scan_ex
 * with ORC_REV_B runs the full scan with them (orc_scan == orc_scan_ex with flags 0). */
#define ORC_REV_B 8u
int orc_is_assert_line_b(const uint8_t* line, uint32_t len);
void orc_statement_b(int ext, const uint8_t* line, uint32_t len, uint32_t* stmt_off, uint32_t* stmt_len);
int orc_classify_b(const uint8_t* line, uint32_t len, uint32_t so, uint32_t sl, uint32_t* ident_off, uint32_t* ident_len);
int orc_scan_ex(const uint8_t* arena, const int32_t* off, const int32_t* len, const uint8_t* ext,
                const uint16_t* grp, int32_t n_files, int32_t n_groups,
                orc_file_stat* stats, int64_t* group_counts, int64_t* global_counts,
                orc_assert_event* aev, int64_t aev_cap, int64_t* n_aev,
                orc_header_event* hev, int64_t hev_cap, int64_t* n_hev,
                uint64_t* line_hash, int64_t* line_base, uint32_t flags);

/* SPEC section 10: per-line statement kinds (0 blank, 1 first line of a statement, 2 continuation).  Fills
 * line_base[n_files+1]; line_end / line_kind hold up to cap lines (file-relative end offset of every line =
 * position of its LF or the file size).  Returns the total number of lines, or -1. */
int64_t orc_statements(const uint8_t* arena, const int32_t* off, const int32_t* len, int32_t n_files,
                       int64_t* line_base, uint32_t* line_end, uint8_t* line_kind, int64_t cap);

/* SPEC section 9: out[f*n_repos + r] = distinct cases with flag f set in repo r;
 * cases_per_repo[r] = distinct cases of repo r.  case ids < n_cases. */
int orc_reduce(const uint8_t* flags, const int32_t* repo, const int32_t* case_id, int32_t n_rows,
               int32_t n_flags, int32_t n_repos, int32_t n_cases, int64_t* out,
               int64_t* cases_per_repo);

/* orc_mt.c: orc_scan over a pool of persistent POSIX threads, files partitioned statically by bytes (the
 * host-cores baseline bench.py times).  n_threads <= 0: one per CPU of the calling thread's affinity mask.
 * orc_mt_create returns the thread count; orc_mt_scan returns 0 / -1 (corpus) / -2 (no pool). */
int orc_mt_affinity_cpus(void);
int orc_mt_create(int n_threads, int32_t max_groups);
void orc_mt_destroy(void);
int orc_mt_scan(const uint8_t* arena, const int32_t* off, const int32_t* len, const uint8_t* ext, const uint16_t* grp,
                int32_t n_files, int32_t n_groups, orc_file_stat* stats, int64_t* group_counts, int64_t* global_counts);
int orc_mt_diff(const uint8_t* arena_old, const int32_t* off_old, const int32_t* len_old, const uint8_t* ext_old,
                const uint8_t* arena_new, const int32_t* off_new, const int32_t* len_new, const uint8_t* ext_new,
                int32_t n_pairs, int64_t* added, int64_t* removed, orc_diff_detail* detail);

#ifdef __cplusplus
}
#endif
#endif
