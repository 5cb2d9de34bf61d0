/* oracle/orc.c - CPU ORACLE, TEST INFRASTRUCTURE ONLY (see orc.h).  Plain C99, one thread,
 * no SIMD, no regex library.  Every function restates one section of docs/SPEC.md and cites the
 * reference artefact that pins it (paths relative to /root/reference). */
#define _GNU_SOURCE
#include "orc.h"
#include <stdlib.h>
#include <string.h>

static const char* const ORC_NAMES[] = {
#include "orc_categories.inc"
};
#define ORC_NAMED ((int)(sizeof(ORC_NAMES) / sizeof(ORC_NAMES[0])))

const char* orc_category_name(int id) {
  if (id >= 0 && id < ORC_NAMED) return ORC_NAMES[id];
  if (id == ORC_CAT_OTHER) return "<other>";
  return "";
}

/* ---------------------------------------------------------------- SPEC section 3: hashing (S9) */
#define M61 ((uint64_t)0x1FFFFFFFFFFFFFFFull)

static inline uint64_t rotl61(uint64_t x, unsigned r) { /* x < 2^61, r < 61 */
  if (r == 0) return x;
  return ((x << r) & M61) | (x >> (61 - r));
}
static inline uint64_t fold61(uint64_t x) { return (x & M61) + (x >> 61); }

static uint64_t mersenne61(const uint8_t* p, uint64_t len) {
  /* N = sum p[i]*256^i; 256^i mod (2^61-1) = 2^(8i mod 61): a rotation. */
  uint64_t acc = 0;
  unsigned r = 0;
  for (uint64_t i = 0; i < len; ++i) {
    acc = fold61(acc + rotl61((uint64_t)p[i], r));
    r += 8;
    if (r >= 61) r -= 61;
  }
  acc = fold61(acc);
  if (acc == M61) acc = 0;
  return acc;
}

static inline uint64_t finalise(uint64_t h61, uint64_t len) {
  uint64_t x = h61 ^ (len * 0x9E3779B97F4A7C15ull);
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

uint64_t orc_bytes_hash(const uint8_t* p, uint64_t len) { return finalise(mersenne61(p, len), len); }

uint64_t orc_line_hash(const uint8_t* line, uint64_t len) {
  if (len && line[len - 1] == 0x0D) --len; /* trailing CR excluded */
  return orc_bytes_hash(line, len);
}

/* SPEC section 3, n-gram hash: the window of up to n consecutive line hashes of one file that starts at line i,
 * G = sum over k of (line_hash[i+k] mod 2^61-1) * 2^(13k)  mod 2^61-1, finalised with the window length.
 * (S9 is a design choice of the north star; no artefact of the package attests it: SURVEY.md section 8a.) */
void orc_ngram_hashes(const uint64_t* line_hash, const int64_t* line_base, int32_t n_files, int32_t n, uint64_t* out) {
  for (int32_t f = 0; f < n_files; ++f)
    for (int64_t i = line_base[f]; i < line_base[f + 1]; ++i) {
      uint64_t acc = 0;
      unsigned r = 0;
      int32_t k = 0;
      for (; k < n && i + k < line_base[f + 1]; ++k) {
        uint64_t h = fold61(fold61(line_hash[i + k]));
        if (h == M61) h = 0;
        acc = fold61(acc + rotl61(h, r));
        r += 13;
        if (r >= 61) r -= 61;
      }
      acc = fold61(acc);
      if (acc == M61) acc = 0;
      out[i] = finalise(acc, (uint64_t)k);
    }
}

/* ---------------------------------------------------------------- helpers */
static inline int is_w(uint8_t c) { return c == 0x20 || c == 0x09 || c == 0x0D || c == 0x0B || c == 0x0C; }
static inline int is_ident(uint8_t c) {
  return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_';
}
static inline uint8_t lower(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c; }

static int contains_cs(const uint8_t* s, uint32_t n, const char* pat) {
  uint32_t m = (uint32_t)strlen(pat);
  if (m > n) return 0;
  return memmem(s, n, pat, m) != NULL;
}
static int contains_ci(const uint8_t* s, uint32_t n, const char* pat_lower) {
  uint32_t m = (uint32_t)strlen(pat_lower);
  if (m > n) return 0;
  for (uint32_t i = 0; i + m <= n; ++i) {
    uint32_t k = 0;
    while (k < m && lower(s[i + k]) == (uint8_t)pat_lower[k]) ++k;
    if (k == m) return 1;
  }
  return 0;
}
static int starts_with(const uint8_t* s, uint32_t n, const char* pat) {
  uint32_t m = (uint32_t)strlen(pat);
  return m <= n && memcmp(s, pat, m) == 0;
}
static void strip(const uint8_t* line, uint32_t len, uint32_t* b, uint32_t* e) {
  uint32_t s = 0, t = len;
  while (s < t && is_w(line[s])) ++s;
  while (t > s && is_w(line[t - 1])) --t;
  *b = s; *e = t;
}

/* ---------------------------------------------------------------- SPEC section 4 (S4)
 * Trigger: Important-files/ML-Testing-v1.xlsx!prefect_tests (4169/4169 rows contain `assert` ci or
 * `EXPECT_`), !MycroftAI_tests:R57 (docstring captured).  Truncation: 0 of 11 981 Rev-A statement
 * cells contain '('. */
int orc_is_assert_line(const uint8_t* line, uint32_t len) {
  return contains_ci(line, len, "assert") || contains_cs(line, len, "EXPECT_");
}

void orc_statement(const uint8_t* line, uint32_t len, uint32_t* stmt_off, uint32_t* stmt_len) {
  uint32_t b, e;
  strip(line, len, &b, &e);
  uint32_t t = b;
  while (t < e && line[t] != '(') ++t;
  while (t > b && is_w(line[t - 1])) --t;
  *stmt_off = b;
  *stmt_len = t - b;
}

/* ---------------------------------------------------------------- SPEC section 6 (S5)
 * Golden G4: statement -> category pairs of the five Rev-A sheets of ML-Testing-v1.xlsx
 * (apollo_tests, prefect_tests, carma-platform_tests, MycroftAI_tests, donkeycar_tests). */
enum { C_EMPTY = 0, C_EQ = 1, C_NE, C_TRUE, C_FALSE, C_GT, C_GE, C_LT, C_LE, C_NEAR, C_FLOAT_EQ, C_DOUBLE_EQ, C_RAISES };

static int stem_lookup(const uint8_t* s, uint32_t n) {
  static const struct { const char* stem; int id; } T[] = {
    {"EQ", C_EQ}, {"NE", C_NE}, {"TRUE", C_TRUE}, {"FALSE", C_FALSE}, {"GT", C_GT}, {"GE", C_GE},
    {"LT", C_LT}, {"LE", C_LE}, {"NEAR", C_NEAR}, {"FLOAT_EQ", C_FLOAT_EQ},
    {"DOUBLE_EQ", C_DOUBLE_EQ}, {"THROW", C_RAISES}};
  for (unsigned i = 0; i < sizeof(T) / sizeof(T[0]); ++i)
    if (strlen(T[i].stem) == n && memcmp(T[i].stem, s, n) == 0) return T[i].id;
  return C_EMPTY;
}

int orc_classify(const uint8_t* t, uint32_t len, uint32_t* ident_off, uint32_t* ident_len) {
  uint32_t i = len;
  while (i > 0 && is_ident(t[i - 1])) --i;
  const uint8_t* L = t + i;
  uint32_t ln = len - i;
  if (ident_off) *ident_off = i;
  if (ident_len) *ident_len = ln;
  /* rule 1: gtest macro (EXPECT_EQ->assertEqual 3435/3435, EXPECT_STREQ->'' 218/218) */
  if (starts_with(L, ln, "EXPECT_") || starts_with(L, ln, "ASSERT_")) return stem_lookup(L + 7, ln - 7);
  /* rule 2: bare assert (prefect_tests: 4235/4235 bare-assert rows) */
  if ((len == 6 && memcmp(t, "assert", 6) == 0) || starts_with(t, len, "assert ")) {
    const uint8_t* e = t + 7;
    uint32_t n = len > 7 ? len - 7 : 0;
    if (starts_with(e, n, "not ")) return C_NE;
    if ((contains_cs(e, n, " not ") && contains_cs(e, n, " in ")) || contains_cs(e, n, " is not ")) return C_FALSE;
    if (contains_cs(e, n, "True")) return C_TRUE;
    if (contains_cs(e, n, "==")) return C_EQ;
    if (contains_cs(e, n, "!=")) return C_NE;
    if (contains_cs(e, n, "<=")) return C_LE;
    if (contains_cs(e, n, ">=")) return C_GE;
    if (contains_cs(e, n, "<")) return C_LT;
    if (contains_cs(e, n, ">")) return C_GT;
    return C_TRUE;
  }
  /* rule 3 */
  if (ln == 7 && memcmp(L, "assert_", 7) == 0) return C_TRUE;
  /* rule 4: verbatim identifier (assertEquals is NOT folded into assertEqual: 113+54 rows) */
  if (starts_with(L, ln, "assert")) {
    for (int k = 1; k < ORC_NAMED; ++k)
      if (strlen(ORC_NAMES[k]) == ln && memcmp(ORC_NAMES[k], L, ln) == 0) return k;
    return ORC_CAT_OTHER;
  }
  return C_EMPTY; /* rule 5: `if`->'' x5, `GPUAssert`->'' x2, `FOR_EACH`->'' */
}

/* ---------------------------------------------------------------- SPEC section 5 (S3)
 * PY: ML-Testing-v1.xlsx!MycroftAI_tests:R13 (`default=` line is a header), :R4 (class header).
 * CJ: ML-Testing-v1.xlsx!apollo_tests:R8-R10 vs src/apollo/v6.0.0/modules/perception/fusion/
 *     common/dst_evidence_test.cc:42,:55,:71 (and :46, which contains "test" but no '{', is not). */
int orc_header_kind(int ext, const uint8_t* line, uint32_t len) {
  if (ext == 0) return 0;
  uint32_t b, e;
  strip(line, len, &b, &e);
  const uint8_t* s = line + b;
  uint32_t n = e - b;
  int hdr;
  if (ext == 1) {
    hdr = contains_cs(line, len, "def") ||
          (starts_with(s, n, "class") && n > 5 && (s[5] == 0x20 || s[5] == 0x09));
  } else {
    hdr = contains_ci(line, len, "test") &&
          (contains_cs(line, len, "{") || contains_cs(line, len, "class") || contains_cs(line, len, "void"));
  }
  if (!hdr) return 0;
  return starts_with(s, n, "TEST_F") ? 3 : 1;
}

uint32_t orc_method_string(int ext, const uint8_t* line, uint32_t len, uint8_t* out, uint32_t cap) {
  uint32_t b, e, o = 0;
  strip(line, len, &b, &e);
#define PUT(c) do { if (o < cap) out[o] = (c); ++o; } while (0)
  if (ext == 1) { /* ML-Testing-v1.xlsx!prefect_tests:R2, !donkeycar_tests:R3 */
    uint32_t i = b;
    if (starts_with(line + i, e - i, "class")) i += 5;
    while (i < e) {
      if (starts_with(line + i, e - i, "def")) { i += 3; continue; }
      if (!is_w(line[i])) PUT(line[i]);
      ++i;
    }
    if (o > 0 && o <= cap && out[o - 1] == ':') --o;
  } else if (ext == 4) { /* ML-Testing-v1.xlsx!carma-platform_tests (java rows) */
    static const char* const words[] = {"public", "private", "protected", "static", "void", "class"};
    uint32_t i = b;
    while (i < e) {
      int hit = 0;
      for (unsigned w = 0; w < 6; ++w)
        if (starts_with(line + i, e - i, words[w])) { i += (uint32_t)strlen(words[w]); hit = 1; break; }
      if (hit) continue;
      if (!is_w(line[i])) PUT(line[i]);
      ++i;
    }
  } else { /* ML-Testing-v1.xlsx!apollo_tests:R4, :R8 */
    uint32_t t = b;
    while (t < e && line[t] != ')') ++t;
    /* delete '{', then strip again: find bounds ignoring W and '{' at the ends */
    uint32_t s = b;
    while (s < t && (is_w(line[s]) || line[s] == '{')) ++s;
    uint32_t u = t;
    while (u > s && (is_w(line[u - 1]) || line[u - 1] == '{')) --u;
    for (uint32_t i = s; i < u; ++i)
      if (line[i] != '{') PUT(line[i]);
  }
#undef PUT
  return o < cap ? o : cap;
}

/* ---------------------------------------------------------------- SPEC section 4b: the later revision of the lost tool ("Rev B")
 * Golden G1: Important-files/ML-Testing-v1.xlsx!DeepSpeech vs src/DeepSpeech/v0.9.3 (the one version-matched count
 * golden): rows `BOOST_CHECK_EQUAL` x 1, 1, 2, 2 of native_client/kenlm/util/bit_packing_test.cc:18,26,35,38;
 * `BOOST_CHECK(!left.full);` (full statement) of lm/partial_test.cc; `assert (audioFormat == 1); // 1 is PCM` (java). */
int orc_is_assert_line_b(const uint8_t* line, uint32_t len) {
  return orc_is_assert_line(line, len) || contains_cs(line, len, "_CHECK") || contains_cs(line, len, "TESTEQUAL") ||
         contains_cs(line, len, "FAIL");
}

static int ident_is(const uint8_t* L, uint32_t ln, const char* name) { return strlen(name) == ln && memcmp(L, name, ln) == 0; }

/* Statement of Rev B: Rev A's T, except the whole stripped line for java files and for the callees BOOST_CHECK / NTA_CHECK. */
void orc_statement_b(int ext, const uint8_t* line, uint32_t len, uint32_t* stmt_off, uint32_t* stmt_len) {
  uint32_t so, sl;
  orc_statement(line, len, &so, &sl);
  uint32_t i = sl;
  while (i > 0 && is_ident(line[so + i - 1])) --i;
  const uint8_t* L = line + so + i;
  const uint32_t ln = sl - i;
  if (ext == 4 || ident_is(L, ln, "BOOST_CHECK") || ident_is(L, ln, "NTA_CHECK")) {
    uint32_t b, e;
    strip(line, len, &b, &e);
    so = b; sl = e - b;
  }
  *stmt_off = so; *stmt_len = sl;
}

/* Category of Rev B from Rev A's statement T (at line + so, length sl): rule 1b BOOST_CHECK_EQUAL -> assertEqual; rule 2b
 * the bare forms - T == "assert" followed by '(' (java `assert (x == 1);`), callee BOOST_CHECK / NTA_CHECK - by the
 * operators of the text behind the '(': leading '!' -> assertFalse, then the operator list of rule 2; nothing found:
 * assertTrue for assert, '' for the two macros.  Everything else: Rev A. */
int orc_classify_b(const uint8_t* line, uint32_t len, uint32_t so, uint32_t sl, uint32_t* ident_off, uint32_t* ident_len) {
  const uint8_t* t = line + so;
  uint32_t i = sl;
  while (i > 0 && is_ident(t[i - 1])) --i;
  const uint8_t* L = t + i;
  const uint32_t ln = sl - i;
  if (ident_is(L, ln, "BOOST_CHECK_EQUAL")) { if (ident_off) *ident_off = i; if (ident_len) *ident_len = ln; return C_EQ; }
  const int macro = ident_is(L, ln, "BOOST_CHECK") || ident_is(L, ln, "NTA_CHECK");
  const int bare = sl == 6 && memcmp(t, "assert", 6) == 0;
  uint32_t p = so + sl;                                   /* first byte behind T: blanks, then '(' ? */
  while (p < len && is_w(line[p])) ++p;
  if ((macro || bare) && p < len && line[p] == '(') {
    if (ident_off) *ident_off = i;
    if (ident_len) *ident_len = ln;
    uint32_t b, e;
    strip(line, len, &b, &e);
    const uint8_t* x = line + p + 1;
    uint32_t n = e > p + 1 ? e - (p + 1) : 0;
    while (n && is_w(*x)) { ++x; --n; }
    if (n && x[0] == '!' && !(n > 1 && x[1] == '=')) return C_FALSE;
    if ((contains_cs(x, n, " not ") && contains_cs(x, n, " in ")) || contains_cs(x, n, " is not ")) return C_FALSE;
    if (contains_cs(x, n, "True")) return C_TRUE;
    if (contains_cs(x, n, "==")) return C_EQ;
    if (contains_cs(x, n, "!=")) return C_NE;
    if (contains_cs(x, n, "<=")) return C_LE;
    if (contains_cs(x, n, ">=")) return C_GE;
    if (contains_cs(x, n, "<")) return C_LT;
    if (contains_cs(x, n, ">")) return C_GT;
    return macro ? C_EMPTY : C_TRUE;
  }
  return orc_classify(t, sl, ident_off, ident_len);
}

/* ---------------------------------------------------------------- full scan */
int orc_scan(const uint8_t* arena, const int32_t* off, const int32_t* len, const uint8_t* ext,
             const uint16_t* grp, int32_t n_files, int32_t n_groups,
             orc_file_stat* stats, int64_t* group_counts, int64_t* global_counts,
             orc_assert_event* aev, int64_t aev_cap, int64_t* n_aev,
             orc_header_event* hev, int64_t hev_cap, int64_t* n_hev,
             uint64_t* line_hash, int64_t* line_base) {
  return orc_scan_ex(arena, off, len, ext, grp, n_files, n_groups, stats, group_counts, global_counts, aev, aev_cap, n_aev,
                     hev, hev_cap, n_hev, line_hash, line_base, 0);
}

int orc_scan_ex(const uint8_t* arena, const int32_t* off, const int32_t* len, const uint8_t* ext,
             const uint16_t* grp, int32_t n_files, int32_t n_groups,
             orc_file_stat* stats, int64_t* group_counts, int64_t* global_counts,
             orc_assert_event* aev, int64_t aev_cap, int64_t* n_aev,
             orc_header_event* hev, int64_t hev_cap, int64_t* n_hev,
             uint64_t* line_hash, int64_t* line_base, uint32_t flags) {
  const int rev_b = (flags & ORC_REV_B) != 0;
  int64_t na = 0, nh = 0, nl = 0;
  if (group_counts) memset(group_counts, 0, sizeof(int64_t) * (size_t)n_groups * ORC_K);
  if (global_counts) memset(global_counts, 0, sizeof(int64_t) * ORC_K);
  for (int32_t f = 0; f < n_files; ++f) {
    if (off[f] < 0 || len[f] < 0 || (off[f] & 127)) return -1;
    if (grp && grp[f] >= n_groups) return -1;
    const uint8_t* p = arena + off[f];
    uint32_t size = (uint32_t)len[f];
    int x = ext ? ext[f] : 0;
    orc_file_stat st = {0, 0, 0, 0, 0};
    if (line_base) line_base[f] = nl;
    uint32_t pos = 0;
    while (pos < size) {
      const uint8_t* nlp = memchr(p + pos, '\n', size - pos);
      uint32_t end = nlp ? (uint32_t)(nlp - p) : size;
      const uint8_t* line = p + pos;
      uint32_t ll = end - pos;
      uint64_t h = orc_line_hash(line, ll);
      st.n_lines++;
      st.digest += h;
      if (line_hash) line_hash[nl] = h;
      ++nl;
      if (x != 0) {
        int hk = orc_header_kind(x, line, ll);
        if (hk) {
          st.n_headers++;
          if (hk & 2) st.n_fixture++;
          if (hev && nh < hev_cap) {
            orc_header_event ev = {(uint32_t)f, pos, ll, (uint32_t)(hk >> 1)};
            hev[nh] = ev;
          }
          ++nh;
        }
        if (rev_b ? orc_is_assert_line_b(line, ll) : orc_is_assert_line(line, ll)) {
          uint32_t so, sl, io, il;
          orc_statement(line, ll, &so, &sl);
          int cat = rev_b ? orc_classify_b(line, ll, so, sl, &io, &il) : orc_classify(line + so, sl, &io, &il);
          if (rev_b) {                                     /* the event carries the Rev-B statement; ident stays Rev A's L */
            const uint32_t so_a = so;
            orc_statement_b(x, line, ll, &so, &sl);
            io += so_a - so;
          }
          st.n_assert++;
          if (group_counts) group_counts[(size_t)(grp ? grp[f] : 0) * ORC_K + cat]++;
          if (global_counts) global_counts[cat]++;
          if (aev && na < aev_cap) {
            orc_assert_event ev;
            memset(&ev, 0, sizeof ev);
            ev.file = (uint32_t)f; ev.line_off = pos; ev.stmt_off = pos + so;
            ev.stmt_len = (uint16_t)(sl > 65535 ? 65535 : sl); ev.cat = (uint16_t)cat;
            ev.ident_off = pos + so + io; ev.ident_len = (uint16_t)(il > 65535 ? 65535 : il);
            ev.stmt_hash = orc_bytes_hash(line + so, sl);
            aev[na] = ev;
          }
          ++na;
        }
      }
      pos = end + 1; /* past the LF (or past the end) */
    }
    if (stats) stats[f] = st;
  }
  if (line_base) line_base[n_files] = nl;
  if (n_aev) *n_aev = na;
  if (n_hev) *n_hev = nh;
  return 0;
}

/* ---------------------------------------------------------------- SPEC section 8 (S8)
 * Important-files/ML-Testing-v1.xlsx!projects:R1 (cloc = added + removed). */
int64_t orc_lcs(const uint64_t* a, int64_t n, const uint64_t* b, int64_t m) {
  int64_t pre = 0;
  while (pre < n && pre < m && a[pre] == b[pre]) ++pre;
  int64_t suf = 0;
  while (suf < n - pre && suf < m - pre && a[n - 1 - suf] == b[m - 1 - suf]) ++suf;
  a += pre; b += pre; n -= pre + suf; m -= pre + suf;
  if (n == 0 || m == 0) return pre + suf;
  if (m > n) { const uint64_t* t = a; a = b; b = t; int64_t k = n; n = m; m = k; }
  int32_t* row = (int32_t*)calloc((size_t)m + 1, sizeof(int32_t));
  if (!row) return -1;
  for (int64_t i = 1; i <= n; ++i) {
    int32_t diag = 0; /* row[i-1][j-1] */
    for (int64_t j = 1; j <= m; ++j) {
      int32_t up = row[j];
      if (a[i - 1] == b[j - 1]) row[j] = diag + 1;
      else if (row[j - 1] > up) row[j] = row[j - 1];
      diag = up;
    }
  }
  int64_t r = row[m];
  free(row);
  return r + pre + suf;
}

static int64_t file_line_hashes(const uint8_t* p, uint32_t size, uint64_t** out) {
  int64_t cap = 64, n = 0;
  uint64_t* h = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)cap);
  uint32_t pos = 0;
  while (h && pos < size) {
    const uint8_t* nlp = memchr(p + pos, '\n', size - pos);
    uint32_t end = nlp ? (uint32_t)(nlp - p) : size;
    if (n == cap) { cap *= 2; h = (uint64_t*)realloc(h, sizeof(uint64_t) * (size_t)cap); if (!h) break; }
    h[n++] = orc_line_hash(p + pos, end - pos);
    pos = end + 1;
  }
  *out = h;
  return h ? n : -1;
}

int orc_diff_pairs(const uint8_t* arena_old, const int32_t* off_old, const int32_t* len_old,
                   const uint8_t* arena_new, const int32_t* off_new, const int32_t* len_new,
                   int32_t n_pairs, int64_t* added, int64_t* removed) {
  for (int32_t i = 0; i < n_pairs; ++i) {
    uint64_t *a = NULL, *b = NULL;
    int64_t n = file_line_hashes(arena_old + off_old[i], (uint32_t)len_old[i], &a);
    int64_t m = file_line_hashes(arena_new + off_new[i], (uint32_t)len_new[i], &b);
    if (n < 0 || m < 0) { free(a); free(b); return -1; }
    int64_t l = orc_lcs(a, n, b, m);
    free(a); free(b);
    if (l < 0) return -1;
    removed[i] = n - l;
    added[i] = m - l;
  }
  return 0;
}

/* SPEC section 8, hunks: Myers' greedy D-path search with the rows of V kept for the canonical backtrack.
 * BASELINE.json configs[4] ("per-hunk diff+classify"); no artefact pins it (parity unpinned). */
int64_t orc_diff_script(const uint64_t* a, int64_t n, const uint64_t* b, int64_t m, const uint8_t* fa,
                        const uint8_t* fb, orc_diff_detail* out) {
  orc_diff_detail d = {0, 0, 0, 0, 0};
  int64_t pre = 0;
  while (pre < n && pre < m && a[pre] == b[pre]) ++pre;
  int64_t suf = 0;
  while (suf < n - pre && suf < m - pre && a[n - 1 - suf] == b[m - 1 - suf]) ++suf;
  a += pre; b += pre; if (fa) fa += pre; if (fb) fb += pre;
  n -= pre + suf; m -= pre + suf;
  int64_t D = 0;
  if (n == 0 || m == 0) {
    D = n + m;
    if (n) { d.hunks_del = 1; if (fa) for (int64_t i = 0; i < n; ++i) d.removed_assert += fa[i] != 0; }
    if (m) { d.hunks_add = 1; if (fb) for (int64_t i = 0; i < m; ++i) d.added_assert += fb[i] != 0; }
    if (out) *out = d;
    return D;
  }
  const int64_t off = n + m + 1;
  int64_t* V = (int64_t*)calloc((size_t)(2 * (n + m) + 3), sizeof(int64_t));
  int64_t** rows = (int64_t**)calloc((size_t)(n + m + 1), sizeof(int64_t*));   /* rows[d][k + d] */
  if (!V || !rows) { free(V); free(rows); return -1; }
  int found = 0;
  V[off + 1] = 0;
  for (D = 0; D <= n + m && !found; ++D) {
    for (int64_t k = -D; k <= D; k += 2) {
      int64_t x = (k == -D || (k != D && V[off + k - 1] < V[off + k + 1])) ? V[off + k + 1] : V[off + k - 1] + 1;
      int64_t y = x - k;
      while (x < n && y < m && a[x] == b[y]) { ++x; ++y; }
      V[off + k] = x;
      if (x >= n && y >= m) found = 1;
    }
    rows[D] = (int64_t*)malloc(sizeof(int64_t) * (size_t)(2 * D + 1));
    if (!rows[D]) { found = -1; break; }
    memcpy(rows[D], V + off - D, sizeof(int64_t) * (size_t)(2 * D + 1));
  }
  if (found == 1) {
    --D;
    /* backtrack: edits from the last to the first; snake_after = matches between this edit and the next */
    int64_t x = n, y = m;
    int in_hunk = 0, has_add = 0, has_del = 0;
    for (int64_t dd = D; dd >= 1; --dd) {
      const int64_t k = x - y;
      const int64_t* P = rows[dd - 1];                    /* P[kk + dd - 1] */
      const int down = (k == -dd || (k != dd && P[k - 1 + dd - 1] < P[k + 1 + dd - 1]));
      const int64_t pk = down ? k + 1 : k - 1;
      const int64_t px = P[pk + dd - 1], py = px - pk;
      const int64_t midx = down ? px : px + 1;
      const int64_t snake_after = x - midx;
      if (in_hunk && snake_after > 0) {                   /* a match separates this edit from the later hunk */
        if (has_add && has_del) d.hunks_mod++; else if (has_add) d.hunks_add++; else d.hunks_del++;
        has_add = has_del = 0;
      }
      in_hunk = 1;
      if (down) { has_add = 1; if (fb) d.added_assert += fb[py] != 0; }
      else { has_del = 1; if (fa) d.removed_assert += fa[px] != 0; }
      x = px; y = py;
    }
    if (in_hunk) { if (has_add && has_del) d.hunks_mod++; else if (has_add) d.hunks_add++; else d.hunks_del++; }
  }
  for (int64_t i = 0; i <= n + m; ++i) free(rows[i]);
  free(rows); free(V);
  if (found != 1) return -1;
  if (out) *out = d;
  return D;
}

static int64_t file_line_hashes_flags(const uint8_t* p, uint32_t size, int ext, uint64_t** out, uint8_t** flags) {
  int64_t n = file_line_hashes(p, size, out);
  if (n < 0) return n;
  uint8_t* f = (uint8_t*)calloc((size_t)(n ? n : 1), 1);
  if (!f) return -1;
  uint32_t pos = 0; int64_t i = 0;
  while (pos < size) {
    const uint8_t* nlp = memchr(p + pos, '\n', size - pos);
    uint32_t end = nlp ? (uint32_t)(nlp - p) : size;
    f[i++] = (uint8_t)(ext != 0 && orc_is_assert_line(p + pos, end - pos));
    pos = end + 1;
  }
  *flags = f;
  return n;
}

int orc_diff_pairs_detail(const uint8_t* arena_old, const int32_t* off_old, const int32_t* len_old, const uint8_t* ext_old,
                          const uint8_t* arena_new, const int32_t* off_new, const int32_t* len_new, const uint8_t* ext_new,
                          int32_t n_pairs, int64_t* added, int64_t* removed, orc_diff_detail* detail) {
  for (int32_t i = 0; i < n_pairs; ++i) {
    uint64_t *a = NULL, *b = NULL; uint8_t *fa = NULL, *fb = NULL;
    int64_t n = file_line_hashes_flags(arena_old + off_old[i], (uint32_t)len_old[i], ext_old ? ext_old[i] : 0, &a, &fa);
    int64_t m = file_line_hashes_flags(arena_new + off_new[i], (uint32_t)len_new[i], ext_new ? ext_new[i] : 0, &b, &fb);
    int64_t D = (n < 0 || m < 0) ? -1 : orc_diff_script(a, n, b, m, fa, fb, &detail[i]);
    free(a); free(b); free(fa); free(fb);
    if (D < 0) return -1;
    const int64_t l = (n + m - D) / 2;
    removed[i] = n - l;
    added[i] = m - l;
  }
  return 0;
}

/* ---------------------------------------------------------------- SPEC section 10 (body statements, golden G2)
 * Important-files/ML-Analysis-v4.xlsx!Apollo:R2-R26 = src/apollo/v6.0.0/modules/common/math/aabox2d_test.cc:27-53. */
int64_t orc_statements(const uint8_t* arena, const int32_t* off, const int32_t* len, int32_t n_files,
                       int64_t* line_base, uint32_t* line_end, uint8_t* line_kind, int64_t cap) {
  int64_t nl = 0;
  for (int32_t f = 0; f < n_files; ++f) {
    const uint8_t* p = arena + off[f];
    const uint32_t size = (uint32_t)len[f];
    if (line_base) line_base[f] = nl;
    int64_t depth = 0;
    uint32_t pos = 0;
    while (pos < size) {
      const uint8_t* nlp = memchr(p + pos, '\n', size - pos);
      const uint32_t end = nlp ? (uint32_t)(nlp - p) : size;
      uint32_t b, e;
      strip(p + pos, end - pos, &b, &e);
      uint8_t kind = 0;
      if (e > b) {
        kind = depth == 0 ? 1 : 2;
        for (uint32_t i = pos; i < end; ++i) depth += (p[i] == '(') - (p[i] == ')');
        if (depth < 0) depth = 0;
      }
      if (nl < cap) { if (line_end) line_end[nl] = end; if (line_kind) line_kind[nl] = kind; }
      ++nl;
      pos = end + 1;
    }
  }
  if (line_base) line_base[n_files] = nl;
  return nl;
}

/* ---------------------------------------------------------------- SPEC section 9 (S10)
 * RQs/taxonomy_test2.csv -> RQs/RQ3/tests_strategy_rq32.csv, RQs/RQ4/tests_methods_v2.csv. */
int orc_reduce(const uint8_t* flags, const int32_t* repo, const int32_t* case_id, int32_t n_rows,
               int32_t n_flags, int32_t n_repos, int32_t n_cases, int64_t* out,
               int64_t* cases_per_repo) {
  size_t words = ((size_t)n_cases + 63) / 64;
  uint64_t* seen = (uint64_t*)calloc(words ? words : 1, sizeof(uint64_t));
  if (!seen) return -1;
  for (int32_t f = -1; f < n_flags; ++f) {
    for (int32_t r = 0; r < n_repos; ++r) {
      memset(seen, 0, words * sizeof(uint64_t));
      int64_t cnt = 0;
      for (int32_t i = 0; i < n_rows; ++i) {
        if (repo[i] != r) continue;
        if (f >= 0 && flags[(size_t)i * n_flags + f] == 0) continue;
        int32_t c = case_id[i];
        if (c < 0 || c >= n_cases) { free(seen); return -1; }
        uint64_t bit = 1ull << (c & 63);
        if (!(seen[c >> 6] & bit)) { seen[c >> 6] |= bit; ++cnt; }
      }
      if (f < 0) { if (cases_per_repo) cases_per_repo[r] = cnt; }
      else out[(size_t)f * n_repos + r] = cnt;
    }
  }
  free(seen);
  return 0;
}
