// tsm_gen.cpp - host-only helpers of libtosemscan.so: arena layout and the deterministic synthetic
// corpus of SURVEY.md section 8d (configs C2-C5 of BASELINE.json).  std::mt19937_64 raw outputs only
// (the standard fixes the engine, not the distributions), one engine per file seeded from
// (seed, file index) so that any rank can generate any subset of one logical corpus.
//
// Content model, parameters measured on the 1 779 bundled test files:
//   line length buckets 23 % <10 B, 7 % 10-19, 14 % 20-29, 14 % 30-39, 12 % 40-49, 10 % 50-59,
//   9 % 60-69, 8 % 70-79, 2 % >= 80;  8.8 % of lines carry an assertion token drawn with the
//   corpus frequencies;  one test-case header every ~48 lines;  ext 50 % py / 40 % cc / 10 % java.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/tosemscan.h"

namespace {

struct Rng {
  std::mt19937_64 e;
  explicit Rng(uint64_t s) : e(s) {}
  uint64_t next() { return e(); }
  uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }   // n < 2^32
  double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }            // [0,1)
};

uint64_t file_seed(uint64_t seed, uint64_t index) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (index + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

const char* const kWords[] = {
    "self", "result", "value", "data", "config", "model", "input", "output", "index", "count", "name", "path",
    "node", "state", "buffer", "size", "len", "range", "return", "if", "else", "for", "in", "not", "None", "True",
    "False", "import", "from", "with", "as", "try", "except", "raise", "lambda", "int", "float", "str", "list",
    "dict", "std", "vector", "string", "const", "auto", "double", "bool", "nullptr", "new", "this", "static",
    "public", "private", "final", "throws", "Exception", "tensor", "shape", "batch", "layer", "weights", "loss",
    "optimizer", "session", "graph", "worker", "actor", "task", "queue", "message", "request", "response",
    "latest", "context", "manager", "process", "address", "passed", "classes", "default", "definition", "voided",
    "tester", "contest", "x", "y", "i", "j", "k", "n", "a", "b", "tmp", "ret", "obj", "args", "kwargs", "np",
    "tf", "os", "sys", "math", "random", "time", "json", "logging", "0", "1", "2", "10", "0.5", "1e-6", "42"};
constexpr int kNumWords = sizeof(kWords) / sizeof(kWords[0]);
const char* const kPunct[] = {" ", " ", " ", " = ", ".", "(", ")", ", ", ": ", "[", "]", " == ", " + ", "_", "->", "::", " < ", "; "};
constexpr int kNumPunct = sizeof(kPunct) / sizeof(kPunct[0]);

struct Tok { const char* text; int weight; };
// corpus frequencies: EXPECT_EQ 4375 : EXPECT_TRUE 1544 : EXPECT_NEAR 1412 : ...
const Tok kAssertCc[] = {{"EXPECT_EQ(", 4375}, {"EXPECT_TRUE(", 1544}, {"EXPECT_NEAR(", 1412}, {"EXPECT_FALSE(", 900},
                         {"EXPECT_DOUBLE_EQ(", 938}, {"EXPECT_FLOAT_EQ(", 393}, {"ASSERT_TRUE(", 700}, {"ASSERT_EQ(", 650},
                         {"EXPECT_NE(", 420}, {"EXPECT_GT(", 150}, {"EXPECT_LE(", 88}, {"EXPECT_STREQ(", 218},
                         {"EXPECT_THROW(", 120}, {"EXPECT_CALL(", 144}, {"ASSERT_NE(", 100}, {"else ASSERT_EQ(", 5},
                         {"// EXPECT_EQ(", 20}, {"RAPIDJSON_ASSERT(", 4}, {"BOOST_CHECK_EQUAL(", 30}};
const Tok kAssertPy[] = {{"self.assertEqual(", 3300}, {"self.assertTrue(", 1400}, {"self.assertAlmostEqual(", 380},
                         {"self.assertFalse(", 500}, {"self.assertIn(", 205}, {"self.assertIsInstance(", 205},
                         {"self.assertRaises(", 283}, {"self.assertListEqual(", 86}, {"self.assertAllClose(", 67},
                         {"self.assert_(", 10}, {"mock.assert_called_once_with(", 36}, {"np.testing.assert_array_equal(", 27},
                         {"assert ", 3800}, {"assert not ", 300}, {"assert(", 40}, {"except AssertionError:", 10},
                         {"\"\"\"Assert that the ", 8}, {"self.assertWeirdCustomThing(", 12}, {"assert_raises(", 52}};
const Tok kAssertJava[] = {{"assertEquals(", 656}, {"assertTrue(", 400}, {"assertFalse(", 150}, {"assertNull(", 18},
                           {"assertNotNull(", 13}, {"assertArrayEquals(", 7}, {"Assert.assertEquals(", 60}, {"assert (", 30},
                           {"assertThat(", 25}};
const char* const kBareTail[] = {"== ", "!= ", "<= ", ">= ", "< ", "> ", "is not ", "not in ", "in ", "is ", "== True and ", ""};

template <int N>
const char* pick(Rng& r, const Tok (&t)[N]) {
  int total = 0;
  for (int i = 0; i < N; ++i) total += t[i].weight;
  int x = (int)r.below((uint32_t)total);
  for (int i = 0; i < N; ++i) { x -= t[i].weight; if (x < 0) return t[i].text; }
  return t[N - 1].text;
}

int line_target(Rng& r) {                       // bucketed length law
  const uint32_t u = r.below(100);
  if (u < 23) return (int)r.below(10);
  if (u < 30) return 10 + (int)r.below(10);
  if (u < 44) return 20 + (int)r.below(10);
  if (u < 58) return 30 + (int)r.below(10);
  if (u < 70) return 40 + (int)r.below(10);
  if (u < 80) return 50 + (int)r.below(10);
  if (u < 89) return 60 + (int)r.below(10);
  if (u < 97) return 70 + (int)r.below(10);
  return 80 + (int)r.below(60);
}

void filler(Rng& r, std::string& s, size_t target) {
  while (s.size() < target) {
    s += kWords[r.below(kNumWords)];
    if (s.size() < target) s += kPunct[r.below(kNumPunct)];
  }
}

// One line (without LF) of a file with the given ext tag.
void make_line(Rng& r, int ext, std::string& s) {
  s.clear();
  const int target = line_target(r);
  const uint32_t kind = r.below(1000);
  if (target >= 4) s.append(2 * r.below(5), ' ');
  if (kind < 88 && target >= 8) {               // 8.8 % assertion lines
    if (ext == TSM_EXT_PY) {
      const char* t = pick(r, kAssertPy);
      s += t;
      if (!strcmp(t, "assert ") || !strcmp(t, "assert not ")) {
        s += kWords[r.below(kNumWords)];
        s += ' ';
        s += kBareTail[r.below(sizeof(kBareTail) / sizeof(kBareTail[0]))];
      }
    } else if (ext == TSM_EXT_JAVA) s += pick(r, kAssertJava);
    else s += pick(r, kAssertCc);
    filler(r, s, (size_t)target);
    if (s.find('(') != std::string::npos) s += (ext == TSM_EXT_PY ? ")" : ");");
  } else if (kind < 88 + 21 && target >= 8) {   // a header every ~48 lines
    const char* w = kWords[r.below(kNumWords)];
    if (ext == TSM_EXT_PY) {
      if (r.below(8) == 0) { s += "class Test"; s += w; s += "(unittest.TestCase):"; }
      else { s += "def test_"; s += w; s += "(self):"; }
    } else if (ext == TSM_EXT_JAVA) { s += "public void test"; s += w; s += "() throws Exception {"; }
    else {
      s += (r.below(3) == 0 ? "TEST_F(" : "TEST(");
      s += w; s += "Test, "; s += kWords[r.below(kNumWords)]; s += ") {";
    }
  } else {
    filler(r, s, (size_t)target);
    if (kind >= 990) s += '\r';                 // a few CRLF lines
  }
}

// Generates file `index`; appends to `out` if non-null; returns the byte count.
int64_t gen_file(uint64_t seed, int64_t index, int size_law, int32_t fixed_size, int ext, std::string* out) {
  Rng r(file_seed(seed, (uint64_t)index));
  int64_t target;
  if (size_law == 0) target = fixed_size;
  else {                                         // pdf ~ x^-1.5 on [128, 2^20]: inverse CDF
    const double a = 1.0 / std::sqrt(128.0), b = 1.0 / std::sqrt(1048576.0);
    const double t = a - r.unit() * (a - b);
    target = (int64_t)(1.0 / (t * t));
    if (target < 128) target = 128;
    if (target > 1048576) target = 1048576;
  }
  int64_t n = 0;
  std::string line;
  while (n < target) {
    make_line(r, ext, line);
    line += '\n';
    if (size_law == 0 && n + (int64_t)line.size() > target) {   // exact size: cut the last line
      line.resize((size_t)(target - n));
      line[line.size() - 1] = '\n';
    }
    if (out) out->append(line);
    n += (int64_t)line.size();
  }
  return n;
}

int gen_ext(uint64_t seed, int64_t index) {
  const uint32_t u = (uint32_t)(file_seed(seed ^ 0xE17ull, (uint64_t)index) % 10);
  return u < 5 ? TSM_EXT_PY : (u < 9 ? TSM_EXT_CC : TSM_EXT_JAVA);
}

}  // namespace

extern "C" int64_t tsm_layout(const int32_t* len, int32_t n_files, int32_t* off) {
  if (n_files < 0 || (n_files && (!len || !off))) return -1;
  int64_t o = 0;
  for (int32_t i = 0; i < n_files; ++i) {
    if (len[i] < 0) return -1;
    off[i] = (int32_t)o;
    o += ((int64_t)len[i] + TSM_ALIGN - 1) / TSM_ALIGN * TSM_ALIGN;
    if (o >= (1ll << 31)) return -1;
  }
  if (off) off[n_files] = (int32_t)o;
  return o;
}

extern "C" int tsm_gen_sizes(uint64_t seed, int32_t n_files, int size_law, int32_t fixed_size,
                             int32_t first_index, int32_t index_stride, int32_t* len, uint8_t* ext,
                             uint16_t* grp, int32_t n_groups) {
  if (n_files < 0 || !len || !ext || n_groups < 1 || (size_law == 0 && fixed_size < 1) || index_stride < 1 ||
      first_index < 0 || size_law < 0 || size_law > 1)
    return TSM_E_ARG;
  for (int32_t i = 0; i < n_files; ++i) {
    const int64_t logical = (int64_t)first_index + (int64_t)i * index_stride;
    ext[i] = (uint8_t)gen_ext(seed, logical);
    len[i] = size_law == 0 ? fixed_size : (int32_t)gen_file(seed, logical, size_law, fixed_size, ext[i], nullptr);
    if (grp) grp[i] = (uint16_t)(file_seed(seed ^ 0x6A0ull, (uint64_t)logical) % (uint64_t)n_groups);
  }
  return TSM_OK;
}

extern "C" int tsm_gen_fill(uint64_t seed, int32_t n_files, int size_law, int32_t first_index,
                            int32_t index_stride, const int32_t* off, const int32_t* len,
                            const uint8_t* ext, uint8_t* arena) {
  if (n_files < 0 || !off || !len || !ext || !arena || index_stride < 1 || first_index < 0 || size_law < 0 ||
      size_law > 1)
    return TSM_E_ARG;
  std::string buf;
  for (int32_t i = 0; i < n_files; ++i) {
    const int64_t logical = (int64_t)first_index + (int64_t)i * index_stride;
    buf.clear();
    gen_file(seed, logical, size_law, len[i], ext[i], &buf);
    if ((int64_t)buf.size() != len[i]) return TSM_E_ARG;   // len[] must come from tsm_gen_sizes
    memcpy(arena + off[i], buf.data(), buf.size());
  }
  return TSM_OK;
}

namespace {
int64_t edit_lines(uint64_t seed, const uint8_t* src, int32_t src_len, double lambda, std::string* out);
}

extern "C" int64_t tsm_gen_edit(uint64_t seed, const uint8_t* src, int32_t src_len, double lambda,
                                uint8_t* dst, int64_t cap) {
  if (!src || src_len < 0 || !dst || cap < 0) return -1;
  std::string out;
  const int64_t n = edit_lines(seed, src, src_len, lambda, &out);
  if (n > cap) return -1;
  memcpy(dst, out.data(), (size_t)n);
  return n;
}

namespace {
int64_t edit_lines(uint64_t seed, const uint8_t* src, int32_t src_len, double lambda, std::string* out) {
  Rng r(file_seed(seed, 0xD1FFull));
  std::vector<std::string> lines;
  for (int32_t p = 0; p < src_len;) {
    const void* nl = memchr(src + p, '\n', (size_t)(src_len - p));
    const int32_t e = nl ? (int32_t)((const uint8_t*)nl - src) + 1 : src_len;
    lines.emplace_back((const char*)src + p, (size_t)(e - p));
    p = e;
  }
  // Poisson(lambda) edits (Knuth), each a Geometric(0.4) run of lines at a uniform position
  int edits = 0;
  for (double L = std::exp(-lambda), pr = r.unit(); pr > L; pr *= r.unit()) ++edits;
  std::string line;
  for (int k = 0; k < edits; ++k) {
    int run = 1;
    while (r.unit() >= 0.4 && run < 64) ++run;
    const uint32_t op = r.below(3);
    const size_t at = lines.empty() ? 0 : r.below((uint32_t)lines.size() + 1);
    if (op == 0 || op == 2) {                    // delete / replace: remove up to `run` lines at `at`
      const size_t a = at < lines.size() ? at : lines.size();
      const size_t b = a + (size_t)run < lines.size() ? a + (size_t)run : lines.size();
      lines.erase(lines.begin() + (long)a, lines.begin() + (long)b);
    }
    if (op == 1 || op == 2) {                    // insert / replace: add `run` fresh lines at `at`
      const size_t a = at < lines.size() ? at : lines.size();
      for (int j = 0; j < run; ++j) {
        make_line(r, TSM_EXT_PY + (int)r.below(2), line);
        line += '\n';
        lines.insert(lines.begin() + (long)a, line);
      }
    }
  }
  int64_t n = 0;
  for (const std::string& s : lines) {
    if (out) out->append(s);
    n += (int64_t)s.size();
  }
  return n;
}

// BASELINE config C5: pair `index` = (old, new); old follows the C4 size law with the target
// clamped to `cap` bytes (whole lines), new = old with Poisson(lambda) line edits.
void gen_pair(uint64_t seed, int64_t index, int32_t cap, double lambda, int ext, std::string& old_s, std::string& new_s) {
  Rng r(file_seed(seed, (uint64_t)index));
  const double a = 1.0 / std::sqrt(128.0), b = 1.0 / std::sqrt(1048576.0);
  const double t = a - r.unit() * (a - b);
  int64_t target = (int64_t)(1.0 / (t * t));
  if (target < 128) target = 128;
  if (target > cap) target = cap;
  old_s.clear();
  std::string line;
  while ((int64_t)old_s.size() < target) { make_line(r, ext, line); line += '\n'; old_s += line; }
  new_s.clear();
  edit_lines(seed ^ (0xC5ull << 56) ^ (uint64_t)index, (const uint8_t*)old_s.data(), (int32_t)old_s.size(), lambda, &new_s);
}
}  // namespace

extern "C" int tsm_gen_pair_sizes(uint64_t seed, int32_t n_pairs, const int32_t* index, int32_t first_index, int32_t index_stride,
                                  int32_t cap, double lambda, int32_t* len_old, int32_t* len_new, uint8_t* ext) {
  if (n_pairs < 0 || !len_old || !len_new || !ext || first_index < 0 || index_stride < 1 || cap < 128 || lambda < 0) return TSM_E_ARG;
  std::string o, n;
  for (int32_t i = 0; i < n_pairs; ++i) {
    const int64_t logical = index ? (int64_t)index[i] : (int64_t)first_index + (int64_t)i * index_stride;
    if (logical < 0) return TSM_E_ARG;
    ext[i] = (uint8_t)gen_ext(seed, logical);
    gen_pair(seed, logical, cap, lambda, ext[i], o, n);
    len_old[i] = (int32_t)o.size();
    len_new[i] = (int32_t)n.size();
  }
  return TSM_OK;
}

extern "C" int tsm_gen_pair_fill(uint64_t seed, int32_t n_pairs, const int32_t* index, int32_t first_index, int32_t index_stride,
                                 int32_t cap, double lambda, const uint8_t* ext, const int32_t* off_old, const int32_t* len_old,
                                 uint8_t* arena_old, const int32_t* off_new, const int32_t* len_new, uint8_t* arena_new) {
  if (n_pairs < 0 || !ext || !off_old || !len_old || !arena_old || !off_new || !len_new || !arena_new || first_index < 0 ||
      index_stride < 1 || cap < 128)
    return TSM_E_ARG;
  std::string o, n;
  for (int32_t i = 0; i < n_pairs; ++i) {
    const int64_t logical = index ? (int64_t)index[i] : (int64_t)first_index + (int64_t)i * index_stride;
    gen_pair(seed, logical, cap, lambda, ext[i], o, n);
    if ((int64_t)o.size() != len_old[i] || (int64_t)n.size() != len_new[i]) return TSM_E_ARG;   // sizes must come from tsm_gen_pair_sizes
    memcpy(arena_old + off_old[i], o.data(), o.size());
    memcpy(arena_new + off_new[i], n.data(), n.size());
  }
  return TSM_OK;
}
