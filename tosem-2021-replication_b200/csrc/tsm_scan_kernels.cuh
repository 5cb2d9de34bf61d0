// tsm_scan_kernels.cuh - hand-written sm_100a kernels of the corpus scan (docs/SPEC.md).
//
//   k_plan      files -> (file, 4 KiB chunk) work units                         [tiny]
//   k_scan      THE hot kernel: every source byte is read from HBM exactly once.  One warp per
//               work unit, chunk staged global->shared by a 1-D TMA bulk copy (cp.async.bulk +
//               mbarrier), pass 1 = SWAR newline table, pass 2 = lane-per-stripe Shift-And
//               automaton + Mersenne-61 running prefix, pass 3 = lane-per-line finalise, then
//               per-file counters, digest and the candidate (assertion-line) list.
//   k_classify  one thread per candidate: statement, last identifier, category (S5), events, the
//               cross-file aggregate into a shared-memory privatised [group][category] table, and
//               the totals of the per-file records.
//
// There is no reference kernel: the reference ships data only.  Rules cite
// docs/SPEC.md, which cites the artefacts.
#pragma once
#include <type_traits>
#include "tsm_device.cuh"

namespace tsm {

__constant__ uint32_t c_lut[256];                       // automaton byte classes (one table, tsm_device.cuh)
__constant__ uint32_t c_elut[256];                      // bare-assert operator automaton (k_classify)
// category tables: read once per block of k_classify into shared memory (coalesced, hence plain device memory)
__device__ uint8_t c_cat_slot[TSM_CAT_SLOTS];            // perfect hash slot -> category id
__device__ uint16_t c_cat_off[TSM_CAT_NAMED + 1];
__device__ char c_cat_blob[TSM_CAT_BLOB_LEN + 1];

// ================================================================================= k_plan
// One lane per file: units = ceil(len / CH); the warp reserves a contiguous range of the unit
// table with one atomic.  Unit order is irrelevant for the results (all outputs are sums or sets).
__global__ void k_plan(ScanParams p) {
  const int f = p.f_begin + blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t nu = 0;
  if (f < p.f_end) {
    nu = ((uint32_t)p.len[f] + CH - 1) / CH;
    if (nu != 1) p.stats[f] = tsm_file_stat{0, 0, 0, 0, 0};   // several chunks add into it (one chunk: k_scan stores), none leave it zero
  }
  uint32_t incl = nu;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t base = 0;
  if (lane == 31 && total) base = atomicAdd(&p.slab->n_units, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  uint32_t at = p.unit_base + base + incl - nu;
  for (uint32_t u = 0; u < nu; ++u, ++at) {
    if (at < p.unit_cap) { p.unit_file[at] = (uint32_t)f; p.unit_begin[at] = u * CH; }
    else p.ctrl->overflow = 1;
  }
}

// ================================================================================= k_scan
// Work unit = (file, 4 KiB chunk).  Per warp, per unit:
//   stage   one 1-D TMA bulk copy (cp.async.bulk + mbarrier) of [chunk-16, chunk+4096+240) into shared
//   pass 1  SWAR newline bits, one 136-byte stripe per lane; one warp scan orders them into the line table
//   pass 2  stripe walk (all lanes busy whatever the line lengths): Shift-And automaton (one LDS per
//           byte), pattern ends -> per-line flag words, Mersenne-61 running prefix (checkpoint every 4 words)
//   pass 2b words with a newline AND a pattern end, byte by byte, one lane per word
//   pass 3  balanced finalise, one lane per line: hash = difference of two prefixes, header / assertion flags
//   pass 4  ballot compaction of candidate (and header-event) lines into the global lists
// SWAR: 16-bit mask of the bytes equal to '\n' in a 16-byte vector.
__device__ __forceinline__ uint32_t nl_word(uint32_t w) {
  const uint32_t y = w ^ 0x0A0A0A0Au;
  const uint32_t t = (y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
  const uint32_t z = ~(t | y | 0x7F7F7F7Fu);            // 0x80 in every byte that was '\n'
  return (z * 0x00204081u) >> 28;                        // gather the four flag bits: 7+21, 15+14, 23+7, 31+0 -> 28..31
}
__device__ __forceinline__ uint32_t nl16(const uint4& v) {
  return nl_word(v.x) | (nl_word(v.y) << 4) | (nl_word(v.z) << 8) | (nl_word(v.w) << 12);
}

// Per-lane state of the line currently walked by this lane.
struct LineState {
  uint32_t s, e;            // [s, e) = line
  uint32_t pos;             // next 8-byte block to process
  uint32_t D, A;            // automaton state / OR of all states
  unsigned long long B;     // Horner accumulator: B_k = B_{k-1} * 2^-64 + X_k  (mod 2^61-1)
};

__device__ __forceinline__ void line_init(LineState& L, uint32_t s, uint32_t e) {
  L.s = s; L.e = e; L.D = 0; L.A = 0; L.B = 0;
  L.pos = (s == e) ? e : (s & ~7u);
}

// One 8-byte block of a lane-per-line walk (long-line slow path, k_hash_lines).  Files without a scannable
// extension run the same code: their pattern ends are simply never looked at.
__device__ __forceinline__ void line_block(LineState& L, unsigned long long w, const uint32_t* lut, uint32_t first) {
  const uint32_t pos = L.pos;
  if (pos < L.s || pos + 8 > L.e) {                      // first / last block: zero the bytes outside the line
    unsigned long long m = ~0ull;
    if (pos < L.s) m <<= 8u * (L.s - pos);
    if (pos + 8 > L.e) m &= ~0ull >> (8u * (pos + 8 - L.e));
    w &= m;
  }
  {
    const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
    uint32_t D = L.D, A = L.A;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      D = ((D + D) | first) & lut[__byte_perm(lo, 0, 0x4440 + k)];
      A |= D;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      D = ((D + D) | first) & lut[__byte_perm(hi, 0, 0x4440 + k)];
      A |= D;
    }
    L.D = D; L.A = A;
  }
  // B = B * 2^-64 + w  (2^-64 = 2^-3 = 2^58 mod 2^61-1: a rotation by 3 to the right)
  const unsigned long long b = L.B;
  const unsigned long long rot = (b >> 3) | ((b & 7ull) << 58);
  L.B = fold61(fold61(rot + fold61(w)));
  L.pos = pos + 8;
}

// Hash-only variant (S8 line hashes): same masking and Horner step, no automaton.
__device__ __forceinline__ void hash_block(LineState& L, unsigned long long w) {
  const uint32_t pos = L.pos;
  if (pos < L.s || pos + 8 > L.e) {
    unsigned long long m = ~0ull;
    if (pos < L.s) m <<= 8u * (L.s - pos);
    if (pos + 8 > L.e) m &= ~0ull >> (8u * (pos + 8 - L.e));
    w &= m;
  }
  const unsigned long long b = L.B;
  const unsigned long long rot = (b >> 3) | ((b & 7ull) << 58);
  L.B = fold61(fold61(rot + fold61(w)));
  L.pos = pos + 8;
}

struct SmemByte {                                        // byte source = the staged chunk
  const uint8_t* b;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return b[i]; }
  // the 8 bytes at i, any alignment (shared memory is readable 8 bytes past any line of the buffer)
  __device__ __forceinline__ unsigned long long load8(uint32_t i) const {
    const uint32_t a = i & ~7u, sh = 8u * (i & 7u);
    const unsigned long long lo = *reinterpret_cast<const unsigned long long*>(b + a);
    const unsigned long long hi = *reinterpret_cast<const unsigned long long*>(b + a + 8);
    return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
  }
  // number of leading 0x20 bytes among the 8 bytes at i
  __device__ __forceinline__ uint32_t spaces8(uint32_t i) const {
    const unsigned long long w = load8(i);
    const unsigned long long x = w ^ 0x2020202020202020ull, k7 = 0x7F7F7F7F7F7F7F7Full;
    const unsigned long long nz = (((x & k7) + k7) | x) & ~k7;           // 0x80 in every byte that is not a space
    return nz ? ((uint32_t)__ffsll((long long)nz) - 1u) >> 3 : 8u;
  }
};
struct GmemByte {                                        // byte source = the file in HBM (slow path)
  const uint8_t* b;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return __ldg(b + i); }
};

template <typename LoadByte>
__device__ __forceinline__ bool starts_with(LoadByte lb, uint32_t s, uint32_t e, const char* pat, int n, bool need_ws) {
  if constexpr (std::is_same<LoadByte, SmemByte>::value) {  // staged bytes: runs of spaces eight at a time
    uint32_t r;
    while (s + 8 <= e && (r = lb.spaces8(s)) != 0) { s += r; if (r < 8) break; }
  }
  while (s < e && is_w(lb(s))) ++s;
  if (s + n + (need_ws ? 1 : 0) > e) return false;
  for (int k = 0; k < n; ++k)
    if (lb(s + k) != (uint8_t)pat[k]) return false;
  if (need_ws) {                                         // the blank must be inside the STRIPPED line:
    const uint32_t c = lb(s + n);                        // some non-blank byte has to follow it
    if (c != 0x20 && c != 0x09) return false;
    for (uint32_t q = s + n + 1; q < e; ++q)
      if (!is_w(lb(q))) return true;
    return false;
  }
  return true;
}

struct Accum { uint32_t lines, asserts, hdrs, fixes; unsigned long long digest; };


// The four facts pass 3 needs about a line, as one nibble: bit 0 = assertion pattern, bits 1..3 = the
// language's header patterns (PY: def, class, TEST_F gate; C family: test, one of { class void, TEST_F gate).
__device__ __forceinline__ uint32_t flag_nibble(uint32_t A, uint32_t g1, uint32_t g2, uint32_t g3) {
  return ((A & (AF_ASSERT | AF_EXPECT)) ? 1u : 0u) | ((A & g1) ? 2u : 0u) | ((A & g2) ? 4u : 0u) | ((A & g3) ? 8u : 0u);
}
__device__ __forceinline__ uint32_t flag_nibble_ext(uint32_t A, int ext) {
  return ext == TSM_EXT_PY ? flag_nibble(A, PY_G1, PY_G2, A_STF) : flag_nibble(A, CJ_G1, CJ_G2, A_STF);
}

// Finish one line: h0 = Mersenne-61 value of its bytes (SPEC section 3, trailing CR still inside), nib = its
// pattern nibble.  Adds the line to the per-file accumulators and returns its LF_* flags (SPEC sections 4/5).
template <typename LoadByte>
__device__ __forceinline__ uint32_t line_finish_h(uint32_t s, uint32_t e, unsigned long long h0, uint32_t nib, int ext,
                                                  LoadByte lb, Accum& ac) {
  uint32_t len = e - s;
  unsigned long long h = 0;
  if (len) {
    h = h0;
    if (lb(e - 1) == 0x0D) {                             // drop one trailing CR: subtract 0x0D * 256^(len-1)
      --len;
      const unsigned long long cr = rotl61(0x0Dull, (8u * len) % 61u);
      h = h >= cr ? h - cr : h + M61 - cr;
    }                                                    // h0 is canonical (< 2^61 - 1) and the CR step keeps it so
  }
  ac.lines++;
  ac.digest += mix_hash(h, len);
  if (ext == 0) return 0;
  uint32_t fl = (nib & 1u) ? LF_CAND : 0;
  bool hdr;
  if (ext == TSM_EXT_PY) {
    hdr = (nib & 2u) != 0;
    if (!hdr && (nib & 4u)) hdr = starts_with(lb, s, e, "class", 5, true);
  } else {
    hdr = (nib & 6u) == 6u;
  }
  if (hdr) { fl |= LF_HDR; if (starts_with(lb, s, e, "TEST_F", 6, false)) fl |= LF_FIX; }   // (headers are ~2 % of the lines)
  ac.asserts += fl & LF_CAND;
  ac.hdrs += (fl >> 1) & 1u;
  ac.fixes += (fl >> 2) & 1u;
  return fl;
}

// Same, from the raw (A, B) of a lane-per-line walk (the long-line slow path and k_hash_lines).
template <typename LoadByte>
__device__ __forceinline__ uint32_t line_finish(uint32_t s, uint32_t e, uint32_t A, unsigned long long B, int ext,
                                                LoadByte lb, Accum& ac) {
  unsigned long long h0 = 0;
  if (e != s) {
    // N * 2^(8*lead) = B * 2^(64*(m-1)), m = number of 8-byte blocks the line touches
    const uint32_t lead = s & 7u, m = ((e - 1) >> 3) - (s >> 3) + 1;
    h0 = rotl61(canon61(B), (3u * (m - 1) + 61u * 8u - 8u * lead) % 61u);
  }
  return line_finish_h(s, e, h0, ext ? flag_nibble_ext(A, ext) : 0u, ext, lb, ac);
}

// Append `n` list entries with one atomic; returns the base slot (broadcast from lane 0).
__device__ __forceinline__ uint32_t warp_reserve(uint32_t* counter, uint32_t n, int lane) {
  uint32_t base = 0;
  if (lane == 0 && n) base = atomicAdd(counter, n);
  return __shfl_sync(0xffffffffu, base, 0);
}

// The automaton table of k_scan sits at the start of the dynamic shared memory: a compile-time address.
__device__ __forceinline__ const uint32_t* scan_lut() {
  extern __shared__ __align__(128) uint8_t smem[];
  return reinterpret_cast<const uint32_t*>(smem);
}

// Eight automaton steps over one 8-byte word; A collects every state of the word.
__device__ __forceinline__ void step8(unsigned long long w, uint32_t& D, uint32_t& A) {
  const uint32_t* lut = scan_lut();
  const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    D = ((D + D) | A_FIRST) & lut[__byte_perm(lo, 0, 0x4440 + k)];
    A |= D;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    D = ((D + D) | A_FIRST) & lut[__byte_perm(hi, 0, 0x4440 + k)];
    A |= D;
  }
}

// OR automaton states into the window's per-line flag words (lines outside the window: dropped).
__device__ __forceinline__ void flag_or(uint8_t* wb, uint32_t rel, uint32_t A) {
  if (rel < FLAG_CAP) atomicOr(reinterpret_cast<uint32_t*>(wb + OFF_FLAGS) + rel, A);
}

// Pass 2b: a word that holds both a pattern end and a newline - walk its bytes one at a time so that
// every match lands on its own line.  entry = word index | line index at the word's first byte << 10.
// The automaton state in front of the word follows from the eight bytes before it (no pattern is longer).
__device__ __noinline__ void resolve_word(uint8_t* wb, uint32_t fin, uint32_t wlo, uint32_t entry) {
  const uint32_t k = entry & 1023u, pos = 8u * k;
  uint32_t idx = (entry >> 10) - wlo;
  const uint32_t l = k / 17u, i = k - 17u * l;
  uint32_t nl8 = (reinterpret_cast<const uint32_t*>(wb + OFF_MSK)[(i >> 2) * 32u + l] >> (8u * (i & 3u))) & 0xFFu;
  uint32_t D = 0, A = 0;
  if (pos > PRE) step8(*reinterpret_cast<const unsigned long long*>(wb + pos - 8), D, A);
  unsigned long long w = *reinterpret_cast<const unsigned long long*>(wb + pos);
  const uint32_t* lut = scan_lut();
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    D = ((D + D) | A_FIRST) & lut[(uint32_t)w & 0xFFu];
    if (D & fin) flag_or(wb, idx, D);
    idx += nl8 & 1u;
    nl8 >>= 1;
    w >>= 8;
  }
}

// Pass 2: every lane walks the 17 words of its own 136-byte stripe (all 32 lanes busy for the whole
// pass, whatever the line lengths are; the bytes outside the chunk's staged range are zeros):
//   automaton   8 steps per word; the OR of the word's states goes to the flag word of the line the
//               word lies in (its index follows from the stripe's newline bits); a word with a
//               newline AND a pattern end goes to the pass-2b queue instead
//   hash        R_k = R_{k-1} * 2^-64 + w_k (mod 2^61-1) is stored behind every word: pass 3 gets the
//               Mersenne value of any byte range as a difference of two such prefixes
// then one warp scan turns the stripe totals into the absolute prefix at every stripe start.
// Not inlined on purpose: the hot loop gets its own register allocation.
__device__ __noinline__ void walk_stripes(uint8_t* wb, uint32_t fin, uint32_t wlo, uint32_t base_all, int lane) {
  const uint32_t pos0 = (uint32_t)lane * STRIPE;
  uint8_t* sp = wb + pos0;
  const uint32_t* msk = reinterpret_cast<const uint32_t*>(wb + OFF_MSK) + lane;
  uint32_t* flags = reinterpret_cast<uint32_t*>(wb + OFF_FLAGS);
  unsigned long long* rw = reinterpret_cast<unsigned long long*>(wb + OFF_RW) + (uint32_t)lane * RW_PER_STRIPE;
  uint32_t D = 0;
  if (lane) {                                            // state in front of the stripe
    uint32_t A = 0;
    step8(*reinterpret_cast<const unsigned long long*>(sp - 8), D, A);
  }
  unsigned long long R = 0;
  uint32_t idxg = base_all - wlo;                        // window-relative line index at the group's first byte
#pragma unroll 1
  for (uint32_t g = 0; g < 5; ++g) {
    const uint32_t mg = msk[g * 32u];
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {
      if (g == 4 && k) break;
      const uint32_t off = 32u * g + 8u * k;
      const unsigned long long w = *reinterpret_cast<const unsigned long long*>(sp + off);
      uint32_t A = 0;
      step8(w, D, A);
      R = (R >> 3) + ((R & 7ull) << 58) + fold61(w);      // lazily reduced: stays below 2^63
      if ((k & (RW_STRIDE - 1u)) == RW_STRIDE - 1u && g < 4)   // checkpoint behind every RW_STRIDE-th word (not the 17th)
        rw[(4u * g + k) >> RW_SHIFT] = R;
      const uint32_t nl8 = (mg >> (8u * k)) & 0xFFu;
      const uint32_t idx = idxg + __popc(mg & ((1u << (8u * k)) - 1u));
      atomicOr(flags + min(idx, FLAG_CAP - 1u), nl8 ? 0u : A);            // entry FLAG_CAP-1 is never a line
      if (nl8 && (A & fin)) {
        const uint32_t slot = atomicAdd(reinterpret_cast<uint32_t*>(wb + OFF_CTL), 1u);
        const uint32_t entry = ((pos0 + off) >> 3) | ((idx + wlo) << 10);
        if (slot < Q_CAP) reinterpret_cast<uint32_t*>(wb + OFF_Q)[slot] = entry;
        else resolve_word(wb, fin, wlo, entry);
      }
    }
    idxg += __popc(mg);
  }
  // stripe totals (frame of the stripe's last word) -> absolute frame -> exclusive scan
  unsigned long long incl = rotl61(canon61(R), (3u * (17u * (uint32_t)lane + 16u)) % 61u);
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl = fold61(incl + t);
  }
  unsigned long long excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0;
  unsigned long long* sbase = reinterpret_cast<unsigned long long*>(wb + OFF_BASE);
  sbase[lane] = excl;
  if (lane == 31) sbase[32] = incl;
  __syncwarp();
  const uint32_t nq = min(*reinterpret_cast<const uint32_t*>(wb + OFF_CTL), Q_CAP);
  for (uint32_t t = (uint32_t)lane; t < nq; t += 32)
    resolve_word(wb, fin, wlo, reinterpret_cast<const uint32_t*>(wb + OFF_Q)[t]);
  __syncwarp();
}

// Mersenne-61 value of the staged bytes [0, x), every byte weighted 256^position; lazily reduced (< 2^62 + 2).
// x = 8k + b, r3k = 3k mod 61.  Starts from the last checkpoint in front of word k and folds the words between.
__device__ __forceinline__ unsigned long long prefix_at(const uint8_t* wb, uint32_t k, uint32_t b, uint32_t r3k) {
  const uint32_t l = k / 17u, i = k - 17u * l, c0 = i >> RW_SHIFT;
  unsigned long long R = 0;                              // prefix inside the stripe behind word k-1, frame of word k-1
  if (c0) R = *reinterpret_cast<const unsigned long long*>(wb + OFF_RW + 8u * (l * RW_PER_STRIPE + c0 - 1u));
#pragma unroll 1
  for (uint32_t t = c0 << RW_SHIFT; t < i; ++t)          // at most RW_STRIDE - 1 words
    R = (R >> 3) + ((R & 7ull) << 58) + fold61(*reinterpret_cast<const unsigned long long*>(wb + 8u * (k - i + t)));
  unsigned long long loc = (R >> 3) + ((R & 7ull) << 58);                // frame of word k
  if (b) loc += *reinterpret_cast<const unsigned long long*>(wb + 8u * k) & ((1ull << (8u * b)) - 1ull);
  return *reinterpret_cast<const unsigned long long*>(wb + OFF_BASE + 8u * l) + rotl61(fold61(fold61(loc)), r3k);
}

// Pass 3: balanced finalise, one lane per line: hash = difference of two prefixes, flags from the
// line's flag word; the starts of the candidate lines are compacted (u16 each) over the flag words
// already consumed.  Returns the number of candidates.
__device__ __noinline__ uint32_t finish_lines(uint8_t* wb, const uint32_t* lc, uint32_t j0, uint32_t cnt, uint32_t ns,
                                              int ext, int lane, Accum& ac) {
  uint16_t* tab = reinterpret_cast<uint16_t*>(wb + OFF_TAB);
  const uint32_t* flags = reinterpret_cast<const uint32_t*>(wb + OFF_FLAGS);
  uint16_t* clist = reinterpret_cast<uint16_t*>(wb + OFF_FLAGS);
  const uint32_t g1 = lc[1], g2 = lc[2], g3 = lc[3];
  const SmemByte lb{wb};
  Accum a = ac;
  uint32_t nc = 0;
  uint32_t cs = j0 ? ((uint32_t)tab[j0 - 1] & TAB_POS) + 1u : ns;       // start of the round's first line
  unsigned long long cP = prefix_at(wb, cs >> 3, cs & 7u, (3u * (cs >> 3)) % 61u);   // ... and the prefix in front of it
  for (uint32_t base = j0; base < cnt; base += 32) {                     // uniform trip count
    const uint32_t j = base + (uint32_t)lane;
    const bool valid = j < cnt;
    const uint32_t e = valid ? (uint32_t)tab[j] & TAB_POS : 0u;
    const uint32_t A = valid ? flags[j] : 0u;
    const uint32_t k = e >> 3, b = e & 7u, r3k = (3u * k) % 61u;
    uint32_t r8e = r3k + 8u * b;                                         // (8 * e) mod 61
    if (r8e >= 61u) r8e -= 61u;
    const unsigned long long Pe = prefix_at(wb, k, b, r3k);
    const unsigned long long Pn = Pe + rotl61(0x0Aull, r8e);             // prefix behind the terminator (< 2^63 - 4)
    uint32_t s = __shfl_up_sync(0xffffffffu, e, 1) + 1u;
    unsigned long long Ps = __shfl_up_sync(0xffffffffu, Pn, 1);
    if (lane == 0) { s = cs; Ps = cP; }
    cs = __shfl_sync(0xffffffffu, e, 31) + 1u;
    cP = __shfl_sync(0xffffffffu, Pn, 31);
    uint32_t fl = 0;
    if (valid) {
      const unsigned long long hr = canon61(Pe + 4ull * M61 - Ps);       // bytes [s, e), weighted from position 0
      const unsigned long long h0 = rotl61(hr, (61u - (8u * s) % 61u) % 61u);
      fl = line_finish_h(s, e, h0, flag_nibble(A, g1, g2, g3), ext, lb, a);
      tab[j] = (uint16_t)(e | (fl << 13));               // flags ride in the 3 spare bits of the entry
    }
    __syncwarp();                                        // every flag word of the round is read: the list may grow over them
    const uint32_t mc = __ballot_sync(0xffffffffu, fl & LF_CAND);
    if (fl & LF_CAND) clist[nc + __popc(mc & ((1u << lane) - 1u))] = (uint16_t)s;
    nc += __popc(mc);
  }
  ac = a;
  __syncwarp();
  return nc;
}

// Passes 3-4 over the current line table: lines j in [j0, cnt), line j = [start_j, tab[j]).
__device__ __forceinline__ void drain(const ScanParams& p, uint8_t* wb, const uint32_t* lc, uint32_t cnt,
                                      bool& skip_first, uint32_t& next_start, uint32_t f, uint32_t cb, int ext, int lane,
                                      Accum& ac) {
  __syncwarp();
  if (cnt == 0) return;
  const uint16_t* tab = reinterpret_cast<const uint16_t*>(wb + OFF_TAB);
  const uint32_t j0 = skip_first ? 1u : 0u;
  const uint32_t ns = next_start;
  const uint32_t nc = finish_lines(wb, lc, j0, cnt, ns, ext, lane, ac);   // pass 3
  // ---- pass 4: candidates (always) and header events (on request) to their global lists
  if (nc) {
    const uint16_t* clist = reinterpret_cast<const uint16_t*>(wb + OFF_FLAGS);
    const uint32_t cbase = warp_reserve(&p.ctrl->n_cand, nc, lane);
    for (uint32_t i = (uint32_t)lane; i < nc; i += 32) {
      const uint32_t slot = cbase + i;
      if (slot < p.cand_cap) p.cand[slot] = ((unsigned long long)f << 32) | (cb + (uint32_t)clist[i] - PRE);
      else p.ctrl->overflow = 1;
    }
  }
  if ((p.flags & TSM_SCAN_HEADER_EVENTS) && ext != 0) {
    uint32_t nh = 0;
    for (uint32_t b = j0; b < cnt; b += 32) {
      const uint32_t j = b + lane;
      const uint32_t fb = j < cnt ? (uint32_t)tab[j] >> 13 : 0u;
      nh += __popc(__ballot_sync(0xffffffffu, fb & LF_HDR));
    }
    if (nh) {
      uint32_t hbase = warp_reserve(&p.ctrl->n_hev, nh, lane);
      for (uint32_t b = j0; b < cnt; b += 32) {
        const uint32_t j = b + lane;
        const uint32_t fb = j < cnt ? (uint32_t)tab[j] >> 13 : 0u;
        uint32_t s = 0, e = 0;
        if (j < cnt) { s = j ? ((uint32_t)tab[j - 1] & TAB_POS) + 1u : ns; e = (uint32_t)tab[j] & TAB_POS; }
        const uint32_t mh = __ballot_sync(0xffffffffu, fb & LF_HDR);
        if (fb & LF_HDR) {
          const uint32_t slot = hbase + __popc(mh & ((1u << lane) - 1u));
          if (slot < p.hev_cap) p.hev[slot] = tsm_header_event{f, cb + s - PRE, e - s, (fb >> 2) & 1u};
          else p.ctrl->overflow = 1;
        }
        hbase += __popc(mh);
      }
    }
  }
  next_start = ((uint32_t)tab[cnt - 1] & TAB_POS) + 1u;
  skip_first = false;
  __syncwarp();
}

// Slow path: a line that starts in this chunk but ends behind the staged bytes.  Walked by lane 0
// straight from HBM (correct for any length; lines longer than 240 B past a chunk edge are rare).
__device__ __noinline__ uint32_t long_line(const ScanParams& p, const uint32_t* lut, uint32_t first, uint32_t f,
                                           uint32_t fo, uint32_t size, int ext, uint32_t s, Accum& ac, uint32_t* fl_out = nullptr) {
  const uint8_t* g = p.arena + fo;
  const GmemByte lb{g};
  uint32_t e = s;
  while (e < size && lb(e) != '\n') ++e;
  LineState L;
  line_init(L, s, e);
  while (L.pos < L.e) {
    const unsigned long long w = __ldg(reinterpret_cast<const unsigned long long*>(g + L.pos));
    line_block(L, w, lut, first);
  }
  const uint32_t fl = line_finish(s, e, L.A, L.B, ext, lb, ac);
  if ((fl & LF_CAND) && p.cand_cap) {
    const uint32_t slot = atomicAdd(&p.ctrl->n_cand, 1u);
    if (slot < p.cand_cap) p.cand[slot] = ((unsigned long long)f << 32) | s;
    else p.ctrl->overflow = 1;
  }
  if ((fl & LF_HDR) && (p.flags & TSM_SCAN_HEADER_EVENTS)) {
    const uint32_t slot = atomicAdd(&p.ctrl->n_hev, 1u);
    if (slot < p.hev_cap) p.hev[slot] = tsm_header_event{f, s, e - s, (fl >> 2) & 1u};
    else p.ctrl->overflow = 1;
  }
  if (fl_out) *fl_out = fl;
  return e;                                              // file-relative end of the line
}

__device__ __forceinline__ void process_chunk(const ScanParams& p, const uint32_t* lc, uint8_t* wb, uint32_t f,
                                              uint32_t cb, uint32_t fo, uint32_t size, int ext, int lane) {
  const uint8_t* buf = wb;
  uint16_t* tab = reinterpret_cast<uint16_t*>(wb + OFF_TAB);
  const uint32_t ce = min(cb + CH, size);
  const uint32_t le = min(ce + EXT, size);
  const uint32_t lim = PRE + (ce - cb);                  // buffer position just past the owned bytes
  const uint32_t lim2 = PRE + (le - cb);                 // ... past the staged bytes
  bool skip_first = (cb != 0) && (buf[PRE - 1] != '\n'); // chunk starts inside a foreign line
  uint32_t next_start = PRE;
  Accum ac{0, 0, 0, 0, 0};
  uint32_t cnt = 0;
  __syncwarp();
  // ---- everything outside the staged range [PRE, lim2) becomes zeros: no pass has to mask its loads
  //      (a zero byte is no newline, matches no pattern and adds nothing to the hash prefix)
  if (lane < 2) reinterpret_cast<unsigned long long*>(wb)[lane] = 0ull;
  {
    const uint32_t za = (lim2 + 7u) & ~7u;
    if ((uint32_t)lane < za - lim2) wb[lim2 + lane] = 0;
    for (uint32_t q = za + 8u * (uint32_t)lane; q < BUF; q += 256u) *reinterpret_cast<unsigned long long*>(wb + q) = 0ull;
  }
  __syncwarp();
  // ---- pass 1: every lane takes one 136-byte stripe of the 4 352 staged bytes (17 conflict-free
  //      LDS.64) and keeps the newline positions of its stripe as a 136-bit mask in registers;
  //      one warp scan then orders them into the u16 line table, NL_CAP entries per window
  const uint32_t pos0 = (uint32_t)lane * STRIPE;
  uint32_t* msk = reinterpret_cast<uint32_t*>(wb + OFF_MSK) + lane;       // msk[g * 32]: newline bits of the stripe's bytes [32g, 32g+32)
  uint32_t own[5], extm[5];                              // newlines at positions [PRE, lim) / [lim, lim2)
  {
    uint32_t m[5] = {0u, 0u, 0u, 0u, 0u};
#pragma unroll
    for (int i = 0; i < 17; ++i) {
      const uint2 w = *reinterpret_cast<const uint2*>(buf + pos0 + 8u * (uint32_t)i);
      m[i >> 2] |= (nl_word(w.x) | (nl_word(w.y) << 4)) << (8 * (i & 3));
    }
    const uint32_t rel = lim > pos0 ? lim - pos0 : 0u;   // owned bytes of this stripe (may exceed 136)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const uint32_t below = rel >= 32u * (j + 1) ? 0xFFFFFFFFu : (rel <= 32u * j ? 0u : (1u << (rel - 32u * j)) - 1u);
      own[j] = m[j] & below;
      extm[j] = m[j] & ~below;
      msk[j * 32] = m[j];                                // pass 2 reads them back
    }
  }
  const uint32_t mine = __popc(own[0]) + __popc(own[1]) + __popc(own[2]) + __popc(own[3]) + __popc(own[4]);
  const uint32_t mine_ext = __popc(extm[0]) + __popc(extm[1]) + __popc(extm[2]) + __popc(extm[3]) + __popc(extm[4]);
  uint32_t incl = mine | (mine_ext << 16);               // both counts in one scan (each < 2^13)
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31) & 0xFFFFu;
  const uint32_t base_idx = (incl & 0xFFFFu) - mine;
  const uint32_t base_all = base_idx + (incl >> 16) - mine_ext;          // newlines (of either kind) in front of the stripe
  uint32_t ext_first = 0xFFFFu;                          // first newline behind the owned bytes
#pragma unroll
  for (int j = 4; j >= 0; --j)
    if (extm[j]) ext_first = pos0 + 32u * (uint32_t)j + (uint32_t)(__ffs(extm[j]) - 1);
  ext_first = __reduce_min_sync(0xffffffffu, ext_first);
  for (uint32_t wstart = 0;; wstart += NL_CAP) {
    cnt = min(total - wstart, NL_CAP);
    uint32_t idx = base_idx - wstart;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      uint32_t b = own[j];
      while (b) {
        if (idx < cnt) tab[idx] = (uint16_t)(pos0 + 32u * (uint32_t)j + (uint32_t)(__ffs(b) - 1));
        ++idx;
        b &= b - 1;
      }
    }
    for (uint32_t i = (uint32_t)lane; i < cnt + 2u; i += 32) reinterpret_cast<uint32_t*>(wb + OFF_FLAGS)[i] = 0u;
    if (lane == 0) *reinterpret_cast<uint32_t*>(wb + OFF_CTL) = 0u;
    __syncwarp();
    walk_stripes(wb, lc[0], wstart, base_all, lane);         // pass 2 (the window's flag words)
    if (wstart + cnt >= total) break;                    // last window: the tail line joins it below
    drain(p, wb, lc, cnt, skip_first, next_start, f, cb, ext, lane, ac);
  }
  // ---- the last owned line: starts in the chunk, may end behind it (its line index = total)
  const uint32_t tail_start = cnt ? ((uint32_t)tab[cnt - 1] & TAB_POS) + 1u : next_start;
  const bool owned = !(skip_first && cnt == 0);
  bool have_tail = false, tail_long = false;
  uint32_t tail_end = 0;
  if (owned && tail_start < lim) {
    if (ce == size) { tail_end = lim; have_tail = true; }              // unterminated last line of the file
    else if (ext_first != 0xFFFFu) { tail_end = ext_first; have_tail = true; }
    else if (le == size) { tail_end = lim2; have_tail = true; }        // file ends inside the staged bytes
    else tail_long = true;
  }
  if (have_tail) {                                       // the table has room for NL_CAP + 1 entries
    if (lane == 0) tab[cnt] = (uint16_t)tail_end;
    ++cnt;
  }
  drain(p, wb, lc, cnt, skip_first, next_start, f, cb, ext, lane, ac);
  if (tail_long && lane == 0) long_line(p, scan_lut(), A_FIRST, f, fo, size, ext, cb + tail_start - PRE, ac);
  // ---- per-file counters: warp reduce (the digest as three partial sums: low halves keep their carries),
  //      then one store (single-chunk file) or one atomic per counter
  ac.lines = __reduce_add_sync(0xffffffffu, ac.lines);
  ac.asserts = __reduce_add_sync(0xffffffffu, ac.asserts);
  ac.hdrs = __reduce_add_sync(0xffffffffu, ac.hdrs);
  ac.fixes = __reduce_add_sync(0xffffffffu, ac.fixes);
  {
    const uint32_t dlo = (uint32_t)ac.digest, dhi = (uint32_t)(ac.digest >> 32);
    const unsigned long long s0 = __reduce_add_sync(0xffffffffu, dlo & 0xFFFFu);
    const unsigned long long s1 = __reduce_add_sync(0xffffffffu, dlo >> 16);
    const unsigned long long s2 = __reduce_add_sync(0xffffffffu, dhi);
    ac.digest = s0 + (s1 << 16) + (s2 << 32);
  }
  if (lane == 0) {
    tsm_file_stat* st = p.stats + f;
    if (size <= CH) {                                    // sole owner of the record: plain store
      *st = tsm_file_stat{ac.lines, ac.asserts, ac.hdrs, ac.fixes, ac.digest};
    } else {
      if (ac.lines) atomicAdd(&st->n_lines, ac.lines);
      if (ac.asserts) atomicAdd(&st->n_assert, ac.asserts);
      if (ac.hdrs) atomicAdd(&st->n_headers, ac.hdrs);
      if (ac.fixes) atomicAdd(&st->n_fixture, ac.fixes);
      if (ac.digest) atomicAdd(reinterpret_cast<unsigned long long*>(&st->digest), ac.digest);
    }
  }
}

// Stage the bytes [max(cb-16,0), min(cb+CH+EXT, size)) of a file so that file byte cb sits at buf+PRE.
__device__ __forceinline__ void issue_load(const ScanParams& p, uint8_t* buf, uint64_t* bar, uint32_t fo,
                                           uint32_t size, uint32_t cb) {
  const uint32_t lb = cb ? cb - PRE : 0u;
  const uint32_t le = min(cb + CH + EXT, size);
  const uint32_t bytes = (le - lb + 15u) & ~15u;          // the pad up to the 128-B file boundary is readable
  mbar_expect_tx(bar, bytes);
  bulk_load(buf + PRE - (cb - lb), p.arena + (size_t)fo + lb, bytes, bar);
}

struct Unit { uint32_t u, f, cb, fo, size; int ext; };

// Claim the next work unit and fetch its metadata (4 dependent global loads: issued one chunk
// ahead so that their latency hides behind the current chunk).
__device__ __forceinline__ Unit claim_unit(const ScanParams& p, uint32_t n_units, int lane) {
  Unit x{0, 0, 0, 0, 0, 0};
  if (lane == 0) x.u = atomicAdd(&p.slab->work, 1u);
  x.u = __shfl_sync(0xffffffffu, x.u, 0);
  if (x.u < n_units) {
    x.f = p.unit_file[p.unit_base + x.u];
    x.cb = p.unit_begin[p.unit_base + x.u];
    x.fo = (uint32_t)p.off[x.f];
    x.size = (uint32_t)p.len[x.f];
    x.ext = p.ext[x.f];
  }
  return x;
}

__global__ void __launch_bounds__(SCAN_WARPS * 32, SCAN_CTAS_PER_SM) k_scan(ScanParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t* lut_all = reinterpret_cast<uint32_t*>(smem);  // [0,256) the automaton table, then 3 x 4 per-language masks
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut_all[i] = c_lut[i];
  if (threadIdx.x < 12) {                                // per language (PY, C family, none): all pattern ends that count, then the header groups
    const int t = threadIdx.x, lang = t >> 2, q = t & 3;
    const uint32_t g1 = lang == 0 ? PY_G1 : CJ_G1, g2 = lang == 0 ? PY_G2 : CJ_G2;
    const uint32_t v = q == 0 ? (AF_ASSERT | AF_EXPECT | g1 | g2 | A_STF) : (q == 1 ? g1 : (q == 2 ? g2 : A_STF));
    lut_all[256 + t] = lang == 2 ? 0u : v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wb = smem + LUT_BYTES + warp * WARP_SMEM;
  uint64_t* bar = reinterpret_cast<uint64_t*>(wb + OFF_CTL + 8);
  if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncwarp();
  const uint32_t n_units = p.slab->n_units;
  uint32_t phase = 0;
  Unit cur = claim_unit(p, n_units, lane);
#if TSM_LOCKSTEP
  // the warps of a CTA start every chunk together: they then run the same pass at about the same time and
  // share its instructions in the SM's instruction caches (the hot code is larger than the 32 KB L1.5;
  // measured -6 % kernel time, profiles/README.md)
  while (__syncthreads_or(cur.u < n_units)) {
    if (cur.u >= n_units) continue;
#else
  while (cur.u < n_units) {
#endif
    fence_proxy_async();                                 // this warp's zero fill and reads of the last chunk come first
    __syncwarp();
    if (lane == 0) issue_load(p, wb, bar, cur.fo, cur.size, cur.cb);
    const Unit nxt = claim_unit(p, n_units, lane);       // metadata of the next unit arrives during this chunk
    while (!mbar_try_wait(bar, phase)) {}
    phase ^= 1;
    const uint32_t lang = cur.ext == 0 ? 2u : (cur.ext == TSM_EXT_PY ? 0u : 1u);
    process_chunk(p, lut_all + 256u + 4u * lang, wb, cur.f, cur.cb, cur.fo, cur.size, cur.ext, lane);
    __syncwarp();
    cur = nxt;
  }
}

// ================================================================================= k_classify
// One thread per candidate line, exactly SPEC sections 4 and 6.  The statement T ends at the first '(';
// one pass over its bytes (8-byte loads, one shared-memory class lookup per byte) finds the
// stripped start, the right-stripped end and the last identifier L; everything else (gtest stem,
// bare-assert operators, table lookup of L, statement hash) touches only the few bytes it needs.
constexpr uint32_t CC_W = 1, CC_IDENT = 2, CC_STOP = 4;  // byte classes: blank, [A-Za-z0-9_], '(' or LF
constexpr uint32_t CLS_CAT_OFF_N = (TSM_CAT_NAMED + 2) & ~1u;            // u16 entries (even count keeps the blob 4-byte aligned)
constexpr uint32_t CLS_CAT_WORDS = ((TSM_CAT_SLOTS + 2 * CLS_CAT_OFF_N + TSM_CAT_BLOB_LEN + 15) / 16) * 4;   // (keeps the queue 16-byte aligned)
constexpr uint32_t CLS_SMEM_BASE = 4 * (512 + CLS_CAT_WORDS + 4 * 1024 + 4);   // bytes in front of the histogram (tables + BQ_CAP queue)
constexpr uint32_t E_FIRST = (1u << 0) | (1u << 5) | (1u << 9) | (1u << 17) | (1u << 21) | (1u << 23) | (1u << 25) |
                             (1u << 27) | (1u << 29) | (1u << 30);

struct FileBytes {                                       // 8-byte buffered reader over one file in HBM
  const unsigned long long* base; uint32_t cur_blk; unsigned long long w;
  // the first words of the candidate line, fetched by independent loads up front: one HBM / L2 latency per
  // candidate instead of one per step of the parse
  uint32_t blk0; unsigned long long w0, w1, w2, w3;
  __device__ __forceinline__ FileBytes(const uint8_t* b, uint32_t first_byte)
      : base(reinterpret_cast<const unsigned long long*>(b)), cur_blk(0xFFFFFFFFu), w(0), blk0(first_byte >> 3) {
    w0 = __ldg(base + blk0); w1 = __ldg(base + blk0 + 1); w2 = __ldg(base + blk0 + 2); w3 = __ldg(base + blk0 + 3);
  }
  __device__ __forceinline__ unsigned long long word(uint32_t blk) const {
    const uint32_t d = blk - blk0;
    if (d < 4u) return d < 2u ? (d == 0u ? w0 : w1) : (d == 2u ? w2 : w3);
    return __ldg(base + blk);
  }
  __device__ __forceinline__ uint32_t get(uint32_t i) {
    const uint32_t blk = i >> 3;
    if (blk != cur_blk) { w = word(blk); cur_blk = blk; }
    return (uint32_t)(w >> (8u * (i & 7u))) & 0xFFu;
  }
  // the 8 bytes at file offset i (unaligned), little-endian
  __device__ __forceinline__ unsigned long long get8(uint32_t i) const {
    const uint32_t blk = i >> 3, sh = 8u * (i & 7u);
    const unsigned long long lo = word(blk);
    if (sh == 0) return lo;
    return (lo >> sh) | (word(blk + 1) << (64u - sh));
  }
};

__device__ __forceinline__ unsigned long long low_bytes(unsigned long long v, uint32_t n) {   // keep n <= 8 bytes
  return n >= 8 ? v : (v & ((1ull << (8u * n)) - 1ull));
}

// gtest stem table (SPEC section 6 rule 1): stem = bytes after "EXPECT_" / "ASSERT_", n = its length
__device__ __forceinline__ int stem_lookup(FileBytes& rd, uint32_t s, uint32_t n) {
  if (n < 2 || n > 9) return 0;
  const unsigned long long v = low_bytes(rd.get8(s), n);
  const uint32_t c9 = n == 9 ? rd.get(s + 8) : 0u;
  switch (n) {
    case 2:
      if (v == 0x5145ull) return 1;            // EQ
      if (v == 0x454Eull) return 2;            // NE
      if (v == 0x5447ull) return 5;            // GT
      if (v == 0x4547ull) return 6;            // GE
      if (v == 0x544Cull) return 7;            // LT
      if (v == 0x454Cull) return 8;            // LE
      return 0;
    case 4:
      if (v == 0x45555254ull) return 3;        // TRUE
      if (v == 0x5241454Eull) return 9;        // NEAR
      return 0;
    case 5:
      if (v == 0x45534C4146ull) return 4;      // FALSE
      if (v == 0x574F524854ull) return 12;     // THROW
      return 0;
    case 8: return v == 0x51455F54414F4C46ull ? 10 : 0;                 // FLOAT_EQ
    case 9: return (v == 0x455F454C42554F44ull && c9 == 'Q') ? 11 : 0;   // DOUBLE_EQ
    default: return 0;
  }
}

// SPEC section 6 rule 2 on e = T[7:] of a bare "assert <expr>": one Shift-And pass over e for the ten
// operator patterns (table built in tsm_api.cu), 8 bytes per step:
//  " not " 0-4 | " in " 5-8 | " is not " 9-16 | "True" 17-20 | "==" 21-22 | "!=" 23-24 | "<=" 25-26 |
//  ">=" 27-28 | "<" 29 | ">" 30   (final bits 4, 8, 16, 20, 22, 24, 26, 28, 29, 30)
__device__ __forceinline__ int bare_assert_category(FileBytes& rd, uint32_t e0, uint32_t en, const uint32_t* elut) {
  if (en >= 4 && (uint32_t)rd.get8(e0) == 0x20746F6Eu) return 2;        // "not "
  uint32_t D = 0, A = 0;
  for (uint32_t a = 0; a < en; a += 8) {                 // bytes behind the end become zeros (match nothing)
    unsigned long long w = rd.get8(e0 + a);
    if (en - a < 8) w &= (1ull << (8u * (en - a))) - 1ull;
    const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
#pragma unroll
    for (int k = 0; k < 4; ++k) { D = ((D + D) | E_FIRST) & elut[__byte_perm(lo, 0, 0x4440 + k)]; A |= D; }
#pragma unroll
    for (int k = 0; k < 4; ++k) { D = ((D + D) | E_FIRST) & elut[__byte_perm(hi, 0, 0x4440 + k)]; A |= D; }
  }
  if (((A & (1u << 4)) && (A & (1u << 8))) || (A & (1u << 16))) return 4;   // not ... in / is not
  if (A & (1u << 20)) return 3;                          // True
  if (A & (1u << 22)) return 1;                          // ==
  if (A & (1u << 24)) return 2;                          // !=
  if (A & (1u << 26)) return 8;                          // <=
  if (A & (1u << 28)) return 6;                          // >=
  if (A & (1u << 29)) return 7;                          // <
  if (A & (1u << 30)) return 5;                          // >
  return 3;
}

constexpr uint32_t BQ_CAP = 1024;                        // bare-assert expressions a block of k_classify defers (16 B each)

__global__ void __launch_bounds__(256) k_classify(ScanParams p) {
  extern __shared__ __align__(16) uint32_t csm[];        // byte classes, operator table, category tables, deferral queue, [n_groups][K] histogram
  uint32_t* cls = csm;
  uint32_t* elut = csm + 256;
  uint8_t* cat_slot = reinterpret_cast<uint8_t*>(csm + 512);             // copies of the constant tables: the lanes of a
  uint16_t* cat_off = reinterpret_cast<uint16_t*>(cat_slot + TSM_CAT_SLOTS);   // warp look up different names (divergent
  uint8_t* cat_blob = reinterpret_cast<uint8_t*>(cat_off + CLS_CAT_OFF_N);     // constant-memory reads would serialise)
  uint4* bq = reinterpret_cast<uint4*>(csm + 512 + CLS_CAT_WORDS);       // deferred bare asserts: file, e0, en, event slot
  uint32_t* bqn = csm + 512 + CLS_CAT_WORDS + 4 * BQ_CAP;
  uint32_t* hist = bqn + 4;
  const bool use_smem = p.n_groups <= 16;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const uint32_t c = (uint32_t)i;
    cls[i] = (is_w(c) ? CC_W : 0u) | (is_ident(c) ? CC_IDENT : 0u) | ((c == '(' || c == '\n') ? CC_STOP : 0u);
    elut[i] = c_elut[i];
  }
  for (int i = threadIdx.x; i < TSM_CAT_SLOTS; i += blockDim.x) cat_slot[i] = c_cat_slot[i];
  for (int i = threadIdx.x; i < TSM_CAT_NAMED + 1; i += blockDim.x) cat_off[i] = c_cat_off[i];
  for (int i = threadIdx.x; i < TSM_CAT_BLOB_LEN; i += blockDim.x) cat_blob[i] = (uint8_t)c_cat_blob[i];
  if (use_smem) for (int i = threadIdx.x; i < p.n_groups * TSM_K; i += blockDim.x) hist[i] = 0;
  if (threadIdx.x == 0) *bqn = 0;
  __syncthreads();
  auto count = [&](uint32_t f, int cat) {                // cross-file aggregate: [group][category]
    const uint32_t g = p.grp ? p.grp[f] : 0u;
    if (use_smem) atomicAdd(&hist[g * TSM_K + cat], 1u);
    else {
      atomicAdd(&p.counts[(size_t)g * TSM_K + cat], 1ull);
      atomicAdd(&p.counts[(size_t)p.n_groups * TSM_K + cat], 1ull);
    }
  };
  const uint32_t n = min(p.ctrl->n_cand, p.cand_cap);
  const bool want_ev = (p.flags & TSM_SCAN_ASSERT_EVENTS) != 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long cd = p.cand[i];
    const uint32_t f = (uint32_t)(cd >> 32), line_off = (uint32_t)cd;
    const uint32_t size = (uint32_t)p.len[f];
    FileBytes rd(p.arena + (size_t)(uint32_t)p.off[f], line_off);
    // ---- T = [t0, last): skip the indentation, find the first '(' / LF a word at a time (SWAR),
    //      strip blanks backwards; L = the identifier run that ends at `last`
    uint32_t q = line_off;
    while (q < size && (cls[rd.get(q)] & CC_W)) ++q;     // LF is not blank: stops at the line end too
    const uint32_t t0 = q;
    uint32_t stop = size;
    for (uint32_t wb = t0 & ~7u; wb < size; wb += 8) {
      unsigned long long w = rd.word(wb >> 3);
      const unsigned long long x1 = w ^ 0x2828282828282828ull, x2 = w ^ 0x0A0A0A0A0A0A0A0Aull;
      const unsigned long long k7 = 0x7F7F7F7F7F7F7F7Full;
      unsigned long long z = (~(((x1 & k7) + k7) | x1 | k7)) | (~(((x2 & k7) + k7) | x2 | k7));   // 0x80 per '(' or LF
      if (wb < t0) z &= ~0ull << (8u * (t0 - wb));
      if (z) { stop = wb + ((uint32_t)__ffsll((long long)z) - 1u) / 8u; break; }
    }
    stop = min(stop, size);
    uint32_t last = stop;
    while (last > t0 && (cls[rd.get(last - 1)] & CC_W)) --last;
    uint32_t Ls = last;
    while (Ls > t0 && (cls[rd.get(Ls - 1)] & CC_IDENT)) --Ls;
    const uint32_t tlen = last - t0, Ln = last - Ls;
    // ---- category (SPEC section 6), first match wins
    int cat = 0;
    bool done = false;
    uint32_t def_e0 = 0, def_en = 0;                     // def_en != 0: category decided by the deferred operator pass
    if (Ln >= 7) {                                       // rule 1: EXPECT_x / ASSERT_x
      const unsigned long long h7 = low_bytes(rd.get8(Ls), 7);
      if (h7 == 0x5F544345505845ull || h7 == 0x5F545245535341ull) { cat = stem_lookup(rd, Ls + 7, Ln - 7); done = true; }
    }
    if (!done && tlen >= 6) {                            // rule 2: T == "assert" or T starts with "assert "
      const unsigned long long h = rd.get8(t0);
      const bool a6 = low_bytes(h, 6) == 0x747265737361ull;
      if (a6 && tlen == 6) { cat = 3; done = true; }
      else if (a6 && tlen >= 8 && ((h >> 48) & 0xFF) == 0x20) {
        done = true;
        def_e0 = t0 + 7; def_en = tlen - 7;              // e = T[7:]: the operator pass runs later, with every lane busy
      }
    }
    if (!done && Ln >= 6) {                              // rules 3-5 on L
      const unsigned long long h = rd.get8(Ls);
      if (low_bytes(h, 6) == 0x747265737361ull) {
        if (Ln == 7 && ((h >> 48) & 0xFF) == '_') cat = 3;
        else {
          cat = TSM_CAT_OTHER;
          uint32_t hh = 0x811C9DC5u;                       // FNV-1a of L, 8 bytes per load
          for (uint32_t j = 0; j < Ln; j += 8) {
            unsigned long long w = rd.get8(Ls + j);
            const uint32_t nb = min(8u, Ln - j);
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
              if (k < nb) hh = (hh ^ ((uint32_t)w & 0xFFu)) * 0x01000193u;
              w >>= 8;
            }
          }
          const int id = cat_slot[(hh * TSM_CAT_HASH_MULT) >> 23];
          if (id && (uint32_t)(cat_off[id + 1] - cat_off[id]) == Ln) {
            const uint8_t* name = cat_blob + cat_off[id];
            bool ok = true;
            for (uint32_t j = 0; j < Ln; j += 8) {
              unsigned long long w = rd.get8(Ls + j);
              const uint32_t nb = min(8u, Ln - j);
              for (uint32_t k = 0; k < nb; ++k) { ok &= ((uint32_t)w & 0xFFu) == (uint32_t)name[j + k]; w >>= 8; }
            }
            if (ok) cat = id;
          }
        }
      }
    }
    // ---- event, then aggregate (or defer)
    uint32_t ev_slot = 0xFFFFFFFFu;
    if (want_ev) {
      unsigned long long hacc = 0; uint32_t hr = 0;     // Mersenne-61 of T (SPEC section 3)
      for (uint32_t j = 0; j < tlen; ++j) {
        hacc = fold61(hacc + rotl61((unsigned long long)rd.get(t0 + j), hr));
        hr += 8; if (hr >= 61) hr -= 61;
      }
      const uint32_t slot = atomicAdd(&p.ctrl->n_aev, 1u);
      if (slot < p.aev_cap) {
        tsm_assert_event ev;
        ev.file = f; ev.line_off = line_off; ev.stmt_off = t0;
        ev.stmt_len = (uint16_t)min(tlen, 65535u); ev.cat = (uint16_t)cat;      // (patched by the deferred pass)
        ev.ident_off = Ls; ev.ident_len = (uint16_t)min(Ln, 65535u); ev.pad = 0;
        ev.stmt_hash = mix_hash(canon61(hacc), tlen);
        p.aev[slot] = ev;
        ev_slot = slot;
      } else p.ctrl->overflow = 1;
    }
    if (def_en) {
      const uint32_t slot = atomicAdd(bqn, 1u);
      if (slot < BQ_CAP) { bq[slot] = make_uint4(f, def_e0, def_en, ev_slot); continue; }
      cat = bare_assert_category(rd, def_e0, def_en, elut);              // queue full: decide here
      if (ev_slot != 0xFFFFFFFFu) p.aev[ev_slot].cat = (uint16_t)cat;
    }
    count(f, cat);
  }
  // ---- deferred bare asserts: the block's queue, one expression per thread (the inline version kept 3 of 32
  //      lanes busy for ~50 bytes of serial automaton each)
  __syncthreads();
  {
    const uint32_t nq = min(*bqn, BQ_CAP);
    for (uint32_t t = threadIdx.x; t < nq; t += blockDim.x) {
      const uint4 e = bq[t];
      FileBytes rd(p.arena + (size_t)(uint32_t)p.off[e.x], e.y);
      const int cat = bare_assert_category(rd, e.y, e.z, elut);
      if (e.w != 0xFFFFFFFFu) p.aev[e.w].cat = (uint16_t)cat;
      count(e.x, cat);
    }
  }
  // ---- totals of the per-file records (lines, assertion lines, headers, fixture headers) behind the table
  {
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < p.n_files; f += gridDim.x * blockDim.x) {
      const tsm_file_stat s = p.stats[f];
      t0 += s.n_lines; t1 += s.n_assert; t2 += s.n_headers; t3 += s.n_fixture;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      t0 += __shfl_xor_sync(0xffffffffu, t0, d); t1 += __shfl_xor_sync(0xffffffffu, t1, d);
      t2 += __shfl_xor_sync(0xffffffffu, t2, d); t3 += __shfl_xor_sync(0xffffffffu, t3, d);
    }
    if ((threadIdx.x & 31) == 0) {
      unsigned long long* tot = p.counts + (size_t)(p.n_groups + 1) * TSM_K;
      if (t0) atomicAdd(&tot[0], t0);
      if (t1) atomicAdd(&tot[1], t1);
      if (t2) atomicAdd(&tot[2], t2);
      if (t3) atomicAdd(&tot[3], t3);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < p.n_groups * TSM_K; i += blockDim.x) {
      const uint32_t v = hist[i];
      if (v) {
        atomicAdd(&p.counts[i], (unsigned long long)v);
        atomicAdd(&p.counts[(size_t)p.n_groups * TSM_K + (i & (TSM_K - 1))], (unsigned long long)v);
      }
    }
  }
}

}  // namespace tsm
