// tsm_scan2_kernels.cuh - k_scan, second generation (round 2): the "streaming stripe walk".
//
// Same contract as the first k_scan (docs/SPEC.md sections 2-5, 7): one warp per
// (file, 4 KiB chunk) work unit, chunk staged global -> shared by one 1-D TMA bulk copy, every source byte
// read from HBM exactly once.  What changed is how the per-line facts are produced:
//
//   * '\n' is a ninth pattern of the Shift-And automaton (state bit 31), so the OR of a word's eight
//     states says for free whether the word holds a newline: no SWAR newline pass, no line table up front;
//   * the walk keeps the OR of the states since the last newline word in a REGISTER and stores it behind
//     every word (one STS): no per-word shared-memory atomicOr, no per-word line index arithmetic;
//   * lines are finished per NEWLINE WORD (one lane per word that holds a newline, balanced over the warp
//     through a small entry table): the line that ends at the word's first newline gets the stored OR,
//     lines that lie inside the word (at most 6 bytes) are walked in place;
//   * a word that holds a newline AND a pattern end ("mixed", a few per chunk) is re-walked byte by byte
//     by a dense pass that splits its states between the line that ends in it and the line that starts.
//
// The hash prefix machinery (Mersenne-61 running prefix, checkpoints, warp scan of the stripe totals) is
// the first generation's.  There is no reference kernel; rules cite docs/SPEC.md.
#pragma once
#include "tsm_scan_kernels.cuh"

namespace tsm {

__constant__ uint32_t c_lut2[256];                       // automaton table of this kernel (bit 31 = '\n')

// Pattern layout: the first generation's (tsm_device.cuh) without the `_F` gate, plus the newline bit.
constexpr uint32_t B_FIRST = (1u << 0) | (1u << 6) | (1u << 13) | (1u << 18) | (1u << 21) | (1u << 25) | (1u << 29) | (1u << 31);

// ---- per-warp shared memory --------------------------------------------------------------------
constexpr uint32_t NWORD = BUF / 8;                      // 544 words of 8 bytes, 17 per stripe
constexpr uint32_t O2_ARUN = BUF;                        // u32[NWORD + 1]  OR of the states since the last newline word, in front of every word
constexpr uint32_t SLOT_TAIL = NWORD;                    //                 (+ one slot: the line that ends with the data)
constexpr uint32_t O2_RW = O2_ARUN + ((NWORD + 1) * 4 + 7) / 8 * 8;   // u64[32 * 4]   hash prefix behind words 3, 7, 11, 15 of every stripe
constexpr uint32_t O2_ENT = O2_RW + 32 * 4 * 8;          // u16[ENT_CAP]    newline words in order (bit 15: mixed); later the candidate list
constexpr uint32_t ENT_CAP = NWORD + 8;
constexpr uint32_t ENT_VIRTUAL = 0x7FFFu;                // entry of the unterminated last line of a file
constexpr uint32_t O2_BASE = O2_ENT + ENT_CAP * 2;       // u64[33]         hash prefix at every stripe start (+ total)
constexpr uint32_t Q2_CAP = 64;
constexpr uint32_t O2_Q = O2_BASE + 34 * 8;              // u16[Q2_CAP]     mixed words
constexpr uint32_t O2_CTL = O2_Q + Q2_CAP * 2;           // u32 queue length, pad, u64 mbarrier
constexpr uint32_t WARP_SMEM2 = ((O2_CTL + 16 + 127) / 128) * 128;
constexpr uint32_t SCAN2_SMEM = LUT_BYTES + SCAN_WARPS * WARP_SMEM2;
static_assert(O2_RW % 8 == 0 && O2_ENT % 8 == 0 && O2_BASE % 8 == 0 && O2_Q % 4 == 0 && O2_CTL % 8 == 0, "alignment");

// Eight automaton steps over one 8-byte word; A collects every state of the word.
__device__ __forceinline__ void step8b(unsigned long long w, uint32_t& D, uint32_t& A) {
  const uint32_t* lut = scan_lut();
  const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    D = ((D + D) | B_FIRST) & lut[__byte_perm(lo, 0, 0x4440 + k)];
    A |= D;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    D = ((D + D) | B_FIRST) & lut[__byte_perm(hi, 0, 0x4440 + k)];
    A |= D;
  }
}

__device__ __forceinline__ uint32_t nl8_of(unsigned long long w) {       // bit b = byte b of w is '\n'
  return nl_word((uint32_t)w) | (nl_word((uint32_t)(w >> 32)) << 4);
}
__device__ __forceinline__ unsigned long long low_mask(uint32_t b) {    // the b low bytes, b in 0..8
  return b >= 8u ? ~0ull : ((1ull << (8u * b)) - 1ull);
}
__device__ __forceinline__ unsigned long long ror3_61(unsigned long long r) { return (r >> 3) + ((r & 7ull) << 58); }

struct WalkOut { uint32_t nlw, tail; };                  // bit k: word k of the stripe holds a newline; states since the stripe's last newline word

// The walk: every lane takes the 17 words of its own 136-byte stripe (all 32 lanes busy whatever the line
// lengths are; the bytes outside the chunk's staged range are zeros).  Per word: 8 automaton steps, the
// Mersenne-61 running prefix R_k = R_{k-1} * 2^-64 + w_k, one store of the running OR.  Then one warp scan
// turns the stripe totals into the absolute hash prefix at every stripe start.
__device__ __noinline__ WalkOut walk2(uint8_t* wb, uint32_t fin, int lane) {
  const uint32_t pos0 = (uint32_t)lane * STRIPE;
  const uint8_t* sp = wb + pos0;
  uint32_t* ar = reinterpret_cast<uint32_t*>(wb + O2_ARUN) + (uint32_t)lane * 17u;
  unsigned long long* rw = reinterpret_cast<unsigned long long*>(wb + O2_RW) + (uint32_t)lane * 4u;
  uint32_t D = 0;
  if (lane) {                                            // state in front of the stripe (no pattern is longer than 7 bytes)
    uint32_t A = 0;
    step8b(*reinterpret_cast<const unsigned long long*>(sp - 8), D, A);
  }
  unsigned long long R = 0;
  uint32_t run = 0, nlr = 0;                             // nlr: newline-word bits, the newest word in bit 0
#pragma unroll 1
  for (uint32_t g = 0; g < 5; ++g) {
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {
      if (g == 4 && k) break;
      const unsigned long long w = *reinterpret_cast<const unsigned long long*>(sp + 32u * g + 8u * k);
      uint32_t A = 0;
      step8b(w, D, A);
      R = ror3_61(R) + fold61(w);                        // lazily reduced: stays below 2^63
      if (k == 3) rw[g] = R;
      ar[4u * g + k] = run;
      nlr = __funnelshift_l(A, nlr, 1);                  // bit 31 of A: the word holds a newline
      const bool nl = (int32_t)A < 0;
      if (nl && (A & fin)) {                             // rare: a pattern ends in a newline word
        const uint32_t slot = atomicAdd(reinterpret_cast<uint32_t*>(wb + O2_CTL), 1u);
        if (slot < Q2_CAP) reinterpret_cast<uint16_t*>(wb + O2_Q)[slot] = (uint16_t)((uint32_t)lane * 17u + 4u * g + k);
      }
      run = nl ? 0u : (run | A);
    }
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
  unsigned long long* sbase = reinterpret_cast<unsigned long long*>(wb + O2_BASE);
  sbase[lane] = excl;
  if (lane == 31) sbase[32] = incl;
  return WalkOut{__brev(nlr) >> 15, run};
}

// A mixed word (newline + pattern end), byte by byte: the states in front of its first newline belong to the
// line that ends there (entry i), the states behind its last newline to the line that ends at the next entry.
// Lines inside the word are walked by the finish pass itself (the entry's bit 15 asks for it).
__device__ __forceinline__ void resolve_mixed(uint8_t* wb, uint32_t g, uint32_t i, uint32_t n_real) {
  const uint32_t* lut = scan_lut();
  uint32_t D = 0, A = 0;
  step8b(*reinterpret_cast<const unsigned long long*>(wb + 8u * g - 8u), D, A);   // g >= 2: the first 16 bytes are zeros
  unsigned long long w = *reinterpret_cast<const unsigned long long*>(wb + 8u * g);
  uint32_t pre = 0, post = 0, seen = 0;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    D = ((D + D) | B_FIRST) & lut[(uint32_t)w & 0xFFu];
    const uint32_t m = (uint32_t)((int32_t)D >> 31);     // all ones at a newline
    pre |= D & ~seen;
    seen |= m;
    post = (post | D) & ~m;
    w >>= 8;
  }
  uint32_t* arun = reinterpret_cast<uint32_t*>(wb + O2_ARUN);
  const uint16_t* ent = reinterpret_cast<const uint16_t*>(wb + O2_ENT);
  atomicOr(arun + g, pre);
  const uint32_t tgt = i + 1u < n_real ? ((uint32_t)ent[i + 1u] & 0x3FFu) : SLOT_TAIL;
  atomicOr(arun + tgt, post);
}

// Hash prefixes (SPEC section 3) of the staged bytes [0, 8g + b1) and [0, 8g + b2), every byte weighted
// 256^position, lazily reduced (< 2^62 + 2).  w = word g.
__device__ __forceinline__ void prefix_pair(const uint8_t* wb, uint32_t g, unsigned long long w, uint32_t b1, uint32_t b2,
                                            unsigned long long& P1, unsigned long long& P2) {
  const uint32_t l = g / 17u, i = g - 17u * l, c0 = i >> 2, ns = i & 3u;
  unsigned long long R = 0;
  if (c0) R = *reinterpret_cast<const unsigned long long*>(wb + O2_RW + 8u * (l * 4u + c0 - 1u));
  const unsigned long long* wp = reinterpret_cast<const unsigned long long*>(wb) + (g - ns);   // words since the checkpoint
  if (ns > 0u) R = ror3_61(R) + fold61(wp[0]);
  if (ns > 1u) R = ror3_61(R) + fold61(wp[1]);
  if (ns > 2u) R = ror3_61(R) + fold61(wp[2]);
  const unsigned long long Q = ror3_61(R);               // frame of word g
  const uint32_t r3g = (3u * g) % 61u;
  const unsigned long long sb = *reinterpret_cast<const unsigned long long*>(wb + O2_BASE + 8u * l);
  P1 = sb + rotl61(fold61(fold61(Q + fold61(w & low_mask(b1)))), r3g);
  P2 = sb + rotl61(fold61(fold61(Q + fold61(w & low_mask(b2)))), r3g);
}

__device__ __forceinline__ void emit_header(const ScanParams& p, uint32_t f, uint32_t line_off, uint32_t len, uint32_t fl) {
  const uint32_t slot = atomicAdd(&p.ctrl->n_hev, 1u);
  if (slot < p.hev_cap) p.hev[slot] = tsm_header_event{f, line_off, len, (fl >> 2) & 1u};
  else p.ctrl->overflow = 1;
}

// Finish pass: one lane per newline word (entry).  Line "A" of an entry ends at the word's first newline and
// starts behind the last newline of the entry in front of it; its hash is the difference of two prefixes, its
// pattern flags are the word's stored OR.  The other newlines of the word end lines of at most 6 bytes that lie
// inside the word.  A line belongs to the chunk its first byte lies in (start < lim).  The starts of the
// assertion lines are compacted (u16 each) over the entries already consumed.  Returns their number.
__device__ __noinline__ uint32_t finish_entries(const ScanParams& p, uint8_t* wb, const uint32_t* lc, uint32_t n_real, uint32_t n_tot,
                                                uint32_t X, uint32_t lim, bool skip_first, uint32_t f, uint32_t cb, int ext,
                                                int lane, Accum& ac) {
  uint16_t* ent = reinterpret_cast<uint16_t*>(wb + O2_ENT);
  const uint32_t* arun = reinterpret_cast<const uint32_t*>(wb + O2_ARUN);
  const uint32_t g1 = lc[1], g2 = lc[2];
  const SmemByte lb{wb};
  const bool want_hev = (p.flags & TSM_SCAN_HEADER_EVENTS) != 0;
  Accum a = ac;
  uint32_t nc = 0;
  uint32_t prev_last = PRE - 1u;                         // newline in front of the next line (the 16 bytes in front of the chunk are zeros)
  unsigned long long prevP = 0;                          // hash prefix of the bytes [0, prev_last]
  for (uint32_t base = 0; base < n_tot; base += 32) {    // uniform trip count
    const uint32_t j = base + (uint32_t)lane;
    const bool valid = j < n_tot;
    const uint32_t ge = valid ? (uint32_t)ent[j] : 0u;
    const bool isv = valid && j >= n_real;               // the unterminated last line of the file: ends at X
    const uint32_t g = isv ? (X - 1u) >> 3 : (ge & 0x3FFu);
    const unsigned long long w = *reinterpret_cast<const unsigned long long*>(wb + 8u * g);
    const uint32_t nl8 = isv ? (1u << (X - 8u * g)) : (valid ? nl8_of(w) : 1u);
    const uint32_t p1 = (uint32_t)__ffs((int)nl8) - 1u, p2 = 31u - (uint32_t)__clz((int)nl8);
    const uint32_t e = 8u * g + p1, eL = 8u * g + p2;
    unsigned long long Pe, Pn;
    prefix_pair(wb, g, w, p1, p2 + 1u, Pe, Pn);
    if (isv) Pe = *reinterpret_cast<const unsigned long long*>(wb + O2_BASE + 8u * 32u);   // everything staged (zeros behind X)
    uint32_t s = __shfl_up_sync(0xffffffffu, eL, 1) + 1u;
    unsigned long long Ps = __shfl_up_sync(0xffffffffu, Pn, 1);
    if (lane == 0) { s = prev_last + 1u; Ps = prevP; }
    prev_last = __shfl_sync(0xffffffffu, eL, 31);
    prevP = __shfl_sync(0xffffffffu, Pn, 31);
    const bool owned = valid && s < lim && !(s == PRE && skip_first);
    uint32_t fl = 0;
    if (owned) {
      const uint32_t A = arun[isv ? SLOT_TAIL : g];
      const unsigned long long hr = canon61(Pe + 4ull * M61 - Ps);       // bytes [s, e), weighted from position 0
      const uint32_t sh = (8u * s) % 61u;
      const unsigned long long h0 = rotl61(hr, sh ? 61u - sh : 0u);
      fl = line_finish_h(s, e, h0, flag_nibble(A, g1, g2, 0u), ext, lb, a);
      if (want_hev && (fl & LF_HDR)) emit_header(p, f, cb + s - PRE, e - s, fl);
    }
    __syncwarp();                                        // every entry of the round is read: the list may grow over them
    const uint32_t mc = __ballot_sync(0xffffffffu, fl & LF_CAND);
    if (fl & LF_CAND) ent[nc + __popc(mc & ((1u << lane) - 1u))] = (uint16_t)s;
    nc += __popc(mc);
    // ---- lines inside the word (behind its first newline): content of at most 6 bytes
    uint32_t rest = (valid && !isv) ? (nl8 & (nl8 - 1u)) : 0u;
    while (__any_sync(0xffffffffu, rest != 0u)) {
      if (rest) {
        const uint32_t q2 = (uint32_t)__ffs((int)rest) - 1u;
        const uint32_t q1 = 31u - (uint32_t)__clz((int)(nl8 & ((1u << q2) - 1u)));
        const uint32_t si = 8u * g + q1 + 1u, len = q2 - q1 - 1u;
        if (si < lim) {
          const unsigned long long v = (w >> (8u * (q1 + 1u))) & low_mask(len);   // < 2^48: canonical as it is
          uint32_t A2 = 0;
          if ((ge & 0x8000u) && len) {                   // mixed word: the line's own states
            const uint32_t* lut = scan_lut();
            uint32_t D = 0;
            unsigned long long t = v;
            for (uint32_t c = 0; c < len; ++c) { D = ((D + D) | B_FIRST) & lut[(uint32_t)t & 0xFFu]; A2 |= D; t >>= 8; }
          }
          const uint32_t fl2 = line_finish_h(si, si + len, v, flag_nibble(A2, g1, g2, 0u), ext, lb, a);
          if (fl2 & LF_CAND) {
            const uint32_t slot = atomicAdd(&p.ctrl->n_cand, 1u);
            if (slot < p.cand_cap) p.cand[slot] = ((unsigned long long)f << 32) | (cb + si - PRE);
            else p.ctrl->overflow = 1;
          }
          if (want_hev && (fl2 & LF_HDR)) emit_header(p, f, cb + si - PRE, len, fl2);
        }
        rest &= rest - 1u;
      }
    }
    __syncwarp();
  }
  ac = a;
  return nc;
}

__device__ __forceinline__ void process_chunk2(const ScanParams& p, const uint32_t* lc, uint8_t* wb, uint32_t f,
                                               uint32_t cb, uint32_t fo, uint32_t size, int ext, int lane) {
  const uint32_t ce = min(cb + CH, size);
  const uint32_t le = min(ce + EXT, size);
  const uint32_t lim = PRE + (ce - cb);                  // buffer position just past the owned bytes
  const uint32_t lim2 = PRE + (le - cb);                 // ... past the staged bytes
  const bool skip_first = (cb != 0) && (wb[PRE - 1] != '\n');   // chunk starts inside a foreign line
  Accum ac{0, 0, 0, 0, 0};
  __syncwarp();
  // ---- everything outside the staged range [PRE, lim2) becomes zeros: no pass has to mask its loads
  //      (a zero byte is no newline, matches no pattern and adds nothing to the hash prefix)
  if (lane < 2) reinterpret_cast<unsigned long long*>(wb)[lane] = 0ull;
  {
    const uint32_t za = (lim2 + 7u) & ~7u;
    if ((uint32_t)lane < za - lim2) wb[lim2 + lane] = 0;
    for (uint32_t q = za + 8u * (uint32_t)lane; q < BUF; q += 256u) *reinterpret_cast<unsigned long long*>(wb + q) = 0ull;
  }
  if (lane == 0) *reinterpret_cast<uint32_t*>(wb + O2_CTL) = 0u;
  __syncwarp();
  const WalkOut wo = walk2(wb, lc[0], lane);
  // ---- newline words behind the owned bytes: only the first one matters (it ends the last owned line)
  const uint32_t w0 = 17u * (uint32_t)lane, lim_w = (lim + 7u) >> 3;
  const uint32_t ownbits = lim_w <= w0 ? 0u : (lim_w - w0 >= 17u ? 0x1FFFFu : (1u << (lim_w - w0)) - 1u);
  const uint32_t extbits = wo.nlw & ~ownbits;
  const uint32_t gx = __reduce_min_sync(0xffffffffu, extbits ? w0 + (uint32_t)__ffs((int)extbits) - 1u : 0xFFFFu);
  uint32_t kept = wo.nlw & ownbits;
  if (gx - w0 < 17u) kept |= 1u << (gx - w0);
  // ---- states of a line that spans stripes: OR of the stripe tails back to the stripe of its first byte
  uint32_t tv = wo.tail, tf = wo.nlw != 0u;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t uv = __shfl_up_sync(0xffffffffu, tv, d), uf = __shfl_up_sync(0xffffffffu, tf, d);
    if (lane >= d) { if (!tf) tv |= uv; tf |= uf; }
  }
  uint32_t carry = __shfl_up_sync(0xffffffffu, tv, 1);
  if (lane == 0) carry = 0;
  const uint32_t tail_all = __shfl_sync(0xffffffffu, tv, 31);
  // ---- entry table: the kept newline words in order
  const uint32_t cnt = __popc(kept);
  uint32_t incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  const uint32_t n_real = __shfl_sync(0xffffffffu, incl, 31), ebase = incl - cnt;
  uint16_t* ent = reinterpret_cast<uint16_t*>(wb + O2_ENT);
  uint32_t* arun = reinterpret_cast<uint32_t*>(wb + O2_ARUN);
  {
    uint32_t b = kept, idx = ebase;
    while (b) { ent[idx++] = (uint16_t)(w0 + (uint32_t)__ffs((int)b) - 1u); b &= b - 1u; }
    if (wo.nlw) arun[w0 + (uint32_t)__ffs((int)wo.nlw) - 1u] |= carry;
    if (lane == 0) arun[SLOT_TAIL] = tail_all;
  }
  __syncwarp();
  // ---- mixed words
  {
    const uint32_t nq_all = *reinterpret_cast<const uint32_t*>(wb + O2_CTL);
    const uint16_t* q = reinterpret_cast<const uint16_t*>(wb + O2_Q);
    if (nq_all <= Q2_CAP) {
      for (uint32_t base = 0; base < nq_all; base += 32) {
        const uint32_t t = base + (uint32_t)lane;
        const uint32_t g = t < nq_all ? (uint32_t)q[t] : 0u;
        const uint32_t l = g / 17u, k = g - 17u * l;
        const uint32_t m = __shfl_sync(0xffffffffu, kept, (int)l), eb = __shfl_sync(0xffffffffu, ebase, (int)l);
        const bool act = t < nq_all && ((m >> k) & 1u);
        const uint32_t i = eb + __popc(m & ((1u << k) - 1u));
        if (act) resolve_mixed(wb, g, i, n_real);
        __syncwarp();
        if (act) ent[i] |= 0x8000u;
        __syncwarp();
      }
    } else if (lc[0]) {                                  // queue overflow: take every newline word as mixed
      for (uint32_t base = 0; base < n_real; base += 32) {
        const uint32_t i = base + (uint32_t)lane;
        if (i < n_real) resolve_mixed(wb, (uint32_t)ent[i] & 0x3FFu, i, n_real);
        __syncwarp();
        if (i < n_real) ent[i] |= 0x8000u;
        __syncwarp();
      }
    }
  }
  // ---- the line behind the last newline: ends with the file (virtual entry), lies in the next chunk, or is long
  uint32_t final_prev = PRE - 1u;
  if (n_real) {
    const uint32_t gl = (uint32_t)ent[n_real - 1u] & 0x3FFu;
    const uint32_t m = nl8_of(*reinterpret_cast<const unsigned long long*>(wb + 8u * gl));
    final_prev = 8u * gl + 31u - (uint32_t)__clz((int)m);
  }
  const uint32_t tail_start = final_prev + 1u;
  bool virt = false, tail_long = false;
  if (tail_start < lim && !(skip_first && n_real == 0u)) {
    if (le == size) virt = tail_start < lim2;            // unterminated last line of the file (the data ends at lim2)
    else tail_long = true;
  }
  const uint32_t nc = finish_entries(p, wb, lc, n_real, n_real + (virt ? 1u : 0u), lim2, lim, skip_first, f, cb, ext, lane, ac);
  // ---- candidates to their global list
  if (nc) {
    const uint32_t cbase = warp_reserve(&p.ctrl->n_cand, nc, lane);
    for (uint32_t i = (uint32_t)lane; i < nc; i += 32) {
      const uint32_t slot = cbase + i;
      if (slot < p.cand_cap) p.cand[slot] = ((unsigned long long)f << 32) | (cb + (uint32_t)ent[i] - PRE);
      else p.ctrl->overflow = 1;
    }
  }
  if (tail_long && lane == 0) long_line(p, scan_lut(), B_FIRST, f, fo, size, ext, cb + tail_start - PRE, ac);
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

__global__ void __launch_bounds__(SCAN_WARPS * 32, SCAN_CTAS_PER_SM) k_scan2(ScanParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t* lut_all = reinterpret_cast<uint32_t*>(smem);  // [0,256) the automaton table, then 3 x 4 per-language masks
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut_all[i] = c_lut2[i];
  if (threadIdx.x < 12) {                                // per language (PY, C family, none): pattern ends that count, then the header groups
    const int t = threadIdx.x, lang = t >> 2, q = t & 3;
    const uint32_t g1 = lang == 0 ? PY_G1 : CJ_G1, g2 = lang == 0 ? PY_G2 : CJ_G2;
    const uint32_t v = q == 0 ? (AF_ASSERT | AF_EXPECT | g1 | g2) : (q == 1 ? g1 : (q == 2 ? g2 : 0u));
    lut_all[256 + t] = lang == 2 ? 0u : v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wb = smem + LUT_BYTES + warp * WARP_SMEM2;
  uint64_t* bar = reinterpret_cast<uint64_t*>(wb + O2_CTL + 8);
  if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncwarp();
  const uint32_t n_units = p.slab->n_units;
  uint32_t phase = 0;
  Unit cur = claim_unit(p, n_units, lane);
#if TSM_LOCKSTEP
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
    process_chunk2(p, lut_all + 256u + 4u * lang, wb, cur.f, cur.cb, cur.fo, cur.size, cur.ext, lane);
    __syncwarp();
    cur = nxt;
  }
}

}  // namespace tsm
