// Kernel definitions; see kernels.cuh for the inventory and engine.cu for launches.
#include "kernels.cuh"
#include "launch.h"

namespace rbgpu {

extern __shared__ __align__(16) unsigned char g_smem[];

__device__ __forceinline__ uint32_t pick_start_rev(const DfaView& d, const uint8_t* t, uint64_t n, uint64_t at) {
  return d.uniform_start ? d.start[32] : d.start[flags_reverse(t, n, at)];
}
__device__ __forceinline__ uint32_t pick_start_fwd(const DfaView& d, const uint8_t* t, uint64_t n, uint64_t at) {
  return d.uniform_start ? d.start[32] : d.start[flags_forward(t, n, at)];
}

// ------------------------------------------------------------ scan_rev_bitmap --
// Bitmap format: bit i is set iff a match STARTS at position i+1, i.e. iff the
// reverse automaton reports a (delayed) match when it consumes text[i].  A match
// starting at position 0 is reported by the EOF step into *flag0.
// Segment t owns bytes [base + t*seg, base + (t+1)*seg) ∩ [0, n).  The state at
// the top of a segment is guessed by running the automaton from a fresh start
// state `warm` bytes further right; verify_segments() checks every guess
// against the neighbour's exact final state and lists the segments to redo.
//
// Generic version: any table size (shared memory when it fits, else L1/L2), any
// alignment, optional UTF-8 boundary mask.  One byte load + two lookups per byte.
__global__ void scan_rev_bitmap(ScanArgs a) {
  const Table T = stage_table(a.dfa, g_smem, a.use_smem);
  const uint64_t total = a.redo_list ? (uint64_t)*a.n_redo : a.n_seg;
  const uint32_t match_lo = a.dfa.match_lo;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t t = a.redo_list ? a.redo_list[idx] : idx;
    const uint64_t lo = min(a.base + t * a.seg, a.n);
    const uint64_t hi = min(lo + a.seg, a.n);  // bytes [lo, hi)
    uint32_t s;
    if (a.redo_list) {
      s = a.fin[t + 1];
    } else {
      const uint64_t w = min(hi + a.warm, a.n);
      s = pick_start_rev(a.dfa, a.text, a.n, w);
      for (uint64_t i = w; i > hi; i--) s = T.step(s, a.text[i - 1]);
    }
    a.guess[t] = (uint16_t)s;
    uint64_t word = 0;
    uint32_t next_byte = hi < a.n ? a.text[hi] : 0;  // text[i+1], for the UTF-8 boundary mask
    for (uint64_t i = hi; i > lo;) {
      i--;
      const uint32_t b = a.text[i];
      s = T.step(s, b);
      bool hit = s >= match_lo;
      if (a.utf8_boundaries && (next_byte & 0xC0) == 0x80) hit = false;
      if (hit) word |= 1ull << (i & 63);
      next_byte = b;
      if ((i & 63) == 0 || i == lo) { a.bitmap[i >> 6] = word; word = 0; }
    }
    a.fin[t] = (uint16_t)s;
    if (lo == 0) {
      bool hit0 = T.step_eof(s) >= match_lo;
      if (a.utf8_boundaries && (next_byte & 0xC0) == 0x80 && a.n > 0) hit0 = false;
      *a.flag0 = hit0;
    }
  }
}

// Fast version for byte-indexed tables that fit shared memory (<= ~200 states):
//   - the table is expanded to [state][256] 32-bit entries; an entry IS the shared-
//     memory address of the successor's row plus a per-row XOR key, so one LOP3
//     ((w >> k) & 0x3FC) ^ entry forms the next address and one LDS fetches the
//     next entry.  The key (row & 31) rotates the bank of a given byte from row
//     to row, which removes the systematic conflicts of small alphabets (DNA).
//   - match rows sit above non-match rows, so "entry >= thr" is the match test.
//   - each lane streams its own segment with 16-byte loads, one 64-byte group
//     prefetched ahead; bits are assembled in registers, one 64-bit store per
//     64 bytes of text.
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 ldg128(const uint8_t* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// Consume the 4 bytes of w from the highest address down; BIT0 = bit index of byte 0.
template <int BIT0>
__device__ __forceinline__ void rev_word(uint32_t w, uint32_t& e, uint32_t& bits, uint32_t thr) {
  e = lds32(((w >> 22) & 0x3FCu) ^ e); if (e >= thr) bits |= 1u << (BIT0 + 3);
  e = lds32(((w >> 14) & 0x3FCu) ^ e); if (e >= thr) bits |= 1u << (BIT0 + 2);
  e = lds32(((w >> 6) & 0x3FCu) ^ e);  if (e >= thr) bits |= 1u << (BIT0 + 1);
  e = lds32(((w << 2) & 0x3FCu) ^ e);  if (e >= thr) bits |= 1u << (BIT0 + 0);
}
template <int BIT0>
__device__ __forceinline__ void rev_block16(const uint4& v, uint32_t& e, uint32_t& bits, uint32_t thr) {
  rev_word<BIT0 + 12>(v.w, e, bits, thr);
  rev_word<BIT0 + 8>(v.z, e, bits, thr);
  rev_word<BIT0 + 4>(v.y, e, bits, thr);
  rev_word<BIT0 + 0>(v.x, e, bits, thr);
}
__device__ __forceinline__ uint32_t fast_entry(uint32_t tbase, uint32_t id) { return tbase + id * 1024u + (id & 31u) * 4u; }
__device__ __forceinline__ uint32_t fast_step(uint32_t e, uint32_t byte) { return lds32((byte << 2) ^ e); }

__global__ void __launch_bounds__(1024, 1) scan_rev_fast(ScanArgs a) {
  // stage: entry(r, b) at word (r << 8) + (b ^ (r & 31))
  const uint32_t tbase = ((uint32_t)__cvta_generic_to_shared(g_smem) + 1023u) & ~1023u;
  {
    const uint32_t n_ent = a.dfa.n_states * 256u;
    for (uint32_t i = threadIdx.x; i < n_ent; i += blockDim.x) {
      const uint32_t r = i >> 8, b = i & 255u;
      const uint32_t nx = a.next256[i];
      const uint32_t addr = tbase + (r << 10) + ((b ^ (r & 31u)) << 2);
      const uint32_t val = fast_entry(tbase, nx);
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(val));
    }
    __syncthreads();
  }
  const uint32_t thr = tbase + a.dfa.match_lo * 1024u;
  const uint64_t total = a.redo_list ? (uint64_t)*a.n_redo : a.n_seg;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t t = a.redo_list ? a.redo_list[idx] : idx;
    const uint64_t lo = min(a.base + t * a.seg, a.n);  // multiple of 64 (or n)
    const uint64_t hi = min(lo + a.seg, a.n);
    uint32_t e;
    if (a.redo_list) {
      e = fast_entry(tbase, a.fin[t + 1]);
    } else {
      const uint64_t w = min(hi + a.warm, a.n);
      e = fast_entry(tbase, pick_start_rev(a.dfa, a.text, a.n, w));
      // warm-up over [hi, w): ragged top bytewise, then whole 64-byte groups
      uint64_t i = w;
      while (i > hi && (i & 63)) { i--; e = fast_step(e, a.text[i]); }
      uint32_t dummy = 0;
      while (i > hi) {
        i -= 64;
        const uint8_t* p = a.text + i;
        const uint4 v3 = ldg128(p + 48), v2 = ldg128(p + 32), v1 = ldg128(p + 16), v0 = ldg128(p);
        rev_block16<0>(v3, e, dummy, 0xFFFFFFFFu);
        rev_block16<0>(v2, e, dummy, 0xFFFFFFFFu);
        rev_block16<0>(v1, e, dummy, 0xFFFFFFFFu);
        rev_block16<0>(v0, e, dummy, 0xFFFFFFFFu);
      }
    }
    a.guess[t] = (uint16_t)((e - tbase) >> 10);
    uint64_t i = hi;
    if (i & 63) {  // ragged top (only the last segment of the haystack)
      uint64_t word = 0;
      while (i > lo && (i & 63)) {
        i--;
        e = fast_step(e, a.text[i]);
        if (e >= thr) word |= 1ull << (i & 63);
      }
      a.bitmap[i >> 6] = word;
    }
    if (i > lo) {
      const uint8_t* p = a.text + i - 64;
      uint4 v3 = ldg128(p + 48), v2 = ldg128(p + 32), v1 = ldg128(p + 16), v0 = ldg128(p);
      while (i > lo) {
        i -= 64;
        const uint4 c3 = v3, c2 = v2, c1 = v1, c0 = v0;
        if (i > lo) {  // prefetch the next (lower) group
          const uint8_t* q = a.text + i - 64;
          v3 = ldg128(q + 48); v2 = ldg128(q + 32); v1 = ldg128(q + 16); v0 = ldg128(q);
        }
        uint32_t bhi = 0, blo = 0;
        rev_block16<16>(c3, e, bhi, thr);
        rev_block16<0>(c2, e, bhi, thr);
        rev_block16<16>(c1, e, blo, thr);
        rev_block16<0>(c0, e, blo, thr);
        a.bitmap[i >> 6] = ((uint64_t)bhi << 32) | blo;
      }
    }
    const uint32_t s_lo = (e - tbase) >> 10;
    a.fin[t] = (uint16_t)s_lo;
    if (lo == 0) *a.flag0 = a.eof[s_lo] >= a.dfa.match_lo;
  }
}

// ------------------------------------------------------------ scan_fwd_reduce --
// Mirror image for forward all-match scans.  Position q is a match END iff the
// automaton, standing at q, reports a match when it consumes text[q] (EOF at n).
// Per segment: the first match end and the OR of the per-state pattern masks.
__global__ void scan_fwd_reduce(ScanArgs a) {
  const Table T = stage_table(a.dfa, g_smem, a.use_smem);
  const uint64_t total = a.redo_list ? (uint64_t)*a.n_redo : a.n_seg;
  const uint32_t match_lo = a.dfa.match_lo, mw = a.dfa.mask_words;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t t = a.redo_list ? a.redo_list[idx] : idx;
    const uint64_t lo = a.base + t * a.seg;
    const uint64_t hi = min(lo + a.seg, a.n + 1);
    uint32_t s;
    if (a.redo_list) {
      s = a.fin[t - 1];
    } else {
      const uint64_t w = (lo - a.base > a.warm) ? lo - a.warm : a.base;
      s = pick_start_fwd(a.dfa, a.text, a.n, w);
      for (uint64_t q = w; q < lo; q++) s = T.step(s, a.text[q]);
    }
    a.guess[t] = (uint16_t)s;
    uint64_t first = kNone;
    uint64_t acc[kMaxMaskWords] = {0, 0, 0, 0};
    for (uint64_t q = lo; q < hi; q++) {
      s = q < a.n ? T.step(s, a.text[q]) : T.step_eof(s);
      if (s >= match_lo) {
        if (first == kNone) first = q;
        if (a.seg_mask) {
#pragma unroll
          for (uint32_t w = 0; w < kMaxMaskWords; w++)
            if (w < mw) acc[w] |= a.dfa.masks[(uint64_t)s * mw + w];
        }
      }
    }
    a.fin[t] = (uint16_t)s;
    a.seg_first[t] = first;
    if (a.seg_mask) {
#pragma unroll
      for (uint32_t w = 0; w < kMaxMaskWords; w++)
        if (w < mw) a.seg_mask[t * mw + w] = acc[w];
    }
  }
}

// guess[t] must equal the exact state the neighbouring segment ended in.
__global__ void verify_segments(const uint16_t* guess, const uint16_t* fin, uint64_t n_seg, int reverse,
                                uint32_t* redo_list, uint32_t* n_redo) {
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n_seg; t += (uint64_t)gridDim.x * blockDim.x) {
    bool bad;
    if (reverse) bad = t + 1 < n_seg && guess[t] != fin[t + 1];
    else bad = t > 0 && guess[t] != fin[t - 1];
    if (bad) redo_list[atomicAdd(n_redo, 1u)] = (uint32_t)t;
  }
}

// result[0] = min first-match position, result[1..] = OR of masks.
__global__ void reduce_segments(const uint64_t* seg_first, const uint64_t* seg_mask, uint64_t n_seg, uint32_t mw,
                                unsigned long long* result) {
  uint64_t first = kNone;
  uint64_t acc[kMaxMaskWords] = {0, 0, 0, 0};
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n_seg; t += (uint64_t)gridDim.x * blockDim.x) {
    first = min(first, seg_first[t]);
    if (seg_mask)
      for (uint32_t w = 0; w < mw; w++) acc[w] |= seg_mask[t * mw + w];
  }
  for (int o = 16; o; o >>= 1) {
    first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    for (uint32_t w = 0; w < kMaxMaskWords; w++) acc[w] |= __shfl_xor_sync(0xffffffffu, acc[w], o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (first != kNone) atomicMin(&result[0], (unsigned long long)first);
    for (uint32_t w = 0; w < mw; w++)
      if (acc[w]) atomicOr(&result[1 + w], (unsigned long long)acc[w]);
  }
}

// ---------------------------------------------------------------- walk_chunks --
// First match start in [pos, limit), or kNone.  Bit i of the bitmap stands for
// position i+1; position 0 is *flag0.
__device__ __forceinline__ uint64_t next_bit(const uint64_t* bm, const uint8_t* flag0, uint64_t pos, uint64_t limit) {
  if (pos >= limit) return kNone;
  if (pos == 0) {
    if (*flag0) return 0;
    pos = 1;
    if (pos >= limit) return kNone;
  }
  const uint64_t b_lo = pos - 1, b_hi = limit - 1;  // bits [b_lo, b_hi)
  uint64_t w = b_lo >> 6;
  const uint64_t last_w = (b_hi - 1) >> 6;
  uint64_t word = bm[w] & (~0ull << (b_lo & 63));
  for (;;) {
    if (word) {
      uint64_t b = (w << 6) + (uint64_t)(__ffsll((long long)word) - 1);
      return b < b_hi ? b + 1 : kNone;
    }
    if (w == last_w) return kNone;
    word = bm[++w];
  }
}

// End of the leftmost-first match anchored at s (src/dfa.rs:576-764 run on the
// anchored program): last position at which a match state was entered, with the
// one-byte delay and the EOF flush.
__device__ __forceinline__ uint64_t anchored_end(const DfaView& d, const uint8_t* text, uint64_t n, uint64_t s) {
  uint32_t st = pick_start_fwd(d, text, n, s);
  uint64_t last = kNone;
  for (uint64_t q = s;; q++) {
    st = q < n ? d.trans[st * d.stride + d.classes[text[q]]] : d.trans[st * d.stride + d.stride - 1];
    if (st >= d.match_lo) last = q;
    if (st == 0 || q >= n) break;
  }
  return last;
}

// Start of the match ending at e, found the way the reference does it: the
// reverse DFA over the slice text[p..] (src/exec.rs:651-657), so position p is
// judged as beginning-of-text (SURVEY.md §8 hazard H1).
__device__ __forceinline__ uint64_t slice_start(const DfaView& d, const uint8_t* text, uint64_t n, uint64_t p, uint64_t e) {
  const uint8_t* t = text + p;
  const uint64_t len = n - p;
  uint64_t at = e - p;
  uint32_t st = d.uniform_start ? d.start[32] : d.start[flags_reverse(t, len, at)];
  uint64_t last = kNone;
  if (st == 0) return kNone;
  while (at > 0) {
    at--;
    st = d.trans[st * d.stride + d.classes[t[at]]];
    if (st == 0) return last == kNone ? kNone : p + last;
    if (st >= d.match_lo) last = at + 1;
  }
  st = d.trans[st * d.stride + d.stride - 1];
  if (st >= d.match_lo) last = 0;
  return last == kNone ? kNone : p + last;
}

__device__ __forceinline__ uint64_t next_utf8(const uint8_t* text, uint64_t n, uint64_t i) {  // src/utf8.rs:24-40
  if (i >= n) return i + 1;
  uint32_t b = text[i];
  return i + (b <= 0x7F ? 1 : b <= 0xDF ? 2 : b <= 0xEF ? 3 : 4);
}

template <bool EMIT>
__global__ void walk_chunks(WalkArgs a) {
  for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < a.n_chunks;
       k += (uint64_t)gridDim.x * blockDim.x) {
    if (!EMIT && a.dirty && !a.dirty[k]) continue;
    const uint64_t c_lo = a.base + k * a.chunk;
    const uint64_t c_hi = min(c_lo + a.chunk, a.n + 1);  // candidate positions [c_lo, c_hi)
    uint64_t p = a.in_p[k], lm = a.in_lm[k];
    const bool spec = p == kSpec;
    if (spec) { p = c_lo; lm = kNone; }
    if (p == kNone) {  // the chain ended upstream (reference iterator returned None)
      if (!EMIT) { a.out_p[k] = kNone; a.out_lm[k] = lm; a.count[k] = 0; }
      continue;
    }
    uint64_t cnt = 0;
    uint64_t w_at = EMIT ? a.offset[k] : 0;
    bool p_is_chain = !spec;  // p is a real restart point of the reference iterator
    for (;;) {
      const uint64_t s = next_bit(a.bitmap, a.flag0, max(p, c_lo), c_hi);
      if (s == kNone) break;
      uint64_t e = anchored_end(a.fwd, a.text, a.n, s);
      if (e == kNone) { p = s + 1; p_is_chain = false; continue; }  // unreachable for consistent tables
      uint64_t ms = s;
      if (a.emulate_slice && p_is_chain && e != p) {
        // exec.rs:647-657: empty match at the restart point short-circuits; otherwise
        // the start comes from the reverse DFA over text[p..].
        ms = slice_start(a.rev, a.text, a.n, p, e);
        if (ms == kNone) { p = kNone; break; }  // NoMatch => find_at None => iteration stops
      }
      if (ms == e) {
        p = a.utf8 ? next_utf8(a.text, a.n, e) : e + 1;
        p_is_chain = true;
        if (e == lm) continue;  // re_trait.rs:210-214
      } else {
        p = e;
        p_is_chain = true;
      }
      lm = e;
      if (EMIT) {
        if (w_at < a.cap) { a.out[2 * w_at] = ms; a.out[2 * w_at + 1] = e; }
        w_at++;
      }
      cnt++;
    }
    if (!EMIT) { a.out_p[k] = p; a.out_lm[k] = lm; a.count[k] = cnt; }
  }
}
template __global__ void walk_chunks<false>(WalkArgs);
template __global__ void walk_chunks<true>(WalkArgs);

__global__ void init_walk_entries(uint64_t* in_p, uint64_t* in_lm, uint64_t n_chunks, uint64_t start) {
  for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n_chunks; k += (uint64_t)gridDim.x * blockDim.x) {
    in_p[k] = k == 0 ? start : kSpec;
    in_lm[k] = kNone;
  }
}

// Compare each chunk's assumed entry state with what its predecessor produced.
__global__ void stitch_check(WalkArgs a, uint32_t* n_dirty) {
  for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < a.n_chunks;
       k += (uint64_t)gridDim.x * blockDim.x) {
    if (k == 0) { a.dirty[0] = 0; continue; }
    const uint64_t tp = a.out_p[k - 1], tl = a.out_lm[k - 1];
    const uint64_t c_lo = a.base + k * a.chunk;
    const uint64_t cp = a.in_p[k], cl = a.in_lm[k];
    bool ok;
    if (cp == kSpec) ok = tp != kNone && (tp < c_lo || (tp == c_lo && !a.emulate_slice && !(a.can_match_empty && tl == c_lo)));
    else ok = cp == tp && cl == tl;
    if (!ok) {
      a.in_p[k] = tp;
      a.in_lm[k] = tl;
      a.dirty[k] = 1;
      atomicAdd(n_dirty, 1u);
    } else {
      a.dirty[k] = 0;
    }
  }
}

// ------------------------------------------------------------- prefix sums ----
__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t* total, uint64_t* warp_sums) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint64_t x = v;
  for (int o = 1; o < 32; o <<= 1) {
    uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint64_t ws = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t y = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= o) ws += y;
    }
    warp_sums[lane] = ws;  // inclusive
  }
  __syncthreads();
  uint64_t warp_off = wid ? warp_sums[wid - 1] : 0;
  *total = warp_sums[(blockDim.x >> 5) - 1];
  return warp_off + x - v;
}
__global__ void scan_counts_local(const uint64_t* in, uint64_t* out, uint64_t* block_sums, uint64_t n) {
  __shared__ uint64_t ws[32];
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t v = i < n ? in[i] : 0, total;
  uint64_t ex = block_exclusive_scan(v, &total, ws);
  if (i < n) out[i] = ex;
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void scan_block_sums(uint64_t* block_sums, uint64_t n_blocks, unsigned long long* grand_total) {
  __shared__ uint64_t ws[32];
  __shared__ uint64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint64_t base = 0; base < n_blocks; base += blockDim.x) {
    uint64_t i = base + threadIdx.x;
    uint64_t v = i < n_blocks ? block_sums[i] : 0, total;
    uint64_t ex = block_exclusive_scan(v, &total, ws);
    uint64_t carry = carry_s;
    if (i < n_blocks) block_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand_total = carry_s;
}
__global__ void scan_add_block_offsets(uint64_t* out, const uint64_t* block_sums, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] += block_sums[blockIdx.x];
}

// ---------------------------------------------------------------- batch mode --
// One thread per record; each record is its own haystack (record-local ^, $, \b).
__global__ void is_match_batch(BatchArgs a) {
  const Table T = stage_table(a.fwd, g_smem, a.use_smem);
  const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  bool matched = false;
  if (r < a.n_rec) {
    const uint64_t lo = a.offsets[r], len = a.offsets[r + 1] - lo;
    const uint8_t* p = a.text + lo;
    uint32_t s = pick_start_fwd(a.fwd, p, len, 0);
    for (uint64_t q = 0; s != 0; q++) {
      s = q < len ? T.step(s, p[q]) : T.step_eof(s);
      if (s >= a.fwd.match_lo) { matched = true; break; }
      if (q >= len) break;
    }
  }
  const uint32_t bits = __ballot_sync(0xffffffffu, matched);
  if ((threadIdx.x & 31) == 0 && r < a.n_rec) a.out_bits[r >> 5] = bits;
}

// exec.rs:632-662 per record: forward leftmost-first end, then reverse longest start.
__global__ void find_batch(BatchArgs a) {
  const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  bool found = false;
  if (r < a.n_rec) {
    const uint64_t lo = a.offsets[r], len = a.offsets[r + 1] - lo;
    const uint8_t* p = a.text + lo;
    const DfaView& f = a.fwd;
    uint32_t s = pick_start_fwd(f, p, len, 0);
    uint64_t e = kNone;
    for (uint64_t q = 0; s != 0; q++) {
      s = q < len ? f.trans[s * f.stride + f.classes[p[q]]] : f.trans[s * f.stride + f.stride - 1];
      if (s >= f.match_lo) e = q;
      if (q >= len) break;
    }
    uint64_t ms = kNone;
    if (e != kNone) ms = e == 0 ? 0 : slice_start(a.rev, p, len, 0, e);
    if (ms != kNone) {
      found = true;
      a.out_spans[2 * r] = ms;
      a.out_spans[2 * r + 1] = e;
    } else {
      a.out_spans[2 * r] = 0;
      a.out_spans[2 * r + 1] = 0;
    }
  }
  const uint32_t bits = __ballot_sync(0xffffffffu, found);
  if ((threadIdx.x & 31) == 0 && r < a.n_rec) a.out_bits[r >> 5] = bits;
}

// dfa.rs:525-570 per record: OR of the per-state pattern masks along the scan.
__global__ void set_matches_batch(BatchArgs a) {
  const Table T = stage_table(a.fwd, g_smem, a.use_smem);
  const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (r >= a.n_rec) return;
  const uint64_t lo = a.offsets[r], len = a.offsets[r + 1] - lo;
  const uint8_t* p = a.text + lo;
  const uint32_t mw = a.fwd.mask_words;
  uint64_t acc[kMaxMaskWords] = {0, 0, 0, 0};
  uint32_t s = pick_start_fwd(a.fwd, p, len, 0);
  for (uint64_t q = 0; s != 0; q++) {
    s = q < len ? T.step(s, p[q]) : T.step_eof(s);
    if (s >= a.fwd.match_lo) {
#pragma unroll
      for (uint32_t w = 0; w < kMaxMaskWords; w++)
        if (w < mw) acc[w] |= a.fwd.masks[(uint64_t)s * mw + w];
    }
    if (q >= len) break;
  }
#pragma unroll
  for (uint32_t w = 0; w < kMaxMaskWords; w++)
    if (w < mw) a.out_masks[r * mw + w] = acc[w];
}

}  // namespace rbgpu
