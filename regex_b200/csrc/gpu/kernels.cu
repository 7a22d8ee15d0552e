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
// Segment t owns positions [base + t*seg, base + (t+1)*seg) ∩ [0, n].  Position q
// is a match START iff the reverse automaton, standing at q, reports a (delayed)
// match when it consumes text[q-1] (or EOF for q == 0).  The state at the top
// of a segment is guessed by running the automaton from a fresh start state
// `warm` bytes further right; verify_segments() checks every guess against the
// neighbour's exact final state and lists the segments to redo.
__global__ void scan_rev_bitmap(ScanArgs a) {
  const Table T = stage_table(a.dfa, g_smem, a.use_smem);
  const uint64_t total = a.redo_list ? (uint64_t)*a.n_redo : a.n_seg;
  const uint32_t match_lo = a.dfa.match_lo;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t t = a.redo_list ? a.redo_list[idx] : idx;
    const uint64_t lo = a.base + t * a.seg;
    const uint64_t hi = min(lo + a.seg, a.n + 1);  // exclusive
    const uint64_t q_hi = hi - 1;
    uint32_t s;
    if (a.redo_list) {
      s = a.fin[t + 1];
    } else {
      const uint64_t w = min(q_hi + a.warm, a.n);
      s = pick_start_rev(a.dfa, a.text, a.n, w);
      for (uint64_t q = w; q > q_hi; q--) s = T.step(s, a.text[q - 1]);
    }
    a.guess[t] = (uint16_t)s;
    uint64_t word = 0;
    uint32_t byte_at_q = q_hi < a.n ? a.text[q_hi] : 0;  // text[q], for the UTF-8 boundary mask
    for (uint64_t q = q_hi;; q--) {
      uint32_t b = 0;
      if (q > 0) { b = a.text[q - 1]; s = T.step(s, b); } else { s = T.step_eof(s); }
      bool hit = s >= match_lo;
      if (a.utf8_boundaries && (byte_at_q & 0xC0) == 0x80) hit = false;
      if (hit) word |= 1ull << (q & 63);
      byte_at_q = b;
      if ((q & 63) == 0) { a.bitmap[q >> 6] = word; word = 0; }
      if (q == lo) break;
    }
    a.fin[t] = (uint16_t)s;
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
// First set bit of the start bitmap in [pos, limit), or kNone.
__device__ __forceinline__ uint64_t next_bit(const uint64_t* bm, uint64_t pos, uint64_t limit) {
  if (pos >= limit) return kNone;
  uint64_t w = pos >> 6;
  const uint64_t last_w = (limit - 1) >> 6;
  uint64_t word = bm[w] & (~0ull << (pos & 63));
  for (;;) {
    if (word) {
      uint64_t q = (w << 6) + (uint64_t)(__ffsll((long long)word) - 1);
      return q < limit ? q : kNone;
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
      const uint64_t s = next_bit(a.bitmap, max(p, c_lo), c_hi);
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
