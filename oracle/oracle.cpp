// TEST INFRASTRUCTURE ONLY -- CPU oracle for the regex_b200 parity tests.
//
// A CPU restatement of the reference's search engines, used as the checker in
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs.  Nothing in the product (regex_b200/, include/) links, imports or calls
// this file; the product path fails loudly when its CUDA library is missing.
//
// Parity status: PINNED.  The reference (Rust, crate regex 0.2.5) cannot be built
// in this environment (no rustc/cargo), so the oracle is pinned against the
// reference's own known-answer vectors instead: tests/*.rs mat!/matiter!/matset!
// cases, src/testdata/*.dat (AT&T suite), bench/src/sherlock.rs counts and
// examples/regexdna-output.txt -- see tests/golden/ and tools/scrape_golden.py.
//
// Engines restated here (each cites the reference file:line it follows):
//   PikeVm      src/pikevm.rs:130-352 with look-around per src/input.rs:268-318
//               -- semantic ground truth ("all engines expose precisely the
//               same matching semantics", HACKING.md:60-61).
//   LazyDfa     src/dfa.rs:576-1580 -- on-line subset construction with a
//               state cache, forward leftmost-first and reverse longest scans,
//               one-byte match delay, EOF sentinel, set scan.
//   DfaSuffix   src/exec.rs:725-794, 1176-1210 -- suffix-literal scan (memmem), reverse
//               DFA from the literal's end, forward DFA from the start; the engine the
//               reference selects for `[a-zA-Z]+ing` and the one the CPU baseline times
//               (engine 2 = "auto"; recognised for concatenations ending in >= 3 literal
//               characters, otherwise the plain pipeline).  The Teddy / Aho-Corasick
//               literal engines of src/literals.rs are not restated.
//   Exec        src/exec.rs:382-514, 632-662, 998-1038 -- shortest/is_match/find
//               pipeline (forward DFA then reverse DFA over &text[start..]).
//   iterate()   src/re_trait.rs:197-220 -- the find_iter chaining rule.
// The byte-level NFA program comes from the shared host front end
// (regex_b200/csrc/frontend, a restatement of regex-syntax + src/compile.rs).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../regex_b200/csrc/frontend/frontend.h"

namespace oracle {

using rb::Inst;
using rb::Look;
using rb::Op;
using rb::Program;

static const size_t NONE = (size_t)-1;

static inline bool is_word_byte(int b) {
  return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_';
}

// Length of the valid UTF-8 scalar at p (0 if none).
static size_t valid_utf8_len(const uint8_t* p, size_t n) {
  if (n == 0) return 0;
  uint8_t b = p[0];
  if (b < 0x80) return 1;
  size_t len = (b >= 0xC2 && b <= 0xDF) ? 2 : (b >= 0xE0 && b <= 0xEF) ? 3 : (b >= 0xF0 && b <= 0xF4) ? 4 : 0;
  if (len == 0 || len > n) return 0;
  for (size_t k = 1; k < len; k++) if ((p[k] & 0xC0) != 0x80) return 0;
  if (len == 3 && ((b == 0xE0 && p[1] < 0xA0) || (b == 0xED && p[1] > 0x9F))) return 0;
  if (len == 4 && ((b == 0xF0 && p[1] < 0x90) || (b == 0xF4 && p[1] > 0x8F))) return 0;
  return len;
}

// ------------------------------------------------------------------ Pike VM --
// src/pikevm.rs:130-352, byte input.  Thread list = insertion-ordered sparse set;
// each thread carries the start offset of its match (capture slot 0).
struct PikeVm {
  const Program& prog;
  struct Threads {
    std::vector<uint32_t> dense, sparse;
    std::vector<size_t> start;  // per ip
    void init(size_t n) { dense.clear(); dense.reserve(n); sparse.assign(n, 0); start.assign(n, 0); }
    bool contains(uint32_t ip) const { uint32_t i = sparse[ip]; return i < dense.size() && dense[i] == ip; }
    void insert(uint32_t ip) { sparse[ip] = (uint32_t)dense.size(); dense.push_back(ip); }
  };
  explicit PikeVm(const Program& p) : prog(p) {}

  // src/input.rs:268-318 (ByteInput); ASCII word boundaries only -- Unicode
  // \b is rejected before we get here.
  static bool look_ok(Look l, const uint8_t* t, size_t n, size_t at) {
    bool w1 = at > 0 && is_word_byte(t[at - 1]);
    bool w2 = at < n && is_word_byte(t[at]);
    switch (l) {
      case Look::StartLine: return at == 0 || t[at - 1] == '\n';
      case Look::EndLine: return at == n || t[at] == '\n';
      case Look::StartText: return at == 0;
      case Look::EndText: return at == n;
      case Look::WordBoundary: case Look::WordBoundaryAscii: return w1 != w2;
      case Look::NotWordBoundary: case Look::NotWordBoundaryAscii: return w1 == w2;
    }
    return false;
  }
  void add(Threads& l, uint32_t ip0, size_t start, const uint8_t* t, size_t n, size_t at, std::vector<uint32_t>& stack) {
    stack.push_back(ip0);
    while (!stack.empty()) {
      uint32_t ip = stack.back();
      stack.pop_back();
      for (;;) {
        if (l.contains(ip)) break;
        l.insert(ip);
        const Inst& in = prog.insts[ip];
        if (in.op == Op::EmptyLook) {
          if (!look_ok(in.look, t, n, at)) break;
          ip = in.a;
        } else if (in.op == Op::Save) {
          ip = in.a;
        } else if (in.op == Op::Split) {
          stack.push_back(in.b);
          ip = in.a;
        } else {
          l.start[ip] = start;
          break;
        }
      }
    }
  }
  // Returns true on match.  matches (size n_patterns) is OR-ed; (ms,me) is the
  // leftmost-first span for single-pattern programs.
  bool exec(const uint8_t* t, size_t n, size_t start, bool quit_after_match, std::vector<bool>* matches,
            size_t* ms, size_t* me) {
    Threads clist, nlist;
    clist.init(prog.insts.size());
    nlist.init(prog.insts.size());
    std::vector<uint32_t> stack;
    bool matched = false, all_matched = false;
    const bool single = prog.matches.size() == 1;
    size_t at = start;
    // `Regex` (only_utf8) programs scan with a lazy AnyChar* prefix (compile.rs:381-395):
    // new threads start only where whole valid UTF-8 scalars have been skipped.
    size_t next_boundary = start;
    for (;;) {
      if (clist.dense.empty()) {
        if ((matched && single) || all_matched || (at != 0 && prog.is_anchored_start)) break;
      }
      bool may_start = true;
      if (prog.only_utf8 && at != start) {
        may_start = at == next_boundary;
      }
      if (prog.only_utf8 && at == next_boundary) {
        size_t l = valid_utf8_len(t + at, n - at);
        next_boundary = l ? at + l : NONE;
      }
      if (may_start && (clist.dense.empty() || (!prog.is_anchored_start && !all_matched)))
        add(clist, prog.start_anchored, at, t, n, at, stack);
      for (size_t i = 0; i < clist.dense.size(); i++) {
        uint32_t ip = clist.dense[i];
        const Inst& in = prog.insts[ip];
        if (in.op == Op::Match) {
          if (matches && in.a < matches->size()) (*matches)[in.a] = true;
          if (ms) { *ms = clist.start[ip]; *me = at; }
          matched = true;
          if (matches) {
            all_matched = true;
            for (bool b : *matches) all_matched = all_matched && b;
          } else {
            all_matched = single;
          }
          if (quit_after_match) return true;
          if (single) break;  // leftmost-first: drop lower-priority threads
        } else if (in.op == Op::Bytes) {
          if (at < n && in.lo <= t[at] && t[at] <= in.hi) add(nlist, in.a, clist.start[ip], t, n, at + 1, stack);
        }
      }
      if (at >= n) break;
      at++;
      std::swap(clist, nlist);
      nlist.dense.clear();
    }
    return matched;
  }
};

// ------------------------------------------------------- Pike VM with captures --
// src/pikevm.rs:130-352 with the per-thread capture slots: what the reference runs to fill
// `Captures` once the DFA has found the match (src/exec.rs:861-875, captures_nfa_with_match:
// the NFA over text[..e'] from the match start, e' two characters past the match end).
struct CapVm {
  const Program& prog;
  size_t n_slots;
  struct Threads {
    std::vector<uint32_t> dense, sparse;
    std::vector<size_t> caps;  // [ip][slot]
    size_t n_slots = 0;
    void init(size_t n, size_t slots) { dense.clear(); sparse.assign(n, 0); caps.assign(n * slots, NONE); n_slots = slots; }
    bool contains(uint32_t ip) const { uint32_t i = sparse[ip]; return i < dense.size() && dense[i] == ip; }
    void insert(uint32_t ip) { sparse[ip] = (uint32_t)dense.size(); dense.push_back(ip); }
  };
  struct Frame { bool is_cap; uint32_t ip_or_slot; size_t pos; };
  std::vector<Frame> stack;
  explicit CapVm(const Program& p) : prog(p), n_slots(2 * (size_t)p.n_captures) {}

  void add(Threads& l, std::vector<size_t>& tc, uint32_t ip0, const uint8_t* t, size_t n, size_t at) {  // pikevm.rs:274-352
    stack.push_back({false, ip0, 0});
    while (!stack.empty()) {
      Frame f = stack.back();
      stack.pop_back();
      if (f.is_cap) { tc[f.ip_or_slot] = f.pos; continue; }
      uint32_t ip = f.ip_or_slot;
      for (;;) {
        if (l.contains(ip)) break;
        l.insert(ip);
        const Inst& in = prog.insts[ip];
        if (in.op == Op::EmptyLook) {
          if (!PikeVm::look_ok(in.look, t, n, at)) break;
          ip = in.a;
        } else if (in.op == Op::Save) {
          if (in.b < n_slots) {
            stack.push_back({true, in.b, tc[in.b]});
            tc[in.b] = at;
          }
          ip = in.a;
        } else if (in.op == Op::Split) {
          stack.push_back({false, in.b, 0});
          ip = in.a;
        } else {
          for (size_t k = 0; k < n_slots; k++) l.caps[ip * n_slots + k] = tc[k];
          break;
        }
      }
    }
  }
  bool exec(const uint8_t* t, size_t n, size_t start, std::vector<size_t>& slots) {  // pikevm.rs:130-228, one regex
    Threads clist, nlist;
    clist.init(prog.insts.size(), n_slots);
    nlist.init(prog.insts.size(), n_slots);
    slots.assign(n_slots, NONE);
    std::vector<size_t> tc(n_slots, NONE);
    bool matched = false;
    for (size_t at = start;; at++) {
      if (clist.dense.empty()) {
        if (matched || (at != 0 && prog.is_anchored_start)) break;  // pikevm.rs:143-158 (at.is_start() is position 0)
      }
      // a new thread at every position until something has matched (pikevm.rs:172-175): the lazy
      // prefix of the program; its slots start out empty (the caller's slots are all None then)
      if (clist.dense.empty() || (!prog.is_anchored_start && !matched)) {
        std::fill(tc.begin(), tc.end(), NONE);
        add(clist, tc, prog.start, t, n, at);
      }
      for (size_t i = 0; i < clist.dense.size(); i++) {
        const uint32_t ip = clist.dense[i];
        const Inst& in = prog.insts[ip];
        if (in.op == Op::Match) {
          for (size_t k = 0; k < n_slots; k++) slots[k] = clist.caps[ip * n_slots + k];
          matched = true;
          break;  // leftmost-first
        }
        if (in.op == Op::Bytes && at < n && in.lo <= t[at] && t[at] <= in.hi) {
          std::vector<size_t> caps_of(clist.caps.begin() + ip * n_slots, clist.caps.begin() + (ip + 1) * n_slots);
          add(nlist, caps_of, in.a, t, n, at + 1);
        }
      }
      if (at >= n) break;
      std::swap(clist, nlist);
      nlist.dense.clear();
    }
    return matched;
  }
};

// ----------------------------------------------------------------- lazy DFA --
// src/dfa.rs.  States are created on demand and cached for the lifetime of the
// object (the reference's cache flush / give-up heuristics, dfa.rs:1282-1320,
// only decide *when* it falls back to the NFA, never the result).
struct LazyDfa {
  struct Result { enum K { Match, NoMatch } k; size_t pos; };
  static constexpr uint32_t UNKNOWN = 0xFFFFFFFFu, DEAD = 0xFFFFFFFEu;

  const Program& prog;
  struct State { std::vector<uint32_t> ips; uint8_t flags; };
  static constexpr uint8_t F_MATCH = 1, F_WORD = 2, F_EMPTY = 4;
  std::vector<State> states;
  // Transition cache laid out like the reference's (dfa.rs:236-262, 373-400): one flat
  // table, row = state, column = byte class; an entry is UNKNOWN, DEAD, or the successor's
  // ROW OFFSET (index * ncls, the reference's premultiplied state pointer) with MATCH_BIT set
  // when the successor is a match state, so the scan loop (dfa.rs:636-657) is two dependent
  // loads and an add per byte and never looks at the state itself.
  static constexpr uint32_t MATCH_BIT = 1u << 30, ID_MASK = MATCH_BIT - 1;
  std::vector<uint32_t> flat;
  std::map<std::pair<uint8_t, std::vector<uint32_t>>, uint32_t> cache;
  uint32_t start_states[128];
  int ncls;
  std::vector<uint32_t> stack;
  uint32_t last_match_off = UNKNOWN;  // row offset of the last match state seen (forward_many reads its ips)

  struct EmptyFlags { bool start = false, end = false, start_line = false, end_line = false, wb = false, nwb = false; };
  struct Sparse {
    std::vector<uint32_t> dense, sparse;
    void init(size_t n) { dense.clear(); sparse.assign(n, 0); }
    bool contains(uint32_t v) const { uint32_t i = sparse[v]; return i < dense.size() && dense[i] == v; }
    void insert(uint32_t v) { sparse[v] = (uint32_t)dense.size(); dense.push_back(v); }
    void clear() { dense.clear(); }
  } qcur, qnext;

  explicit LazyDfa(const Program& p) : prog(p) {
    ncls = p.num_classes + 1;  // + EOF (dfa.rs:1528-1531)
    for (auto& s : start_states) s = UNKNOWN;
    qcur.init(p.insts.size());
    qnext.init(p.insts.size());
  }
  bool continue_past_first_match() const { return prog.is_reverse || prog.matches.size() > 1; }  // dfa.rs:1557

  void follow_epsilons(uint32_t ip0, Sparse& q, const EmptyFlags& f) {  // dfa.rs:1073-1134
    stack.push_back(ip0);
    while (!stack.empty()) {
      uint32_t ip = stack.back();
      stack.pop_back();
      if (q.contains(ip)) continue;
      q.insert(ip);
      const Inst& in = prog.insts[ip];
      switch (in.op) {
        case Op::Match: case Op::Bytes: break;
        case Op::Save: stack.push_back(in.a); break;
        case Op::Split: stack.push_back(in.b); stack.push_back(in.a); break;
        case Op::EmptyLook: {
          bool ok = false;
          switch (in.look) {
            case Look::StartLine: ok = f.start_line; break;
            case Look::EndLine: ok = f.end_line; break;
            case Look::StartText: ok = f.start; break;
            case Look::EndText: ok = f.end; break;
            case Look::WordBoundary: case Look::WordBoundaryAscii: ok = f.wb; break;
            case Look::NotWordBoundary: case Look::NotWordBoundaryAscii: ok = f.nwb; break;
          }
          if (ok) stack.push_back(in.a);
        }
      }
    }
  }
  uint32_t cached_state(const Sparse& q, uint8_t flags) {  // dfa.rs:1154-1244
    std::vector<uint32_t> ips;
    for (uint32_t ip : q.dense) {
      const Inst& in = prog.insts[ip];
      if (in.op == Op::Save || in.op == Op::Split) continue;
      ips.push_back(ip);
      if (in.op == Op::EmptyLook) flags |= F_EMPTY;
      if (in.op == Op::Match && !continue_past_first_match()) break;
    }
    if (ips.empty() && !(flags & F_MATCH)) return DEAD;
    auto key = std::make_pair(flags, ips);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    uint32_t id = (uint32_t)states.size();
    states.push_back(State{std::move(ips), flags});
    flat.resize(states.size() * (size_t)ncls, UNKNOWN);
    cache.emplace(std::move(key), id);
    return id;
  }
  // b in [0,255], 256 = EOF.  dfa.rs:910-1048.
  uint32_t exec_byte(uint32_t si, int b) {
    qcur.clear();
    for (uint32_t ip : states[si].ips) qcur.insert(ip);
    const bool eof = b == 256;
    const bool is_word_last = states[si].flags & F_WORD;
    const bool is_word = !eof && is_word_byte(b);
    if (states[si].flags & F_EMPTY) {
      EmptyFlags f;
      if (eof) { f.end = true; f.end_line = true; }
      else if (b == '\n') f.end_line = true;
      if (is_word_last == is_word) f.nwb = true; else f.wb = true;
      qnext.clear();
      for (uint32_t ip : qcur.dense) follow_epsilons(ip, qnext, f);
      std::swap(qcur, qnext);
    }
    EmptyFlags f;
    f.start_line = !eof && b == '\n';
    uint8_t flags = is_word ? F_WORD : 0;
    qnext.clear();
    for (uint32_t ip : qcur.dense) {
      const Inst& in = prog.insts[ip];
      if (in.op == Op::Match) {
        flags |= F_MATCH;
        if (!continue_past_first_match()) break;
        if (prog.matches.size() > 1 && !qnext.contains(ip)) qnext.insert(ip);  // dfa.rs:988-993
      } else if (in.op == Op::Bytes) {
        if (!eof && in.lo <= b && b <= in.hi) follow_epsilons(in.a, qnext, f);
      }
    }
    bool cache_it = true;
    if (eof && prog.matches.size() > 1) { std::swap(qcur, qnext); cache_it = false; }  // dfa.rs:1004-1015
    uint32_t next = cached_state(qnext, flags);
    if (cache_it) flat[(size_t)si * ncls + cls_of(b)] = encode(next);
    return next;
  }
  int cls_of(int b) const { return b == 256 ? ncls - 1 : prog.byte_classes[b]; }
  uint32_t encode(uint32_t id) const {
    return id == DEAD ? DEAD : (id * (uint32_t)ncls) | ((states[id].flags & F_MATCH) ? MATCH_BIT : 0u);
  }
  uint32_t next_state(uint32_t si, int b) {  // dfa.rs:1345-1361
    if (si == DEAD) return DEAD;
    uint32_t n = flat[(size_t)si * ncls + cls_of(b)];
    if (n == UNKNOWN) return exec_byte(si, b);
    return n == DEAD ? DEAD : (n & ID_MASK) / (uint32_t)ncls;
  }
  // One byte of the scan loops from row offset `off`: the cached entry if there is one,
  // else compute (and cache) it.  DEAD and UNKNOWN are the two largest values.
  inline uint32_t step_cached(uint32_t off, uint8_t b) {
    uint32_t v = flat[off + prog.byte_classes[b]];
    if (v == UNKNOWN) v = encode(exec_byte(off / (uint32_t)ncls, b));
    return v;
  }
  uint32_t start_state(int flagi) {  // dfa.rs:1370-1409
    if (start_states[flagi] != UNKNOWN) return start_states[flagi];
    EmptyFlags f;
    f.start = flagi & 1; f.end = flagi & 2; f.start_line = flagi & 4; f.end_line = flagi & 8;
    f.wb = flagi & 16; f.nwb = flagi & 32;
    qcur.clear();
    follow_epsilons(prog.start, qcur, f);
    return start_states[flagi] = cached_state(qcur, (flagi & 64) ? F_WORD : 0);
  }
  static int flags_forward(const uint8_t* t, size_t n, size_t at) {  // dfa.rs:1415-1434
    int f = 0;
    if (at == 0) f |= 1;
    if (n == 0) f |= 2 | 8;
    if (at == 0 || t[at - 1] == '\n') f |= 4;
    bool last = at > 0 && is_word_byte(t[at - 1]);
    bool cur = at < n && is_word_byte(t[at]);
    f |= (cur == last) ? 32 : 16;
    if (last) f |= 64;
    return f;
  }
  static int flags_reverse(const uint8_t* t, size_t n, size_t at) {  // dfa.rs:1440-1464
    int f = 0;
    if (at == n) f |= 1;
    if (n == 0) f |= 2 | 8;
    if (at == n || t[at] == '\n') f |= 4;
    bool last = at < n && is_word_byte(t[at]);
    bool cur = at > 0 && is_word_byte(t[at - 1]);
    f |= (cur == last) ? 32 : 16;
    if (last) f |= 64;
    return f;
  }
  // dfa.rs:576-764 without the unrolling / prefix-skip fast paths.
  Result forward(const uint8_t* t, size_t n, size_t at, bool quit_after_match) {
    last_match_off = UNKNOWN;
    uint32_t si = start_state(flags_forward(t, n, at));
    Result r{Result::NoMatch, at};
    if (si == DEAD) return r;
    uint32_t off = si * (uint32_t)ncls;
    while (at < n) {
      const uint32_t v = step_cached(off, t[at]);
      at++;
      if (v == DEAD) { if (r.k == Result::NoMatch) r.pos = at; return r; }
      off = v & ID_MASK;
      if (v & MATCH_BIT) {
        r = Result{Result::Match, at - 1};
        if (quit_after_match) return r;
        last_match_off = off;
        if (prog.matches.size() > 1) {  // dfa.rs:675-682
          si = off / (uint32_t)ncls;
          bool just = true;
          for (uint32_t ip : states[si].ips) just = just && prog.insts[ip].op == Op::Match;
          if (just) return r;
        }
      }
    }
    si = next_state(off / (uint32_t)ncls, 256);
    if (si == DEAD) { if (r.k == Result::NoMatch) r.pos = n; return r; }
    if (states[si].flags & F_MATCH) { last_match_off = si * (uint32_t)ncls; r = Result{Result::Match, n}; }
    return r;
  }
  // dfa.rs:768-866
  Result reverse(const uint8_t* t, size_t n, size_t at, bool quit_after_match) {
    last_match_off = UNKNOWN;
    uint32_t si = start_state(flags_reverse(t, n, at));
    Result r{Result::NoMatch, at};
    if (si == DEAD) return r;
    uint32_t off = si * (uint32_t)ncls;
    while (at > 0) {
      at--;
      const uint32_t v = step_cached(off, t[at]);
      if (v == DEAD) { if (r.k == Result::NoMatch) r.pos = at; return r; }
      off = v & ID_MASK;
      if (v & MATCH_BIT) {
        r = Result{Result::Match, at + 1};
        if (quit_after_match) return r;
        last_match_off = off;
      }
    }
    si = next_state(off / (uint32_t)ncls, 256);
    if (si == DEAD) { if (r.k == Result::NoMatch) r.pos = 0; return r; }
    if (states[si].flags & F_MATCH) { last_match_off = si * (uint32_t)ncls; r = Result{Result::Match, 0}; }
    return r;
  }
  // dfa.rs:525-570
  bool forward_many(const uint8_t* t, size_t n, size_t at, std::vector<bool>& matches) {
    Result r = forward(t, n, at, false);
    if (r.k != Result::Match) return false;
    if (matches.size() == 1) { matches[0] = true; return true; }
    for (uint32_t ip : states[last_match_off / (uint32_t)ncls].ips)
      if (prog.insts[ip].op == Op::Match) matches[prog.insts[ip].a] = true;
    return true;
  }
};

// ----------------------------------------------------------------- executor --
struct Regex {
  std::vector<std::string> pats;
  Program nfa, dfa, dfa_rev;
  Program caps;  // the forward program (lazy prefix included) WITH Save instructions for every group (captures)
  std::unique_ptr<LazyDfa> fwd, rev;
  bool only_utf8 = false;
  std::string error;

  // Longest common suffix literal when the reference would pick MatchType::DfaSuffix
  // (exec.rs:1176-1210: at least 3 chars and longer than the common prefix literal).  Only
  // the simple case is recognised here -- a concatenation that ends in plain literal
  // characters and does not start with one -- which covers the headline `[a-zA-Z]+ing`.
  std::string lcs;

  // exec.rs:725-756: scan for the suffix literal, then the reverse DFA from its end over the
  // slice that starts at the previous literal's end.  Returns 0 = no match, 1 = match start
  // in *ms, 2 = gave up (the reverse scan reached the slice start: avoid quadratic time).
  int exec_dfa_reverse_suffix(const uint8_t* t, size_t n, size_t original_start, size_t* ms) {
    size_t start = original_start, end = start;
    while (end <= n) {
      start = end;
      const void* hit = memmem(t + end, n - end, lcs.data(), lcs.size());
      if (!hit) return 0;
      end = (size_t)((const uint8_t*)hit - t) + lcs.size();
      auto r = rev->reverse(t + start, end - start, end - start, false);
      if (r.pos == 0) return 2;  // Match(0) | NoMatch(0)
      if (r.k == LazyDfa::Result::Match) { *ms = r.pos + start; return 1; }
    }
    return 0;
  }

  // engine: 0 = default pipeline (lazy DFA forward + reverse-on-slice), 1 = Pike VM,
  // 2 = the engine the reference itself selects: DfaSuffix when `lcs` is set (exec.rs:764-794),
  // else the default pipeline
  bool find_at(int engine, const uint8_t* t, size_t n, size_t start, size_t* s, size_t* e) {
    if (engine == 1) {
      PikeVm vm(nfa);
      return vm.exec(t, n, start, false, nullptr, s, e);
    }
    if (engine == 2 && !lcs.empty()) {
      size_t ms;
      const int k = exec_dfa_reverse_suffix(t, n, start, &ms);
      if (k == 0) return false;
      if (k == 1) {
        auto r = fwd->forward(t, n, ms, false);  // the literal only gives the earliest possible end
        if (r.k != LazyDfa::Result::Match) return false;
        *s = ms;
        *e = r.pos;
        return true;
      }
      // gave up: fall through to find_dfa_forward
    }
    // src/exec.rs:632-662
    auto r = fwd->forward(t, n, start, false);
    if (r.k != LazyDfa::Result::Match) return false;
    if (r.pos == start) { *s = start; *e = start; return true; }
    auto r2 = rev->reverse(t + start, n - start, r.pos - start, false);
    if (r2.k != LazyDfa::Result::Match) return false;
    *s = start + r2.pos;
    *e = r.pos;
    return true;
  }
  bool shortest_at(int engine, const uint8_t* t, size_t n, size_t start, size_t* e) {  // exec.rs:382-420
    if (engine == 1) {
      PikeVm vm(nfa);
      size_t s;
      return vm.exec(t, n, start, true, nullptr, &s, e);
    }
    auto r = fwd->forward(t, n, start, true);
    if (r.k != LazyDfa::Result::Match) return false;
    *e = r.pos;
    return true;
  }
  bool many_at(int engine, const uint8_t* t, size_t n, size_t start, std::vector<bool>& m) {  // exec.rs:998-1038
    if (engine == 1) {
      PikeVm vm(nfa);
      return vm.exec(t, n, start, false, &m, nullptr, nullptr);
    }
    return fwd->forward_many(t, n, start, m);
  }
  // exec.rs:527-590 read_captures_at with MatchType::Dfa: the DFA pipeline finds the match, the
  // NFA fills the groups over the narrowed window (exec.rs:861-875).
  bool captures_at(const uint8_t* t, size_t n, size_t start, std::vector<size_t>& slots) {
    size_t s, e;
    if (!find_at(0, t, n, start, &s, &e)) return false;
    auto next_utf8 = [&](size_t i) {  // src/utf8.rs:24-40
      if (i >= n) return i + 1;
      const uint8_t b = t[i];
      return i + (b <= 0x7F ? 1 : b <= 0xDF ? 2 : b <= 0xEF ? 3 : 4);
    };
    const size_t e2 = std::min(next_utf8(next_utf8(e)), n);
    CapVm vm(caps);
    return vm.exec(t, e2, s, slots);
  }
  size_t next_after_empty(const uint8_t* t, size_t n, size_t i) const {
    if (!only_utf8) return i + 1;  // exec.rs:375-377
    if (i >= n) return i + 1;      // utf8.rs:24-40
    uint8_t b = t[i];
    return i + (b <= 0x7F ? 1 : b <= 0xDF ? 2 : b <= 0xEF ? 3 : 4);
  }
};

static Regex* build(const std::vector<std::string>& pats, uint32_t flags, bool only_utf8, size_t size_limit) {
  auto re = std::make_unique<Regex>();
  re->pats = pats;
  re->only_utf8 = only_utf8;
  rb::Flags f;
  f.casei = flags & 1; f.multi = flags & 2; f.dotnl = flags & 4; f.swap_greed = flags & 8;
  f.ignore_space = flags & 16; f.unicode = flags & 32;
  f.allow_bytes = !only_utf8;  // exec.rs:224-225: .unicode(..) then .allow_bytes(!only_utf8) wins
  std::vector<rb::Expr> exprs(pats.size());
  rb::Error err;
  for (size_t i = 0; i < pats.size(); i++)
    if (!rb::parse(pats[i], f, 200, &exprs[i], &err)) { re->error = err.msg; return re.release(); }
  if (pats.empty()) return re.release();
  rb::CompileOptions o;
  o.only_utf8 = only_utf8;
  o.size_limit = size_limit;
  o.unanchored_prefix = false; o.saves = false;
  if (!rb::compile(exprs, o, &re->nfa, &err)) { re->error = err.msg; return re.release(); }
  if (exprs.size() == 1) {
    o.saves = true;
    o.unanchored_prefix = true;
    if (!rb::compile(exprs, o, &re->caps, &err)) { re->error = err.msg; return re.release(); }
    o.saves = false;
  }
  o.unanchored_prefix = true;
  if (!rb::compile(exprs, o, &re->dfa, &err)) { re->error = err.msg; return re.release(); }
  o.reverse = true; o.unanchored_prefix = false;
  if (!rb::compile(exprs, o, &re->dfa_rev, &err)) { re->error = err.msg; return re.release(); }
  if (re->dfa.has_unicode_word_boundary) {
    // The reference handles Unicode \b by quitting the DFA on non-ASCII bytes and
    // re-running the NFA with Unicode-aware look-around (dfa.rs:1491-1496); on
    // pure-ASCII haystacks both reduce to the ASCII rule implemented here.
  }
  re->fwd = std::make_unique<LazyDfa>(re->dfa);
  re->rev = std::make_unique<LazyDfa>(re->dfa_rev);
  if (exprs.size() == 1) {
    const rb::Expr* x = &exprs[0];
    while (x->kind == rb::EK::Group && x->es.size() == 1) x = &x->es[0];
    if (x->kind == rb::EK::Concat && !x->es.empty()) {
      auto literal_bytes = [](const rb::Expr& c, std::string* out) {
        if (c.casei) return false;
        if (c.kind == rb::EK::Literal) {
          for (uint32_t ch : c.chars) {  // UTF-8 encode
            if (ch < 0x80) out->push_back((char)ch);
            else if (ch < 0x800) { out->push_back((char)(0xC0 | (ch >> 6))); out->push_back((char)(0x80 | (ch & 0x3F))); }
            else if (ch < 0x10000) { out->push_back((char)(0xE0 | (ch >> 12))); out->push_back((char)(0x80 | ((ch >> 6) & 0x3F))); out->push_back((char)(0x80 | (ch & 0x3F))); }
            else { out->push_back((char)(0xF0 | (ch >> 18))); out->push_back((char)(0x80 | ((ch >> 12) & 0x3F))); out->push_back((char)(0x80 | ((ch >> 6) & 0x3F))); out->push_back((char)(0x80 | (ch & 0x3F))); }
          }
          return true;
        }
        if (c.kind == rb::EK::LiteralBytes) { out->append((const char*)c.bytes.data(), c.bytes.size()); return true; }
        return false;
      };
      std::string head;
      const bool prefix_literal = literal_bytes(x->es.front(), &head);
      size_t first = x->es.size();
      while (first > 0) {
        std::string tmp;
        if (!literal_bytes(x->es[first - 1], &tmp)) break;
        first--;
      }
      std::string suffix;
      for (size_t i = first; i < x->es.size(); i++) literal_bytes(x->es[i], &suffix);
      size_t chars = 0;
      for (unsigned char c : suffix) chars += (c & 0xC0) != 0x80;
      if (first > 0 && !prefix_literal && chars >= 3 && !re->nfa.is_anchored_start && !re->nfa.is_anchored_end) re->lcs = suffix;
    }
  }
  return re.release();
}

}  // namespace oracle

// ------------------------------------------------------------------ C entry --
extern "C" {

void* oracle_compile(const char* const* pats, const size_t* lens, size_t n, uint32_t flags, int only_utf8,
                     size_t size_limit) {
  std::vector<std::string> v;
  for (size_t i = 0; i < n; i++) v.emplace_back(pats[i], lens[i]);
  return oracle::build(v, flags, only_utf8 != 0, size_limit);
}
void oracle_free(void* h) { delete (oracle::Regex*)h; }
const char* oracle_error(void* h) {
  auto* re = (oracle::Regex*)h;
  return re->error.empty() ? nullptr : re->error.c_str();
}
int oracle_needs_unicode_word_boundary(void* h) { return ((oracle::Regex*)h)->dfa.has_unicode_word_boundary; }

int oracle_find_at(void* h, int engine, const uint8_t* t, size_t n, size_t start, size_t* s, size_t* e) {
  return ((oracle::Regex*)h)->find_at(engine, t, n, start, s, e);
}
int oracle_shortest_match_at(void* h, int engine, const uint8_t* t, size_t n, size_t start, size_t* e) {
  return ((oracle::Regex*)h)->shortest_at(engine, t, n, start, e);
}
int oracle_is_match_at(void* h, int engine, const uint8_t* t, size_t n, size_t start) {
  size_t e;
  auto* re = (oracle::Regex*)h;
  if (re->dfa.n_patterns > 1) {  // exec.rs:427-468 uses the shortest DFA for sets too
    std::vector<bool> m(re->dfa.n_patterns, false);
    return re->many_at(engine, t, n, start, m);
  }
  return re->shortest_at(engine, t, n, start, &e);
}
// find_iter (re_trait.rs:197-220).  Writes up to cap spans; returns the total count.
size_t oracle_find_iter(void* h, int engine, const uint8_t* t, size_t n, size_t* spans, size_t cap) {
  auto* re = (oracle::Regex*)h;
  size_t last_end = 0, last_match = oracle::NONE, count = 0;
  for (;;) {
    if (last_end > n) break;
    size_t s, e;
    if (!re->find_at(engine, t, n, last_end, &s, &e)) break;
    if (s == e) {
      last_end = re->next_after_empty(t, n, e);
      if (e == last_match) continue;
    } else {
      last_end = e;
    }
    last_match = e;
    if (count < cap) { spans[2 * count] = s; spans[2 * count + 1] = e; }
    count++;
  }
  return count;
}
// RegexSet::matches (re_set.rs:184-213): out[i] = 0/1 per pattern; returns any.
int oracle_set_matches(void* h, int engine, const uint8_t* t, size_t n, size_t start, uint8_t* out) {
  auto* re = (oracle::Regex*)h;
  std::vector<bool> m(re->pats.size(), false);
  bool any = re->pats.empty() ? false : re->many_at(engine, t, n, start, m);
  for (size_t i = 0; i < m.size(); i++) out[i] = m[i];
  return any;
}
// Captures::get(i) for every group: slots[2i], slots[2i+1] (SIZE_MAX = the group did not take part).
// Returns the number of groups (incl. group 0) on a match, 0 otherwise.
size_t oracle_captures_at(void* h, const uint8_t* t, size_t n, size_t start, size_t* slots, size_t max_slots) {
  auto* re = (oracle::Regex*)h;
  std::vector<size_t> v;
  if (re->pats.size() != 1 || !re->captures_at(t, n, start, v)) return 0;
  for (size_t k = 0; k < v.size() && k < max_slots; k++) slots[k] = v[k];
  return v.size() / 2;
}
size_t oracle_num_patterns(void* h) { return ((oracle::Regex*)h)->pats.size(); }

// Timed CPU baseline helper: count find_iter matches over `n` bytes split into
// `threads` byte ranges cut at '\n' (exact for patterns that cannot match '\n').
// Each thread owns a private compiled copy (the reference keeps per-thread DFA
// caches, exec.rs:1066-1072).
size_t oracle_count_parallel(const char* pat, size_t pat_len, uint32_t flags, int only_utf8, const uint8_t* t, size_t n,
                             int threads) {
  std::vector<size_t> cuts{0};
  for (int k = 1; k < threads; k++) {
    size_t c = n / threads * k;
    while (c < n && t[c] != '\n') c++;
    if (c < n) c++;
    if (c > cuts.back()) cuts.push_back(c);
  }
  cuts.push_back(n);
  std::vector<size_t> counts(cuts.size() - 1, 0);
  std::vector<std::thread> pool;
  for (size_t k = 0; k + 1 < cuts.size(); k++) {
    pool.emplace_back([&, k] {
      const char* p[1] = {pat};
      size_t l[1] = {pat_len};
      void* h = oracle_compile(p, l, 1, flags, only_utf8, 10u << 20);
      counts[k] = oracle_find_iter(h, 2, t + cuts[k], cuts[k + 1] - cuts[k], nullptr, 0);  // the engine the reference would select
      oracle_free(h);
    });
  }
  for (auto& th : pool) th.join();
  size_t total = 0;
  for (size_t c : counts) total += c;
  return total;
}

}  // extern "C"
