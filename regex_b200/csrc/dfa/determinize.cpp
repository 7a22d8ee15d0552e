// Offline subset construction + minimisation.  See determinize.h.
#include "determinize.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <unordered_map>

namespace rb {

namespace {

struct EmptyFlags {  // dfa.rs:408-416
  bool start = false, end = false, start_line = false, end_line = false,
       word_boundary = false, not_word_boundary = false;
};

inline bool is_word_byte(int b) {
  return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_';
}
inline bool is_word_look(Look l) {
  return l == Look::WordBoundary || l == Look::NotWordBoundary || l == Look::WordBoundaryAscii ||
         l == Look::NotWordBoundaryAscii;
}

// Insertion-ordered set of instruction pointers (the reference's SparseSet,
// src/sparse.rs): order encodes thread priority.
struct OrderedSet {
  std::vector<uint32_t> dense;
  std::vector<uint32_t> sparse;
  explicit OrderedSet(size_t n) : sparse(n, 0) { dense.reserve(n); }
  bool contains(uint32_t v) const {
    uint32_t i = sparse[v];
    return i < dense.size() && dense[i] == v;
  }
  void insert(uint32_t v) { sparse[v] = (uint32_t)dense.size(); dense.push_back(v); }
  void clear() { dense.clear(); }
};

struct VecHash {
  size_t operator()(const std::vector<uint32_t>& v) const {
    uint64_t h = 1469598103934665603ull;
    for (uint32_t x : v) { h ^= x; h *= 1099511628211ull; }
    return (size_t)h;
  }
};

struct Builder {
  const Program& prog;
  const DfaOptions& opt;
  uint32_t n_byte_classes, n_classes, mask_words;
  int rep[256];  // representative byte of each byte class
  bool prog_has_word_looks = false;

  // raw states: key layout = [word, mask words (lo,hi pairs)..., ips...]
  std::unordered_map<std::vector<uint32_t>, uint32_t, VecHash> index;
  std::vector<std::vector<uint32_t>> keys;
  std::vector<uint32_t> trans;  // raw [state][n_classes], 0 = dead
  std::vector<uint32_t> stack;
  OrderedSet qcur, qnext;
  std::vector<uint64_t> ipbits;  // scratch for intern()
  bool too_big = false;

  Builder(const Program& p, const DfaOptions& o)
      : prog(p), opt(o), qcur(p.insts.size()), qnext(p.insts.size()) {
    n_byte_classes = (uint32_t)p.num_classes;
    n_classes = n_byte_classes + 1;
    mask_words = (uint32_t)((p.n_patterns + 63) / 64);
    for (int c = 0; c < 256; c++) rep[c] = -1;
    for (int b = 255; b >= 0; b--) rep[p.byte_classes[b]] = b;
    for (const auto& in : p.insts)
      if (in.op == Op::EmptyLook && is_word_look(in.look)) prog_has_word_looks = true;
    keys.emplace_back();  // state 0 = dead
    trans.assign(n_classes, 0);
  }

  size_t key_header() const { return 1 + 2 * (size_t)mask_words; }

  // dfa.rs:1073-1134
  void follow_epsilons(uint32_t ip0, OrderedSet& q, const EmptyFlags& f) {
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
            case Look::WordBoundary: case Look::WordBoundaryAscii: ok = f.word_boundary; break;
            case Look::NotWordBoundary: case Look::NotWordBoundaryAscii: ok = f.not_word_boundary; break;
          }
          if (ok) stack.push_back(in.a);
          break;
        }
      }
    }
  }

  // dfa.rs:1196-1244 (+ canonicalisation that cannot change results: the word
  // flag is kept only when a word-boundary look can still read it -- i.e. when the state
  // holds any look-around at all, since a word look may be reached through another look
  // at the next re-closure (`(?-u:\b)^`: the reverse program meets `^` first) -- and
  // priority order is dropped when no leftmost-first cut will ever use it).
  uint32_t intern(const OrderedSet& q, bool word, const std::vector<uint64_t>& mask) {
    std::vector<uint32_t> key(key_header(), 0);
    bool any_mask = false;
    for (uint32_t w = 0; w < mask_words; w++) {
      key[1 + 2 * w] = (uint32_t)mask[w];
      key[2 + 2 * w] = (uint32_t)(mask[w] >> 32);
      any_mask = any_mask || mask[w];
    }
    bool word_look = false;
    if (opt.leftmost_first) {
      for (uint32_t ip : q.dense) {
        const Inst& in = prog.insts[ip];
        if (in.op == Op::Save || in.op == Op::Split) continue;
        key.push_back(ip);
        if (in.op == Op::EmptyLook) word_look = prog_has_word_looks;  // a word look may sit behind this look
        if (in.op == Op::Match) break;
      }
    } else {
      // order is irrelevant without the leftmost-first cut: canonical form = ascending ips,
      // produced by a bitmap sweep (sorting thousands of ips per transition dominated \w+)
      if (ipbits.empty()) ipbits.assign((prog.insts.size() + 63) / 64, 0);
      for (uint32_t ip : q.dense) {
        const Inst& in = prog.insts[ip];
        if (in.op == Op::Save || in.op == Op::Split) continue;
        ipbits[ip >> 6] |= 1ull << (ip & 63);
        if (in.op == Op::EmptyLook) word_look = prog_has_word_looks;  // a word look may sit behind this look
      }
      for (size_t w = 0; w < ipbits.size(); w++) {
        uint64_t m = ipbits[w];
        ipbits[w] = 0;
        while (m) {
          key.push_back((uint32_t)(w * 64 + __builtin_ctzll(m)));
          m &= m - 1;
        }
      }
    }
    if (key.size() == key_header() && !any_mask) return 0;  // dead
    key[0] = (word && word_look) ? 1 : 0;
    auto it = index.find(key);
    if (it != index.end()) return it->second;
    uint32_t id = (uint32_t)keys.size();
    if (id > 400000 || (size_t)(id + 1) * n_classes * 2 > opt.max_table_bytes * 8) { too_big = true; return 0; }
    index.emplace(key, id);
    keys.push_back(std::move(key));
    trans.resize((size_t)(id + 1) * n_classes, 0xFFFFFFFFu);
    return id;
  }

  // dfa.rs:910-1048.  b in [0,255] or 256 for EOF.
  uint32_t step(uint32_t si, int b) {
    const std::vector<uint32_t> key = keys[si];  // copy: intern() may reallocate keys
    bool is_word_last = key[0] != 0;
    bool has_empty = false;
    qcur.clear();
    for (size_t k = key_header(); k < key.size(); k++) {
      qcur.insert(key[k]);
      if (prog.insts[key[k]].op == Op::EmptyLook) has_empty = true;
    }
    bool eof = b == 256;
    bool is_word = !eof && is_word_byte(b);
    if (has_empty) {
      EmptyFlags f;
      if (eof) { f.end = true; f.end_line = true; }
      else if (b == '\n') f.end_line = true;
      if (is_word_last == is_word) f.not_word_boundary = true; else f.word_boundary = true;
      qnext.clear();
      for (uint32_t ip : qcur.dense) follow_epsilons(ip, qnext, f);
      std::swap(qcur.dense, qnext.dense);
      std::swap(qcur.sparse, qnext.sparse);
    }
    EmptyFlags f;
    f.start_line = !eof && b == '\n';
    std::vector<uint64_t> mask(mask_words, 0);
    qnext.clear();
    for (uint32_t ip : qcur.dense) {
      const Inst& in = prog.insts[ip];
      if (in.op == Op::Match) {
        mask[in.a / 64] |= 1ull << (in.a % 64);
        if (opt.leftmost_first) break;
      } else if (in.op == Op::Bytes) {
        if (!eof && in.lo <= b && b <= in.hi) follow_epsilons(in.a, qnext, f);
      }
    }
    return intern(qnext, is_word, mask);
  }

  // All transitions of one raw state.  Same results, in the same order, as calling step()
  // for every class -- but in one pass per state: each Bytes instruction is appended to the classes its byte range covers
  // (instead of testing every instruction against every class), and classes that end up
  // with the same target list share one closure + intern (Unicode-aware classes have ~110
  // byte classes but a few dozen distinct successors).  \w+ compiled in 6.6 s before.
  std::vector<std::vector<uint32_t>> per_class;
  std::unordered_map<std::vector<uint32_t>, uint32_t, VecHash> memo;
  std::vector<uint32_t> memo_key;
  void step_all(uint32_t si, uint32_t* row) {
    const std::vector<uint32_t> key = keys[si];  // copy: intern() may reallocate keys
    const bool is_word_last = key[0] != 0;
    bool has_empty = false;
    for (size_t k = key_header(); k < key.size(); k++)
      if (prog.insts[key[k]].op == Op::EmptyLook) has_empty = true;
    // The ordered instruction list the byte is applied to.  With look-arounds in the state
    // it is first re-closed under the flags the byte implies (dfa.rs:933-957) -- and those
    // depend only on whether the byte is a newline / a word byte / the end of text, so
    // there are at most four variants: 0 = other byte, 1 = word byte, 2 = newline, 3 = EOF.
    auto variant_list = [&](int g, std::vector<uint32_t>* out) {
      out->clear();
      if (!has_empty) {
        out->assign(key.begin() + key_header(), key.end());
        return;
      }
      qcur.clear();
      for (size_t k = key_header(); k < key.size(); k++) qcur.insert(key[k]);
      EmptyFlags f;
      if (g == 3) { f.end = true; f.end_line = true; }
      else if (g == 2) f.end_line = true;
      const bool is_word = g == 1;
      if (is_word_last == is_word) f.not_word_boundary = true; else f.word_boundary = true;
      qnext.clear();
      for (uint32_t ip : qcur.dense) follow_epsilons(ip, qnext, f);
      *out = qnext.dense;
    };
    auto group_of = [&](uint32_t c) { const int b = rep[c]; return b == '\n' ? 2 : is_word_byte(b) ? 1 : 0; };
    per_class.resize(n_byte_classes);
    for (auto& v : per_class) v.clear();
    std::vector<uint64_t> mask[4];
    bool present[3] = {false, false, false};
    for (uint32_t c = 0; c < n_byte_classes; c++) present[group_of(c)] = true;
    std::vector<uint32_t> list;
    for (int g = 0; g < 4; g++) {
      mask[g].assign(mask_words, 0);
      if (g < 3 && !present[g]) continue;
      variant_list(g, &list);
      for (uint32_t ip : list) {
        const Inst& in = prog.insts[ip];
        if (in.op == Op::Match) {
          mask[g][in.a / 64] |= 1ull << (in.a % 64);
          if (opt.leftmost_first) break;
        } else if (in.op == Op::Bytes && g < 3) {
          for (uint32_t c = prog.byte_classes[in.lo]; c <= prog.byte_classes[in.hi]; c++)
            if (group_of(c) == g) per_class[c].push_back(in.a);
        }
      }
    }
    // classes of one variant with the same Bytes targets have the same successor; so do
    // equal (targets, variant, mask) combinations met in other states
    struct Seen { uint32_t cls; int g; uint32_t to; };
    std::vector<Seen> seen;
    for (uint32_t c = 0; c < n_byte_classes; c++) {
      const int g = group_of(c);
      uint32_t to = 0xFFFFFFFFu;
      for (const Seen& e : seen)
        if (e.g == g && per_class[e.cls] == per_class[c]) { to = e.to; break; }
      if (to == 0xFFFFFFFFu) {
        memo_key.clear();
        memo_key.push_back((uint32_t)g);
        for (uint64_t m : mask[g]) { memo_key.push_back((uint32_t)m); memo_key.push_back((uint32_t)(m >> 32)); }
        memo_key.insert(memo_key.end(), per_class[c].begin(), per_class[c].end());
        auto hit = memo.find(memo_key);
        if (hit != memo.end()) {
          to = hit->second;
        } else {
          EmptyFlags f;
          f.start_line = g == 2;
          qnext.clear();
          for (uint32_t target : per_class[c]) follow_epsilons(target, qnext, f);
          to = intern(qnext, g == 1, mask[g]);
          if (!too_big) memo.emplace(memo_key, to);
        }
        seen.push_back(Seen{c, g, to});
      }
      row[c] = to;
    }
    qnext.clear();
    row[n_byte_classes] = intern(qnext, false, mask[3]);  // EOF: no byte is consumed
  }

  uint32_t start_state(int flagi) {  // dfa.rs:1370-1409
    EmptyFlags f;
    f.start = flagi & 1; f.end = flagi & 2; f.start_line = flagi & 4; f.end_line = flagi & 8;
    f.word_boundary = flagi & 16; f.not_word_boundary = flagi & 32;
    bool word = flagi & 64;
    qnext.clear();
    uint32_t entry = (opt.anchored || !prog.has_prefix) ? prog.start_anchored : prog.start;
    follow_epsilons(entry, qnext, f);
    return intern(qnext, word, std::vector<uint64_t>(mask_words, 0));
  }
};

}  // namespace

int start_flag_index_forward(const uint8_t* text, size_t len, size_t at) {
  int f = 0;
  if (at == 0) f |= 1;
  if (len == 0) f |= 2 | 8;
  if (at == 0 || text[at - 1] == '\n') f |= 4;
  bool last = at > 0 && is_word_byte(text[at - 1]);
  bool cur = at < len && is_word_byte(text[at]);
  f |= (last == cur) ? 32 : 16;
  if (last) f |= 64;
  return f;
}
int start_flag_index_reverse(const uint8_t* text, size_t len, size_t at) {
  int f = 0;
  if (at == len) f |= 1;
  if (len == 0) f |= 2 | 8;
  if (at == len || text[at] == '\n') f |= 4;
  bool last = at < len && is_word_byte(text[at]);
  bool cur = at > 0 && is_word_byte(text[at - 1]);
  f |= (last == cur) ? 32 : 16;
  if (last) f |= 64;
  return f;
}

// The lazy DFA resolves `^`-type looks only in the closure taken right AFTER a byte and `$`-type
// and word looks only in the re-closure right BEFORE the next byte (dfa.rs:933-957, 971-999).
// Two adjacent look-arounds (no byte in between) in the "wrong" order never pass inside a scan:
//   word look, then multi-line `^`   -- forward scans miss it (the reference's DFA and NFA disagree);
//   multi-line `$`, then word look   -- the reverse program reads "word look, then start-of-line
//                                       look", so the start bitmap would silently miss matches.
// Neither occurs in the reference's tests or in practice (`\b(?m:^)`, `(?m:$)\b`); both are rejected
// instead of answered differently from the reference.
static bool has_unresolvable_look_pair(const Program& prog) {
  if (prog.is_reverse) return false;  // the forward programs of the same regex carry the check
  std::vector<uint32_t> stack;
  std::vector<uint8_t> seen;
  for (uint32_t ip0 = 0; ip0 < prog.insts.size(); ip0++) {
    const Inst& first = prog.insts[ip0];
    if (first.op != Op::EmptyLook) continue;
    const bool from_end_line = first.look == Look::EndLine, from_word = is_word_look(first.look);
    if (!from_end_line && !from_word) continue;
    seen.assign(prog.insts.size(), 0);
    stack.assign(1, first.a);
    while (!stack.empty()) {
      const uint32_t ip = stack.back();
      stack.pop_back();
      if (seen[ip]) continue;
      seen[ip] = 1;
      const Inst& in = prog.insts[ip];
      switch (in.op) {
        case Op::Save: stack.push_back(in.a); break;
        case Op::Split: stack.push_back(in.a); stack.push_back(in.b); break;
        case Op::EmptyLook:
          if (from_end_line && is_word_look(in.look)) return true;
          if (from_word && in.look == Look::StartLine) return true;
          stack.push_back(in.a);
          break;
        default: break;
      }
    }
  }
  return false;
}

bool determinize(const Program& prog, const DfaOptions& opt, Dfa* out, Error* err) {
  if (has_unresolvable_look_pair(prog)) {
    err->kind = Error::UnresolvableLookPair;  // its own class: not one of the two rejections the north star grants
    err->msg = "look-around sequence not supported by the B200 DFA backend: a word boundary directly followed by a "
               "multi-line `^`, or a multi-line `$` directly followed by a word boundary, cannot be resolved inside a "
               "DFA scan (the reference's own DFA and NFA engines disagree on such patterns).";
    return false;
  }
  if (prog.has_unicode_word_boundary) {
    err->kind = Error::UnicodeWordBoundary;
    err->msg = "Unicode word boundaries (\\b, \\B without (?-u)) need a look-around engine; "
               "the B200 DFA backend does not support them (use (?-u:\\b)).";
    return false;
  }
  Builder b(prog, opt);
  uint32_t raw_start[128];
  for (int f = 0; f < 128; f++) raw_start[f] = b.start_state(f);
  std::vector<uint32_t> row(b.n_classes);
  for (uint32_t si = 1; si < b.keys.size() && !b.too_big; si++) {
    b.step_all(si, row.data());
    std::copy(row.begin(), row.end(), b.trans.begin() + (size_t)si * b.n_classes);  // trans may have grown meanwhile
  }
  if (b.too_big) {
    err->kind = Error::DfaTooBig;
    err->msg = "Compiled regex exceeds size limit: determinized DFA is larger than the " +
               std::to_string(opt.max_table_bytes) + " byte table budget.";
    return false;
  }
  const uint32_t n = (uint32_t)b.keys.size(), nc = b.n_classes, mw = b.mask_words;

  // ---- minimise (Moore refinement), initial partition = delayed match mask ----
  std::vector<uint32_t> block(n, 0);
  uint32_t n_blocks;
  {
    std::unordered_map<std::vector<uint32_t>, uint32_t, VecHash> first;
    first.emplace(std::vector<uint32_t>(2 * mw + 1, 0xFFFFFFFFu), 0);  // dead alone
    for (uint32_t s = 1; s < n; s++) {
      std::vector<uint32_t> sig(b.keys[s].begin() + 1, b.keys[s].begin() + 1 + 2 * mw);
      auto it = first.emplace(sig, (uint32_t)first.size()).first;
      block[s] = it->second;
    }
    n_blocks = (uint32_t)first.size();
  }
  for (;;) {
    std::unordered_map<std::vector<uint32_t>, uint32_t, VecHash> sigs;
    std::vector<uint32_t> nb(n);
    std::vector<uint32_t> sig(nc + 1);
    for (uint32_t s = 0; s < n; s++) {
      sig[0] = block[s];
      for (uint32_t c = 0; c < nc; c++) sig[c + 1] = block[b.trans[(size_t)s * nc + c]];
      auto it = sigs.emplace(sig, (uint32_t)sigs.size()).first;
      nb[s] = it->second;
    }
    bool stable = sigs.size() == n_blocks;
    n_blocks = (uint32_t)sigs.size();
    block.swap(nb);
    if (stable) break;
  }
  // A live block that can never reach a match is equivalent to dead only if it
  // carries no mask itself; Moore already merged those with state 0.

  // ---- renumber: dead, live..., match... (BFS order inside each group) ----
  std::vector<uint32_t> repr(n_blocks, 0xFFFFFFFFu);
  for (uint32_t s = 0; s < n; s++) if (repr[block[s]] == 0xFFFFFFFFu) repr[block[s]] = s;
  auto has_mask = [&](uint32_t s) {
    for (uint32_t w = 0; w < 2 * mw; w++) if (b.keys[s].size() > 1 + w && b.keys[s][1 + w]) return true;
    return false;
  };
  std::vector<uint32_t> new_id(n_blocks, 0);
  uint32_t next_id = 0;
  new_id[block[0]] = next_id++;
  for (uint32_t bl = 0; bl < n_blocks; bl++)
    if (bl != block[0] && !has_mask(repr[bl])) new_id[bl] = next_id++;
  uint32_t match_lo = next_id;
  for (uint32_t bl = 0; bl < n_blocks; bl++)
    if (bl != block[0] && has_mask(repr[bl])) new_id[bl] = next_id++;
  if (n_blocks > 65535 || (size_t)n_blocks * nc * 2 > opt.max_table_bytes) {
    err->kind = Error::DfaTooBig;
    err->msg = "Compiled regex exceeds size limit: determinized DFA (" + std::to_string(n_blocks) +
               " states x " + std::to_string(nc) + " classes) is larger than the " +
               std::to_string(opt.max_table_bytes) + " byte table budget.";
    return false;
  }

  Dfa& d = *out;
  d = Dfa();
  d.n_states = n_blocks;
  d.n_classes = nc;
  d.match_lo = match_lo;
  d.mask_words = mw;
  d.raw_states = n;
  d.reverse = prog.is_reverse;
  d.has_looks = prog.has_looks;
  std::memcpy(d.classes, prog.byte_classes, 256);
  d.trans.assign((size_t)n_blocks * nc, 0);
  d.masks.assign((size_t)n_blocks * mw, 0);
  for (uint32_t bl = 0; bl < n_blocks; bl++) {
    uint32_t s = repr[bl], id = new_id[bl];
    for (uint32_t c = 0; c < nc; c++) d.trans[(size_t)id * nc + c] = (uint16_t)new_id[block[b.trans[(size_t)s * nc + c]]];
    if (s != 0)
      for (uint32_t w = 0; w < mw; w++)
        d.masks[(size_t)id * mw + w] = (uint64_t)b.keys[s][1 + 2 * w] | ((uint64_t)b.keys[s][2 + 2 * w] << 32);
  }
  d.uniform_start = true;
  for (int f = 0; f < 128; f++) d.start[f] = (uint16_t)new_id[block[raw_start[f]]];
  // Only flag combinations that can occur matter for uniformity: exactly one of
  // word-boundary / not-word-boundary is set, end implies end-line, start implies start-line.
  int first_feasible = -1;
  for (int f = 0; f < 128; f++) {
    bool wb = f & 16, nwb = f & 32;
    if (wb == nwb) continue;
    if (((f & 2) != 0) != ((f & 8) != 0)) continue;
    if ((f & 1) && !(f & 4)) continue;
    if (first_feasible < 0) first_feasible = f;
    else if (d.start[f] != d.start[first_feasible]) d.uniform_start = false;
  }
  return true;
}

std::string dump_dfa(const Dfa& d) {
  std::string s;
  char buf[96];
  snprintf(buf, sizeof buf, "states=%u classes=%u match_lo=%u raw=%zu uniform_start=%d\n", d.n_states,
           d.n_classes, d.match_lo, d.raw_states, (int)d.uniform_start);
  s += buf;
  for (uint32_t i = 0; i < d.n_states; i++) {
    snprintf(buf, sizeof buf, "%5u%s:", i, d.is_match((uint16_t)i) ? "*" : " ");
    s += buf;
    for (uint32_t c = 0; c < d.n_classes; c++) {
      snprintf(buf, sizeof buf, " %u", d.trans[(size_t)i * d.n_classes + c]);
      s += buf;
    }
    s += "\n";
  }
  return s;
}

}  // namespace rb
