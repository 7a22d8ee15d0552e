// AST -> byte-level NFA program + byte classes.  See frontend.h.
//
// Follows src/compile.rs:124-1102 for everything that affects the matched
// language or thread priority: Split goto1-before-goto2 ordering
// (prog.rs:313-320), greedy/lazy orientation (compile.rs:574-578), the lazy
// any* prefix (compile.rs:381-395), reversed concatenations and swapped
// look-arounds for reverse programs (compile.rs:294-321, 352-357), UTF-8
// automata for Unicode classes (compile.rs:875-969) and the byte-class
// boundaries (compile.rs:1051-1102).  Instruction numbering differs from the
// reference (holes are patched through explicit lists and the UTF-8 suffix
// cache is a plain map); neither changes which threads exist nor their order.
#include "frontend.h"

#include <algorithm>
#include <cstdio>
#include <map>
#include <tuple>

namespace rb {

// ------------------------------------------------------ utf8 sequences ----
// Splits [lo,hi] (scalar values) into byte-range sequences.  Algorithm: the
// classic split-at-encoding-length-boundaries / split-at-continuation-byte-
// boundaries recursion (Russ Cox, RE2 "rune range"), which is what the
// utf8-ranges 1.x crate implements (reference dependency, Cargo.toml:35).
static int encode_utf8(uint32_t c, uint8_t* out) {
  if (c < 0x80) { out[0] = (uint8_t)c; return 1; }
  if (c < 0x800) { out[0] = 0xC0 | (c >> 6); out[1] = 0x80 | (c & 0x3F); return 2; }
  if (c < 0x10000) { out[0] = 0xE0 | (c >> 12); out[1] = 0x80 | ((c >> 6) & 0x3F); out[2] = 0x80 | (c & 0x3F); return 3; }
  out[0] = 0xF0 | (c >> 18); out[1] = 0x80 | ((c >> 12) & 0x3F); out[2] = 0x80 | ((c >> 6) & 0x3F); out[3] = 0x80 | (c & 0x3F);
  return 4;
}
static void utf8_split(uint32_t lo, uint32_t hi, std::vector<Utf8Seq>* out) {
  if (lo > hi) return;
  if (lo < 0xE000 && hi > 0xD7FF) {  // step around the surrogate gap
    utf8_split(lo, 0xD7FF, out);
    utf8_split(0xE000, hi, out);
    return;
  }
  static const uint32_t kMax[3] = {0x7F, 0x7FF, 0xFFFF};
  for (uint32_t m : kMax)
    if (lo <= m && m < hi) { utf8_split(lo, m, out); utf8_split(m + 1, hi, out); return; }
  if (hi < 0x80) { Utf8Seq s; s.n = 1; s.r[0] = {(uint8_t)lo, (uint8_t)hi}; out->push_back(s); return; }
  for (int i = 1; i < 4; i++) {
    uint32_t m = (1u << (6 * i)) - 1;
    if ((lo & ~m) != (hi & ~m)) {
      if ((lo & m) != 0) { utf8_split(lo, lo | m, out); utf8_split((lo | m) + 1, hi, out); return; }
      if ((hi & m) != m) { utf8_split(lo, (hi & ~m) - 1, out); utf8_split(hi & ~m, hi, out); return; }
    }
  }
  uint8_t a[4], b[4];
  int n = encode_utf8(lo, a);
  encode_utf8(hi, b);
  Utf8Seq s; s.n = n;
  for (int i = 0; i < n; i++) s.r[i] = {a[i], b[i]};
  out->push_back(s);
}
void utf8_sequences(uint32_t lo, uint32_t hi, std::vector<Utf8Seq>* out) { utf8_split(lo, hi, out); }

// ---------------------------------------------------------- match length ----
static int utf8_len(uint32_t c) { return c < 0x80 ? 1 : c < 0x800 ? 2 : c < 0x10000 ? 3 : 4; }
static uint64_t sat_add(uint64_t a, uint64_t b) { return (a == kUnbounded || b == kUnbounded || a + b < a) ? kUnbounded : a + b; }
static uint64_t sat_mul(uint64_t a, uint64_t b) {
  if (a == 0 || b == 0) return 0;
  if (a == kUnbounded || b == kUnbounded || a > (kUnbounded - 1) / b) return kUnbounded;
  return a * b;
}
static void class_len(const CharClass& c, uint64_t* mn, uint64_t* mx) {
  *mn = 4; *mx = 1;
  for (const auto& r : c) {
    *mn = std::min<uint64_t>(*mn, utf8_len(r.lo));
    *mx = std::max<uint64_t>(*mx, utf8_len(r.hi));
  }
}
void expr_len_range(const Expr& e, uint64_t* mn, uint64_t* mx) {
  switch (e.kind) {
    case EK::Empty: case EK::StartLine: case EK::EndLine: case EK::StartText: case EK::EndText:
    case EK::WordBoundary: case EK::NotWordBoundary: case EK::WordBoundaryAscii: case EK::NotWordBoundaryAscii:
      *mn = *mx = 0; return;
    case EK::Literal: {
      *mn = *mx = 0;
      for (uint32_t ch : e.chars) {
        CharClass cls{{ch, ch}};
        if (e.casei) cls = class_case_fold(cls);
        uint64_t a, b;
        class_len(cls, &a, &b);
        *mn += a; *mx += b;
      }
      return;
    }
    case EK::LiteralBytes: *mn = *mx = e.bytes.size(); return;
    case EK::AnyChar: case EK::AnyCharNoNL: *mn = 1; *mx = 4; return;
    case EK::AnyByte: case EK::AnyByteNoNL: case EK::ClassBytes: *mn = *mx = 1; return;
    case EK::Class: class_len(e.cls, mn, mx); return;
    case EK::Group: expr_len_range(e.es[0], mn, mx); return;
    case EK::Concat: {
      *mn = *mx = 0;
      for (const auto& s : e.es) {
        uint64_t a, b;
        expr_len_range(s, &a, &b);
        *mn = sat_add(*mn, a); *mx = sat_add(*mx, b);
      }
      return;
    }
    case EK::Alternate: {
      *mn = kUnbounded; *mx = 0;
      for (const auto& s : e.es) {
        uint64_t a, b;
        expr_len_range(s, &a, &b);
        *mn = std::min(*mn, a); *mx = std::max(*mx, b);
      }
      return;
    }
    case EK::Repeat: {
      uint64_t a, b;
      expr_len_range(e.es[0], &a, &b);
      uint64_t lo = 0, hi = kUnbounded;
      switch (e.rep) {
        case RepKind::ZeroOrOne: lo = 0; hi = 1; break;
        case RepKind::ZeroOrMore: lo = 0; break;
        case RepKind::OneOrMore: lo = 1; break;
        case RepKind::Range: lo = e.rmin; hi = e.has_max ? e.rmax : kUnbounded; break;
      }
      *mn = sat_mul(a, lo);
      *mx = (b == 0) ? 0 : sat_mul(b, hi);
      return;
    }
  }
}

// ------------------------------------------------------------ compiler ----
namespace {

constexpr uint32_t kUnset = 0xFFFFFFFFu;
// size_of::<Inst>() in the reference (prog.rs:253-256): used for size_limit only.
constexpr size_t kRefInstBytes = 40;

struct HoleRef { uint32_t pc; uint8_t which; };  // which: 0 = a, 1 = b
using Holes = std::vector<HoleRef>;
struct Patch { Holes holes; uint32_t entry; };
struct TooBig {};

struct Compiler {
  Program& p;
  const CompileOptions& opt;
  bool byte_set[256] = {false};
  std::map<std::tuple<uint32_t, uint8_t, uint8_t>, uint32_t> suffix_cache;

  Compiler(Program& prog, const CompileOptions& o) : p(prog), opt(o) {}

  uint32_t pc() const { return (uint32_t)p.insts.size(); }
  void check_size() {
    if (p.insts.size() * kRefInstBytes > opt.size_limit) throw TooBig{};
  }
  uint32_t emit(Inst in) { p.insts.push_back(in); return pc() - 1; }
  uint32_t emit_split() {
    Inst sp; sp.op = Op::Split; sp.a = kUnset; sp.b = kUnset;
    return emit(sp);
  }
  void fill(const Holes& hs, uint32_t target) {
    for (auto h : hs) (h.which ? p.insts[h.pc].b : p.insts[h.pc].a) = target;
  }
  static void append(Holes& dst, const Holes& src) { dst.insert(dst.end(), src.begin(), src.end()); }

  void set_range(uint8_t lo, uint8_t hi) {  // compile.rs:1058-1064
    if (lo > 0) byte_set[lo - 1] = true;
    byte_set[hi] = true;
  }
  void set_word_boundary() {  // compile.rs:1066-1080
    auto isw = [](int b) { return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_'; };
    int b1 = 0;
    while (b1 <= 255) {
      int b2 = b1 + 1;
      while (b2 <= 255 && isw(b1) == isw(b2)) b2++;
      set_range((uint8_t)b1, (uint8_t)(b2 - 1));
      b1 = b2;
    }
  }

  Patch c_bytes_inst(uint8_t lo, uint8_t hi) {
    set_range(lo, hi);
    Inst in; in.op = Op::Bytes; in.lo = lo; in.hi = hi; in.a = kUnset;
    uint32_t at = emit(in);
    return {{{at, 0}}, at};
  }
  // Alternation of single-byte ranges (compile.rs:472-496).
  Patch c_class_bytes(const ByteClass& ranges) {
    Patch out; out.entry = kUnset;
    uint32_t prev_split = kUnset;
    for (size_t k = 0; k < ranges.size(); k++) {
      if (k + 1 < ranges.size()) {
        uint32_t s = emit_split();
        if (prev_split != kUnset) p.insts[prev_split].b = s;
        if (out.entry == kUnset) out.entry = s;
        Patch b = c_bytes_inst(ranges[k].lo, ranges[k].hi);
        p.insts[s].a = b.entry;
        prev_split = s;
        append(out.holes, b.holes);
      } else {
        Patch b = c_bytes_inst(ranges[k].lo, ranges[k].hi);
        if (prev_split != kUnset) p.insts[prev_split].b = b.entry;
        if (out.entry == kUnset) out.entry = b.entry;
        append(out.holes, b.holes);
      }
    }
    return out;
  }
  // One UTF-8 sequence; forward programs emit the trailing byte first so that
  // common suffixes can be shared (compile.rs:924-968).
  Patch c_utf8_seq(const Utf8Seq& s) {
    uint32_t from = kUnset;
    Holes last;
    for (int k = 0; k < s.n; k++) {
      const ByteRange& r = p.is_reverse ? s.r[k] : s.r[s.n - 1 - k];
      auto key = std::make_tuple(from, r.lo, r.hi);
      auto it = suffix_cache.find(key);
      if (it != suffix_cache.end()) { from = it->second; continue; }
      set_range(r.lo, r.hi);
      Inst in; in.op = Op::Bytes; in.lo = r.lo; in.hi = r.hi; in.a = from;
      uint32_t at = emit(in);
      if (from == kUnset) last = {{at, 0}};
      suffix_cache[key] = at;
      from = at;
    }
    return {last, from};
  }
  // Unicode class -> alternation of UTF-8 sequences (compile.rs:880-922).
  Patch c_class(const CharClass& ranges) {
    std::vector<Utf8Seq> seqs;
    for (const auto& r : ranges) utf8_sequences(r.lo, r.hi, &seqs);
    suffix_cache.clear();
    Patch out; out.entry = kUnset;
    uint32_t prev_split = kUnset;
    for (size_t k = 0; k < seqs.size(); k++) {
      check_size();
      if (k + 1 < seqs.size()) {
        uint32_t s = emit_split();
        if (prev_split != kUnset) p.insts[prev_split].b = s;
        if (out.entry == kUnset) out.entry = s;
        Patch q = c_utf8_seq(seqs[k]);
        p.insts[s].a = q.entry;
        prev_split = s;
        append(out.holes, q.holes);
      } else {
        Patch q = c_utf8_seq(seqs[k]);
        if (prev_split != kUnset) p.insts[prev_split].b = q.entry;
        if (out.entry == kUnset) out.entry = q.entry;
        append(out.holes, q.holes);
      }
    }
    return out;
  }
  Patch c_look(Look l) {
    Inst in; in.op = Op::EmptyLook; in.look = l; in.a = kUnset;
    uint32_t at = emit(in);
    p.has_looks = true;
    return {{{at, 0}}, at};
  }
  Patch c_concat_n(const Expr& e, uint32_t n) {
    Patch out; out.entry = pc();
    bool first = true;
    for (uint32_t k = 0; k < n; k++) {
      Patch q = c(e);
      if (first) { out.entry = q.entry; first = false; }
      else fill(out.holes, q.entry);
      out.holes = std::move(q.holes);
    }
    return out;
  }
  Patch c_concat(const std::vector<const Expr*>& es) {
    Patch out; out.entry = pc();
    bool first = true;
    for (const Expr* e : es) {
      Patch q = c(*e);
      if (first) { out.entry = q.entry; first = false; }
      else fill(out.holes, q.entry);
      out.holes = std::move(q.holes);
    }
    return out;
  }
  // compile.rs:565-616.  Split.a is the preferred branch.
  Patch c_question(const Expr& e, bool greedy) {
    Inst sp; sp.op = Op::Split; sp.a = kUnset; sp.b = kUnset;
    uint32_t s = emit(sp);
    Patch q = c(e);
    Patch out; out.entry = s; out.holes = std::move(q.holes);
    if (greedy) { p.insts[s].a = q.entry; out.holes.push_back({s, 1}); }
    else { p.insts[s].b = q.entry; out.holes.push_back({s, 0}); }
    return out;
  }
  Patch c_star(const Expr& e, bool greedy) {
    Inst sp; sp.op = Op::Split; sp.a = kUnset; sp.b = kUnset;
    uint32_t s = emit(sp);
    Patch q = c(e);
    fill(q.holes, s);
    Patch out; out.entry = s;
    if (greedy) { p.insts[s].a = q.entry; out.holes.push_back({s, 1}); }
    else { p.insts[s].b = q.entry; out.holes.push_back({s, 0}); }
    return out;
  }
  Patch c_plus(const Expr& e, bool greedy) {
    Patch q = c(e);
    Inst sp; sp.op = Op::Split; sp.a = kUnset; sp.b = kUnset;
    uint32_t s = emit(sp);
    fill(q.holes, s);
    Patch out; out.entry = q.entry;
    if (greedy) { p.insts[s].a = q.entry; out.holes.push_back({s, 1}); }
    else { p.insts[s].b = q.entry; out.holes.push_back({s, 0}); }
    return out;
  }
  Patch c_repeat(const Expr& r) {
    const Expr& e = r.es[0];
    switch (r.rep) {
      case RepKind::ZeroOrOne: return c_question(e, r.greedy);
      case RepKind::ZeroOrMore: return c_star(e, r.greedy);
      case RepKind::OneOrMore: return c_plus(e, r.greedy);
      case RepKind::Range: break;
    }
    if (!r.has_max) {  // e{n,}  ==  e^n e*   (compile.rs:618-629)
      Patch head = c_concat_n(e, r.rmin);
      Patch tail = c_star(e, r.greedy);
      if (r.rmin == 0) return tail;
      fill(head.holes, tail.entry);
      return {tail.holes, head.entry};
    }
    // e{n,m} == e^n (e (e (...)?)?)?  with every optional exit leaving the
    // whole repetition (compile.rs:631-678).
    Patch head = c_concat_n(e, r.rmin);
    if (r.rmin == r.rmax) return head;
    Holes exits;
    Holes prev = head.holes;
    uint32_t entry = head.entry;
    bool have_entry = r.rmin > 0;
    for (uint32_t k = r.rmin; k < r.rmax; k++) {
      check_size();
      Inst sp; sp.op = Op::Split; sp.a = kUnset; sp.b = kUnset;
      uint32_t s = emit(sp);
      if (!have_entry) { entry = s; have_entry = true; }
      fill(prev, s);
      Patch q = c(e);
      prev = std::move(q.holes);
      if (r.greedy) { p.insts[s].a = q.entry; exits.push_back({s, 1}); }
      else { p.insts[s].b = q.entry; exits.push_back({s, 0}); }
    }
    append(exits, prev);
    return {exits, entry};
  }
  Patch c_alternate(const std::vector<Expr>& es) {  // compile.rs:521-544
    Patch out; out.entry = kUnset;
    uint32_t prev_split = kUnset;
    for (size_t k = 0; k < es.size(); k++) {
      if (k + 1 < es.size()) {
        uint32_t s = emit_split();
        if (prev_split != kUnset) p.insts[prev_split].b = s;
        if (out.entry == kUnset) out.entry = s;
        Patch q = c(es[k]);
        p.insts[s].a = q.entry;
        prev_split = s;
        append(out.holes, q.holes);
      } else {
        Patch q = c(es[k]);
        if (prev_split != kUnset) p.insts[prev_split].b = q.entry;
        if (out.entry == kUnset) out.entry = q.entry;
        append(out.holes, q.holes);
      }
    }
    return out;
  }

  Patch c(const Expr& e) {
    check_size();
    const bool rev = p.is_reverse;
    switch (e.kind) {
      case EK::Empty: return {{}, pc()};
      case EK::Literal: {
        Patch out; out.entry = pc();
        bool first = true;
        size_t n = e.chars.size();
        for (size_t k = 0; k < n; k++) {
          uint32_t ch = rev ? e.chars[n - 1 - k] : e.chars[k];
          CharClass cls{{ch, ch}};
          if (e.casei) cls = class_case_fold(cls);
          Patch q = c_class(cls);
          if (first) { out.entry = q.entry; first = false; } else fill(out.holes, q.entry);
          out.holes = std::move(q.holes);
        }
        return out;
      }
      case EK::LiteralBytes: {
        Patch out; out.entry = pc();
        bool first = true;
        size_t n = e.bytes.size();
        for (size_t k = 0; k < n; k++) {
          uint8_t b = rev ? e.bytes[n - 1 - k] : e.bytes[k];
          ByteClass cls{{b, b}};
          if (e.casei) cls = bclass_case_fold(cls);
          Patch q = c_class_bytes(cls);
          if (first) { out.entry = q.entry; first = false; } else fill(out.holes, q.entry);
          out.holes = std::move(q.holes);
        }
        return out;
      }
      case EK::AnyChar: return c_class({{0, 0x10FFFF}});
      case EK::AnyCharNoNL: return c_class({{0, 9}, {0xB, 0x10FFFF}});
      case EK::AnyByte: return c_class_bytes({{0, 0xFF}});
      case EK::AnyByteNoNL: return c_class_bytes({{0, 9}, {0xB, 0xFF}});
      case EK::Class: return c_class(e.cls);
      case EK::ClassBytes: return c_class_bytes(e.bcls);
      case EK::StartLine: set_range('\n', '\n'); return c_look(rev ? Look::EndLine : Look::StartLine);
      case EK::EndLine: set_range('\n', '\n'); return c_look(rev ? Look::StartLine : Look::EndLine);
      case EK::StartText: return c_look(rev ? Look::EndText : Look::StartText);
      case EK::EndText: return c_look(rev ? Look::StartText : Look::EndText);
      case EK::WordBoundary:
        p.has_unicode_word_boundary = true; set_word_boundary(); return c_look(Look::WordBoundary);
      case EK::NotWordBoundary:
        p.has_unicode_word_boundary = true; set_word_boundary(); return c_look(Look::NotWordBoundary);
      case EK::WordBoundaryAscii: set_word_boundary(); return c_look(Look::WordBoundaryAscii);
      case EK::NotWordBoundaryAscii: set_word_boundary(); return c_look(Look::NotWordBoundaryAscii);
      case EK::Group: {  // compile.rs:345-362: Save pairs only in the capture program (DFA programs drop them)
        if (!opt.saves || e.cap <= 0 || rev) return c(e.es[0]);
        p.n_captures = std::max(p.n_captures, e.cap + 1);
        Inst s0; s0.op = Op::Save; s0.a = kUnset; s0.b = 2 * (uint32_t)e.cap;
        const uint32_t a = emit(s0);
        Patch q = c(e.es[0]);
        p.insts[a].a = q.entry;
        Inst s1; s1.op = Op::Save; s1.a = kUnset; s1.b = 2 * (uint32_t)e.cap + 1;
        const uint32_t b = emit(s1);
        fill(q.holes, b);
        return {{{b, 0}}, a};
      }
      case EK::Concat: {
        std::vector<const Expr*> es;
        if (rev) for (size_t k = e.es.size(); k-- > 0;) es.push_back(&e.es[k]);
        else for (const auto& s : e.es) es.push_back(&s);
        return c_concat(es);
      }
      case EK::Alternate: return c_alternate(e.es);
      case EK::Repeat: return c_repeat(e);
    }
    return {{}, pc()};
  }

  Patch c_dotstar() {  // compile.rs:381-395
    Expr any; any.kind = opt.only_utf8 ? EK::AnyChar : EK::AnyByte;
    return c_star(any, /*greedy=*/false);
  }
  Patch c_top(const Expr& e) {  // group 0 Save pair only for the Pike VM oracle
    if (!opt.saves) return c(e);
    Inst s0; s0.op = Op::Save; s0.a = kUnset; s0.b = 0;  // a = goto, b = slot
    uint32_t a = emit(s0);
    Patch q = c(e);
    p.insts[a].a = q.entry;
    Inst s1; s1.op = Op::Save; s1.a = kUnset; s1.b = 1;
    uint32_t b = emit(s1);
    fill(q.holes, b);
    return {{{b, 0}}, a};
  }

  void finish() {  // compile.rs:1082-1101
    int cls = 0;
    for (int i = 0; i < 256; i++) {
      p.byte_classes[i] = (uint8_t)cls;
      if (i < 255 && byte_set[i]) cls++;
    }
    p.num_classes = p.byte_classes[255] + 1;
  }
};

}  // namespace

bool compile(const std::vector<Expr>& exprs, const CompileOptions& opt, Program* out, Error* err) {
  Program& p = *out;
  p = Program();
  p.only_utf8 = opt.only_utf8;
  p.is_reverse = opt.reverse;
  p.n_patterns = exprs.size();
  p.is_anchored_start = true;
  p.is_anchored_end = true;
  for (const auto& e : exprs) {
    p.is_anchored_start = p.is_anchored_start && e.is_anchored_start();
    p.is_anchored_end = p.is_anchored_end && e.is_anchored_end();
  }
  // forward programs anchor on ^, reverse programs never get the reference's
  // prefix (prog.rs:128-130); our reverse *unanchored* scan asks for one.
  bool anchored = opt.reverse ? false : p.is_anchored_start;
  p.has_prefix = opt.unanchored_prefix && !anchored;
  Compiler c(p, opt);
  try {
    Patch dot;
    if (p.has_prefix) {
      dot = c.c_dotstar();
      p.start = dot.entry;
    }
    if (exprs.size() == 1) {  // compile.rs:137-160
      Patch q = c.c_top(exprs[0]);
      p.start_anchored = q.entry;
      if (p.has_prefix) c.fill(dot.holes, q.entry); else p.start = q.entry;
      c.fill(q.holes, c.pc());
      p.matches.push_back(c.pc());
      Inst m; m.op = Op::Match; m.a = 0;
      c.emit(m);
    } else {  // split chain, one Match(i) per pattern (compile.rs:162-198)
      uint32_t prev_split = kUnset;
      for (size_t i = 0; i < exprs.size(); i++) {
        uint32_t s = kUnset;
        if (i + 1 < exprs.size()) {
          s = c.emit_split();
          if (prev_split != kUnset) p.insts[prev_split].b = s;
          if (i == 0) {
            p.start_anchored = s;
            if (p.has_prefix) c.fill(dot.holes, s); else p.start = s;
          }
        }
        Patch q = c.c(exprs[i]);
        if (s != kUnset) { p.insts[s].a = q.entry; prev_split = s; }
        else if (prev_split != kUnset) p.insts[prev_split].b = q.entry;
        c.fill(q.holes, c.pc());
        p.matches.push_back(c.pc());
        Inst m; m.op = Op::Match; m.a = (uint32_t)i;
        c.emit(m);
      }
    }
  } catch (TooBig&) {
    err->kind = Error::CompiledTooBig;
    err->msg = "Compiled regex exceeds size limit of " + std::to_string(opt.size_limit) + " bytes.";
    return false;
  }
  c.finish();
  return true;
}

std::string dump_program(const Program& p) {
  static const char* looks[] = {"StartLine", "EndLine", "StartText", "EndText", "WordBoundary",
                                "NotWordBoundary", "WordBoundaryAscii", "NotWordBoundaryAscii"};
  std::string s;
  char buf[128];
  for (size_t i = 0; i < p.insts.size(); i++) {
    const Inst& in = p.insts[i];
    switch (in.op) {
      case Op::Match: snprintf(buf, sizeof buf, "%04zu Match(%u)", i, in.a); break;
      case Op::Save: snprintf(buf, sizeof buf, "%04zu Save(%u) -> %u", i, in.b, in.a); break;
      case Op::Split: snprintf(buf, sizeof buf, "%04zu Split(%u, %u)", i, in.a, in.b); break;
      case Op::EmptyLook: snprintf(buf, sizeof buf, "%04zu %s -> %u", i, looks[(int)in.look], in.a); break;
      case Op::Bytes: snprintf(buf, sizeof buf, "%04zu Bytes(%02x, %02x) -> %u", i, in.lo, in.hi, in.a); break;
    }
    s += buf;
    if (i == p.start) s += " (start)";
    if (i == p.start_anchored) s += " (anchored start)";
    s += "\n";
  }
  return s;
}

}  // namespace rb
