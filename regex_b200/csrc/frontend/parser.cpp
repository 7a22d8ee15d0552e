// Regex syntax parser + class algebra.  See frontend.h for the reference map.
//
// Structure: a frame stack (one frame per open group) instead of the
// reference's single mixed expression/marker stack; the accepted language, the
// produced AST and the error classification follow
// regex-syntax/src/parser.rs:106-1038 and :1098-1293.
#include "frontend.h"

#include <algorithm>
#include <cstring>

namespace rb {

#include "unicode_data.inc"

// ------------------------------------------------------------ unicode ----
static int uni_find(const std::string& name) {
  int lo = 0, hi = (int)kUniNamed - 1;
  while (lo <= hi) {
    int mid = (lo + hi) / 2;
    int c = std::strcmp(kUniNames[mid], name.c_str());
    if (c == 0) return mid;
    if (c < 0) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}
static CharClass uni_table(int idx) {
  CharClass c;
  for (uint32_t i = kUniOffsets[idx]; i < kUniOffsets[idx + 1]; i++)
    c.push_back({kUniRanges[2 * i], kUniRanges[2 * i + 1]});
  return c;
}
static bool uni_contains(int idx, uint32_t cp) {
  uint32_t lo = kUniOffsets[idx], hi = kUniOffsets[idx + 1];
  while (lo < hi) {
    uint32_t mid = (lo + hi) / 2;
    if (kUniRanges[2 * mid + 1] < cp) lo = mid + 1;
    else if (kUniRanges[2 * mid] > cp) hi = mid;
    else return true;
  }
  return false;
}
bool is_unicode_whitespace(uint32_t c) {
  static const int ws = uni_find("White_Space");
  return uni_contains(ws, c);
}
bool is_unicode_word_char(uint32_t c) { return uni_contains((int)kUniNamed, c); }

static uint32_t inc_char(uint32_t c) {
  if (c == 0x10FFFF) return c;
  if (c == 0xD7FF) return 0xE000;
  return c + 1;
}
static uint32_t dec_char(uint32_t c) {
  if (c == 0) return 0;
  if (c == 0xE000) return 0xD7FF;
  return c - 1;
}

// ------------------------------------------------------ class algebra ----
// lib.rs:687-704 (sorted, non-overlapping, non-adjacent)
CharClass class_canonicalize(CharClass c) {
  std::sort(c.begin(), c.end());
  CharClass out;
  for (const auto& r : c) {
    if (!out.empty()) {
      auto& last = out.back();
      if (std::max(last.lo, r.lo) <= inc_char(std::min(last.hi, r.hi))) {
        last.lo = std::min(last.lo, r.lo);
        last.hi = std::max(last.hi, r.hi);
        continue;
      }
    }
    out.push_back(r);
  }
  return out;
}
// lib.rs:745-768
CharClass class_negate(CharClass c) {
  if (c.empty()) return {{0, 0x10FFFF}};
  c = class_canonicalize(std::move(c));
  CharClass inv;
  if (c[0].lo > 0) inv.push_back({0, dec_char(c[0].lo)});
  for (size_t i = 0; i + 1 < c.size(); i++) {
    uint32_t a = inc_char(c[i].hi), b = dec_char(c[i + 1].lo);
    inv.push_back(a <= b ? ClassRange{a, b} : ClassRange{b, a});
  }
  if (c.back().hi < 0x10FFFF) inv.push_back({inc_char(c.back().hi), 0x10FFFF});
  return inv;
}
// lib.rs:776-789, 871-914: r ∪ {fold partners of members of r}
CharClass class_case_fold(const CharClass& c) {
  CharClass out;
  for (const auto& r : c) {
    size_t lo = 0, hi = kFoldPairs;
    while (lo < hi) {
      size_t mid = (lo + hi) / 2;
      if (kFold[2 * mid] < r.lo) lo = mid + 1; else hi = mid;
    }
    for (size_t i = lo; i < kFoldPairs && kFold[2 * i] <= r.hi; i++)
      out.push_back({kFold[2 * i + 1], kFold[2 * i + 1]});
    out.push_back(r);
  }
  return class_canonicalize(std::move(out));
}
// lib.rs:709-739
static CharClass class_intersect(const CharClass& a, const CharClass& b) {
  CharClass out;
  size_t i = 0, j = 0;
  while (i < a.size() && j < b.size()) {
    uint32_t lo = std::max(a[i].lo, b[j].lo), hi = std::min(a[i].hi, b[j].hi);
    if (lo <= hi) out.push_back({lo, hi});
    if (a[i].hi < b[j].hi) i++; else j++;
  }
  return class_canonicalize(std::move(out));
}
ByteClass bclass_canonicalize(ByteClass c) {
  std::sort(c.begin(), c.end());
  ByteClass out;
  for (const auto& r : c) {
    if (!out.empty()) {
      auto& last = out.back();
      int mn = std::min(last.hi, r.hi);
      if (std::max(last.lo, r.lo) <= std::min(mn + 1, 255)) {
        last.lo = std::min(last.lo, r.lo);
        last.hi = std::max(last.hi, r.hi);
        continue;
      }
    }
    out.push_back(r);
  }
  return out;
}
// lib.rs:1062-1068, 1119-1133 (ASCII only)
ByteClass bclass_case_fold(const ByteClass& c) {
  ByteClass out;
  for (const auto& r : c) {
    out.push_back(r);
    if (std::max<int>(r.lo, 'a') <= std::min<int>(r.hi, 'z'))
      out.push_back({(uint8_t)(std::max<int>(r.lo, 'a') - 32), (uint8_t)(std::min<int>(r.hi, 'z') - 32)});
    if (std::max<int>(r.lo, 'A') <= std::min<int>(r.hi, 'Z'))
      out.push_back({(uint8_t)(std::max<int>(r.lo, 'A') + 32), (uint8_t)(std::min<int>(r.hi, 'Z') + 32)});
  }
  return bclass_canonicalize(std::move(out));
}
// lib.rs:669-674, 818-826
static ByteClass to_byte_class(const CharClass& c) {
  ByteClass out;
  for (const auto& r : c)
    if (r.lo <= 0xFF) out.push_back({(uint8_t)r.lo, (uint8_t)std::min<uint32_t>(r.hi, 0xFF)});
  return bclass_canonicalize(std::move(out));
}

// ------------------------------------------------------- Expr queries ----
bool Expr::can_repeat() const {  // lib.rs:411-423
  switch (kind) {
    case EK::Empty: case EK::Repeat: case EK::Concat: case EK::Alternate: return false;
    default: return true;
  }
}
static bool rep_matches_empty(const Expr& e) {
  switch (e.rep) {
    case RepKind::ZeroOrOne: case RepKind::ZeroOrMore: return true;
    case RepKind::OneOrMore: return false;
    case RepKind::Range: return e.rmin == 0;
  }
  return false;
}
bool Expr::is_anchored_start() const {  // lib.rs:518-529
  switch (kind) {
    case EK::Repeat: return !rep_matches_empty(*this) && es[0].is_anchored_start();
    case EK::Group: return es[0].is_anchored_start();
    case EK::Concat: return es[0].is_anchored_start();
    case EK::Alternate:
      for (const auto& e : es) if (!e.is_anchored_start()) return false;
      return true;
    case EK::StartText: return true;
    default: return false;
  }
}
bool Expr::is_anchored_end() const {  // lib.rs:548-559
  switch (kind) {
    case EK::Repeat: return !rep_matches_empty(*this) && es[0].is_anchored_end();
    case EK::Group: return es[0].is_anchored_end();
    case EK::Concat: return es.back().is_anchored_end();
    case EK::Alternate:
      for (const auto& e : es) if (!e.is_anchored_end()) return false;
      return true;
    case EK::EndText: return true;
    default: return false;
  }
}
bool Expr::has_bytes() const {  // lib.rs:578-590
  switch (kind) {
    case EK::Repeat: case EK::Group: return es[0].has_bytes();
    case EK::Concat: case EK::Alternate:
      for (const auto& e : es) if (e.has_bytes()) return true;
      return false;
    case EK::LiteralBytes: case EK::AnyByte: case EK::AnyByteNoNL: case EK::ClassBytes:
    case EK::WordBoundaryAscii: case EK::NotWordBoundaryAscii: return true;
    default: return false;
  }
}

// ------------------------------------------------------------- parser ----
namespace {

bool is_punct(uint32_t c) {  // parser.rs:1378-1384
  switch (c) {
    case '\\': case '.': case '+': case '*': case '?': case '(': case ')': case '|':
    case '[': case ']': case '{': case '}': case '^': case '$': case '#': case '&':
    case '-': case '~': return true;
    default: return false;
  }
}
bool is_ascii_word(uint32_t c) {
  return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_' || (c >= '0' && c <= '9');
}
void append_utf8(std::string& s, uint32_t c) {
  if (c < 0x80) s += (char)c;
  else if (c < 0x800) { s += (char)(0xC0 | (c >> 6)); s += (char)(0x80 | (c & 0x3F)); }
  else if (c < 0x10000) { s += (char)(0xE0 | (c >> 12)); s += (char)(0x80 | ((c >> 6) & 0x3F)); s += (char)(0x80 | (c & 0x3F)); }
  else { s += (char)(0xF0 | (c >> 18)); s += (char)(0x80 | ((c >> 12) & 0x3F)); s += (char)(0x80 | ((c >> 6) & 0x3F)); s += (char)(0x80 | (c & 0x3F)); }
}
bool decode_utf8_string(const std::string& s, std::vector<uint32_t>* out) {
  size_t i = 0, n = s.size();
  while (i < n) {
    uint8_t b = (uint8_t)s[i];
    uint32_t c; int len;
    if (b < 0x80) { c = b; len = 1; }
    else if (b >= 0xC2 && b <= 0xDF) { c = b & 0x1F; len = 2; }
    else if (b >= 0xE0 && b <= 0xEF) { c = b & 0x0F; len = 3; }
    else if (b >= 0xF0 && b <= 0xF4) { c = b & 0x07; len = 4; }
    else return false;
    if (i + len > n) return false;
    for (int k = 1; k < len; k++) {
      uint8_t cb = (uint8_t)s[i + k];
      if ((cb & 0xC0) != 0x80) return false;
      c = (c << 6) | (cb & 0x3F);
    }
    if ((len == 3 && c < 0x800) || (len == 4 && c < 0x10000) || c > 0x10FFFF || (c >= 0xD800 && c <= 0xDFFF)) return false;
    out->push_back(c);
    i += len;
  }
  return true;
}

struct ParseFail {};

struct AsciiClassDef { const char* name; ClassRange r[6]; int n; };
const AsciiClassDef kAscii[] = {  // parser.rs:1443-1457
  {"alnum", {{'0','9'},{'A','Z'},{'a','z'}}, 3},
  {"alpha", {{'A','Z'},{'a','z'}}, 2},
  {"ascii", {{0,0x7F}}, 1},
  {"blank", {{' ',' '},{'\t','\t'}}, 2},
  {"cntrl", {{0,0x1F},{0x7F,0x7F}}, 2},
  {"digit", {{'0','9'}}, 1},
  {"graph", {{'!','~'}}, 1},
  {"lower", {{'a','z'}}, 1},
  {"print", {{' ','~'}}, 1},
  {"punct", {{'!','/'},{':','@'},{'[','`'},{'{','~'}}, 4},
  {"space", {{'\t','\t'},{'\n','\n'},{0x0B,0x0B},{0x0C,0x0C},{'\r','\r'},{' ',' '}}, 6},
  {"upper", {{'A','Z'}}, 1},
  {"word", {{'0','9'},{'A','Z'},{'_','_'},{'a','z'}}, 4},
  {"xdigit", {{'0','9'},{'A','F'},{'a','f'}}, 3},
};
bool ascii_class(const std::string& name, CharClass* out) {
  for (const auto& d : kAscii)
    if (name == d.name) { out->assign(d.r, d.r + d.n); return true; }
  return false;
}

struct Frame {
  std::vector<Expr> alts;
  std::vector<Expr> concat;
  bool saw_alt = false;
  int cap = 0;
  std::string name;
  size_t chari = 0;
  Flags old_flags;
};

struct Parser {
  std::vector<uint32_t> ch;
  size_t i = 0;
  Flags flags;
  int caps = 0;
  std::vector<std::string> names;
  std::vector<Frame> frames;
  Error* err;

  bool eof() const { return i >= ch.size(); }
  uint32_t cur() const { return ch[i]; }
  uint32_t bump() { return ch[i++]; }
  bool peek_is(uint32_t c) const { return !eof() && ch[i] == c; }
  bool bump_if(uint32_t c) { if (peek_is(c)) { i++; return true; } return false; }
  bool peek_str(const char* s) const {
    size_t k = 0;
    for (; s[k]; k++) if (i + k >= ch.size() || ch[i + k] != (uint32_t)(uint8_t)s[k]) return false;
    return true;
  }
  bool bump_if_str(const char* s) { if (peek_str(s)) { i += std::strlen(s); return true; } return false; }

  [[noreturn]] void fail_at(size_t pos, const std::string& what) {
    size_t s = std::max<size_t>(5, pos) - 5, e = std::min(ch.size(), pos + 5);
    std::string sur;
    for (size_t k = s; k < e; k++) append_utf8(sur, ch[k]);
    err->kind = Error::Syntax;
    err->msg = "Error parsing regex near '" + sur + "' at character offset " + std::to_string(pos) + ": " + what;
    throw ParseFail{};
  }
  [[noreturn]] void fail(const std::string& what) { fail_at(i, what); }

  void ignore_space() {  // parser.rs:894-916
    if (!flags.ignore_space) return;
    while (!eof()) {
      uint32_t c = cur();
      if (c == '#') {
        i++;
        while (!eof()) if (bump() == '\n') break;
      } else if (is_unicode_whitespace(c)) {
        i++;
      } else {
        return;
      }
    }
  }

  uint8_t one_byte(uint32_t c) {  // parser.rs:993-1000
    if (c > 0x7F) fail("Unicode features are not allowed when the Unicode (u) flag is not set.");
    return (uint8_t)c;
  }
  Expr lit(uint32_t c) {  // parser.rs:1025-1037
    Expr e;
    e.casei = flags.casei;
    if (flags.unicode) { e.kind = EK::Literal; e.chars.push_back(c); }
    else { e.kind = EK::LiteralBytes; e.bytes.push_back(one_byte(c)); }
    return e;
  }
  Expr u32_to_one_byte(uint32_t b) {  // parser.rs:1008-1020
    if (b > 0xFF) fail("Unicode features are not allowed when the Unicode (u) flag is not set.");
    if (!flags.allow_bytes && b > 0x7F) fail("Matching arbitrary bytes is not allowed.");
    Expr e;
    e.kind = EK::LiteralBytes;
    e.casei = flags.casei;
    e.bytes.push_back((uint8_t)b);
    return e;
  }
  CharClass class_transform(bool negate, CharClass cls) {  // parser.rs:979-987
    if (flags.casei) cls = class_case_fold(cls);
    if (negate) cls = class_negate(std::move(cls));
    return cls;
  }
  static Expr simple(EK k) { Expr e; e.kind = k; return e; }
  static Expr class_expr(CharClass c) { Expr e; e.kind = EK::Class; e.cls = std::move(c); return e; }

  std::string take_while_word() {
    std::string s;
    while (!eof() && is_ascii_word(cur())) s += (char)bump();
    return s;
  }

  // parser.rs:194-242.  Positioned on the backslash.
  Expr parse_escape() {
    i++;
    if (eof()) fail("Started an escape sequence that didn't finish before the end of the regex.");
    uint32_t c = cur();
    if (is_punct(c) || (flags.ignore_space && is_unicode_whitespace(c))) { i++; return lit(c); }
    switch (c) {
      case 'a': i++; return lit(0x07);
      case 'f': i++; return lit(0x0C);
      case 't': i++; return lit('\t');
      case 'n': i++; return lit('\n');
      case 'r': i++; return lit('\r');
      case 'v': i++; return lit(0x0B);
      case 'A': i++; return simple(EK::StartText);
      case 'z': i++; return simple(EK::EndText);
      case 'b': i++; return simple(flags.unicode ? EK::WordBoundary : EK::WordBoundaryAscii);
      case 'B': i++; return simple(flags.unicode ? EK::NotWordBoundary : EK::NotWordBoundaryAscii);
      case '0': case '1': case '2': case '3': case '4': case '5': case '6': case '7': {
        uint32_t n = 0;  // up to three octal digits (parser.rs:469-488)
        for (int k = 0; k < 3 && !eof() && cur() >= '0' && cur() <= '7'; k++) n = n * 8 + (bump() - '0');
        if (!flags.unicode) return u32_to_one_byte(n);
        return lit(n);
      }
      case 'x': i++; return parse_hex();
      case 'p': case 'P': i++; return class_expr(parse_unicode_class(c == 'P'));
      case 'd': case 's': case 'w': case 'D': case 'S': case 'W': i++; return class_expr(parse_perl_class(c));
      default: {
        std::string s = "Unrecognized escape sequence: '\\";
        append_utf8(s, c);
        fail(s + "'.");
      }
    }
  }
  static bool parse_radix(const std::string& s, int radix, uint32_t* out) {
    if (s.empty()) return false;
    uint64_t v = 0;
    for (char c : s) {
      int d;
      if (c >= '0' && c <= '9') d = c - '0';
      else if (c >= 'a' && c <= 'z') d = c - 'a' + 10;
      else if (c >= 'A' && c <= 'Z') d = c - 'A' + 10;
      else return false;
      if (d >= radix) return false;
      v = v * radix + d;
      if (v > 0xFFFFFFFFull) return false;
    }
    *out = (uint32_t)v;
    return true;
  }
  Expr parse_hex() {  // parser.rs:499-553
    ignore_space();
    uint32_t n = 0;
    if (bump_if('{')) {
      ignore_space();
      std::string s = take_while_word();
      if (!parse_radix(s, 16, &n)) fail("Not a valid base 16 number: '" + s + "'");
      ignore_space();
      if (!bump_if('}')) fail("Unclosed hexadecimal literal (missing a '}').");
      if (!flags.unicode) return u32_to_one_byte(n);
      if (n > 0x10FFFF || (n >= 0xD800 && n <= 0xDFFF))
        fail("Number does not correspond to a Unicode scalar value: '" + std::to_string(n) + "'.");
      return lit(n);
    }
    std::string s;
    size_t taken = 0;
    while (taken < 2 && !eof()) { append_utf8(s, bump()); taken++; }
    if (s.size() < 2) fail("Unexpected end of two digit hexadecimal literal.");
    if (!parse_radix(s, 16, &n)) fail("Not a valid base 16 number: '" + s + "'");
    if (!flags.unicode) return u32_to_one_byte(n);
    return lit(n);
  }
  CharClass parse_unicode_class(bool neg) {  // parser.rs:821-850
    ignore_space();
    std::string name;
    if (bump_if('{')) {
      ignore_space();
      name = take_while_word();
      ignore_space();
      if (name.empty() || !bump_if('}')) fail("Unclosed Unicode literal (missing a '}').");
    } else {
      if (eof()) fail("Started an escape sequence that didn't finish before the end of the regex.");
      append_utf8(name, bump());
    }
    int idx = uni_find(name);
    if (idx < 0) fail("Unrecognized Unicode class name: '" + name + "'.");
    if (!flags.unicode) fail("Unicode features are not allowed when the Unicode (u) flag is not set.");
    return class_transform(neg, uni_table(idx));
  }
  CharClass parse_perl_class(uint32_t name) {  // parser.rs:857-875
    static const int nd = uni_find("Nd"), ws = uni_find("White_Space");
    CharClass cls;
    bool negate = (name == 'D' || name == 'S' || name == 'W');
    uint32_t low = name | 0x20;
    if (flags.unicode) {
      cls = uni_table(low == 'd' ? nd : low == 's' ? ws : (int)kUniNamed);
    } else {
      ascii_class(low == 'd' ? "digit" : low == 's' ? "space" : "word", &cls);
    }
    return class_transform(negate, std::move(cls));
  }

  // ---- classes (parser.rs:562-876, 1249-1292) ----
  struct Bracket { enum K { Left, Set, Inter } k; bool negated = false; CharClass set; };

  void open_bracket(std::vector<Bracket>& st) {  // parser.rs:653-677
    i++;
    ignore_space();
    bool negated = bump_if('^');
    ignore_space();
    CharClass c;
    while (bump_if('-')) { c.push_back({'-', '-'}); ignore_space(); }
    if (c.empty() && bump_if(']')) { c.push_back({']', ']'}); ignore_space(); }
    Bracket b; b.k = Bracket::Left; b.negated = negated;
    st.push_back(b);
    if (!c.empty()) { Bracket s; s.k = Bracket::Set; s.set = c; st.push_back(s); }
  }
  bool maybe_parse_ascii(CharClass* out) {  // parser.rs:791-808
    size_t save = i;
    i++;
    if (bump_if(':')) {
      bool negate = bump_if('^');
      std::string name;
      while (!eof() && cur() != ':') append_utf8(name, bump());
      if (!name.empty() && bump_if_str(":]")) {
        CharClass cls;
        if (ascii_class(name, &cls)) { *out = class_transform(negate, cls); return true; }
      }
    }
    i = save;
    return false;
  }
  CharClass union_transform(CharClass c) {
    return flags.casei ? class_case_fold(c) : class_canonicalize(std::move(c));
  }
  CharClass close_bracket(std::vector<Bracket>& st) {
    CharClass uni;
    std::vector<CharClass> inter;
    for (;;) {
      Bracket b = std::move(st.back());
      st.pop_back();
      if (b.k == Bracket::Set) {
        uni.insert(uni.end(), b.set.begin(), b.set.end());
      } else if (b.k == Bracket::Inter) {
        inter.push_back(union_transform(std::move(uni)));
        uni.clear();
      } else {
        CharClass cls = union_transform(std::move(uni));
        for (const auto& c : inter) cls = class_intersect(cls, c);
        if (b.negated) cls = class_negate(std::move(cls));
        if (cls.empty()) fail("Empty character classes are not allowed.");
        return cls;
      }
    }
  }
  CharClass literal_of(const Expr& e, bool* ok) {
    *ok = true;
    if (e.kind == EK::Literal) return {{e.chars[0], e.chars[0]}};
    if (e.kind == EK::LiteralBytes) return {{e.bytes[0], e.bytes[0]}};
    *ok = false;
    return {};
  }
  CharClass parse_class_range(uint32_t start) {  // parser.rs:724-776
    ignore_space();
    if (!bump_if('-')) return {{start, start}};
    ignore_space();
    if (eof()) fail("Character class was not closed before the end of the regex (missing a ']').");
    if (peek_is(']')) return {{start, start}, {'-', '-'}};
    uint32_t end;
    if (cur() == '\\') {
      Expr e = parse_escape();
      bool ok;
      CharClass l = literal_of(e, &ok);
      if (!ok) fail("Invalid escape sequence in character class.");
      end = l[0].lo;
    } else {
      end = bump();
      if (end == '-') fail("Use of unescaped '-' in character class is not allowed.");
      if (!flags.unicode) one_byte(end);
    }
    if (end < start) {
      std::string s = "Invalid character class range '", a, b;
      append_utf8(a, start); append_utf8(b, end);
      fail(s + a + "-" + b + "'. Character class ranges must start with the smaller character, but " + a + " > " + b);
    }
    return {{start, end}};
  }
  CharClass parse_class_as_chars() {  // parser.rs:586-643
    std::vector<Bracket> st;
    open_bracket(st);
    for (;;) {
      ignore_space();
      if (eof()) fail("Character class was not closed before the end of the regex (missing a ']').");
      uint32_t c = cur();
      Bracket s; s.k = Bracket::Set;
      if (c == '[') {
        if (maybe_parse_ascii(&s.set)) st.push_back(std::move(s));
        else open_bracket(st);
      } else if (c == ']') {
        i++;
        CharClass cls = close_bracket(st);
        if (st.empty()) return cls;
        s.set = std::move(cls);
        st.push_back(std::move(s));
      } else if (c == '\\') {
        Expr e = parse_escape();
        if (e.kind == EK::Class) {
          s.set = std::move(e.cls);
        } else {
          bool ok;
          CharClass l = literal_of(e, &ok);
          if (!ok) fail("Invalid escape sequence in character class.");
          s.set = parse_class_range(l[0].lo);
        }
        st.push_back(std::move(s));
      } else if (c == '&' && peek_str("&&")) {
        i += 2;
        Bracket b; b.k = Bracket::Inter;
        st.push_back(b);
      } else {
        if (!flags.unicode) one_byte(c);
        i++;
        if ((c == '~' || c == '-') && peek_is(c)) {
          std::string m = "Use of unescaped '";
          m += (char)c;
          fail(m + "' in character class is not allowed.");
        }
        s.set = parse_class_range(c);
        st.push_back(std::move(s));
      }
    }
  }
  Expr parse_class() {  // parser.rs:562-580
    CharClass cls = parse_class_as_chars();
    if (flags.unicode) return class_expr(std::move(cls));
    ByteClass bc = to_byte_class(cls);
    if (bc.empty()) fail("Empty character classes are not allowed.");
    Expr e; e.kind = EK::ClassBytes; e.bcls = std::move(bc);
    return e;
  }

  // ---- groups / repeats ----
  static Expr make_concat(std::vector<Expr>& es) {
    if (es.empty()) return Expr{};
    if (es.size() == 1) return std::move(es[0]);
    Expr e; e.kind = EK::Concat; e.es = std::move(es);
    return e;
  }
  // '(' ... parser.rs:253-352
  void parse_group() {
    size_t chari = i;
    i++;
    ignore_space();
    Frame f;
    f.chari = chari;
    f.old_flags = flags;
    if (bump_if_str("?P<")) {
      std::string name;
      while (!eof() && !peek_is('>')) append_utf8(name, bump());
      if (eof()) fail("Capture name group for '" + name + "' is not closed. (Missing a '>'.)");
      if (name.empty()) fail("Capture names cannot be empty.");
      bool valid = !(name[0] >= '0' && name[0] <= '9');
      for (unsigned char c : name) if (!is_ascii_word(c)) valid = false;
      if (!valid)
        fail("Invalid capture name: '" + name + "'. Capture names must consist of [_a-zA-Z0-9] and are not allowed to start with with a number.");
      i++;  // '>'
      for (const auto& n : names) if (n == name) fail("Capture name '" + name + "' is used more than once.");
      names.push_back(name);
      f.name = name;
      f.cap = ++caps;
    } else if (bump_if('?')) {
      bool sign = true, saw_flag = false;
      for (;;) {
        if (eof()) fail("Inline flag settings was not closed before the end of the regex (missing a ')' or ':').");
        uint32_t c = cur();
        switch (c) {
          case 'i': flags.casei = sign; saw_flag = true; break;
          case 'm': flags.multi = sign; saw_flag = true; break;
          case 's': flags.dotnl = sign; saw_flag = true; break;
          case 'U': flags.swap_greed = sign; saw_flag = true; break;
          case 'x': flags.ignore_space = sign; saw_flag = true; break;
          case 'u': flags.unicode = sign; saw_flag = true; break;
          case '-':
            if (!sign) fail("Only one negation symbol is allowed in flags.");
            sign = false; saw_flag = false; break;
          case ')':
            if (!saw_flag) fail("Flag negation requires setting at least one flag.");
            i++;
            return;  // flags now apply to the rest of the enclosing group
          case ':':
            if (!sign && !saw_flag) fail("Flag negation requires setting at least one flag.");
            i++;
            frames.push_back(std::move(f));
            return;
          default: {
            std::string m = "Unrecognized flag: '";
            append_utf8(m, c);
            fail(m + "'. (Allowed flags: i, m, s, U, u, x.)");
          }
        }
        i++;
      }
    } else {
      f.cap = ++caps;
    }
    frames.push_back(std::move(f));
  }
  void close_group() {  // parser.rs:1153-1192
    if (frames.size() <= 1) fail("Unopened parenthesis.");
    Frame f = std::move(frames.back());
    frames.pop_back();
    if (f.concat.empty()) {
      if (f.saw_alt) fail("Alternations cannot be empty.");
      fail("Empty regex groups (e.g., '()') are not allowed.");
    }
    Expr inner;
    if (f.saw_alt) {
      f.alts.push_back(make_concat(f.concat));
      inner.kind = EK::Alternate;
      inner.es = std::move(f.alts);
    } else {
      inner = make_concat(f.concat);
    }
    Expr g;
    g.kind = EK::Group;
    g.cap = f.cap;
    g.name = f.name;
    g.es.push_back(std::move(inner));
    i++;
    flags = f.old_flags;
    frames.back().concat.push_back(std::move(g));
  }
  void alternate() {  // parser.rs:1098-1129
    Frame& f = frames.back();
    if (f.concat.empty()) fail("Alternations cannot be empty.");
    f.alts.push_back(make_concat(f.concat));
    f.concat.clear();
    f.saw_alt = true;
    i++;
  }
  Expr pop_for_repeat() {
    Frame& f = frames.back();
    if (f.concat.empty()) {
      // reference pops the pending Alternate and rejects it as unrepeatable;
      // with nothing at all on the stack it reports a missing expression.
      if (f.saw_alt) fail("Invalid application of repetition operator.");
      fail("Missing expression for repetition operator.");
    }
    if (!f.concat.back().can_repeat()) fail("Invalid application of repetition operator.");
    Expr e = std::move(f.concat.back());
    f.concat.pop_back();
    return e;
  }
  void push_repeat(Expr sub, RepKind k, uint32_t mn, uint32_t mx, bool has_max) {
    Expr r;
    r.kind = EK::Repeat;
    r.rep = k; r.rmin = mn; r.rmax = mx; r.has_max = has_max;
    r.greedy = (!bump_if('?')) ^ flags.swap_greed;
    r.es.push_back(std::move(sub));
    frames.back().concat.push_back(std::move(r));
  }
  uint32_t parse_decimal() {  // parser.rs:451-463
    std::string s;
    while (!eof() && (is_ascii_word(cur()) || is_unicode_whitespace(cur()))) append_utf8(s, bump());
    if (s.empty()) fail("Missing maximum in counted repetition operator.");
    size_t a = 0, b = s.size();
    auto ws = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); };
    while (a < b && ws(s[a])) a++;
    while (b > a && ws(s[b - 1])) b--;
    std::string t = s.substr(a, b - a);
    uint32_t n;
    if (!parse_radix(t, 10, &n)) fail("Not a valid base 10 number: '" + t + "'");
    return n;
  }
  void parse_counted_repeat() {  // parser.rs:387-424
    Expr e = pop_for_repeat();
    i++;
    ignore_space();
    uint32_t mn = parse_decimal(), mx = mn;
    bool has_max = true;
    ignore_space();
    if (bump_if(',')) {
      ignore_space();
      if (peek_is('}')) {
        has_max = false;
      } else {
        mx = parse_decimal();
        if (mn > mx)
          fail("Invalid counted repetition range: {" + std::to_string(mn) + ", " + std::to_string(mx) +
               "}. Counted repetition ranges must start with the minimum, but " + std::to_string(mn) + " > " + std::to_string(mx));
      }
    }
    ignore_space();
    if (!bump_if('}')) fail("Unclosed counted repetition (missing a '}').");
    push_repeat(std::move(e), RepKind::Range, mn, mx, has_max);
  }
  void simple_repeat(RepKind k) {
    Expr e = pop_for_repeat();
    i++;
    push_repeat(std::move(e), k, 0, 0, false);
  }
  void push(Expr e) { frames.back().concat.push_back(std::move(e)); }
  void one(EK k) { i++; push(simple(k)); }

  Expr run() {  // parser.rs:122-188
    frames.emplace_back();
    for (;;) {
      ignore_space();
      if (eof()) break;
      uint32_t c = cur();
      switch (c) {
        case '\\': push(parse_escape()); break;
        case '|': alternate(); break;
        case '?': simple_repeat(RepKind::ZeroOrOne); break;
        case '*': simple_repeat(RepKind::ZeroOrMore); break;
        case '+': simple_repeat(RepKind::OneOrMore); break;
        case '{': parse_counted_repeat(); break;
        case '[': push(parse_class()); break;
        case '^': one(flags.multi ? EK::StartLine : EK::StartText); break;
        case '$': one(flags.multi ? EK::EndLine : EK::EndText); break;
        case '.':
          if (!flags.unicode && !flags.allow_bytes) fail("Matching arbitrary bytes is not allowed.");
          if (flags.dotnl) one(flags.unicode ? EK::AnyChar : EK::AnyByte);
          else one(flags.unicode ? EK::AnyCharNoNL : EK::AnyByteNoNL);
          break;
        case '(': parse_group(); break;
        case ')': close_group(); break;
        default: i++; push(lit(c)); break;
      }
    }
    if (frames.size() > 1) fail_at(frames.back().chari, "Unclosed parenthesis.");
    Frame& f = frames.back();
    if (f.saw_alt) {
      if (f.concat.empty()) fail("Alternations cannot be empty.");
      f.alts.push_back(make_concat(f.concat));
      Expr a; a.kind = EK::Alternate; a.es = std::move(f.alts);
      return a;
    }
    return make_concat(f.concat);
  }
};

// lib.rs:425-500
Expr simplify(Expr e, size_t depth, size_t limit, Error* err) {
  if (depth > limit) {
    err->kind = Error::Syntax;
    err->msg = "Error parsing regex: Exhausted space required to parse regex with too much nesting.";
    throw ParseFail{};
  }
  switch (e.kind) {
    case EK::Repeat:
      e.es[0] = simplify(std::move(e.es[0]), depth + 1, limit, err);
      return e;
    case EK::Group: {
      Expr inner = simplify(std::move(e.es[0]), depth + 1, limit, err);
      if (e.cap == 0 && e.name.empty() && inner.can_repeat()) return inner;
      e.es[0] = std::move(inner);
      return e;
    }
    case EK::Concat: {
      std::vector<Expr> out;
      for (auto& sub : e.es) {
        Expr s = simplify(std::move(sub), depth + 1, limit, err);
        if (!out.empty() && out.back().kind == s.kind && out.back().casei == s.casei &&
            (s.kind == EK::Literal || s.kind == EK::LiteralBytes)) {
          out.back().chars.insert(out.back().chars.end(), s.chars.begin(), s.chars.end());
          out.back().bytes.insert(out.back().bytes.end(), s.bytes.begin(), s.bytes.end());
        } else {
          out.push_back(std::move(s));
        }
      }
      if (out.size() == 1) return std::move(out[0]);
      e.es = std::move(out);
      return e;
    }
    case EK::Alternate:
      for (auto& sub : e.es) sub = simplify(std::move(sub), depth + 1, limit, err);
      return e;
    default:
      return e;
  }
}

}  // namespace

bool parse(const std::string& pat, Flags flags, size_t nest_limit, Expr* out, Error* err) {
  Parser p;
  p.flags = flags;
  p.err = err;
  if (!decode_utf8_string(pat, &p.ch)) {
    err->kind = Error::InvalidUtf8Pattern;
    err->msg = "invalid utf-8 in pattern";
    return false;
  }
  try {
    *out = simplify(p.run(), 0, nest_limit, err);
  } catch (ParseFail&) {
    return false;
  }
  return true;
}

}  // namespace rb
