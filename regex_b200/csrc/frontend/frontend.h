// Host front end: regex syntax -> AST -> byte-level NFA program + byte classes.
//
// Restates (in C++17, from scratch) the behaviour of the reference's front end so
// the determinizer sees the same language and the same thread priorities:
//   - syntax / flags / classes:   regex-syntax/src/parser.rs:106-1038, 1077-1293
//   - AST + simplification:       regex-syntax/src/lib.rs:97-179, 425-500
//   - class algebra, case fold:   regex-syntax/src/lib.rs:610-915, 951-1134
//   - NFA program + byte classes: src/compile.rs:124-1102, src/prog.rs:261-425
// The front end is shared by the product (determinizer input) and by the test
// oracle under oracle/ (Pike VM / lazy-DFA input); it is not on the GPU hot path.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rb {

// ---------------------------------------------------------------- AST ----
struct ClassRange {
  uint32_t lo, hi;  // inclusive Unicode scalar values
  bool operator<(const ClassRange& o) const { return lo != o.lo ? lo < o.lo : hi < o.hi; }
  bool operator==(const ClassRange& o) const { return lo == o.lo && hi == o.hi; }
};
using CharClass = std::vector<ClassRange>;
struct ByteRange {
  uint8_t lo, hi;
  bool operator<(const ByteRange& o) const { return lo != o.lo ? lo < o.lo : hi < o.hi; }
};
using ByteClass = std::vector<ByteRange>;

enum class EK : uint8_t {
  Empty, Literal, LiteralBytes, AnyChar, AnyCharNoNL, AnyByte, AnyByteNoNL,
  Class, ClassBytes, StartLine, EndLine, StartText, EndText,
  WordBoundary, NotWordBoundary, WordBoundaryAscii, NotWordBoundaryAscii,
  Group, Repeat, Concat, Alternate,
};
enum class RepKind : uint8_t { ZeroOrOne, ZeroOrMore, OneOrMore, Range };

struct Expr {
  EK kind = EK::Empty;
  std::vector<uint32_t> chars;  // Literal
  std::vector<uint8_t> bytes;   // LiteralBytes
  bool casei = false;
  CharClass cls;                // Class
  ByteClass bcls;               // ClassBytes
  int cap = 0;                  // Group: capture index (0 = non-capturing)
  std::string name;             // Group: capture name ("" = unnamed)
  RepKind rep = RepKind::ZeroOrOne;
  uint32_t rmin = 0, rmax = 0;
  bool has_max = false;
  bool greedy = true;
  std::vector<Expr> es;         // children (Group/Repeat: es[0])

  bool can_repeat() const;
  bool is_anchored_start() const;
  bool is_anchored_end() const;
  bool has_bytes() const;
};

struct Flags {
  bool casei = false, multi = false, dotnl = false, swap_greed = false,
       ignore_space = false, unicode = true, allow_bytes = false;
};

struct Error {
  enum Kind { None = 0, Syntax, CompiledTooBig, DfaTooBig, UnicodeWordBoundary, InvalidUtf8Pattern, UnresolvableLookPair } kind = None;
  std::string msg;
  explicit operator bool() const { return kind != None; }
};

// Parses `pat` (UTF-8).  On failure returns false and fills err (Syntax).
bool parse(const std::string& pat, Flags flags, size_t nest_limit, Expr* out, Error* err);

// class helpers (exposed for tests / compiler)
CharClass class_canonicalize(CharClass c);
CharClass class_negate(CharClass c);
CharClass class_case_fold(const CharClass& c);
ByteClass bclass_canonicalize(ByteClass c);
ByteClass bclass_case_fold(const ByteClass& c);
bool is_unicode_whitespace(uint32_t c);
bool is_unicode_word_char(uint32_t c);

// UTF-8 range splitting: a scalar range -> sequences of byte ranges whose
// concatenation language is exactly the UTF-8 encodings of the range
// (what the reference gets from the utf8-ranges crate, src/compile.rs:885-895).
struct Utf8Seq { int n; ByteRange r[4]; };
void utf8_sequences(uint32_t lo, uint32_t hi, std::vector<Utf8Seq>* out);

// ------------------------------------------------------------ program ----
enum class Look : uint8_t {
  StartLine, EndLine, StartText, EndText,
  WordBoundary, NotWordBoundary, WordBoundaryAscii, NotWordBoundaryAscii,
};
enum class Op : uint8_t { Match, Save, Split, EmptyLook, Bytes };
struct Inst {
  Op op;
  uint32_t a = 0;   // Match: pattern index; Split: goto1; Save/EmptyLook/Bytes: goto
  uint32_t b = 0;   // Split: goto2; Save: slot
  Look look = Look::StartLine;
  uint8_t lo = 0, hi = 0;  // Bytes
};

struct Program {
  std::vector<Inst> insts;
  std::vector<uint32_t> matches;   // ip of Match(i)
  uint32_t start = 0;              // entry incl. the lazy any* prefix (if compiled)
  uint32_t start_anchored = 0;     // entry of the pattern proper (no prefix)
  uint8_t byte_classes[256];
  int num_classes = 1;             // byte classes only (EOF class is num_classes)
  bool only_utf8 = true;
  bool is_reverse = false;
  bool has_prefix = false;         // lazy any* prefix present
  bool is_anchored_start = false, is_anchored_end = false;
  bool has_unicode_word_boundary = false;
  bool has_looks = false;          // any EmptyLook instruction
  size_t n_patterns = 1;
  int n_captures = 1;              // incl. group 0 (only meaningful when saves=true)
};

struct CompileOptions {
  bool only_utf8 = true;
  bool reverse = false;
  // Prepend the lazy `(?s:.)*?` (AnyChar/AnyByte per only_utf8) unless every
  // pattern is anchored at the start (compile.rs:142-153, prog.rs:128-130).
  bool unanchored_prefix = true;
  bool saves = false;              // emit Save for group 0 (Pike VM oracle only)
  size_t size_limit = 10u << 20;
};
bool compile(const std::vector<Expr>& exprs, const CompileOptions& opt, Program* out, Error* err);

std::string dump_program(const Program& p);

// Shortest / longest possible match length in bytes (kUnbounded = no bound).
constexpr uint64_t kUnbounded = ~0ull;
void expr_len_range(const Expr& e, uint64_t* mn, uint64_t* mx);

}  // namespace rb
