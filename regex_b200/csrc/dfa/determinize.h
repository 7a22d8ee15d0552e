// Offline determinizer: byte-level NFA program -> dense byte-class DFA table.
//
// This is the host half of the B200 search path.  It performs, eagerly and to
// completion, the subset construction that the reference performs lazily while
// scanning (src/dfa.rs:910-1048 exec_byte, :1073-1134 follow_epsilons,
// :1196-1244 cached_state_key, :1370-1434 start states), then minimises the
// automaton and renumbers states so the kernels can classify a state with one
// compare:   0 = dead,  [1, match_lo) = live,  [match_lo, n_states) = match.
//
// Match semantics are the reference's: a state is a *match state* when the
// state it was reached FROM contained a Match instruction (after the pre-byte
// re-closure of look-arounds), i.e. matches are reported one byte late and the
// extra EOF class flushes a match that ends at the end of the haystack
// (dfa.rs:963-968, 748-763).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../frontend/frontend.h"

namespace rb {

struct DfaOptions {
  bool anchored = false;        // start at the pattern proper, ignoring any lazy any* prefix
  bool leftmost_first = false;  // drop lower-priority threads after a Match (dfa.rs:984-988)
  size_t max_table_bytes = 32u << 20;
};

struct Dfa {
  uint32_t n_states = 0;        // incl. dead state 0
  uint32_t n_classes = 0;       // byte classes + 1; the last class is EOF
  uint32_t match_lo = 0;        // states >= match_lo carry a (delayed) match
  uint32_t mask_words = 1;      // 64-bit words per state in `masks`
  std::vector<uint16_t> trans;  // [n_states][n_classes]
  std::vector<uint64_t> masks;  // [n_states][mask_words]: patterns matched (delayed)
  uint8_t classes[256];
  // start state by flag index (dfa.rs:1381-1390):
  //   bit0 start-of-text, bit1 end-of-text (empty haystack), bit2 start-line,
  //   bit3 end-line (empty haystack), bit4 word-boundary, bit5 not-word-boundary,
  //   bit6 previous byte is an ASCII word byte
  uint16_t start[128];
  bool uniform_start = true;    // all feasible flag combinations share one start state
  bool reverse = false;
  bool has_looks = false;
  bool matches_empty = false;   // some start state leads to a match without consuming a byte
  size_t raw_states = 0;        // before minimisation (diagnostics)

  uint16_t next(uint16_t s, uint8_t byte) const { return trans[(size_t)s * n_classes + classes[byte]]; }
  uint16_t next_eof(uint16_t s) const { return trans[(size_t)s * n_classes + (n_classes - 1)]; }
  bool is_match(uint16_t s) const { return s >= match_lo; }
};

// Returns false with err->kind == DfaTooBig when the budget is exceeded, or
// UnicodeWordBoundary when the program needs Unicode-aware \b (north star: no
// fallback engine, explicit error).
bool determinize(const Program& prog, const DfaOptions& opt, Dfa* out, Error* err);

// Feasible start-flag index for a forward scan starting at `at`
// (dfa.rs:1415-1434) / reverse scan starting at `at` (dfa.rs:1440-1464).
int start_flag_index_forward(const uint8_t* text, size_t len, size_t at);
int start_flag_index_reverse(const uint8_t* text, size_t len, size_t at);

std::string dump_dfa(const Dfa& d);

}  // namespace rb
