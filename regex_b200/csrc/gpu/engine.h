// Host side of the B200 search backend: owns the determinized tables on the
// device and orchestrates the kernels.  This is the analogue of the reference's
// ExecNoSync (src/exec.rs:365-596): same operations, same results, but every
// search runs on the GPU -- there is no CPU matching path in this library.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../dfa/determinize.h"
#include "../frontend/frontend.h"

namespace rbgpu {

struct CompileOptions {
  uint32_t flags = 32;               // RURE_FLAG_* bits (rure.h:56-68); default = UNICODE
  size_t size_limit = 10u << 20;     // program size limit (re_builder.rs:29)
  size_t dfa_size_limit = 2u << 20;  // scaled x16 into the dense-table budget
  bool only_utf8 = false;            // false = bytes::Regex, true = Regex (str)
  bool as_set = false;               // RegexSet semantics even for one pattern
};

enum DfaKind { kFwdAnchoredLF = 0, kRevUnanchoredAll, kFwdUnanchoredAll, kRevAnchoredLongest, kFwdUnanchoredLF, kNumDfaKinds };

struct Tuning {
  uint32_t seg = 0;        // bytes per scan segment (multiple of 64); 0 = automatic
  bool force_generic = false;  // tests: use the generic scan kernel even when the fast one applies
  bool fuse = true;            // walk each segment's chain inside the fast scan kernel
  bool tensor_tma = true;      // feed the fast scan kernel with 2-D tiled TMA loads
  uint32_t chunk = 2048;   // bitmap bits per chain-walk chunk (one thread; multiple of 256)
  uint32_t warm = 0;       // 0 = automatic (bounded patterns: max match length; else 128)
  uint32_t block = 256;
  uint32_t blocks_per_sm = 8;
  uint64_t wave0 = 64ull << 20;  // first wave of a forward search in bytes (x16 per wave); 0 = one wave
  bool narrow_sets = true;
  int batch_refill = 1;          // 0 off, 1 is_match on large batches, 2 always and also find: batched is_match: lanes take the next record as soon as theirs is decided (batch_refill) instead of one record per lane per round
  int prefilter = 1;             // literal_scan instead of the DFA scan: 0 never, 1 when the scanned byte is estimated rare enough
                                 // to win (one byte, <= ~0.12 % of the haystack), 2 whenever the pattern qualifies structurally       // RegexSet::matches: continue with the automaton of the still-unmatched patterns
  uint32_t max_stitch_rounds = 16;  // re-walk rounds before the stitch falls back to one sequential pass
  uint32_t max_redo_rounds = 3;     // scan redo rounds before segment entry states are solved by state-map composition
};

struct Stats {  // filled by the last single-haystack call (diagnostics, bench roofline)
  uint64_t scan_redo_rounds = 0, scan_redo_segments = 0;
  uint64_t stitch_rounds = 0, stitch_dirty_chunks = 0, sequential_passes = 0, map_passes = 0, waves = 0, long_runs = 0;
  float scan_ms = 0, walk_ms = 0, total_ms = 0;
  bool fused = false;  // the scan kernel also walked the chains (scan_ms covers both)
  int path = 0;        // last find_all: 0 generic scan, 1 fast scan, 2 fused scan + walk, 3 literal prefilter
};

constexpr uint32_t kNoState = 0xFFFFFFFFu;

// One byte-range shard of a larger haystack (multi-GPU).  The buffer is
// [left context | owned bytes | right halo]; positions are buffer-relative.
struct ShardIO {
  // in
  uint64_t own_lo = 0, own_hi = 0;  // owns match starts at positions (own_lo, own_hi] (+ position 0 if own_lo == 0)
  bool is_first = true, is_last = true;  // buffer begins / ends where the haystack does
  uint32_t rev_entry = kNoState;    // exact reverse-scan state at own_hi from the right neighbour
  uint64_t chain_p = 0, chain_lm = ~0ull;  // iterator state entering the shard (kSpec = speculate)
  bool reuse_scan = false;          // keep the start bitmap of the previous call on this buffer
  bool chain_clamped = false;       // chain_p stands in for a restart point left of the buffer (look-around patterns:
                                    // the call fails with left_ctx_short if the answer would depend on bytes before it)
  // out
  uint32_t rev_guess = 0;           // state assumed at own_hi
  uint32_t rev_left = 0;            // exact state at own_lo (what the left neighbour must assume)
  uint64_t exit_p = 0, exit_lm = 0; // iterator state leaving the shard
  uint64_t n_matches = 0;
  bool halo_overflow = false;
  bool left_ctx_short = false;
};

class DeviceBuf {
 public:
  ~DeviceBuf();
  void* ensure(size_t bytes);  // grow-only
  void* ptr = nullptr;
  size_t cap = 0;
};

class Regex {
 public:
  static Regex* compile(const std::vector<std::string>& patterns, const CompileOptions& opt, rb::Error* err);
  ~Regex();

  size_t n_patterns() const { return patterns_.size(); }
  bool is_set() const { return is_set_; }
  // Exclusive use of this object's device scratch for one call (the entry points lock mu_ themselves; the C ABI
  // uses these to find an idle engine when several threads share one rure*).
  bool try_acquire() { return mu_.try_lock(); }
  void acquire() { mu_.lock(); }
  void release() { mu_.unlock(); }
  const rb::Dfa* host_dfa(DfaKind k, rb::Error* err);  // builds lazily (also used by tests to inspect tables)

  // ---- single haystack, device-resident text --------------------------------
  // find_iter (re_trait.rs:197-220): writes up to cap {start,end} pairs to d_out
  // (device memory, may be null to count only); *total = number of matches.
  int find_all_device(const uint8_t* d_text, uint64_t n, uint64_t start, uint64_t* d_out, uint64_t cap, uint64_t* total);
  // the same over one shard of a sharded haystack; spans are buffer-relative
  int find_all_shard_device(const uint8_t* d_text, uint64_t n, ShardIO* io, uint64_t* d_out, uint64_t cap);
  // find_at (exec.rs:473-514)
  int find_at_device(const uint8_t* d_text, uint64_t n, uint64_t start, bool* found, uint64_t* s, uint64_t* e);
  // shortest_match_at / is_match_at (exec.rs:382-468); for sets: any pattern.
  int shortest_match_device(const uint8_t* d_text, uint64_t n, uint64_t start, bool* found, uint64_t* end);
  // RegexSet::matches (re_set.rs:184-213): masks = ceil(n_patterns/64) words (host).
  int set_matches_device(const uint8_t* d_text, uint64_t n, uint64_t start, bool* any, uint64_t* masks);

  // one shard of a forward search over a sharded haystack (see engine.cu)
  int forward_shard_device(const uint8_t* d_text, uint64_t n, uint64_t own_lo, uint64_t own_hi, bool is_first, bool is_last, uint32_t entry,
                           bool want_masks, bool* found, uint64_t* first_end, uint64_t* masks, uint32_t* entry_used, uint32_t* exit_state);

  // ---- bulk operations over the span list (re_bytes.rs:316-360 split/splitn, :476-535 replacen) ----
  // rep: replacement; expand = `$0`, `${0}`, `$$` as src/expand.rs (a reference to another group fails: captures
  // beyond group 0 are out of scope; NoExpand = expand false); limit 0 = all matches.  d_out may be null to size.
  int replace_device(const uint8_t* d_text, uint64_t n, const uint8_t* rep, uint64_t rep_len, bool expand, uint64_t limit,
                     uint8_t* d_out, uint64_t out_cap, uint64_t* out_len);
  int replace_host(const uint8_t* text, uint64_t n, const uint8_t* rep, uint64_t rep_len, bool expand, uint64_t limit,
                   uint8_t* out, uint64_t out_cap, uint64_t* out_len);
  // pieces as (start, end) pairs; has_limit/limit as splitn
  int split_device(const uint8_t* d_text, uint64_t n, bool has_limit, uint64_t limit, uint64_t* d_pieces, uint64_t cap, uint64_t* n_pieces);
  int split_host(const uint8_t* text, uint64_t n, bool has_limit, uint64_t limit, uint64_t* pieces, uint64_t cap, uint64_t* n_pieces);

  // ---- capture groups (exec.rs:527-590 read_captures_at, :861-875 captures_nfa_with_match) ----
  int n_groups() const { return n_groups_; }
  const std::vector<std::pair<std::string, int>>& group_names() const { return group_name_index_; }
  // slots of every match in d_spans: d_slots[m][2 * n_groups], kNone where a group did not take part
  int captures_device(const uint8_t* d_text, uint64_t n, const uint64_t* d_spans, uint64_t m, uint64_t* d_slots);
  // Regex::captures at `start` on a host haystack: slots[2 * n_groups]
  int captures_at_host(const uint8_t* text, uint64_t n, uint64_t start, bool* found, uint64_t* slots);
  // every match with its groups: slots[min(cap, *m)][2 * n_groups]
  int captures_all_host(const uint8_t* text, uint64_t n, uint64_t* slots, uint64_t cap, uint64_t* m);

  // ---- batched records, device-resident text + offsets[n_rec+1] --------------
  int is_match_batch_device(const uint8_t* d_text, const uint64_t* d_offsets, uint64_t n_rec, uint32_t* d_bits);
  int find_batch_device(const uint8_t* d_text, const uint64_t* d_offsets, uint64_t n_rec, uint64_t* d_spans, uint32_t* d_bits);
  int set_matches_batch_device(const uint8_t* d_text, const uint64_t* d_offsets, uint64_t n_rec, uint64_t* d_masks);

  // ---- host-buffer wrappers (copies inside; the e2e path) ---------------------
  int find_all_host(const uint8_t* text, uint64_t n, uint64_t start, uint64_t* out, uint64_t cap, uint64_t* total);
  int find_at_host(const uint8_t* text, uint64_t n, uint64_t start, bool* found, uint64_t* s, uint64_t* e);
  int shortest_match_host(const uint8_t* text, uint64_t n, uint64_t start, bool* found, uint64_t* end);
  int set_matches_host(const uint8_t* text, uint64_t n, uint64_t start, bool* any, uint64_t* masks);
  int is_match_batch_host(const uint8_t* text, const uint64_t* offsets, uint64_t n_rec, uint8_t* out_bits);
  int find_batch_host(const uint8_t* text, const uint64_t* offsets, uint64_t n_rec, uint64_t* spans, uint8_t* out_bits);
  int set_matches_batch_host(const uint8_t* text, const uint64_t* offsets, uint64_t n_rec, uint64_t* masks);

  // Run on a caller-owned CUDA stream (e.g. torch's current stream) instead of the private one.
  void set_stream(void* cuda_stream) { ext_stream_ = cuda_stream; use_ext_stream_ = true; }
  const std::string& last_error() const { return error_; }
  Tuning tuning;
  Stats stats;
  uint64_t min_len = 0, max_len = 0;  // match length range in bytes (max = kUnbounded)
  bool can_match_empty = false;
  bool has_looks = false;
  bool only_utf8 = false;

 private:
  Regex() = default;
  struct DeviceDfa;
  int init_device();
  int d2h(void* dst, const void* src, size_t bytes);
  int ensure(DfaKind k, DeviceDfa** out);
  int fail(const std::string& msg);
  int check(int cuda_err, const char* what);
  struct ScanPlan { bool fast = false; uint32_t seg = 0, warm = 0; uint64_t n_seg = 0; };
  ScanPlan plan_scan(const uint8_t* d_text, uint64_t base, uint64_t limit, bool fast_table);
  int scan_starts(const uint8_t* d_text, uint64_t n, uint64_t base, uint64_t limit, ShardIO* io, const ScanPlan& plan,
                  const void* fused_walk);
  int solve_entries(const void* scan_args, bool reverse);
  int resolve_long_run(const uint8_t* d_text, uint64_t n, uint64_t s, bool text_continues, uint64_t* e_out);
  int all_spans_device(const uint8_t* d_text, uint64_t n, uint64_t** d_spans, uint64_t* m);
  int ensure_capture_program();
  int replace_prepare(const uint8_t* d_text, uint64_t n, const uint8_t* rep, uint64_t rep_len, bool expand, uint64_t limit, uint64_t* out_len);
  int replace_emit(uint8_t* d_out, uint64_t out_cap);
  bool plan_prefilter();  // fills pf_words_ (launch.h PfArgs) when every match has one of <= 4 bytes at a fixed offset
  int forward_reduce(const uint8_t* d_text, uint64_t n, uint64_t start, bool want_masks, uint64_t* result_host, uint64_t end = ~0ull,
                     uint32_t entry = 0xFFFFFFFFu, uint32_t* exit_out = nullptr);
  int forward_range(const uint8_t* d_text, uint64_t n, uint64_t start, uint64_t limit, uint32_t entry, bool want_masks,
                    uint64_t* result_host, uint32_t* exit_state);
  int subset(const std::vector<uint32_t>& members, Regex** out);
  const uint8_t* upload_text(const uint8_t* text, uint64_t n, int* rc);
  int find_all_host_pipelined(const uint8_t* text, uint64_t n, uint8_t* d, uint64_t* d_out, uint64_t* out, uint64_t cap, uint64_t* total);

  struct LazyUpload { const uint8_t* src; uint8_t* dst; uint64_t n, done; };  // host haystack uploaded wave by wave
  LazyUpload* lazy_ = nullptr;
  std::vector<std::string> patterns_;
  bool is_set_ = false;
  bool sparse_set_ = false;  // a narrowed RegexSet: none of its patterns matched in the first waves
  std::map<std::vector<uint32_t>, std::unique_ptr<Regex>> subsets_;
  CompileOptions opt_;
  std::vector<rb::Expr> exprs_;
  std::unique_ptr<rb::Dfa> host_[kNumDfaKinds];
  DeviceDfa* dev_[kNumDfaKinds] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  std::string error_;
  std::recursive_mutex mu_;
  void* stream_ = nullptr;
  void* own_stream_ = nullptr;
  void* copy_stream_ = nullptr;  // host -> device pieces of a pipelined find_all
  void* back_stream_ = nullptr;  // device -> host spans
  void* ext_stream_ = nullptr;
  bool use_ext_stream_ = false;
  // scratch (grow-only)
  DeviceBuf text_, offsets_, bitmap_, guess_, fin_, redo_, counters_, seg_first_, seg_mask_;
  DeviceBuf long_tab_;
  static constexpr uint32_t kMaxLongRuns = 256;
  DeviceBuf spans_all_, lens_, lits_, rep_out_, pieces_, reps_before_;
  uint64_t rep_totals_[2] = {0, 0};  // matched bytes, replacement bytes of the pending replace call
  std::vector<uint8_t> rep_args_, rep_lits_;  // launch.h ReplaceArgs image + literal bytes of the pending replace call
  int n_groups_ = 1;                      // capture groups incl. group 0
  std::vector<std::string> group_names_;  // names of the named groups
  std::vector<std::pair<std::string, int>> group_name_index_;  // (name, group index)
  DeviceBuf cap_insts_, cap_scratch_, cap_slots_, cap_span_;
  uint32_t cap_n_insts_ = 0, cap_start_ = 0;  // capture program on the device (0 = not uploaded yet)
  bool cap_anchored_ = false;
  DeviceBuf in_p_, in_lm_, out_p_, out_lm_, count_, offset_, dirty_, first_cand_, skip_, meta_, excl_, btot_, present_, kidx_, kstates_, maps_, comp_, bentry_, exact_, stage_, block_sums_, out_, bits_, masks_;
  void* pinned_ = nullptr;  // small pinned staging area for counters / scalars
  void* timing_events_[3] = {nullptr, nullptr, nullptr};
  void *fork_event_ = nullptr, *join_event_ = nullptr;  // forward scans: the generic edge kernel runs beside the fast one
  int pf_state_ = 0;               // 0 = not decided yet, 1 = prefilter applies, -1 = it does not
  uint32_t pf_freq_ = 0;           // estimated frequency of the scanned bytes, parts per 65536
  std::vector<uint32_t> pf_words_; // PfArgs image (engine.cu)
};

uint64_t kernel_launches();  // total kernels launched by this library in this process
int device_sm_count();

}  // namespace rbgpu
