// Host orchestration of the search kernels.  See engine.h.
#include "engine.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "launch.h"

namespace rbgpu {

// cuTensorMapEncodeTiled through the runtime (the library links cudart statically, not libcuda).
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}


static std::atomic<uint64_t> g_launches{0};
uint64_t kernel_launches() { return g_launches.load(); }

int device_sm_count() {
  static int sms = -1;
  if (sms < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
  }
  return sms;
}

DeviceBuf::~DeviceBuf() { if (ptr) cudaFree(ptr); }
void* DeviceBuf::ensure(size_t bytes) {
  if (bytes <= cap) return ptr;
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
  cap = 0;
  size_t want = std::max<size_t>(bytes, 256);
  if (cudaMalloc(&ptr, want) != cudaSuccess) { ptr = nullptr; return nullptr; }
  cap = want;
  return ptr;
}

// Hot tables of up to this many rows (288 bytes per row, one-byte successor ids) are
// staged in shared memory by the fast kernels.
static constexpr uint32_t kFastStates = 200;
static size_t hot_bytes(uint32_t rows) { return ((size_t)rows * RB_HOT_ROW + 255) / 256 * 256; }  // kernels.cu hot_table_bytes
// scan_rev_fast's reverse table with signed row ids (kernels.cu hot_signed_bytes): match rows and the others are rounded apart
static size_t hot_bytes_signed(const HotView& h) { return hot_bytes(h.n - h.match_lo) + hot_bytes(h.match_lo); }
static bool hot_signed_ok(const HotView& h) { return h.n != 0 && h.match_lo <= 128 && h.n - h.match_lo <= 128; }

struct Regex::DeviceDfa {
  DfaView view;
  void* trans = nullptr;
  void* classes = nullptr;
  void* masks = nullptr;
  // byte-indexed expansion over the hot states for the fast kernels (kernels.cuh HotView);
  // hot.n == 0 when even the ASCII-reachable part has more than kFastStates states
  void* next256 = nullptr;
  void* eof = nullptr;
  void* hot2full = nullptr;
  void* full2hot = nullptr;
  void* view_dev = nullptr;  // copy of `view` in global memory (cold fallbacks take a pointer)
  HotView hot{};
  ~DeviceDfa() {
    cudaFree(trans); cudaFree(classes); cudaFree(masks); cudaFree(next256); cudaFree(eof);
    cudaFree(hot2full); cudaFree(full2hot); cudaFree(view_dev);
  }
};

// ------------------------------------------------------------------ compile --
Regex* Regex::compile(const std::vector<std::string>& patterns, const CompileOptions& opt, rb::Error* err) {
  std::unique_ptr<Regex> re(new Regex());
  re->patterns_ = patterns;
  re->opt_ = opt;
  re->only_utf8 = opt.only_utf8;
  re->is_set_ = opt.as_set || patterns.size() != 1;
  rb::Flags f;
  f.casei = opt.flags & 1; f.multi = opt.flags & 2; f.dotnl = opt.flags & 4; f.swap_greed = opt.flags & 8;
  f.ignore_space = opt.flags & 16; f.unicode = opt.flags & 32;
  f.allow_bytes = !opt.only_utf8;  // exec.rs:224-225
  re->exprs_.resize(patterns.size());
  re->min_len = rb::kUnbounded;
  re->max_len = 0;
  for (size_t i = 0; i < patterns.size(); i++) {
    if (!rb::parse(patterns[i], f, 200, &re->exprs_[i], err)) return nullptr;
    uint64_t mn, mx;
    rb::expr_len_range(re->exprs_[i], &mn, &mx);
    re->min_len = std::min(re->min_len, mn);
    re->max_len = std::max(re->max_len, mx);
  }
  if (patterns.size() == 1) {  // capture groups of the pattern (only their existence matters: replacement templates)
    std::vector<const rb::Expr*> stack{&re->exprs_[0]};
    while (!stack.empty()) {
      const rb::Expr* x = stack.back();
      stack.pop_back();
      if (x->kind == rb::EK::Group && x->cap > 0) {
        re->n_groups_ = std::max(re->n_groups_, x->cap + 1);
        if (!x->name.empty()) { re->group_names_.push_back(x->name); re->group_name_index_.emplace_back(x->name, x->cap); }
      }
      for (const rb::Expr& c : x->es) stack.push_back(&c);
    }
  }
  if (patterns.empty()) { re->min_len = 0; return re.release(); }
  re->can_match_empty = re->min_len == 0;
  // Validate eagerly what the reference validates at build time (program size)
  // plus the two explicit errors of this backend (Unicode \b, table budget).
  rb::Program probe;
  rb::CompileOptions co;
  co.only_utf8 = opt.only_utf8;
  co.size_limit = opt.size_limit;
  if (!rb::compile(re->exprs_, co, &probe, err)) return nullptr;
  re->has_looks = probe.has_looks;
  if (probe.has_unicode_word_boundary) {
    err->kind = rb::Error::UnicodeWordBoundary;
    err->msg = "Unicode word boundaries (\\b, \\B) are not supported by the B200 DFA backend; use (?-u:\\b).";
    return nullptr;
  }
  if (patterns.size() > 64 * kMaxMaskWords) {
    err->kind = rb::Error::DfaTooBig;
    err->msg = "regex sets larger than 256 patterns exceed the per-state mask budget";
    return nullptr;
  }
  // Determinize the tables every object needs so budget errors surface at compile time.
  if (!re->host_dfa(kFwdUnanchoredAll, err)) return nullptr;
  if (!re->is_set_) {
    if (!re->host_dfa(kFwdAnchoredLF, err)) return nullptr;
    if (!re->host_dfa(kRevUnanchoredAll, err)) return nullptr;
  }
  return re.release();
}

Regex::~Regex() {
  for (auto*& d : dev_) { delete d; d = nullptr; }
  if (pinned_) cudaFreeHost(pinned_);
  for (void* e : timing_events_) if (e) cudaEventDestroy((cudaEvent_t)e);
  if (fork_event_) cudaEventDestroy((cudaEvent_t)fork_event_);
  if (join_event_) cudaEventDestroy((cudaEvent_t)join_event_);
  if (copy_stream_) cudaStreamDestroy((cudaStream_t)copy_stream_);
  if (back_stream_) cudaStreamDestroy((cudaStream_t)back_stream_);
  if (own_stream_) cudaStreamDestroy((cudaStream_t)own_stream_);
}

const rb::Dfa* Regex::host_dfa(DfaKind k, rb::Error* err) {
  if (host_[k]) return host_[k].get();
  if (patterns_.empty()) {  // an empty RegexSet matches nothing and has no automaton
    err->kind = rb::Error::Syntax;
    err->msg = "an empty pattern set has no automaton";
    return nullptr;
  }
  rb::CompileOptions co;
  co.only_utf8 = opt_.only_utf8;
  co.size_limit = opt_.size_limit;
  rb::DfaOptions dopt;
  dopt.max_table_bytes = opt_.dfa_size_limit * 16;
  switch (k) {
    case kFwdAnchoredLF: co.reverse = false; co.unanchored_prefix = false; dopt.anchored = true; dopt.leftmost_first = true; break;
    case kRevUnanchoredAll: co.reverse = true; co.unanchored_prefix = true; dopt.anchored = false; dopt.leftmost_first = false; break;
    case kFwdUnanchoredAll: co.reverse = false; co.unanchored_prefix = true; dopt.anchored = false; dopt.leftmost_first = false; break;
    case kRevAnchoredLongest: co.reverse = true; co.unanchored_prefix = false; dopt.anchored = true; dopt.leftmost_first = false; break;
    case kFwdUnanchoredLF: co.reverse = false; co.unanchored_prefix = true; dopt.anchored = false; dopt.leftmost_first = true; break;
    default: return nullptr;
  }
  if (is_set_ || patterns_.size() != 1) dopt.leftmost_first = false;  // dfa.rs:1557-1559
  rb::Program prog;
  if (!rb::compile(exprs_, co, &prog, err)) return nullptr;
  std::unique_ptr<rb::Dfa> d(new rb::Dfa());
  if (!rb::determinize(prog, dopt, d.get(), err)) return nullptr;
  host_[k] = std::move(d);
  return host_[k].get();
}

int Regex::fail(const std::string& msg) {
  error_ = msg;
  return -1;
}
int Regex::check(int e, const char* what) {
  if (e == cudaSuccess) return 0;
  return fail(std::string("CUDA error in ") + what + ": " + cudaGetErrorString((cudaError_t)e));
}
#define RB_CUDA(call)                                   \
  do {                                                  \
    int rc__ = check((int)(call), #call);               \
    if (rc__) return rc__;                              \
  } while (0)
#define RB_LAUNCH_CHECK(name)                           \
  do {                                                  \
    g_launches++;                                       \
    int rc__ = check((int)cudaGetLastError(), name);    \
    if (rc__) return rc__;                              \
  } while (0)

// Stream, pinned scalars and timing events of this object.  Every copy and kernel of the
// library is issued on stream_ (or on the two pipeline streams, ordered by events).  The
// library's own stream is a BLOCKING stream: it is ordered with the legacy default stream,
// which is where a caller that never heard of streams (or torch's default stream) produced the
// device buffers it hands to the *_device entry points -- a non-blocking stream raced with
// `torch.zeros(...)` of the output buffer in the tests.  Callers on another stream pass it
// with rure_b200_set_stream.
int Regex::init_device() {
  if (!pinned_) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
      return fail("no CUDA device available: regex_b200 has no CPU matching path");
    cudaStream_t s;
    RB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamDefault));
    own_stream_ = s;
    RB_CUDA(cudaMallocHost(&pinned_, 4096));
    for (void*& e : timing_events_) {
      cudaEvent_t ev;
      RB_CUDA(cudaEventCreate(&ev));
      e = ev;
    }
  }
  stream_ = use_ext_stream_ ? ext_stream_ : own_stream_;
  return 0;
}

int Regex::ensure(DfaKind k, DeviceDfa** out) {
  if (int rc = init_device()) return rc;
  if (dev_[k]) { *out = dev_[k]; return 0; }
  cudaStream_t up = (cudaStream_t)stream_;
  rb::Error err;
  const rb::Dfa* h = host_dfa(k, &err);
  if (!h) return fail(err.msg);
  std::unique_ptr<DeviceDfa> d(new DeviceDfa());
  const size_t tb = h->trans.size() * sizeof(uint16_t);
  RB_CUDA(cudaMalloc(&d->trans, std::max<size_t>(tb, 16)));
  RB_CUDA(cudaMalloc(&d->classes, 256));
  RB_CUDA(cudaMalloc(&d->masks, std::max<size_t>(h->masks.size() * 8, 16)));
  RB_CUDA(cudaMemcpyAsync(d->trans, h->trans.data(), tb, cudaMemcpyHostToDevice, up));
  RB_CUDA(cudaMemcpyAsync(d->classes, h->classes, 256, cudaMemcpyHostToDevice, up));
  RB_CUDA(cudaMemcpyAsync(d->masks, h->masks.data(), h->masks.size() * 8, cudaMemcpyHostToDevice, up));
  DfaView& v = d->view;
  v.trans = (const uint16_t*)d->trans;
  v.classes = (const uint8_t*)d->classes;
  v.masks = (const uint64_t*)d->masks;
  v.n_states = h->n_states;
  v.stride = h->n_classes;
  v.match_lo = h->match_lo;
  v.mask_words = h->mask_words;
  v.table_bytes = (uint32_t)tb;
  std::memcpy(v.start, h->start, sizeof v.start);
  v.uniform_start = h->uniform_start;
  // Hot set: every state when they all fit, else the closure of the start states under
  // ASCII bytes (what an ASCII haystack can reach).  Row 1 is the trap.
  {
    std::vector<uint8_t> is_hot(h->n_states, 0);
    {
      std::vector<uint16_t> stack;
      auto push = [&](uint16_t s) { if (!is_hot[s]) { is_hot[s] = 1; stack.push_back(s); } };
      push(0);
      for (uint16_t s : h->start) push(s);
      while (!stack.empty()) {
        const uint16_t s = stack.back();
        stack.pop_back();
        for (int b = 0; b < 128; b++) push(h->next(s, (uint8_t)b));
      }
    }
    // All states when they fit -- unless the ASCII closure is a small part of the automaton: Unicode `\d`
    // in `(\d{4})-(\d{2})-(\d{2})` makes 183 states of which 15 are reachable on ASCII, and a 51 KB table
    // costs a block per SM and most of the L1 where 4 KB will do (non-ASCII bytes take the trap row).
    const size_t n_closure = (size_t)std::count(is_hot.begin(), is_hot.end(), 1);
    if (h->n_states + 1 <= kFastStates && n_closure * 4 > h->n_states) std::fill(is_hot.begin(), is_hot.end(), 1);
    const size_t n_hot = (size_t)std::count(is_hot.begin(), is_hot.end(), 1) + 1;
    if (n_hot <= kFastStates) {
      std::vector<uint16_t> full2hot(h->n_states, 0xFFFF), hot2full;
      hot2full.push_back(0);       // dead
      hot2full.push_back(0xFFFF);  // trap
      full2hot[0] = 0;
      for (uint32_t s = 1; s < h->n_states; s++)  // full ids are already ordered non-match < match
        if (is_hot[s]) { full2hot[s] = (uint16_t)hot2full.size(); hot2full.push_back((uint16_t)s); }
      uint32_t hot_match_lo = (uint32_t)hot2full.size();
      for (uint32_t i = 2; i < hot2full.size(); i++)
        if (hot2full[i] >= h->match_lo) { hot_match_lo = i; break; }
      std::vector<uint16_t> n256(hot2full.size() * 256), eof(hot2full.size(), 0);
      for (uint32_t r = 0; r < hot2full.size(); r++) {
        for (int b = 0; b < 256; b++) {
          uint16_t to = 1;  // trap row: absorbing
          if (r != 1) {
            const uint16_t t = full2hot[h->next(hot2full[r], (uint8_t)b)];
            to = t == 0xFFFF ? 1 : t;
          }
          n256[(size_t)r * 256 + b] = to;
        }
        if (r != 1) eof[r] = h->next_eof(hot2full[r]);
      }
      RB_CUDA(cudaMalloc(&d->next256, n256.size() * 2));
      RB_CUDA(cudaMalloc(&d->eof, eof.size() * 2));
      RB_CUDA(cudaMalloc(&d->hot2full, hot2full.size() * 2));
      RB_CUDA(cudaMalloc(&d->full2hot, full2hot.size() * 2));
      RB_CUDA(cudaMemcpyAsync(d->next256, n256.data(), n256.size() * 2, cudaMemcpyHostToDevice, up));
      RB_CUDA(cudaMemcpyAsync(d->eof, eof.data(), eof.size() * 2, cudaMemcpyHostToDevice, up));
      RB_CUDA(cudaMemcpyAsync(d->hot2full, hot2full.data(), hot2full.size() * 2, cudaMemcpyHostToDevice, up));
      RB_CUDA(cudaMemcpyAsync(d->full2hot, full2hot.data(), full2hot.size() * 2, cudaMemcpyHostToDevice, up));
      RB_CUDA(cudaStreamSynchronize(up));  // the host vectors above die with this scope
      d->hot.next256 = (const uint16_t*)d->next256;
      d->hot.eof = (const uint16_t*)d->eof;
      d->hot.hot2full = (const uint16_t*)d->hot2full;
      d->hot.full2hot = (const uint16_t*)d->full2hot;
      d->hot.n = (uint32_t)hot2full.size();
      d->hot.match_lo = hot_match_lo;
      d->hot.start = full2hot[h->start[32]];
    }
  }
  RB_CUDA(cudaMalloc(&d->view_dev, sizeof(DfaView)));
  RB_CUDA(cudaMemcpyAsync(d->view_dev, &d->view, sizeof(DfaView), cudaMemcpyHostToDevice, up));
  RB_CUDA(cudaStreamSynchronize(up));
  dev_[k] = d.release();
  *out = dev_[k];
  return 0;
}

// Shared-memory staging decision: table + class map must fit the opt-in limit.
static size_t smem_for(const DfaView& v) {
  size_t need = ((size_t)v.table_bytes + 15) / 16 * 16 + 256;
  return need <= 200 * 1024 ? need : 0;
}
// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute belongs to the function, not to the
// caller: it is only ever RAISED (per-kernel maximum under a lock) -- two regex objects with different table
// sizes searching from two threads would otherwise lower it under each other's launches ("invalid argument").
// min_bytes: kernels with static shared memory on top pass 48 KB so that the sum may exceed the default limit.
template <typename K>
static cudaError_t allow_smem(K kernel, size_t bytes, size_t min_bytes = 0) {
  bytes = std::max(bytes, min_bytes);
  if (bytes <= 48 * 1024 && min_bytes == 0) return cudaSuccess;
  static std::mutex mu;
  static std::unordered_map<const void*, size_t> granted;
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = granted[(const void*)kernel];
  if (bytes <= have) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}
static uint32_t grid_for(uint64_t threads_needed, uint32_t block, uint32_t blocks_per_sm) {
  uint64_t blocks = (threads_needed + block - 1) / block;
  uint64_t cap = (uint64_t)std::max(1, device_sm_count()) * blocks_per_sm;
  return (uint32_t)std::max<uint64_t>(1, std::min(blocks, cap));
}

static size_t fast_scan_smem(size_t table_bytes) {
  return table_bytes + 1024 + (1024 / 32) * (2 * 2048 + 64);  // kernels.cu kBoxStages, kRingBarBytes  // kernels.cu: tables, 512-aligned rings, mbarriers
}

static uint32_t pick_warm(const Regex& re) {
  if (re.tuning.warm) return (re.tuning.warm + 15) / 16 * 16;
  if (re.max_len != rb::kUnbounded) return (uint32_t)std::min<uint64_t>((std::max<uint64_t>(re.max_len, 1) + 15) / 16 * 16, 4096);
  return 128;
}

// ---------------------------------------------------------- literal prefilter --
// Estimated byte frequencies, parts per 65536 (tools/gen_byte_freq.py).
static const uint16_t kByteFreq[256] = {
#include "byte_freq.inc"
};
// A pattern qualifies when every match must have one of at most four bytes of small total
// estimated frequency at some fixed offset o < min_len (the reference scans for its rarest
// literal byte, src/literals.rs:466-489; here the bytes come from the automaton, so character
// classes and alternations qualify as well: `Sher[a-z]+|Hol[a-z]+` scans for S and H).
// Measured on the B200 (profiles/r02_prefilter.md): the word-wide byte tests cost ~2.3
// instructions per haystack byte for two byte values, and every hit costs a queue push plus its
// share of a verification round, against 5 instructions per byte for the shared-memory DFA scan.
// `Holmes|Watson` over the C2 corpus (an H or W every 238 bytes, a match every 1.1 KB) runs at
// 1.33 TB/s here and at 2.93 TB/s on the fused DFA path, so the automatic choice is conservative:
static constexpr uint32_t kPrefilterMaxFreq = 600;   // structural limit (tuning.prefilter == 2): ~0.9 % of the bytes
static constexpr uint32_t kPrefilterAutoFreq = 80;   // automatic (tuning.prefilter == 1): ~0.12 %, i.e. one rare byte
static constexpr uint32_t kPrefilterState = 0xFFFEu; // what a prefilter shard reports as its boundary states
bool Regex::plan_prefilter() {
  if (pf_state_) return pf_state_ > 0;
  pf_state_ = -1;
  if (is_set_ || has_looks || min_len == 0 || patterns_.size() != 1) return false;
  rb::Error err;
  const rb::Dfa* d = host_dfa(kFwdAnchoredLF, &err);
  if (!d || !d->uniform_start) return false;
  const uint32_t depth = (uint32_t)std::min<uint64_t>(min_len, 8);
  std::vector<std::vector<uint8_t>> allowed(depth, std::vector<uint8_t>(256, 0));
  std::vector<uint8_t> reach(d->n_states, 0), next(d->n_states, 0);
  reach[d->start[32]] = 1;
  for (uint32_t dep = 0; dep < depth; dep++) {
    std::fill(next.begin(), next.end(), 0);
    for (uint32_t st = 1; st < d->n_states; st++) {
      if (!reach[st]) continue;
      for (int b = 0; b < 256; b++) {
        const uint16_t t = d->next((uint16_t)st, (uint8_t)b);
        if (t) { allowed[dep][b] = 1; next[t] = 1; }
      }
    }
    reach.swap(next);
  }
  uint32_t best_o = 0, best_f = ~0u;
  for (uint32_t o = 0; o < depth; o++) {
    uint32_t cnt = 0, f = 0;
    for (int b = 0; b < 256; b++) if (allowed[o][b]) { cnt++; f += kByteFreq[b]; }
    if (cnt >= 1 && cnt <= 4 && f < best_f) { best_f = f; best_o = o; }
  }
  if (best_f > kPrefilterMaxFreq) return false;
  PfArgs a{};
  a.o = best_o;
  for (int b = 0; b < 256; b++) if (allowed[best_o][b]) a.bcast[a.n_bytes++] = 0x01010101u * (uint32_t)b;
  a.n_sets = std::min<uint32_t>(depth, 4);
  for (uint32_t dep = 0; dep < a.n_sets; dep++)
    for (int b = 0; b < 256; b++) if (allowed[dep][b]) a.sets[dep][b >> 5] |= 1u << (b & 31);
  pf_words_.resize(sizeof(PfArgs) / 4);
  std::memcpy(pf_words_.data(), &a, sizeof a);
  pf_freq_ = best_f;
  pf_state_ = 1;
  return true;
}

// ------------------------------------------------------------ long matches ----
// End of the leftmost-first match anchored at s when it is too long for one thread: the
// anchored automaton over [s, n) in 4 KiB segments, every segment entered in its exact state
// (state maps from EVERY state of the automaton, composed -- the states of `(?s)foo.*bar`
// after `foo` never converge, so no warm-up guess would do), then a maximum over the last match
// positions.  (states + 2) generic passes over the run instead of ~10 s per GiB for one thread.
int Regex::resolve_long_run(const uint8_t* d_text, uint64_t n, uint64_t s, bool text_continues, uint64_t* e_out) {
  DeviceDfa* fwd;
  if (int rc = ensure(kFwdAnchoredLF, &fwd)) return rc;
  cudaStream_t st = (cudaStream_t)stream_;
  const uint32_t k_max = fwd->view.n_states;
  const uint32_t seg = 4096;
  ScanArgs g{};
  g.dfa = fwd->view;
  const size_t smem = smem_for(fwd->view);
  g.use_smem = smem != 0;
  g.text = d_text;
  g.n = n;
  g.base = s;
  g.seg = seg;
  uint32_t* counters = (uint32_t*)counters_.ptr;
  uint16_t* d_states = (uint16_t*)kstates_.ensure(k_max * 2);
  uint16_t* d_kidx = (uint16_t*)kidx_.ensure(65536 * 2);
  if (!d_states || !d_kidx) return fail("out of device memory (long run)");
  // the start state at s: the look-behind flags come from the bytes next to s (pick_start_fwd)
  uint16_t start_state = fwd->view.start[32];
  if (!fwd->view.uniform_start) {
    uint8_t around[2] = {0, 0};
    const uint64_t lo = s ? s - 1 : 0, hi = std::min(n, s + 1);
    RB_CUDA(d2h(around, d_text + lo, hi - lo));
    const bool has_prev = s > 0, has_cur = s < n;
    const uint8_t prev = has_prev ? around[0] : 0, cur = has_cur ? around[has_prev ? 1 : 0] : 0;
    auto word = [](uint8_t b) { return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_'; };
    int f = 0;
    if (s == 0) f |= 1;
    if (n == 0) f |= 2 | 8;
    if (s == 0 || prev == '\n') f |= 4;
    const bool last = has_prev && word(prev), cw = has_cur && word(cur);
    f |= (last == cw) ? 32 : 16;
    if (last) f |= 64;
    start_state = fwd->view.start[f];
  }
  // K: the states the run is in at segment boundaries, closed under the per-segment maps, found as in
  // solve_entries: start from the run's start state, run the segments from the states known so far, add
  // the states their maps lead to, repeat (`(?s)foo.*bar`: 6 of 25 states)
  uint8_t* present = (uint8_t*)present_.ensure(65536);
  uint16_t* d_first = (uint16_t*)(counters + 11);
  if (!present) return fail("out of device memory (long run)");
  RB_CUDA(cudaMemsetAsync(present, 0, 65536, st));
  {
    const uint8_t one = 1;
    RB_CUDA(cudaMemcpyAsync(present + start_state, &one, 1, cudaMemcpyHostToDevice, st));
    RB_CUDA(cudaMemcpyAsync(d_first, &start_state, 2, cudaMemcpyHostToDevice, st));
    RB_CUDA(cudaStreamSynchronize(st));
  }
  std::vector<uint8_t> h_present(65536);
  std::vector<uint16_t> K, kidx;
  RB_CUDA(allow_smem(scan_map, smem));
  RB_CUDA(allow_smem(scan_last_match, smem));
  // growing windows from s: a long match is rarely the whole rest of the haystack
  for (uint64_t window = 8ull << 20;; window *= 8) {
    const uint64_t end = window >= n - s ? n : s + window;
    const uint64_t n_seg = std::max<uint64_t>(1, (end - s + seg - 1) / seg);
    const uint64_t n_blocks = (n_seg + kMapBlock - 1) / kMapBlock;
    g.limit = end;
    g.n_seg = n_seg;
    uint32_t k = 0;
    uint16_t* maps = nullptr;
    for (;;) {
      RB_CUDA(d2h(h_present.data(), present, 65536));
      K.clear();
      kidx.assign(65536, 0xFFFF);
      for (uint32_t v = 0; v < k_max; v++)
        if (h_present[v]) { kidx[v] = (uint16_t)K.size(); K.push_back((uint16_t)v); }
      k = (uint32_t)K.size();
      maps = (uint16_t*)maps_.ensure(n_seg * k * 2);
      if (!maps) return fail("out of device memory (long run)");
      RB_CUDA(cudaMemcpyAsync(d_states, K.data(), k * 2, cudaMemcpyHostToDevice, st));
      RB_CUDA(cudaMemcpyAsync(d_kidx, kidx.data(), 65536 * 2, cudaMemcpyHostToDevice, st));
      scan_map<<<grid_for(n_seg * k, 256, tuning.blocks_per_sm), 256, smem, st>>>(g, 0, d_states, k, maps);
      RB_LAUNCH_CHECK("scan_map");
      stats.map_passes += k;
      RB_CUDA(cudaMemsetAsync(counters + 8, 0, 4, st));
      closure_check<<<grid_for(n_seg * k, 256, 8), 256, 0, st>>>(maps, n_seg * k, d_kidx, present, counters + 8);
      RB_LAUNCH_CHECK("closure_check");
      uint32_t n_new = 0;
      RB_CUDA(d2h(&n_new, counters + 8, 4));
      if (n_new == 0) break;
    }
    uint16_t* comp = (uint16_t*)comp_.ensure(n_blocks * k * 2);
    uint16_t* bentry = (uint16_t*)bentry_.ensure(n_blocks * 2);
    uint16_t* exact = (uint16_t*)exact_.ensure(n_seg * 2);
    if (!comp || !bentry || !exact) return fail("out of device memory (long run)");
    compose_blocks<<<(uint32_t)((n_blocks * k + 255) / 256), 256, 0, st>>>(maps, d_kidx, d_states, k, n_seg, 0, comp);
    RB_LAUNCH_CHECK("compose_blocks");
    compose_top<<<1, 32, 0, st>>>(comp, d_kidx, k, n_blocks, 0, d_first, bentry);
    RB_LAUNCH_CHECK("compose_top");
    compose_fill<<<(uint32_t)((n_blocks + 255) / 256), 256, 0, st>>>(maps, d_kidx, k, n_seg, 0, bentry, exact);
    RB_LAUNCH_CHECK("compose_fill");
    unsigned long long* best = (unsigned long long*)(counters + 8);
    RB_CUDA(cudaMemsetAsync(best, 0, 12, st));  // best, then the alive flag at counters[10]
    scan_last_match<<<grid_for(n_seg, 256, tuning.blocks_per_sm), 256, smem, st>>>(g, exact, best, counters + 10);
    RB_LAUNCH_CHECK("scan_last_match");
    uint32_t h[4] = {0, 0, 0, 0};
    RB_CUDA(d2h(h, best, 12));
    const bool alive = h[2] != 0;
    if (alive && end < n) continue;
    const uint64_t found = (uint64_t)h[0] | ((uint64_t)h[1] << 32);
    if (alive && text_continues) {
      // the run meets the end of a shard's halo: the same verdict as the walk kernels give (err_flag), read
      // by the caller once the search is through
      const uint32_t one = 1;
      RB_CUDA(cudaMemcpyAsync(counters + 28, &one, 4, cudaMemcpyHostToDevice, st));
      RB_CUDA(cudaStreamSynchronize(st));
      *e_out = found ? found - 1 : n;
      return 0;
    }
    if (found == 0) return fail("internal error: a candidate start has no match (long run)");
    *e_out = found - 1;
    return 0;
  }
}

// ------------------------------------------------- bounded fix-up of entry states --
// Exact entry state of every segment by state-map composition (kernels.cu, "exact entry states"):
// K = the states seen at segment boundaries, closed under the per-segment maps; on return
// fin[] holds, next to every segment, the exact state it is entered with, so that one
// verify + redo round completes the scan.  Used when the plain redo rounds do not converge.
int Regex::solve_entries(const void* scan_args, bool reverse) {
  ScanArgs g = *(const ScanArgs*)scan_args;
  cudaStream_t st = (cudaStream_t)stream_;
  const size_t smem = smem_for(g.dfa);
  g.use_smem = smem != 0;
  g.redo_list = nullptr;
  RB_CUDA(allow_smem(scan_map, smem));
  const uint64_t n_seg = g.n_seg;
  uint8_t* present = (uint8_t*)present_.ensure(65536);
  uint16_t* d_kidx = (uint16_t*)kidx_.ensure(65536 * 2);
  uint32_t* counters = (uint32_t*)counters_.ptr;
  if (!present || !d_kidx) return fail("out of device memory (state maps)");
  RB_CUDA(cudaMemsetAsync(present, 0, 65536, st));
  mark_states<<<grid_for(n_seg, 256, 8), 256, 0, st>>>(g.guess, n_seg, present);
  RB_LAUNCH_CHECK("mark_states");
  mark_states<<<grid_for(n_seg, 256, 8), 256, 0, st>>>(g.fin, n_seg, present);
  RB_LAUNCH_CHECK("mark_states");
  std::vector<uint8_t> h_present(65536);
  std::vector<uint16_t> K, kidx;
  uint16_t *d_states = nullptr, *maps = nullptr;
  for (;;) {
    RB_CUDA(d2h(h_present.data(), present, 65536));
    K.clear();
    kidx.assign(65536, 0xFFFF);
    for (uint32_t v = 0; v < 65536; v++)
      if (h_present[v] && v < g.dfa.n_states) { kidx[v] = (uint16_t)K.size(); K.push_back((uint16_t)v); }
    const uint32_t k = (uint32_t)K.size();
    if (k > 512) return fail("the automaton reaches more than 512 distinct states at segment boundaries; state-map composition refused");
    d_states = (uint16_t*)kstates_.ensure(k * 2);
    maps = (uint16_t*)maps_.ensure(n_seg * k * 2);
    if (!d_states || !maps) return fail("out of device memory (state maps)");
    RB_CUDA(cudaMemcpyAsync(d_states, K.data(), k * 2, cudaMemcpyHostToDevice, st));
    RB_CUDA(cudaMemcpyAsync(d_kidx, kidx.data(), 65536 * 2, cudaMemcpyHostToDevice, st));
    scan_map<<<grid_for(n_seg * k, 256, tuning.blocks_per_sm), 256, smem, st>>>(g, reverse ? 1 : 0, d_states, k, maps);
    RB_LAUNCH_CHECK("scan_map");
    stats.map_passes += k;
    RB_CUDA(cudaMemsetAsync(counters + 8, 0, 4, st));
    closure_check<<<grid_for(n_seg * k, 256, 8), 256, 0, st>>>(maps, n_seg * k, d_kidx, present, counters + 8);
    RB_LAUNCH_CHECK("closure_check");
    uint32_t n_new = 0;
    RB_CUDA(d2h(&n_new, counters + 8, 4));  // also orders the host vectors above before they are rebuilt
    if (n_new == 0) break;
  }
  const uint32_t k = (uint32_t)K.size();
  const uint64_t n_blocks = (n_seg + kMapBlock - 1) / kMapBlock;
  uint16_t* comp = (uint16_t*)comp_.ensure(n_blocks * k * 2);
  uint16_t* bentry = (uint16_t*)bentry_.ensure(n_blocks * 2);
  uint16_t* exact = (uint16_t*)exact_.ensure(n_seg * 2);
  if (!comp || !bentry || !exact) return fail("out of device memory (state maps)");
  compose_blocks<<<(uint32_t)((n_blocks * k + 255) / 256), 256, 0, st>>>(maps, d_kidx, d_states, k, n_seg, reverse ? 1 : 0, comp);
  RB_LAUNCH_CHECK("compose_blocks");
  compose_top<<<1, 32, 0, st>>>(comp, d_kidx, k, n_blocks, reverse ? 1 : 0, reverse ? g.guess + (n_seg - 1) : g.guess, bentry);
  RB_LAUNCH_CHECK("compose_top");
  compose_fill<<<(uint32_t)((n_blocks + 255) / 256), 256, 0, st>>>(maps, d_kidx, k, n_seg, reverse ? 1 : 0, bentry, exact);
  RB_LAUNCH_CHECK("compose_fill");
  publish_exact<<<grid_for(n_seg, 256, 8), 256, 0, st>>>(exact, n_seg, reverse ? 1 : 0, g.fin);
  RB_LAUNCH_CHECK("publish_exact");
  return 0;
}

// ------------------------------------------------------------- start bitmap --
Regex::ScanPlan Regex::plan_scan(const uint8_t* d_text, uint64_t base, uint64_t limit, bool fast_table) {
  ScanPlan p;
  const bool utf8_mask = only_utf8 && can_match_empty;
  p.fast = fast_table && !utf8_mask && ((uintptr_t)d_text & 15) == 0 && !tuning.force_generic;
  // segment length: long enough to amortise the warm-up, short enough to fill the GPU
  uint32_t seg = tuning.seg;
  if (seg == 0) {
    const uint64_t lanes = (uint64_t)std::max(1, device_sm_count()) * 1024 * 2;
    uint64_t want = ((limit - base) / lanes + 255) / 256 * 256;
    seg = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 256), p.fast ? 4096 : 1024);
  }
  p.seg = seg;
  p.n_seg = std::max<uint64_t>(1, (limit - base + seg - 1) / seg);
  p.warm = pick_warm(*this);
  if (p.fast) p.warm = (p.warm + 63) / 64 * 64;
  return p;
}

int Regex::scan_starts(const uint8_t* d_text, uint64_t n, uint64_t base, uint64_t limit, ShardIO* io, const ScanPlan& plan,
                       const void* fused_walk) {
  DeviceDfa* rev;
  if (int rc = ensure(kRevUnanchoredAll, &rev)) return rc;
  cudaStream_t st = (cudaStream_t)stream_;
  const bool utf8_mask = only_utf8 && can_match_empty;
  const bool fast = plan.fast;
  const uint32_t seg = plan.seg;
  const uint64_t n_seg = plan.n_seg;
  if (n_seg >= 0xFFFFFFFFull) return fail("haystack too large for one scan (segment index overflow)");
  ScanArgs a{};
  a.dfa = rev->view;
  a.text = d_text;
  a.n = n;
  a.limit = limit;
  a.base = base;
  a.n_seg = n_seg;
  a.seg = seg;
  a.warm = plan.warm;
  a.bitmap = (uint64_t*)bitmap_.ensure(((n >> 6) + 2) * 8);
  a.guess = (uint16_t*)guess_.ensure(n_seg * 2);
  a.fin = (uint16_t*)fin_.ensure((n_seg + 1) * 2);
  uint32_t* redo = (uint32_t*)redo_.ensure(n_seg * 4);
  uint32_t* counters = (uint32_t*)counters_.ensure(128);
  if (!a.bitmap || !a.guess || !a.fin || !redo || !counters) return fail("out of device memory (scan scratch)");
  a.flag0 = (uint8_t*)(counters + 24);
  a.utf8_boundaries = utf8_mask;
  a.hot = rev->hot;
  size_t smem;
  uint32_t block;
  const WalkArgs* fw = (const WalkArgs*)fused_walk;
  WalkArgs no_walk{};
  if (fast) {
    block = 1024;
    const bool fw_fixed = fw && fw->fixed_len != 0;
    smem = fast_scan_smem(hot_bytes_signed(rev->hot) + (fw && !fw_fixed ? hot_bytes(fw->fwd_hot.n) : 0));
    {  // kernels.cu: rings and mbarriers of the block's warps first, then the tables (hot_signed_below / hot_signed_bytes)
      const uint32_t after_rings = (block / 32) * (2 * 2048 + 64);
      a.tbase_off = after_rings + (uint32_t)hot_bytes(rev->hot.n - rev->hot.match_lo);
      a.fbase_off = after_rings + (uint32_t)hot_bytes_signed(rev->hot);
    }
    if (fw_fixed) RB_CUDA(allow_smem(scan_rev_fast<2>, smem));
    else if (fw) RB_CUDA(allow_smem(scan_rev_fast<1>, smem));
    else RB_CUDA(allow_smem(scan_rev_fast<0>, smem));
  } else {
    smem = smem_for(rev->view);
    a.use_smem = smem != 0;
    block = tuning.block;
    RB_CUDA(allow_smem(scan_rev_bitmap, smem));
  }
  // 2-D view of the haystack for the tiled TMA loads: rows = full segments, cols = bytes
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof tmap);
  if (fast && tuning.tensor_tma && encode_tiled() && seg >= 64 && seg % 16 == 0 && a.warm <= seg) {
    const uint64_t rows = (n - base) / seg;  // only rows that lie entirely inside the buffer
    if (rows >= 34) {
      cuuint64_t dims[2] = {seg, rows};
      cuuint64_t strides[1] = {seg};
      cuuint32_t box[2] = {64, 32};  // one box = the 64-byte group of all 32 lanes' segments
      cuuint32_t estr[2] = {1, 1};
      CUresult r = encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)(d_text + base), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r == CUDA_SUCCESS) a.tmap_rows = rows;
    }
  }
  auto launch = [&](const ScanArgs& args, uint64_t work) {
    if (fast && fw && fw->fixed_len) scan_rev_fast<2><<<grid_for(work, block, 1), block, smem, st>>>(args, *fw, tmap);
    else if (fast && fw) scan_rev_fast<1><<<grid_for(work, block, 1), block, smem, st>>>(args, *fw, tmap);
    else if (fast) scan_rev_fast<0><<<grid_for(work, block, 1), block, smem, st>>>(args, no_walk, tmap);
    else scan_rev_bitmap<<<grid_for(work, block, tuning.blocks_per_sm), block, smem, st>>>(args);
  };
  const bool reuse = io && io->reuse_scan;
  if (!reuse) {
    launch(a, n_seg);
    RB_LAUNCH_CHECK("scan_rev");
  }
  stats.scan_redo_rounds = stats.scan_redo_segments = stats.map_passes = 0;
  auto redo_round = [&](uint32_t n_redo) -> int {
    stats.scan_redo_rounds++;
    stats.scan_redo_segments += n_redo;
    ScanArgs r = a;
    r.redo_list = redo;
    r.n_redo = counters;
    launch(r, n_redo);
    RB_LAUNCH_CHECK("scan_rev(redo)");
    return 0;
  };
  // A shard may be told the exact state at its top edge by its right neighbour.
  uint16_t* h16 = (uint16_t*)pinned_ + 512;
  const bool shard = io && (!io->is_first || !io->is_last || io->rev_entry != kNoState || io->own_hi < n);
  if (shard) {
    RB_CUDA(cudaMemcpyAsync(h16, a.guess + (n_seg - 1), 2, cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    io->rev_guess = h16[0];
    if (io->rev_entry != kNoState && io->rev_entry != io->rev_guess) {
      h16[1] = (uint16_t)io->rev_entry;
      uint32_t* h32 = (uint32_t*)pinned_ + 300;
      h32[0] = (uint32_t)(n_seg - 1);
      h32[1] = 1;
      RB_CUDA(cudaMemcpyAsync(a.fin + n_seg, h16 + 1, 2, cudaMemcpyHostToDevice, st));
      RB_CUDA(cudaMemcpyAsync(redo, h32, 4, cudaMemcpyHostToDevice, st));
      RB_CUDA(cudaMemcpyAsync(counters, h32 + 1, 4, cudaMemcpyHostToDevice, st));
      if (int rc = redo_round(1)) return rc;
      io->rev_guess = io->rev_entry;
    }
  }
  for (uint32_t round = 0;; round++) {
    RB_CUDA(cudaMemsetAsync(counters, 0, 4, st));
    verify_segments<<<grid_for(n_seg, 256, 8), 256, 0, st>>>(a.guess, a.fin, n_seg, 1, redo, counters);
    RB_LAUNCH_CHECK("verify_segments");
    RB_CUDA(cudaMemcpyAsync(pinned_, counters, 4, cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    const uint32_t n_redo = *(uint32_t*)pinned_;
    if (n_redo == 0) break;
    if (round == tuning.max_redo_rounds) {
      // wrong guesses keep cascading: solve every entry state at once, then one more redo round
      if (int rc = solve_entries(&a, true)) return rc;
      continue;
    }
    if (int rc = redo_round(n_redo)) return rc;
  }
  if (shard) {
    RB_CUDA(cudaMemcpyAsync(h16, a.fin, 2, cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    io->rev_left = h16[0];
  }
  return 0;
}

// ------------------------------------------------------------------ find_all --
int Regex::find_all_device(const uint8_t* d_text, uint64_t n, uint64_t start, uint64_t* d_out, uint64_t cap, uint64_t* total) {
  *total = 0;
  if (start > n) return 0;  // re_trait.rs:198-200
  ShardIO io;
  io.own_lo = start ? ((start - 1) & ~255ull) : 0;  // bit i <-> position i+1
  io.own_hi = n;
  io.chain_p = start;
  io.chain_lm = kNone;
  int rc = find_all_shard_device(d_text, n, &io, d_out, cap);
  *total = io.n_matches;
  return rc;
}

int Regex::find_all_shard_device(const uint8_t* d_text, uint64_t n, ShardIO* io, uint64_t* d_out, uint64_t cap) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  io->n_matches = 0;
  io->halo_overflow = false;
  io->left_ctx_short = false;
  if (is_set_) return fail("find requires exactly one pattern (RegexSet cannot be used with find, exec.rs:510-512)");
  if (io->own_hi > n || io->own_lo > io->own_hi || (io->own_lo & 255) || (!io->is_last && ((io->own_hi & 255) || io->own_hi == n)))
    return fail("bad shard geometry: own_lo/own_hi must be multiples of 256 inside the buffer (own_hi == n only for the last shard)");
  DeviceDfa *fwd, *rev = nullptr;
  if (int rc = ensure(kFwdAnchoredLF, &fwd)) return rc;
  const bool emulate = has_looks;
  if (emulate) if (int rc = ensure(kRevAnchoredLongest, &rev)) return rc;
  DeviceDfa* revall;
  if (int rc = ensure(kRevUnanchoredAll, &revall)) return rc;
  cudaStream_t st = (cudaStream_t)stream_;
  cudaEvent_t ev[3] = {(cudaEvent_t)timing_events_[0], (cudaEvent_t)timing_events_[1], (cudaEvent_t)timing_events_[2]};
  RB_CUDA(cudaEventRecord(ev[0], st));
  const ScanPlan plan = plan_scan(d_text, io->own_lo, io->own_hi,
                                   hot_signed_ok(revall->hot) && fast_scan_smem(hot_bytes_signed(revall->hot)) <= 227 * 1024);
  // runner: 2 = fixed-length (no haystack access), 1 = byte-indexed shared-memory table
  // (uniform start state, 8-byte aligned text), 0 = generic
  const bool wfixed = min_len == max_len && min_len > 0 && !emulate && !tuning.force_generic;
  const bool wfast = !wfixed && fwd->hot.n != 0 && fwd->view.uniform_start && ((uintptr_t)d_text & 15) == 0 && !tuning.force_generic;
  const int wkind = wfixed ? 2 : wfast ? 1 : 0;
  // literal prefilter: no start bitmap at all -- literal_scan finds, verifies and chains the candidates
  const bool use_pf = tuning.prefilter && !tuning.force_generic && ((uintptr_t)d_text & 15) == 0 && plan_prefilter() &&
                      (tuning.prefilter >= 2 || pf_freq_ <= kPrefilterAutoFreq);
  PfArgs pf{};
  if (use_pf) std::memcpy(&pf, pf_words_.data(), sizeof pf);
  const bool pf_fast = use_pf && fwd->hot.n != 0 && fwd->view.uniform_start;
  // fused: every lane of the fast scan kernel also walks its own segment (chunk == segment)
  const bool fused = !use_pf && plan.fast && wkind != 0 && tuning.fuse && !io->reuse_scan &&
                     fast_scan_smem(hot_bytes_signed(revall->hot) + (wkind == 1 ? hot_bytes(fwd->hot.n) : 0)) <= 227 * 1024;

  WalkArgs w{};
  w.fwd = fwd->view;
  if (rev) w.rev = rev->view;
  w.text = d_text;
  w.n = n;
  if (!bitmap_.ensure(((n >> 6) + 2) * 8) || !counters_.ensure(128)) return fail("out of device memory (bitmap)");
  w.bitmap = (const uint64_t*)bitmap_.ptr;
  w.flag0 = (const uint8_t*)((uint32_t*)counters_.ptr + 24);
  w.base = io->own_lo;
  w.limit = io->own_hi;
  w.text_continues = !io->is_last;
  w.chunk = use_pf ? 8192 : fused ? plan.seg : std::max<uint32_t>(256, (tuning.chunk + 255) / 256 * 256);
  w.stage_cap = std::max<uint32_t>(4, w.chunk / 64);
  w.n_chunks = std::max<uint64_t>(1, (w.limit - w.base + w.chunk - 1) / w.chunk);
  const uint64_t nc = w.n_chunks;
  w.in_p = (uint64_t*)in_p_.ensure(nc * 8);
  w.in_lm = (uint64_t*)in_lm_.ensure(nc * 8);
  w.out_p = (uint64_t*)out_p_.ensure(nc * 8);
  w.out_lm = (uint64_t*)out_lm_.ensure(nc * 8);
  w.count = (uint64_t*)count_.ensure(nc * 8);
  uint64_t* offset = (uint64_t*)offset_.ensure(nc * 8);
  uint32_t* dirty_list = (uint32_t*)dirty_.ensure(nc * 4);
  w.first_cand = (uint64_t*)first_cand_.ensure(nc * 8);
  w.skip = (uint32_t*)skip_.ensure(nc * 4);
  w.meta = (uint32_t*)meta_.ensure(nc * 4);
  w.stage = (uint64_t*)stage_.ensure(nc * (uint64_t)w.stage_cap * 16);
  const uint64_t n_blocks = (nc + 1023) / 1024;
  uint64_t* block_sums = (uint64_t*)block_sums_.ensure(n_blocks * 8);
  uint32_t* counters = (uint32_t*)counters_.ensure(128);
  if (!w.in_p || !w.in_lm || !w.out_p || !w.out_lm || !w.count || !offset || !dirty_list || !w.first_cand || !w.skip || !w.meta || !w.stage || !block_sums || !counters)
    return fail("out of device memory (walk scratch)");
  w.offset = offset;
  w.out = d_out;
  w.cap = d_out ? cap : 0;
  w.utf8 = only_utf8;
  w.emulate_slice = emulate;
  w.can_match_empty = can_match_empty;
  // entry states: chunk 0 starts the real chain at `start`; the rest speculate.
  w.err_flag = counters + 28;
  w.floor_flag = counters + 29;
  // long anchored runs go to resolve_long_run (automata of up to 128 states); counters[6..7] = the request slot
  const bool can_resolve = fwd->view.n_states <= 128;
  uint64_t* long_tab = (uint64_t*)long_tab_.ensure(kMaxLongRuns * 16);
  if (!long_tab) return fail("out of device memory (long runs)");
  uint32_t n_long = 0;
  w.exact_cap = can_resolve ? kExactRunCap : kNone;
  w.long_tab = long_tab;
  w.n_long = 0;
  w.long_req = (unsigned long long*)(counters + 6);
  RB_CUDA(cudaMemsetAsync(w.long_req, 0xFF, 8, st));
  w.clamp_p = io->chain_clamped ? io->chain_p : kNone;
  RB_CUDA(cudaMemsetAsync(w.err_flag, 0, 8, st));
  init_walk_entries<<<grid_for(nc, 256, 8), 256, 0, st>>>(w.in_p, w.in_lm, w.skip, nc, io->chain_p, io->chain_lm);
  RB_LAUNCH_CHECK("init_walk_entries");
  size_t wsmem = 0;
  if (wfixed) {
    w.fixed_len = min_len;
  } else if (wfast) {
    wsmem = hot_bytes(fwd->hot.n) + 256;
    w.fwd_hot = fwd->hot;
    RB_CUDA(allow_smem(walk_chunks<1>, wsmem));
    RB_CUDA(allow_smem(compact_spans<1>, wsmem));
  } else {
    wsmem = smem_for(fwd->view);
    w.use_smem = wsmem != 0;
    RB_CUDA(allow_smem(walk_chunks<0>, wsmem));
    RB_CUDA(allow_smem(compact_spans<0>, wsmem));
  }
  if (use_pf) {  // the prefilter's verifier: table runner (shared-memory hot table) or generic
    w.fixed_len = 0;
    if (pf_fast) {
      wsmem = hot_bytes(fwd->hot.n) + 256;
      w.fwd_hot = fwd->hot;
    } else {
      wsmem = smem_for(fwd->view);
      w.use_smem = wsmem != 0;
    }
  }
  auto launch_pf = [&](const WalkArgs& args, uint64_t chunks, int mode) -> cudaError_t {
    const uint32_t g = mode == 2 ? 1 : grid_for(chunks * 32, 256, 3);  // one warp per chunk
#define RB_PF(F, N)                                                                \
    {                                                                              \
      /* the kernel also has ~9 KB of static shared memory: opt in whenever the sum could pass 48 KB */ \
      cudaError_t e__ = allow_smem(literal_scan<F, N>, wsmem, 48 * 1024); \
      if (e__ != cudaSuccess) return e__;                                          \
      literal_scan<F, N><<<g, 256, wsmem, st>>>(args, pf, mode);                   \
      return cudaSuccess;                                                          \
    }
    if (pf_fast) switch (pf.n_bytes) { case 1: RB_PF(1, 1) case 2: RB_PF(1, 2) case 3: RB_PF(1, 3) default: RB_PF(1, 4) }
    switch (pf.n_bytes) { case 1: RB_PF(0, 1) case 2: RB_PF(0, 2) case 3: RB_PF(0, 3) default: RB_PF(0, 4) }
#undef RB_PF
  };
  auto launch_walk = [&](const WalkArgs& args, uint64_t work) {
    if (use_pf) { (void)launch_pf(args, work, 1); return; }
    const uint32_t g = grid_for(work, 256, 6);
    if (wkind == 2) walk_chunks<2><<<g, 256, 0, st>>>(args);
    else if (wkind == 1) walk_chunks<1><<<g, 256, wsmem, st>>>(args);
    else walk_chunks<0><<<g, 256, wsmem, st>>>(args);
  };
  if (use_pf) {
    RB_CUDA(launch_pf(w, nc, 0));
    RB_LAUNCH_CHECK("literal_scan");
    io->rev_guess = io->rev_entry != kNoState ? io->rev_entry : kPrefilterState;  // no reverse scan, nothing to guess
    io->rev_left = kPrefilterState;
    stats.scan_redo_rounds = stats.scan_redo_segments = stats.map_passes = 0;
    RB_CUDA(cudaEventRecord(ev[1], st));
  } else {
    if (int rc = scan_starts(d_text, n, io->own_lo, io->own_hi, io, plan, fused ? &w : nullptr)) return rc;
    RB_CUDA(cudaEventRecord(ev[1], st));
    if (!fused) {
      launch_walk(w, nc);
      RB_LAUNCH_CHECK("walk_chunks");
    }
  }
  // ---- stitch: bring the speculative chunk walks into agreement with the sequential iterator ----
  stats.stitch_rounds = stats.stitch_dirty_chunks = stats.sequential_passes = stats.long_runs = 0;
  const bool strict = emulate || can_match_empty;
  bool general = strict;
  uint32_t* hc = (uint32_t*)pinned_ + 64;  // host copy of counters[0..3]
  if (!strict) {
    RB_CUDA(cudaMemsetAsync(counters, 0, 16, st));
    stitch_fast<<<grid_for(nc, 256, 8), 256, 0, st>>>(w, counters);
    RB_LAUNCH_CHECK("stitch_fast");
    RB_CUDA(cudaMemcpyAsync(hc, counters, 16, cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    general = hc[2] != 0;  // (a deferred chunk -- long run -- asks for the general loop too)
  }
  ChainKey* grand_key = nullptr;
  if (general) {
    ChainKey* excl = (ChainKey*)excl_.ensure(nc * sizeof(ChainKey));
    ChainKey* btot = (ChainKey*)btot_.ensure((n_blocks + 1) * sizeof(ChainKey));
    if (!excl || !btot) return fail("out of device memory (stitch scratch)");
    grand_key = btot + n_blocks;
    WalkArgs wd = w;
    wd.dirty_list = dirty_list;
    wd.n_dirty = counters;
    uint32_t dirty_rounds = 0;
    for (;;) {
      RB_CUDA(cudaMemsetAsync(counters + 12, 0xFF, 4, st));  // [12] leftmost deferred chunk
      entries_local<<<(uint32_t)n_blocks, 1024, 0, st>>>(w, excl, btot, counters + 12);
      RB_LAUNCH_CHECK("entries_local");
      entries_blocks<<<1, 1024, 0, st>>>(btot, n_blocks, grand_key);
      RB_LAUNCH_CHECK("entries_blocks");
      RB_CUDA(cudaMemsetAsync(counters, 0, 12, st));       // [0] chunks to walk again, [1] changed decisions
      RB_CUDA(cudaMemsetAsync(counters + 3, 0xFF, 4, st));  // [3] smallest chunk index to walk again
      stitch_resolve<<<grid_for(nc, 256, 8), 256, 0, st>>>(wd, excl, btot, counters);
      RB_LAUNCH_CHECK("stitch_resolve");
      RB_CUDA(cudaMemcpyAsync(hc, counters, 32, cudaMemcpyDeviceToHost, st));
      RB_CUDA(cudaStreamSynchronize(st));
      const uint32_t n_dirty = hc[0], n_changed = hc[1];
      const uint64_t long_s = *(const uint64_t*)(hc + 6);
      bool serve = long_s != kNone;
      if (serve) {
        // the request of a chunk that this round found covered by an earlier match is dropped
        uint32_t m = 0;
        RB_CUDA(d2h(&m, w.meta + (long_s - w.base) / w.chunk, 4));
        serve = (m >> 30) == kChunkDeferred;
        if (!serve) RB_CUDA(cudaMemsetAsync(w.long_req, 0xFF, 8, st));
      }
      if (serve) {
        // a chunk walked from its exact entry met a match longer than kExactRunCap: measure it in parallel,
        // remember it, and let the stitch walk the deferred chunk again
        if (n_long == kMaxLongRuns) return fail("more than 256 matches longer than 256 KiB in one search");
        uint64_t e = kNone;
        if (int rc = resolve_long_run(d_text, n, long_s, !io->is_last, &e)) return rc;
        const uint64_t pair[2] = {long_s, e};
        RB_CUDA(cudaMemcpyAsync(long_tab + 2 * n_long, pair, 16, cudaMemcpyHostToDevice, st));
        RB_CUDA(cudaMemsetAsync(w.long_req, 0xFF, 8, st));
        RB_CUDA(cudaStreamSynchronize(st));
        n_long++;
        w.n_long = wd.n_long = n_long;
        stats.long_runs++;
      }
      if (n_dirty == 0 && n_changed == 0 && !serve) break;
      stats.stitch_rounds++;
      stats.stitch_dirty_chunks += n_dirty;
      if (n_dirty == 0) continue;
      if (++dirty_rounds > tuning.max_stitch_rounds) {
        // chains that do not meet again: one sequential pass from the leftmost such chunk
        wd.seq_from = hc[3];
        if (use_pf) (void)launch_pf(wd, 1, 2);
        else if (wkind == 2) walk_sequential<2><<<1, 256, 0, st>>>(wd);
        else if (wkind == 1) walk_sequential<1><<<1, 256, wsmem, st>>>(wd);
        else walk_sequential<0><<<1, 256, wsmem, st>>>(wd);
        RB_LAUNCH_CHECK("walk_sequential");
        stats.sequential_passes++;
        dirty_rounds = 0;
      } else {
        launch_walk(wd, n_dirty);
        RB_LAUNCH_CHECK("walk_chunks(dirty)");
      }
    }
  }
  unsigned long long* grand = (unsigned long long*)(counters + 4);
  scan_counts_local<<<(uint32_t)n_blocks, 1024, 0, st>>>(w.count, offset, block_sums, nc);
  RB_LAUNCH_CHECK("scan_counts_local");
  scan_block_sums<<<1, 1024, 0, st>>>(block_sums, n_blocks, grand);
  RB_LAUNCH_CHECK("scan_block_sums");
  scan_add_block_offsets<<<(uint32_t)n_blocks, 1024, 0, st>>>(offset, block_sums, nc);
  RB_LAUNCH_CHECK("scan_add_block_offsets");
  if (w.cap > 0 && use_pf) {
    compact_staged<<<grid_for(nc, 256, 6), 256, 0, st>>>(w);
    RB_LAUNCH_CHECK("compact_staged");
    RB_CUDA(launch_pf(w, nc, 3));  // chunks with more matches than staging slots: straight into the output
    RB_LAUNCH_CHECK("literal_scan(overflow)");
  } else if (w.cap > 0) {
    const uint32_t g = grid_for(nc, 256, 6);
    if (wkind == 2) compact_spans<2><<<g, 256, 0, st>>>(w);
    else if (wkind == 1) compact_spans<1><<<g, 256, wsmem, st>>>(w);
    else compact_spans<0><<<g, 256, wsmem, st>>>(w);
    RB_LAUNCH_CHECK("compact_spans");
  }
  uint64_t* h = (uint64_t*)pinned_;
  RB_CUDA(cudaMemcpyAsync(h, grand, 8, cudaMemcpyDeviceToHost, st));
  if (grand_key) {  // exit state = the last contribution of the whole range
    RB_CUDA(cudaMemcpyAsync(h + 1, grand_key, 16, cudaMemcpyDeviceToHost, st));
  } else {
    RB_CUDA(cudaMemcpyAsync(h + 1, w.out_p + (nc - 1), 8, cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaMemcpyAsync(h + 2, w.out_lm + (nc - 1), 8, cudaMemcpyDeviceToHost, st));
  }
  RB_CUDA(cudaMemcpyAsync(h + 3, w.err_flag, 8, cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaEventRecord(ev[2], st));
  RB_CUDA(cudaStreamSynchronize(st));
  io->n_matches = h[0];
  if (grand_key) {  // ChainKey: 0 = nothing in this range moved the iterator (speculative shard without candidates)
    io->exit_p = h[1] == 0 ? kSpec : (h[1] == ~0ull ? kNone : h[1] - 1);
    io->exit_lm = h[1] == 0 ? kNone : h[2];
  } else {
    io->exit_p = h[1];
    io->exit_lm = h[2];
  }
  io->halo_overflow = ((uint32_t*)(h + 3))[0] != 0;
  io->left_ctx_short = ((uint32_t*)(h + 3))[1] != 0;
  stats.fused = fused;
  stats.path = use_pf ? 3 : fused ? 2 : plan.fast ? 1 : 0;
  cudaEventElapsedTime(&stats.scan_ms, ev[0], ev[1]);
  cudaEventElapsedTime(&stats.walk_ms, ev[1], ev[2]);
  cudaEventElapsedTime(&stats.total_ms, ev[0], ev[2]);
  if (io->halo_overflow) return fail("a match runs past the end of the shard halo; enlarge the halo");
  if (io->left_ctx_short)
    return fail("the reverse-on-slice scan of a look-around pattern reaches the start of the shard buffer; enlarge the left context");
  return 0;
}

int Regex::find_at_device(const uint8_t* d_text, uint64_t n, uint64_t start, bool* found, uint64_t* s, uint64_t* e) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *found = false;
  uint64_t* d_out = (uint64_t*)out_.ensure(16);
  if (!d_out) return fail("out of device memory");
  uint64_t total = 0;
  // The first element of the find_iter chain started at `start` is find_at(start).
  if (int rc = find_all_device(d_text, n, start, d_out, 1, &total)) return rc;
  if (total == 0) return 0;
  uint64_t h[2];
  RB_CUDA(cudaMemcpyAsync(h, d_out, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
  RB_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  *found = true;
  *s = h[0];
  *e = h[1];
  return 0;
}

// ------------------------------------------------------ forward reductions ----
// One wave of a forward all-match scan: positions [start, limit) of the haystack (limit == n + 1
// includes the end-of-text step), entered in state `entry` (kNoEntry: the start state for the
// flags at `start`).  result_host[0] = first match end, [1..] = OR of the pattern masks;
// *exit_state = exact state after the last position (the next wave's entry).
int Regex::forward_range(const uint8_t* d_text, uint64_t n, uint64_t start, uint64_t limit, uint32_t entry, bool want_masks,
                         uint64_t* result_host, uint32_t* exit_state) {
  DeviceDfa* fwd;
  if (int rc = ensure(kFwdUnanchoredAll, &fwd)) return rc;
  cudaStream_t st = (cudaStream_t)stream_;
  // fast path (scan_fwd_fast): hot table, 16-byte aligned haystack and start; RegexSets only when
  // matches are known to be rare (a narrowed set, see forward_reduce): the fast kernel finds the
  // patterns of a match by redoing that 64-byte group on the full table
  const bool fast_ok = (!want_masks || sparse_set_) && fwd->hot.n != 0 && !tuning.force_generic && tuning.tensor_tma && encode_tiled() &&
                       ((uintptr_t)d_text & 15) == 0 && (start & 63) == 0 &&
                       fast_scan_smem(hot_bytes(fwd->hot.n)) <= 227 * 1024;
  const uint32_t seg = tuning.seg ? tuning.seg : (fast_ok ? 4096 : 1024);
  const uint64_t n_seg = (limit - start + seg - 1) / seg;
  if (n_seg >= 0xFFFFFFFFull) return fail("haystack too large for one scan (segment index overflow)");
  const uint32_t mw = fwd->view.mask_words;
  ScanArgs a{};
  a.dfa = fwd->view;
  const size_t smem = smem_for(fwd->view);
  a.use_smem = smem != 0;
  a.text = d_text;
  a.n = n;
  a.base = start;
  a.fwd_limit = limit;
  a.entry0 = entry;
  a.n_seg = n_seg;
  a.seg = seg;
  a.warm = pick_warm(*this);
  if (fast_ok) a.warm = (a.warm + 63) / 64 * 64;
  a.seg_first = (uint64_t*)seg_first_.ensure(n_seg * 8);
  a.seg_mask = want_masks ? (uint64_t*)seg_mask_.ensure(n_seg * 8 * mw) : nullptr;
  a.guess = (uint16_t*)guess_.ensure(n_seg * 2);
  a.fin = (uint16_t*)fin_.ensure(n_seg * 2);
  uint32_t* redo = (uint32_t*)redo_.ensure(n_seg * 4);
  uint32_t* counters = (uint32_t*)counters_.ensure(128);
  if (!a.seg_first || (want_masks && !a.seg_mask) || !a.guess || !a.fin || !redo || !counters)
    return fail("out of device memory (scan scratch)");
  RB_CUDA(allow_smem(scan_fwd_reduce, smem));
  // whole warps of full segments go to the fast kernel; segment 0, the ragged end and the EOF step stay generic
  const uint64_t n_full = (std::min(limit, n) - start) / seg;
  bool fast_launched = false;
  if (!fork_event_) {
    cudaEvent_t e1, e2;
    RB_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
    RB_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    fork_event_ = e1;
    join_event_ = e2;
  }
  if (fast_ok && seg % 64 == 0 && a.warm <= seg && n_full >= 34) {
    const uint64_t skip_lo = 1, skip_hi = 1 + (n_full - 1) / 32 * 32;
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof tmap);
    cuuint64_t dims[2] = {seg, skip_hi - skip_lo + 1};  // row 0 = segment skip_lo - 1 (warm-up source of the first warp)
    cuuint64_t strides[1] = {seg};
    cuuint32_t box[2] = {64, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)(d_text + start + (skip_lo - 1) * (uint64_t)seg), dims, strides,
                                box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) {
      a.hot = fwd->hot;
      a.skip_lo = skip_lo;
      a.skip_hi = skip_hi;
      const size_t fsm = fast_scan_smem(hot_bytes(fwd->hot.n));
      RB_CUDA(allow_smem(scan_fwd_fast, fsm));
      RB_CUDA(cudaEventRecord((cudaEvent_t)fork_event_, st));  // the side stream starts once everything before this point is done
      scan_fwd_fast<<<grid_for(skip_hi - skip_lo, 1024, 1), 1024, fsm, st>>>(a, tmap);
      RB_LAUNCH_CHECK("scan_fwd_fast");
      fast_launched = true;
    }
  }
  // The generic kernel takes what the fast one leaves (segment 0, the ragged end, the end-of-text
  // step): a handful of 4 KiB segments walked by single threads, ~0.25 ms of pure latency (ncu
  // profiles/r02: 10 % issue, 3 % of the warps).  It runs beside the fast kernel on a second stream.
  if (fast_launched) {
    if (!copy_stream_) {
      cudaStream_t s1, s2;
      RB_CUDA(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
      RB_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
      copy_stream_ = s1;
      back_stream_ = s2;
    }
    cudaStream_t side = (cudaStream_t)back_stream_;
    RB_CUDA(cudaStreamWaitEvent(side, (cudaEvent_t)fork_event_, 0));
    scan_fwd_reduce<<<grid_for(n_seg, tuning.block, tuning.blocks_per_sm), tuning.block, smem, side>>>(a);
    RB_LAUNCH_CHECK("scan_fwd_reduce");
    RB_CUDA(cudaEventRecord((cudaEvent_t)join_event_, side));
    RB_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)join_event_, 0));
  } else {
    scan_fwd_reduce<<<grid_for(n_seg, tuning.block, tuning.blocks_per_sm), tuning.block, smem, st>>>(a);
    RB_LAUNCH_CHECK("scan_fwd_reduce");
  }
  for (uint32_t round = 0;; round++) {
    RB_CUDA(cudaMemsetAsync(counters, 0, 4, st));
    verify_segments<<<grid_for(n_seg, 256, 8), 256, 0, st>>>(a.guess, a.fin, n_seg, 0, redo, counters);
    RB_LAUNCH_CHECK("verify_segments");
    RB_CUDA(cudaMemcpyAsync(pinned_, counters, 4, cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    const uint32_t n_redo = *(uint32_t*)pinned_;
    if (n_redo == 0) break;
    stats.scan_redo_rounds++;
    stats.scan_redo_segments += n_redo;
    if (round == tuning.max_redo_rounds) {
      if (int rc = solve_entries(&a, false)) return rc;
      continue;
    }
    ScanArgs r = a;
    r.redo_list = redo;
    r.n_redo = counters;
    scan_fwd_reduce<<<grid_for(n_redo, tuning.block, tuning.blocks_per_sm), tuning.block, smem, st>>>(r);
    RB_LAUNCH_CHECK("scan_fwd_reduce(redo)");
  }
  unsigned long long* result = (unsigned long long*)(counters + 4);
  uint64_t* h = (uint64_t*)pinned_ + 16;
  h[0] = kNone;
  for (uint32_t w = 0; w < kMaxMaskWords; w++) h[1 + w] = 0;
  RB_CUDA(cudaMemcpyAsync(result, h, 8 * (1 + kMaxMaskWords), cudaMemcpyHostToDevice, st));
  reduce_segments<<<grid_for(n_seg, 256, 4), 256, 0, st>>>(a.seg_first, a.seg_mask, n_seg, mw, result);
  RB_LAUNCH_CHECK("reduce_segments");
  RB_CUDA(cudaMemcpyAsync(h, result, 8 * (1 + kMaxMaskWords), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaMemcpyAsync(h + 8, a.fin + (n_seg - 1), 2, cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  for (uint32_t w = 0; w < 1 + kMaxMaskWords; w++) result_host[w] = h[w];
  *exit_state = *(uint16_t*)(h + 8);
  return 0;
}

// is_match / shortest_match / RegexSet::matches over one haystack, in waves of growing size
// (64 MiB, x16 each) so that the search stops early the way the reference does:
//   - a single pattern stops at the first wave that holds a match (dfa.rs:658-667, quit_after_match);
//   - a set stops once every pattern has matched (dfa.rs:675-682);
//   - a set ALSO narrows: patterns that matched in the first waves need no further tracking,
//     so the rest of the haystack is scanned with the automaton of the patterns that are still
//     open (usually a handful of rare ones: small, shared-memory resident, fast kernel) instead
//     of the product automaton of all of them.  That automaton knows nothing about the bytes
//     already passed, so it scans from `start` again; narrowing happens after the first two
//     waves only, which bounds the repeated work to 1/16 of the haystack.
// end: positions [start, end) are searched (n + 1 = to the end of the text, the default); entry:
// exact automaton state at `start` (a shard, kNoEntry = a fresh search); *exit_state: the exact
// state after the last position (what the right-hand shard must be entered with).
int Regex::forward_reduce(const uint8_t* d_text, uint64_t n, uint64_t start, bool want_masks, uint64_t* result_host, uint64_t end,
                          uint32_t entry, uint32_t* exit_out) {
  result_host[0] = kNone;
  for (uint32_t w = 0; w < kMaxMaskWords; w++) result_host[1 + w] = 0;
  const uint32_t n_pat = (uint32_t)patterns_.size();
  uint64_t wave = tuning.wave0 ? (tuning.wave0 + 4095) / 4096 * 4096 : 0;
  uint64_t lo = start;
  end = std::min(end, n + 1);
  const bool whole = end == n + 1 && entry == kNoEntry;  // narrowing restarts the search at `start` with a fresh automaton
  if (exit_out) *exit_out = entry == kNoEntry ? kNoState : entry;
  stats.scan_redo_rounds = stats.scan_redo_segments = stats.map_passes = stats.waves = 0;
  for (uint32_t wi = 0;; wi++) {
    uint64_t limit = end;
    if (wave && end - lo > 2 * wave) limit = (lo + wave) / 4096 * 4096;
    uint64_t r[1 + kMaxMaskWords];
    uint32_t exit_state = 0;
    if (lazy_) {  // host haystack: bring in what this wave reads (plus a little for the flags at its end)
      const uint64_t want = std::min(lazy_->n, limit + 64);
      if (want > lazy_->done) {
        RB_CUDA(cudaMemcpyAsync(lazy_->dst + lazy_->done, lazy_->src + lazy_->done, want - lazy_->done, cudaMemcpyHostToDevice, (cudaStream_t)stream_));
        lazy_->done = want;
      }
    }
    if (int rc = forward_range(d_text, n, lo, limit, entry, want_masks, r, &exit_state)) return rc;
    stats.waves++;
    result_host[0] = std::min(result_host[0], r[0]);
    bool fresh = false;
    for (uint32_t w = 0; w < kMaxMaskWords; w++) {
      fresh = fresh || (r[1 + w] & ~result_host[1 + w]);
      result_host[1 + w] |= r[1 + w];
    }
    if (exit_out) *exit_out = exit_state;
    if (limit == end) break;
    if (!want_masks && result_host[0] != kNone) break;
    if (exit_state == 0) break;  // dead state: an anchored search that can no longer match
    if (want_masks) {
      std::vector<uint32_t> open;
      for (uint32_t i = 0; i < n_pat; i++)
        if (!((result_host[1 + i / 64] >> (i % 64)) & 1)) open.push_back(i);
      if (open.empty()) break;
      if (tuning.narrow_sets && whole && wi < 2 && fresh && open.size() < n_pat) {
        Regex* sub = nullptr;
        if (int rc = subset(open, &sub)) return rc;
        uint64_t sr[1 + kMaxMaskWords];
        sub->tuning.wave0 = tuning.wave0;
        sub->set_stream(stream_);
        sub->lazy_ = lazy_;
        std::lock_guard<std::recursive_mutex> lock(sub->mu_);
        const int src = sub->forward_reduce(d_text, n, start, true, sr);
        sub->lazy_ = nullptr;
        if (src) return fail(sub->last_error());
        stats.waves += sub->stats.waves;
        for (uint32_t j = 0; j < open.size(); j++)
          if ((sr[1 + j / 64] >> (j % 64)) & 1) result_host[1 + open[j] / 64] |= 1ull << (open[j] % 64);
        break;
      }
    }
    lo = limit;
    entry = exit_state;
    wave *= 16;
  }
  return 0;
}

// The RegexSet of a subset of this set's patterns (cached; used by forward_reduce's narrowing).
int Regex::subset(const std::vector<uint32_t>& members, Regex** out) {
  auto it = subsets_.find(members);
  if (it == subsets_.end()) {
    std::vector<std::string> pats;
    for (uint32_t i : members) pats.push_back(patterns_[i]);
    CompileOptions o = opt_;
    o.as_set = true;
    rb::Error err;
    std::unique_ptr<Regex> sub(Regex::compile(pats, o, &err));
    if (!sub) return fail("narrowed pattern set: " + err.msg);
    sub->sparse_set_ = true;
    sub->tuning = tuning;
    it = subsets_.emplace(members, std::move(sub)).first;
  }
  *out = it->second.get();
  return 0;
}


int Regex::shortest_match_device(const uint8_t* d_text, uint64_t n, uint64_t start, bool* found, uint64_t* end) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *found = false;
  if (patterns_.empty() || start > n) return 0;
  uint64_t res[1 + kMaxMaskWords];
  if (int rc = forward_reduce(d_text, n, start, false, res)) return rc;
  if (res[0] != kNone) { *found = true; *end = res[0]; }
  return 0;
}

int Regex::set_matches_device(const uint8_t* d_text, uint64_t n, uint64_t start, bool* any, uint64_t* masks) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *any = false;
  const uint32_t mw = (uint32_t)((patterns_.size() + 63) / 64);
  for (uint32_t w = 0; w < mw; w++) masks[w] = 0;
  if (patterns_.empty() || start > n) return 0;
  uint64_t res[1 + kMaxMaskWords];
  if (int rc = forward_reduce(d_text, n, start, true, res)) return rc;
  for (uint32_t w = 0; w < mw; w++) { masks[w] = res[1 + w]; if (res[1 + w]) *any = true; }
  return 0;
}

// One byte-range shard of a forward search (is_match / shortest_match / RegexSet::matches over a
// sharded haystack).  The automaton state flows left to right: a shard that is not
// told its entry state guesses it by running the automaton over the left context in front of
// own_lo (host side, <= 256 bytes) from the start state that belongs to that position; the
// ranks compare each guess with the left neighbour's exact exit state and search again when
// they differ (regex_b200/sharded.py: forward_sharded).
int Regex::forward_shard_device(const uint8_t* d_text, uint64_t n, uint64_t own_lo, uint64_t own_hi, bool is_first, bool is_last, uint32_t entry,
                                bool want_masks, bool* found, uint64_t* first_end, uint64_t* masks, uint32_t* entry_used,
                                uint32_t* exit_state) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *found = false;
  *first_end = kNone;
  const uint32_t mw = (uint32_t)std::max<size_t>(1, (patterns_.size() + 63) / 64);
  for (uint32_t w = 0; w < mw; w++) masks[w] = 0;
  *entry_used = *exit_state = entry;
  if (patterns_.empty()) return 0;
  if (own_lo > own_hi || own_hi > n) return fail("bad shard geometry");
  DeviceDfa* fwd;
  if (int rc = ensure(kFwdUnanchoredAll, &fwd)) return rc;
  if (entry == kNoEntry && own_lo > 0) {
    rb::Error err;
    const rb::Dfa* h = host_dfa(kFwdUnanchoredAll, &err);
    if (!h) return fail(err.msg);
    // run the automaton over the last <= 255 bytes in front of own_lo; its start state takes the
    // flags of a position inside a long text (dfa.rs:1415-1434), so one more byte is looked at
    uint64_t from = own_lo > 255 ? own_lo - 255 : 0;
    if (from == 0 && !is_first) from = 1;            // buffer byte 0 is not the start of the haystack
    const uint64_t copy_lo = from ? from - 1 : 0;
    const uint64_t copy_n = std::min<uint64_t>(n, own_lo + 1) - copy_lo;
    std::vector<uint8_t> left(copy_n + 1, 0);
    RB_CUDA(d2h(left.data(), d_text + copy_lo, copy_n));
    uint32_t st = h->start[rb::start_flag_index_forward(left.data(), copy_n, from - copy_lo) & 127];
    for (uint64_t i = from; i < own_lo; i++) st = h->next((uint16_t)st, left[i - copy_lo]);
    entry = st;
  }
  *entry_used = entry;
  uint64_t res[1 + kMaxMaskWords];
  const uint64_t end = is_last ? n + 1 : own_hi;
  if (int rc = forward_reduce(d_text, n, own_lo, want_masks, res, end, entry, exit_state)) return rc;
  if (res[0] != kNone) { *found = true; *first_end = res[0]; }
  for (uint32_t w = 0; w < mw; w++) { masks[w] = res[1 + w]; if (want_masks && res[1 + w]) *found = true; }
  return 0;
}

// -------------------------------------------------------------------- batch ----
int Regex::is_match_batch_device(const uint8_t* d_text, const uint64_t* d_offsets, uint64_t n_rec, uint32_t* d_bits) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (n_rec == 0) return 0;
  if (patterns_.empty()) {
    if (int rc = init_device()) return rc;
    RB_CUDA(cudaMemsetAsync(d_bits, 0, (n_rec + 31) / 32 * 4, (cudaStream_t)stream_));
    RB_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
    return 0;
  }
  DeviceDfa* fwd;
  if (int rc = ensure(kFwdUnanchoredAll, &fwd)) return rc;
  BatchArgs a{};
  a.fwd = fwd->view;
  const size_t smem = smem_for(fwd->view);
  a.use_smem = smem != 0;
  a.text = d_text;
  a.offsets = d_offsets;
  a.n_rec = n_rec;
  a.out_bits = d_bits;
  if (fwd->hot.n && ((uintptr_t)d_text & 15) == 0 && !tuning.force_generic) {
    a.fwd_hot = fwd->hot;
    a.fwd_g = (const DfaView*)fwd->view_dev;
    const size_t fsm = hot_bytes(fwd->hot.n) + 256;
    const uint32_t per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / (fsm + 8192)));
    // tasks of 96..2048 records, at least four per warp of the grid; smaller batches keep one record per lane
    const uint32_t grid0 = grid_for(n_rec, 512, per_sm);
    const uint64_t per_task = n_rec / ((uint64_t)grid0 * 16 * 4);
    a.task_recs = (uint32_t)std::min<uint64_t>(2048, per_task / 32 * 32);
    if (tuning.batch_refill >= 2) a.task_recs = std::max<uint32_t>(a.task_recs, 32);  // tests: whatever the size
    if (tuning.batch_refill && a.task_recs >= (tuning.batch_refill >= 2 ? 32u : 96u)) {
      uint32_t* counters = (uint32_t*)counters_.ensure(128);
      if (!counters) return fail("out of device memory (batch scratch)");
      a.task_counter = (unsigned long long*)(counters + 16);
      RB_CUDA(cudaMemsetAsync(a.task_counter, 0, 8, (cudaStream_t)stream_));
      // (8 KB of static shared memory for the result words on top of the table: opt in whenever the sum may pass 48 KB)
      RB_CUDA(allow_smem(batch_refill<0>, fsm, 48 * 1024));
      batch_refill<0><<<grid_for(n_rec, 512, per_sm), 512, fsm, (cudaStream_t)stream_>>>(a);
      RB_LAUNCH_CHECK("batch_refill<0>");
    } else {
      RB_CUDA(allow_smem(batch_fast<0>, fsm));
      batch_fast<0><<<grid_for(n_rec, 512, per_sm), 512, fsm, (cudaStream_t)stream_>>>(a);
      RB_LAUNCH_CHECK("batch_fast<0>");
    }
  } else {
    RB_CUDA(allow_smem(is_match_batch, smem));
    is_match_batch<<<(uint32_t)((n_rec + 255) / 256), 256, smem, (cudaStream_t)stream_>>>(a);
    RB_LAUNCH_CHECK("is_match_batch");
  }
  RB_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  return 0;
}

int Regex::find_batch_device(const uint8_t* d_text, const uint64_t* d_offsets, uint64_t n_rec, uint64_t* d_spans, uint32_t* d_bits) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (n_rec == 0) return 0;
  if (is_set_) return fail("find requires exactly one pattern");
  DeviceDfa *fwd, *rev;
  if (int rc = ensure(kFwdUnanchoredLF, &fwd)) return rc;
  if (int rc = ensure(kRevAnchoredLongest, &rev)) return rc;
  BatchArgs a{};
  a.fwd = fwd->view;
  a.rev = rev->view;
  a.text = d_text;
  a.offsets = d_offsets;
  a.n_rec = n_rec;
  a.out_bits = d_bits;
  a.out_spans = d_spans;
  if (fwd->hot.n && rev->hot.n && ((uintptr_t)d_text & 15) == 0 && !tuning.force_generic) {
    a.fwd_hot = fwd->hot;
    a.rev_hot = rev->hot;
    a.fwd_g = (const DfaView*)fwd->view_dev;
    a.rev_g = (const DfaView*)rev->view_dev;
    const size_t fsm = hot_bytes(fwd->hot.n) + hot_bytes(rev->hot.n) + 256;
    const uint32_t per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / (fsm + 16384)));
    const uint64_t per_task = n_rec / ((uint64_t)grid_for(n_rec, 512, per_sm) * 16 * 4);
    a.task_recs = (uint32_t)std::max<uint64_t>(32, std::min<uint64_t>(2048, per_task / 32 * 32));
    if (tuning.batch_refill > 1) {  // find: measured slower than batch_fast<1> on log lines, off unless asked for
      uint32_t* counters = (uint32_t*)counters_.ensure(128);
      if (!counters) return fail("out of device memory (batch scratch)");
      a.task_counter = (unsigned long long*)(counters + 16);
      RB_CUDA(cudaMemsetAsync(a.task_counter, 0, 8, (cudaStream_t)stream_));
      RB_CUDA(allow_smem(batch_refill<1>, fsm, 48 * 1024));
      batch_refill<1><<<grid_for(n_rec, 512, per_sm), 512, fsm, (cudaStream_t)stream_>>>(a);
      RB_LAUNCH_CHECK("batch_refill<1>");
    } else {
      RB_CUDA(allow_smem(batch_fast<1>, fsm));
      batch_fast<1><<<grid_for(n_rec, 512, per_sm), 512, fsm, (cudaStream_t)stream_>>>(a);
      RB_LAUNCH_CHECK("batch_fast<1>");
    }
  } else {
    find_batch<<<(uint32_t)((n_rec + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(a);
    RB_LAUNCH_CHECK("find_batch");
  }
  RB_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  return 0;
}

int Regex::set_matches_batch_device(const uint8_t* d_text, const uint64_t* d_offsets, uint64_t n_rec, uint64_t* d_masks) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (n_rec == 0) return 0;
  const uint32_t mw = (uint32_t)std::max<size_t>(1, (patterns_.size() + 63) / 64);
  if (patterns_.empty()) {
    if (int rc = init_device()) return rc;
    RB_CUDA(cudaMemsetAsync(d_masks, 0, n_rec * 8 * mw, (cudaStream_t)stream_));
    RB_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
    return 0;
  }
  DeviceDfa* fwd;
  if (int rc = ensure(kFwdUnanchoredAll, &fwd)) return rc;
  BatchArgs a{};
  a.fwd = fwd->view;
  const size_t smem = smem_for(fwd->view);
  a.use_smem = smem != 0;
  a.text = d_text;
  a.offsets = d_offsets;
  a.n_rec = n_rec;
  a.out_masks = d_masks;
  RB_CUDA(allow_smem(set_matches_batch, smem));
  set_matches_batch<<<(uint32_t)((n_rec + 255) / 256), 256, smem, (cudaStream_t)stream_>>>(a);
  RB_LAUNCH_CHECK("set_matches_batch");
  RB_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  return 0;
}

// ------------------------------------------------------- replace_all / split ----
// All spans of the haystack in spans_all_ (device); *m = their number.
int Regex::all_spans_device(const uint8_t* d_text, uint64_t n, uint64_t** d_spans, uint64_t* m) {
  uint64_t total = 0;
  if (int rc = find_all_device(d_text, n, 0, nullptr, 0, &total)) return rc;
  uint64_t* sp = (uint64_t*)spans_all_.ensure(std::max<uint64_t>(total, 1) * 16);
  if (!sp) return fail("out of device memory (spans)");
  if (total) {
    uint64_t again = 0;
    if (int rc = find_all_device(d_text, n, 0, sp, total, &again)) return rc;
  }
  *d_spans = sp;
  *m = total;
  return 0;
}

int Regex::replace_device(const uint8_t* d_text, uint64_t n, const uint8_t* rep, uint64_t rep_len, bool expand, uint64_t limit,
                          uint8_t* d_out, uint64_t out_cap, uint64_t* out_len) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (int rc = replace_prepare(d_text, n, rep, rep_len, expand, limit, out_len)) return rc;
  if (!d_out) return 0;
  return replace_emit(d_out, out_cap);
}

// Template, spans and prefix sums of one replace call; leaves everything replace_emit needs in
// rep_args_ / rep_lits_ and the device scratch.
int Regex::replace_prepare(const uint8_t* d_text, uint64_t n, const uint8_t* rep, uint64_t rep_len, bool expand, uint64_t limit,
                           uint64_t* out_len) {
  if (is_set_) return fail("replace requires exactly one pattern");
  // ---- the replacement template (src/expand.rs:50-90) ----
  rep_args_.assign(sizeof(ReplaceArgs), 0);
  ReplaceArgs& a = *reinterpret_cast<ReplaceArgs*>(rep_args_.data());
  a = ReplaceArgs{};
  std::vector<uint8_t>& lits = rep_lits_;
  lits.clear();
  auto literal = [&](const uint8_t* p, uint64_t len) -> bool {
    if (len == 0) return true;
    if (a.n_parts && !(a.part_len[a.n_parts - 1] & kRepGroup) && a.part_off[a.n_parts - 1] + a.part_len[a.n_parts - 1] == lits.size()) {
      a.part_len[a.n_parts - 1] += (uint32_t)len;  // extend the previous literal part
    } else {
      if (a.n_parts == kMaxRepParts) return false;
      a.part_off[a.n_parts] = (uint32_t)lits.size();
      a.part_len[a.n_parts++] = (uint32_t)len;
    }
    lits.insert(lits.end(), p, p + len);
    a.lit_total += len;
    return true;
  };
  bool fits = true, needs_groups = false;
  if (!expand) {
    fits = literal(rep, rep_len);
  } else {
    uint64_t i = 0;
    auto cap_letter = [](uint8_t b) { return (b >= '0' && b <= '9') || (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || b == '_'; };
    while (i < rep_len && fits) {
      uint64_t j = i;
      while (j < rep_len && rep[j] != '$') j++;
      fits = literal(rep + i, j - i);
      i = j;
      if (i >= rep_len) break;
      if (i + 1 < rep_len && rep[i + 1] == '$') { fits = fits && literal(rep + i, 1); i += 2; continue; }
      // find_cap_ref (expand.rs:128-167)
      uint64_t k = i + 1;
      bool brace = false;
      if (k < rep_len && rep[k] == '{') { brace = true; k++; }
      uint64_t ce = k;
      while (ce < rep_len && cap_letter(rep[ce])) ce++;
      bool ok = rep_len - i > 1 && ce > k;
      std::string name;
      if (ok) {
        name.assign((const char*)rep + k, ce - k);
        if (brace) { if (ce < rep_len && rep[ce] == '}') ce++; else ok = false; }
      }
      if (!ok) { fits = fits && literal(rep + i, 1); i += 1; continue; }  // a lone '$'
      i = ce;
      // expand.rs:78-87: a number is a group index, anything else a group name; a group the pattern
      // does not have expands to nothing
      long g = -1;
      if (name.find_first_not_of("0123456789") == std::string::npos) {
        g = name.size() <= 9 ? std::atol(name.c_str()) : -1;
      } else {
        for (const auto& gn : group_name_index_) if (gn.first == name) g = gn.second;
      }
      if (g >= 0 && g < n_groups_) {
        if (a.n_parts == kMaxRepParts) { fits = false; break; }
        a.part_len[a.n_parts] = kRepGroup | (uint32_t)g;
        a.part_off[a.n_parts++] = 0;
        if (g > 0) needs_groups = true;
      }
    }
  }
  if (!fits) return fail("replacement template has too many parts");
  uint64_t* spans;
  uint64_t m;
  if (int rc = all_spans_device(d_text, n, &spans, &m)) return rc;
  if (limit && m > limit) m = limit;
  cudaStream_t st = (cudaStream_t)stream_;
  uint64_t* lens = (uint64_t*)lens_.ensure(std::max<uint64_t>(m, 1) * 8);
  uint64_t* before = (uint64_t*)offset_.ensure(std::max<uint64_t>(m, 1) * 8);
  const uint64_t n_blocks = (m + 1023) / 1024;
  uint64_t* block_sums = (uint64_t*)block_sums_.ensure(std::max<uint64_t>(n_blocks, 1) * 8);
  uint32_t* counters = (uint32_t*)counters_.ensure(128);
  uint8_t* d_lits = (uint8_t*)lits_.ensure(std::max<size_t>(lits.size(), 16));
  if (!lens || !before || !block_sums || !counters || !d_lits) return fail("out of device memory (replace scratch)");
  uint64_t* reps_before = (uint64_t*)reps_before_.ensure(std::max<uint64_t>(m, 1) * 8);
  if (!reps_before) return fail("out of device memory (replace scratch)");
  a.text = d_text;
  a.n = n;
  a.spans = spans;
  a.n_matches = m;
  a.lens_before = before;
  a.reps_before = reps_before;
  a.lits = d_lits;
  if (needs_groups && m) {  // `$1`, `$name`: the groups of every match (capture pass on the narrowed windows)
    a.n_slots = 2 * (uint32_t)n_groups_;
    uint64_t* d_slots = (uint64_t*)cap_slots_.ensure(m * a.n_slots * 8);
    if (!d_slots) return fail("out of device memory (captures)");
    if (int rc = captures_device(d_text, n, spans, m, d_slots)) return rc;
    a.slots = d_slots;
  }
  uint64_t matched = 0, replaced = 0;
  if (m) {
    unsigned long long* grand = (unsigned long long*)(counters + 4);
    auto prefix = [&](uint64_t* out, uint64_t* total) -> int {  // exclusive prefix sum of lens[] into out[]
      scan_counts_local<<<(uint32_t)n_blocks, 1024, 0, st>>>(lens, out, block_sums, m);
      RB_LAUNCH_CHECK("scan_counts_local");
      scan_block_sums<<<1, 1024, 0, st>>>(block_sums, n_blocks, grand);
      RB_LAUNCH_CHECK("scan_block_sums");
      scan_add_block_offsets<<<(uint32_t)n_blocks, 1024, 0, st>>>(out, block_sums, m);
      RB_LAUNCH_CHECK("scan_add_block_offsets");
      RB_CUDA(d2h(total, grand, 8));
      return 0;
    };
    span_lengths<<<grid_for(m, 256, 8), 256, 0, st>>>(spans, m, lens);
    RB_LAUNCH_CHECK("span_lengths");
    if (int rc = prefix(before, &matched)) return rc;
    replace_lengths<<<grid_for(m, 256, 8), 256, 0, st>>>(a, lens);
    RB_LAUNCH_CHECK("replace_lengths");
    if (int rc = prefix(reps_before, &replaced)) return rc;
  }
  rep_totals_[0] = matched;
  rep_totals_[1] = replaced;
  *out_len = n - matched + replaced;
  return 0;
}

int Regex::replace_emit(uint8_t* d_out, uint64_t out_cap) {
  ReplaceArgs a = *reinterpret_cast<ReplaceArgs*>(rep_args_.data());
  const std::vector<uint8_t>& lits = rep_lits_;
  cudaStream_t st = (cudaStream_t)stream_;
  uint8_t* d_lits = (uint8_t*)lits_.ptr;
  const uint64_t n = a.n, m = a.n_matches;
  if (!lits.empty()) RB_CUDA(cudaMemcpyAsync(d_lits, lits.data(), lits.size(), cudaMemcpyHostToDevice, st));
  a.out = d_out;
  a.out_cap = out_cap;
  if (n) {
    replace_gaps<<<grid_for(((n + 2047) / 2048) * 32, 256, 8), 256, 0, st>>>(a, rep_totals_[0], rep_totals_[1]);
    RB_LAUNCH_CHECK("replace_gaps");
  }
  if (m && a.n_parts) {
    replace_matches<<<grid_for(m * 32, 256, 8), 256, 0, st>>>(a);
    RB_LAUNCH_CHECK("replace_matches");
  }
  RB_CUDA(cudaStreamSynchronize(st));  // `lits` and the caller's buffers may go away now
  return 0;
}

int Regex::replace_host(const uint8_t* text, uint64_t n, const uint8_t* rep, uint64_t rep_len, bool expand, uint64_t limit, uint8_t* out,
                        uint64_t out_cap, uint64_t* out_len) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  int rc;
  const uint8_t* d = upload_text(text, n, &rc);
  if (rc) return rc;
  if ((rc = replace_prepare(d, n, rep, rep_len, expand, limit, out_len))) return rc;
  if (!out) return 0;
  uint8_t* d_out = (uint8_t*)rep_out_.ensure(std::max<uint64_t>(*out_len, 16));
  if (!d_out) return fail("out of device memory (replace output)");
  if ((rc = replace_emit(d_out, *out_len))) return rc;
  const uint64_t k = std::min(out_cap, *out_len);
  if (k) RB_CUDA(d2h(out, d_out, k));
  return 0;
}

int Regex::split_device(const uint8_t* d_text, uint64_t n, bool has_limit, uint64_t limit, uint64_t* d_pieces, uint64_t cap, uint64_t* n_pieces) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *n_pieces = 0;
  if (is_set_) return fail("split requires exactly one pattern");
  if (has_limit && limit == 0) return 0;  // re_bytes.rs:738-740
  uint64_t* spans;
  uint64_t m;
  if (int rc = all_spans_device(d_text, n, &spans, &m)) return rc;
  uint64_t last_end = 0;
  if (m) RB_CUDA(d2h(&last_end, spans + 2 * m - 1, 8));
  // Split yields one piece per match and the rest of the text when it is not empty (re_bytes.rs:702-720)
  const uint64_t p_split = m + (last_end < n ? 1 : 0);
  uint64_t pieces = p_split;
  int last_is_rest = 0;
  if (has_limit && p_split >= limit - 1) {  // SplitN: the limit-th piece is whatever remains, empty or not (re_bytes.rs:737-748)
    pieces = limit;
    last_is_rest = 1;
  }
  *n_pieces = pieces;
  if (d_pieces && pieces && cap) {
    split_pieces<<<grid_for(pieces, 256, 8), 256, 0, (cudaStream_t)stream_>>>(spans, m, n, pieces, last_is_rest, d_pieces, cap);
    RB_LAUNCH_CHECK("split_pieces");
    RB_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  }
  return 0;
}

int Regex::split_host(const uint8_t* text, uint64_t n, bool has_limit, uint64_t limit, uint64_t* pieces, uint64_t cap, uint64_t* n_pieces) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  int rc;
  const uint8_t* d = upload_text(text, n, &rc);
  if (rc) return rc;
  if ((rc = split_device(d, n, has_limit, limit, nullptr, 0, n_pieces))) return rc;
  if (!pieces || !cap || !*n_pieces) return 0;
  const uint64_t k = std::min(cap, *n_pieces);
  uint64_t* d_p = (uint64_t*)pieces_.ensure(k * 16);
  if (!d_p) return fail("out of device memory (pieces)");
  if ((rc = split_device(d, n, has_limit, limit, d_p, k, n_pieces))) return rc;
  RB_CUDA(d2h(pieces, d_p, k * 16));
  return 0;
}

// ------------------------------------------------------------ capture groups ----
int Regex::ensure_capture_program() {
  if (cap_n_insts_) return 0;
  if (is_set_ || patterns_.size() != 1) return fail("captures require exactly one pattern (exec.rs:587-589)");
  if (int rc = init_device()) return rc;
  rb::CompileOptions co;
  co.only_utf8 = opt_.only_utf8;
  co.size_limit = opt_.size_limit;
  co.unanchored_prefix = true;
  co.saves = true;
  rb::Program prog;
  rb::Error err;
  if (!rb::compile(exprs_, co, &prog, &err)) return fail(err.msg);
  std::vector<NfaInst> insts(prog.insts.size());
  for (size_t i = 0; i < insts.size(); i++) {
    const rb::Inst& in = prog.insts[i];
    insts[i].op_look_lo_hi = (uint32_t)in.op | ((uint32_t)in.look << 8) | ((uint32_t)in.lo << 16) | ((uint32_t)in.hi << 24);
    insts[i].a = in.a;
    insts[i].b = in.b;
  }
  NfaInst* d = (NfaInst*)cap_insts_.ensure(insts.size() * sizeof(NfaInst));
  if (!d) return fail("out of device memory (capture program)");
  RB_CUDA(cudaMemcpyAsync(d, insts.data(), insts.size() * sizeof(NfaInst), cudaMemcpyHostToDevice, (cudaStream_t)stream_));
  RB_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  cap_n_insts_ = (uint32_t)insts.size();
  cap_start_ = prog.start;
  cap_anchored_ = prog.is_anchored_start;
  return 0;
}

int Regex::captures_device(const uint8_t* d_text, uint64_t n, const uint64_t* d_spans, uint64_t m, uint64_t* d_slots) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  if (int rc = ensure_capture_program()) return rc;
  if (m == 0) return 0;
  CapArgs a{};
  a.insts = (const NfaInst*)cap_insts_.ptr;
  a.n_insts = cap_n_insts_;
  a.start_ip = cap_start_;
  a.anchored_start = cap_anchored_ ? 1 : 0;
  a.n_slots = 2 * (uint32_t)n_groups_;
  a.text = d_text;
  a.n = n;
  a.spans = d_spans;
  a.n_matches = m;
  a.slots = d_slots;
  // kernels.cu pike_captures: two slot tables, thread caps, stack positions, four index arrays, stack tags
  const uint64_t ni = a.n_insts, ns = a.n_slots;
  a.per_thread = (2 * ni * ns * 8 + ns * 8 + (2 * ni + 4) * 8 + 4 * ni * 4 + (2 * ni + 4) * 4 + 15) / 16 * 16;
  if (a.per_thread > (8u << 20)) return fail("the pattern is too large for capture extraction on the device (program x groups)");
  uint64_t threads = std::min<uint64_t>((m + 63) / 64 * 64, 16384);
  while (threads > 64 && threads * a.per_thread > (1ull << 30)) threads /= 2;
  a.scratch = (uint8_t*)cap_scratch_.ensure(threads * a.per_thread);
  if (!a.scratch) return fail("out of device memory (capture scratch)");
  pike_captures<<<(uint32_t)(threads / 64), 64, 0, (cudaStream_t)stream_>>>(a);
  RB_LAUNCH_CHECK("pike_captures");
  RB_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  return 0;
}

int Regex::captures_at_host(const uint8_t* text, uint64_t n, uint64_t start, bool* found, uint64_t* slots) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  uint64_t s = 0, e = 0;
  if (int rc = find_at_host(text, n, start, found, &s, &e)) return rc;  // the DFA path finds the match (exec.rs:547-556)
  if (!*found) return 0;
  // the window the capture pass reads: one byte of look-behind, two characters of look-ahead
  uint8_t* d = (uint8_t*)text_.ensure(n + 64);
  const uint64_t lo = s ? s - 1 : 0, hi = std::min<uint64_t>(n, e + 8);
  cudaStream_t st = (cudaStream_t)stream_;
  RB_CUDA(cudaMemcpyAsync(d + lo, text + lo, hi - lo, cudaMemcpyHostToDevice, st));
  uint64_t* d_span = (uint64_t*)cap_span_.ensure(16);
  const uint32_t ns = 2 * (uint32_t)n_groups_;
  uint64_t* d_slots = (uint64_t*)cap_slots_.ensure(ns * 8);
  if (!d_span || !d_slots) return fail("out of device memory (captures)");
  const uint64_t span[2] = {s, e};
  RB_CUDA(cudaMemcpyAsync(d_span, span, 16, cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaStreamSynchronize(st));
  if (int rc = captures_device(d, hi, d_span, 1, d_slots)) return rc;  // (hi bounds a window the pass never leaves)
  RB_CUDA(d2h(slots, d_slots, ns * 8));
  *found = slots[0] != kNone;  // exec.rs:896-906: the result is what the NFA found in the window
  return 0;
}

int Regex::captures_all_host(const uint8_t* text, uint64_t n, uint64_t* slots, uint64_t cap, uint64_t* m) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  int rc;
  const uint8_t* d = upload_text(text, n, &rc);
  if (rc) return rc;
  uint64_t* spans;
  if ((rc = all_spans_device(d, n, &spans, m))) return rc;
  const uint64_t k = std::min(cap, *m);
  if (!slots || !k) return 0;
  const uint32_t ns = 2 * (uint32_t)n_groups_;
  uint64_t* d_slots = (uint64_t*)cap_slots_.ensure(k * ns * 8);
  if (!d_slots) return fail("out of device memory (captures)");
  if ((rc = captures_device(d, n, spans, k, d_slots))) return rc;
  RB_CUDA(d2h(slots, d_slots, k * ns * 8));
  // CaptureMatches (re_trait.rs:236-263) iterates on what read_captures_at returns, i.e. on the NFA's
  // group 0.  That is the find_iter span except where the reverse-on-slice quirk gave
  // the DFA a start no match begins at; from the first such match on, iterate like the reference.
  std::vector<uint64_t> h_spans(2 * k);
  RB_CUDA(d2h(h_spans.data(), spans, k * 16));
  uint64_t j = 0;
  while (j < k && slots[j * ns] == h_spans[2 * j] && slots[j * ns + 1] == h_spans[2 * j + 1]) j++;
  if (j == k && *m <= cap) return 0;
  if (j == k) return 0;  // the caller's buffer ends before any divergence
  auto next_after_empty = [&](uint64_t i) -> uint64_t {  // exec.rs:335-337, 375-377
    if (!only_utf8 || i >= n) return i + 1;
    const uint8_t b = text[i];
    return i + (b <= 0x7F ? 1 : b <= 0xDF ? 2 : b <= 0xEF ? 3 : 4);
  };
  uint64_t last_end = 0, last_match = kNone;
  if (j) {
    const uint64_t ps = h_spans[2 * (j - 1)], pe = h_spans[2 * (j - 1) + 1];
    last_end = ps == pe ? next_after_empty(pe) : pe;
    last_match = pe;
  }
  std::vector<uint64_t> one(ns);
  uint64_t count = j;
  while (last_end <= n) {
    bool found = false;
    if ((rc = captures_at_host(text, n, last_end, &found, one.data()))) return rc;
    if (!found) break;
    const uint64_t s = one[0], e = one[1];
    if (s == e) {
      last_end = next_after_empty(e);
      if (e == last_match) continue;
    } else {
      last_end = e;
    }
    last_match = e;
    if (count < cap) std::copy(one.begin(), one.end(), slots + count * ns);
    count++;
  }
  *m = count;
  return 0;
}

// ------------------------------------------------------------ host wrappers ----
// device -> host on the library's stream, complete on return
int Regex::d2h(void* dst, const void* src, size_t bytes) {
  int e = (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream_);
  if (e != cudaSuccess) return e;
  return (int)cudaStreamSynchronize((cudaStream_t)stream_);
}

const uint8_t* Regex::upload_text(const uint8_t* text, uint64_t n, int* rc) {
  if ((*rc = init_device())) return nullptr;
  uint8_t* d = (uint8_t*)text_.ensure(n + 64);
  if (!d) { *rc = fail("out of device memory (haystack)"); return nullptr; }
  if (n) {  // stream-ordered with the kernels that read it
    int e = (int)cudaMemcpyAsync(d, text, n, cudaMemcpyHostToDevice, (cudaStream_t)stream_);
    if (e != cudaSuccess) { *rc = check(e, "haystack upload"); return nullptr; }
  }
  return d;
}

// Bytes per uploaded piece of a pipelined host find_all (RB200_PIPELINE_PIECE overrides; 0 = off).
// Measured on 4 GiB of pinned host text: 43.8 GB/s upload-then-search, 51.7 GB/s pipelined
// (the PCIe rate).  tests/test_gpu_parity.py::test_pipelined_host_find_all and
// tools/micro/pipeline_probe.py compare it with the plain path.
static constexpr uint64_t kPipelinePieceDefault = 64ull << 20;

// Host haystack, pipelined: the upload is cut into pieces on a copy stream and every piece is
// searched as a byte-range shard of the (partly resident) device buffer as soon as it and
// the halo behind it have arrived, so the search hides under the PCIe transfer and the
// spans of a piece travel back while later pieces are still going up.  Pieces are taken
// left to right, so each enters the find_iter chain with its predecessor's exact exit
// state; the reverse-scan state a piece assumed at its top is checked against the exact
// one its successor computes.  Anything unusual (a wrong assumption, a match running past
// the halo) abandons the pipeline: the caller then runs the plain path on the resident text.
// Returns 0 = done, 1 = not applicable / abandoned (text is fully resident), < 0 = error.
int Regex::find_all_host_pipelined(const uint8_t* text, uint64_t n, uint8_t* d, uint64_t* d_out, uint64_t* out, uint64_t cap,
                                   uint64_t* total) {
  uint64_t piece = kPipelinePieceDefault;
  if (const char* e = getenv("RB200_PIPELINE_PIECE")) piece = (uint64_t)atoll(e) / 256 * 256;
  const uint64_t halo = 64 << 10;
  if (piece == 0 || n < 2 * piece || is_set_ || patterns_.empty()) {
    RB_CUDA(cudaMemcpyAsync(d, text, n, cudaMemcpyHostToDevice, (cudaStream_t)stream_));
    return 1;
  }
  if (!copy_stream_) {
    cudaStream_t s1, s2;
    RB_CUDA(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    RB_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    copy_stream_ = s1;
    back_stream_ = s2;
  }
  cudaStream_t up = (cudaStream_t)copy_stream_, back = (cudaStream_t)back_stream_;
  const uint64_t n_pieces = (n + piece - 1) / piece;
  std::vector<cudaEvent_t> arrived(n_pieces, nullptr);
  auto cleanup = [&](int rc) {  // every exit goes through here
    cudaStreamSynchronize(up);  // the plain path needs the whole text anyway
    cudaStreamSynchronize(back);
    for (auto& e : arrived) if (e) cudaEventDestroy(e);
    return rc;
  };
#define RB_PIPE(call)                                    \
  do {                                                   \
    int rc__ = check((int)(call), #call);                \
    if (rc__) return cleanup(rc__);                      \
  } while (0)
  for (uint64_t i = 0; i < n_pieces; i++) {
    const uint64_t lo = i * piece, len = std::min(piece, n - lo);
    RB_PIPE(cudaMemcpyAsync(d + lo, text + lo, len, cudaMemcpyHostToDevice, up));
    RB_PIPE(cudaEventCreateWithFlags(&arrived[i], cudaEventDisableTiming));
    RB_PIPE(cudaEventRecord(arrived[i], up));
  }
  // scratch sized for the whole haystack up front (DeviceBuf::ensure reallocates on growth)
  if (!bitmap_.ensure(((n >> 6) + 2) * 8)) return cleanup(fail("out of device memory (bitmap)"));
  uint64_t done = 0, chain_p = 0, chain_lm = kNone;
  uint32_t prev_guess = kNoState;
  for (uint64_t i = 0; i < n_pieces; i++) {
    const uint64_t lo = i * piece, hi = std::min(n, lo + piece);
    const uint64_t n_i = std::min(n, hi + halo);  // resident prefix this piece may read
    RB_PIPE(cudaEventSynchronize(arrived[std::min(n_pieces - 1, (n_i - 1) / piece)]));
    ShardIO io;
    io.own_lo = lo;
    io.own_hi = hi;
    io.is_first = i == 0;
    io.is_last = n_i == n;
    io.chain_p = chain_p;
    io.chain_lm = chain_lm;
    uint64_t* dst = d_out && cap > done ? d_out + 2 * done : nullptr;
    const uint64_t room = cap > done ? cap - done : 0;
    const int rc = find_all_shard_device(d, n_i, &io, dst, room);
    if (rc) {
      if (io.halo_overflow) return cleanup(1);  // a match longer than the halo: do it in one piece
      return cleanup(rc);
    }
    if (i > 0 && prev_guess != io.rev_left) return cleanup(1);  // the previous piece guessed its top state wrong
    prev_guess = hi < n ? io.rev_guess : kNoState;
    if (dst) {
      const uint64_t k = std::min(room, io.n_matches);
      if (k) RB_PIPE(cudaMemcpyAsync(out + 2 * done, dst, k * 16, cudaMemcpyDeviceToHost, back));
    }
    done += io.n_matches;
    chain_p = io.exit_p;
    chain_lm = io.exit_lm;
    if (chain_p == kNone) {  // the iteration ended (a failed slice emulation): nothing further matches
      break;
    }
  }
  *total = done;
  return cleanup(0);
#undef RB_PIPE
}

int Regex::find_all_host(const uint8_t* text, uint64_t n, uint64_t start, uint64_t* out, uint64_t cap, uint64_t* total) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  int rc;
  if (start == 0 && n >= (8u << 20) && out && cap) {
    if ((rc = init_device())) return rc;
    uint8_t* dd = (uint8_t*)text_.ensure(n + 64);
    uint64_t* d_out = (uint64_t*)out_.ensure(cap * 16);
    if (!dd || !d_out) return fail("out of device memory (haystack / spans)");
    DeviceDfa* warm;  // picks the stream, creates the pinned scratch
    if ((rc = ensure(kFwdAnchoredLF, &warm))) return rc;
    *total = 0;
    rc = find_all_host_pipelined(text, n, dd, d_out, out, cap, total);
    if (rc <= 0) return rc < 0 ? rc : 0;
    // not applicable or abandoned: the text is resident now, search it in one piece
    if ((rc = find_all_device(dd, n, 0, d_out, cap, total))) return rc;
    const uint64_t k = std::min(cap, *total);
    if (k) RB_CUDA(d2h(out, d_out, k * 16));
    return 0;
  }
  const uint8_t* d = upload_text(text, n, &rc);
  if (rc) return rc;
  uint64_t* d_out = nullptr;
  if (out && cap) {
    d_out = (uint64_t*)out_.ensure(cap * 16);
    if (!d_out) return fail("out of device memory (spans)");
  }
  if ((rc = find_all_device(d, n, start, d_out, cap, total))) return rc;
  const uint64_t k = std::min(cap, *total);
  if (d_out && k) RB_CUDA(d2h(out, d_out, k * 16));
  return 0;
}
// find_at on a host haystack (rure_find): the reference's cost is the distance to the next match
// (src/exec.rs:473-514), so the haystack is uploaded and searched in windows that grow from
// 64 KiB (x8 each) starting at `start`, every window as a byte-range shard entered with the
// exact iterator state of the one before, and the search stops at the first span.  A C loop
// `while (rure_find(re, h, n, pos, &m)) pos = m.end;` therefore moves each byte over PCIe
// about once per match near it instead of once per call.
int Regex::find_at_host(const uint8_t* text, uint64_t n, uint64_t start, bool* found, uint64_t* s, uint64_t* e) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *found = false;
  if (start > n) return 0;
  if (int rc = init_device()) return rc;
  if (is_set_) return fail("find requires exactly one pattern (RegexSet cannot be used with find, exec.rs:510-512)");
  uint8_t* d = (uint8_t*)text_.ensure(n + 64);
  uint64_t* d_out = (uint64_t*)out_.ensure(16);
  if (!d || !d_out) return fail("out of device memory (haystack)");
  cudaStream_t st = (cudaStream_t)stream_;
  const uint64_t buf_lo = start > 256 ? (start - 256) & ~255ull : 0;  // 256 bytes of left context for look-behind
  uint64_t resident = buf_lo;                                         // bytes [buf_lo, resident) are on the device
  uint64_t own_lo = start ? ((start - 1) & ~255ull) : 0;              // bit i <-> position i + 1
  uint64_t win = 64 << 10, halo = 64 << 10;
  uint64_t p = start, lm = kNone;
  for (;;) {
    uint64_t own_hi = std::min<uint64_t>(n, (own_lo + win + 255) & ~(uint64_t)255);
    if (n - own_hi < 4096) own_hi = n;
    const uint64_t hi = std::min(n, own_hi + halo);
    if (hi == n) own_hi = n;  // the buffer ends with the text: the last window owns everything up to it
    if (hi > resident) {
      RB_CUDA(cudaMemcpyAsync(d + resident, text + resident, hi - resident, cudaMemcpyHostToDevice, st));
      resident = hi;
    }
    ShardIO io;
    io.own_lo = own_lo - buf_lo;
    io.own_hi = own_hi - buf_lo;
    io.is_first = buf_lo == 0;
    io.is_last = hi == n;
    io.chain_p = p - buf_lo;
    io.chain_lm = lm == kNone ? kNone : lm - buf_lo;
    const int rc = find_all_shard_device(d + buf_lo, hi - buf_lo, &io, d_out, 1);
    if (rc) {
      if (!io.halo_overflow) return rc;
      if (hi == n) return rc;
      halo *= 8;  // a match longer than the halo: look further
      continue;
    }
    if (io.n_matches) {
      uint64_t h[2];
      RB_CUDA(d2h(h, d_out, 16));
      *found = true;
      *s = h[0] + buf_lo;
      *e = h[1] + buf_lo;
      return 0;
    }
    if (own_hi == n || io.exit_p == kNone) return 0;
    p = io.exit_p + buf_lo;
    lm = io.exit_lm == kNone ? kNone : io.exit_lm + buf_lo;
    own_lo = own_hi;
    win *= 8;
  }
}
// is_match / shortest_match / set matches on a host haystack: the waves of forward_reduce upload
// what they are about to read, so an early exit also ends the transfer (dfa.rs:658-667).
int Regex::shortest_match_host(const uint8_t* text, uint64_t n, uint64_t start, bool* found, uint64_t* end) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *found = false;
  if (int rc = init_device()) return rc;
  uint8_t* d = (uint8_t*)text_.ensure(n + 64);
  if (!d) return fail("out of device memory (haystack)");
  LazyUpload up{text, d, n, start > 256 ? (start - 256) & ~255ull : 0};
  lazy_ = &up;
  const int rc = shortest_match_device(d, n, start, found, end);
  lazy_ = nullptr;
  return rc;
}
int Regex::set_matches_host(const uint8_t* text, uint64_t n, uint64_t start, bool* any, uint64_t* masks) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  *any = false;
  if (int rc = init_device()) return rc;
  uint8_t* d = (uint8_t*)text_.ensure(n + 64);
  if (!d) return fail("out of device memory (haystack)");
  LazyUpload up{text, d, n, start > 256 ? (start - 256) & ~255ull : 0};
  lazy_ = &up;
  const int rc = set_matches_device(d, n, start, any, masks);
  lazy_ = nullptr;
  return rc;
}
int Regex::is_match_batch_host(const uint8_t* text, const uint64_t* offsets, uint64_t n_rec, uint8_t* out_bits) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  int rc;
  const uint64_t n = n_rec ? offsets[n_rec] : 0;
  const uint8_t* d = upload_text(text, n, &rc);
  if (rc) return rc;
  uint64_t* d_off = (uint64_t*)offsets_.ensure((n_rec + 1) * 8);
  uint32_t* d_bits = (uint32_t*)bits_.ensure((n_rec + 31) / 32 * 4 + 4);
  if (!d_off || !d_bits) return fail("out of device memory (batch)");
  RB_CUDA(cudaMemcpyAsync(d_off, offsets, (n_rec + 1) * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream_));
  if ((rc = is_match_batch_device(d, d_off, n_rec, d_bits))) return rc;
  RB_CUDA(d2h(out_bits, d_bits, (n_rec + 7) / 8));
  return 0;
}
int Regex::find_batch_host(const uint8_t* text, const uint64_t* offsets, uint64_t n_rec, uint64_t* spans, uint8_t* out_bits) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  int rc;
  const uint64_t n = n_rec ? offsets[n_rec] : 0;
  const uint8_t* d = upload_text(text, n, &rc);
  if (rc) return rc;
  uint64_t* d_off = (uint64_t*)offsets_.ensure((n_rec + 1) * 8);
  uint32_t* d_bits = (uint32_t*)bits_.ensure((n_rec + 31) / 32 * 4 + 4);
  uint64_t* d_spans = (uint64_t*)out_.ensure(std::max<uint64_t>(n_rec, 1) * 16);
  if (!d_off || !d_bits || !d_spans) return fail("out of device memory (batch)");
  RB_CUDA(cudaMemcpyAsync(d_off, offsets, (n_rec + 1) * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream_));
  if ((rc = find_batch_device(d, d_off, n_rec, d_spans, d_bits))) return rc;
  RB_CUDA(d2h(out_bits, d_bits, (n_rec + 7) / 8));
  RB_CUDA(d2h(spans, d_spans, n_rec * 16));
  return 0;
}
int Regex::set_matches_batch_host(const uint8_t* text, const uint64_t* offsets, uint64_t n_rec, uint64_t* masks) {
  std::lock_guard<std::recursive_mutex> lock(mu_);
  int rc;
  const uint64_t n = n_rec ? offsets[n_rec] : 0;
  const uint8_t* d = upload_text(text, n, &rc);
  if (rc) return rc;
  const uint32_t mw = (uint32_t)std::max<size_t>(1, (patterns_.size() + 63) / 64);
  uint64_t* d_off = (uint64_t*)offsets_.ensure((n_rec + 1) * 8);
  uint64_t* d_masks = (uint64_t*)masks_.ensure(std::max<uint64_t>(n_rec, 1) * 8 * mw);
  if (!d_off || !d_masks) return fail("out of device memory (batch)");
  RB_CUDA(cudaMemcpyAsync(d_off, offsets, (n_rec + 1) * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream_));
  if ((rc = set_matches_batch_device(d, d_off, n_rec, d_masks))) return rc;
  RB_CUDA(d2h(masks, d_masks, n_rec * 8 * mw));
  return 0;
}

}  // namespace rbgpu
