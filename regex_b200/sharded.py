"""Byte-range sharding of one haystack across GPUs.

One process per GPU.  Rank g keeps [left context | its bytes | right halo] of the
haystack in its own HBM and runs the ordinary single-GPU kernels on it through
rure_b200_find_all_shard_device.  The only data that crosses NVLink are three tiny
per-shard values, exchanged in ONE all_gather per round (NCCL on GPUs, gloo in the CPU
tests; a single round when every boundary guess was right):

  1. the reverse-scan state at the shard's left edge -- the right-hand neighbour of a
     shard boundary knows it exactly, the left-hand one guessed it from its halo;
     a wrong guess triggers a redo of the affected segments only;
  2. the find_iter iterator state (next search position, previous match end) leaving
     each shard -- a shard that speculated "the iterator enters at my first byte" re-walks
     (no re-scan) when a match of its left neighbour reaches into it;
  3. the match counts, whose exclusive prefix sum gives every rank its global offset.

The reference has nothing to mirror here (the crate is single-threaded per search,
exec.rs:1066-1072); results equal find_iter over the concatenated haystack.
"""
from dataclasses import dataclass

NO_STATE = 0xFFFFFFFF
NONE = (1 << 64) - 1
SPEC = NONE - 1
ALIGN = 256


@dataclass
class ShardGeometry:
    rank: int
    world: int
    total: int      # length of the whole haystack
    a: int          # this rank owns match starts at global positions (a, b]  (+ 0 for rank 0)
    b: int
    buf_lo: int     # global index of buffer byte 0
    buf_hi: int     # global index one past the last buffer byte

    @property
    def own_lo(self):
        return self.a - self.buf_lo

    @property
    def own_hi(self):
        return self.b - self.buf_lo

    @property
    def n_buf(self):
        return self.buf_hi - self.buf_lo

    @property
    def is_first(self):
        return self.rank == 0

    @property
    def is_last(self):
        """The buffer ends where the haystack does (end-of-text semantics apply at its end).
        True for the last rank, and for an earlier rank whose halo reaches the end of a short
        haystack; the protocol itself goes by rank order."""
        return self.buf_hi >= self.total


def plan(total, world, rank, halo=1 << 16, left_ctx=ALIGN):
    """Boundaries at multiples of 256 bytes; `halo` bytes of the right neighbour and
    `left_ctx` bytes of the left neighbour are replicated (look-behind / long matches)."""
    assert halo % 16 == 0 and left_ctx % ALIGN == 0 and left_ctx >= ALIGN

    def edge(g):
        if g <= 0:
            return 0
        if g >= world:
            return total
        return min(total, (total * g // world) // ALIGN * ALIGN)

    a, b = edge(rank), edge(rank + 1)
    if rank < world - 1 and b >= total:  # tiny haystacks: trailing ranks own nothing
        b = a if a >= total else b
    buf_lo = a - left_ctx if rank > 0 and a > 0 else a
    buf_hi = total if rank == world - 1 else min(total, b + halo)
    return ShardGeometry(rank, world, total, a, b, buf_lo, buf_hi)


class TorchDistComm:
    """all_gather of small integer vectors over torch.distributed (nccl or gloo)."""

    def __init__(self, device, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.device = device
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_gather(self, values):
        """One collective, one upload and one download per call: the vector goes out from a pinned host
        buffer, `all_gather_into_tensor` fills one device tensor for all ranks, and one copy brings it
        back (per-rank output tensors cost a device-to-host copy and a sync per rank)."""
        import torch
        n = len(values)
        bufs = self._bufs.get(n) if hasattr(self, "_bufs") else None
        if bufs is None:
            if not hasattr(self, "_bufs"):
                self._bufs = {}
            on_gpu = torch.device(self.device).type == "cuda"
            h_in = torch.empty(n, dtype=torch.int64, pin_memory=on_gpu)
            h_out = torch.empty(self.world * n, dtype=torch.int64, pin_memory=on_gpu)
            d_in = torch.empty(n, dtype=torch.int64, device=self.device)
            d_out = torch.empty(self.world * n, dtype=torch.int64, device=self.device)
            bufs = self._bufs[n] = (h_in, h_out, d_in, d_out, on_gpu)
        h_in, h_out, d_in, d_out, on_gpu = bufs
        # 64-bit patterns travel as int64 (two's complement keeps NONE / SPEC intact)
        a = h_in.numpy()
        for i, v in enumerate(values):
            a[i] = v - (1 << 64) if v >= (1 << 63) else v
        d_in.copy_(h_in, non_blocking=True)
        self.dist.all_gather_into_tensor(d_out, d_in, group=self.group)
        h_out.copy_(d_out, non_blocking=True)
        if on_gpu:
            torch.cuda.current_stream(d_out.device).synchronize()
        flat = h_out.numpy()
        return [[int(x) & ((1 << 64) - 1) for x in flat[g * n:(g + 1) * n]] for g in range(self.world)]


class ThreadComm:
    """In-process stand-in for the collectives: `world` threads, each owning one shard
    (tests drive several shards through one GPU this way)."""

    def __init__(self, world):
        import threading
        self.world = world
        self._barrier = threading.Barrier(world)
        self._slots = [None] * world

    def view(self, rank):
        parent = self

        class _View:
            world = parent.world

            def __init__(self):
                self.rank = rank

            def all_gather(self, values):
                parent._slots[rank] = list(values)
                parent._barrier.wait()
                out = [list(v) for v in parent._slots]
                parent._barrier.wait()
                return out

        return _View()


class SingleComm:
    rank, world = 0, 1

    def all_gather(self, values):
        return [list(values)]


def _spec_ok(tp, tl, c_first, can_match_empty, has_looks):
    """Shard-granularity mirror of stitch_resolve() in csrc/gpu/kernels.cu: may a shard that
    assumed "the iterator enters at my first position" keep its result when the iterator really
    enters at (tp, tl)?  Look-around patterns never may: the START of the shard's first match
    comes from the reverse DFA over text[tp..] (src/exec.rs:651-657), so such a shard is always
    walked again from the exact entry state (the kernels patch the first span in place; across
    ranks a re-walk is simpler and costs no re-scan)."""
    if tp == NONE or has_looks:
        return False
    if can_match_empty:
        return tp < c_first or (tp == c_first and tl != c_first)
    return tp <= c_first


def find_all_sharded(engine, geom, comm, can_match_empty, has_looks, start=0):
    """Run the boundary protocol.  `engine.run(io: dict) -> dict` executes one shard search
    (see GpuShardEngine) and returns rev_guess, rev_left, exit_p, exit_lm, n_matches with
    positions relative to the shard buffer.

    Every round is ONE all_gather of ten integers per rank: the gathered vectors let every
    rank evaluate every rank's boundary conditions, so all ranks agree on who has to redo
    what without a second exchange.  In the common case (every guess right) the whole
    protocol is a single collective.

    Returns (n_local, global_offset, global_total, rounds); the engine holds the local spans
    (buffer-relative; add geom.buf_lo for global positions)."""
    world, me = geom.world, geom.rank
    to_buf = lambda v: v if v in (NONE, SPEC) else v - geom.buf_lo
    to_glob = lambda v: v if v in (NONE, SPEC) else v + geom.buf_lo
    owns_bytes = not (geom.a >= geom.b and not geom.is_first)
    io = dict(own_lo=geom.own_lo, own_hi=geom.own_hi, is_first=geom.is_first, is_last=geom.is_last,
              rev_entry=NO_STATE, reuse_scan=False, chain_clamped=False,
              chain_p=to_buf(start) if geom.is_first else SPEC, chain_lm=NONE)
    res = engine.run(io) if owns_bytes else None
    rounds = 0
    while True:
        if res is None:
            mine = [NO_STATE, NO_STATE, SPEC, SPEC, 0, SPEC, NONE, 0]
        else:
            mine = [res["rev_left"], res["rev_guess"], to_glob(res["exit_p"]), to_glob(res["exit_lm"]), res["n_matches"],
                    to_glob(io["chain_p"]), to_glob(io["chain_lm"]), 1]
        allv = comm.all_gather(mine + [geom.buf_lo, geom.a])  # [8] = first buffer byte, [9] = first owned byte (global)
        # ---- 1. reverse-scan states flow right-to-left: who guessed its right edge wrong? ----
        redo_rev = {}
        for g in range(world):
            if not allv[g][7] or g == world - 1:
                continue
            want = next((allv[h][0] for h in range(g + 1, world) if allv[h][7] and allv[h][0] != NO_STATE), NO_STATE)
            if want != NO_STATE and want != allv[g][1]:
                redo_rev[g] = want
        if redo_rev:
            if me in redo_rev:
                io.update(rev_entry=redo_rev[me], reuse_scan=True)
                res = engine.run(io)
            rounds += 1
            continue
        # ---- 2. iterator state flows left-to-right: whose entry assumption was wrong? -------
        redo_chain = {}
        for g in range(1, world):
            if not allv[g][7]:
                continue
            buf_lo_g, a_g = allv[g][8], allv[g][9]
            tp, tl = next(((allv[h][2], allv[h][3]) for h in range(g - 1, -1, -1) if allv[h][2] != SPEC))
            cp, cl = allv[g][5], allv[g][6]
            # an entry left of the buffer is "before my first byte" whatever its value; look-around
            # patterns then enter at the buffer's first byte as a stand-in (the engine fails loudly
            # if a reverse-on-slice scan gets that far: the left context was too short)
            clamped = tp != NONE and tp < buf_lo_g
            entry = (buf_lo_g, NONE) if clamped else (tp, NONE if (tl != NONE and tl < buf_lo_g) else tl)
            if cp == SPEC:
                ok = (clamped and not has_looks) or _spec_ok(entry[0], entry[1], a_g + 1, can_match_empty, has_looks)
            else:
                ok = (cp, cl) == entry
            if not ok:
                redo_chain[g] = entry + (clamped,)
        if redo_chain:
            if me in redo_chain:
                p_g, l_g, clamped = redo_chain[me]
                io.update(chain_p=to_buf(p_g), chain_lm=to_buf(l_g), chain_clamped=clamped, reuse_scan=True)
                res = engine.run(io)
            rounds += 1
            continue
        # ---- 3. counts -> global offsets ------------------------------------------------
        counts = [v[4] for v in allv]
        find_all_sharded.last_counts = counts  # per-rank counts (gather_spans wants them)
        return counts[me], sum(counts[:me]), sum(counts), rounds


class GpuShardEngine:
    """Shard search through the C ABI on a CUDA buffer (torch uint8 tensor).

    Give every shard its own compiled regex object: a redo round (`reuse_scan`) relies on the
    start bitmap the object kept from this shard's previous call."""

    def __init__(self, regex, d_buffer, cap):
        import torch
        self.regex = regex
        self.d_buffer = d_buffer
        self.spans = torch.empty((max(cap, 1), 2), dtype=torch.int64, device=d_buffer.device)
        self.last = None

    def run(self, io):
        self.last = self.regex.find_all_shard_device(self.d_buffer, io, self.spans)
        return self.last


def forward_sharded(run, geom, comm):
    """is_match / shortest_match / RegexSet::matches over a byte-range sharded haystack.

    `run(io) -> dict` searches this rank's shard (Regex.forward_shard_device /
    RegexSet.forward_shard_device, or the CPU stand-in of the tests) with
    io = {own_lo, own_hi, is_first, is_last, entry_state}.  The forward automaton's state flows
    left to right: every rank first guesses its entry state from its left context; ONE all_gather
    per round carries (entry used, exit state, found, first match end, four mask words); a rank
    whose guess differs from its left neighbour's exit state searches again with the exact
    state.  Returns (first_end or None, mask words OR-ed over the ranks, rounds): shortest_match
    is the smallest first_end, is_match is `first_end is not None`, RegexSet::matches is the
    mask (all-gather + OR; NCCL has no bitwise-OR reduction)."""
    world, me = geom.world, geom.rank
    owns_bytes = not (geom.a >= geom.b and not geom.is_first)
    io = dict(own_lo=geom.own_lo, own_hi=geom.own_hi, is_first=geom.is_first, is_last=geom.is_last, entry_state=NO_STATE)
    res = run(io) if owns_bytes else None
    rounds = 0
    while True:
        if res is None:
            mine = [NO_STATE, NO_STATE, 0, NONE, 0, 0, 0, 0, 0]
        else:
            fe = res["first_end"] + geom.buf_lo if res["found"] and res["first_end"] != NONE else NONE
            mine = [res["entry_used"], res["exit_state"], int(res["found"]), fe] + list(res["masks"]) + [1]
        allv = comm.all_gather(mine)
        redo = {}
        prev_exit = None  # exact exit state of the nearest rank to the left that searched
        for g in range(world):
            if not allv[g][8]:
                continue
            if prev_exit is not None and allv[g][0] != prev_exit:
                redo[g] = prev_exit
                break  # everything to the right depends on this rank's new exit state
            prev_exit = allv[g][1]
        if redo:
            if me in redo:
                io["entry_state"] = redo[me]
                res = run(io)
            rounds += 1
            continue
        ends = [v[3] for v in allv if v[8] and v[2] and v[3] != NONE]
        masks = [0, 0, 0, 0]
        for v in allv:
            if v[8]:
                for w in range(4):
                    masks[w] |= v[4 + w]
        return (min(ends) if ends else None), masks, rounds


def gather_spans(spans_local, n_local, counts, buf_lo, rank, world, dst_rank=0):
    """All ranks' spans in one buffer on `dst_rank`, in haystack order with global positions
    ("spans gathered to GPU 0 ... only if the caller wants a single buffer").
    spans_local: (cap, 2) int64 CUDA tensor of buffer-relative spans; counts: per-rank match
    counts from find_all_sharded.  NCCL point-to-point; returns the (total, 2) tensor on dst_rank,
    None elsewhere."""
    import torch
    import torch.distributed as dist
    mine = (spans_local[:n_local] + buf_lo).contiguous()
    if world == 1:
        return mine
    if rank == dst_rank:
        out = torch.empty((sum(counts), 2), dtype=torch.int64, device=spans_local.device)
        offs = [sum(counts[:g]) for g in range(world)]
        out[offs[rank]:offs[rank] + counts[rank]] = mine
        reqs = [dist.irecv(out[offs[g]:offs[g] + counts[g]], src=g) for g in range(world) if g != dst_rank and counts[g]]
        for r in reqs:
            r.wait()
        return out
    if n_local:
        dist.send(mine, dst=dst_rank)
    return None
