"""CPU stand-in for rure_b200_find_all_shard_device (TEST HELPER ONLY).

Walks the product's exported tables over one shard buffer exactly like the kernels do
(tests/dfa_sim.py), including a deliberately weak guess of the reverse-scan state at
the shard's top edge, so the boundary protocol in regex_b200/sharded.py can be tested
under gloo without a GPU."""
import regex_b200 as R
from dfa_sim import Sim, flags_forward, flags_reverse
from regex_b200.sharded import NO_STATE, NONE, SPEC


class SimShardEngine:
    def __init__(self, regex, buf, warm=0):
        self.sim = Sim(regex)
        self.buf = buf
        self.warm = warm
        self.spans = []

    def run(self, io):
        s, t, n = self.sim, self.buf, len(self.buf)
        lo, hi = io["own_lo"], io["own_hi"]
        d = s.d(R.DFA_REV_UNANCHORED_ALL)
        # reverse scan of bytes [lo, hi): entry state guessed from `warm` halo bytes unless told
        if io["rev_entry"] != NO_STATE:
            st = io["rev_entry"]
            guess = st
        else:
            w = min(hi + self.warm, n) if not io["is_last"] else n
            st = int(d["start"][flags_reverse(t, w)])
            for i in range(w - 1, hi - 1, -1):
                st = s.step(d, st, t[i])
            guess = st
        S = {}
        for i in range(hi - 1, lo - 1, -1):
            st = s.step(d, st, t[i])
            if st >= d["match_lo"]:
                S[i + 1] = True
        rev_left = st
        if lo == 0 and io["is_first"] and s.eof(d, st) >= d["match_lo"]:
            S[0] = True
        # chain walk over the owned candidates
        info = s.info
        p, lm = io["chain_p"], io["chain_lm"]
        chain = p != SPEC
        if not chain:
            p, lm = lo + 1, NONE
        out = []
        cands = sorted(S)
        spec_entry = not chain
        moved = False
        while p != NONE:
            c = next((q for q in cands if q >= p), None)
            if c is None:
                break
            moved = True
            e = s.anchored_end(t, c) if io["is_last"] else self._anchored_end_halo(t, c)
            ms = c
            if info["has_looks"] and chain and e != p:
                ms = self._slice_start(t, p, e, io.get("chain_clamped") and p == io["chain_p"])
                if ms is None:
                    p = NONE
                    break
            chain = True
            if ms == e:
                p = s.next_after_empty(t, e)
                if e == lm:
                    continue
            else:
                p = e
            lm = e
            out.append((ms, e))
        self.spans = out
        if spec_entry and not moved:  # no candidate at all: the iterator passes through (engine: exit SPEC)
            p, lm = SPEC, NONE
        return dict(rev_guess=guess, rev_left=rev_left, exit_p=p, exit_lm=lm, n_matches=len(out))

    def _slice_start(self, t, p, e, clamped):
        """dfa_sim.Sim.slice_start; with a clamped entry, a reverse scan that reaches p alive would
        depend on bytes left of the buffer (engine: left_ctx_short error)."""
        s = self.sim
        d = s.d(R.DFA_REV_ANCHORED_LONGEST)
        sl = t[p:]
        at = e - p
        st = int(d["start"][flags_reverse(sl, at)])
        last = None
        if st == 0:
            return None
        while at > 0:
            at -= 1
            st = s.step(d, st, sl[at])
            if st == 0:
                return None if last is None else p + last
            if st >= d["match_lo"]:
                last = at + 1
        assert not clamped, "left context too short"
        st = s.eof(d, st)
        if st >= d["match_lo"]:
            last = 0
        return None if last is None else p + last

    def _anchored_end_halo(self, t, c):
        """anchored_end where running into the end of the buffer is a halo overflow, not EOF."""
        s = self.sim
        d = s.d(R.DFA_FWD_ANCHORED_LF)
        st = int(d["start"][flags_forward(t, c)])
        last, q = None, c
        while True:
            assert q < len(t), "halo too short for this match"
            st = s.step(d, st, t[q])
            if st >= d["match_lo"]:
                last = q
            if st == 0:
                return last
            q += 1


class SimFwdShard:
    """CPU stand-in for rure_b200_shortest_match_shard_device / rure_b200_set_matches_shard_device:
    the forward all-match automaton over the shard's own bytes, entry state guessed from the left
    context unless given (same rule as Regex::forward_shard_device)."""

    def __init__(self, regex, buf, weak_guess=False):
        self.sim = Sim(regex)
        self.buf = buf
        self.weak_guess = weak_guess  # guess from a single context byte: wrong more often (tests)

    def run(self, io):
        s, t, n = self.sim, self.buf, len(self.buf)
        d = s.d(R.DFA_FWD_UNANCHORED_ALL)
        lo, hi = io["own_lo"], io["own_hi"]
        st = io["entry_state"]
        if st == NO_STATE:
            if lo == 0:
                st = int(d["start"][flags_forward(t, 0)])
            else:
                frm = max(lo - (1 if self.weak_guess else 255), 0)
                if frm == 0 and not io["is_first"]:
                    frm = 1
                st = int(d["start"][flags_forward(t, frm)]) if frm > 0 or io["is_first"] else None
                for i in range(frm, lo):
                    st = s.step(d, st, t[i])
        used = st
        first, masks = None, [0, 0, 0, 0]
        end = n + 1 if io["is_last"] else hi
        for q in range(lo, end):
            st = s.step(d, st, t[q]) if q < n else s.eof(d, st)
            if st >= d["match_lo"]:
                if first is None:
                    first = q
                for w in range(d["masks"].shape[1]):
                    masks[w] |= int(d["masks"][st, w])
        return dict(entry_used=used, exit_state=st, found=first is not None, first_end=first if first is not None else NONE, masks=masks)
