"""CPU simulation of the GPU search formulation over the product's dense tables.

TEST HELPER ONLY.  It walks the tables exported by rure_b200_dfa_export exactly
the way the kernels do (start bitmap from the reverse all-match automaton,
anchored leftmost-first runs, find_iter chain, slice emulation), but sequentially
and in Python, so the determinizer and the decomposition can be checked against
the oracle without a GPU.  Nothing in the product imports this.
"""
import regex_b200 as R

NONE = None


def _is_word(b):
    return (97 <= b <= 122) or (65 <= b <= 90) or (48 <= b <= 57) or b == 95


def flags_forward(t, at):
    n = len(t)
    f = 0
    if at == 0:
        f |= 1
    if n == 0:
        f |= 2 | 8
    if at == 0 or t[at - 1] == 10:
        f |= 4
    last = at > 0 and _is_word(t[at - 1])
    cur = at < n and _is_word(t[at])
    f |= 32 if last == cur else 16
    if last:
        f |= 64
    return f


def flags_reverse(t, at):
    n = len(t)
    f = 0
    if at == n:
        f |= 1
    if n == 0:
        f |= 2 | 8
    if at == n or t[at] == 10:
        f |= 4
    last = at < n and _is_word(t[at])
    cur = at > 0 and _is_word(t[at - 1])
    f |= 32 if last == cur else 16
    if last:
        f |= 64
    return f


class Sim:
    def __init__(self, regex):
        self.re = regex
        self.info = regex.pattern_info() if hasattr(regex, "pattern_info") else None
        self.utf8 = getattr(regex, "_only_utf8", False)
        self._d = {}

    def d(self, kind):
        if kind not in self._d:
            self._d[kind] = self.re.dfa(kind)
        return self._d[kind]

    @staticmethod
    def step(d, s, b):
        return int(d["trans"][s, d["classes"][b]])

    @staticmethod
    def eof(d, s):
        return int(d["trans"][s, d["trans"].shape[1] - 1])

    def start_bitmap(self, t):
        """S[q] for q in 0..n: some match starts at q."""
        d = self.d(R.DFA_REV_UNANCHORED_ALL)
        n = len(t)
        s = int(d["start"][flags_reverse(t, n)])
        S = [False] * (n + 1)
        for q in range(n, -1, -1):
            s = self.step(d, s, t[q - 1]) if q > 0 else self.eof(d, s)
            hit = s >= d["match_lo"]
            if hit and self.utf8 and self.info["can_match_empty"] and q < n and (t[q] & 0xC0) == 0x80:
                hit = False
            S[q] = hit
        return S

    def anchored_end(self, t, s0):
        d = self.d(R.DFA_FWD_ANCHORED_LF)
        n = len(t)
        st = int(d["start"][flags_forward(t, s0)])
        last = None
        q = s0
        while True:
            st = self.step(d, st, t[q]) if q < n else self.eof(d, st)
            if st >= d["match_lo"]:
                last = q
            if st == 0 or q >= n:
                break
            q += 1
        return last

    def slice_start(self, t, p, e):
        d = self.d(R.DFA_REV_ANCHORED_LONGEST)
        sl = t[p:]
        at = e - p
        st = int(d["start"][flags_reverse(sl, at)])
        last = None
        if st == 0:
            return None
        while at > 0:
            at -= 1
            st = self.step(d, st, sl[at])
            if st == 0:
                return None if last is None else p + last
            if st >= d["match_lo"]:
                last = at + 1
        st = self.eof(d, st)
        if st >= d["match_lo"]:
            last = 0
        return None if last is None else p + last

    def next_after_empty(self, t, i):
        if not self.utf8 or i >= len(t):
            return i + 1
        b = t[i]
        return i + (1 if b <= 0x7F else 2 if b <= 0xDF else 3 if b <= 0xEF else 4)

    def find_iter(self, t, start=0, limit=None):
        n = len(t)
        S = self.start_bitmap(t)
        emulate = self.info["has_looks"]
        out = []
        p, lm = start, None
        while p <= n:
            s = next((q for q in range(p, n + 1) if S[q]), None)
            if s is None:
                break
            e = self.anchored_end(t, s)
            assert e is not None, "start bitmap and anchored automaton disagree"
            ms = s
            if emulate and e != p:
                ms = self.slice_start(t, p, e)
                if ms is None:
                    break
            if ms == e:
                p = self.next_after_empty(t, e)
                if e == lm:
                    continue
            else:
                p = e
            lm = e
            out.append((ms, e))
            if limit and len(out) >= limit:
                break
        return out

    def find_at(self, t, start=0):
        r = self.find_iter(t, start, limit=1)
        return r[0] if r else None

    def forward_scan(self, t, start=0):
        """(first match end or None, OR of masks) over the forward all-match automaton."""
        d = self.d(R.DFA_FWD_UNANCHORED_ALL)
        n = len(t)
        if start > n:
            return None, 0
        st = int(d["start"][flags_forward(t, start)])
        first, acc = None, 0
        for q in range(start, n + 1):
            st = self.step(d, st, t[q]) if q < n else self.eof(d, st)
            if st >= d["match_lo"]:
                if first is None:
                    first = q
                for w in range(d["masks"].shape[1]):
                    acc |= int(d["masks"][st, w]) << (64 * w)
            if st == 0:
                break
        return first, acc

    def batch_find(self, t):
        """Reference algorithm per record: forward leftmost-first end, reverse longest start."""
        d = self.d(R.DFA_FWD_UNANCHORED_LF)
        n = len(t)
        st = int(d["start"][flags_forward(t, 0)])
        e = None
        q = 0
        while st != 0:
            st = self.step(d, st, t[q]) if q < n else self.eof(d, st)
            if st >= d["match_lo"]:
                e = q
            if q >= n:
                break
            q += 1
        if e is None:
            return None
        if e == 0:
            return (0, 0)
        s = self.slice_start(t, 0, e)
        return None if s is None else (s, e)
