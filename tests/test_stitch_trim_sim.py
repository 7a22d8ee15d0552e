"""CPU model of the chunk-parallel find_iter walk with speculation, trimming and re-walks
(csrc/gpu/kernels.cu: chunk_walk_simple, stitch_check, walk_chunks, compact_spans), fuzzed
against the oracle.  The kernels' argument -- "a speculative chain and the real one meet at
the first staged span that starts at or after the real entry point, provided the span before
it ended by then" -- is checked here on thousands of small cases with tiny chunks, passes
applied in parallel (every chunk reads its left neighbour's state of the previous pass)."""
import numpy as np
import pytest

import regex_b200 as R
from dfa_sim import Sim
from helpers import xorshift_bytes
from oracle import oracle as O

SPEC = "spec"


def chunked_find_iter(sim, text, chunk):
    n = len(text)
    S = sim.start_bitmap(text)            # S[q]: a match starts at q
    A = {}

    def end_from(s):
        if s not in A:
            A[s] = sim.anchored_end(text, s)
        return A[s]

    n_chunks = max(1, (n + chunk - 1) // chunk)

    def walk(k, p):
        """Chain over chunk k (positions cb+1..ce, plus 0 for chunk 0) entered with next search position p."""
        cb, ce = k * chunk, min((k + 1) * chunk, n)
        spans, fc = [], None
        while True:
            lo = max(p, cb + 1)
            if k == 0 and p == 0 and S[0]:
                s = 0
            else:
                s = next((q for q in range(lo, ce + 1) if S[q]), None)
            if s is None:
                break
            if fc is None:
                fc = s
            e = end_from(s)
            assert e is not None and e > s
            spans.append((s, e))
            p = e
        return spans, p, fc

    in_p = [0] + [SPEC] * (n_chunks - 1)
    state = []
    for k in range(n_chunks):
        spans, out_p, fc = walk(k, 0 if k == 0 else k * chunk + 1)
        state.append(dict(spans=spans, skip=0, out_p=out_p, fc=fc))
    rewalks = trims = 0
    for _ in range(4 * n_chunks + 8):
        outs = [st["out_p"] for st in state]
        dirty, changed = [], 0
        for k in range(1, n_chunks):
            tp, cp, st = outs[k - 1], in_p[k], state[k]
            c_first = k * chunk + 1
            if cp == SPEC:
                ok = (st["fc"] is None or tp <= st["fc"]) and tp <= c_first + chunk
            else:
                ok = cp == tp
            if ok:
                continue
            trimmed = False
            if cp == SPEC:
                live = st["spans"]
                i = next((j for j, (s, _) in enumerate(live) if s >= tp), len(live))
                if i == 0 or live[i - 1][1] <= tp:
                    trimmed = True
                    trims += 1
                    st["skip"] = i
                    if i == len(live):
                        st["out_p"] = tp
                        changed += 1
            in_p[k] = tp
            if not trimmed:
                dirty.append(k)
        for k in dirty:
            spans, out_p, fc = walk(k, in_p[k])
            state[k] = dict(spans=spans, skip=0, out_p=out_p, fc=fc)
            rewalks += 1
        if not dirty and not changed:
            break
    else:
        raise AssertionError("stitch did not converge")
    out = []
    for st in state:
        out += st["spans"][st["skip"]:]
    return out, trims, rewalks


def test_trimmed_speculation_equals_find_iter():
    rng = np.random.Generator(np.random.PCG64(0x7819))
    atoms = ["a", "b", "c", "[ab]", "[^a\\n]", "(?:ab|c)", "b+", "a{2,3}", "[a-c]{1,2}", "(?:a|bc)+", "c+?b", "a.c", "[bc]+a"]
    fixed = [r"[a-c]+b", r"a+", r"ab|abc|bca", r"[ab]+c[ab]*", r"(?:abc)+", r"c[ab]*c", r"[abc]{3}", r"b[^b]*b"]
    total_trims = total_rewalks = cases = 0
    for it in range(260):
        if it < len(fixed):
            pat = fixed[it]
        else:
            pat = "".join(atoms[int(i)] for i in rng.integers(0, len(atoms), size=int(rng.integers(1, 4))))
            if rng.random() < 0.25:
                pat += "|" + atoms[int(rng.integers(0, len(atoms)))]
        r = R.BytesRegex("(?-u)" + pat)
        info = r.pattern_info()
        if info["can_match_empty"] or info["has_looks"]:
            continue
        sim = Sim(r)
        exp_re = O.OracleRegex("(?-u)" + pat)
        for seed, n in ((it, 180), (it + 1000, 97)):
            text = xorshift_bytes(seed, n, b"abc" if seed % 3 else b"aabbc\n")
            exp = exp_re.find_iter(text)
            for chunk in (4, 8, 24):
                got, trims, rewalks = chunked_find_iter(sim, text, chunk)
                assert got == exp, (pat, chunk, text, got[:6], exp[:6])
                total_trims += trims
                total_rewalks += rewalks
                cases += 1
    assert cases > 400 and total_trims > 200, (cases, total_trims, total_rewalks)
