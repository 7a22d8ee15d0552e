"""SURVEY 8a row 11 (choose_match_type): the reference picks one of several engines per pattern and
they are meant to return identical results.  The oracle restates two of them: the default
pipeline (`Dfa`: forward leftmost-first, reverse on the slice, src/exec.rs:632-662) and the
suffix-literal engine (`DfaSuffix`, src/exec.rs:725-794).  This file pins where they agree --
every BASELINE config and every sherlock benchmark pattern, which is what makes the product's
"Dfa semantics only" design bit-exact there -- and documents where the reference's DfaSuffix
engine itself departs from its own Dfa engine, so that the divergence of this backend (which
always returns the Dfa engine's spans) is explicit rather than silent."""
import numpy as np

from helpers import sherlock_counts, sherlock_text, xorshift_bytes
from oracle import oracle as O


def test_suffix_engine_equals_default_engine_on_the_benchmark_patterns():
    text = sherlock_text()
    pats = [x["re"] for x in sherlock_counts()] + [r"[a-zA-Z]+ing", r"\w+ing", r"[a-z]+ly\b".replace(r"\b", ""), r"\s[a-z]+tion"]
    used = 0
    for p in pats:
        o = O.OracleRegex(p, only_utf8=True)
        a, d = o.find_iter(text, engine=O.ENGINE_AUTO), o.find_iter(text, engine=O.ENGINE_DFA)
        assert a == d, p
        used += 1
    assert used > 35


def test_where_the_reference_suffix_engine_departs_from_its_default_engine():
    """exec.rs:737-754 continues its literal scan AFTER the end of a rejected occurrence, so an
    occurrence of a self-overlapping literal that starts inside the rejected one is never tried
    (`\\pLccc` on ` cccc`: the match is `cccc` = letter + `ccc`, but the first `ccc` is rejected
    -- a space precedes it -- and the scan resumes behind it); and its reverse scan only sees the
    text since the previous occurrence.  The Dfa engine finds these matches; so does this backend."""
    o = O.OracleRegex(r"\pLccc")
    t = b"ab cccc\naac"
    assert o.find_iter(t, engine=O.ENGINE_DFA) == [(3, 7)]
    assert o.find_iter(t, engine=O.ENGINE_AUTO) == []
    # how often: random patterns with a literal suffix of three or more characters
    from test_fuzz_tables_vs_oracle import _pattern
    rng = np.random.Generator(np.random.PCG64(99))
    sufs = ["abc", "bca", "cab", "aab", "b c", "abca", "ccc"]
    n = diff = overlapping = 0
    while n < 3000:
        p = _pattern(rng)
        if all(x not in p for x in ("α", "é", "3b1", "pL")):
            p = "(?-u)" + p
        suf = sufs[int(rng.integers(0, len(sufs)))]
        try:
            o = O.OracleRegex("(?:" + p + ")" + suf)
        except O.OracleError:
            continue
        text = xorshift_bytes(int(rng.integers(0, 99999)), int(rng.integers(20, 400)), b"abc \n")
        a, d = o.find_iter(text, engine=O.ENGINE_AUTO), o.find_iter(text, engine=O.ENGINE_DFA)
        n += 1
        if a != d:
            diff += 1
            overlapping += suf in ("ccc", "abca")
    assert 0 < diff < 0.02 * n, (diff, n)  # rare, and real: this is the reference's behaviour, not the oracle's
