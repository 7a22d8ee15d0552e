"""Seeded differential fuzz of the product's dense tables (walked on the CPU by dfa_sim.Sim
exactly as the kernels walk them) against the oracle's restatement of the reference's lazy-DFA
pipeline: find_iter spans, shortest_match and the batched per-record find, over random
patterns with classes, lazy/greedy repeats, alternations, Unicode classes and every kind of
look-around, on ASCII, UTF-8 and invalid-UTF-8 haystacks.  Found: the word flag of a state
was dropped when the word look sat behind another look (`(?-u:\\b)^`), and two look-around
orders no DFA scan can resolve, which are now explicit compile errors."""
import numpy as np
import pytest

import regex_b200 as R
from dfa_sim import Sim
from helpers import xorshift_bytes
from oracle import oracle as O

ATOMS = ["a", "b", "c", "\\n", " ", "[ab]", "[^a]", "[^\\n]", ".", "(?s:.)", "\\w", "\\W", "\\d", "\\s", "(?:ab|c)", "(?:a|ab)",
         "^", "$", "(?m:^)", "(?m:$)", "(?-u:\\b)", "(?-u:\\B)", "\\pL", "[a-c]", "é", "(?i:b)", "x", "\\x{3b1}",
         "[α-ω]", "(?-u:[\\x80-\\xff])"]
LOOKS = {"^", "$", "(?m:^)", "(?m:$)", "(?-u:\\b)", "(?-u:\\B)"}
QUANTS = ["", "", "", "*", "+", "?", "*?", "+?", "??", "{2}", "{1,3}", "{0,2}?", "{2,}"]
TEXTS = [xorshift_bytes(1, 120, b"abc \n"), xorshift_bytes(2, 90, b"ab1 _\n"), "aéb cé\nab aé αβ ".encode() * 3, b"", b"a", b"\n\n",
         b"ab\xff\xc3a\xa9 b", "x αβγ1\n".encode() * 4]


def _pattern(rng):
    parts = []
    for _ in range(int(rng.integers(1, 5))):
        a = ATOMS[int(rng.integers(0, len(ATOMS)))]
        q = "" if a in LOOKS else QUANTS[int(rng.integers(0, len(QUANTS)))]
        parts.append(a + q)
    p = "".join(parts)
    if rng.random() < 0.25:
        p += "|" + ATOMS[int(rng.integers(0, len(ATOMS)))]
    if rng.random() < 0.1:
        p = "(?:" + p + ")+"
    return p


def test_tables_equal_the_reference_pipeline_on_random_patterns():
    rng = np.random.Generator(np.random.PCG64(20261018))
    cases = rejected = 0
    for _ in range(420):
        p = _pattern(rng)
        for cls, utf8 in ((R.BytesRegex, False), (R.Regex, True)):
            try:
                r = cls(p)
            except R.Error as e:
                assert any(k in str(e) for k in ("look-around sequence", "exceeds size limit", "Error parsing regex")), (p, str(e))
                rejected += 1
                continue
            o = O.OracleRegex(p, only_utf8=utf8)
            sim = Sim(r)
            for t in TEXTS:
                if utf8:
                    try:
                        t.decode()
                    except UnicodeDecodeError:
                        continue
                assert sim.find_iter(t) == o.find_iter(t), (p, utf8, t)
                assert sim.forward_scan(t)[0] == o.shortest_match_at(t), (p, utf8, t)
                assert sim.batch_find(t) == o.find_at(t), (p, utf8, t)
                cases += 1
    assert cases > 3000, (cases, rejected)


@pytest.mark.parametrize("pat,ok", [(r"(?-u:\b)(?m:^)", False), (r"(?m:$)(?-u:\b)x?", False), (r"(?-u:\B) ??(?m:^)", False),
                                    (r"(?-u:\b)^", True), (r"$^", True), (r"(?m)(?:^\d+$\n?)+", True), (r"(?m)^(?-u:\b)\w+(?-u:\b)$", True)])
def test_unresolvable_look_pairs_are_explicit_errors(pat, ok):
    if ok:
        R.BytesRegex(pat)
    else:
        with pytest.raises(R.Error, match="look-around sequence"):
            R.BytesRegex(pat)


def test_set_tables_equal_forward_many_on_random_sets():
    """RegexSet: the OR of the per-state pattern masks along the forward scan must equal
    RegexSet::matches (dfa.rs:525-570) for random pattern sets."""
    rng = np.random.Generator(np.random.PCG64(0x5E7))
    cases = 0
    for _ in range(160):
        pats = [_pattern(rng) for _ in range(int(rng.integers(2, 6)))]
        for cls, utf8 in ((R.BytesRegexSet, False), (R.RegexSet, True)):
            try:
                s = cls(pats)
            except R.Error:
                continue
            o = O.OracleRegex(pats, only_utf8=utf8)
            sim = Sim(s)
            for t in TEXTS:
                if utf8:
                    try:
                        t.decode()
                    except UnicodeDecodeError:
                        continue
                _, acc = sim.forward_scan(t)
                got = [i for i in range(len(pats)) if (acc >> i) & 1]
                assert got == list(o.set_matches(t)), (pats, utf8, t)
                cases += 1
    assert cases > 800, cases
