"""Independent cross-check of the shared front end (parser, Unicode tables, compiler): the oracle's
find_iter spans against the third-party `regex` module -- its own parser, its own Unicode
database, a backtracking matcher -- on the reference's sherlock benchmark patterns
(bench/src/sherlock.rs:31-196) and on random patterns.  The oracle links the product's front end
(oracle/oracle.cpp), so a parser or class-table bug would be common to product and oracle; it is
not common to this engine.  Patterns that can match the empty string are left out: the
reference's iteration rule for empty matches differs from Perl-style engines."""
import bisect

import numpy as np
import pytest

regex = pytest.importorskip("regex")

from helpers import sherlock_counts, sherlock_text, xorshift_bytes
from oracle import oracle as O

SLOW_FOR_BACKTRACKING = {"holmes_coword_watson", "holmes_cochar_watson", "repeated_class_negation", "everything_greedy", "everything_greedy_nl"}


def _byte_offsets(s):
    """char index -> byte index in the UTF-8 encoding of s."""
    acc, out = 0, [0]
    for ch in s:
        acc += len(ch.encode("utf-8"))
        out.append(acc)
    return out


def test_sherlock_patterns_spans_equal_python_regex():
    raw = sherlock_text()
    text = raw.decode("utf-8")
    off = _byte_offsets(text)
    n = 0
    for x in sherlock_counts():
        if x["name"] in SLOW_FOR_BACKTRACKING:
            continue
        exp = [(off[m.start()], off[m.end()]) for m in regex.finditer(x["re"], text)]
        assert len(exp) == x["count"], x["name"]  # and both equal the count the reference publishes
        got = O.OracleRegex(x["re"], only_utf8=True).find_iter(raw)
        assert got == exp, (x["name"], got[:3], exp[:3])
        n += 1
    assert n >= 30


ATOMS = ["a", "b", "c", "[ab]", "[^a]", "[^\\n]", ".", "\\w", "\\W", "\\d", "\\s", "(?:ab|c)", "(?:a|ab)", "\\pL", "[a-c]", "é", "(?i:b)", "x",
         "α", "[α-ω]", "\\p{Greek}", "[^\\W\\d]", "\\S", "(?i:é)"]
QUANTS = ["", "", "", "+", "{2}", "{1,3}", "{2,}", "+?", "{1,2}?"]
TEXTS = [xorshift_bytes(1, 400, b"abc \n").decode(), xorshift_bytes(2, 300, b"ab1 _\n").decode(), "aéb cé\nab aé αβ Éx ".replace("", "") * 6,
         "x αβγ1\nÀb ß ".replace("", "") * 8]


def test_random_patterns_spans_equal_python_regex():
    rng = np.random.Generator(np.random.PCG64(77))
    cases = 0
    for _ in range(400):
        parts = [ATOMS[int(rng.integers(0, len(ATOMS)))] + QUANTS[int(rng.integers(0, len(QUANTS)))] for _ in range(int(rng.integers(1, 4)))]
        p = "".join(parts)
        if rng.random() < 0.25:
            p += "|" + ATOMS[int(rng.integers(0, len(ATOMS)))]
        try:
            o = O.OracleRegex(p, only_utf8=True)
        except O.OracleError:
            continue
        py = regex.compile(p)
        for t in TEXTS:
            off = _byte_offsets(t)
            exp = [(off[m.start()], off[m.end()]) for m in py.finditer(t)]
            got = o.find_iter(t.encode("utf-8"))
            assert got == exp, (p, t[:40], got[:4], exp[:4])
            cases += 1
    assert cases > 1200, cases
