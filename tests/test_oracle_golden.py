"""The oracle against the reference's own known-answer tests (CPU only).

This is what pins the oracle: every mat!/matiter!/matset!/ismatch!/
shortmat! vector of tests/*.rs on both restated engines (lazy-DFA pipeline and Pike VM),
the sherlock bench counts (bench/src/sherlock.rs:31-196) and the regex-dna shootout
output (examples/regexdna-output.txt).
"""
import os
import re as pyre

import pytest

from helpers import GOLDEN, misc_is_match, sherlock_counts, sherlock_text, vectors
from oracle import oracle as O


def _run(x, mode, engine):
    text = bytes.fromhex(x["text_hex"])
    only_utf8 = mode == "str"
    if x["kind"] in ("matset", "nomatset"):
        return O.OracleRegex(x["res"], only_utf8=only_utf8).set_matches(text, engine=engine)
    r = O.OracleRegex(x["re"], only_utf8=only_utf8)
    if r.needs_unicode_word_boundary and any(b >= 0x80 for b in text):
        pytest.skip("Unicode \\b on non-ASCII text needs the reference's NFA fallback")
    if x["kind"] == "mat":
        got = r.find_at(text, engine=engine)
        return list(got) if got else None
    if x["kind"] == "matiter":
        return [list(t) for t in r.find_iter(text, engine=engine)]
    if x["kind"] == "ismatch":
        return r.is_match_at(text, engine=engine)
    if x["kind"] == "shortmat":
        return r.shortest_match_at(text, engine=O.ENGINE_DFA)  # engine-defined
    raise AssertionError(x["kind"])


@pytest.mark.parametrize("engine", [O.ENGINE_DFA, O.ENGINE_PIKEVM, O.ENGINE_AUTO], ids=["lazydfa", "pikevm", "auto-dfasuffix"])
def test_reference_vectors(engine):
    bad = []
    n = 0
    for x in vectors():
        for mode in x["modes"]:
            try:
                got = _run(x, mode, engine)
            except pytest.skip.Exception:
                continue
            n += 1
            if got != x["expected"]:
                bad.append((x["file"], x["name"], mode, x["expected"], got))
    assert n > 1100
    assert not bad, bad[:10]


def test_sherlock_counts():
    text = sherlock_text()
    assert len(text) == 594933
    for x in sherlock_counts():
        r = O.OracleRegex(x["re"], only_utf8=True)
        assert r.count(text) == x["count"], x
        # the engine the reference itself selects (DfaSuffix for e.g. [a-zA-Z]+ing, exec.rs:1176-1210);
        # this is the one bench.py times as the CPU baseline
        assert r.count(text, engine=O.ENGINE_AUTO) == x["count"], x


def test_dfasuffix_engine_equals_dfa_engine():
    """exec.rs:725-794 (suffix literal scan + reverse DFA + forward DFA) must give the spans of
    the plain DFA pipeline, including when the quadratic-behaviour guard makes it give up."""
    text = sherlock_text()[:200000]
    adversarial = b"inginginginginging" * 50 + b" xing ing singing"
    for pat in [r"[a-zA-Z]+ing", r"\w+ing", r"[a-z]+ation", r"(?s).{0,5}olmes", r"[A-Z][a-z]*son", r"\s[a-z]+ing\s",
                r"[a-z]*ing", r"x+ing|y+ing", r"(?i)[a-z]+ing", r"[a-zA-Z]+ing$"]:
        for only_utf8 in (False, True):
            r = O.OracleRegex(pat, only_utf8=only_utf8)
            for t in (text, adversarial):
                assert r.find_iter(t, engine=O.ENGINE_AUTO) == r.find_iter(t, engine=O.ENGINE_DFA), (pat, only_utf8)


def test_misc_is_match_known_answers():
    """bench/src/misc.rs bench_match!/bench_not_match! outcomes (str Regex), both oracle engines."""
    for name, pat, expected, hay in misc_is_match():
        o = O.OracleRegex(pat, only_utf8=True)
        for engine in (O.ENGINE_DFA, O.ENGINE_PIKEVM):
            assert o.is_match_at(hay, 0, engine) == expected, (name, engine)


def test_regexdna_shootout():
    """examples/shootout-regex-dna.rs:18-34 against examples/regexdna-output.txt."""
    seq = open(os.path.join(GOLDEN, "regexdna-input.txt"), "rb").read()
    expect = open(os.path.join(GOLDEN, "regexdna-output.txt")).read().split("\n")
    clean_re = O.OracleRegex(r">[^\n]*\n|\n", only_utf8=True)
    spans = clean_re.find_iter(seq)
    out, last = [], 0
    for s, e in spans:
        out.append(seq[last:s])
        last = e
    out.append(seq[last:])
    clean = b"".join(out)
    assert len(seq) == int(expect[10]) and len(clean) == int(expect[11])
    variants = [
        "agggtaaa|tttaccct", "[cgt]gggtaaa|tttaccc[acg]", "a[act]ggtaaa|tttacc[agt]t",
        "ag[act]gtaaa|tttac[agt]ct", "agg[act]taaa|ttta[agt]cct", "aggg[acg]aaa|ttt[cgt]ccct",
        "agggt[cgt]aa|tt[acg]accct", "agggta[cgt]a|t[acg]taccct", "agggtaa[cgt]|[acg]ttaccct",
    ]
    for i, v in enumerate(variants):
        name, count = expect[i].rsplit(" ", 1)
        assert name == v
        assert O.OracleRegex(v, only_utf8=True).count(clean) == int(count)


def test_empty_match_iteration():
    """tests/api.rs:1-69 (hand-ported: those tests are plain #[test] fns, not macros)."""
    cases = [
        (r"", "", [(0, 0)]), (r"", "a", [(0, 0), (1, 1)]), (r"", "ab", [(0, 0), (1, 1), (2, 2)]),
        (r"a*", "", [(0, 0)]), (r"a*", "a", [(0, 1)]), (r"a*", "b", [(0, 0), (1, 1)]),
        (r"a*", "ab", [(0, 1), (2, 2)]), (r"a*", "ba", [(0, 0), (1, 2)]),
        (r"a|b", "ab", [(0, 1), (1, 2)]), (r"\s*", "  a ", [(0, 2), (3, 4)]),
    ]
    for pat, text, exp in cases:
        for eng in (O.ENGINE_DFA, O.ENGINE_PIKEVM):
            assert O.OracleRegex(pat, only_utf8=True).find_iter(text.encode(), engine=eng) == exp, (pat, text)


def test_python_re_crosscheck():
    """Independent cross-check on a syntax subset where Python's backtracker agrees
    (non-empty matches, no Unicode classes): guards the shared front end."""
    text = sherlock_text()[:60000]
    for pat in [r"Sherlock|Holmes", r"[a-zA-Z]+ing", r"(?i)the", r"Sher[a-z]+|Hol[a-z]+", r"[0-9]+", r"th[a-z]{2,4}e",
                r"(?:a|b)+c", r"[^a-z \r\n]{3,}", r"q[a-z]*u?x?"]:
        exp = [(m.start(), m.end()) for m in pyre.finditer(pat.encode(), text)]
        # Python's bytes patterns are byte-oriented: compare with the (?-u) flavour.
        assert O.OracleRegex("(?-u)" + pat, only_utf8=False).find_iter(text) == exp, pat


def test_oracle_captures_equal_the_reference_vectors():
    """Every capture group the reference's mat! vectors list (tests/macros.rs:22-52 compares as many
    groups as the test names): the oracle's Pike VM with slots (src/pikevm.rs:130-352) after the
    DFA pipeline (src/exec.rs:527-590, 861-875)."""
    from helpers import vectors
    n = multi = 0
    for x in vectors():
        if x["kind"] != "mat":
            continue
        text = bytes.fromhex(x["text_hex"])
        for mode in x["modes"]:
            try:
                o = O.OracleRegex(x["re"], only_utf8=mode == "str")
            except O.OracleError:
                continue
            if o.needs_unicode_word_boundary:
                continue
            got = o.captures_at(text)
            exp = x["groups"]
            if exp[0] is None:
                assert got is None, (x["name"], got)
                continue
            assert got is not None, x["name"]
            got = [list(g) if g is not None else None for g in got][:len(exp)]
            assert got == exp, (x["name"], x["re"], got, exp)
            n += 1
            multi += len(exp) > 1
    assert n > 700 and multi > 300, (n, multi)
