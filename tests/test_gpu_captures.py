"""Capture groups on the GPU (the reference runs its NFA only over the window the DFA
narrowed down, src/exec.rs:527-590, 861-875): rure_find_captures / rure_iter_next_captures /
rure_b200_captures_all against the reference's mat! vectors (every listed group) and against
the oracle's Pike VM on larger haystacks."""
import numpy as np
import pytest

import regex_b200 as R
from helpers import sherlock_text, vectors, xorshift_bytes
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_reference_vectors_all_groups():
    n = multi = 0
    for x in vectors():
        if x["kind"] != "mat":
            continue
        text = bytes.fromhex(x["text_hex"])
        for mode in x["modes"]:
            try:
                r = (R.Regex if mode == "str" else R.BytesRegex)(x["re"])
            except R.Error as e:
                assert "word boundar" in str(e), (x["name"], str(e))
                continue
            got = r.captures(text)
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


def test_captures_iter_against_the_oracle():
    text = sherlock_text()[:400000]
    cases = [r"(\w+)\s+(Holmes|Watson)", r"(?P<first>[A-Z][a-z]+) (?P<last>[A-Z][a-z]+)", r"(a+)(b*)|(c)", r"(\d+)(?:st|nd|rd|th)?",
             r"([a-z]+)(ing|ed)\b".replace(r"\b", ""), r"(?m)^(\w+)(.*)$", r"((((t)h)e))", r"(x*)", r"(?i)(sherlock)|(holmes)"]
    for pat in cases:
        for cls, utf8 in ((R.BytesRegex, False), (R.Regex, True)):
            r, o = cls(pat), O.OracleRegex(pat, only_utf8=utf8)
            got = r.captures_iter(text)
            spans = o.find_iter(text)
            assert [g[0] for g in got] == spans, pat
            for i in list(range(0, len(spans), max(1, len(spans) // 300)))[:400]:
                assert got[i] == o.captures_at(text, spans[i][0]), (pat, i)
            assert r.captures_len() == len(got[0]) if got else True
    r = R.BytesRegex(r"(?P<year>\d{4})-(?P<month>\d{2})-(?P<day>\d{2})")
    assert r.capture_name_index("month") == 2 and r.capture_name_index("nope") is None
    assert r.captures(b"on 2024-05-17, then") == [(3, 13), (3, 7), (8, 10), (11, 13)]
    assert r.captures_at(b"2024-05-17 2025-06-18", 1) == [(11, 21), (11, 15), (16, 18), (19, 21)]
    assert r.captures(b"no date") is None


def test_captures_random_patterns():
    from test_fuzz_tables_vs_oracle import _pattern
    rng = np.random.Generator(np.random.PCG64(0xCA9))
    n = 0
    while n < 250:
        parts = ["(" + _pattern(rng) + ")" if rng.random() < 0.7 else _pattern(rng) for _ in range(int(rng.integers(1, 4)))]
        p = "".join(parts)
        if all(x not in p for x in ("α", "é", "3b1", "pL")):
            p = "(?-u)" + p
        try:
            r = R.BytesRegex(p)
        except R.Error:
            continue
        o = O.OracleRegex(p)
        text = xorshift_bytes(int(rng.integers(0, 1 << 30)), int(rng.integers(50, 600)), b"abc \n" if rng.random() < 0.7 else b"ab1 _\n")
        assert r.captures_iter(text) == o.captures_iter(text), (p, text)
        assert r.captures(text) == o.captures_at(text), p
        n += 1
