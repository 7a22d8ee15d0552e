"""The chunk-stitch algorithm of the kernels (tests/stitch_model.py mirrors resolve_entries /
stitch_resolve / walk_chunks) against the oracle, with tiny chunks so that every boundary case
occurs: look-arounds (reverse-on-slice rule at speculative entries, SURVEY hazard H1), empty
matches, matches longer than several chunks, chunks without candidates."""
import numpy as np
import pytest

import regex_b200 as R
from dfa_sim import Sim
from helpers import xorshift_bytes
from oracle import oracle as O
from stitch_model import chunked_find_iter
from test_fuzz_tables_vs_oracle import _pattern


def _ascii(p):
    return "(?-u)" + p if all(x not in p for x in ("α", "é", "3b1", "pL")) else p


def test_slice_rule_at_speculative_chunk_entries():
    """The round-1 gap: `^[ab]{2,}\\w*?|(?m:$)` must give (255,259)-style spans, not chunk-local ones."""
    pat = r"^[ab]{2,}\w*?|(?m:$)"
    r = R.BytesRegex(pat)
    sim, o = Sim(r), O.OracleRegex(pat)
    for seed in range(12):
        text = xorshift_bytes(seed, 300, b"abc \n")
        exp = o.find_iter(text)
        for chunk in (8, 24, 64):
            got, _ = chunked_find_iter(sim, text, chunk)
            assert got == exp, (seed, chunk)


@pytest.mark.parametrize("looks", [True, False])
def test_chunk_stitch_equals_find_iter_on_random_patterns(looks):
    rng = np.random.Generator(np.random.PCG64(0xC0FFEE + looks))
    cases = with_looks = 0
    stats = {}
    tries = 0
    while cases < (260 if looks else 200) and tries < 4000:
        tries += 1
        p = _ascii(_pattern(rng))
        try:
            r = R.BytesRegex(p)
        except R.Error:
            continue
        sim = Sim(r)
        if sim.info["has_looks"] != looks:
            continue
        o = O.OracleRegex(p)
        for seed in (1, 2):
            text = xorshift_bytes(int(rng.integers(0, 999)), 150, b"abc \n" if seed == 1 else b"ab1 _\n")
            exp = o.find_iter(text)
            for chunk in (8, 24):
                got, _ = chunked_find_iter(sim, text, chunk, stats=stats)
                assert got == exp, (p, chunk, text, got[:6], exp[:6])
                cases += 1
        with_looks += looks
    assert cases >= 200, cases


def test_long_matches_and_empty_runs_resolve_in_few_rounds():
    """A match covering many chunks, and long runs of chunks without candidates, must not cost
    one round per chunk (the prefix over contributing chunks resolves them at once)."""
    for pat, text in [(r"(?s)a.*b", b"xx a" + b"-" * 900 + b"b yy ab " + b"." * 500),
                      (r"\w+", b"ab " + b"x" * 1000 + b" cd"),
                      (r"(?m)^z$", b"q" * 700 + b"\nz\n" + b"q" * 500),
                      (r"(?-u:\b)z(?-u:\b)", b"q " * 400 + b" z " + b"q " * 300)]:
        r = R.BytesRegex(pat)
        stats = {}
        got, _ = chunked_find_iter(Sim(r), text, 8, stats=stats)
        assert got == O.OracleRegex(pat).find_iter(text), pat
        assert stats["rounds"] <= 4, (pat, stats)
