"""Literal prefilter kernel (literal_scan; reference analogue: LiteralSearcher / memchr sets / Teddy,
src/literals.rs:92-102,353-371,466-489, src/simd_accel/teddy128.rs:435-675, and the prefix skip
of the DFA loop, src/dfa.rs:700-711): prefilter on == DFA path (force_generic) == oracle."""
import numpy as np
import pytest

import regex_b200 as R
from helpers import sherlock_text, tiled_corpus, xorshift_bytes
from oracle import oracle as O

pytestmark = pytest.mark.gpu

ELIGIBLE = [r"Holmes|Watson", r"Sherlock|Holmes", r"Sher[a-z]+|Hol[a-z]+", r"Sherlock Holmes", r"Watson", r"[HW]atson",
            r"(?-u)Q\w+", r"Zz+|Xx", r"Mr\. \w+"]
NOT_ELIGIBLE = [r"[a-zA-Z]+ing", r"the\s+\w+", r"(?i)Sherlock|Holmes|Watson", r"\w+", r"agggtaaa|tttaccct", r"(?m)^Holmes", r"H*"]


def _spans(a):
    return [tuple(int(v) for v in r) for r in np.asarray(a).reshape(-1, 2).tolist()]


def _forced(r):
    r.set_option("prefilter", 2)  # whenever the pattern qualifies structurally (automatic mode wants ONE rare byte)
    return r


def test_prefilter_is_chosen_for_rare_byte_patterns_only():
    text = sherlock_text()[:200000]
    for pat in ELIGIBLE:
        r = _forced(R.BytesRegex(pat))
        r.find_all(text)
        assert r.last_stats()["path"] == 3, pat
        r.set_option("prefilter", 0)
        r.find_all(text)
        assert r.last_stats()["path"] != 3, pat
    for pat in NOT_ELIGIBLE:
        r = _forced(R.BytesRegex(pat))
        r.find_all(text)
        assert r.last_stats()["path"] != 3, pat
    # automatic mode: one byte of small estimated frequency
    for pat, auto in ((r"Qu[a-z]+", True), (r"@\w+", True), (r"Zz+|Xx", False), (r"Holmes|Watson", False), (r"Watson", False)):
        r = R.BytesRegex(pat)
        assert _spans(r.find_all(text)) == O.OracleRegex(pat).find_iter(text)
        assert (r.last_stats()["path"] == 3) == auto, pat


@pytest.mark.parametrize("utf8", [False, True])
def test_prefilter_equals_dfa_path_and_oracle(utf8):
    text = sherlock_text()
    cls = R.Regex if utf8 else R.BytesRegex
    for pat in ELIGIBLE:
        exp = O.OracleRegex(pat, only_utf8=utf8).find_iter(text)
        r = _forced(cls(pat))
        got = _spans(r.find_all(text))
        assert r.last_stats()["path"] == 3
        assert got == exp, (pat, got[:4], exp[:4])
        g = cls(pat)
        g.force_generic(True)
        assert _spans(g.find_all(text)) == exp, pat
        for cut in (0, 1, 15, 16, 17, 511, 512, 2047, 2048, 2049, 8191, 8192, 8193, 20000):  # ragged ends, chunk edges
            assert _spans(r.find_all(text[:cut])) == O.OracleRegex(pat, only_utf8=utf8).find_iter(text[:cut]), (pat, cut)
        for start in (3, 7, 600, 8190):
            assert r.find_at(text, start) == O.OracleRegex(pat, only_utf8=utf8).find_at(text, start), (pat, start)


def test_prefilter_matches_at_every_alignment_and_chunk_edge():
    """Matches planted so that they straddle 16-byte pieces, 512-byte warp rows, 2 KiB iterations and
    8 KiB chunks; candidates whose fingerprint fails; the byte at every offset of a piece."""
    rng = np.random.Generator(np.random.PCG64(5))
    for pat, lit in ((r"Holmes|Watson", [b"Holmes", b"Watson", b"Holmez", b"Wat", b"H", b"WWatson", b"HolmesHolmes"]),
                     (r"Sher[a-z]+|Hol[a-z]+", [b"Sherlock", b"Hol", b"Holm", b"SherX", b"Sher", b"Holabcdefghijklmnopqrstuvwxyzabcdefghij"])):
        buf = bytearray(xorshift_bytes(9, 70000, b"abcdefgh ijklmn\n"))
        for edge in list(range(8192, 65536, 8192)) + [512, 2048, 4096, 16, 6000]:
            for d in range(-9, 3):
                w = lit[int(rng.integers(0, len(lit)))]
                pos = edge + d
                buf[pos:pos + len(w)] = w
        text = bytes(buf)
        exp = O.OracleRegex(pat).find_iter(text)
        r = _forced(R.BytesRegex(pat))
        assert _spans(r.find_all(text)) == exp, pat
        assert r.last_stats()["path"] == 3


def test_prefilter_dense_and_overlapping_candidates():
    """Adversarial haystacks: every byte is the scanned byte (lane-by-lane path), more matches per chunk
    than staging slots (overflow pass), a match longer than a chunk (general stitch loop)."""
    cases = [(r"HH", b"H" * 50001), (r"HHH|HW", b"HHWHHHHW" * 5000), (r"Holmes", b"Holmes" * 20000),
             (r"Hx*", b"ab" + b"H" + b"x" * 40000 + b" H Hxx " + b"y" * 9000 + b"Hx"), (r"Ho+|Wo", b"Hooo Wo " * 9000 + b"H" + b"o" * 20000)]
    for pat, text in cases:
        exp = O.OracleRegex(pat).find_iter(text)
        r = _forced(R.BytesRegex(pat))
        got = _spans(r.find_all(text))
        assert r.last_stats()["path"] == 3, pat
        assert got == exp, (pat, got[:5], exp[:5], len(got), len(exp))
        assert r.count_all(text) == len(exp)


@pytest.mark.parametrize("world", [2, 3])
def test_prefilter_shards(world):
    import torch
    from test_gpu_parity import _run_gpu_shards
    text = tiled_corpus(1 << 20) + b"Holmes" * 3000 + tiled_corpus(1 << 19, seed=3)
    d = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    for pat in (r"Holmes|Watson", r"Sher[a-z]+|Hol[a-z]+"):
        exp = O.OracleRegex(pat).find_iter(text)
        assert _run_gpu_shards(pat, text, d, world, len(exp), halo=4096, options={"prefilter": 2}) == exp, pat


def test_prefilter_large_haystack_and_pipelined_host_path():
    import torch
    base = tiled_corpus(16 << 20)
    reps = 16
    d = torch.frombuffer(bytearray(base), dtype=torch.uint8).cuda().repeat(reps)
    for pat in (r"Holmes|Watson", r"Sherlock|Holmes"):
        base_spans = O.OracleRegex(pat).find_iter(base)
        r = _forced(R.BytesRegex(pat))
        out = torch.empty((len(base_spans) * reps + 4096, 2), dtype=torch.int64, device="cuda")
        total = r.find_all_device(d, out)
        assert r.last_stats()["path"] == 3
        got = out[:total].cpu().numpy()
        assert (got[:, 0] < got[:, 1]).all() and (got[1:, 0] >= got[:-1, 1]).all()
        first = got[got[:, 1] <= len(base) - 64]
        exp = np.array([s for s in base_spans if s[1] <= len(base) - 64], dtype=np.int64).reshape(-1, 2)
        assert first.shape == exp.shape and (first == exp).all(), pat
        g = R.BytesRegex(pat)
        g.set_option("prefilter", 0)
        assert g.find_all_device(d) == total
        assert r.find_all_device(d) == total  # count-only mode
    # host haystack through the pipelined upload (pieces are shards)
    host = base * 10
    r = _forced(R.BytesRegex(r"Holmes|Watson"))
    got = r.find_all(host)
    one = np.array(O.OracleRegex(r"Holmes|Watson").find_iter(base + base[:64]), dtype=np.int64).reshape(-1, 2)
    g = R.BytesRegex(r"Holmes|Watson")
    g.set_option("prefilter", 0)
    ref = g.find_all(host)
    assert got.shape == ref.shape and (np.asarray(got) == np.asarray(ref)).all()
    k = int((one[:, 1] <= len(base)).sum())
    assert (np.asarray(got[:k], dtype=np.int64) == one[:k]).all()
