"""Every BASELINE.json config as a parity point (the workloads bench.py and tools/bench_configs.py
time): C1 regex-dna is tests/test_gpu_parity.py::test_regexdna_shootout and C2 spans are
test_large_haystack_properties / bench.py's own span-exact windows; here C3 (>= 1 M log lines,
every line against the oracle), C4 (the exact 64-pattern RegexSet: whole-haystack mask and
per-line masks on 64 MiB) and C5 (regex-dna variants over 256 MiB of uniform acgt, every span)."""
import os
import sys

import numpy as np
import pytest

import regex_b200 as R
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


def test_c3_log_lines_every_line_against_the_oracle():
    import torch
    import bench_configs as BC
    dev = torch.device("cuda", 0)
    n_lines = 1_000_000
    text, offsets = BC.log_lines(n_lines, dev)
    host = text.cpu().numpy().tobytes()
    offs = offsets.cpu().numpy()
    for pat in (r"(\d{4})-(\d{2})-(\d{2})", r"(?-u)(\d{4})-(\d{2})-(\d{2})"):
        r = R.BytesRegex(pat)
        bits = torch.zeros((n_lines + 31) // 32, dtype=torch.int32, device=dev)
        spans = torch.empty((n_lines, 2), dtype=torch.int64, device=dev)
        r.is_match_batch_device(text, offsets, bits)
        is_bits = np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder="little")[:n_lines].astype(bool)
        r.find_batch_device(text, offsets, spans, bits)
        found = np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder="little")[:n_lines].astype(bool)
        sp = spans.cpu().numpy()
        o = O.OracleRegex(pat)
        exp_found = np.zeros(n_lines, dtype=bool)
        exp_sp = np.zeros((n_lines, 2), dtype=np.int64)
        for i in range(n_lines):
            m = o.find_at(host[offs[i]:offs[i + 1]])
            if m is not None:
                exp_found[i] = True
                exp_sp[i] = m
        assert (found == exp_found).all() and (is_bits == exp_found).all(), pat
        assert (sp[exp_found] == exp_sp[exp_found]).all(), pat
        assert 0.6 < exp_found.mean() < 0.8  # ISO dates on ~70 % of the lines


def test_c4_regex_set_of_64_whole_haystack_and_per_line():
    import torch
    import bench as B
    import bench_configs as BC
    dev = torch.device("cuda", 0)
    pats = BC.c4_patterns()
    assert len(pats) == 64
    n = 64 << 20
    d = B.device_corpus(n, B.SEED, dev)
    text = d.cpu().numpy().tobytes()
    s = R.BytesRegexSet(pats)
    # whole haystack: RegexSet::matches is "pattern i matches somewhere" (re_set.rs:184-191)
    exp = [i for i, p in enumerate(pats) if O.OracleRegex(p).is_match_at(text)]
    for narrow in (1, 0):
        s.set_option("narrow_sets", narrow)
        assert s.matches_device(d) == exp, narrow
    small = text[:1 << 20]
    assert s.matches(small) == list(O.OracleRegex(pats).set_matches(small)) == [i for i, p in enumerate(pats) if O.OracleRegex(p).is_match_at(small)]
    # per line (record = line with its "\n")
    nl = torch.nonzero(d == 10).flatten() + 1
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), nl])
    n_lines = offsets.numel() - 1
    masks = torch.zeros((n_lines, 1), dtype=torch.int64, device=dev)
    s.matches_batch_device(d, offsets, masks)
    got = masks.cpu().numpy().view(np.uint64)[:, 0]
    offs = offsets.cpu().numpy()
    # (a) literal members: every line, by plain substring search over the whole text
    arr = np.frombuffer(text, dtype=np.uint8)
    checked = 0
    for i, p in enumerate(pats):
        if not p.isalpha():
            continue
        lit = p.encode()
        hits, at = [], text.find(lit)
        while at >= 0:
            hits.append(at)
            at = text.find(lit, at + 1)
        exp_lines = np.zeros(n_lines, dtype=bool)
        if hits:
            h = np.asarray(hits)
            h = h[h + len(lit) <= offs[-1]]
            exp_lines[np.searchsorted(offs, h, side="right") - 1] = True
        assert (((got >> np.uint64(i)) & np.uint64(1)).astype(bool) == exp_lines).all(), p
        checked += 1
    assert checked >= 35
    # (b) all 64 members against the oracle's RegexSet on a sample of lines spread over the haystack
    o = O.OracleRegex(pats)
    rng = np.random.Generator(np.random.PCG64(4))
    for i in np.concatenate([np.arange(1500), rng.integers(0, n_lines, size=2500)]):
        line = text[offs[i]:offs[i + 1]]
        e = sum(1 << j for j in o.set_matches(line))
        assert int(got[i]) == e, (int(i), line)


def test_c5_dna_variants_every_span():
    import torch
    import bench_configs as BC
    n = 256 << 20
    rng = np.random.Generator(np.random.PCG64(0x5EED0005))
    arr = np.frombuffer(b"acgt", dtype=np.uint8)[rng.integers(0, 4, size=n, dtype=np.uint8)]
    text = arr.tobytes()
    d = torch.from_numpy(arr).cuda()
    for v in BC.DNA_VARIANTS:
        exp = np.array(O.OracleRegex(v).find_iter(text), dtype=np.int64).reshape(-1, 2)
        r = R.BytesRegex(v)
        out = torch.empty((exp.shape[0] + 64, 2), dtype=torch.int64, device="cuda")
        total = r.find_all_device(d, out)
        assert total == exp.shape[0], v
        assert (out[:total].cpu().numpy() == exp).all(), v
        assert abs(total / n - (2 if v == BC.DNA_VARIANTS[0] else 6) / 4 ** 8) < 2e-5
