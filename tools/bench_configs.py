#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (C1, C3, C4, C5) on one GPU, device-resident inputs.

bench.py measures the headline C2 metric; this script prints one JSON line per other config
so that their throughput is on record (profiles/).  Results are checked against the oracle
on a bounded sample where that is cheap.  Usage:
    python tools/bench_configs.py [--lines 20000000] [--gib 4] [--skip c1,c3,c4,c5]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import regex_b200 as R
import bench as B

GIB = 1 << 30
DNA_VARIANTS = ["agggtaaa|tttaccct", "[cgt]gggtaaa|tttaccc[acg]", "a[act]ggtaaa|tttacc[agt]t", "ag[act]gtaaa|tttac[agt]ct",
                "agg[act]taaa|ttta[agt]cct", "aggg[acg]aaa|ttt[cgt]ccct", "agggt[cgt]aa|tt[acg]accct", "agggta[cgt]a|t[acg]taccct",
                "agggtaa[cgt]|[acg]ttaccct"]


def timed(fn, steps=5, warmup=3):
    """Wall clock around `steps` calls with a device-wide synchronize on both sides (the
    library may launch on its own stream, which torch events would not see)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3, out


def use_torch_stream(r):
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    return r


def c1(dev):
    raw = open(os.path.join(ROOT, "tests", "golden", "regexdna-input.txt"), "rb").read()
    import re
    seq = re.sub(rb">[^\n]*\n|\n", b"", raw)
    expected = {}
    for line in open(os.path.join(ROOT, "tests", "golden", "regexdna-output.txt")):
        parts = line.split()
        if len(parts) == 2 and parts[1].isdigit() and not parts[0].isdigit():
            expected[parts[0]] = int(parts[1])
    d = torch.frombuffer(bytearray(seq), dtype=torch.uint8).to(dev)
    out = {}
    for v in DNA_VARIANTS:
        r = use_torch_stream(R.BytesRegex(v))
        ms, cnt = timed(lambda: r.find_all_device(d), steps=20)
        assert cnt == expected[v], (v, cnt, expected[v])
        out[v] = {"us": round(ms * 1e3, 1), "count": cnt}
    print(json.dumps({"config": "C1 regex-dna 9 variants, find_iter count over the cleaned 100 KB input (latency)", "results": out}), flush=True)


def log_lines(n_lines, dev, seed=0x5EED0003):
    """Synthetic log corpus built on the device: fixed-width fields so it vectorises.
    line = '<ts> host-hhhh svc[ppppp]: <msg>\\n', ts = ISO date-time w.p. 0.7 else epoch seconds."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    vocab = sorted({w for w in open(os.path.join(ROOT, "tests", "golden", "sherlock.txt"), "rb").read().split() if w.isalpha() and 3 <= len(w) <= 8})[:4096]
    W = 9
    vt = torch.zeros((len(vocab), W), dtype=torch.uint8)
    vl = torch.zeros(len(vocab), dtype=torch.int64)
    for i, w in enumerate(vocab):
        vt[i, :len(w)] = torch.frombuffer(bytearray(w), dtype=torch.uint8)
        vt[i, len(w)] = 32
        vl[i] = len(w) + 1
    vt, vl = vt.to(dev), vl.to(dev)
    chunks, lens = [], []
    step = 2_000_000
    for lo in range(0, n_lines, step):
        m = min(step, n_lines - lo)
        iso = torch.rand(m, generator=g, device=dev) < 0.7
        digits = torch.randint(0, 10, (m, 20), generator=g, device=dev, dtype=torch.int64) + 48
        ts = digits.clone()
        # YYYY-MM-DDTHH:MM:SSZ
        for pos, ch in ((4, "-"), (7, "-"), (10, "T"), (13, ":"), (16, ":"), (19, "Z")):
            ts[:, pos] = ord(ch)
        ts[:, 0] = 50; ts[:, 1] = 48  # 20xx
        epoch = digits.clone()
        epoch[:, 10:] = 0  # 10 digits, rest removed
        head = torch.where(iso[:, None], ts, epoch).to(torch.uint8)
        head_len = torch.where(iso, 20, 10)
        mid = torch.frombuffer(bytearray(b" host-0000 svc[00000]: "), dtype=torch.uint8).to(dev).repeat(m, 1)
        mid[:, 6:10] = (torch.randint(0, 10, (m, 4), generator=g, device=dev) + 48).to(torch.uint8)
        mid[:, 15:20] = (torch.randint(0, 10, (m, 5), generator=g, device=dev) + 48).to(torch.uint8)
        nw = torch.randint(3, 13, (m,), generator=g, device=dev)
        wid = torch.randint(0, len(vocab), (m, 12), generator=g, device=dev)
        words = vt[wid]                       # m x 12 x W
        wl = vl[wid]                          # m x 12
        wl = torch.where(torch.arange(12, device=dev)[None, :] < nw[:, None], wl, 0)
        # assemble into a padded row, then compact by a keep mask
        row = torch.cat([head, mid, words.reshape(m, 12 * W), torch.full((m, 1), 10, dtype=torch.uint8, device=dev)], dim=1)
        keep = torch.zeros_like(row, dtype=torch.bool)
        keep[:, :20] = torch.arange(20, device=dev)[None, :] < head_len[:, None]
        keep[:, 20:20 + mid.shape[1]] = True
        wkeep = torch.arange(W, device=dev)[None, None, :] < wl[:, :, None]
        keep[:, 20 + mid.shape[1]:-1] = wkeep.reshape(m, 12 * W)
        keep[:, -1] = True
        chunks.append(row[keep])
        lens.append(keep.sum(dim=1))
        del row, keep, words, wkeep
    text = torch.cat(chunks)
    ln = torch.cat(lens)
    offsets = torch.zeros(n_lines + 1, dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(ln, 0)
    return text, offsets


def c3(dev, n_lines):
    text, offsets = log_lines(n_lines, dev)
    nbytes = text.numel()
    res = {}
    for pat in (r"(\d{4})-(\d{2})-(\d{2})", r"(?-u)(\d{4})-(\d{2})-(\d{2})"):
        r = use_torch_stream(R.BytesRegex(pat))
        bits = torch.zeros((n_lines + 31) // 32, dtype=torch.int32, device=dev)
        spans = torch.empty((n_lines, 2), dtype=torch.int64, device=dev)
        ms_is, _ = timed(lambda: r.is_match_batch_device(text, offsets, bits))
        n_match = int(sum(bin(int(x) & (2**32 - 1)).count("1") for x in bits[:8192].tolist()))
        ms_find, _ = timed(lambda: r.find_batch_device(text, offsets, spans, bits))
        # spot check against the oracle on the first 2000 lines
        from oracle import oracle as O
        o = O.OracleRegex([pat])
        host = text[:int(offsets[2000])].cpu().numpy().tobytes()
        offs = offsets[:2001].tolist()
        sp = spans[:2000].tolist()
        bw = bits[:64].tolist()
        bad = 0
        for i in range(2000):
            m = o.find_at(host[offs[i]:offs[i + 1]])
            hit = (int(bw[i // 32]) >> (i % 32)) & 1
            if (m is None) != (not hit) or (m is not None and tuple(m) != tuple(sp[i])):
                bad += 1
        res[pat] = {"is_match_Mlines_s": round(n_lines / ms_is / 1e3, 1), "is_match_GB_s": round(nbytes / ms_is / 1e6, 1),
                    "find_Mlines_s": round(n_lines / ms_find / 1e3, 1), "find_GB_s": round(nbytes / ms_find / 1e6, 1),
                    "matches_in_first_262144_lines": n_match, "oracle_mismatches_first_2000": bad}
    print(json.dumps({"config": f"C3 batched per-line is_match + find over {n_lines} synthetic log lines ({nbytes / 1e9:.2f} GB)", "results": res}), flush=True)


def c4_patterns():
    words = [w.decode() for w in open(os.path.join(ROOT, "tests", "golden", "sherlock.txt"), "rb").read().split() if w.isalpha()]
    from collections import Counter
    ranked = [w for w, _ in Counter(words).most_common(200)]
    lits = ranked[50:90]
    alts = ["|".join(ranked[90 + 3 * i:90 + 3 * i + (2 + i % 2)]) for i in range(10)]
    pats = [r"\w+", r"\d+", r"\s+", r"[A-Z][a-z]+"] + lits + alts + [r"^The", r"\.$", r"(?m)^$", r"[0-9]{4}", r"(?i)holmes",
                                                                                   r"(?m)^Sherlock", r"Mr\.", r"Mrs\.", r"[a-z]+'s", r"(?-u)\bBaker\b"]
    return pats[:64]


def c4(dev, gib):
    n = int(gib * GIB)
    text = B.device_corpus(n, B.SEED, dev)
    pats = c4_patterns()
    s = R.BytesRegexSet(pats)
    ms, mask = timed(lambda: s.matches_mask_device(text))
    res = {"n_patterns": len(pats), "whole_haystack_GB_s": round(n / ms / 1e6, 1), "mask": [hex(m) for m in mask]}
    # per-line masks
    nl = torch.nonzero(text == 10).flatten() + 1
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), nl])
    n_lines = offsets.numel() - 1
    masks = torch.zeros((n_lines, (len(pats) + 63) // 64), dtype=torch.int64, device=dev)
    ms2, _ = timed(lambda: s.matches_batch_device(text, offsets, masks))
    res.update({"per_line_GB_s": round(int(offsets[-1]) / ms2 / 1e6, 1), "per_line_Mlines_s": round(n_lines / ms2 / 1e3, 1), "lines": n_lines})
    print(json.dumps({"config": f"C4 RegexSet x{len(pats)} over {gib:g} GiB of the C2 corpus (bytes::RegexSet)", "results": res}), flush=True)


def c5(dev, gib):
    n = int(gib * GIB)
    g = torch.Generator(device=dev)
    g.manual_seed(0x5EED0005)
    lut = torch.tensor([97, 99, 103, 116], dtype=torch.uint8, device=dev)
    text = torch.empty(n, dtype=torch.uint8, device=dev)
    step = 1 << 28
    for lo in range(0, n, step):
        m = min(step, n - lo)
        text[lo:lo + m] = lut[torch.randint(0, 4, (m,), generator=g, device=dev)]
    res = {}
    for v in DNA_VARIANTS:
        r = use_torch_stream(R.BytesRegex(v))
        cnt = r.find_all_device(text)
        spans = torch.empty((cnt + 16, 2), dtype=torch.int64, device=dev)
        ms, c2 = timed(lambda: r.find_all_device(text, spans), steps=3, warmup=2)
        assert c2 == cnt
        res[v] = {"GB_s": round(n / ms / 1e6, 1), "matches": cnt, "density_per_byte": cnt / n}
    print(json.dumps({"config": f"C5 DNA {gib:g} GiB uniform acgt on one GPU, regex-dna variants, find_iter spans", "results": res}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=20_000_000)
    ap.add_argument("--gib", type=float, default=4.0)
    ap.add_argument("--skip", default="")
    a = ap.parse_args()
    skip = set(a.skip.split(","))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    if "c1" not in skip:
        c1(dev)
    if "c3" not in skip:
        c3(dev, a.lines)
    if "c4" not in skip:
        c4(dev, a.gib)
    if "c5" not in skip:
        c5(dev, a.gib)


if __name__ == "__main__":
    main()
