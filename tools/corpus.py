"""Synthetic inputs of the BASELINE.json configs, shared by bench.py,
tools/bench_configs.py and the GPU tests.  Everything is generated from the committed
tests/golden/sherlock.txt and fixed seeds; nothing here reads /root/reference."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0x5EED0001
GIB = 1 << 30
DNA_VARIANTS = ["agggtaaa|tttaccct", "[cgt]gggtaaa|tttaccc[acg]", "a[act]ggtaaa|tttacc[agt]t", "ag[act]gtaaa|tttac[agt]ct",
                "agg[act]taaa|ttta[agt]cct", "aggg[acg]aaa|ttt[cgt]ccct", "agggt[cgt]aa|tt[acg]accct", "agggta[cgt]a|t[acg]taccct",
                "agggtaa[cgt]|[acg]ttaccct"]


def sherlock_lines():
    text = open(os.path.join(ROOT, "tests", "golden", "sherlock.txt"), "rb").read()
    lines = text.split(b"\n")[:-1]
    return [l + b"\n" for l in lines]


def host_corpus(n_bytes, seed=SEED):
    """Same distribution as the device corpus, built on the host (CPU baseline sample)."""
    import numpy as np
    lines = sherlock_lines()
    lens = np.array([len(l) for l in lines], dtype=np.int64)
    rng = np.random.Generator(np.random.PCG64(seed))
    parts, total = [], 0
    while total < n_bytes:
        for i in rng.integers(0, len(lines), size=8192):
            parts.append(lines[i])
            total += int(lens[i])
            if total >= n_bytes:
                break
    return b"".join(parts)[:n_bytes]


def device_corpus(n_bytes, seed, device):
    """Lines of sherlock.txt sampled with replacement, materialised directly in HBM."""
    import numpy as np
    import torch
    lines = sherlock_lines()
    lens = torch.tensor([len(l) for l in lines], dtype=torch.int64, device=device)
    starts = torch.cumsum(lens, 0) - lens
    flat = torch.frombuffer(bytearray(b"".join(lines)), dtype=torch.uint8).to(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.empty(n_bytes, dtype=torch.uint8, device=device)
    block = 64 << 20
    mean = float(lens.float().mean())
    pos = 0
    while pos < n_bytes:
        want = min(block, n_bytes - pos)
        k = int(want / mean * 1.05) + 64
        pick = torch.randint(0, len(lines), (k,), generator=gen, device=device)
        l = lens[pick]
        cum = torch.cumsum(l, 0)
        have = int(cum[-1])
        take = min(want, have)
        idx = torch.arange(take, device=device, dtype=torch.int64)
        line = torch.searchsorted(cum, idx, right=True)
        src = starts[pick[line]] + (idx - (cum[line] - l[line]))
        out[pos:pos + take] = flat[src]
        pos += take
        del idx, line, src, cum, l, pick
    return out


def log_lines(n_lines, dev, seed=0x5EED0003):
    import torch
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


def c4_patterns():
    words = [w.decode() for w in open(os.path.join(ROOT, "tests", "golden", "sherlock.txt"), "rb").read().split() if w.isalpha()]
    from collections import Counter
    ranked = [w for w, _ in Counter(words).most_common(200)]
    lits = ranked[50:90]
    alts = ["|".join(ranked[90 + 3 * i:90 + 3 * i + (2 + i % 2)]) for i in range(10)]
    pats = [r"\w+", r"\d+", r"\s+", r"[A-Z][a-z]+"] + lits + alts + [r"^The", r"\.$", r"(?m)^$", r"[0-9]{4}", r"(?i)holmes",
                                                                                   r"(?m)^Sherlock", r"Mr\.", r"Mrs\.", r"[a-z]+'s", r"(?-u)\bBaker\b"]
    return pats[:64]


def device_dna(n_bytes, dev, seed=0x5EED0005):
    """C5: uniform i.i.d. lower-case acgt, generated in place on the device."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lut = torch.tensor([97, 99, 103, 116], dtype=torch.uint8, device=dev)
    text = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    step = 1 << 28
    for lo in range(0, n_bytes, step):
        m = min(step, n_bytes - lo)
        text[lo:lo + m] = lut[torch.randint(0, 4, (m,), generator=g, device=dev)]
    return text
