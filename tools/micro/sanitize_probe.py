"""Small end-to-end run (meant for compute-sanitizer memcheck where that tool is available): exact-size device buffers at sizes
around every boundary the kernels special-case (16 / 64 / 4096 bytes, 34 full segments)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import regex_b200 as R

dev = torch.device("cuda", 0)
text = open(os.path.join(os.path.dirname(R.__file__), "..", "tests", "golden", "sherlock.txt"), "rb").read()
mixed = ("naïve café ünï 日本語 2014-01-02 ".encode() * 4)
sizes = [0, 1, 15, 16, 17, 63, 64, 65, 255, 4095, 4096, 4097, 8191, 34 * 4096 - 1, 34 * 4096, 35 * 4096 + 77, 300000]
pats = [r"[a-zA-Z]+ing", r"Holmes|Watson", r"the\s+\w+", r"\w+", r"(?m)^\w+", r"\d+", r"x*", r"(?i)sher\w*"]
total = 0
hays = []
for n in sizes:
    hay = (text[:n // 2] + mixed + text[n // 2:])[:n]
    d = torch.frombuffer(bytearray(hay), dtype=torch.uint8).to(dev) if n else torch.empty(0, dtype=torch.uint8, device=dev)
    try:
        hay.decode("utf-8")
        is_utf8 = True
    except UnicodeDecodeError:
        is_utf8 = False
    hays.append((d, is_utf8))
for pat in pats:  # compile once per pattern: determinising the Unicode classes takes a moment
    for cls in (R.BytesRegex, R.Regex):
        try:
            r = cls(pat)
        except R.Error:
            continue
        for d, is_utf8 in hays:
            if cls is R.Regex and not is_utf8:
                continue
            c = r.find_all_device(d)
            out = torch.empty((c + 1, 2), dtype=torch.int64, device=dev)
            assert r.find_all_device(d, out) == c
            r.shortest_match_device(d)
            total += c
lines = text[:200000].split(b"\n")
off = np.concatenate([[0], np.cumsum([len(l) for l in lines])]).astype(np.uint64)
body = b"".join(lines)
for pat in [r"(\d{4})-(\d{2})-(\d{2})", r"Holmes", r"\w+ing\b".replace(r"\b", r"(?-u:\b)"), r"^$"]:
    r = R.BytesRegex(pat)
    r.is_match_batch(body, off)
    r.find_batch(body, off)
s = R.BytesRegexSet([r"\w+", "Holmes", r"\d+", r"^The"])
s.matches(body)
s.matches_batch(body, off)
torch.cuda.synchronize()
print("sanitize probe done, matches:", total)
