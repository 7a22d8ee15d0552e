"""Headline pattern over a few GiB with the scan and the walk fused / apart, and the scan alone
(count-only call): where the time of scan_rev_fast<1> goes.
    python tools/micro/fuse_probe.py [gib]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch

import corpus as C
import regex_b200 as R

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
text = C.device_corpus(int(gib * (1 << 30)), C.SEED, dev)
pat = sys.argv[2] if len(sys.argv) > 2 else r"[a-zA-Z]+ing"
r = R.BytesRegex(pat)
n = r.find_all_device(text)
out = torch.empty((n + 16, 2), dtype=torch.int64, device=dev)
for fuse in (1, 0):
    r.set_fuse(bool(fuse))
    best = None
    for _ in range(6):
        r.find_all_device(text, out)
        st = r.last_stats()
        if best is None or st["total_ms"] < best["total_ms"]:
            best = st
    gbs = text.numel() / best["total_ms"] / 1e6
    print(f"fuse={fuse} n={n} scan_ms={best['scan_ms']:.3f} walk_ms={best['walk_ms']:.3f} total_ms={best['total_ms']:.3f} {gbs:.0f} GB/s path={best['path']}")
