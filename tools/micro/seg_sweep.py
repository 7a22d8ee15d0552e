"""Segment length x warm-up sweep for the fused find_all path (4 GiB of the C2 corpus by default)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch

import corpus as C
import regex_b200 as R

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4
pat = sys.argv[2] if len(sys.argv) > 2 else r"[a-zA-Z]+ing"
dev = torch.device("cuda", 0)
text = C.device_corpus(int(gib * (1 << 30)), C.SEED, dev)
r0 = R.BytesRegex(pat)
n = r0.find_all_device(text)
out = torch.empty((n + 16, 2), dtype=torch.int64, device=dev)
for seg in (1024, 2048, 3072, 4096):
    for warm in (64, 128):
        r = R.BytesRegex(pat)
        r.set_tuning(seg=seg, warm=warm)
        best = None
        for _ in range(5):
            got = r.find_all_device(text, out)
            st = r.last_stats()
            if best is None or st["total_ms"] < best["total_ms"]:
                best = st
        assert got == n
        print(f"seg={seg} warm={warm} scan_ms={best['scan_ms']:.3f} tail_ms={best['walk_ms']:.3f} total_ms={best['total_ms']:.3f} redo={best['scan_redo_segments']:.0f} path={best['path']:.0f}")
