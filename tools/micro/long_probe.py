"""`(?s)foo.*bar` over 1 GiB: one match as long as the haystack (reverse state maps + parallel long-run pass)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import regex_b200 as R
from helpers import tiled_corpus

base = tiled_corpus(16 << 20)
d = torch.frombuffer(bytearray(base), dtype=torch.uint8).cuda().repeat(64)
r = R.BytesRegex(r"(?s)foo.*bar")
out = torch.empty((16, 2), dtype=torch.int64, device="cuda")
for _ in range(3):
    n = r.find_all_device(d, out)
    print(n, out[:1].tolist(), {k: round(v, 2) for k, v in r.last_stats().items()})
