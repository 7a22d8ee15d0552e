"""One pass over each secondary workload at a size that suits `ncu --set full` (profiles/):
    python tools/micro/profile_targets.py c3|c4|fused|launches
c3: batch_fast<0|1> on 5 M log lines; c4: the 64-pattern RegexSet over 1 GiB (product automaton in
the first wave, narrowed automaton on scan_fwd_fast afterwards); fused: `[a-zA-Z]+ing` over 1 GiB."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch

import corpus as C
import regex_b200 as R

what = sys.argv[1] if len(sys.argv) > 1 else "fused"
dev = torch.device("cuda", 0)
if what == "c3":
    text, offsets = C.log_lines(5_000_000, dev)
    r = R.BytesRegex(r"(\d{4})-(\d{2})-(\d{2})")
    bits = torch.zeros((5_000_000 + 31) // 32, dtype=torch.int32, device=dev)
    spans = torch.empty((5_000_000, 2), dtype=torch.int64, device=dev)
    for _ in range(2):
        r.is_match_batch_device(text, offsets, bits)
        r.find_batch_device(text, offsets, spans, bits)
elif what == "c4":
    text = C.device_corpus(1 << 30, C.SEED, dev)
    s = R.BytesRegexSet(C.c4_patterns())
    for _ in range(2):
        print([hex(m) for m in s.matches_mask_device(text)], s.last_stats()["waves"])
else:
    text = C.device_corpus(1 << 30, C.SEED, dev)
    r = R.BytesRegex(r"[a-zA-Z]+ing")
    n = r.find_all_device(text)
    out = torch.empty((n + 16, 2), dtype=torch.int64, device=dev)
    for _ in range(3):
        r.find_all_device(text, out)
    print(n, r.last_stats())
