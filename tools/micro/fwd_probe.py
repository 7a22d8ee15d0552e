"""Whole-haystack forward scans (is_match / shortest_match) on 4 GiB of the C2 corpus."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import regex_b200 as R
import bench as B
dev = torch.device("cuda", 0)
n = 4 << 30
corpus = B.device_corpus(n, B.SEED, dev)
for pat in ["zzzzq", "[a-zA-Z]+ing", "Holmes|Watson", r"the\s+\w+"]:
    r = R.BytesRegex(pat)
    for _ in range(2):
        got = r.shortest_match_device(corpus)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        got = r.shortest_match_device(corpus)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"shortest_match {pat:16s} -> {got}  {dt * 1e3:.3f} ms = {n / dt / 1e9:.0f} GB/s", flush=True)
