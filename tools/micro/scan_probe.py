"""Scan-kernel probes (scan only, pattern without matches), 4 GiB haystacks:
  - one repeated byte: every lane in the same state reading the same table word (no bank conflicts)
  - the C2 corpus
  - RB200_PROBE_SKIP_TABLE=1: same kernel moving the bytes through the TMA ring without the automaton"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import regex_b200 as R
import bench as B

dev = torch.device("cuda", 0)
n = 4 << 30
def run(name, text, pat):
    r = R.BytesRegex(pat)
    r.set_fuse(False)
    for _ in range(3):
        c = r.find_all_device(text)
    st = r.last_stats()
    print(f"{name:28s} {pat:16s} matches {c:9d} scan {st['scan_ms']:.3f} ms = {n / st['scan_ms'] / 1e6:.0f} GB/s", flush=True)

if os.environ.get("RB200_PROBE_SKIP_TABLE") != "1":
    same = torch.full((n,), 97, dtype=torch.uint8, device=dev)
    run("one repeated byte", same, "zzzzq")
    del same
corpus = B.device_corpus(n, B.SEED, dev)
run("C2 corpus", corpus, "zzzzq")
run("C2 corpus", corpus, "Holmes|Watson")
