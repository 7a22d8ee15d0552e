"""Pipelined host find_all (pieces searched while the upload is still running) against the
plain upload-then-search path, same library, same inputs; then the e2e rate of both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import regex_b200 as R
import bench as B

def run(r, host, cap, piece):
    os.environ["RB200_PIPELINE_PIECE"] = str(piece)
    out = np.empty((cap, 2), dtype=np.uint64)
    tot = R.ctypes.c_size_t()
    ok = R.lib().rure_b200_find_all(r._h, host.data_ptr(), host.numel(), out.ctypes.data, cap, R.byref(tot))
    assert ok, R._last_error()
    return tot.value, out[:min(tot.value, cap)].copy()

n = 24 << 20
corpus = B.host_corpus(n)
special = bytearray(corpus)
special[5_000_000:5_300_000] = b"a" * 300_000            # one match far longer than the 64 KiB halo
special[9 << 20:(9 << 20) + 40] = b"xxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxx"
bad = 0
for name, data in (("corpus", corpus), ("long run", bytes(special))):
    host = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    for pat in [r"[a-zA-Z]+ing", r"Holmes|Watson", r"\w+", r"a+", r"(?m)^\w+", r"a*", r"(?-u:\b)the(?-u:\b)", r"[^\n]{40,}"]:
        r = R.BytesRegex(pat)
        c0, s0 = run(r, host, 6_000_000, 0)          # plain
        c1, s1 = run(r, host, 6_000_000, 1 << 20)    # 1 MiB pieces
        c2, s2 = run(r, host, 6_000_000, 3 << 20)    # 3 MiB pieces (ragged last piece)
        same = c0 == c1 == c2 and np.array_equal(s0, s1) and np.array_equal(s0, s2)
        bad += not same
        print(f"{name:9s} {pat:24s} matches {c0:8d} pipelined == plain: {same}", flush=True)
print("MISMATCHES:", bad)
n = 4 << 30
host = torch.frombuffer(bytearray(B.host_corpus(n)), dtype=torch.uint8).pin_memory() if False else None
