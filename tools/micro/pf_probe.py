"""Times one find_all_device pass of a pattern over GIB GiB of the C2 corpus (prefilter on / off).
Usage: python tools/micro/pf_probe.py [GIB] [PATTERN]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import regex_b200 as R
import bench
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
pat = sys.argv[2] if len(sys.argv) > 2 else r"Holmes|Watson"
n = int(gib * (1 << 30))
d = bench.device_corpus(n, bench.SEED, torch.device("cuda", 0))
for pf in (2, 0):
    r = R.BytesRegex(pat)
    r.set_option("prefilter", pf)
    total = r.find_all_device(d)
    out = torch.empty((total + 16, 2), dtype=torch.int64, device="cuda")
    for _ in range(3):
        r.find_all_device(d, out)
    st = r.last_stats()
    print(f"prefilter={pf} path={st['path']:.0f} matches={total} total_ms={st['total_ms']:.3f} scan_ms={st['scan_ms']:.3f} walk_ms={st['walk_ms']:.3f} GB/s={n / st['total_ms'] / 1e6:.1f}", flush=True)
