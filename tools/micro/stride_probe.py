"""Hot-table row stride experiment: run from a build directory whose library was compiled with -DRB_HOT_ROW=<n>."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import regex_b200 as R
import bench as B
dev = torch.device("cuda", 0)
n = 4 << 30
corpus = B.device_corpus(n, B.SEED, dev)
out = []
for pat in ["[a-zA-Z]+ing", "Holmes|Watson", r"the\s+\w+"]:
    r = R.BytesRegex(pat)
    for _ in range(3):
        c = r.find_all_device(corpus)
    out.append(f"{pat}: {n / r.last_stats()['total_ms'] / 1e6:.0f}")
del corpus
g = torch.Generator(device=dev); g.manual_seed(5)
lut = torch.tensor([97, 99, 103, 116], dtype=torch.uint8, device=dev)
dna = lut[torch.randint(0, 4, (n,), generator=g, device=dev)]
for pat in ["agggtaaa|tttaccct", "[cgt]gggtaaa|tttaccc[acg]", "agggtaa[cgt]|[acg]ttaccct"]:
    r = R.BytesRegex(pat)
    for _ in range(3):
        c = r.find_all_device(dna)
    out.append(f"{pat}: {n / r.last_stats()['total_ms'] / 1e6:.0f}")
print(os.path.basename(os.getcwd()), " | ".join(out), flush=True)
