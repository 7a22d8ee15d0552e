"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def vectors():
    with open(os.path.join(GOLDEN, "reference_tests.json")) as f:
        return json.load(f)


def sherlock_text():
    with open(os.path.join(GOLDEN, "sherlock.txt"), "rb") as f:
        return f.read()


def sherlock_counts():
    with open(os.path.join(GOLDEN, "sherlock_counts.json")) as f:
        # the two cfg(feature = "re-re2"/"re-dphobos") variants are other engines' counts
        return [x for x in json.load(f) if "feature = \"re-" not in x["cfg"] or "not(" in x["cfg"]]


def misc_is_match():
    """bench/src/misc.rs known answers: (name, pattern, expected, haystack bytes)."""
    with open(os.path.join(GOLDEN, "misc_is_match.json")) as f:
        cases = json.load(f)
    out = []
    for c in cases:
        parts = []
        for part in c["haystack"]:
            if "s" in part:
                parts.append(part["s"].encode("utf-8"))
            elif "rep" in part:
                parts.append(part["rep"][0].encode("utf-8") * part["rep"][1])
            else:
                with open(os.path.join(GOLDEN, part["file"]), "rb") as g:
                    parts.append(g.read())
        out.append((c["name"], c["re"], c["is_match"], b"".join(parts)))
    return out


def xorshift_bytes(seed, n, alphabet):
    """Deterministic pseudo-random bytes over `alphabet` (xorshift64*)."""
    out = np.empty(n, dtype=np.uint8)
    x = np.uint64(seed)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    # vectorised: generate 64-bit words with a simple LCG stream seeded by xorshift
    rng = np.random.Generator(np.random.PCG64(int(seed)))
    idx = rng.integers(0, len(alpha), size=n, dtype=np.int64)
    out[:] = alpha[idx]
    return out.tobytes()


def tiled_corpus(n_bytes, seed=0x5EED0001):
    """Lines of sherlock.txt sampled with replacement until n_bytes."""
    text = sherlock_text()
    lines = text.split(b"\n")
    lines = [l + b"\n" for l in lines[:-1]]
    lens = np.array([len(l) for l in lines])
    rng = np.random.Generator(np.random.PCG64(seed))
    parts, total = [], 0
    while total < n_bytes:
        k = rng.integers(0, len(lines), size=4096)
        for i in k:
            parts.append(lines[i])
            total += lens[i]
            if total >= n_bytes:
                break
    return b"".join(parts)[:n_bytes]
