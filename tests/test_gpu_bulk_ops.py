"""replace / replace_all / replacen / split / splitn on the GPU (bulk forms of the reference's thin
loops over find_iter, src/re_bytes.rs:316-360, 440-535, 699-749; expansion src/expand.rs:50-167).
Expected values: the reference's own shootout output (examples/regexdna-output.txt), its doc
examples and unit tests, and a Python model of the reference loops over the ORACLE's spans."""
import os

import numpy as np
import pytest

import regex_b200 as R
from helpers import GOLDEN, sherlock_text, tiled_corpus, xorshift_bytes
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _model_replacen(pat, text, limit, rep_fn, utf8=False):
    """re_bytes.rs:476-535 over the oracle's find_iter."""
    spans = O.OracleRegex(pat, only_utf8=utf8).find_iter(text)
    if not spans:
        return text
    out, last = [], 0
    for i, (s, e) in enumerate(spans):
        if limit > 0 and i >= limit:
            break
        out.append(text[last:s])
        out.append(rep_fn(text[s:e]))
        last = e
    out.append(text[last:])
    return b"".join(out)


def _model_split(pat, text, limit=None):
    """re_bytes.rs:699-749 over the oracle's find_iter."""
    spans = iter(O.OracleRegex(pat).find_iter(text))
    state = {"last": 0}

    def split_next():
        m = next(spans, None)
        if m is None:
            if state["last"] >= len(text):
                return None
            s = text[state["last"]:]
            state["last"] = len(text)
            return s
        piece = text[state["last"]:m[0]]
        state["last"] = m[1]
        return piece

    out = []
    if limit is None:
        while True:
            p = split_next()
            if p is None:
                return out
            out.append(p)
    n = limit
    while n > 0:
        n -= 1
        if n == 0:
            out.append(text[state["last"]:])
            break
        p = split_next()
        if p is None:
            break
        out.append(p)
    return out


def test_regexdna_shootout_end_to_end():
    """examples/shootout-regex-dna.rs: clean-up replace_all, nine counts, eleven IUB substitutions;
    lengths and counts from examples/regexdna-output.txt (101745 / 100000 / 133640)."""
    raw = open(os.path.join(GOLDEN, "regexdna-input.txt"), "rb").read()
    assert len(raw) == 101745
    seq = R.Regex(">[^\n]*\n|\n").replace_all(raw, "")
    assert len(seq) == 100000
    variants = ["agggtaaa|tttaccct", "[cgt]gggtaaa|tttaccc[acg]", "a[act]ggtaaa|tttacc[agt]t", "ag[act]gtaaa|tttac[agt]ct",
                "agg[act]taaa|ttta[agt]cct", "aggg[acg]aaa|ttt[cgt]ccct", "agggt[cgt]aa|tt[acg]accct", "agggta[cgt]a|t[acg]taccct",
                "agggtaa[cgt]|[acg]ttaccct"]
    assert [R.Regex(v).count_all(seq) for v in variants] == [0, 3, 9, 8, 10, 3, 4, 3, 5]
    substs = [("B", "(c|g|t)"), ("D", "(a|g|t)"), ("H", "(a|c|t)"), ("K", "(g|t)"), ("M", "(a|c)"), ("N", "(a|c|g|t)"), ("R", "(a|g)"),
              ("S", "(c|g)"), ("V", "(a|c|g)"), ("W", "(a|t)"), ("Y", "(c|t)")]
    model = seq
    for pat, rep in substs:
        seq = R.Regex(pat).replace_all(seq, rep)
        model = model.replace(pat.encode(), rep.encode())
        assert seq == model, pat
    assert len(seq) == 133640


def test_reference_doc_examples_and_expansion():
    # re_bytes.rs doc examples
    assert R.BytesRegex(r"[ \t]+").split(b"a b \t  c\td    e") == [b"a", b"b", b"c", b"d", b"e"]
    assert R.BytesRegex(r"\W+").splitn(b"Hey! How are you?", 3) == [b"Hey", b"How", b"are you?"]
    assert R.BytesRegex(r"[^01]+").replace(b"1078910", b"") == b"1010"
    assert R.BytesRegex(r"\W+").splitn(b"a b", 0) == []
    assert R.BytesRegex(r"\s").splitn(b"a b", 3) == [b"a", b"b", b""]  # SplitN's last piece is the rest, empty or not
    assert R.BytesRegex(r"x").split(b"") == [] and R.BytesRegex(r"x").split(b"axbx") == [b"a", b"b"]
    # expansion (src/expand.rs): $0, ${0}, $$, a lone $, groups that do not exist; NoExpand
    r = R.BytesRegex(r"b+")
    assert r.replace_all(b"abbc abc", b"<$0>") == b"a<bb>c a<b>c"
    assert r.replace_all(b"abbc abc", b"${0}${0}$$") == b"abbbb$c abb$c"
    assert r.replace_all(b"abbc", b"$ $9 ${name}x") == b"a$  xc"
    assert r.replace_all(b"abbc", b"$0", expand=False) == b"a$0c"
    assert r.replacen(b"b b b b", 2, b"X") == b"X X b b"
    assert r.replace_all(b"nothing here", b"X") == b"nothing here"
    assert R.BytesRegex(r"(a)(?P<last>b)").replace_all(b"ab ab", b"$2$1") == b"ba ba"
    assert R.BytesRegex(r"(a)(?P<last>b)").replace_all(b"ab", b"<$last>") == b"<b>"
    assert R.BytesRegex(r"(a)(?P<last>b)").replace_all(b"ab", b"$2$1", expand=False) == b"$2$1"


def test_replace_and_split_against_the_model_over_oracle_spans():
    rng = np.random.Generator(np.random.PCG64(11))
    texts = [sherlock_text()[:300000], xorshift_bytes(4, 50000, b"abc \n"), b"", b"aaaa", tiled_corpus(1 << 20)]
    cases = [(r"[a-zA-Z]+ing", b"<$0>"), (r"Holmes|Watson", b"X"), (r"\s+", b" "), (r"a*", b"-"), (r"", b"|"), (r"(?m)^", b"> "),
             (r"\w+", b"$0$0"), (r"[^\n]*\n", b""), (r"e", b"${0}E$$")]
    for pat, rep in cases:
        r = R.BytesRegex(pat)

        def rep_fn(m, rep=rep):
            return rep.replace(b"${0}", b"\0M").replace(b"$$", b"\0D").replace(b"$0", b"\0M").replace(b"\0M", m).replace(b"\0D", b"$")

        for text in texts:
            for limit in (0, 1, 3):
                assert r.replacen(text, limit, rep) == _model_replacen(pat, text, limit, rep_fn), (pat, limit, len(text))
            assert r.split(text) == _model_split(pat, text), (pat, len(text))
            for limit in (0, 1, 2, 5, 10 ** 9):
                assert r.splitn(text, limit) == _model_split(pat, text, limit), (pat, limit, len(text))


def test_replace_all_device_resident_large():
    import torch
    base = tiled_corpus(16 << 20)
    reps = 16
    d = torch.frombuffer(bytearray(base + b"\n"), dtype=torch.uint8).cuda().repeat(reps)
    r = R.BytesRegex(r"[a-zA-Z]+ing")
    n_out = r.replace_all_device(d, b"[$0]")
    out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    assert r.replace_all_device(d, b"[$0]", out) == n_out
    one = _model_replacen(r"[a-zA-Z]+ing", base + b"\n", 0, lambda m: b"[" + m + b"]")
    assert n_out == len(one) * reps
    got = out.cpu().numpy()
    assert got[:len(one)].tobytes() == one and got[-len(one):].tobytes() == one


def test_replacement_templates_with_capture_groups():
    """`$1`, `${2}`, `$name` (src/expand.rs:50-90) through the capture pass; the reference's doc
    examples (re_bytes.rs:383-440) and a model over the oracle's captures."""
    r = R.BytesRegex(r"(?P<last>[^,\s]+),\s+(?P<first>\S+)")
    assert r.replace(b"Springsteen, Bruce", b"$first $last") == b"Bruce Springsteen"
    assert r.replace(b"Springsteen, Bruce", b"$2 $1") == b"Bruce Springsteen"
    assert r.replace(b"Springsteen, Bruce", b"${first}_$last") == b"Bruce_Springsteen"
    assert r.replace(b"Springsteen, Bruce", b"$first_$last") == b"Springsteen"  # `$first_` is a group that does not exist
    assert r.replace(b"Springsteen, Bruce", b"$2 $last", expand=False) == b"$2 $last"
    pat, rep = r"(\w+)\s+(Holmes|Watson)(,)?", b"[$2|$1$3]"
    text = sherlock_text()[:300000]
    o = O.OracleRegex(pat)

    def expand(c):
        g = [text[x[0]:x[1]] if x is not None else b"" for x in c]
        return b"[" + g[2] + b"|" + g[1] + g[3] + b"]"

    out, last = [], 0
    for c in o.captures_iter(text):
        out.append(text[last:c[0][0]])
        out.append(expand(c))
        last = c[0][1]
    out.append(text[last:])
    assert R.BytesRegex(pat).replace_all(text, rep) == b"".join(out)
