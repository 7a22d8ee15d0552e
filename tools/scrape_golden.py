#!/usr/bin/env python3
"""Scrape the reference's own known-answer tests into tests/golden/*.json.

Sources (all plain text in the reference tree; nothing is compiled or run):
  tests/*.rs                    mat!/matiter!/ismatch!/matset!/nomatset!/shortmat!
                                (macro definitions: tests/macros.rs:1-149,
                                 tests/shortest_match.rs:1-10)
  bench/src/sherlock.rs:31-196  sherlock!(name, pattern, find_iter count)
  bench/src/data/sherlock.txt   the haystack those counts refer to
  examples/regexdna-{input,output}.txt + examples/shootout-regex-dna.rs:26-34
  tests/noparse.rs              patterns that must fail to compile

mat! cases keep group 0 in "expected" and every listed group in "groups" (the macro compares
as many groups as the test lists, tests/macros.rs:22-52).  Haystacks are stored as hex so
raw-byte cases (tests/bytes.rs R(b"...")) survive JSON.

Usage: python tools/scrape_golden.py [/root/reference]
"""
import json
import re
import shutil
import sys
from pathlib import Path

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
out_dir = Path(__file__).resolve().parent.parent / "tests" / "golden"
out_dir.mkdir(parents=True, exist_ok=True)


# ------------------------------------------------------------ Rust literals --
def decode_escapes(body: str, is_bytes: bool) -> bytes:
    out = bytearray()
    i = 0
    while i < len(body):
        c = body[i]
        if c != "\\":
            out += c.encode("utf-8")
            i += 1
            continue
        i += 1
        e = body[i]
        i += 1
        if e == "n":
            out.append(10)
        elif e == "r":
            out.append(13)
        elif e == "t":
            out.append(9)
        elif e == "0":
            out.append(0)
        elif e in "\\'\"":
            out += e.encode()
        elif e == "x":
            v = int(body[i:i + 2], 16)
            i += 2
            out += bytes([v]) if is_bytes else chr(v).encode("utf-8")
        elif e == "u":
            j = body.index("}", i)
            out += chr(int(body[i + 1:j], 16)).encode("utf-8")
            i = j + 1
        elif e == "\n":
            while i < len(body) and body[i] in " \t\n\r":
                i += 1
        else:
            raise ValueError(f"unknown escape \\{e}")
    return bytes(out)


class Lit:
    def __init__(self, value: bytes):
        self.value = value


def read_literal(src: str, i: int):
    """Parse a Rust string literal starting at src[i]; return (bytes, next index)."""
    is_bytes = False
    if src[i] == "b":
        is_bytes = True
        i += 1
    if src[i] == "r":
        j = i + 1
        hashes = 0
        while src[j] == "#":
            hashes += 1
            j += 1
        assert src[j] == '"'
        end = src.index('"' + "#" * hashes, j + 1)
        return src[j + 1:end].encode("utf-8"), end + 1 + hashes
    assert src[i] == '"', src[i:i + 20]
    j = i + 1
    while src[j] != '"':
        j += 2 if src[j] == "\\" else 1
    return decode_escapes(src[i + 1:j], is_bytes), j + 1


LIT_START = re.compile(r'b?r#*"|b?"')


def split_args(src: str, i: int):
    """src[i] is '(' of a macro call.  Returns (list of raw arg strings with
    literals replaced by Lit objects in a side list, end index)."""
    depth = 0
    args, cur, lits = [], [], []
    while True:
        c = src[i]
        m = LIT_START.match(src, i)
        if m and (c in 'b"' or c == "r") and not (cur and (cur[-1].isalnum() or cur[-1] == "_")):
            val, i = read_literal(src, i)
            lits.append(val)
            cur.append(f"\x00{len(lits) - 1}\x00")
            continue
        if c == "/" and src[i + 1] == "/":
            i = src.index("\n", i)
            continue
        if c in "([{":
            depth += 1
            if depth > 1:
                cur.append(c)
        elif c in ")]}":
            depth -= 1
            if depth == 0:
                args.append("".join(cur).strip())
                return args, lits, i + 1
            cur.append(c)
        elif c == "," and depth == 1:
            args.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
        i += 1


def lit_of(arg: str, lits):
    m = re.search(r"\x00(\d+)\x00", arg)
    assert m, arg
    return lits[int(m.group(1))]


MACRO = re.compile(r"^\s*(mat|matiter|ismatch|matset|nomatset|shortmat)!\(", re.M)
FILES = {
    # file -> modes it is compiled under (tests/test_default.rs:53-69, test_default_bytes.rs:40-54)
    "api.rs": ["str", "bytes"], "api_str.rs": ["str"], "bytes.rs": ["bytes"], "crazy.rs": ["str", "bytes"],
    "flags.rs": ["str", "bytes"], "fowler.rs": ["str", "bytes"], "misc.rs": ["str"], "multiline.rs": ["str", "bytes"],
    "regression.rs": ["str", "bytes"], "set.rs": ["str", "bytes"], "shortest_match.rs": ["str", "bytes"],
    "suffix_reverse.rs": ["str", "bytes"], "unicode.rs": ["str", "bytes"], "word_boundary.rs": ["str", "bytes"],
    "word_boundary_ascii.rs": ["bytes"], "word_boundary_unicode.rs": ["str"],
}

vectors = []
for fname, modes in FILES.items():
    src = (ref / "tests" / fname).read_text()
    for m in MACRO.finditer(src):
        kind = m.group(1)
        args, lits, _ = split_args(src, m.end() - 1)
        if args and args[-1] == "":
            args.pop()
        name = args[0]
        v = {"file": fname, "kind": kind, "name": name, "modes": modes}
        if kind in ("matset", "nomatset"):
            v["res"] = [lits[int(k)].decode("utf-8") for k in re.findall(r"\x00(\d+)\x00", args[1])]
            v["text_hex"] = lit_of(args[2], lits).hex()
            v["expected"] = [int(a) for a in args[3:]] if kind == "matset" else []
        else:
            v["re"] = lit_of(args[1], lits).decode("utf-8")
            v["text_hex"] = lit_of(args[2], lits).hex()
            rest = ",".join(args[3:])
            if kind == "mat":
                first = args[3]
                if first.startswith("None"):
                    v["expected"] = None
                else:
                    a, b = re.match(r"Some\(\((\d+),\s*(\d+)\)\)", ",".join(args[3:5])).groups()
                    v["expected"] = [int(a), int(b)]
                v["groups"] = [[int(g.group(1)), int(g.group(2))] if g.group(1) is not None else None
                               for g in re.finditer(r"Some\(\((\d+),\s*(\d+)\)\)|None", rest)]
            elif kind == "matiter":
                v["expected"] = [[int(a), int(b)] for a, b in re.findall(r"\((\d+),\s*(\d+)\)", rest)]
            elif kind == "ismatch":
                v["expected"] = args[3] == "true"
            elif kind == "shortmat":
                mm = re.match(r"Some\((\d+)\)", args[3])
                v["expected"] = int(mm.group(1)) if mm else None
        vectors.append(v)

(out_dir / "reference_tests.json").write_text(json.dumps(vectors, indent=0, ensure_ascii=True) + "\n")
print(f"reference_tests.json: {len(vectors)} vectors")

# ------------------------------------------------------------------ noparse --
src = (ref / "tests" / "noparse.rs").read_text()
noparse = []
for m in re.finditer(r"^\s*noparse!\(", src, re.M):
    args, lits, _ = split_args(src, m.end() - 1)
    noparse.append({"name": args[0], "re": lit_of(args[1], lits).decode("utf-8")})
(out_dir / "noparse.json").write_text(json.dumps(noparse, indent=0) + "\n")
print(f"noparse.json: {len(noparse)} patterns")

# ----------------------------------------------------------------- sherlock --
src = (ref / "bench/src/sherlock.rs").read_text()
sher = []
for m in re.finditer(r"^(#\[cfg\([^\n]*\)\]\n)?\s*sherlock!\(", src, re.M):
    cfg = m.group(1) or ""
    args, lits, _ = split_args(src, m.end() - 1)
    if args and args[-1] == "":
        args.pop()
    sher.append({"name": args[0], "re": lit_of(args[1], lits).decode("utf-8"), "count": int(args[2]),
                 "cfg": cfg.strip()})
(out_dir / "sherlock_counts.json").write_text(json.dumps(sher, indent=0, ensure_ascii=True) + "\n")
shutil.copyfile(ref / "bench/src/data/sherlock.txt", out_dir / "sherlock.txt")
print(f"sherlock_counts.json: {len(sher)} benches; sherlock.txt copied (test corpus, public-domain text)")

# ---------------------------------------------------------------- regex-dna --
shutil.copyfile(ref / "examples/regexdna-input.txt", out_dir / "regexdna-input.txt")
shutil.copyfile(ref / "examples/regexdna-output.txt", out_dir / "regexdna-output.txt")
print("regexdna-input/output copied")

# ------------------------------------------------- bench/src/misc.rs is_match --
# bench_match! / bench_not_match! (bench/src/bench.rs:131-152) assert the is_match outcome
# before timing, so every entry is a known answer.  Patterns and haystack recipes are
# transcribed from bench/src/misc.rs:24-195 (haystacks as part lists: literal / repeat / file).
R = lambda s, n: {"rep": [s, n]}
S = lambda s: {"s": s}
F = lambda name: {"file": name}
ABC = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"
misc = [
    ("no_exponential", "a?" * 100 + "a" * 100, True, [R("a", 100)], "misc.rs:24-29"),
    ("literal", r"y", True, [R("x", 50), S("y")], "misc.rs:31-33"),
    ("not_literal", r".y", True, [R("x", 50), S("y")], "misc.rs:35-37"),
    ("match_class", "[abcdw]", True, [R("xxxx", 20), S("w")], "misc.rs:39-41"),
    ("match_class_in_range", "[ac]", True, [R("bbbb", 20), S("c")], "misc.rs:43-45"),
    ("match_class_unicode", r"\p{L}", True, [R("☃5☃5", 20), S("a")], "misc.rs:49-51"),
    ("anchored_literal_short_non_match", r"^zbc(d|e)", False, [S("abcdefghijklmnopqrstuvwxyz")], "misc.rs:53-55"),
    ("anchored_literal_long_non_match", r"^zbc(d|e)", False, [R("abcdefghijklmnopqrstuvwxyz", 15)], "misc.rs:57-59"),
    ("anchored_literal_short_match", r"^.bc(d|e)", True, [S("abcdefghijklmnopqrstuvwxyz")], "misc.rs:61-63"),
    ("anchored_literal_long_match", r"^.bc(d|e)", True, [R("abcdefghijklmnopqrstuvwxyz", 15)], "misc.rs:65-67"),
    ("one_pass_short", r"^.bc(d|e)*$", True, [S("abcddddddeeeededd")], "misc.rs:69-71"),
    ("one_pass_short_not", r".bc(d|e)*$", True, [S("abcddddddeeeededd")], "misc.rs:73-75"),
    ("one_pass_long_prefix", r"^abcdefghijklmnopqrstuvwxyz.*$", True, [S("abcdefghijklmnopqrstuvwxyz")], "misc.rs:77-79"),
    ("one_pass_long_prefix_not", r"^.bcdefghijklmnopqrstuvwxyz.*$", True, [S("abcdefghijklmnopqrstuvwxyz")], "misc.rs:81-83"),
    ("long_needle1", "a" * 30 + "b", True, [R("a", 100000), S("b")], "misc.rs:85-87"),
    ("long_needle2", "b" * 30 + "a", True, [R("b", 100000), S("a")], "misc.rs:89-91"),
    ("reverse_suffix_no_quadratic", r"[r-z].*bcdefghijklmnopq", False, [R("bcdefghijklmnopq", 500)], "misc.rs:97-99"),
]
for size in ("32", "1K", "32K"):
    misc += [
        (f"easy0_{size}", ABC + "$", True, [F(f"{size}.txt"), S(ABC)], "misc.rs:124-129"),
        (f"easy1_{size}", r"A[AB]B[BC]C[CD]D[DE]E[EF]F[FG]G[GH]H[HI]I[IJ]J$", True, [F(f"{size}.txt"), S("AABCCCDEEEFGGHHHIJJ")], "misc.rs:135-142"),
        (f"medium_{size}", r"[XYZ]" + ABC + "$", True, [F(f"{size}.txt"), S("X" + ABC)], "misc.rs:148-153"),
        (f"hard_{size}", r"[ -~]*" + ABC + "$", True, [F(f"{size}.txt"), S(ABC)], "misc.rs:159-164"),
        (f"reallyhard_{size}", r"[ -~]*" + ABC + ".*", True, [F(f"{size}.txt"), S(ABC)], "misc.rs:170-184"),
    ]
misc.append(("reallyhard2_1K", r"\w+\s+Holmes", True, [F("1K.txt"), S("Sherlock Holmes")], "misc.rs:191-194"))
(out_dir / "misc_is_match.json").write_text(json.dumps(
    [{"name": n, "re": p, "is_match": m, "haystack": h, "src": src_} for n, p, m, h, src_ in misc], indent=0, ensure_ascii=True) + "\n")
for size in ("32", "1K", "32K"):
    shutil.copyfile(ref / f"bench/src/data/{size}.txt", out_dir / f"{size}.txt")
print(f"misc_is_match.json: {len(misc)} is_match known answers; 32/1K/32K corpora copied")
