"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every
symbol the public headers declare, and reports compile errors like the reference."""
import ctypes
import json
import os
import re

import pytest

import regex_b200 as R
from helpers import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rure_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "regex_b200", "librure_b200.so"))
    names = _declared("rure.h") + _declared("rure_b200.h")
    assert len(names) > 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_noparse_patterns_fail_to_compile():
    """tests/noparse.rs: every pattern must be rejected."""
    for x in json.load(open(os.path.join(GOLDEN, "noparse.json"))):
        with pytest.raises(R.Error):
            R.Regex(x["re"])


def test_error_messages_match_reference_text():
    # regex-capi/ctest/test.c:306 and :335 test these substrings
    with pytest.raises(R.Error, match="Unclosed parenthesis"):
        R.BytesRegex("(")
    with pytest.raises(R.Error, match="exceeds size"):
        R.BytesRegex(r"\w{100}", size_limit=0)
    with pytest.raises(R.Error, match="exceeds size"):
        R.BytesRegexSet([r"\w{100}"], size_limit=0)


def test_explicit_errors_of_this_backend():
    with pytest.raises(R.Error, match="word boundar"):
        R.Regex(r"\bfoo\b")
    R.Regex(r"(?-u:\b)foo(?-u:\b)")  # ASCII word boundaries are fine
    with pytest.raises(R.Error, match="exceeds size limit"):
        R.BytesRegex(r"[ab]*a[ab]{14}", dfa_size_limit=1 << 10)  # 2^15 states > tiny budget
    # str-mode regexes refuse byte-oriented patterns (tests/test_default.rs:73-76)
    for pat in [r"(?-u)\xFF", r"(?-u).", r"(?-u)[\xFF]", "(?-u)\u2603"]:
        with pytest.raises(R.Error):
            R.Regex(pat)
    R.BytesRegex(r"(?-u)\xFF")


def test_pattern_info_and_tables():
    r = R.BytesRegex("agggtaaa|tttaccct")
    assert r.pattern_info() == {"min_len": 8, "max_len": 8, "can_match_empty": False, "has_looks": False}
    d = r.dfa(R.DFA_REV_UNANCHORED_ALL)
    assert d["trans"].shape[1] == d["classes"].max() + 2  # byte classes + EOF
    assert (d["trans"][0] == 0).all()  # state 0 is dead
    assert R.Regex(r"a*").pattern_info()["can_match_empty"]
    assert R.Regex(r"(?m)^a$").pattern_info()["has_looks"]


def test_empty_set_has_no_automaton_but_does_not_crash():
    """An empty RegexSet compiles (re_set.rs:96-104) and matches nothing; asking for its
    table is an error, not a crash."""
    import pytest
    import regex_b200 as R
    s = R.BytesRegexSet([])
    assert len(s) == 0
    with pytest.raises(R.Error):
        s.dfa()
