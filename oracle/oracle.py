"""ctypes binding for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module; the product package never does.
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, byref, c_char_p, c_int, c_size_t, c_uint32, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FLAG_CASEI, FLAG_MULTI, FLAG_DOTNL, FLAG_SWAP_GREED, FLAG_SPACE, FLAG_UNICODE = 1, 2, 4, 8, 16, 32
ENGINE_DFA, ENGINE_PIKEVM, ENGINE_AUTO = 0, 1, 2  # AUTO: DfaSuffix where the reference would select it, else DFA


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.oracle_compile.restype = c_void_p
        L.oracle_compile.argtypes = [POINTER(c_char_p), POINTER(c_size_t), c_size_t, c_uint32, c_int, c_size_t]
        L.oracle_free.argtypes = [c_void_p]
        L.oracle_error.restype = c_char_p
        L.oracle_error.argtypes = [c_void_p]
        L.oracle_needs_unicode_word_boundary.argtypes = [c_void_p]
        L.oracle_find_at.argtypes = [c_void_p, c_int, c_char_p, c_size_t, c_size_t, POINTER(c_size_t), POINTER(c_size_t)]
        L.oracle_shortest_match_at.argtypes = [c_void_p, c_int, c_char_p, c_size_t, c_size_t, POINTER(c_size_t)]
        L.oracle_is_match_at.argtypes = [c_void_p, c_int, c_char_p, c_size_t, c_size_t]
        L.oracle_find_iter.restype = c_size_t
        L.oracle_find_iter.argtypes = [c_void_p, c_int, c_char_p, c_size_t, POINTER(c_size_t), c_size_t]
        L.oracle_set_matches.argtypes = [c_void_p, c_int, c_char_p, c_size_t, c_size_t, POINTER(c_uint8)]
        L.oracle_captures_at.restype = c_size_t
        L.oracle_captures_at.argtypes = [c_void_p, c_char_p, c_size_t, c_size_t, POINTER(c_size_t), c_size_t]
        L.oracle_num_patterns.restype = c_size_t
        L.oracle_num_patterns.argtypes = [c_void_p]
        L.oracle_count_parallel.restype = c_size_t
        L.oracle_count_parallel.argtypes = [c_char_p, c_size_t, c_uint32, c_int, c_char_p, c_size_t, c_int]
        _LIB = L
    return _LIB


class OracleError(Exception):
    pass


class OracleRegex:
    """One pattern or a set.  only_utf8=True mirrors `Regex` (str API), False mirrors
    `bytes::Regex` (re_builder.rs:171-172)."""

    def __init__(self, patterns, flags=FLAG_UNICODE, only_utf8=False, size_limit=10 << 20):
        if isinstance(patterns, (str, bytes)):
            patterns = [patterns]
        pats = [p.encode("utf-8") if isinstance(p, str) else p for p in patterns]
        arr = (c_char_p * max(1, len(pats)))(*pats)
        lens = (c_size_t * max(1, len(pats)))(*[len(p) for p in pats])
        self._h = lib().oracle_compile(arr, lens, len(pats), flags, int(only_utf8), size_limit)
        err = lib().oracle_error(self._h)
        if err is not None:
            msg = err.decode("utf-8", "replace")
            lib().oracle_free(self._h)
            self._h = None
            raise OracleError(msg)
        self.n = len(pats)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_free(self._h)
            self._h = None

    @property
    def needs_unicode_word_boundary(self):
        return bool(lib().oracle_needs_unicode_word_boundary(self._h))

    def find_at(self, text, start=0, engine=ENGINE_DFA):
        s, e = c_size_t(), c_size_t()
        ok = lib().oracle_find_at(self._h, engine, text, len(text), start, byref(s), byref(e))
        return (s.value, e.value) if ok else None

    def shortest_match_at(self, text, start=0, engine=ENGINE_DFA):
        e = c_size_t()
        ok = lib().oracle_shortest_match_at(self._h, engine, text, len(text), start, byref(e))
        return e.value if ok else None

    def is_match_at(self, text, start=0, engine=ENGINE_DFA):
        return bool(lib().oracle_is_match_at(self._h, engine, text, len(text), start))

    def captures_at(self, text, start=0):
        """`Regex::captures` (group spans, None for groups that did not take part), or None."""
        buf = (c_size_t * 512)()
        g = lib().oracle_captures_at(self._h, text, len(text), start, buf, 512)
        if g == 0:
            return None
        none = (1 << 64) - 1
        return [None if buf[2 * i] == none else (buf[2 * i], buf[2 * i + 1]) for i in range(g)]

    def captures_iter(self, text, only_utf8=False):
        """`Regex::captures_iter` (re_trait.rs:236-263): the find_iter rule over read_captures_at."""
        out, last_end, last_match, n = [], 0, None, len(text)
        while last_end <= n:
            c = self.captures_at(text, last_end)
            if c is None:
                break
            s, e = c[0]
            if s == e:
                if only_utf8 and e < n:
                    b = text[e]
                    last_end = e + (1 if b <= 0x7F else 2 if b <= 0xDF else 3 if b <= 0xEF else 4)
                else:
                    last_end = e + 1
                if e == last_match:
                    continue
            else:
                last_end = e
            last_match = e
            out.append(c)
        return out

    def find_iter(self, text, engine=ENGINE_DFA, cap=None):
        n = lib().oracle_find_iter(self._h, engine, text, len(text), None, 0) if cap is None else cap
        buf = (c_size_t * (2 * max(1, n)))()
        total = lib().oracle_find_iter(self._h, engine, text, len(text), buf, n)
        k = min(n, total)
        return [(buf[2 * i], buf[2 * i + 1]) for i in range(k)]

    def count(self, text, engine=ENGINE_DFA):
        return lib().oracle_find_iter(self._h, engine, text, len(text), None, 0)

    def set_matches(self, text, start=0, engine=ENGINE_DFA):
        out = (c_uint8 * max(1, self.n))()
        lib().oracle_set_matches(self._h, engine, text, len(text), start, out)
        return [i for i in range(self.n) if out[i]]


def count_parallel(pattern, text, threads, flags=FLAG_UNICODE, only_utf8=False):
    p = pattern.encode("utf-8") if isinstance(pattern, str) else pattern
    return lib().oracle_count_parallel(p, len(p), flags, int(only_utf8), text, len(text), threads)
