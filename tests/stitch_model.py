"""CPU model of the chunk-parallel find_iter resolution (TEST HELPER ONLY).

Mirrors csrc/gpu/kernels.cu: every chunk walks its candidates speculatively ("the iterator
enters at my first position"); `resolve_entries` + `stitch_resolve` then bring the chunks into
agreement with the sequential iterator of the reference (src/re_trait.rs:197-220):

  * every chunk's true entry state (p, last_match) is the exit of the last chunk to its left
    that is neither IDENT (no candidate bit at all: the iterator passes through unchanged, so
    it must hand on what it RECEIVED) nor COVERED (the iterator jumps over it) -- an exclusive
    prefix "max by p" over the contributing chunks, computed by a scan on the GPU;
  * a speculative chunk is kept when the true entry lies before it; for patterns with
    look-arounds the START of its first span is re-derived with the reference's
    reverse-on-slice rule from the true entry (src/exec.rs:651-657, SURVEY hazard H1);
  * patterns without empty matches / look-arounds also keep a speculative chunk entered late,
    minus the spans in front of the point where the real and the speculative chain meet;
  * anything else is walked again from the true entry; rounds repeat until nothing changes.
"""
SPEC = "spec"
OK, IDENT, COVERED = 0, 1, 2


def walk(sim, S, text, k, chunk, p, lm, chain, emulate):
    """kernels.cu chunk_walk: the chain over chunk k (positions cb+1..ce, plus 0 for chunk 0)."""
    n = len(text)
    cb, ce = k * chunk, min((k + 1) * chunk, n)
    spans, fc = [], None
    while p is not None:
        lo = max(p, cb + 1)
        if k == 0 and p == 0 and S[0]:
            s = 0
        else:
            s = next((q for q in range(lo, ce + 1) if S[q]), None)
        if s is None:
            break
        if fc is None:
            fc = s
        e = sim.anchored_end(text, s)
        ms = s
        if emulate and chain and e != p:
            ms = sim.slice_start(text, p, e)
            if ms is None:
                p = None
                break
        chain = True
        if ms == e:
            p = sim.next_after_empty(text, e)
            if e == lm:
                continue
        else:
            p = e
        lm = e
        spans.append((ms, e))
    return spans, p, lm, fc


def _key(p):  # None (iteration over) sorts above every position, like kNone on the device
    return float("inf") if p is None else p


def chunked_find_iter(sim, text, chunk, start=0, max_rounds=None, stats=None):
    info = sim.info
    emulate, cme = info["has_looks"], info["can_match_empty"]
    strict = emulate or cme
    n = len(text)
    S = sim.start_bitmap(text)
    nc = max(1, (n + chunk - 1) // chunk)
    st = []
    for k in range(nc):
        if k == 0:
            spans, p, lm, fc = walk(sim, S, text, 0, chunk, start, None, True, emulate)
            st.append(dict(spans=spans, skip=0, cnt=len(spans), out=(p, lm), fc=fc, entry=(start, None), state=OK))
        else:
            spans, p, lm, fc = walk(sim, S, text, k, chunk, k * chunk + 1, None, False, emulate)
            st.append(dict(spans=spans, skip=0, cnt=len(spans), out=(p, lm), fc=fc, entry=SPEC,
                           state=IDENT if fc is None else OK))
    rounds = 0
    while True:
        rounds += 1
        assert rounds <= (max_rounds or 4 * nc + 8), "stitch did not converge"
        # exclusive prefix: entry of chunk k = exit of the last contributing chunk to its left
        T, cur = [None] * nc, st[0]["out"]
        for k in range(1, nc):
            T[k] = cur
            if st[k]["state"] == OK and _key(st[k]["out"][0]) >= _key(cur[0]):
                cur = st[k]["out"]
        changed, dirty = 0, []
        for k in range(1, nc):
            c = st[k]
            if c["state"] == IDENT:
                continue
            tp, tl = T[k]
            cb, ce = k * chunk, min((k + 1) * chunk, n)
            c_first = cb + 1
            new = dict(state=OK, skip=0, cnt=len(c["spans"]))
            rewalk = False
            if tp is None or tp > ce:
                new = dict(state=COVERED, skip=0, cnt=0)
            elif c["entry"] == SPEC:
                if strict:
                    ok = tp < c_first or (tp == c_first and not emulate and not (cme and tl == c_first))
                    if ok and emulate and c["spans"]:
                        s0, e0 = c["spans"][0]
                        ms = sim.slice_start(text, tp, e0)  # e0 >= c_first > tp
                        if ms is None or (ms == e0) != (s0 == e0):
                            ok = False
                        else:
                            c["spans"][0] = (ms, e0)
                    rewalk = not ok
                else:
                    if not (c["fc"] is None or tp <= c["fc"]):
                        live = c["spans"]
                        i = next((j for j, (s, _) in enumerate(live) if s >= tp), len(live))
                        if i == 0 or live[i - 1][1] <= tp:
                            new = dict(state=OK, skip=i, cnt=len(live) - i)
                        else:
                            rewalk = True
            else:
                rewalk = c["entry"] != (tp, tl) if strict else c["entry"][0] != tp
            if rewalk:
                dirty.append((k, tp, tl))
                continue
            if (c["state"], c["skip"], c["cnt"]) != (new["state"], new["skip"], new["cnt"]):
                changed += 1
                c.update(new)
        for k, tp, tl in dirty:
            spans, p, lm, fc = walk(sim, S, text, k, chunk, tp, tl, True, emulate)
            st[k] = dict(spans=spans, skip=0, cnt=len(spans), out=(p, lm), fc=fc, entry=(tp, tl), state=OK)
        if stats is not None:
            stats["rewalks"] = stats.get("rewalks", 0) + len(dirty)
            stats["rounds"] = rounds
        if not dirty and not changed:
            break
    out = []
    for c in st:
        if c["state"] == OK:
            out += c["spans"][c["skip"]:c["skip"] + c["cnt"]]
    # exit state of the whole range (what a shard hands to its right neighbour)
    cur = st[0]["out"]
    for k in range(1, nc):
        if st[k]["state"] == OK and _key(st[k]["out"][0]) >= _key(cur[0]):
            cur = st[k]["out"]
    return out, cur
