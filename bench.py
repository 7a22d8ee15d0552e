#!/usr/bin/env python3
"""Headline benchmark: haystack GB/s scanned by find_iter with bit-exact spans.

Contract (task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line from rank 0.  A "step" is one full find_iter pass (all spans written to HBM) over
this rank's resident haystack shard.

Workload at N=1 = BASELINE.json configs[1]: sherlock-style patterns over 16 GiB of
synthetic English text (lines of bench/src/data/sherlock.txt sampled with replacement,
SURVEY.md §8d C2), headline pattern `[a-zA-Z]+ing`.  At N>1 every rank holds its own
16 GiB shard of a 16*N GiB corpus (weak scaling, byte-range sharding); see
regex_b200/sharded.py for the boundary exchange.

`--impl reference` times the CPU path (the oracle's restatement of the reference's lazy
DFA pipeline -- the reference itself is Rust and cannot be built here) on the box's
host cores over a bounded sample of the same corpus.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PATTERN = r"[a-zA-Z]+ing"
ALSO = [r"Holmes|Watson", r"Sherlock|Holmes", r"Sher[a-z]+|Hol[a-z]+", r"(?i)Sherlock|Holmes|Watson", r"the\s+\w+"]
SEED = 0x5EED0001
GIB = 1 << 30
# ncu --set full, scan_rev_fast<1> on 1 GiB of this corpus: 1.762 GB read + 0.223 GB written (profiles/r01_ncu_fused_scan.md)
TRAFFIC_PER_BYTE = (1.762146e9 + 0.22295552e9) / (1 << 30)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gib", type=float, default=16.0, help="haystack GiB per GPU")
    ap.add_argument("--pattern", default=PATTERN)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--seg", type=int, default=0, help="scan segment bytes (0 = automatic)")
    ap.add_argument("--no-fuse", action="store_true")
    ap.add_argument("--no-tensor-tma", action="store_true")
    ap.add_argument("--also", action="store_true", help="time the other C2 patterns once each (extra keys)")
    return ap.parse_args()


# ------------------------------------------------------------------ corpus ----
def sherlock_lines():
    text = open(os.path.join(ROOT, "tests", "golden", "sherlock.txt"), "rb").read()
    lines = text.split(b"\n")[:-1]
    return [l + b"\n" for l in lines]


def host_corpus(n_bytes, seed=SEED):
    """Same distribution as the device corpus, built on the host (CPU baseline sample)."""
    import numpy as np
    lines = sherlock_lines()
    lens = np.array([len(l) for l in lines], dtype=np.int64)
    rng = np.random.Generator(np.random.PCG64(seed))
    parts, total = [], 0
    while total < n_bytes:
        for i in rng.integers(0, len(lines), size=8192):
            parts.append(lines[i])
            total += int(lens[i])
            if total >= n_bytes:
                break
    return b"".join(parts)[:n_bytes]


def device_corpus(n_bytes, seed, device):
    """Lines of sherlock.txt sampled with replacement, materialised directly in HBM."""
    import numpy as np
    import torch
    lines = sherlock_lines()
    lens = torch.tensor([len(l) for l in lines], dtype=torch.int64, device=device)
    starts = torch.cumsum(lens, 0) - lens
    flat = torch.frombuffer(bytearray(b"".join(lines)), dtype=torch.uint8).to(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.empty(n_bytes, dtype=torch.uint8, device=device)
    block = 64 << 20
    mean = float(lens.float().mean())
    pos = 0
    while pos < n_bytes:
        want = min(block, n_bytes - pos)
        k = int(want / mean * 1.05) + 64
        pick = torch.randint(0, len(lines), (k,), generator=gen, device=device)
        l = lens[pick]
        cum = torch.cumsum(l, 0)
        have = int(cum[-1])
        take = min(want, have)
        idx = torch.arange(take, device=device, dtype=torch.int64)
        line = torch.searchsorted(cum, idx, right=True)
        src = starts[pick[line]] + (idx - (cum[line] - l[line]))
        out[pos:pos + take] = flat[src]
        pos += take
        del idx, line, src, cum, l, pick
    return out


# ------------------------------------------------------------------ clocks ----
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(int(float(s[0])) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(s[3 + j].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.samples[0][1])), "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------- reference ----
def cpu_reference(pattern, sample, threads, steps, warmup):
    from oracle import oracle as O
    O.build()
    for _ in range(warmup):
        O.count_parallel(pattern, sample[: len(sample) // 8], threads)
    t0 = time.perf_counter()
    count = 0
    for _ in range(steps):
        count = O.count_parallel(pattern, sample, threads)
    dt = (time.perf_counter() - t0) / steps
    return len(sample) / dt / 1e9, dt, count


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_bytes = 256 << 20
    sample = host_corpus(sample_bytes)
    gbs, dt, count = cpu_reference(args.pattern, sample, threads, max(1, args.steps), min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "haystack GB/s scanned (find_iter, bit-exact spans)", "value": round(gbs, 4),
        "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"C2 sherlock-style find_iter `{args.pattern}` over {args.gib:g} GiB/GPU synthetic English text "
                               f"(reference arm: each step scans a bounded {sample_bytes >> 20} MiB sample of that corpus on the host cores)",
                   "haystack_bytes_per_gpu": int(args.gib * GIB), "sample_bytes": sample_bytes},
        "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": f"{sample_bytes >> 20} MiB of the C2 corpus, {threads} threads, byte ranges cut at newlines, oracle engine 'auto' "
                                   f"(DfaSuffix where the reference selects it); matches={count}"},
        "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------- ours ----
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import regex_b200 as R
    from regex_b200 import sharded

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (regex_b200 has no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = int(args.gib * GIB)          # bytes owned per GPU (multiple of 256)
    total_len = n * world            # one haystack, byte-range sharded (weak scaling)
    halo = 1 << 16
    own = device_corpus(n, SEED + rank, dev)
    geom = sharded.plan(total_len, world, rank, halo=halo)
    assert geom.b - geom.a == n
    if world > 1:
        # replicate 256 B of left context and `halo` bytes of the right neighbour (NCCL p2p)
        left = torch.empty(geom.a - geom.buf_lo, dtype=torch.uint8, device=dev)
        right = torch.empty(geom.buf_hi - geom.b, dtype=torch.uint8, device=dev)
        ops = []
        if rank + 1 < world:
            ops += [dist.P2POp(dist.isend, own[-256:].contiguous(), rank + 1), dist.P2POp(dist.irecv, right, rank + 1)]
        if rank > 0:
            ops += [dist.P2POp(dist.isend, own[:halo].contiguous(), rank - 1), dist.P2POp(dist.irecv, left, rank - 1)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        text = torch.cat([left, own, right])
        del left, right
        comm = sharded.TorchDistComm(dev)
    else:
        text = own
        comm = sharded.SingleComm()
    del own

    re_ = R.BytesRegex(args.pattern)
    re_.set_stream(torch.cuda.current_stream().cuda_stream)  # so torch events bracket the kernels
    if args.seg:
        re_.set_tuning(seg=args.seg)
    if args.no_fuse:
        re_.set_fuse(False)
    if args.no_tensor_tma:
        re_.set_tensor_tma(False)
    info = re_.pattern_info()
    # count pass sizes the span buffer
    probe = sharded.GpuShardEngine(re_, text, cap=0)
    n_local, _, _, _ = sharded.find_all_sharded(probe, geom, comm, info["can_match_empty"], info["has_looks"])
    engine = sharded.GpuShardEngine(re_, text, cap=n_local + 1024)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return sharded.find_all_sharded(engine, geom, comm, info["can_match_empty"], info["has_looks"])

    scan_ms, walk_ms = [], []
    with ClockSampler(local_rank) as clk:  # samples cover the warm-up steps and the timed steps
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        launches0 = R.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            got, offset, grand_total, rounds = step()
            st = re_.last_stats()
            scan_ms.append(st["scan_ms"]); walk_ms.append(st["walk_ms"])
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
    launches = R.kernel_launches() - launches0
    assert got == n_local
    dev_s = e0.elapsed_time(e1) / 1e3
    t = torch.tensor([dev_s, wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall = float(t[0]), float(t[1])
    ms_per_step = dev_s / args.steps * 1e3
    value = total_len / (dev_s / args.steps) / 1e9

    result = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        scan = sum(scan_ms) / len(scan_ms)
        fused = bool(st["fused"])
        # 1 byte read per haystack byte + 16 bytes written per emitted span; the start
        # bitmap, staging and stitch traffic are this design's overhead and earn no credit.  The fused
        # kernel (reverse scan + chain walk of each segment) emits the spans; without fusion the
        # dominant kernel is the scan alone and the spans belong to walk_chunks.
        alg_bytes = n + (16 * n_local if fused else 0)
        achieved = alg_bytes / (scan / 1e3) / 1e9
        result = {
            "metric": "haystack GB/s scanned (find_iter, bit-exact spans)", "value": round(value, 2), "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"C2 sherlock-style find_iter `{args.pattern}` over {args.gib:g} GiB/GPU synthetic English text "
                                   f"({world} byte-range shard(s) of one {args.gib * world:g} GiB haystack)",
                       "haystack_bytes_per_gpu": n, "matches_rank0": n_local, "matches_total": grand_total,
                       "l2": "each shard (16 GiB) far exceeds the 126 MB L2; no flush needed",
                       "timing": "torch CUDA events on the stream the library launches on (set_stream), barrier + synchronize on both sides, "
                                 "max over ranks; includes the boundary all_gathers; wall clock %.3f s" % wall,
                       "boundary_fixup_rounds": rounds},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": TRAFFIC_PER_BYTE * n if fused and args.pattern == PATTERN else None,
                         "traffic_source": "ncu --set full dram__bytes_read+write of scan_rev_fast<1> on a 1 GiB haystack (profiles/r01_ncu_fused_scan.md), scaled per byte",
                         "kernel": "scan_rev_fast<fused scan+walk>" if fused else "scan_rev_fast<scan only>", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": round(scan, 4),
                         "other_kernels_ms": round(sum(walk_ms) / len(walk_ms), 4)},
            "clocks": clk.summary(),
            "fixups": {k: st[k] for k in ("scan_redo_rounds", "scan_redo_segments", "stitch_rounds", "stitch_dirty_chunks")},
        }

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    if not args.no_e2e:
        e2e_n = min(n, 4 * GIB)
        lo = geom.own_lo
        host = torch.empty(e2e_n, dtype=torch.uint8, pin_memory=True)
        host.copy_(text[lo:lo + e2e_n])
        cap = int(n_local * (e2e_n / n) * 1.1) + 4096
        out = np.empty((cap, 2), dtype=np.uint64)
        tot = R.ctypes.c_size_t()
        for _ in range(2):
            R.lib().rure_b200_find_all(re_._h, host.data_ptr(), e2e_n, out.ctypes.data, cap, R.byref(tot))
        barrier()
        k = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(k):
            ok = R.lib().rure_b200_find_all(re_._h, host.data_ptr(), e2e_n, out.ctypes.data, cap, R.byref(tot))
            assert ok
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            result["e2e"] = {"value": round(e2e_n * world / float(t[0]) / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": e2e_n,
                             "d2h_bytes_per_step": int(min(tot.value, cap) * 16 + 8), "haystack_bytes": e2e_n,
                             "note": "rure_b200_find_all on pinned host memory per rank (independent haystacks); H2D of the haystack and "
                                     "D2H of all spans inside the timed region; the library pipelines upload, search and "
                                     "download in 64 MiB pieces"}
        del host

    if rank == 0:
        if args.also:
            also = {}
            spans = torch.empty((n_local * 2 + 4096, 2), dtype=torch.int64, device=dev)
            local = text[geom.own_lo:geom.own_lo + n]
            for pat in ALSO:
                r2 = R.BytesRegex(pat)
                r2.find_all_device(local)
                c = r2.find_all_device(local, spans)
                also[pat] = {"GB/s": round(n / (r2.last_stats()["total_ms"] / 1e3) / 1e9, 1), "matches": c}
            result["also"] = also
        if not args.no_cpu:
            from oracle import oracle as O
            sample_bytes = 128 << 20
            sample = text[geom.own_lo:geom.own_lo + sample_bytes].cpu().numpy().tobytes()
            last_nl = sample.rfind(b"\n") + 1
            threads = os.cpu_count() or 1
            gbs, dt, count = cpu_reference(args.pattern, sample[:last_nl], threads, 1, 1)
            gpu_count = R.BytesRegex(args.pattern).find_all_device(text[geom.own_lo:geom.own_lo + last_nl].contiguous())
            assert gpu_count == count, ("parity spot check failed", gpu_count, count)
            result["cpu_baseline"] = {"value": round(gbs, 4), "unit": "GB/s", "cores": threads, "kind": "port",
                                      "sample": f"first {last_nl >> 20} MiB of rank 0's shard, {threads} threads cut at newlines, oracle engine 'auto' "
                                                f"(DfaSuffix as the reference selects for this pattern, exec.rs:1176-1210), count {count} == GPU count"}
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
