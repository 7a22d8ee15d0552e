"""In-tree build of the CUDA library (regex_b200/librure_b200.so) for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the build container; the
resulting .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librure_b200.so")

SOURCES = [
    "frontend/parser.cpp",
    "frontend/compiler.cpp",
    "dfa/determinize.cpp",
    "gpu/kernels.cu",
    "gpu/engine.cu",
    "capi/rure_capi.cpp",
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-pthread", "-rdc=false",
]


def _newest_source_mtime():
    newest = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            newest = max(newest, os.path.getmtime(os.path.join(root, f)))
    for f in os.listdir(os.path.join(HERE, "..", "include")):
        newest = max(newest, os.path.getmtime(os.path.join(HERE, "..", "include", f)))
    return newest


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_source_mtime():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace("/", "_") + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("RB200_EXTRA_NVCC_FLAGS", "").split(), "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out.decode())
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-shared", "-o", OUT, *objs, "-lcudart"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
