"""Build graphvqa_b200/lib/libgvqa_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m graphvqa_b200.build [--force] [--verbose]

The library has no PyTorch dependency: plain ``extern "C"`` entry points declared in
``include/gvqa_b200.h``.  Objects are cached under ``graphvqa_b200/csrc/build/`` keyed by a hash of
the source + headers + flags.
"""
import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libgvqa_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the gvqa_b200 CUDA library cannot be built")
    return exe


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(src):
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    for path in [os.path.join(CSRC, src)] + headers + [os.path.join(INCLUDE, "gvqa_b200.h")]:
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    objs, jobs = [], []
    for src in _sources():
        obj = os.path.join(OBJ_DIR, "%s.%s.o" % (src[:-3], _digest(src)))
        objs.append(obj)
        if force or not os.path.exists(obj):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                  ["-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            for src, r in pool.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write("== %s\n%s%s" % (src, r.stdout, r.stderr))
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed on %s" % src)
    stale = not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs)
    tag = os.path.join(OBJ_DIR, "linked.txt")
    want = "\n".join(objs)
    if not stale and os.path.exists(tag) and open(tag).read() != want:
        stale = True
    if jobs or stale or force:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libgvqa_b200.so failed")
        with open(tag, "w") as f:
            f.write(want)
        # drop objects of older source revisions
        for f in os.listdir(OBJ_DIR):
            p = os.path.join(OBJ_DIR, f)
            if f.endswith(".o") and p not in objs:
                os.remove(p)
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
