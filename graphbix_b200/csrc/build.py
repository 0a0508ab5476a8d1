"""Builds graphbix_b200/libgraphbix_cuda.so with nvcc for sm_100a (in-tree, so it travels to the GPU box).

    python graphbix_b200/csrc/build.py [--q 2,3,8] [--force]

One object per number of groups q (gbx_sbm_q.cu, -DGBX_Q=q) compiled in parallel, plus the host/C-ABI object,
linked into one shared library with the static CUDA runtime.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libgraphbix_cuda.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"] + ARCH + os.environ.get("GBX_EXTRA_FLAGS", "").split()
ALL_Q = list(range(1, 17))


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def build(qs=None, force=False, verbose=False):
    qs = sorted(set(qs or ALL_Q))
    os.makedirs(OBJ, exist_ok=True)
    srcs = [os.path.join(HERE, f) for f in ("gbx_sbm_q.cu", "gbx_sbm_host.cu", "gbx_graph_build.cu", "gbx_pool.cu", "gbx_lsbm.cu", "gbx_common.cuh",
                                            "gbx_sbm_internal.h", "gbx_sbm_pull.cuh")]
    srcs.append(os.path.join(os.path.dirname(PKG), "include", "graphbix_cuda.h"))
    stamp = os.path.join(OBJ, "stamp.txt")
    want = _digest(srcs, extra=",".join(map(str, qs)) + " ".join(FLAGS))
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return OUT
    jobs = []
    for q in qs:
        for model, tag in ((0, "sbm"), (1, "mod")):
            o = os.path.join(OBJ, f"gbx_{tag}_q{q}.o")
            jobs.append((o, [NVCC, *FLAGS, f"-DGBX_Q={q}", f"-DGBX_MODEL={model}", "-c",
                             os.path.join(HERE, "gbx_sbm_q.cu"), "-o", o] + (["-Xptxas", "-v"] if verbose else [])))
    foreach = "-DGBX_FOR_EACH_Q(X)=" + " ".join(f"X({q})" for q in qs)
    ho = os.path.join(OBJ, "gbx_sbm_host.o")
    jobs.append((ho, [NVCC, *FLAGS, foreach, "-c", os.path.join(HERE, "gbx_sbm_host.cu"), "-o", ho]))
    go = os.path.join(OBJ, "gbx_graph_build.o")
    jobs.append((go, [NVCC, *FLAGS, "-c", os.path.join(HERE, "gbx_graph_build.cu"), "-o", go]))
    po = os.path.join(OBJ, "gbx_pool.o")
    jobs.append((po, [NVCC, *FLAGS, "-c", os.path.join(HERE, "gbx_pool.cu"), "-o", po]))
    lo = os.path.join(OBJ, "gbx_lsbm.o")
    jobs.append((lo, [NVCC, *FLAGS, "-c", os.path.join(HERE, "gbx_lsbm.cu"), "-o", lo]))
    logs = []
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        for out in ex.map(lambda j: _run(j[1]), jobs):
            logs.append(out)
    _run([NVCC, *ARCH, "-shared", "-o", OUT, *[j[0] for j in jobs]])
    with open(stamp, "w") as f:
        f.write(want)
    if verbose:
        sys.stdout.write("\n".join(logs))
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--q", default="")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    qs = [int(x) for x in a.q.split(",") if x] or None
    print(build(qs, a.force, a.verbose))
