"""Build libdiffute_b200.so in-tree with nvcc for sm_100a (no torch types in the ABI, loaded via ctypes)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdiffute_b200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v" if os.environ.get("DFU_PTXAS_V") else "-O3"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/diffute_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and (f.endswith((".cu", ".cuh", ".h"))):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


TRACE_LIB = os.path.join(HERE, "libdiffute_b200_trace.so")


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    """trace=True builds the -DDFU_TRACE diagnostic variant (in-kernel timeline records, scripts/trace_step.py);
    the product library never contains that code."""
    LIB = TRACE_LIB if trace else globals()["LIB"]
    STAMP = globals()["STAMP"] + (".trace" if trace else "")
    FLAGS = globals()["FLAGS"] + (["-DDFU_TRACE"] if trace else [])
    bdir = os.path.join(HERE, "build", "trace") if trace else os.path.join(HERE, "build")
    dig = _digest() + ("+trace" if trace else "")

    def fresh():
        return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig

    if not force and fresh():
        return LIB
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            # a box without nvcc can only use the library that travelled with the snapshot — say so loudly when the
            # sources it was built from are not the sources in this tree
            sys.stderr.write(f"diffute_b200: WARNING: {os.path.basename(LIB)} does not match the source digest and nvcc "
                             f"is not available to rebuild it; running the STALE binary\n")
            return LIB
        raise RuntimeError(f"nvcc not found and no prebuilt {os.path.basename(LIB)}")
    # one builder at a time (every torchrun rank calls build()): the others wait for the lock, then find a fresh library
    import fcntl
    os.makedirs(bdir, exist_ok=True)
    with open(os.path.join(bdir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():
                return LIB
            return _compile(LIB, STAMP, FLAGS, bdir, dig, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _compile(LIB, STAMP, FLAGS, bdir, dig, verbose):
    objs = []
    procs = []
    os.makedirs(bdir, exist_ok=True)
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    ok = True
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            ok = False
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif verbose and out.strip():
            print(out)
    if not ok:
        raise RuntimeError("nvcc compilation failed")
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, trace="--trace" in sys.argv))
