"""In-tree build of libaread_sm100.so (nvcc, sm_100a only).

    python aread-multi-domain-recommendation_b200/build.py [--force] [--verbose]

The library is built next to this file so that it travels with the repo snapshot to the GPU
box; nothing is JIT-compiled at import time.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(PKG_DIR, "libaread_sm100.so")
STAMP_PATH = os.path.join(PKG_DIR, "build", "stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr", "--expt-extended-lambda",
    "-Xptxas", "-v",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libaread_sm100.so cannot be built")
    return nvcc


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files += sorted(os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE))
    for path in files:
        h.update(path.encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as fh:
        return fh.read().strip() == _digest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Objects are compiled in parallel."""
    if not force and is_current():
        return LIB_PATH
    nvcc = _nvcc()
    obj_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)
    procs, objs = [], []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        log.append(f"==== {os.path.basename(src)}\n{out}")
        failed |= proc.returncode != 0
    with open(os.path.join(obj_dir, "nvcc.log"), "w") as fh:
        fh.write("\n".join(log))
    if failed or verbose:
        print("\n".join(log), file=sys.stderr if failed else sys.stdout)
    if failed:
        raise RuntimeError("nvcc failed, see log above")
    link = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
            "-Xcompiler", "-fPIC"]
    subprocess.run(link, check=True)
    with open(STAMP_PATH, "w") as fh:
        fh.write(_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
