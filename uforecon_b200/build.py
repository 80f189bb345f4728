"""In-tree build of ``libuforecon_b200.so`` (sm_100a only) and of nothing else.

    python -m uforecon_b200.build [--force]

The shared object lands next to this file so that it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("UFO_LIB_PATH", os.path.join(HERE, "libuforecon_b200.so"))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
OBJ_DIR = os.environ.get("UFO_OBJ_DIR", os.path.join(HERE, "build"))
EXTRA = os.environ.get("UFO_NVCC_EXTRA", "").split()


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    if "UFO_LIB_PATH" in os.environ:      # an explicitly selected (A/B) library is used as it is
        return False
    t = os.path.getmtime(OUT)
    hdr = os.path.join(os.path.dirname(HERE), "include", "uforecon_b200.h")
    return any(os.path.getmtime(s) > t for s in _sources() + [hdr])


def _compile(args):
    nvcc, src, obj, verbose = args
    cmd = [nvcc] + NVCC_FLAGS + EXTRA + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu of csrc/ for sm_100a (one nvcc per translation unit, in parallel) and link the .so."""
    if not force and not needs_build():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    cus = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs = [(nvcc, os.path.join(CSRC, f), os.path.join(OBJ_DIR, f[:-3] + ".o"), verbose) for f in cus]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
        results = list(ex.map(_compile, jobs))
    log = ""
    for src, rc, out in results:
        log += out
        if rc != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {os.path.basename(src)}")
    r = subprocess.run([nvcc, "-shared", "-o", OUT] + [j[2] for j in jobs] + ["-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libuforecon_b200.so failed")
    if verbose:
        sys.stderr.write(log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
