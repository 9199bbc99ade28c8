"""Build libroft_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The shared library has a plain C ABI (include/roft_b200.h) and depends only on the CUDA runtime;
it is what travels to the GPU box.  ``python -m roft_b200.build`` rebuilds it.
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "roft_b200", "csrc")
LIB = os.path.join(ROOT, "roft_b200", "libroft_b200.so")
SOURCES = ["roftb_api.cu", "mask_sync.cu", "worklist.cu", "velocity_track.cu", "ukf_batch.cu", "extract.cu", "render.cu"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "roft_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(ROOT, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation of roft_b200 failed")
    # (the arch on the link line keeps nvcc from adding a default sm_52 stub image to the library)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
