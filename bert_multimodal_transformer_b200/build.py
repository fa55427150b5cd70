"""Builds bert_multimodal_transformer_b200/libmagb200.so in-tree with nvcc for sm_100a.

    python -m bert_multimodal_transformer_b200.build [--force]

The shared library is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmagb200.so")
STAMP = os.path.join(PKG_DIR, "csrc", ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math" if False else "-DMAGB200_NO_FAST_MATH",   # erf/exp accuracy matters for 1e-3 parity
    "-Xcompiler", "-fPIC", "-shared",
    "--threads", "8",
]


def _find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libmagb200.so cannot be built")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))
                   + glob.glob(os.path.join(CSRC, "*.h"))
                   + [os.path.join(os.path.dirname(PKG_DIR), "include", "magb200.h")])
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh():
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _digest()


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ into libmagb200.so (no-op when sources are unchanged)."""
    if not force and is_fresh():
        return LIB_PATH
    nvcc = _find_nvcc()
    objs = []
    build_dir = os.path.join(CSRC, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in _sources():
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif verbose and out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("building libmagb200.so failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    subprocess.run(link, check=True)
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB_PATH


def build_variant(out_path, extra_flags):
    """A second library with extra compile flags (A/B measurements, e.g. -DMAGB200_PHILOX); objects go to a
    separate directory, the default library and its stamp are untouched.  Select it with MAGB200_LIB_PATH."""
    nvcc = _find_nvcc()
    build_dir = os.path.join(CSRC, "build", "variant_" + hashlib.sha256(" ".join(extra_flags).encode()).hexdigest()[:8])
    os.makedirs(build_dir, exist_ok=True)
    procs, objs = [], []
    for src in _sources():
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out_path] + objs, check=True)
    return out_path


if __name__ == "__main__":
    if "--philox" in sys.argv:
        print(build_variant(os.path.join(PKG_DIR, "libmagb200_philox.so"), ["-DMAGB200_PHILOX"]))
    else:
        path = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
        print(path)
