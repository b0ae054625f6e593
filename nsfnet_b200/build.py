"""Builds nsfnet_b200/libnsf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["nsf_ffma.cu", "nsf_capi.cu", "nsf_umma.cu", "nsf_pm_jet.cu", "nsf_value_fwd.cu", "nsf_aux.cu", "nsf_eval.cu"]
OUT = os.path.join(HERE, "libnsf_b200.so")


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "nsf_b200.h"), os.path.abspath(__file__)]
    return d


def build_cuda(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """`out` / `defines`: kernel experiments build variants of the library beside the product (scripts/_bin/)"""
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in _deps()):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-diag-suppress", "177", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-D" + d for d in defines]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out, "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libnsf_b200.so")
    return out


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
