"""Build libloe_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python cs-304-speech-recognition-code_b200/build.py [--force] [--verbose]

The library has no dependency on torch or Python: it is a plain C-ABI shared object
(include/loe_b200.h) that the host package loads with ctypes.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libloe_b200.so")
SOURCES = ["common.cu", "mfcc.cu", "mfcc_ex.cu", "emission.cu", "emission_tc.cu", "viterbi.cu", "viterbi_warp.cu", "kmeans.cu", "mstep.cu", "vad.cu", "dtw.cu", "decoder.cu", "emission_h16.cu", "emission_gmm.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
SPILL_SENSITIVE = {"emission_h16.cu"}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "loe_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--use_fast_math=false"]
        cmd = [c for c in cmd if c != "--use_fast_math=false"]
        if verbose or src in SPILL_SENSITIVE:
            cmd += ["-Xptxas", "-v"]
        cmd += ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src}\n{out}\n")
        failed |= p.returncode != 0
        if p.returncode == 0 and src in SPILL_SENSITIVE:
            # the warp-specialised tensor-core kernels lose ~30 % when ptxas spills inside their role loops (the
            # setmaxnreg split leaves no slack): make a regression visible at build time
            import re
            for m in re.finditer(r"Function properties for (\S+)\s*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores", out):
                if int(m.group(3)) > 0:
                    sys.stderr.write(f"WARNING: {src}: {m.group(1)} spills {m.group(3)} bytes -- rebalance kProducerRegs / kEpilogueRegs\n")
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [_nvcc(), *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
