"""Build libysp.so (C ABI + sm_100a CUDA kernels) in-tree with nvcc.  No torch dependency in the library."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libysp.so")
SOURCES = ["engine.cu", "kernels_simt.cu", "kernels_fused.cu", "kernels_dlc_tc.cu", "kernels_dlc32.cu", "kernels_stem_attn.cu", "nms.cu", "conv_tc.cu", "conv_tc32.cu", "conv_halo.cu", "conv_halo32.cu", "attention_tc.cu",
           "kernels_train.cu", "kernels_train_tc.cu", "train.cu", "kernels_ghost.cu"]
HEADERS = ["common.cuh", "kernels.h", "kernels_train.h", os.path.join("..", "..", "include", "ysp.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    objs = []
    procs = []
    odir = os.path.join(HERE, "build")
    os.makedirs(odir, exist_ok=True)
    for s in SOURCES:
        o = os.path.join(odir, s.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
    tmp = SO + ".tmp"
    subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp, *objs, "-lcuda"])
    os.replace(tmp, SO)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
