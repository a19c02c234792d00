"""Builds the engine's shared library in-tree with nvcc for sm_100a.

    python -m oscen_b200.build            # -> oscen_b200/lib/liboscen_b200.so

nvcc cross-compiles without a GPU.  Flags that matter for parity: -fmad=false (rustc never fuses
a*b+c; every fused op in the kernels is an explicit fma()), default IEEE div/sqrt, no fast-math,
and -ffp-contract=off for the host-side coefficient code."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "liboscen_b200.so")
SOURCES = ["engine.cu"]
DEPS = ["engine.cu", "engine_types.h", "nodes.cuh", "om_math.h", "kernels_common.cuh", "kernel_fm.cuh",
        "kernel_others.cuh", "kernel_resample.cuh", "kernel_nodes.cuh", "midi_front.hpp", "midi_device.cuh", "plan.hpp", "graphs.hpp", "dsl.hpp", "codegen.hpp", "jit.hpp", "state_io.cuh", "rtc_prelude.h"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-cudart", "static"]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, d) for d in DEPS] + [os.path.join(HERE, "..", "include", "oscen_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
