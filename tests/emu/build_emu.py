"""Builds tests/emu/libgf_emu.so: the product's kernels and API layer compiled by g++ against
the test-only SIMT emulator (cuda_emu.h).  Used by the CPU test-suite to check kernel logic
where there is no GPU.  The product package never loads this library."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cudaimageprocessing_b200", "csrc")
LIB = os.path.join(HERE, "libgf_emu.so")


def build_emu(force: bool = False) -> str:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_emu.cpp")] + \
           [os.path.join(ROOT, "include", "gf_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["/usr/bin/g++", "-O1", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-DGF_CPU_EMU",
           "-I", HERE, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-x", "c++", os.path.join(CSRC, "gf_api.cu"), os.path.join(HERE, "cuda_emu.cpp"), "-o", LIB]
    if os.path.exists(os.path.join(CSRC, "gf_fast.cuh")):
        cmd.insert(1, "-DGF_HAVE_FAST")
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_emu(force=True))
