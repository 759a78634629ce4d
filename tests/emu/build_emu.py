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
    fresh = lambda: os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps)
    if not force and fresh():
        return LIB
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:     # several test processes may get here at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():
                return LIB
            return _build(deps)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build(deps) -> str:
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    flags = ["-O1", "-std=c++17", "-fopenmp", "-fPIC", "-DGF_CPU_EMU", "-I", HERE, "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if os.path.exists(os.path.join(CSRC, "gf_fast.cuh")):
        flags.insert(0, "-DGF_HAVE_FAST")
    units = [os.path.join(CSRC, f) for f in ("gf_api.cu", "gf_tu_s8.cu", "gf_tu_ws.cu", "gf_tu_c4.cu")] + [os.path.join(HERE, "cuda_emu.cpp")]
    with tempfile.TemporaryDirectory(prefix="gfemu_") as tmp:
        def one(src):
            obj = os.path.join(tmp, os.path.basename(src).rsplit(".", 1)[0] + ".o")
            subprocess.check_call(["/usr/bin/g++"] + flags + ["-x", "c++", "-c", src, "-o", obj])
            return obj
        with ThreadPoolExecutor(max_workers=5) as ex:          # one translation unit per kernel family, in parallel
            objs = list(ex.map(one, units))
        tmp_lib = LIB + f".tmp{os.getpid()}"
        subprocess.check_call(["/usr/bin/g++", "-shared", "-fopenmp", "-o", tmp_lib] + objs)
        os.replace(tmp_lib, LIB)
    return LIB


if __name__ == "__main__":
    print(build_emu(force=True))
