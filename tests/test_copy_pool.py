"""The host path's copy pool (csrc/gf_copy_pool.h: persistent threads that copy one block together, used to stage
pageable caller buffers) is plain C++: built with g++ and stress-tested here without a GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_copy_pool_stress(tmp_path):
    exe = tmp_path / "copy_pool_stress"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I", os.path.join(ROOT, "cudaimageprocessing_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "copy_pool_stress.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe), "200"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "copy pool ok" in out.stdout
