"""CPU tests of the boundary: the product library builds for sm_100a, loads, and exports every
symbol include/gf_b200.h declares (no compute calls -- there is no GPU here); the C++ drop-in
shims export the reference's mangled names; argument checking happens before any launch."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from cudaimageprocessing_b200.build import build
    return build()


def test_header_symbols_all_exported(lib_path):
    from cudaimageprocessing_b200._capi import SIGNATURES, GfApi
    hdr = open(os.path.join(ROOT, "include", "gf_b200.h")).read()
    declared = set(re.findall(r"\b(gf_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(SIGNATURES), declared ^ set(SIGNATURES)
    api = GfApi(ctypes.CDLL(lib_path))          # raises AttributeError on a missing symbol
    assert api.cdll.gf_version() >= 100


def test_shim_exports_reference_signatures(lib_path):
    """The reference's main.cpp links against these exact C++ symbols
    (guided_filter.h:19,30; guided_filter_d.h:6-21)."""
    out = subprocess.check_output(["nm", "-D", "--defined-only", "-C", lib_path], text=True)
    for sig in ["GuidedFilter::init(int, int, int, int)", "GuidedFilter::run(float*, float*, float*, int, float)",
                "hBoxFilter(float*, float*, float*, int4 const&, int4 const&, int)",
                "hMultiply(float*, float*, float*, int4 const&, int4 const&)",
                "hCalcA(float*, float*, float*, float*, float*, int4 const&, int4 const&, float)",
                "hCalcB(float*, float*, float*, float*, int4 const&, int4 const&)",
                "hLinearTransform(float*, float*, float*, float*, int4 const&, int4 const&)",
                "hGuidedFilter(float*, float*, float*, float*, float*, float, int, int, int, int)"]:
        assert sig in out, sig


def test_sass_is_sm100a(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_argument_errors_before_launch(lib_path):
    """Bad arguments are refused with a status and a message, never a crash (and never a launch)."""
    from cudaimageprocessing_b200._capi import GF_ERR_INVALID, GF_ERR_UNSUPPORTED, GfApi, GfError
    api = GfApi(ctypes.CDLL(lib_path))
    n0 = api.launch_count()
    a = np.zeros((8, 8), np.float32)
    p = a.ctypes.data
    with pytest.raises(GfError) as e:
        api.call("gf_guided_gray", None, p, p, None, None, 8, 8, 0, 0, 0, 0, 2, 0.1, 0, None)
    assert e.value.status == GF_ERR_INVALID
    with pytest.raises(GfError) as e:
        api.call("gf_guided_gray", p, p, p, None, None, 8, 8, 4, 0, 0, 0, 2, 0.1, 0, None)   # stride < width
    assert e.value.status == GF_ERR_INVALID
    h = ctypes.c_void_p()
    with pytest.raises(GfError) as e:
        api.call("gf_create", ctypes.addressof(h), 8, 8, 2, 1)
    assert e.value.status == GF_ERR_UNSUPPORTED and "Do not support channel" in str(e.value)
    with pytest.raises(GfError) as e:   # strip without its halo rows
        api.call("gf_guided_gray_strip", p, p, p, 8, 64, 16, 8, 16, 8, 0, 0, 0, 2, 0.1, 0, None)
    assert e.value.status == GF_ERR_INVALID and "halo" in str(e.value)
    assert api.launch_count() == n0


def test_no_gpu_means_error_not_fallback(lib_path):
    """Without a device a compute call must fail loudly (GF_ERR_CUDA), not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cudaimageprocessing_b200._capi import GF_ERR_CUDA, GF_ERR_NOMEM, GfApi, GfError
    api = GfApi(ctypes.CDLL(lib_path))
    a = np.ones((64, 64), np.float32)
    q = np.zeros((64, 64), np.float32)
    with pytest.raises(GfError) as e:
        api.call("gf_guided_gray", a.ctypes.data, a.ctypes.data, q.ctypes.data, None, None, 64, 64, 0, 0, 0, 0, 2, 0.1, 0, None)
    assert e.value.status in (GF_ERR_CUDA, GF_ERR_NOMEM)
    assert not q.any()


def test_dropin_demo_compiles_against_shims(lib_path, tmp_path):
    """A host program written against the reference's headers (the call sequences of
    main.cpp:141-150 and :257) compiles and links against include/ + libgf_b200.so unchanged."""
    exe = tmp_path / "dropin_demo"
    cmd = ["nvcc", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "dropin", "dropin_demo.cpp"),
           "-x", "cu", "-o", str(exe), "-L", os.path.dirname(lib_path), "-lgf_b200", "-Xlinker", "-rpath=" + os.path.dirname(lib_path)]
    cmd = [c for c in cmd if c not in ("-x", "cu")]
    subprocess.check_call(cmd)
    assert exe.exists()
