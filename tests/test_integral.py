"""Integral/ module: the summed-area table of a uint8 image.
CPU part: the oracle against cv2.integral (what Integral/main.cpp:124 compares the reference with)
and the kernels under the emulator.  GPU part: bit-exact against the oracle at the sizes the
reference's res.log lists, the padded (hAligned4Integral) form, int32 wrap-around and the int64 form."""
import ctypes

import numpy as np
import pytest

from oracle import gf_oracle as O


def test_oracle_matches_cv2_integral():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for (h, w) in ((1, 1), (7, 13), (211, 307), (600, 801)):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ref = cv2.integral(img)[1:, 1:]
        assert np.array_equal(O.integral_u8(img, np.int64), ref.astype(np.int64))
        assert np.array_equal(O.integral_u8(img, np.int32), ref.astype(np.int32))


def _run(api, up, down, img, i64=False, pad_to=None, stride_pad=0, with_scratch=False):
    h, w = img.shape
    ss = w + stride_pad
    src = np.zeros((h, ss), np.uint8)
    src[:, :w] = img
    d_src = up(src)
    dt = np.int64 if i64 else np.int32
    if pad_to:
        dh, dw = pad_to
        d_out = up(np.full((dh, dw), -1, dt))
        api.call("gf_integral_u8_i32_padded", d_src["ptr"], d_out["ptr"], w, h, ss, dw, dh, None)
        return down(d_out)
    d_out = up(np.full((h, w + stride_pad), -1, dt))
    d_scr = up(np.zeros(((h + 15) // 16 + 1, w), dt)) if with_scratch else {"ptr": None}
    api.call("gf_integral_u8_i64" if i64 else "gf_integral_u8_i32", d_src["ptr"], d_out["ptr"], d_scr["ptr"], w, h, ss, w + stride_pad, None)
    return down(d_out)[:, :w]


def _emu():
    from gf_backend import EmuBackend
    be = EmuBackend()

    def up(a):
        b = be._aligned(a.shape).view(np.float32)          # 64-byte aligned host memory
        raw = np.empty(a.nbytes + 64, np.uint8)
        off = (-raw.ctypes.data) % 64
        buf = raw[off:off + a.nbytes].view(a.dtype).reshape(a.shape)
        buf[...] = a
        return {"ptr": buf.ctypes.data, "buf": buf, "keep": raw}
    return be.api, up, (lambda d: d["buf"])


@pytest.mark.parametrize("shape", [(1, 1), (5, 9), (40, 256), (33, 300), (70, 520), (17, 1030)])
def test_integral_emulated(shape):
    api, up, down = _emu()
    rng = np.random.default_rng(sum(shape))
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(_run(api, up, down, img), O.integral_u8(img, np.int32))
    assert np.array_equal(_run(api, up, down, img, i64=True, with_scratch=True), O.integral_u8(img, np.int64))
    h, w = shape
    dh, dw = (h + 3) // 4 * 4, (w + 3) // 4 * 4
    out = _run(api, up, down, img, pad_to=(dh, dw))
    ext = np.zeros((dh, dw), np.uint8)
    ext[:h, :w] = img
    assert np.array_equal(out, O.integral_u8(ext, np.int32))


@pytest.mark.parametrize("opts", [{"GF_SAT_TWO_PASS": 0}, {"GF_SAT_TWO_PASS": 0, "GF_SAT_HB": 3}, {"GF_SAT_TWO_PASS": 0, "GF_SAT_HB": 1000},
                                  {"GF_SAT_TWO_PASS": 1}, {"GF_SAT_TWO_PASS": 1, "GF_SAT_HB": 5}])
def test_integral_emulated_forms(opts):
    """both forms forced (reduce-then-scan / two-pass; the library picks by size and type), many short bands, one band"""
    api, up, down = _emu()
    for k, v in opts.items():
        api.set_option(k, v)
    try:
        for shape in ((70, 520), (33, 300), (19, 8)):
            img = np.random.default_rng(sum(shape)).integers(0, 256, shape, dtype=np.uint8)
            assert np.array_equal(_run(api, up, down, img), O.integral_u8(img, np.int32))
            assert np.array_equal(_run(api, up, down, img, i64=True, with_scratch=True), O.integral_u8(img, np.int64))
    finally:
        for k in opts:
            api.set_option(k, -1)


def _cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import cudaimageprocessing_b200 as pkg
    api = pkg.api()

    def up(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        return {"ptr": t.data_ptr(), "t": t}

    def down(d):
        torch.cuda.synchronize()
        return d["t"].cpu().numpy()
    return api, up, down


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2160, 3840), (5941, 5910), (5837, 4151), (1, 7), (1080, 1921), (333, 4097)])
def test_integral_gpu_bit_exact(shape):
    """sizes include the first two of the reference's Integral/res.log (5910x5941, 4151x5837: 'Max difference
    of NPPI and CUDA: 0'); int32 results equal the int64 oracle modulo 2^32, bit for bit."""
    api, up, down = _cuda()
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(_run(api, up, down, img), O.integral_u8(img, np.int32))
    assert api.last_kernel() == "integral_i32"
    assert np.array_equal(_run(api, up, down, img, stride_pad=8, with_scratch=True), O.integral_u8(img, np.int32))
    h, w = shape
    dh, dw = (h + 3) // 4 * 4, (w + 3) // 4 * 4
    ext = np.zeros((dh, dw), np.uint8)
    ext[:h, :w] = img
    assert np.array_equal(_run(api, up, down, img, pad_to=(dh, dw)), O.integral_u8(ext, np.int32))


@pytest.mark.gpu
def test_integral_gpu_overflow_and_int64():
    """an all-255 image of 3000x3000 px sums to 2.3e9 > 2^31: the int32 table wraps exactly, the int64 one does not."""
    api, up, down = _cuda()
    img = np.full((3000, 3000), 255, np.uint8)
    assert np.array_equal(_run(api, up, down, img), O.integral_u8(img, np.int32))
    out = _run(api, up, down, img, i64=True)
    assert out[-1, -1] == 255 * 3000 * 3000 and np.array_equal(out, O.integral_u8(img, np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("two_pass", [0, 1])
def test_integral_gpu_both_forms(two_pass):
    """the reduce-then-scan form and the two-pass form, each forced, on an aligned 4K table, an unaligned one and the
    int64 table (the library chooses between them by size and type: csrc/gf_integral.cuh, gf_sat_launch)."""
    api, up, down = _cuda()
    api.set_option("GF_SAT_TWO_PASS", two_pass)
    try:
        for shape in ((2160, 3840), (2161, 3001)):
            img = np.random.default_rng(shape[1]).integers(0, 256, shape, dtype=np.uint8)
            assert np.array_equal(_run(api, up, down, img), O.integral_u8(img, np.int32))
            assert np.array_equal(_run(api, up, down, img, i64=True, with_scratch=True), O.integral_u8(img, np.int64))
    finally:
        api.set_option("GF_SAT_TWO_PASS", -1)
