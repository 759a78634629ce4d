"""uint8 in / uint8 out: gf_guided_gray_u8 = convertTo(1/255) -> guided filter ->
convertTo(CV_8U, 255) fused into one kernel.  Checked at the uint8 level against the oracle's
restatement of the reference demo pipeline, on BASELINE config 1's image (the bundled picture,
gray self-guide, r=8, eps=1e-2) and on random planes.  As with the float KAT, results sitting on a
rounding knife-edge may differ by 1 LSB; we allow 1 LSB on at most 0.01 % of the pixels."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import gf_oracle as O


def test_u8_integer_domain_is_the_same_filter():
    """The uint8 build filters I' = 255 I, p' = 255 p with eps' = 255^2 eps and rounds q' = 255 q directly
    (no conversion instructions; integer-valued sums below 2^24 are exact in float32).  In exact arithmetic
    that is the reference pipeline; here: float64 on both sides agrees to rounding."""
    rng = np.random.default_rng(0)
    I = rng.integers(0, 256, (40, 50), dtype=np.uint8)
    p = rng.integers(0, 256, (40, 50), dtype=np.uint8)
    q1 = O.guided_filter_gray(I.astype(np.float64) / 255, p.astype(np.float64) / 255, 4, 1e-2, 0, np.float64) * 255
    q2 = O.guided_filter_gray(I.astype(np.float64), p.astype(np.float64), 4, 1e-2 * 65025, 0, np.float64)
    assert np.abs(q1 - q2).max() < 1e-9


def _call(api, up, down, I, p, r, eps, border, pad=0):
    h, w = I.shape
    s = w + pad

    def pitched(a):
        b = np.zeros((h, s), np.uint8)
        b[:, :w] = a
        return b
    dI, dp, dq = up(pitched(I)), up(pitched(p)), up(np.zeros((h, s), np.uint8))
    api.call("gf_guided_gray_u8", dI["ptr"], dp["ptr"], dq["ptr"], w, h, s, s, s, r, eps, border, None)
    return down(dq)[:, :w]


def _check(q, ref, max_frac=1e-4):
    d = q.astype(int) - ref.astype(int)
    assert np.abs(d).max() <= 1, np.abs(d).max()
    assert np.count_nonzero(d) <= max(2, int(max_frac * d.size)), np.count_nonzero(d)


def _bundled():
    cv2 = pytest.importorskip("cv2")
    g = cv2.imread(os.path.join(GOLDEN, "adobe_guide_gray_u8.png"), cv2.IMREAD_UNCHANGED)
    s = cv2.imread(os.path.join(GOLDEN, "adobe_src_gray_u8.png"), cv2.IMREAD_UNCHANGED)
    return g, s


@pytest.mark.parametrize("border", [0, 1])
def test_u8_emulated(border):
    from test_integral import _emu
    api, up, down = _emu()
    rng = np.random.default_rng(3)
    for (h, w, r) in ((40, 264, 4), (60, 256, 8), (70, 300, 7)):
        if border == 1 and w % 8:
            continue
        I = rng.integers(0, 256, (h, w), dtype=np.uint8)
        p = np.clip(I.astype(int) + rng.integers(-20, 21, (h, w)), 0, 255).astype(np.uint8)
        q = _call(api, up, down, I, p, r, 1e-2, border, pad=(-w) % 8)
        assert api.last_kernel() == f"s8u8_r{r}"
        _check(q, O.guided_filter_gray_u8(I, p, r, 1e-2, border))
    with pytest.raises(Exception):                      # no silent fallback: unsupported radius is an error
        _call(api, up, down, I, p, 5, 1e-2, 0, pad=(-w) % 8)


@pytest.mark.gpu
def test_u8_config1_bundled_image_gpu():
    """BASELINE config 1: the bundled image, gray self-guide, r=8, eps=1e-2 -- uint8 in, uint8 out."""
    from test_integral import _cuda
    api, up, down = _cuda()
    g, s = _bundled()
    h, w = g.shape
    w8 = w // 8 * 8                                     # the planes are 8-byte aligned after the crop to a multiple of 8
    for (I, p, r, eps) in ((g[:, :w8], g[:, :w8], 8, 1e-2), (g[:, :w8], s[:h, :w8] if s.shape == g.shape else g[:, :w8], 7, 0.3)):
        I, p = np.ascontiguousarray(I), np.ascontiguousarray(p)
        q = _call(api, up, down, I, p, r, eps, 0)
        assert api.last_kernel() == f"s8u8_r{r}"
        _check(q, O.guided_filter_gray_u8(I, p, r, eps, 0))


@pytest.mark.gpu
@pytest.mark.parametrize("r", [1, 2, 3, 4, 5, 6, 7])
def test_u8_reference_sweep_radii_gpu(r):
    """the radii of the reference's sweep (GuidedFilter/run.py:4-6: r = 1..7, eps 0.3) on uint8 planes"""
    from test_integral import _cuda
    api, up, down = _cuda()
    rng = np.random.default_rng(40 + r)
    I = rng.integers(0, 256, (1080, 1920), dtype=np.uint8)
    p = rng.integers(0, 256, (1080, 1920), dtype=np.uint8)
    q = _call(api, up, down, I, p, r, 0.3, 0)
    assert api.last_kernel() == f"s8u8_r{r}"
    _check(q, O.guided_filter_gray_u8(I, p, r, 0.3, 0))


@pytest.mark.gpu
@pytest.mark.parametrize("border", [0, 1, 2])
def test_u8_random_4k_gpu(border):
    from test_integral import _cuda
    api, up, down = _cuda()
    rng = np.random.default_rng(11)
    I = rng.integers(0, 256, (2160, 3840), dtype=np.uint8)
    p = rng.integers(0, 256, (2160, 3840), dtype=np.uint8)
    q = _call(api, up, down, I, p, 8, 1e-2, border)
    assert api.last_kernel() == "s8u8_r8"
    from oracle import c_oracle as C
    ref = O.to_u8(C.guided_gray_f32(O.u8_to_f32(I), O.u8_to_f32(p), 8, 1e-2, border, max(1, (os.cpu_count() or 2) - 1)))
    _check(q, ref)


# ---- north_star: "bit-exact integral sums for uint8 input (64-bit accumulators ...)" --------------------------------
def _window_sums(api, up, down, I, p, r, border):
    h, w = I.shape
    dI, dp = up(I), up(p)
    outs = [up(np.zeros((h, w), np.int64)) for _ in range(4)]
    api.call("gf_window_sums_u8", dI["ptr"], dp["ptr"], outs[0]["ptr"], outs[1]["ptr"], outs[2]["ptr"], outs[3]["ptr"], w, h, 0, 0, r,
             border, None)
    return [down(o) for o in outs]


def _check_window_sums(api, up, down, shape, r, border, seed, bright=False):
    rng = np.random.default_rng(seed)
    lo = 246 if bright else 0          # bright planes: sum(I p) and sum(I I) exceed 2^24, where float32 would round
    I = rng.integers(lo, 256, shape, dtype=np.uint8)
    p = rng.integers(lo, 256, shape, dtype=np.uint8)
    si, sp, sip, sii = _window_sums(api, up, down, I, p, r, border)
    I64, p64 = I.astype(np.int64), p.astype(np.int64)

    def ref(a):          # box_sum_u8 takes uint8; the products go through the same exact int64 code path
        return O.box_sum_u8(a, r, border) if a.dtype == np.uint8 else _box_sum_i64(a, r, border)
    assert np.array_equal(si, O.box_sum_u8(I, r, border))
    assert np.array_equal(sp, O.box_sum_u8(p, r, border))
    assert np.array_equal(sip, _box_sum_i64(I64 * p64, r, border))
    assert np.array_equal(sii, _box_sum_i64(I64 * I64, r, border))
    if bright:
        assert sii.max() > 2 ** 24


def _box_sum_i64(a, r, border):
    """exact int64 window sums of an int64 plane (same construction as oracle.box_sum_u8)"""
    def axis_sum(a, axis):
        n = a.shape[axis]
        a = np.moveaxis(a, axis, 0)
        if border == O.BORDER_TRUNCATE:
            pad = np.zeros((r,) + a.shape[1:], dtype=np.int64)
            e = np.concatenate([pad, a, pad], axis=0)
        else:
            e = a[O.border_index(np.arange(-r, n + r), n, border)]
        c = np.cumsum(e, axis=0)
        c = np.concatenate([np.zeros((1,) + c.shape[1:], dtype=np.int64), c], axis=0)
        return np.moveaxis(c[2 * r + 1: 2 * r + 1 + n] - c[0:n], 0, axis)
    return axis_sum(axis_sum(a, 1), 0)


@pytest.mark.parametrize("shape,r,border", [((40, 70), 8, 0), ((33, 2100), 16, 2), ((50, 64), 8, 1), ((21, 30), 20, 0)])
def test_window_sums_u8_emulated(shape, r, border):
    from test_integral import _emu
    api, up, down = _emu()
    _check_window_sums(api, up, down, shape, r, border, seed=7, bright=True)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,r,border", [((2160, 3840), 8, 0), ((2160, 3840), 16, 0), ((1080, 1920), 8, 1), ((777, 1301), 32, 2)])
def test_window_sums_u8_gpu_bit_exact(shape, r, border):
    """The four stage-1 window sums of uint8 planes, bit for bit against int64 numpy, at r = 8 and 16 on 4K planes whose
    sums of products lie above 2^24."""
    from test_integral import _cuda
    api, up, down = _cuda()
    _check_window_sums(api, up, down, shape, r, border, seed=11, bright=True)
