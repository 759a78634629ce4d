"""GaussianFilter/ module: separable Gaussian blur of a float32 gray image.
CPU part: the oracle pinned to cv2.getGaussianKernel / cv2.GaussianBlur (the host result the reference's
gaussian.cu:441 compares every kernel with; fixtures in tests/golden/gauss_cv2.npz, generator next to
them) and the kernel under the emulator.  GPU part: parity through the C ABI.
Tolerance: 2e-6 absolute on [0, 1] data (float32 accumulation of <= 2*129 taps; the reference prints its
own max-diff against OpenCV, gaussian.cu:625-639, and accepts the same order of magnitude)."""
import os

import numpy as np
import pytest

from oracle import gf_oracle as O

TOL = 2e-6
GOLD = os.path.join(os.path.dirname(__file__), "golden", "gauss_cv2.npz")


def _golden():
    z = np.load(GOLD)
    for i in range(int(z["n"][0])):
        r, s = z[f"par{i}"]
        yield z[f"img{i}"], int(r), float(s), z[f"blur{i}"], z[f"taps{i}"]


def test_oracle_pinned_to_cv2():
    for img, r, s, blur, taps in _golden():
        assert np.array_equal(O.gaussian_kernel_1d(r, s), taps), (r, s)
        assert np.abs(O.gaussian_blur_gray(img, r, s) - blur).max() <= 1e-6, (r, s)


def _run(api, up, down, img, r, sigma, spad=0, dpad=0):
    h, w = img.shape
    src = np.zeros((h, w + spad), np.float32)
    src[:, :w] = img
    d_src, d_dst = up(src), up(np.full((h, w + dpad), np.nan, np.float32))
    api.call("gf_gaussian_gray", d_src["ptr"], d_dst["ptr"], w, h, w + spad, w + dpad, r, sigma, None)
    out = down(d_dst)
    assert np.isnan(out[:, w:]).all()            # the stride padding is never written
    return out[:, :w]


def _emu():
    from gf_backend import EmuBackend
    be = EmuBackend()

    def up(a):
        raw = np.empty(a.nbytes + 64, np.uint8)             # 64-byte aligned host memory
        off = (-raw.ctypes.data) % 64
        buf = raw[off:off + a.nbytes].view(a.dtype).reshape(a.shape)
        buf[...] = a
        return {"ptr": buf.ctypes.data, "buf": buf, "keep": raw}
    return be.api, up, (lambda d: d["buf"])


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


def _pads(img, fast):
    """fast: row strides that are multiples of 4 floats (the 4-columns-per-thread kernel, radii 1..16);
    otherwise odd strides (the thread-per-column kernel)"""
    w = img.shape[1]
    return ((-w) % 4 + 4, (-w) % 4 + 8) if fast else ((-w) % 2 + 3, (-w) % 2 + 5)


@pytest.mark.parametrize("fast", [True, False])
def test_gaussian_emulated_golden(fast):
    api, up, down = _emu()
    for img, r, s, blur, _ in _golden():
        spad, dpad = _pads(img, fast)
        q = _run(api, up, down, img, r, s, spad=spad, dpad=dpad)
        assert api.last_kernel() == ("gauss4" if fast and 1 <= r <= 16 else "gauss")
        assert np.abs(q - blur).max() <= TOL, (r, s)
        assert np.abs(q - O.gaussian_blur_gray(img, r, s)).max() <= TOL, (r, s)


@pytest.mark.parametrize("shape,r,sigma", [((1, 1), 0, 1.0), ((3, 5), 4, 2.0), ((40, 700), 8, 3.0), ((150, 260), 2, 0.8),
                                           ((30, 64), 20, 6.0), ((200, 33), 64, 20.0), ((90, 1100), 8, 2.0), ((70, 1003), 5, 1.2),
                                           ((300, 40), 7, 2.0), ((2, 3), 8, 3.0), ((60, 520), 1, 0.5), ((80, 600), 13, 4.0), ((50, 1100), 16, 6.0)])
def test_gaussian_emulated_shapes(shape, r, sigma):
    """1x1, images narrower than the radius (repeated reflection), several strips and bands, the largest
    radius, widths that are not a multiple of 4 (partial last vector), both kernels where both apply"""
    api, up, down = _emu()
    img = np.random.default_rng(sum(shape) + r).random(shape, dtype=np.float32)
    ref = O.gaussian_blur_gray(img, r, sigma)
    for fast in (True, False):
        spad, dpad = _pads(img, fast)
        q = _run(api, up, down, img, r, sigma, spad=spad, dpad=dpad)
        assert api.last_kernel() == ("gauss4" if fast and 1 <= r <= 16 else "gauss")
        assert np.abs(q - ref).max() <= TOL


def test_gaussian_errors_are_loud():
    api, up, down = _emu()
    img = np.zeros((8, 8), np.float32)
    d = up(img)
    for args in ((d["ptr"], d["ptr"], 8, 8, 8, 8, 1, 1.0, None), (d["ptr"], None, 8, 8, 8, 8, 1, 1.0, None),
                 (d["ptr"], up(img)["ptr"], 8, 8, 8, 8, 65, 1.0, None), (d["ptr"], up(img)["ptr"], 8, 8, 4, 8, 1, 1.0, None)):
        with pytest.raises(Exception):
            api.call("gf_gaussian_gray", *args)


@pytest.mark.gpu
def test_gaussian_gpu_golden():
    api, up, down = _cuda()
    for fast in (True, False):
        for img, r, s, blur, _ in _golden():
            spad, dpad = _pads(img, fast)
            q = _run(api, up, down, img, r, s, spad=spad, dpad=dpad)
            assert api.last_kernel() == ("gauss4" if fast and 1 <= r <= 16 else "gauss")
            assert np.abs(q - blur).max() <= TOL, (r, s)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,r,sigma", [((2160, 3840), 1, 0.5), ((2160, 3840), 8, 3.0), ((1080, 1921), 4, 0.0), ((333, 4097), 16, 5.0),
                                           ((700, 500), 64, 25.0), ((1, 1), 3, 1.0), ((5, 3000), 7, 2.0)])
def test_gaussian_gpu_parity(shape, r, sigma):
    """the reference's own default case (3840x2160, r=1, sigma=0.5: gaussian.cu:412-416) and wider ones against the f64 oracle;
    a constant image stays constant (the taps sum to 1)"""
    api, up, down = _cuda()
    img = np.random.default_rng(shape[0] + r).random(shape, dtype=np.float32)
    ref = O.gaussian_blur_gray(img, r, sigma)
    for fast in (True, False):
        spad, dpad = _pads(img, fast)
        q = _run(api, up, down, img, r, sigma, spad=spad, dpad=dpad)
        assert api.last_kernel() == ("gauss4" if fast and 1 <= r <= 16 else "gauss")
        assert np.abs(q - ref).max() <= TOL
    c = _run(api, up, down, np.full(shape, 0.625, np.float32), r, sigma, spad=_pads(img, True)[0], dpad=_pads(img, True)[1])
    assert np.abs(c - 0.625).max() <= 1e-6
