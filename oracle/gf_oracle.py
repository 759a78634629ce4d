"""CPU oracle for the GuidedFilter hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, in numpy, the algorithm of the reference's guided filter so
that the CUDA path can be checked against it.  It is never imported by the
product package (`cudaimageprocessing_b200`); only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may use it, and there only as the checker.

What each function follows in the reference (paths relative to /root/reference):

* gray guide, BORDER_REFLECT101, fixed divisor (2r+1)^2
    GuidedFilter/main.cpp:236-252  (the "mycv" cv::blur composition; same maths
    at main.cpp:30-47) and the fused GPU path GuidedFilter/guided_filter_d.cu:
    421-858 (`gCalcAB` + `gWeightByABm`, border rule `reflectBorder` :415-418).
* gray / per-channel, BORDER_TRUNCATE (window clipped to the image, divided by
  the true pixel count)
    GuidedFilter/guided_filter.cpp:28-66 (`GuidedFilter::run`) with
    guided_filter_d.cu:251-262 (`gIntegralToMean` clip + area) and the
    point-wise formulas :306-323 (a), :349-362 (b), :382-395 (q).
* BORDER_REFLECT is what `cv::ximgproc::guidedFilter` (main.cpp:234) uses; that
  function lives in OpenCV-contrib (un-vendored, version unpinned:
  GuidedFilter/CMakeLists.txt:8), so it is restated from He et al., "Guided
  Image Filtering", TPAMI 2013, eqs. (5), (6), (8) / (19)-(21).
* colour guide (3x3 covariance): absent from the reference's own code (SURVEY
  fact 5); restated from He et al. eqs. (19)-(21).  PARITY UNPINNED for this
  variant: no golden vector of the reference exercises it.

Pinning: `tests/test_oracle_golden.py` checks `guided_filter_gray` (float32
mode) against the reference's golden PNG `adobe_image_4_myres.png`
(r=7, eps=0.3; GuidedFilter/main.cpp:295-304) through the committed fixtures in
`tests/golden/`.
"""
from __future__ import annotations

import numpy as np

BORDER_REFLECT101 = 0   # gfedcb|abcdefgh|gfedcba   (cv::BORDER_DEFAULT, reflectBorder)
BORDER_TRUNCATE = 1     # window clipped to the image, divide by true count (gIntegralToMean)
BORDER_REFLECT = 2      # fedcba|abcdefgh|hgfedcb   (cv::ximgproc::guidedFilter)

_BORDER_NAMES = {BORDER_REFLECT101: "reflect101", BORDER_TRUNCATE: "truncate", BORDER_REFLECT: "reflect"}


def border_index(idx, n: int, mode: int) -> np.ndarray:
    """Map (possibly out-of-range) coordinates to source coordinates.

    REFLECT101 follows `reflectBorder` (guided_filter_d.cu:415-418) for a single
    overshoot and extends it periodically (period 2n-2) the way OpenCV's
    borderInterpolate does, so it stays defined when r >= n.  TRUNCATE returns -1
    for coordinates outside the image (they contribute nothing).
    """
    idx = np.asarray(idx, dtype=np.int64)
    if mode == BORDER_TRUNCATE:
        return np.where((idx >= 0) & (idx < n), idx, -1)
    if n == 1:
        return np.zeros_like(idx)
    if mode == BORDER_REFLECT101:
        period = 2 * n - 2
        m = np.mod(idx, period)
        return np.where(m < n, m, period - m)
    if mode == BORDER_REFLECT:
        period = 2 * n
        m = np.mod(idx, period)
        return np.where(m < n, m, period - 1 - m)
    raise ValueError(f"unknown border mode {mode}")


def _window_sum_axis(a: np.ndarray, r: int, mode: int, axis: int) -> np.ndarray:
    """Sum over [i-r, i+r] along `axis` with the border rule, float64 accumulate."""
    n = a.shape[axis]
    ext = border_index(np.arange(-r, n + r), n, mode)
    a = np.moveaxis(a, axis, 0)
    if mode == BORDER_TRUNCATE:
        pad = np.zeros((r,) + a.shape[1:], dtype=a.dtype)
        e = np.concatenate([pad, a, pad], axis=0)
    else:
        e = a[ext]
    c = np.cumsum(e, axis=0, dtype=np.float64)
    c = np.concatenate([np.zeros((1,) + c.shape[1:], dtype=np.float64), c], axis=0)
    s = c[2 * r + 1: 2 * r + 1 + n] - c[0:n]
    return np.moveaxis(s, 0, axis)


def window_count(n: int, r: int, mode: int) -> np.ndarray:
    """Number of pixels the 1-D window [i-r, i+r] covers (true count for TRUNCATE)."""
    if mode == BORDER_TRUNCATE:
        i = np.arange(n)
        return (np.minimum(n - 1, i + r) - np.maximum(0, i - r) + 1).astype(np.float64)
    return np.full(n, 2 * r + 1, dtype=np.float64)


def box_sum(img: np.ndarray, r: int, mode: int) -> np.ndarray:
    """(2r+1)^2 window sums of an HxW or HxWxC array, float64."""
    a = np.asarray(img, dtype=np.float64)
    return _window_sum_axis(_window_sum_axis(a, r, mode, 1), r, mode, 0)


def box_mean(img: np.ndarray, r: int, mode: int, dtype=np.float64) -> np.ndarray:
    """Box mean.

    REFLECT101 == cv::blur(img, Size(2r+1,2r+1)) (main.cpp:241-244): float32 data,
    double running sums, one multiply by 1/(2r+1)^2, cast back.  TRUNCATE ==
    hBoxFilter (guided_filter_d.cu:868-924) with exact sums instead of the
    reference's float32 integral image.
    """
    img = np.asarray(img)
    h, w = img.shape[:2]
    s = box_sum(img, r, mode)
    cnt = np.outer(window_count(h, r, mode), window_count(w, r, mode))
    if s.ndim == 3:
        cnt = cnt[:, :, None]
    return (s * (1.0 / cnt)).astype(dtype)


def guided_filter_gray(I, p, r: int, eps: float, mode: int = BORDER_REFLECT101,
                       dtype=np.float64, return_ab: bool = False):
    """q = mean(a)*I + mean(b) with a = cov(I,p)/(var(I)+eps), b = mean(p) - a*mean(I).

    Follows GuidedFilter/main.cpp:236-252 statement by statement.  With
    dtype=float32 every Mat is float32 as in the reference (the KAT mode); with
    float64 it is the ground truth the CUDA path is held to (<= 1e-4).
    I and p may be HxW, or HxWxC with equal C (channels filtered independently --
    `channel1 == channel2` branches, guided_filter_d.cu:968-971), or I HxW with
    p HxWxC (`CN1` branches, :972-975).
    """
    I = np.asarray(I).astype(dtype)
    p = np.asarray(p).astype(dtype)
    if I.ndim == 2 and p.ndim == 3:
        I = I[:, :, None]
    eps = dtype(eps)
    pm = box_mean(p, r, mode, dtype)
    im = box_mean(I, r, mode, dtype)
    ipm = box_mean(p * I, r, mode, dtype)
    iim = box_mean(I * I, r, mode, dtype)
    a = (ipm - pm * im) / (iim - im * im + eps)
    b = pm - a * im
    am = box_mean(a, r, mode, dtype)
    bm = box_mean(b, r, mode, dtype)
    q = (am * I + bm).astype(dtype)
    if return_ab:
        return q, a.astype(dtype), b.astype(dtype)
    return q


def guided_filter_color(I, p, r: int, eps: float, mode: int = BORDER_REFLECT101,
                        dtype=np.float64):
    """Colour-guide guided filter (He et al. 2013, eqs. 19-21).  PARITY UNPINNED.

    I: HxWx3 guide.  p: HxW or HxWxC.  a_k = (Sigma_k + eps*U)^-1 (mean(I p) -
    mu_k mean(p)), b_k = mean(p) - a_k^T mu_k, q = mean(a)^T I + mean(b).
    """
    I = np.asarray(I).astype(dtype)
    p = np.asarray(p).astype(dtype)
    squeeze = p.ndim == 2
    if squeeze:
        p = p[:, :, None]
    h, w, _ = I.shape
    mu = box_mean(I, r, mode, dtype)                         # H W 3
    pm = box_mean(p, r, mode, dtype)                         # H W C
    # 6 unique second moments
    idx = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    sig = np.empty((h, w, 3, 3), dtype=dtype)
    for (i, j) in idx:
        m = box_mean(I[:, :, i] * I[:, :, j], r, mode, dtype) - mu[:, :, i] * mu[:, :, j]
        sig[:, :, i, j] = m
        sig[:, :, j, i] = m
    sig = sig + dtype(eps) * np.eye(3, dtype=dtype)
    out = np.empty_like(p)
    inv = np.linalg.inv(sig.astype(np.float64)).astype(dtype)
    for c in range(p.shape[2]):
        pc = p[:, :, c]
        cov = np.stack([box_mean(I[:, :, i] * pc, r, mode, dtype) - mu[:, :, i] * pm[:, :, c]
                        for i in range(3)], axis=-1)         # H W 3
        a = np.einsum("hwij,hwj->hwi", inv, cov).astype(dtype)
        b = pm[:, :, c] - np.sum(a * mu, axis=-1)
        am = box_mean(a, r, mode, dtype)
        bm = box_mean(b, r, mode, dtype)
        out[:, :, c] = np.sum(am * I, axis=-1) + bm
    return out[:, :, 0] if squeeze else out


def guided_filter_class_run(I, p, r: int, eps: float, dtype=np.float64):
    """`GuidedFilter::run` (guided_filter.cpp:28-66): TRUNCATE border, channel
    combinations (1,1), (3,3) per channel, (guide 1, src 3).  The reference's
    `gCalcBCN1` bug (guided_filter_d.cu:371-372) is NOT reproduced."""
    return guided_filter_gray(I, p, r, eps, BORDER_TRUNCATE, dtype)


def to_u8(q) -> np.ndarray:
    """Mat::convertTo(CV_8U, 255.0) (main.cpp:295-297): saturate_cast<uchar>(cvRound(q*255)).
    OpenCV scales a CV_32F Mat in float32 (cvt32f8u: `src*a + b` with float a, b) and cvRound
    is round-half-to-even, so a float32 input is multiplied in float32 here too -- the KAT
    has thousands of pixels sitting on x.5 after the 5x bilinear upscale."""
    q = np.asarray(q)
    if q.dtype == np.float32:
        v = q * np.float32(255.0)
    else:
        v = q.astype(np.float64) * 255.0
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def box_sum_u8(img_u8: np.ndarray, r: int, mode: int) -> np.ndarray:
    """Exact integer (int64) window sums of a uint8 image."""
    a = np.asarray(img_u8, dtype=np.int64)

    def axis_sum(a, axis):
        n = a.shape[axis]
        a = np.moveaxis(a, axis, 0)
        if mode == BORDER_TRUNCATE:
            pad = np.zeros((r,) + a.shape[1:], dtype=np.int64)
            e = np.concatenate([pad, a, pad], axis=0)
        else:
            e = a[border_index(np.arange(-r, n + r), n, mode)]
        c = np.cumsum(e, axis=0)
        c = np.concatenate([np.zeros((1,) + c.shape[1:], dtype=np.int64), c], axis=0)
        return np.moveaxis(c[2 * r + 1: 2 * r + 1 + n] - c[0:n], 0, axis)

    return axis_sum(axis_sum(a, 1), 0)


# ---- the four element-wise launchers of path A ----------------------------------
# float32 restatements of guided_filter_d.cu:273-412 as driven by hMultiply / hCalcA / hCalcB /
# hLinearTransform (:927-1044).  "Guide-shaped" operands have the source's channel count or ONE channel
# (the reference's CN1 kernels broadcast them over the source channels).
def _fma32(a, b, c):
    """__fmaf_rn on float32 operands (the float64 product of two float32 values is exact)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def _bcast(g, like):
    g = np.asarray(g, np.float32)
    like = np.asarray(like)
    if like.ndim == 3 and g.ndim == 2:
        return np.repeat(g[:, :, None], like.shape[2], axis=2)
    return g


def pw_multiply(a, b):
    """hMultiply: c = __fmul_rn(a, b) (gMultiply :273-287 / gMultiplyCN1 :290-303)."""
    a = np.asarray(a, np.float32)
    return (a * _bcast(b, a)).astype(np.float32)


def pw_calc_a(pm, im, ipm, iim, eps):
    """hCalcA: a = fma(pm, -im, ipm) / fma(-im, im, iim + eps); eps joins iim BEFORE im^2 is subtracted
    (gCalcA :306-323, gCalcACN1 :326-346)."""
    pm = np.asarray(pm, np.float32)
    vim = _bcast(im, pm)
    viim = (_bcast(iim, pm) + np.float32(eps)).astype(np.float32)
    num = _fma32(pm, -vim, np.asarray(ipm, np.float32))
    den = _fma32(-vim, vim, viim)
    return (num / den).astype(np.float32)


def pw_calc_b(a, pm, im):
    """hCalcB: b = fma(a, -im, pm) (gCalcB :349-362).  The reference's 1-channel-guide variant truncates -im
    to int and reads im at the source index (gCalcBCN1 :371-372): that bug is NOT reproduced, the CN1 case
    broadcasts im like the other CN1 kernels."""
    a = np.asarray(a, np.float32)
    return _fma32(a, -_bcast(im, a), np.asarray(pm, np.float32))


def pw_linear_transform(src, a, b):
    """hLinearTransform: dst = fma(src, a, b), src guide-shaped (gLinearTransform :382-395 / CN1 :398-412)."""
    a = np.asarray(a, np.float32)
    return _fma32(_bcast(src, a), a, np.asarray(b, np.float32))


def class_run_steps(I, p, r: int, eps: float):
    """`GuidedFilter::run` step by step (guided_filter.cpp:28-66) with every intermediate plane, float32
    element-wise steps around float64-accurate TRUNCATE box means rounded to float32 (the reference's float32
    integral image is NOT reproduced: SURVEY fact 4)."""
    I = np.asarray(I, np.float32)
    p = np.asarray(p, np.float32)
    bm_ = lambda x: box_mean(x, r, BORDER_TRUNCATE, np.float64).astype(np.float32)
    s = {}
    s["pm"] = bm_(p)
    s["im"] = bm_(I)
    s["ip"] = pw_multiply(p, I)
    s["ii"] = pw_multiply(I, I)
    s["ipm"] = bm_(s["ip"])
    s["iim"] = bm_(s["ii"])
    s["a"] = pw_calc_a(s["pm"], s["im"], s["ipm"], s["iim"], eps)
    s["b"] = pw_calc_b(s["a"], s["pm"], s["im"])
    s["am"] = bm_(s["a"])
    s["bm"] = bm_(s["b"])
    s["q"] = pw_linear_transform(I, s["am"], s["bm"])
    return s


# ---- Integral/ module ---------------------------------------------------------
def integral_u8(img, dtype=np.int64):
    """Inclusive summed-area table of a uint8 image, W x H (no zero row/column): the layout of the
    reference's hIntegral (Integral/integral_d.cu:863-893; Integral/main.cpp:124 compares it with
    cv::integral's [1:, 1:]).  dtype=np.int32 wraps modulo 2^32 like the reference's int accumulators."""
    sat = np.cumsum(np.cumsum(img.astype(np.int64), axis=0), axis=1)
    if np.dtype(dtype) == np.int32:
        return (sat & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
    return sat.astype(dtype)


# ---- uint8 in / uint8 out -----------------------------------------------------
def u8_to_f32(img_u8):
    """Mat::convertTo(CV_32F, 1.0/255.0) (main.cpp:121-122,205-206): float(x * (1.0/255.0)), the product in double."""
    return (np.asarray(img_u8).astype(np.float64) * (1.0 / 255.0)).astype(np.float32)


def guided_filter_gray_u8(I_u8, p_u8, r, eps, border=BORDER_REFLECT101):
    """The reference demo's whole pipeline on 8-bit planes: convertTo float, the float32 CPU composition
    (main.cpp:236-252), convertTo(CV_8U, 255) (main.cpp:295-297)."""
    return to_u8(guided_filter_gray(u8_to_f32(I_u8), u8_to_f32(p_u8), r, eps, border, np.float32))


# ---- GaussianFilter/ module -----------------------------------------------------
_SMALL_GAUSS = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
                7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125]}


def gaussian_kernel_1d(radius: int, sigma: float) -> np.ndarray:
    """cv::getGaussianKernel(2r+1, sigma, CV_32F) restated -- the taps GaussianFilter/gaussian.cu:437 builds
    for every kernel it launches (and cv::GaussianBlur uses for the host result it compares with, :441).
    Returned as float32 (the values the filters multiply with)."""
    n = 2 * radius + 1
    if sigma <= 0 and n in _SMALL_GAUSS:
        return np.asarray(_SMALL_GAUSS[n], np.float32)
    if sigma <= 0:
        sigma = 0.3 * ((n - 1) * 0.5 - 1.0) + 0.8
    x = np.arange(n, dtype=np.float64) - radius
    t = np.exp(-0.5 / (sigma * sigma) * x * x)        # OpenCV >= 4.x: all in double, one rounding to float at the end
    return (t * (1.0 / t.sum())).astype(np.float32)


def gaussian_blur_gray(img: np.ndarray, radius: int, sigma: float) -> np.ndarray:
    """Separable Gaussian blur with REFLECT101 borders (reflectBorder, GaussianFilter/gaussian.h; the
    separable form is gGaussSplit / gGaussOptim, gaussian.cu:129-306), float64 accumulation of the
    float32 taps: the value cv::GaussianBlur and every reference kernel approximate in float32."""
    k = gaussian_kernel_1d(radius, sigma).astype(np.float64)
    a = np.asarray(img, np.float64)
    h, w = a.shape

    def refl(i, n):
        if n == 1:
            return np.zeros_like(i)
        i = np.abs(i)
        period = 2 * n - 2
        i = i % period
        return np.where(i >= n, period - i, i)
    xs = refl(np.arange(-radius, w + radius), w)
    ys = refl(np.arange(-radius, h + radius), h)
    ap = a[:, xs]
    hor = sum(k[j] * ap[:, j:j + w] for j in range(2 * radius + 1))
    hp = hor[ys, :]
    return sum(k[j] * hp[j:j + h, :] for j in range(2 * radius + 1))
