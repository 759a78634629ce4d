"""hMultiply, hCalcA, hCalcB, hLinearTransform (guided_filter_d.cu:273-412, 927-1044)
through the C ABI, under the emulator (CPU suite) and on the GPU, against the oracle's float32 restatements.
What the cases pin: the one-channel-guide broadcast of the CN1 kernels (:288-303, :327-346, :398-412), eps
joining iim BEFORE im^2 is subtracted (:318, :336), the corrected gCalcBCN1 (:371-372), and that the eleven
launcher calls of GuidedFilter::run (guided_filter.cpp:28-66) chained by hand give what gf_run gives."""
import numpy as np
import pytest

from oracle import gf_oracle as O

ULP = 2.0 ** -22      # two float32 ulps at 1.0: fma contraction vs the float64-emulated fma


def _planes(shape_s, shape_g, seed):
    rng = np.random.default_rng(seed)
    return rng.random(shape_s, dtype=np.float32), rng.random(shape_g, dtype=np.float32)


SHAPES = [((19, 37), (19, 37)), ((19, 37, 3), (19, 37, 3)), ((19, 37, 3), (19, 37)), ((3, 300, 3), (3, 300))]


def _check_launchers(be):
    for ss, sg in SHAPES:
        s, g = _planes(ss, sg, 3)
        s2, g2 = _planes(ss, sg, 4)
        # hMultiply
        assert np.abs(be.multiply(s, g) - O.pw_multiply(s, g)).max() <= ULP
        # hCalcA: eps = 0.3 makes the order of "+ eps" and "- im^2" visible in the last bits
        pm, im = s, g
        ipm, iim = s2, (g2 * g2 + g * g).astype(np.float32)
        ref = O.pw_calc_a(pm, im, ipm, iim, 0.3)
        got = be.calc_a(pm, im, ipm, iim, 0.3)
        assert np.abs(got - ref).max() <= 4 * ULP * max(1.0, np.abs(ref).max())
        # hCalcB (CN1: NOT the reference's truncated-to-int variant)
        a = (s2 * 3 - 1.5).astype(np.float32)
        ref = O.pw_calc_b(a, pm, im)
        assert np.abs(be.calc_b(a, pm, im) - ref).max() <= 2 * ULP
        if len(ss) == 3 and len(sg) == 2:
            buggy = (a * (-im).astype(np.int32)[:, :, None] + pm)        # what gCalcBCN1 computes: int(-im) = 0
            assert np.abs(ref - buggy).max() > 0.1
        # hLinearTransform: src is guide-shaped
        ref = O.pw_linear_transform(g, a, s)
        assert np.abs(be.linear_transform(g, a, s) - ref).max() <= 4 * ULP


def _check_chain(be, shape_i, shape_p, r=3, eps=0.05):
    rng = np.random.default_rng(9)
    I = rng.random(shape_i, dtype=np.float32)
    p = rng.random(shape_p, dtype=np.float32)
    got = be.class_run_by_launchers(I, p, r, eps)
    ref = O.class_run_steps(I, p, r, eps)
    for k in ("pm", "im", "ipm", "iim"):
        assert np.abs(got[k] - ref[k]).max() <= 1e-6, k
    for k in ("a", "b", "am", "bm", "q"):
        assert np.abs(got[k] - ref[k]).max() <= 1e-4, k      # a = cov / (var + eps) amplifies 1e-7 by 1 / 0.05
    q_run = be.class_run(I, p, r, eps)                       # the fused gf_run on the same planes
    assert np.abs(got["q"] - q_run).max() <= 1e-5
    assert np.abs(q_run - O.guided_filter_class_run(I, p, r, eps)).max() <= 1e-4


def test_oracle_pointwise_matches_float64():
    """The float32 restatements agree with the float64 composition (no transcription slip)."""
    rng = np.random.default_rng(1)
    I = rng.random((24, 31), dtype=np.float32)
    p3 = rng.random((24, 31, 3), dtype=np.float32)
    s = O.class_run_steps(I, p3, 4, 0.02)
    assert np.abs(s["q"] - O.guided_filter_class_run(I, p3, 4, 0.02)).max() <= 1e-5


def test_launchers_emulated():
    from gf_backend import EmuBackend
    _check_launchers(EmuBackend())


@pytest.mark.parametrize("shape_i,shape_p", [((26, 40), (26, 40)), ((26, 40, 3), (26, 40, 3)), ((26, 40), (26, 40, 3))])
def test_chain_emulated(shape_i, shape_p):
    from gf_backend import EmuBackend
    _check_chain(EmuBackend(), shape_i, shape_p)


@pytest.mark.gpu
def test_launchers_gpu():
    from gf_backend import CudaBackend
    _check_launchers(CudaBackend())


@pytest.mark.gpu
@pytest.mark.parametrize("shape_i,shape_p", [((270, 480), (270, 480)), ((135, 240, 3), (135, 240, 3)), ((135, 240), (135, 240, 3))])
def test_chain_gpu(shape_i, shape_p):
    from gf_backend import CudaBackend
    _check_chain(CudaBackend(), shape_i, shape_p, r=7, eps=0.3)
