import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_kat_full():
    """The reference's r=7 / eps=0.3 4K KAT (main.cpp:193-252).  Inputs are rebuilt from the
    committed decoded gray planes with cv2.resize; the planes' hashes are pinned, so a cv2
    whose resize drifted makes the full-frame KAT skip (the crops in kat_crops.npz still run)."""
    import hashlib
    import json
    cv2 = pytest.importorskip("cv2")
    meta = json.load(open(os.path.join(GOLDEN, "kat_meta.json")))
    src = cv2.imread(os.path.join(GOLDEN, "adobe_src_gray_u8.png"), cv2.IMREAD_UNCHANGED)
    gui = cv2.imread(os.path.join(GOLDEN, "adobe_guide_gray_u8.png"), cv2.IMREAD_UNCHANGED)
    P = cv2.resize(src.astype(np.float32) * np.float32(1.0 / 255.0), (meta["width"], meta["height"]))
    I = cv2.resize(gui.astype(np.float32) * np.float32(1.0 / 255.0), (meta["width"], meta["height"]))
    if (hashlib.sha256(P.tobytes()).hexdigest() != meta["sha256_P_f32"]
            or hashlib.sha256(I.tobytes()).hexdigest() != meta["sha256_I_f32"]):
        pytest.skip("cv2.resize output differs from the pinned planes (cv2 %s)" % cv2.__version__)
    gold = cv2.imread(os.path.join(GOLDEN, "adobe_image_4_myres.png"), cv2.IMREAD_UNCHANGED)
    other = np.load(os.path.join(GOLDEN, "kat_other_goldens.npz"))
    cures = gold.copy().reshape(-1)
    cures[other["cures_idx"]] = other["cures_val"]
    cvres = gold.copy().reshape(-1)
    cvres[other["cvres_idx"]] = other["cvres_val"]
    return dict(I=I, P=P, gold=gold, cures=cures.reshape(gold.shape), cvres=cvres.reshape(gold.shape),
                r=meta["r"], eps=meta["eps"])


def load_kat_crops():
    z = np.load(os.path.join(GOLDEN, "kat_crops.npz"))
    names = sorted({k.rsplit("_", 1)[0] for k in z.files})
    out = []
    for n in names:
        out.append(dict(name=n, P=z[n + "_P"], I=z[n + "_I"], gold=z[n + "_gold"], cures=z[n + "_cures"],
                        off=tuple(int(v) for v in z[n + "_off"])))
    return out


def synth_pair(h, w, seed=0, kind="noise"):
    """Synthetic I, p in [0,1] float32 (SURVEY 8(d) config 2 inputs)."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        I = rng.random((h, w), dtype=np.float32)
        p = np.random.default_rng(seed + 1).random((h, w), dtype=np.float32)
    else:  # structured: low-frequency sinusoid + step edges + small noise (var ~ 0 regions)
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        I = 0.5 + 0.25 * np.sin(xx / 37.0) * np.cos(yy / 53.0)
        I += 0.2 * ((xx // 97 + yy // 61) % 2)
        I = np.clip(I + rng.normal(0, 0.02, (h, w)).astype(np.float32), 0, 1).astype(np.float32)
        p = np.clip(I * 0.8 + 0.1 + rng.normal(0, 0.05, (h, w)).astype(np.float32), 0, 1).astype(np.float32)
    return I, p


@pytest.fixture
def knob():
    """Sets launch-path options through the C ABI (gf_set_option) for one test and restores the defaults:
    knob(be, "GF_S8_HB", 40).  (The library does not read the environment per launch.)"""
    done = []

    def _set(be, name, value):
        be.api.set_option(name, value)
        done.append((be, name))
    yield _set
    for be, name in done:
        be.api.set_option(name, -1)
