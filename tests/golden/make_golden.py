"""Regenerates tests/golden/* from the reference's bundled data.  Run in the build container
only (`python tests/golden/make_golden.py`): it reads /root/reference, which does not exist on
the GPU box.  The KAT it pins is `cudaSmallGuidedDemo` (GuidedFilter/main.cpp:178-312) as run
by GuidedFilter/run.py:5-6 (last iteration: r=7, eps=0.3), whose CPU result is
data/adobe_image_4_myres.png and whose GPU result is data/adobe_image_4_cures.png.
"""
import hashlib
import json
import os
import shutil
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gf_oracle as O  # noqa: E402

REF = "/root/reference/GuidedFilter/data/"
R, EPS, W, H = 7, 0.3, 3840, 2160


def main():
    src = cv2.imread(REF + "adobe_image_4.jpg", cv2.IMREAD_GRAYSCALE)       # main.cpp:193
    gui = cv2.imread(REF + "adobe_gt_4.jpg", cv2.IMREAD_GRAYSCALE)          # main.cpp:199
    cv2.imwrite(os.path.join(HERE, "adobe_src_gray_u8.png"), src)            # pins the JPEG decode
    cv2.imwrite(os.path.join(HERE, "adobe_guide_gray_u8.png"), gui)
    P = cv2.resize(src.astype(np.float32) * np.float32(1.0 / 255.0), (W, H))  # main.cpp:205-211
    I = cv2.resize(gui.astype(np.float32) * np.float32(1.0 / 255.0), (W, H))
    myres = cv2.imread(REF + "adobe_image_4_myres.png", cv2.IMREAD_UNCHANGED)
    cures = cv2.imread(REF + "adobe_image_4_cures.png", cv2.IMREAD_UNCHANGED)
    cvres = cv2.imread(REF + "adobe_image_4_cvres.png", cv2.IMREAD_UNCHANGED)
    shutil.copyfile(REF + "adobe_image_4_myres.png", os.path.join(HERE, "adobe_image_4_myres.png"))

    # the other two goldens as sparse differences against _myres (34 px and 11761 px)
    def sparse(a):
        idx = np.flatnonzero(a != myres)
        return idx.astype(np.int32), a.reshape(-1)[idx]
    ci, cv_ = sparse(cures)
    xi, xv = sparse(cvres)
    np.savez_compressed(os.path.join(HERE, "kat_other_goldens.npz"), cures_idx=ci, cures_val=cv_,
                        cvres_idx=xi, cvres_val=xv)

    # the oracle must reproduce the KAT before anything is written
    q = O.guided_filter_gray(I, P, R, EPS, O.BORDER_REFLECT101, np.float32)
    nd = int(np.count_nonzero(O.to_u8(q) != myres))
    assert nd == 0, f"oracle does not reproduce _myres.png: {nd} px differ"

    meta = {
        "r": R, "eps": EPS, "width": W, "height": H, "cv2": cv2.__version__,
        "sha256_P_f32": hashlib.sha256(P.tobytes()).hexdigest(),
        "sha256_I_f32": hashlib.sha256(I.tobytes()).hexdigest(),
        "oracle_f32_vs_myres_diff_px": nd,
        "cures_vs_myres_diff_px": int(ci.size), "cvres_vs_myres_diff_px": int(xi.size),
    }

    # self-contained crops (no cv2 needed to replay): input window with a 2r apron where the
    # image continues, none where the image border is (so the border rule is exercised).
    n, ap = 96, 2 * R
    crops = {}
    for name, (y0, x0) in {"tl": (0, 0), "br": (H - n, W - n), "top": (0, 1900), "left": (1000, 0),
                           "mid": (700, 2300), "mid2": (1500, 900)}.items():
        ya, yb = max(0, y0 - ap), min(H, y0 + n + ap)
        xa, xb = max(0, x0 - ap), min(W, x0 + n + ap)
        crops[name + "_P"] = P[ya:yb, xa:xb].copy()
        crops[name + "_I"] = I[ya:yb, xa:xb].copy()
        crops[name + "_gold"] = myres[y0:y0 + n, x0:x0 + n].copy()
        crops[name + "_cures"] = cures[y0:y0 + n, x0:x0 + n].copy()
        crops[name + "_off"] = np.array([y0 - ya, x0 - xa], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "kat_crops.npz"), **crops)
    with open(os.path.join(HERE, "kat_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
