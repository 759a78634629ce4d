"""Generates tests/golden/gauss_cv2.npz: cv2.GaussianBlur outputs (the host result the reference's
GaussianFilter/gaussian.cu:441 compares every kernel with) on seeded float32 images, plus
cv2.getGaussianKernel taps.  Run in the build container (cv2 is not needed at test time)."""
import os

import cv2
import numpy as np

out = {}
rng = np.random.default_rng(2024)
cases = [(37, 53, 1, 0.5), (64, 200, 3, 1.0), (90, 310, 8, 2.5), (50, 70, 2, 0.0), (120, 96, 5, 0.0), (40, 300, 16, 5.0), (9, 12, 4, 1.5)]
for i, (h, w, r, s) in enumerate(cases):
    img = rng.random((h, w), dtype=np.float32)
    out[f"img{i}"] = img
    out[f"par{i}"] = np.asarray([r, s], np.float64)
    out[f"blur{i}"] = cv2.GaussianBlur(img, (2 * r + 1, 2 * r + 1), s, sigmaY=s, borderType=cv2.BORDER_REFLECT_101)
    out[f"taps{i}"] = cv2.getGaussianKernel(2 * r + 1, s, cv2.CV_32F).ravel()
out["n"] = np.asarray([len(cases)])
np.savez_compressed(os.path.join(os.path.dirname(__file__), "gauss_cv2.npz"), **out)
print("cv2", cv2.__version__, "cases", len(cases))
