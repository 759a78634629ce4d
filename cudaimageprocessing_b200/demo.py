"""`cuda_guided_filter`-compatible demo driver: the reference's cudaSmallGuidedDemo
(GuidedFilter/main.cpp:178-312) on the B200 library.

    python -m cudaimageprocessing_b200.demo [radius] [eps] [nrepeats] [src_path] [guided_path]

Same positional arguments, defaults and steps as the reference executable (main.cpp:181-190): both images
are read as 8-bit gray (a missing guide = 3x3 median of the source, :199-203), converted to float32 / 255
(:205-206), resized to 3840x2160 (:207-211), filtered `nrepeats` times after 100 warm-up calls (:253-266)
with A and B written, timed on the device, and the result is saved as `<src>_cures.png` through the same
saturating x255 conversion (:295-305).  The reference's host arms (`_cvres.png` from OpenCV-contrib's
ximgproc, `_myres.png` from its CPU composition) are not reproduced: the library has no CPU path.
`GuidedFilter/run.py` works unchanged with a one-line `build/cuda_guided_filter` wrapper:
    #!/bin/sh
    exec python -m cudaimageprocessing_b200.demo "$@"
"""
from __future__ import annotations

import ctypes
import os
import sys


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    radius = int(argv[0]) if len(argv) > 0 else 1
    eps = float(argv[1]) if len(argv) > 1 else 0.3
    nrepeats = int(argv[2]) if len(argv) > 2 else 1
    src_path = argv[3] if len(argv) > 3 else "../data/adobe_image_4.jpg"
    guided_path = argv[4] if len(argv) > 4 else "../data/adobe_gt_4.jpg"
    try:
        import cv2
    except ImportError:
        print("this demo reads and writes images with OpenCV's Python module (cv2), which is not installed", file=sys.stderr)
        return 2
    import numpy as np
    import torch

    from . import api as get_api

    if not torch.cuda.is_available():
        print("no CUDA device: the library has no CPU path", file=sys.stderr)
        return 2
    h_src = cv2.imread(src_path, cv2.IMREAD_GRAYSCALE)
    if h_src is None:
        print(f"Can not read source image from: {src_path}")
        return 0                                           # the reference returns without an error code too (main.cpp:194-198)
    h_guided = cv2.imread(guided_path, cv2.IMREAD_GRAYSCALE)
    if h_guided is None:
        print("Guided image is missing. We use median-filtered image as guided-image")
        h_guided = cv2.medianBlur(h_src, 3)
    width, height = 3840, 2160
    src = cv2.resize(h_src.astype(np.float32) * np.float32(1.0 / 255.0), (width, height))
    guided = cv2.resize(h_guided.astype(np.float32) * np.float32(1.0 / 255.0), (width, height))

    api = get_api()
    d_src, d_guided = torch.from_numpy(src).cuda(), torch.from_numpy(guided).cuda()
    d_dst, d_A, d_B = (torch.empty_like(d_src) for _ in range(3))
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    def run():
        api.call("gf_guided_gray", d_guided.data_ptr(), d_src.data_ptr(), d_dst.data_ptr(), d_A.data_ptr(), d_B.data_ptr(),
                 width, height, 0, 0, 0, 0, radius, eps, 0, sp)
    for _ in range(100):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(nrepeats):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"Time cost of CUDA guided filter: {e0.elapsed_time(e1) / max(nrepeats, 1):f}ms")
    # convertTo(CV_8U, 255.0) (main.cpp:296): OpenCV scales a float32 Mat in float32, rounds half to even, saturates
    out = np.clip(np.rint(d_dst.cpu().numpy() * np.float32(255.0)), 0, 255).astype(np.uint8)
    cures_path = os.path.splitext(src_path)[0] + "_cures.png"
    cv2.imwrite(cures_path, out)
    print(f"kernel: {api.last_kernel()}  result: {cures_path}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
