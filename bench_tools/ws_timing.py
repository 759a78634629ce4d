"""Per-warp durations of one gf_ws launch (needs a -DGF_WS_TIMING build: GF_LIB_PATH=.../libgf_timing.so).
    GF_LIB_PATH=$PWD/cudaimageprocessing_b200/libgf_timing.so GF_WS=1 python bench_tools/ws_timing.py 3840 2160 8"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import cudaimageprocessing_b200 as pkg  # noqa: E402


def main():
    w, h, r = (int(x) for x in sys.argv[1:4])
    api = pkg.api()
    g = torch.Generator(device="cuda").manual_seed(0)
    sets = [(torch.rand((h, w), device="cuda", generator=g), torch.rand((h, w), device="cuda", generator=g), torch.empty((h, w), device="cuda"))
            for _ in range(4)]
    NS = 4
    dbg = torch.zeros((4096, 3 * NS, 4), dtype=torch.int64, device="cuda")
    p = dbg.data_ptr()
    api.set_option("GF_WS_DBG_LO", p & 0x7FFFFFFF)
    api.set_option("GF_WS_DBG_HI", p >> 31)
    for i in range(8):
        a, b, c = sets[i % 4]
        dbg.zero_()
        api.call("gf_guided_gray", a.data_ptr(), b.data_ptr(), c.data_ptr(), None, None, w, h, 0, 0, 0, 0, r, 1e-2, 0, None)
        torch.cuda.synchronize()
    d = dbg.cpu().numpy()
    used = d[:, :, 1] > 0
    ncta = int(used.any(axis=1).sum())
    t0 = d[:, :, 0][used].min()
    rows = []
    for cta in range(ncta):
        u = used[cta]
        if not u.any():
            continue
        st = d[cta][u]
        strip, band = int(st[0, 2]) >> 16, int(st[0, 2]) & 0xffff
        dur = (st[:, 1] - st[:, 0])
        rows.append({"cta": cta, "strip": strip, "band": band, "start": int(st[:, 0].min() - t0), "end": int(st[:, 1].max() - t0),
                     "warp_clk": [int(x) for x in dur], "L": [int(x) >> 16 for x in st[:, 3]], "s1rows": [int(x) & 0xffff for x in st[:, 3]]})
    ends = np.array([r_["end"] for r_ in rows])
    print(json.dumps({"kernel": api.last_kernel(), "ctas": ncta, "end_min": int(ends.min()), "end_mean": float(ends.mean()), "end_max": int(ends.max())}))
    for r_ in sorted(rows, key=lambda x: -x["end"])[:12]:
        print(json.dumps(r_))
    print("...")
    for r_ in sorted(rows, key=lambda x: x["end"])[:6]:
        print(json.dumps(r_))
    # by class
    ns = max(r_["strip"] for r_ in rows) + 1
    nb = max(r_["band"] for r_ in rows) + 1
    for name, sel in (("edge strips", lambda r_: r_["strip"] in (0, ns - 1)), ("interior strips", lambda r_: 0 < r_["strip"] < ns - 1),
                      ("first/last band", lambda r_: r_["band"] in (0,) ), ):
        e = [r_["end"] - r_["start"] for r_ in rows if sel(r_)]
        if e:
            print(name, "n", len(e), "mean", int(np.mean(e)), "max", max(e), "min", min(e))
    # producer vs consumer durations
    s1 = [max(r_["warp_clk"][:NS]) for r_ in rows]
    print("producer max per cta: mean", int(np.mean(s1)), "max", max(s1))


if __name__ == "__main__":
    main()
