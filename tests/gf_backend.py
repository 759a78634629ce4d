"""Two ways to drive the same C ABI from the tests:

* CudaBackend -- the product library (libgf_b200.so) on a real GPU, device memory via torch.
* EmuBackend  -- tests/emu/libgf_emu.so, the same kernels under the test-only SIMT emulator
                 on host memory (numpy).  Only for kernel-logic tests without a GPU.

Both expose the same helper methods and return numpy arrays, so parity cases are written once.
"""
from __future__ import annotations

import ctypes

import numpy as np

from cudaimageprocessing_b200._capi import GfApi


class _Base:
    api: GfApi

    # --- memory helpers, overridden ---
    def up(self, a: np.ndarray):
        raise NotImplementedError

    def empty(self, shape):
        raise NotImplementedError

    def ptr(self, buf) -> int:
        raise NotImplementedError

    def down(self, buf) -> np.ndarray:
        raise NotImplementedError

    def sync(self):
        pass

    # --- calls ---
    def guided_gray(self, I, p, r, eps, border, want_ab=False, pad=0):
        """pad > 0 gives every plane a row stride of width+pad floats (pitched layout)."""
        h, w = I.shape
        s = w + pad

        def pitched(a):
            b = np.zeros((h, s), np.float32)
            b[:, :w] = a
            return b
        dI, dp = self.up(pitched(I)), self.up(pitched(p))
        dq = self.empty((h, s))
        dA = self.empty((h, s)) if want_ab else None
        dB = self.empty((h, s)) if want_ab else None
        self.api.call("gf_guided_gray", self.ptr(dI), self.ptr(dp), self.ptr(dq), self.ptr(dA) if want_ab else None,
                      self.ptr(dB) if want_ab else None, w, h, s, s, s, s, r, eps, border, None)
        self.sync()
        q = self.down(dq)[:, :w]
        if want_ab:
            return q, self.down(dA)[:, :w], self.down(dB)[:, :w]
        return q

    def guided_color(self, I3, p, r, eps, border):
        h, w, _ = I3.shape
        pc = 1 if p.ndim == 2 else p.shape[2]
        dI, dp = self.up(I3), self.up(p)
        dq = self.empty(p.shape)
        self.api.call("gf_guided_color", self.ptr(dI), self.ptr(dp), self.ptr(dq), w, h, pc, 0, 0, 0, r, eps, border, None)
        self.sync()
        return self.down(dq)

    def class_run(self, I, p, r, eps, border=1):
        h, w = I.shape[:2]
        gch = 1 if I.ndim == 2 else I.shape[2]
        sch = 1 if p.ndim == 2 else p.shape[2]
        hnd = ctypes.c_void_p()
        self.api.call("gf_create", ctypes.addressof(hnd), w, h, gch, sch)
        try:
            dI, dp = self.up(I), self.up(p)
            dq = self.empty(p.shape)
            self.api.call("gf_run", hnd, self.ptr(dI), self.ptr(dp), self.ptr(dq), r, eps, border, 0, 0, 0, None)
            self.sync()
            return self.down(dq)
        finally:
            self.api.call("gf_destroy", hnd)

    def batch(self, I, p, r, eps, border):
        n, h, w = p.shape
        gch = 1 if I.ndim == 3 else I.shape[3]
        dI, dp = self.up(I), self.up(p)
        dq = self.empty(p.shape)
        self.api.call("gf_guided_batch", self.ptr(dI), self.ptr(dp), self.ptr(dq), n, w, h, gch, 0, 0, 0, 0, 0, 0,
                      r, eps, border, None)
        self.sync()
        return self.down(dq)

    def strip(self, I_buf, p_buf, width, global_h, buf_y0, out_y0, out_rows, r, eps, border):
        dI, dp = self.up(I_buf), self.up(p_buf)
        dq = self.empty((out_rows, width))
        self.api.call("gf_guided_gray_strip", self.ptr(dI), self.ptr(dp), self.ptr(dq), width, global_h, buf_y0,
                      I_buf.shape[0], out_y0, out_rows, 0, 0, 0, r, eps, border, None)
        self.sync()
        return self.down(dq)

    def run_strips(self, I, p, world, r, eps, border):
        """gf_run_strips with `world` ranks emulated in ONE process: every rank's strip buffers hold only its own
        rows (halo rows poisoned), the call pulls the halos out of the neighbours' buffers."""
        from cudaimageprocessing_b200._capi import StripPeer
        H, w = I.shape
        ranks = []
        for g in range(world):
            y0, y1 = H * g // world, H * (g + 1) // world
            top, bot = ctypes.c_int(), ctypes.c_int()
            self.api.call("gf_strip_layout", H, y0, y1 - y0, r, ctypes.addressof(top), ctypes.addressof(bot))
            bufs = []
            for a in (I, p):
                b = np.full((top.value + y1 - y0 + bot.value, w), np.nan, np.float32)
                b[top.value:top.value + y1 - y0] = a[y0:y1]
                bufs.append(self.up(b))
            ranks.append((y0, y1, top.value, bufs))
        out = np.empty_like(I)
        for g, (y0, y1, top, bufs) in enumerate(ranks):
            def peer(k):
                py0, py1, ptop, pb = ranks[k]
                return StripPeer(self.ptr(pb[0]), self.ptr(pb[1]), w, w, ptop, py1 - py0)
            up = peer(g - 1) if g > 0 else None
            dn = peer(g + 1) if g < world - 1 else None
            dq = self.empty((y1 - y0, w))
            self.api.call("gf_run_strips", self.ptr(bufs[0]), self.ptr(bufs[1]), self.ptr(dq), w, H, y0, y1 - y0, w, w, w, r, eps, border,
                          ctypes.byref(up) if up is not None else None, ctypes.byref(dn) if dn is not None else None, None)
            self.sync()
            out[y0:y1] = self.down(dq)
        return out

    def box(self, a, r, border, inplace=False):
        h, w = a.shape[:2]
        c = 1 if a.ndim == 2 else a.shape[2]
        d = self.up(a)
        o = d if inplace else self.empty(a.shape)
        self.api.call("gf_box_filter", self.ptr(d), self.ptr(o), w, h, c, 0, 0, r, border, None)
        self.sync()
        return self.down(o)

    # --- the four element-wise launchers of path A (hMultiply, hCalcA, hCalcB, hLinearTransform) ---
    @staticmethod
    def _ch(a):
        return 1 if np.ndim(a) == 2 else np.shape(a)[2]

    def multiply(self, a, b):
        h, w = a.shape[:2]
        da, db = self.up(a), self.up(b)
        o = self.empty(a.shape)
        self.api.call("gf_multiply", self.ptr(da), self.ptr(db), self.ptr(o), w, h, self._ch(a), self._ch(b), 0, 0, None)
        self.sync()
        return self.down(o)

    def calc_a(self, pm, im, ipm, iim, eps):
        h, w = pm.shape[:2]
        d = [self.up(x) for x in (pm, im, ipm, iim)]
        o = self.empty(pm.shape)
        self.api.call("gf_calc_a", self.ptr(o), self.ptr(d[0]), self.ptr(d[1]), self.ptr(d[2]), self.ptr(d[3]), w, h,
                      self._ch(pm), self._ch(im), 0, 0, eps, None)
        self.sync()
        return self.down(o)

    def calc_b(self, a, pm, im):
        h, w = a.shape[:2]
        d = [self.up(x) for x in (a, pm, im)]
        o = self.empty(a.shape)
        self.api.call("gf_calc_b", self.ptr(o), self.ptr(d[0]), self.ptr(d[1]), self.ptr(d[2]), w, h, self._ch(a), self._ch(im),
                      0, 0, None)
        self.sync()
        return self.down(o)

    def linear_transform(self, src, a, b):
        h, w = a.shape[:2]
        d = [self.up(x) for x in (src, a, b)]
        o = self.empty(a.shape)
        self.api.call("gf_linear_transform", self.ptr(d[0]), self.ptr(o), self.ptr(d[1]), self.ptr(d[2]), w, h, self._ch(a),
                      self._ch(src), 0, 0, None)
        self.sync()
        return self.down(o)

    def class_run_by_launchers(self, I, p, r, eps):
        """The eleven launcher calls of GuidedFilter::run (guided_filter.cpp:28-66), in its order, through the C ABI."""
        pm = self.box(p, r, 1)
        im = self.box(I, r, 1)
        ipm = self.box(self.multiply(p, I), r, 1)
        iim = self.box(self.multiply(I, I), r, 1)
        a = self.calc_a(pm, im, ipm, iim, eps)
        b = self.calc_b(a, pm, im)
        am = self.box(a, r, 1)
        bm = self.box(b, r, 1)
        return {"pm": pm, "im": im, "ipm": ipm, "iim": iim, "a": a, "b": b, "am": am, "bm": bm,
                "q": self.linear_transform(I, am, bm)}


class EmuBackend(_Base):
    name = "emu"

    def __init__(self):
        from emu.build_emu import build_emu
        self.api = GfApi(ctypes.CDLL(build_emu()))

    @staticmethod
    def _aligned(shape):
        """64-byte aligned float32 array (what cudaMalloc gives the real kernels)."""
        n = int(np.prod(shape))
        raw = np.empty(n * 4 + 64, np.uint8)
        off = (-raw.ctypes.data) % 64
        return raw[off:off + n * 4].view(np.float32).reshape(shape)

    def up(self, a):
        b = self._aligned(np.shape(a))
        b[...] = a
        return b

    def empty(self, shape):
        b = self._aligned(shape)
        b[...] = np.nan
        return b

    def ptr(self, buf):
        return None if buf is None else buf.ctypes.data

    def down(self, buf):
        return buf


class CudaBackend(_Base):
    name = "cuda"

    def __init__(self):
        import torch
        import cudaimageprocessing_b200 as pkg
        self.torch = torch
        self.api = pkg.api()

    def up(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()

    def empty(self, shape):
        return self.torch.full(tuple(shape), float("nan"), dtype=self.torch.float32, device="cuda")

    def ptr(self, buf):
        return None if buf is None else buf.data_ptr()

    def down(self, buf):
        return buf.cpu().numpy()

    def sync(self):
        self.torch.cuda.synchronize()
