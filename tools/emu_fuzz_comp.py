#!/usr/bin/env python
"""tools/emu_fuzz_comp.py [seconds] [seed] — randomised differential run of the compositor: the reference build
(oracle/_ref, when present) against the oracle, and the oracle against the EMULATED product kernels (tests/cudaemu).
Packed-RGB (4 byte orders) and 4:2:0 (4 formats, both ranges) canvases from 1x1 up, 0..40 pads that may lie anywhere
(far outside, covering the whole canvas, 1 pixel wide), every operator, alphas from 0 to 1 incl. values next to the
rounding points, every background.  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "cudaemu"))

from oracle import bindings as ob   # noqa: E402

RGB = ["RGBA", "BGRA", "ARGB", "ABGR"]
YUV = [2, 3, 23, 24]
ALPHAS = [0.0, 1.0, 0.5, 0.25, 0.999, 0.004, 1 / 255, 0.5 / 255, 254.5 / 255, 0.75, 0.1]


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    import build_emu
    from gstreamer_b200 import _lib
    emu = C.CDLL(build_emu.build())
    for name, (res, args) in _lib._SIGS.items():
        if hasattr(emu, name):
            fn = getattr(emu, name)
            fn.restype, fn.argtypes = res, args
    o = ob.oracle()
    try:
        r = ob.ref()
    except Exception:
        r = None
    t0, n_cases, bad = time.time(), 0, 0
    while time.time() - t0 < budget:
        yuv = rng.random() < 0.5
        big = rng.random() < 0.15
        W, H = (int(v) for v in rng.integers(1, 260 if big else 90, 2))
        bg = int(rng.integers(0, 4))
        n = int(rng.integers(0, 7)) if rng.random() < 0.85 else int(rng.integers(7, 41))
        opads = (ob.OraclePad * max(n, 1))()
        keep = []
        geo = []
        fmt = int(rng.choice(YUV)) if yuv else ob.FMT[str(rng.choice(RGB))]
        for i in range(n):
            k = rng.random()
            if k < 0.1:
                w, h = W + int(rng.integers(0, 40)), H + int(rng.integers(0, 40))       # covers the canvas
            elif k < 0.2:
                w, h = 1, int(rng.integers(1, 60))
            else:
                w, h = int(rng.integers(1, 120)), int(rng.integers(1, 90))
            x, y = int(rng.integers(-w - 5, W + 5)), int(rng.integers(-h - 5, H + 5))
            al = float(rng.choice(ALPHAS)) if rng.random() < 0.7 else float(rng.random())
            op = int(rng.integers(0, 3))
            if yuv:
                a = rng.integers(0, 256, o.oracle_compositor_yuv_size(fmt, w, h), dtype=np.uint8)
                stride = 0
            else:
                a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
                if rng.random() < 0.3:
                    a[..., 0 if fmt in (ob.FMT["ARGB"], ob.FMT["ABGR"]) else 3] = rng.choice([0, 255, 1, 254])
                stride = w * 4
            keep.append(a)
            geo.append((w, h, x, y, al, op))
            opads[i].data, opads[i].width, opads[i].height, opads[i].stride = a.ctypes.data, w, h, stride
            opads[i].xpos, opads[i].ypos, opads[i].alpha, opads[i].op = x, y, al, op
        desc = ("yuv" if yuv else "rgb", fmt, W, H, bg, n)
        hc = C.c_void_p()
        if emu.b200_comp_create(fmt, W, H, 0, C.byref(hc)) != 0:
            print("CREATE", desc, flush=True)
            bad += 1
            continue
        if yuv:
            rg = int(rng.integers(0, 2))
            sz = o.oracle_compositor_yuv_size(fmt, W, H)
            want = np.full(sz, 0x77, dtype=np.uint8)
            o.oracle_compositor_yuv(fmt, want.ctypes.data, W, H, bg, rg, opads, n)
            if r:
                w2 = np.full(sz, 0x77, dtype=np.uint8)
                r.ref_compositor_yuv(fmt, w2.ctypes.data, W, H, bg, rg, opads, n)
                if not np.array_equal(w2, want):
                    print("ORACLE != REF", desc, rg, geo, flush=True)
                    bad += 1
            oi = _lib.VideoInfoC()
            emu.b200_video_info_set_format(C.byref(oi), fmt, W, H)
            oi.color_range = 2 if rg else 1
            pads = (_lib.CompPadYuvC * max(n, 1))()
            for i, (w, h, x, y, al, op) in enumerate(geo):
                emu.b200_video_info_set_format(C.byref(pads[i].info), fmt, w, h)
                pads[i].data, pads[i].xpos, pads[i].ypos, pads[i].alpha, pads[i].op = keep[i].ctypes.data, x, y, al, op
            got = np.full(sz, 0x77, dtype=np.uint8)
            st = emu.b200_comp_blend_yuv(hc, got.ctypes.data, C.byref(oi), bg, pads, n, None)
        else:
            want = np.full((H, W, 4), 0x77, dtype=np.uint8)
            o.oracle_compositor(fmt, want.ctypes.data, W, H, W * 4, bg, opads, n)
            if r:
                w2 = np.full((H, W, 4), 0x77, dtype=np.uint8)
                r.ref_compositor(fmt, w2.ctypes.data, W, H, W * 4, bg, opads, n)
                if not np.array_equal(w2, want):
                    print("ORACLE != REF", desc, geo, flush=True)
                    bad += 1
            pads = (_lib.CompPadC * max(n, 1))()
            for i, (w, h, x, y, al, op) in enumerate(geo):
                pads[i].data, pads[i].width, pads[i].height, pads[i].stride = keep[i].ctypes.data, w, h, w * 4
                pads[i].xpos, pads[i].ypos, pads[i].alpha, pads[i].op = x, y, al, op
            got = np.full((H, W, 4), 0x77, dtype=np.uint8)
            st = emu.b200_comp_blend(hc, got.ctypes.data, W * 4, bg, pads, n, None)
        emu.b200_comp_destroy(hc)
        if st != 0 or not np.array_equal(got, want):
            print("EMU != ORACLE", st, desc, geo, flush=True)
            bad += 1
        n_cases += 1
    print(f"{n_cases} layouts, {bad} mismatches, seed {seed}, reference {'present' if r else 'absent'}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
