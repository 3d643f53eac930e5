#!/usr/bin/env python
"""tools/emu_fuzz.py [seconds] [seed] — randomised differential run of the EMULATED convert+scale kernels (tests/cudaemu:
the product's kernel sources compiled for the host) against the oracle: random sizes (biased small and odd), methods,
all format pairs incl. the opt-in ones, chroma sitings, colorimetry, destination rectangles; where the reference build
(oracle/_ref) is present, half of the cases also compare the oracle with the reference's own converter, leaving out the
reference's two known defect classes.  TEST INFRASTRUCTURE.
Meant to be run under the sanitizer builds as well:
    B200_EMU_ASAN=1 LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:log_path=/tmp/asan python tools/emu_fuzz.py 300
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "cudaemu"))

from oracle import bindings as ob   # noqa: E402
import test_emu_kernels as T        # noqa: E402


ONLY = [k for k in os.environ.get("B200_FUZZ_KINDS", "").split(",") if k]      # e.g. B200_FUZZ_KINDS=422-420,rgb-yuv


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    import build_emu
    from gstreamer_b200 import _lib
    emu = C.CDLL(build_emu.build())
    for name, (res, args) in _lib._SIGS.items():
        if hasattr(emu, name):
            fn = getattr(emu, name)
            fn.restype, fn.argtypes = res, args
    t0, n, bad, kinds = time.time(), 0, 0, {}
    n_ref = n_defect = 0
    try:
        ob.ref()
        ref_ok = True
    except Exception:                               # noqa: BLE001  (oracle/_ref absent: the GPU box)
        ref_ok = False
    while time.time() - t0 < budget:
        big = rng.random() < 0.2
        hi = 300 if big else 70
        iw, ih, W, H = (int(v) for v in rng.integers(1, hi, 4))
        r = rng.random()
        if r < 0.15:
            W = iw
        elif r < 0.3:
            H = ih
        elif r < 0.4:
            W, H = max(1, iw // 2), max(1, ih // 2)
        method = int(rng.integers(0, 10))
        kind = str(rng.choice(ONLY or ["yuv-rgb", "yuv-rgb", "same", "cross", "rgb-yuv", "rgb-rgb", "422-rgb", "422-420", "4xx-420"]))
        if kind == "yuv-rgb":
            fi, fo = str(rng.choice(T.YUV)), str(rng.choice(T.RGB))
        elif kind == "same":
            fi = str(rng.choice(T.YUV))
            fo = fi if fi in ("NV12", "NV21") else str(rng.choice(["I420", "YV12"]))
        elif kind == "cross":
            fi = str(rng.choice(T.YUV))
            fo = str(rng.choice([f for f in T.YUV if not ((fi in ("I420", "YV12")) == (f in ("I420", "YV12")) and (fi == f or fi in ("I420", "YV12")))]))
        elif kind == "rgb-yuv":
            fi, fo = str(rng.choice(T.RGB)), str(rng.choice(T.YUV))
        elif kind == "rgb-rgb":
            fi, fo = str(rng.choice(T.RGB)), str(rng.choice(T.RGB))
        elif kind == "422-420":                     # capture -> encoder: table rows at an unchanged size, the chain otherwise
            fi, fo = str(rng.choice(["YUY2", "UYVY", "YVYU"])), str(rng.choice(T.YUV))
            if rng.random() < 0.3:
                W, H = iw, ih
        elif kind == "4xx-420":                     # planar 4:2:2 / 4:4:4 -> planar 4:2:0: plane-scaling rows
            fi, fo = str(rng.choice(["Y42B", "Y444"])), str(rng.choice(T.YUV))    # ... to NV12 / NV21: the chain
            if rng.random() < 0.3:
                W, H = iw, ih
        else:
            fi, fo = str(rng.choice(T.YUV_422_444)), str(rng.choice(T.RGB))
        site = int(rng.choice([1, 2, 4, 6]))
        out_site = int(rng.choice([1, 2, 4, 6])) if kind == "cross" else None
        if kind == "422-420" or (kind == "4xx-420" and fo in ("NV12", "NV21")):   # the output keeps the default site of its own (frame) size
            out_site = 2 if H > 576 else 1
        dest = None
        if rng.random() < 0.25:
            dw, dh = int(rng.integers(1, W + 1)), int(rng.integers(1, H + 1))
            dest = (int(rng.integers(0, W - dw + 1)), int(rng.integers(0, H - dh + 1)), dw, dh)
        frame = T.frame_for(fi, iw, ih, n)
        size = (iw, ih, W, H)
        generic = not (fi in T.RGB and fo in T.RGB) and kind not in ("same", "4xx-420")
        try:
            kw = dict(site=site if fi not in T.RGB else None, out_site=out_site, dest=dest)
            want = T.expected(fi, fo, size, method, frame, **kw)
            if ref_ok and rng.random() < 0.5:       # ... and the oracle itself against the reference build
                rkw = {}
                if fi not in T.RGB:
                    rkw["site"] = site
                if fi not in T.RGB and fo not in T.RGB:
                    d0 = ob.vcs_desc(iw, ih, W, H, method, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo], site=site)
                    rkw.update(matrix=d0.in_matrix, out_matrix=d0.in_matrix, out_site=site if out_site is None else out_site)
                if dest:
                    rkw.update(dest=dest, border_argb=0xff000000)
                rv = ob.RefVcs(iw, ih, W, H, method, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo], **rkw)
                ref_out = rv.convert(frame, np.full(want.size, 0x5A, dtype=np.uint8))
                rv.close()
                n_ref += 1
                if not np.array_equal(ref_out, want):
                    dw, dh = (dest[2], dest[3]) if dest else (W, H)
                    vfirst = ih != dh and (iw == dw or dw * ih > iw * dh)
                    if not (fi in T.RGB and fo in T.RGB) and (vfirst or (method == 0 or ih == 1) and dh > ih):
                        n_defect += 1               # the two reference defect classes, not reproduced
                    else:
                        bad += 1
                        print("ORACLE != REF", kind, fi, fo, size, "m", method, "site", site, out_site, "dest", dest,
                              int(np.count_nonzero(ref_out != want)), "of", want.size, flush=True)
            got = T.run(emu, fi, fo, size, method, frame, force_generic=False, **kw)
            if rng.random() < 0.3 and generic:      # the generic kernel on shapes the fast kernels would take
                got2 = T.run(emu, fi, fo, size, method, frame, force_generic=True, **kw)
                if not np.array_equal(got2, want):
                    bad += 1
                    print("MISMATCH (generic)", kind, fi, fo, size, "m", method, "site", site, out_site, "dest", dest, flush=True)
        except Exception as e:                      # noqa: BLE001
            print("ERROR", kind, fi, fo, size, "m", method, "dest", dest, repr(e)[:200], flush=True)
            bad += 1
            continue
        n += 1
        kinds[kind] = kinds.get(kind, 0) + 1
        if not np.array_equal(got, want):
            bad += 1
            print("MISMATCH", kind, fi, fo, size, "m", method, "site", site, out_site, "dest", dest,
                  int(np.count_nonzero(got != want)), "of", got.size, flush=True)
    print(f"emu_fuzz: {n} cases, {bad} problems, {kinds}; oracle vs reference build on {n_ref} of them "
          f"({n_defect} in the reference's two known defect classes)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
