#!/usr/bin/env python
"""Generates tests/golden/*.npz from the reference's OWN code (oracle/_ref/libgstref.so, built by
oracle/Makefile from /root/reference sources).  Run in the build container only:

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by oracle.bindings.nv12_random_frame (deterministic LCG), so
only seeds, parameters and the reference's output bytes (or their SHA-256 for large cases) are
stored.  The reference's tests hold no value-level vectors for these paths, hence
fixtures produced by running the reference itself.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def video():
    cases = []
    small = [(16, 16, 8, 8), (64, 48, 32, 24), (32, 24, 64, 48), (33, 17, 20, 11), (20, 11, 33, 17),
             (64, 48, 64, 48), (1, 1, 1, 1), (2, 2, 1, 1), (1, 1, 2, 2), (90, 40, 31, 40), (40, 90, 40, 31)]
    arrays = {}
    for (iw, ih, ow, oh) in small:
        for m in range(10):
            r = ob.RefVcs(iw, ih, ow, oh, m)
            frame = ob.nv12_random_frame(iw, ih, seed=iw * 131 + ih * 7 + m)
            out = r.convert(frame)
            key = f"v_{iw}x{ih}_{ow}x{oh}_m{m}"
            arrays[key] = out
            cases.append({"key": key, "in": [iw, ih], "out": [ow, oh], "method": m, "seed": iw * 131 + ih * 7 + m})
            r.close()
    big = []
    for (iw, ih, ow, oh, m) in [(1920, 1080, 1280, 720, 1), (3840, 2160, 1920, 1080, 3), (640, 480, 320, 240, 0),
                                (641, 481, 111, 30, 1), (320, 240, 640, 480, 3)]:
        r = ob.RefVcs(iw, ih, ow, oh, m)
        frame = ob.nv12_random_frame(iw, ih, seed=99)
        out = r.convert(frame)
        big.append({"in": [iw, ih], "out": [ow, oh], "method": m, "seed": 99,
                    "sha256": hashlib.sha256(out.tobytes()).hexdigest(), "first16": out[:16].tolist()})
        r.close()
    np.savez_compressed(os.path.join(HERE, "video_small.npz"), **arrays)
    json.dump({"small": cases, "big": big}, open(os.path.join(HERE, "video_cases.json"), "w"), indent=1)


def compositor():
    import ctypes as C
    rng = np.random.default_rng(2024)
    arrays, cases = {}, []
    for t in range(24):
        W, H = int(rng.integers(4, 80)), int(rng.integers(4, 60))
        fmt, bg = int(rng.choice([11, 12, 13, 14])), int(rng.integers(0, 4))
        n = int(rng.integers(1, 5))
        pads = (ob.OraclePad * n)()
        spec, keep = [], []
        for i in range(n):
            w, h = int(rng.integers(1, 50)), int(rng.integers(1, 40))
            seed = int(rng.integers(0, 1 << 30))
            a = np.random.default_rng(seed).integers(0, 256, (h, w, 4), dtype=np.uint8)
            keep.append(a)
            x, y = int(rng.integers(-20, W)), int(rng.integers(-20, H))
            al, op = float(rng.choice([0.25, 0.5, 1.0, 0.9])), int(rng.integers(0, 3))
            pads[i].data, pads[i].width, pads[i].height, pads[i].stride = a.ctypes.data, w, h, w * 4
            pads[i].xpos, pads[i].ypos, pads[i].alpha, pads[i].op = x, y, al, op
            spec.append([w, h, x, y, al, op, seed])
        dst = np.zeros((H, W, 4), dtype=np.uint8)
        ob.ref().ref_compositor(fmt, dst.ctypes.data, W, H, W * 4, bg, pads, n)
        arrays[f"c_{t}"] = dst
        cases.append({"key": f"c_{t}", "W": W, "H": H, "fmt": fmt, "bg": bg, "pads": spec})
    np.savez_compressed(os.path.join(HERE, "comp.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "comp_cases.json"), "w"), indent=1)


def audio():
    r = ob.ref()
    arrays, cases = {}, []
    for t, (a, b, ch, q, bufs) in enumerate([(48000, 44100, 2, 4, [480, 480, 480]), (44100, 48000, 1, 4, [441, 441]),
                                             (8000, 16000, 1, 4, [160, 160, 160, 160]), (48000, 44100, 4, 10, [1024, 512]),
                                             (12345, 54321, 2, 4, [600]), (96000, 8000, 1, 2, [4000, 4000])]):
        h = r.ref_ars_new(a, b, ch, q)
        rng = np.random.default_rng(500 + t)
        outs, counts = [], []
        for n in bufs:
            x = (rng.standard_normal((n, ch)) * 0.5).astype(np.float32)
            cap = int(n * b / a) + 64
            o = np.zeros((cap, ch), dtype=np.float32)
            k = r.ref_ars_process(h, x.ctypes.data, n, o.ctypes.data, cap)
            outs.append(o[:k].copy())
            counts.append(int(k))
        r.ref_ars_free(h)
        arrays[f"a_{t}"] = np.concatenate(outs) if outs else np.zeros((0, ch), np.float32)
        cases.append({"key": f"a_{t}", "in_rate": a, "out_rate": b, "ch": ch, "quality": q, "bufs": bufs,
                      "counts": counts, "seed": 500 + t})
    np.savez_compressed(os.path.join(HERE, "audio.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "audio_cases.json"), "w"), indent=1)


def video_planar():
    """I420 / YV12 inputs (unpack_I420 chain and the same-size convert_I420_BGRA-family fast path)"""
    arrays, cases = {}, []
    for fmt in ("I420", "YV12"):
        for (iw, ih, ow, oh) in [(16, 16, 8, 8), (64, 48, 32, 24), (32, 24, 64, 48), (33, 17, 19, 11), (64, 48, 64, 48),
                                 (65, 49, 65, 49), (1, 1, 1, 1), (2, 2, 2, 2), (20, 11, 33, 19)]:
            for m, ofmt in [(0, "BGRA"), (1, "RGBA"), (3, "ARGB"), (9, "xBGR")]:
                seed = iw * 17 + oh + m
                r = ob.RefVcs(iw, ih, ow, oh, m, in_fmt=ob.FMT[fmt], out_fmt=ob.FMT[ofmt])
                out = r.convert(ob.i420_random_frame(iw, ih, seed))
                r.close()
                key = f"p_{fmt}_{iw}x{ih}_{ow}x{oh}_m{m}_{ofmt}"
                arrays[key] = out
                cases.append({"key": key, "in_fmt": fmt, "out_fmt": ofmt, "in": [iw, ih], "out": [ow, oh], "method": m, "seed": seed})
    np.savez_compressed(os.path.join(HERE, "video_planar.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "video_planar_cases.json"), "w"), indent=1)


def audio_interpolated():
    """rate pairs whose phase table exceeds 1 MiB: interpolated filter mode"""
    r = ob.ref()
    arrays, cases = {}, []
    for t, (a, b, ch, q, bufs) in enumerate([(44100, 48001, 2, 4, [441, 441, 100]), (48000, 44101, 1, 6, [480, 480]),
                                             (96000, 8001, 2, 3, [4000, 1000]), (7999, 48000, 3, 10, [300, 300])]):
        h = r.ref_ars_new(a, b, ch, q)
        rng = np.random.default_rng(900 + t)
        outs, counts = [], []
        for n in bufs:
            x = (rng.standard_normal((n, ch)) * 0.5).astype(np.float32)
            cap = int(n * b / a) + 64
            o = np.zeros((cap, ch), dtype=np.float32)
            k = r.ref_ars_process(h, x.ctypes.data, n, o.ctypes.data, cap)
            outs.append(o[:k].copy())
            counts.append(int(k))
        r.ref_ars_free(h)
        arrays[f"ai_{t}"] = np.concatenate(outs)
        cases.append({"key": f"ai_{t}", "in_rate": a, "out_rate": b, "ch": ch, "quality": q, "bufs": bufs,
                      "counts": counts, "seed": 900 + t})
    np.savez_compressed(os.path.join(HERE, "audio_interp.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "audio_interp_cases.json"), "w"), indent=1)


def audio_formats():
    """S16 / S32 / F64 sample formats, FULL and interpolated filter modes"""
    r = ob.ref()
    arrays, cases = {}, []
    t = 0
    for fmt in ("S16", "S32", "F64"):
        _, gfmt, dt, _ = ob.AUDIO_FORMATS[fmt]
        for (a, b, ch, q, bufs) in [(48000, 44100, 2, 4, [480, 480]), (8000, 16000, 1, 4, [160, 160, 160]),
                                    (44100, 48001, 2, 4, [441, 100]), (96000, 8000, 1, 7, [4000])]:
            h = r.ref_ars_new_fmt(a, b, ch, q, gfmt)
            rng = np.random.default_rng(1200 + t)
            outs, counts = [], []
            for n in bufs:
                x = ob.audio_test_signal(rng, n, ch, fmt)
                cap = int(n * b / a) + 64
                o = np.zeros((cap, ch), dtype=dt)
                k = r.ref_ars_process(h, x.ctypes.data, n, o.ctypes.data, cap)
                outs.append(o[:k].copy())
                counts.append(int(k))
            r.ref_ars_free(h)
            arrays[f"af_{t}"] = np.concatenate(outs)
            cases.append({"key": f"af_{t}", "fmt": fmt, "in_rate": a, "out_rate": b, "ch": ch, "quality": q, "bufs": bufs,
                          "counts": counts, "seed": 1200 + t})
            t += 1
    np.savez_compressed(os.path.join(HERE, "audio_formats.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "audio_formats_cases.json"), "w"), indent=1)


def video_yuv():
    """YUV -> same YUV family plane scaling"""
    arrays, cases = {}, []
    for fi, fo in [("NV12", "NV12"), ("I420", "I420"), ("I420", "YV12"), ("NV21", "NV21")]:
        for (iw, ih, ow, oh) in [(64, 48, 32, 24), (64, 48, 100, 70), (65, 49, 33, 25), (64, 48, 64, 24), (32, 24, 64, 48),
                                 (33, 17, 20, 9)]:
            for m in (0, 1, 3, 9):
                frame = ob.i420_random_frame(iw, ih, m + iw) if fi in ("I420", "YV12") else ob.nv12_random_frame(iw, ih, m + iw)
                d = ob.vcs_desc(iw, ih, ow, oh, m, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo])
                size = ob.vcs_sizes(d)[1]
                r = ob.RefVcs(iw, ih, ow, oh, m, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo])
                out = r.convert(frame, np.zeros(size, dtype=np.uint8))
                r.close()
                key = f"y_{fi}_{fo}_{iw}x{ih}_{ow}x{oh}_m{m}"
                arrays[key] = out
                cases.append({"key": key, "in_fmt": fi, "out_fmt": fo, "in": [iw, ih], "out": [ow, oh], "method": m, "seed": m + iw})
    np.savez_compressed(os.path.join(HERE, "video_yuv.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "video_yuv_cases.json"), "w"), indent=1)


def video_cross():
    """4:2:0 -> the other 4:2:0 family: generic chain with chroma down-sampling (horizontal-first geometries, where the
    reference's temp-line ring does not alias)"""
    arrays, cases = {}, []
    for fi, fo in [("NV12", "I420"), ("I420", "NV12"), ("NV12", "NV21"), ("YV12", "NV21")]:
        for (iw, ih, ow, oh) in [(64, 48, 32, 24), (64, 48, 96, 72), (65, 49, 33, 25), (33, 17, 20, 31), (50, 21, 50, 21),
                                 (57, 35, 29, 35)]:
            for m, site, osite in ((1, 2, 2), (3, 1, 1), (9, 2, 1), (0, 1, 2)):
                if m == 0 and oh > ih:
                    continue    # nearest vertical repeats + in-place down-sampling: reference defect class
                frame = ob.i420_random_frame(iw, ih, m + iw) if fi in ("I420", "YV12") else ob.nv12_random_frame(iw, ih, m + iw)
                d = ob.vcs_desc(iw, ih, ow, oh, m, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo], site=site)
                size = ob.vcs_sizes(d)[1]
                r = ob.RefVcs(iw, ih, ow, oh, m, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo], site=site, matrix=d.in_matrix,
                              out_matrix=d.in_matrix, out_site=osite)
                out = r.convert(frame, np.zeros(size, dtype=np.uint8))
                r.close()
                key = f"x_{fi}_{fo}_{iw}x{ih}_{ow}x{oh}_m{m}_s{site}{osite}"
                arrays[key] = out
                cases.append({"key": key, "in_fmt": fi, "out_fmt": fo, "in": [iw, ih], "out": [ow, oh], "method": m,
                              "site": site, "out_site": osite, "seed": m + iw})
    np.savez_compressed(os.path.join(HERE, "video_cross.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "video_cross_cases.json"), "w"), indent=1)


def video_rgb_in():
    """packed RGB -> 4:2:0: unpack, scalers, RGB -> YUV table matrix, chroma down-sampling, pack"""
    arrays, cases = {}, []
    for fi, fo in [("BGRA", "NV12"), ("RGBA", "I420"), ("ARGB", "NV21"), ("xBGR", "YV12")]:
        for (iw, ih, ow, oh) in [(64, 48, 32, 24), (64, 48, 96, 72), (65, 49, 33, 25), (33, 17, 20, 31), (50, 21, 50, 21),
                                 (40, 90, 40, 31)]:
            for m, omat, orng, osite in ((1, 0, 0, 0), (3, 3, 2, 2), (9, 4, 1, 1), (0, 6, 2, 6)):
                if m == 0 and oh > ih:
                    continue    # nearest vertical repeats + in-place stages: reference defect class
                frame = np.random.default_rng(m + iw).integers(0, 256, iw * ih * 4, dtype=np.uint8)
                d = ob.vcs_desc(iw, ih, ow, oh, m, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo])
                size = ob.vcs_sizes(d)[1]
                r = ob.RefVcs(iw, ih, ow, oh, m, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo], out_matrix=omat or -1,
                              out_rng=orng or -1, out_site=osite or -1)
                out = r.convert(frame, np.zeros(size, dtype=np.uint8))
                r.close()
                key = f"r_{fi}_{fo}_{iw}x{ih}_{ow}x{oh}_m{m}"
                arrays[key] = out
                cases.append({"key": key, "in_fmt": fi, "out_fmt": fo, "in": [iw, ih], "out": [ow, oh], "method": m,
                              "out_matrix": omat, "out_range": orng, "out_site": osite, "seed": m + iw})
    np.savez_compressed(os.path.join(HERE, "video_rgb_in.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "video_rgb_in_cases.json"), "w"), indent=1)


def video_rgb_rgb():
    """packed RGB -> packed RGB: same format (one-plane scaling) and another byte order (chain without matrix)"""
    arrays, cases = {}, []
    for fi, fo in [("BGRA", "BGRA"), ("RGBA", "RGBA"), ("xRGB", "xRGB"), ("BGRA", "RGBA"), ("ARGB", "BGRx"), ("RGBx", "ABGR")]:
        for (iw, ih, ow, oh) in [(64, 48, 32, 24), (40, 30, 64, 48), (65, 49, 33, 25), (33, 17, 20, 31), (40, 90, 40, 31)]:
            for m in (0, 1, 3, 9):
                frame = np.random.default_rng(m + iw).integers(0, 256, iw * ih * 4, dtype=np.uint8)
                r = ob.RefVcs(iw, ih, ow, oh, m, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo])
                out = r.convert(frame, np.zeros(ow * oh * 4, dtype=np.uint8))
                r.close()
                key = f"q_{fi}_{fo}_{iw}x{ih}_{ow}x{oh}_m{m}"
                arrays[key] = out
                cases.append({"key": key, "in_fmt": fi, "out_fmt": fo, "in": [iw, ih], "out": [ow, oh], "method": m, "seed": m + iw})
    np.savez_compressed(os.path.join(HERE, "video_rgb_rgb.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "video_rgb_rgb_cases.json"), "w"), indent=1)


def video_422():
    """packed / planar 4:2:2 and 4:4:4 inputs -> packed RGB"""
    arrays, cases = {}, []
    for fi, fo in [("YUY2", "BGRA"), ("UYVY", "RGBA"), ("YVYU", "xRGB"), ("Y42B", "ARGB"), ("Y444", "BGRx")]:
        for (iw, ih, ow, oh) in [(64, 48, 32, 24), (64, 48, 96, 72), (65, 49, 33, 25), (33, 17, 20, 31), (50, 21, 50, 21), (40, 90, 40, 31)]:
            for m, site in ((1, 2), (3, 1), (9, 2), (0, 1)):
                if m == 0 and oh > ih and ow * oh <= iw * ih:
                    continue
                d = ob.vcs_desc(iw, ih, ow, oh, m, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo], site=site)
                frame = np.random.default_rng(m + iw).integers(0, 256, ob.vcs_sizes(d)[0], dtype=np.uint8)
                r = ob.RefVcs(iw, ih, ow, oh, m, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo], site=site)
                out = r.convert(frame, np.zeros(ow * oh * 4, dtype=np.uint8))
                r.close()
                key = f"s_{fi}_{fo}_{iw}x{ih}_{ow}x{oh}_m{m}"
                arrays[key] = out
                cases.append({"key": key, "in_fmt": fi, "out_fmt": fo, "in": [iw, ih], "out": [ow, oh], "method": m, "site": site,
                              "seed": m + iw})
    np.savez_compressed(os.path.join(HERE, "video_422.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "video_422_cases.json"), "w"), indent=1)


def video_422_420():
    """packed 4:2:2 -> 4:2:0 (capture -> encoder): the YUY2 / UYVY -> I420 / YV12 table rows at an unchanged size, the chain otherwise;
    output chroma-site = the default of the output size (the element's fixation across a sub-sampling change)"""
    arrays, cases = {}, []
    for fi, fo in [("YUY2", "I420"), ("UYVY", "YV12"), ("YVYU", "I420"), ("YUY2", "NV12"), ("UYVY", "NV21"),
                   ("Y42B", "I420"), ("Y444", "YV12"),          # planar -> planar: plane-scaling rows
                   ("Y42B", "NV12"), ("Y444", "NV21")]:         # planar -> semi-planar: the chain
        for (iw, ih, ow, oh) in [(64, 48, 64, 48), (33, 17, 33, 17), (50, 21, 50, 21), (64, 48, 32, 24), (64, 48, 96, 72), (33, 17, 20, 31)]:
            for m, site in ((1, 2), (3, 1), (9, 2)):
                d = ob.vcs_desc(iw, ih, ow, oh, m, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo], site=site)
                frame = np.random.default_rng(m + iw).integers(0, 256, ob.vcs_sizes(d)[0], dtype=np.uint8)
                out_site = 2 if oh > 576 else 1
                r = ob.RefVcs(iw, ih, ow, oh, m, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo], site=site, matrix=d.in_matrix,
                              out_matrix=d.in_matrix, out_site=out_site)
                out = r.convert(frame, np.zeros(ob.vcs_sizes(d)[1], dtype=np.uint8))
                r.close()
                key = f"t_{fi}_{fo}_{iw}x{ih}_{ow}x{oh}_m{m}"
                arrays[key] = out
                cases.append({"key": key, "in_fmt": fi, "out_fmt": fo, "in": [iw, ih], "out": [ow, oh], "method": m, "site": site,
                              "seed": m + iw})
    np.savez_compressed(os.path.join(HERE, "video_422_420.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "video_422_420_cases.json"), "w"), indent=1)


def compositor_420():
    """I420 / YV12 / NV12 / NV21 output"""
    o, r = ob.oracle(), ob.ref()
    rng = np.random.default_rng(4242)
    arrays, cases = {}, []
    for t in range(24):
        fmt = [2, 3, 23, 24][t % 4]
        W, H, bg, rg = int(rng.integers(2, 80)), int(rng.integers(2, 60)), int(rng.integers(0, 4)), int(rng.integers(0, 2))
        n = int(rng.integers(1, 5))
        pads = (ob.OraclePad * n)()
        spec, keep = [], []
        for i in range(n):
            w, h = int(rng.integers(1, 50)), int(rng.integers(1, 40))
            seed = int(rng.integers(0, 1 << 30))
            a = np.random.default_rng(seed).integers(0, 256, o.oracle_compositor_yuv_size(fmt, w, h), dtype=np.uint8)
            keep.append(a)
            x, y = int(rng.integers(-20, W)), int(rng.integers(-20, H))
            al, op = float(rng.choice([0.25, 0.5, 1.0, 0.9])), int(rng.integers(0, 3))
            pads[i].data, pads[i].width, pads[i].height, pads[i].stride = a.ctypes.data, w, h, 0
            pads[i].xpos, pads[i].ypos, pads[i].alpha, pads[i].op = x, y, al, op
            spec.append([w, h, x, y, al, op, seed])
        dst = np.zeros(o.oracle_compositor_yuv_size(fmt, W, H), dtype=np.uint8)
        r.ref_compositor_yuv(fmt, dst.ctypes.data, W, H, bg, rg, pads, n)
        arrays[f"cy_{t}"] = dst
        cases.append({"key": f"cy_{t}", "W": W, "H": H, "fmt": fmt, "bg": bg, "range": rg, "pads": spec})
    np.savez_compressed(os.path.join(HERE, "comp_420.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "comp_420_cases.json"), "w"), indent=1)


if __name__ == "__main__":
    assert ob.have_ref(), "needs oracle/_ref/libgstref.so (make -C oracle ref)"
    only = set(sys.argv[1:])
    for name, fn in [("video", video), ("compositor", compositor), ("audio", audio), ("video_planar", video_planar),
                     ("audio_interpolated", audio_interpolated), ("audio_formats", audio_formats), ("compositor_420", compositor_420), ("video_yuv", video_yuv),
                     ("video_cross", video_cross), ("video_rgb_in", video_rgb_in), ("video_rgb_rgb", video_rgb_rgb), ("video_422", video_422), ("video_422_420", video_422_420)]:
        if not only or name in only:
            fn()
    print("golden fixtures written to", HERE)
