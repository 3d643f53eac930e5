#!/usr/bin/env python
"""tools/planar_in_bench.py — planar (I420) against semi-planar (NV12) input on the two RGB bench shapes: which kernel runs, us/frame."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import gstreamer_b200 as g  # noqa: E402
from gstreamer_b200 import _lib  # noqa: E402
from oracle import bindings as ob  # noqa: E402

for (fmt, IW, IH, OW, OH, m) in [(23, 3840, 2160, 1920, 1080, 3), (2, 3840, 2160, 1920, 1080, 3), (23, 1920, 1080, 1280, 720, 1),
                                 (2, 1920, 1080, 1280, 720, 1), (2, 1920, 1080, 1280, 720, 3)]:
    el = g.CudaVideoConvertScale(method=m)
    ii, oi = g.VideoInfo(fmt, IW, IH), g.VideoInfo(12, OW, OH)
    el.set_info(ii, oi)
    per = 32
    gen = ob.i420_random_frame if fmt in (2, 3) else ob.nv12_random_frame
    base = [torch.from_numpy(gen(IW, IH, s)).cuda() for s in range(2)]
    rin = [base[k % 2].clone() for k in range(2 * per)]
    rout = [torch.empty(oi.size, dtype=torch.uint8, device="cuda") for _ in range(2 * per)]
    s = torch.cuda.Stream()
    step = lambda i: el.transform_frames(rin[(i % 2) * per:(i % 2 + 1) * per], rout[(i % 2) * per:(i % 2 + 1) * per], s)
    with torch.cuda.stream(s):
        for i in range(3):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for i in range(10):
            step(i)
        e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"config": f"{g.VideoFormat(fmt).name} {IW}x{IH} -> BGRA {OW}x{OH} method {m}",
                      "kernel": _lib.lib.b200_vcs_kernel_name(el._h).decode(), "us_per_frame": round(ms * 1e3 / per, 3)}), flush=True)
    del rin, rout, base
