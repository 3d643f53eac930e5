#!/usr/bin/env python
"""Device check of the chain + chroma down-sampling paths (kernel_variant 5) against the oracle: the 4:2:0 -> other-4:2:0
family pairs, and with `rgb` as second argument the (opt-in) packed RGB -> 4:2:0 direction.
Runs every case of tests/test_vcs_cross_gpu.py plus a random sweep, records EVERY outcome (it does not stop at the
first mismatch) in gpurun_out/cross_check.json, and times one 1080p -> 720p NV12 -> I420 conversion."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np   # noqa: E402


def rgb_mode(budget, t_start, out, flush):
    """random packed RGB -> 4:2:0 configurations"""
    import test_vcs_rgbin_gpu as R
    import test_vcs_cross_gpu as T
    rng = np.random.default_rng(2)
    while time.time() - t_start < budget and len(out["errors"]) <= 5:
        iw, ih, ow, oh = (int(v) for v in rng.integers(1, 200, 4))
        if rng.random() < 0.2:
            ow = iw
        if rng.random() < 0.2:
            oh = ih
        fi, fo = R.RGB_IN[int(rng.integers(0, 8))], R.YUV_OUT[int(rng.integers(0, 4))]
        method = int(rng.integers(0, 10))
        col = (int(rng.choice([2, 3, 4, 5, 6])), int(rng.choice([1, 2])), int(rng.choice([1, 2, 4, 6]))) if rng.random() < 0.5 else None
        tag = f"{fi}-{fo} {iw}x{ih}-{ow}x{oh} m{method} {col}"
        out["cases"] += 1
        try:
            frame = R.rgb_frame(iw, ih, int(rng.integers(0, 1000)))
            want = R.expected((iw, ih, ow, oh), method, frame, fi, fo, col)
            (got,), oi = R.convert((iw, ih, ow, oh), method, frame, fi, fo, col)
            bad = T.planes_equal(got, want, oi, ow, oh, fo in ("NV12", "NV21"))
            if bad:
                out["failures"].append(tag + ": " + "; ".join(bad))
            else:
                out["ok"] += 1
        except Exception as e:                                   # noqa: BLE001
            out["errors"].append(tag + ": " + repr(e)[:300])
    flush()
    print(json.dumps({k: (v if not isinstance(v, list) else v[:8]) for k, v in out.items()}, indent=1))
    return 0 if not out["failures"] and not out["errors"] else 1


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    t_start = time.time()
    import torch
    import gstreamer_b200 as g
    import test_vcs_cross_gpu as T
    from oracle import bindings as ob
    out = {"cases": 0, "ok": 0, "odd_height_cases": 0, "failures": [], "errors": []}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)

    def flush():
        out["seconds"] = round(time.time() - t_start, 1)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cross_check.json"), "w"), indent=1)

    def one(pair, size, method, site, out_site, seed=5):
        iw, ih, ow, oh = size
        tag = f"{pair[0]}-{pair[1]} {iw}x{ih}-{ow}x{oh} m{method} site{site}->{out_site}"
        out["cases"] += 1
        try:
            frame = T.random_frame(pair, iw, ih, seed)
            odd = (oh & 1) and ih == oh and ((iw, ih) != (ow, oh) or site != out_site) and not (site & 4) and not (out_site & 4)
            if odd and ih >= 3:
                out["odd_height_cases"] += 1
            want = T.expected(size, method, frame, pair, site, out_site)
            (got,), oi = T.convert(size, method, frame, pair, site, out_site)
            bad = T.planes_equal(got, want, oi, ow, oh, pair[1] in ("NV12", "NV21"))
            if bad:
                out["failures"].append(tag + ": " + "; ".join(bad))
            else:
                out["ok"] += 1
        except Exception as e:                                   # noqa: BLE001 - record and go on
            out["errors"].append(tag + ": " + repr(e)[:300])

    odd_only = len(sys.argv) > 2 and sys.argv[2] == "odd"
    if len(sys.argv) > 2 and sys.argv[2] == "rgb":
        return rgb_mode(budget, t_start, out, flush)
    # 1. the pytest matrix (small shapes first)
    for size in [] if odd_only else sorted(T.SIZES, key=lambda s: s[0] * s[1]):
        for pair in T.PAIRS:
            for method in (0, 1, 3, 4, 9):
                if size[0] * size[1] > 500_000 and (pair != ("NV12", "I420") or method not in (1, 3)):
                    continue
                for site, out_site in ((2, 2), (1, 1), (2, 1), (6, 4)):
                    one(pair, size, method, site, out_site)
        flush()
        if time.time() - t_start > budget or len(out["errors"]) > 5:
            break
    # 2. random sweep with whatever time is left
    rng = np.random.default_rng(1)
    while time.time() - t_start < budget and len(out["errors"]) <= 5:
        iw, ih, ow, oh = (int(v) for v in rng.integers(1, 200, 4))
        if rng.random() < 0.2:
            ow = iw
        if rng.random() < 0.2 or odd_only:
            oh = ih = ih | 1
        pair = T.PAIRS[int(rng.integers(0, len(T.PAIRS)))]
        one(pair, (iw, ih, ow, oh), int(rng.integers(0, 10)), int(rng.choice([1, 2, 4, 6])), int(rng.choice([1, 2, 4, 6])),
            seed=int(rng.integers(0, 1000)))
    flush()
    if odd_only:
        print(json.dumps({k: (v if not isinstance(v, list) else v[:8]) for k, v in out.items()}, indent=1))
        return 0 if not out["failures"] and not out["errors"] else 1
    # 3. timing: 1080p -> 720p NV12 -> I420 lanczos, 16 frames per launch pair
    try:
        size, pair = (1920, 1080, 1280, 720), ("NV12", "I420")
        el = g.CudaVideoConvertScale(add_borders=False, method=3, cuda_device_id=0)
        ii, oi = g.VideoInfo(23, 1920, 1080), g.VideoInfo(2, 1280, 720)
        oi.set_colorimetry(matrix=ii.c.color_matrix, chroma_site=ii.c.chroma_site)
        el.set_info(ii, oi)
        n = 16
        src = [torch.randint(0, 256, (ii.size,), dtype=torch.uint8, device="cuda") for _ in range(n)]
        dst = [torch.empty(oi.size, dtype=torch.uint8, device="cuda") for _ in range(n)]
        for _ in range(3):
            el.transform_frames(src, dst)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            el.transform_frames(src, dst)
        e1.record()
        torch.cuda.synchronize()
        out["timing_1080p_720p_nv12_i420_lanczos_us_per_frame"] = round(e0.elapsed_time(e1) * 1000 / (10 * n), 2)
    except Exception as e:                                       # noqa: BLE001
        out["errors"].append("timing: " + repr(e)[:300])
    flush()
    print(json.dumps({k: (v if not isinstance(v, list) else v[:8]) for k, v in out.items()}, indent=1))
    return 0 if not out["failures"] and not out["errors"] else 1


if __name__ == "__main__":
    sys.exit(main())
