#!/usr/bin/env python
"""tools/gpu_fuzz.py [seconds] — randomised differential run of the CUDA convert+scale path (through the C-ABI)
against the CPU oracle: random sizes, methods, input/output formats, colorimetry, chroma siting.  Development aid
(run on a GPU box); prints every mismatch and a summary, exits non-zero on any."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import gstreamer_b200 as g  # noqa: E402
from oracle import bindings as ob  # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    t0, n, bad, variants = time.time(), 0, 0, {}
    while time.time() - t0 < budget:
        iw, ih = int(rng.integers(1, 500)), int(rng.integers(1, 300))
        if rng.random() < 0.5:
            ow, oh = int(rng.integers(1, 500)), int(rng.integers(1, 300))
        else:
            f = rng.uniform(0.2, 3.0)
            ow, oh = max(1, int(iw * f)), max(1, int(ih * f))
        m = int(rng.integers(0, 10))
        fi = int(rng.choice([23, 24, 2, 3]))
        r = rng.random()
        if r < 0.2:
            fo = {23: 23, 24: 24, 2: int(rng.choice([2, 3])), 3: int(rng.choice([2, 3]))}[fi]
        elif r < 0.35:
            fo = int(rng.choice([23, 24, 2, 3]))            # includes the cross-family pairs (chain + chroma down-sampling)
        else:
            fo = int(rng.choice([7, 8, 9, 10, 11, 12, 13, 14]))
        site, mat, rg = int(rng.choice([1, 2, 4, 6])), int(rng.choice([3, 4, 6, 2, 5])), int(rng.choice([1, 2]))
        frame = ob.i420_random_frame(iw, ih, n) if fi in (2, 3) else ob.nv12_random_frame(iw, ih, n)
        out_site = int(rng.choice([1, 2, 4, 6]))
        d = ob.vcs_desc(iw, ih, ow, oh, m, in_fmt=fi, out_fmt=fo, site=site, matrix=mat, rng=rg)
        d.out_chroma_site = out_site
        want = ob.oracle_vcs_convert(d, frame)
        el = g.CudaVideoConvertScale(add_borders=False, method=m)
        ii = g.VideoInfo(fi, iw, ih)
        ii.set_colorimetry(matrix=mat, range=rg, chroma_site=site)
        oi = g.VideoInfo(fo, ow, oh)
        if fo in (23, 24, 2, 3):
            oi.set_colorimetry(matrix=mat, range=rg, chroma_site=out_site)   # caps fixation carries the input colorimetry over
        el.set_info(ii, oi)
        v = int(el.plan_info().kernel_variant)
        variants[v] = variants.get(v, 0) + 1
        dst = torch.zeros(oi.size, dtype=torch.uint8, device="cuda")
        el.transform_frame(torch.from_numpy(frame).cuda(), dst)
        torch.cuda.synchronize()
        got = dst.cpu().numpy()
        n += 1
        # YUV outputs: compare pixel bytes only (row padding is not written by either side, both start at zero)
        if not np.array_equal(got, want):
            bad += 1
            print("MISMATCH", (iw, ih, ow, oh), "method", m, "fmt", fi, "->", fo, "site", site, "matrix", mat, "range", rg,
                  "variant", v, int(np.count_nonzero(got != want)), "of", got.size, flush=True)
    print(f"gpu_fuzz: {n} cases, {bad} mismatches, kernel variants {variants}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
