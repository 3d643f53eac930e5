"""Destination rectangle + border fill (the element's add-borders; GST_VIDEO_CONVERTER_OPT_DEST_*): the input is scaled
into a rectangle of the output frame by the ordinary kernels (the plan describes the rectangle: same strides, shifted
plane origins) and vcs_border_kernel fills the rest with the border colour (setup_borderline / convert_fill_border,
video-converter.c:2189-2258, :7190-7300)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import bindings as ob

pytestmark = [pytest.mark.gpu]

PAIRS = [("NV12", "BGRA"), ("I420", "RGBA"), ("NV12", "NV12"), ("I420", "YV12"), ("NV12", "I420"), ("YV12", "NV21")]


def run(pair, size, frame_size, dest, method, border=0xff000000, seed=1):
    import torch
    import gstreamer_b200 as g
    from gstreamer_b200 import _lib
    fi, fo = pair
    iw, ih = size
    W, H = frame_size
    frame = ob.i420_random_frame(iw, ih, seed) if fi in ("I420", "YV12") else ob.nv12_random_frame(iw, ih, seed)
    d = ob.vcs_desc(iw, ih, W, H, method, in_fmt=ob.FMT[fi], out_fmt=ob.FMT[fo])
    want = ob.oracle_vcs_convert_dest(d, frame, dest, border, fill=0x5A)
    ii, oi = g.VideoInfo(ob.FMT[fi], iw, ih), g.VideoInfo(ob.FMT[fo], W, H)
    g.transfer_colorimetry_from_input(ii, oi)
    cfg = _lib.VcsConfigC()
    g.lib.b200_vcs_config_init(C.byref(cfg))
    cfg.method = method
    cfg.dest_x, cfg.dest_y, cfg.dest_width, cfg.dest_height = dest
    cfg.border_argb = border
    h = C.c_void_p()
    g._lib.check(g.lib.b200_vcs_create(C.byref(ii.c), C.byref(oi.c), C.byref(cfg), 0, C.byref(h)), "b200_vcs_create")
    try:
        src = torch.from_numpy(frame).cuda()
        dst = torch.full((oi.size,), 0x5A, dtype=torch.uint8, device="cuda")
        g._lib.check(g.lib.b200_vcs_convert(h, src.data_ptr(), dst.data_ptr(), None), "b200_vcs_convert")
        torch.cuda.synchronize()
        got = dst.cpu().numpy()
    finally:
        g.lib.b200_vcs_destroy(h)
    return got, want


@pytest.mark.parametrize("pair", PAIRS, ids=lambda p: "%s-%s" % p)
@pytest.mark.parametrize("method", [0, 1, 3, 9], ids=["nearest", "bilinear", "lanczos", "mitchell"])
def test_add_borders_rectangles_match_oracle(cuda_device, pair, method):
    for size, frame_size in [((64, 48), (80, 80)), ((64, 48), (40, 90)), ((100, 100), (150, 50)), ((33, 17), (70, 21)),
                             ((1920, 1080), (1280, 1024)), ((640, 480), (1920, 1080))]:
        if size[0] * size[1] > 500_000 and (method not in (1, 3) or pair[0] != "NV12"):
            continue
        dest = ob.vcs_borders(*size, *frame_size)
        got, want = run(pair, size, frame_size, dest, method)
        bad = np.argwhere(got != want)
        assert bad.size == 0, f"{size}->{frame_size} dest {dest}: {len(bad)} bytes differ, first at {bad[:4].ravel().tolist()}"


@pytest.mark.parametrize("pair", PAIRS, ids=lambda p: "%s-%s" % p)
def test_arbitrary_rectangles_and_colours(cuda_device, pair):
    rng = np.random.default_rng(5)
    for t in range(12):
        iw, ih, W, H = (int(v) for v in rng.integers(2, 90, 4))
        dw, dh = int(rng.integers(1, W + 1)), int(rng.integers(1, H + 1))
        dest = (int(rng.integers(0, W - dw + 1)), int(rng.integers(0, H - dh + 1)), dw, dh)
        border = [0xff000000, 0x80ff4020, 0xff10c0f0][t % 3]
        got, want = run(pair, (iw, ih), (W, H), dest, [1, 3, 0, 9][t % 4], border, seed=t)
        bad = np.argwhere(got != want)
        assert bad.size == 0, f"{(iw, ih)}->{(W, H)} dest {dest}: {len(bad)} bytes differ, first at {bad[:4].ravel().tolist()}"
