"""I420 / YV12 input: unpack_I420 in front of the same chain, and the reference's
convert_I420_BGRA-family fast path (nearest chroma) when the size does not change."""
import numpy as np
import pytest

from oracle import bindings as ob

pytestmark = pytest.mark.gpu

SIZES = [(640, 480, 320, 240), (320, 240, 640, 480), (641, 481, 111, 30), (64, 48, 64, 48), (65, 49, 65, 49),
         (1, 1, 1, 1), (2, 2, 2, 2), (3, 5, 7, 2), (17, 33, 64, 7), (1920, 1080, 1280, 720), (30, 111, 641, 481),
         (1920, 1080, 1920, 1080), (1920, 1080, 854, 480), (3840, 2160, 1920, 1080), (250, 140, 167, 93)]


def _convert(iw, ih, ow, oh, method, frame, in_fmt, variant=None, out_fmt=12, site=None, matrix=None, rng=None,
             layout=None):
    import torch
    import gstreamer_b200 as g
    el = g.CudaVideoConvertScale(add_borders=False, method=method, cuda_device_id=0)
    ii = g.VideoInfo(in_fmt, iw, ih)
    ii.set_colorimetry(matrix=matrix, range=rng, chroma_site=site)
    if layout:
        ii.set_layout(*layout)
    oi = g.VideoInfo(out_fmt, ow, oh)
    el.set_info(ii, oi)
    if variant is not None:
        el.set_kernel_variant(variant)
    src = torch.from_numpy(frame).cuda()
    dst = torch.full((oi.size,), 0x5A, dtype=torch.uint8, device="cuda")
    el.transform_frame(src, dst)
    torch.cuda.synchronize()
    return dst.cpu().numpy(), int(el.plan_info().kernel_variant)


@pytest.mark.parametrize("method", [0, 1, 3, 9], ids=["nearest", "bilinear", "lanczos", "mitchell"])
@pytest.mark.parametrize("size", SIZES, ids=lambda s: "%dx%d-%dx%d" % s)
@pytest.mark.parametrize("in_fmt", ["I420", "YV12"])
def test_planar_matches_oracle(cuda_device, in_fmt, size, method):
    iw, ih, ow, oh = size
    fmt = ob.FMT[in_fmt]
    frame = ob.i420_random_frame(iw, ih, seed=iw + oh + method)
    want = ob.oracle_vcs_convert(ob.vcs_desc(iw, ih, ow, oh, method, in_fmt=fmt), frame)
    got, variant = _convert(iw, ih, ow, oh, method, frame, fmt)
    assert variant in (1, 2, 3), "default-layout frames must take a fast kernel"     # 1: exact 2:1 with 8 taps (second form, planar)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, f"variant {variant}: {bad.size} bytes differ, first at {bad[:8]}: got {got[bad[:8]]} want {want[bad[:8]]}"
    if iw * ih <= 700 * 500:
        generic, v0 = _convert(iw, ih, ow, oh, method, frame, fmt, variant=0)
        assert v0 == 0 and np.array_equal(generic, want)


@pytest.mark.parametrize("in_fmt", ["I420", "YV12"])
@pytest.mark.parametrize("out_fmt", ["RGBx", "BGRx", "xRGB", "xBGR", "RGBA", "BGRA", "ARGB", "ABGR"])
@pytest.mark.parametrize("size", [(98, 66, 45, 37), (98, 66, 98, 66), (98, 66, 150, 101)], ids=["down", "same-size", "up"])
@pytest.mark.parametrize("matrix,rng,site", [(3, 2, 2), (4, 1, 1), (6, 2, 6), (2, 1, 4)])
def test_planar_formats_and_colorimetry(cuda_device, in_fmt, out_fmt, size, matrix, rng, site):
    iw, ih, ow, oh = size
    frame = ob.i420_random_frame(iw, ih, 5)
    kw = dict(in_fmt=ob.FMT[in_fmt], out_fmt=ob.FMT[out_fmt], site=site, matrix=matrix, rng=rng)
    want = ob.oracle_vcs_convert(ob.vcs_desc(iw, ih, ow, oh, 3, **kw), frame)
    got, variant = _convert(iw, ih, ow, oh, 3, frame, kw["in_fmt"], out_fmt=kw["out_fmt"], site=site, matrix=matrix, rng=rng)
    assert variant in (2, 3) and np.array_equal(got, want)


def test_planar_distinct_plane_pitches_fall_back(cuda_device):
    """U and V planes with different pitches (legal in a GstVideoMeta): generic kernel, same bytes"""
    iw, ih, ow, oh = 122, 60, 80, 40
    frame = ob.i420_random_frame(iw, ih, 4)
    want = ob.oracle_vcs_convert(ob.vcs_desc(iw, ih, ow, oh, 1, in_fmt=2), frame)
    sy, sc, hh = (iw + 3) & ~3, ((iw + 1) // 2 + 3) & ~3, (ih + 1) & ~1
    py, pu, pv = 128, 72, 96
    padded = np.zeros(py * hh + (pu + pv) * (hh // 2) + 64, dtype=np.uint8)
    ou, ov = py * hh + 16, py * hh + 16 + pu * (hh // 2) + 16
    padded[: py * hh].reshape(hh, py)[:, :sy] = frame[: sy * hh].reshape(hh, sy)
    padded[ou: ou + pu * (hh // 2)].reshape(hh // 2, pu)[:, :sc] = frame[sy * hh: sy * hh + sc * (hh // 2)].reshape(hh // 2, sc)
    padded[ov: ov + pv * (hh // 2)].reshape(hh // 2, pv)[:, :sc] = frame[sy * hh + sc * (hh // 2):].reshape(hh // 2, sc)
    got, variant = _convert(iw, ih, ow, oh, 1, padded, 2, layout=([py, pu, pv], [0, ou, ov]))
    assert variant == 0 and np.array_equal(got, want)
