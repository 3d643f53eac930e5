"""The n-tap kernel (vcs_ntap.cuh: dp4a FIRs at any ratio) against the oracle and the generic kernel."""
import numpy as np
import pytest

from oracle import bindings as ob

pytestmark = pytest.mark.gpu

SIZES = [
    (1920, 1080, 1280, 720), (1280, 720, 1920, 1080), (3840, 2160, 1280, 720), (720, 576, 360, 288),
    (720, 480, 1280, 720), (642, 362, 320, 180), (322, 242, 1000, 730), (131, 77, 130, 76),
    (64, 64, 640, 640), (2048, 64, 256, 8), (9, 7, 8, 6), (4, 4, 8, 8), (8, 8, 4, 4), (254, 100, 127, 50),
    (1000, 600, 100, 60), (640, 480, 640, 360), (640, 360, 320, 360), (1920, 1080, 480, 270),
    (1920, 1080, 854, 480), (1280, 720, 854, 480), (100, 300, 250, 310), (3840, 2160, 3840, 1000),   # vertical pass first
]
METHODS = [2, 3, 4, 5, 6, 7, 8, 9]      # every n-tap method of GstVideoScaleMethod


def _convert(iw, ih, ow, oh, method, frame, variant=None, in_fmt=23, out_fmt=12, site=None, matrix=None, rng=None):
    import torch
    import gstreamer_b200 as g
    el = g.CudaVideoConvertScale(add_borders=False, method=method, cuda_device_id=0)
    ii = g.VideoInfo(in_fmt, iw, ih)
    ii.set_colorimetry(matrix=matrix, range=rng, chroma_site=site)
    oi = g.VideoInfo(out_fmt, ow, oh)
    el.set_info(ii, oi)
    if variant is not None:
        el.set_kernel_variant(variant)
    src = torch.from_numpy(frame).cuda()
    dst = torch.full((oi.size,), 0x5A, dtype=torch.uint8, device="cuda")
    el.transform_frame(src, dst)
    torch.cuda.synchronize()
    return dst.cpu().numpy(), int(el.plan_info().kernel_variant)


@pytest.mark.parametrize("method", [3, 2, 6], ids=["lanczos", "cubic", "sinc"])
@pytest.mark.parametrize("size", SIZES, ids=lambda s: "%dx%d-%dx%d" % s)
def test_ntap_matches_oracle(cuda_device, size, method):
    iw, ih, ow, oh = size
    frame = ob.nv12_random_frame(iw, ih, seed=iw + 3 * oh + method)
    want = ob.oracle_vcs_convert(ob.vcs_desc(iw, ih, ow, oh, method), frame)
    got, variant = _convert(iw, ih, ow, oh, method, frame)
    assert variant in (1, 3)                     # 1: the exact-2:1 kernel takes (254,100)->(127,50)-like shapes
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, f"variant {variant}: {bad.size} bytes differ, first at {bad[:8]}: got {got[bad[:8]]} want {want[bad[:8]]}"


@pytest.mark.parametrize("method", METHODS)
def test_ntap_is_selected_and_equals_generic(cuda_device, method):
    iw, ih, ow, oh = 1280, 720, 852, 480
    frame = ob.nv12_smpte_like_frame(iw, ih, method)
    got, variant = _convert(iw, ih, ow, oh, method, frame)
    assert variant == 3
    generic, v0 = _convert(iw, ih, ow, oh, method, frame, variant=0)
    assert v0 == 0 and np.array_equal(got, generic)
    assert np.array_equal(got, ob.oracle_vcs_convert(ob.vcs_desc(iw, ih, ow, oh, method), frame))


@pytest.mark.parametrize("site", [1, 2, 4, 6])
@pytest.mark.parametrize("in_fmt", ["NV12", "NV21"])
@pytest.mark.parametrize("out_fmt", ["BGRA", "RGBx", "ARGB", "xBGR"])
@pytest.mark.parametrize("size", [(322, 182, 200, 120), (200, 120, 322, 194), (320, 180, 320, 100), (320, 180, 200, 180)],
                         ids=["down", "up", "v-only", "h-only"])
def test_ntap_formats_and_siting(cuda_device, site, in_fmt, out_fmt, size):
    iw, ih, ow, oh = size
    frame = ob.nv12_random_frame(iw, ih, seed=site)
    d = ob.vcs_desc(iw, ih, ow, oh, 3, in_fmt=ob.FMT[in_fmt], out_fmt=ob.FMT[out_fmt], site=site)
    want = ob.oracle_vcs_convert(d, frame)
    got, variant = _convert(iw, ih, ow, oh, 3, frame, in_fmt=ob.FMT[in_fmt], out_fmt=ob.FMT[out_fmt], site=site)
    assert variant == 3                          # either pass order
    assert np.array_equal(got, want)


@pytest.mark.parametrize("matrix,rng", [(3, 2), (4, 2), (4, 1), (6, 2), (2, 1), (5, 2)])
def test_ntap_colorimetry(cuda_device, matrix, rng):
    iw, ih, ow, oh = 330, 200, 220, 134
    frame = ob.nv12_random_frame(iw, ih, seed=matrix)
    want = ob.oracle_vcs_convert(ob.vcs_desc(iw, ih, ow, oh, 3, matrix=matrix, rng=rng), frame)
    got, variant = _convert(iw, ih, ow, oh, 3, frame, matrix=matrix, rng=rng)
    assert variant == 3 and np.array_equal(got, want)


def test_ntap_extreme_content(cuda_device):
    """saturating content (0/255 checker): overshoot clamps in both passes"""
    iw, ih, ow, oh = 640, 360, 426, 240
    st = (iw + 3) & ~3
    frame = np.zeros(st * ih * 3 // 2, dtype=np.uint8)
    yy, xx = np.mgrid[0:ih, 0:st]
    frame[: st * ih] = (((xx // 3 + yy // 2) & 1) * 255).astype(np.uint8).ravel()
    frame[st * ih:] = np.tile(np.array([0, 255, 255, 0], dtype=np.uint8), st * ih // 8)
    for method in (3, 6, 9):
        want = ob.oracle_vcs_convert(ob.vcs_desc(iw, ih, ow, oh, method), frame)
        got, variant = _convert(iw, ih, ow, oh, method, frame)
        assert variant == 3 and np.array_equal(got, want)


def test_ntap_batch_full_size(cuda_device):
    import torch
    import gstreamer_b200 as g
    iw, ih, ow, oh = 3840, 2160, 1280, 720
    frames = [ob.nv12_random_frame(iw, ih, s) for s in range(2)]
    d = ob.vcs_desc(iw, ih, ow, oh, 3)
    el = g.CudaVideoConvertScale(add_borders=False, method=3)
    ii, oi = g.VideoInfo(23, iw, ih), g.VideoInfo(12, ow, oh)
    el.set_info(ii, oi)
    assert int(el.plan_info().kernel_variant) == 3
    src = [torch.from_numpy(f).cuda() for f in frames]
    dst = [torch.zeros(oi.size, dtype=torch.uint8, device="cuda") for _ in frames]
    el.transform_frames(src, dst)
    torch.cuda.synchronize()
    for f, o in zip(frames, dst):
        assert np.array_equal(o.cpu().numpy(), ob.oracle_vcs_convert(d, f))
