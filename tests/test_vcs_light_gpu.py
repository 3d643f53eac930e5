"""The light kernel (copy / 2-tap axes, vcs_light.cuh) against the oracle and against the generic kernel.

C1 (1920x1080 NV12 -> 1280x720 BGRA, method=bilinear: the element default) is in this class."""
import numpy as np
import pytest

from oracle import bindings as ob

pytestmark = pytest.mark.gpu

SIZES = [
    (1920, 1080, 1280, 720), (1280, 720, 1920, 1080), (640, 480, 640, 480), (720, 576, 360, 288),
    (720, 480, 1280, 720), (642, 362, 320, 180), (322, 242, 1000, 700), (131, 77, 130, 76),
    (64, 64, 640, 640), (2048, 64, 256, 8), (9, 7, 8, 6), (4, 4, 8, 8), (8, 8, 4, 4), (254, 100, 127, 50),
    (1000, 600, 100, 60), (1920, 1080, 854, 480), (250, 140, 167, 93), (640, 480, 640, 200), (100, 300, 250, 310),
]


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


@pytest.mark.parametrize("method", [0, 1], ids=["nearest", "bilinear"])
@pytest.mark.parametrize("size", SIZES, ids=lambda s: "%dx%d-%dx%d" % s)
def test_light_matches_oracle_and_generic(cuda_device, size, method):
    iw, ih, ow, oh = size
    frame = ob.nv12_random_frame(iw, ih, seed=iw + 3 * oh + method)
    want = ob.oracle_vcs_convert(ob.vcs_desc(iw, ih, ow, oh, method), frame)
    got, variant = _convert(iw, ih, ow, oh, method, frame)
    assert variant == 2                          # either pass order
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, f"{bad.size} bytes differ, first at {bad[:8]}: got {got[bad[:8]]} want {want[bad[:8]]}"
    generic, v0 = _convert(iw, ih, ow, oh, method, frame, variant=0)
    assert v0 == 0 and np.array_equal(generic, want)


@pytest.mark.parametrize("site", [1, 2, 4, 6])
@pytest.mark.parametrize("in_fmt", ["NV12", "NV21"])
@pytest.mark.parametrize("out_fmt", ["BGRA", "RGBx", "ARGB", "xBGR"])
@pytest.mark.parametrize("size", [(322, 182, 200, 120), (200, 120, 322, 194)], ids=["down", "up"])
def test_light_formats_and_siting(cuda_device, site, in_fmt, out_fmt, size):
    iw, ih, ow, oh = size
    frame = ob.nv12_random_frame(iw, ih, seed=site)
    d = ob.vcs_desc(iw, ih, ow, oh, 1, in_fmt=ob.FMT[in_fmt], out_fmt=ob.FMT[out_fmt], site=site)
    want = ob.oracle_vcs_convert(d, frame)
    got, variant = _convert(iw, ih, ow, oh, 1, frame, in_fmt=ob.FMT[in_fmt], out_fmt=ob.FMT[out_fmt], site=site)
    assert variant == 2 and np.array_equal(got, want)


@pytest.mark.parametrize("matrix,rng", [(3, 2), (4, 2), (4, 1), (6, 2), (2, 1), (5, 2)])
def test_light_colorimetry(cuda_device, matrix, rng):
    iw, ih, ow, oh = 330, 200, 220, 134
    frame = ob.nv12_random_frame(iw, ih, seed=matrix)
    want = ob.oracle_vcs_convert(ob.vcs_desc(iw, ih, ow, oh, 1, matrix=matrix, rng=rng), frame)
    got, variant = _convert(iw, ih, ow, oh, 1, frame, matrix=matrix, rng=rng)
    assert variant == 2 and np.array_equal(got, want)


def test_light_pitched_layout(cuda_device):
    """common-pitch GstCudaMemory layout with plane offsets (gstcudamemory.cpp:194-345)"""
    import torch
    import gstreamer_b200 as g
    iw, ih, ow, oh = 250, 140, 166, 93
    frame = ob.nv12_random_frame(iw, ih, 2)
    want = ob.oracle_vcs_convert(ob.vcs_desc(iw, ih, ow, oh, 1), frame).reshape(oh, ow * 4)
    st = (iw + 3) & ~3
    pitch, opitch = 512, 1024
    padded = np.full(pitch * (ih + ih // 2) + 256, 0x33, dtype=np.uint8)
    padded[256: 256 + pitch * ih].reshape(ih, pitch)[:, :st] = frame[: st * ih].reshape(ih, st)
    padded[256 + pitch * ih:].reshape(ih // 2, pitch)[:, :st] = frame[st * ih:].reshape(ih // 2, st)
    ii = g.VideoInfo(23, iw, ih).set_layout([pitch, pitch], [256, 256 + pitch * ih])
    oi = g.VideoInfo(12, ow, oh).set_layout([opitch], [64])
    el = g.CudaVideoConvertScale(add_borders=False, method=1)
    el.set_info(ii, oi)
    assert int(el.plan_info().kernel_variant) == 2
    dst = torch.full((64 + opitch * oh,), 0x77, dtype=torch.uint8, device="cuda")
    el.transform_frame(torch.from_numpy(padded).cuda(), dst)
    torch.cuda.synchronize()
    got = dst.cpu().numpy()
    assert np.array_equal(got[64:].reshape(oh, opitch)[:, : ow * 4], want)
    assert np.all(got[:64] == 0x77) and np.all(got[64:].reshape(oh, opitch)[:, ow * 4:] == 0x77)


def test_light_unaligned_layout_falls_back(cuda_device):
    """a luma pitch that is not a multiple of 4 cannot use 32-bit row loads: generic kernel, same bytes"""
    import torch
    import gstreamer_b200 as g
    iw, ih, ow, oh = 122, 60, 80, 40
    frame = ob.nv12_random_frame(iw, ih, 4)
    want = ob.oracle_vcs_convert(ob.vcs_desc(iw, ih, ow, oh, 1), frame)
    st = (iw + 3) & ~3
    pitch = 126
    padded = np.zeros(pitch * (ih + ih // 2), dtype=np.uint8)
    padded[: pitch * ih].reshape(ih, pitch)[:, :iw] = frame[: st * ih].reshape(ih, st)[:, :iw]
    padded[pitch * ih:].reshape(ih // 2, pitch)[:, :iw] = frame[st * ih:].reshape(ih // 2, st)[:, :iw]
    ii = g.VideoInfo(23, iw, ih).set_layout([pitch, pitch], [0, pitch * ih])
    oi = g.VideoInfo(12, ow, oh)
    el = g.CudaVideoConvertScale(add_borders=False, method=1)
    el.set_info(ii, oi)
    assert int(el.plan_info().kernel_variant) == 0
    dst = torch.zeros(oi.size, dtype=torch.uint8, device="cuda")
    el.transform_frame(torch.from_numpy(padded).cuda(), dst)
    torch.cuda.synchronize()
    assert np.array_equal(dst.cpu().numpy(), want)


def test_light_batch(cuda_device):
    import torch
    import gstreamer_b200 as g
    iw, ih, ow, oh = 1920, 1080, 1280, 720
    frames = [ob.nv12_smpte_like_frame(iw, ih, s) for s in range(3)]
    d = ob.vcs_desc(iw, ih, ow, oh, 1)
    el = g.CudaVideoConvertScale(add_borders=False, method=1)
    ii, oi = g.VideoInfo(23, iw, ih), g.VideoInfo(12, ow, oh)
    el.set_info(ii, oi)
    src = [torch.from_numpy(f).cuda() for f in frames]
    dst = [torch.zeros(oi.size, dtype=torch.uint8, device="cuda") for _ in frames]
    el.transform_frames(src, dst)
    torch.cuda.synchronize()
    for f, o in zip(frames, dst):
        assert np.array_equal(o.cpu().numpy(), ob.oracle_vcs_convert(d, f))
