"""The specialised 2:1 / 8-tap kernel must produce the same bytes as the oracle (and as the
generic kernel) on every shape it declares itself eligible for."""
import numpy as np
import pytest

from oracle import bindings as ob

pytestmark = pytest.mark.gpu

SIZES = [(16, 16), (64, 48), (256, 144), (480, 272), (488, 264), (1920, 1080), (1928, 1096), (3840, 2160)]
METHODS = [3, 5, 6, 7, 8, 9]          # every element method that yields 8 taps at 2:1


def _convert(iw, ih, method, frame, variant, in_fmt=23, out_fmt=12, matrix=None, rng=None):
    import torch
    import gstreamer_b200 as g
    el = g.CudaVideoConvertScale(add_borders=False, method=method)
    ii = g.VideoInfo(in_fmt, iw, ih).set_colorimetry(chroma_site=2, matrix=matrix, range=rng)
    oi = g.VideoInfo(out_fmt, iw // 2, ih // 2)
    el.set_info(ii, oi)
    assert el.plan_info().kernel_variant == 1, "shape should be eligible for the specialised kernel"
    el.set_kernel_variant(variant)
    src = torch.from_numpy(frame).cuda()
    dst = torch.full((oi.size,), 0x5A, dtype=torch.uint8, device="cuda")
    el.transform_frame(src, dst)
    torch.cuda.synchronize()
    return dst.cpu().numpy()


def _explain(got, want, ow):
    bad = np.nonzero(got != want)[0]
    if bad.size == 0:
        return ""
    px = bad // 4
    rows, cols = np.unique(px // ow), np.unique(px % ow)
    return (f"{bad.size} bytes differ; rows {rows[:12]}.. ({rows.size}), cols {cols[:12]}.. ({cols.size}), "
            f"channels {np.unique(bad % 4)}, first got {got[bad[:6]]} want {want[bad[:6]]}")


# the full-size shapes run for the headline method only (the CPU oracle needs seconds per frame there)
L2_CASES = [(s, m) for m in METHODS for s in SIZES if m == 3 or s[0] <= 2000]


@pytest.mark.parametrize("size,method", L2_CASES, ids=lambda v: "%dx%d" % v if isinstance(v, tuple) else str(v))
def test_specialised_matches_oracle(cuda_device, size, method):
    iw, ih = size
    d = ob.vcs_desc(iw, ih, iw // 2, ih // 2, method, site=2)
    frame = ob.nv12_random_frame(iw, ih, seed=iw + method)
    want = ob.oracle_vcs_convert(d, frame)
    got = _convert(iw, ih, method, frame, 1)
    assert np.array_equal(got, want), _explain(got, want, iw // 2)


@pytest.mark.parametrize("in_fmt", ["NV12", "NV21"])
@pytest.mark.parametrize("out_fmt", ["BGRA", "RGBA", "ARGB", "ABGR"])
def test_specialised_formats(cuda_device, in_fmt, out_fmt):
    iw, ih = 496, 280
    d = ob.vcs_desc(iw, ih, iw // 2, ih // 2, 3, in_fmt=ob.FMT[in_fmt], out_fmt=ob.FMT[out_fmt], site=2,
                    matrix=4, rng=1)
    frame = ob.nv12_random_frame(iw, ih, seed=3)
    want = ob.oracle_vcs_convert(d, frame)
    got = _convert(iw, ih, 3, frame, 1, in_fmt=ob.FMT[in_fmt], out_fmt=ob.FMT[out_fmt], matrix=4, rng=1)
    assert np.array_equal(got, want), _explain(got, want, iw // 2)


@pytest.mark.parametrize("in_fmt", ["I420", "YV12"])
@pytest.mark.parametrize("size,method,out_fmt", [((64, 48), 9, "RGBA"), ((488, 264), 3, "BGRA"), ((496, 280), 6, "ARGB"),
                                                 ((1000, 520), 3, "xBGR"), ((1928, 1096), 3, "BGRA"), ((3840, 2160), 3, "BGRA")],
                         ids=lambda v: "%dx%d" % v if isinstance(v, tuple) else str(v))
def test_planar_input_runs_the_second_form(cuda_device, in_fmt, size, method, out_fmt):
    """I420 / YV12 input on the exact-2:1 kernel (vcs_lanczos2_v2_kernel<PLANAR>): separate U and V planes with their own
    strides, every tile class (left / right border table path, top / bottom clamped rows, interior), byte orders with and
    without a compile-time selector - against the oracle"""
    import gstreamer_b200 as g
    from gstreamer_b200 import _lib
    iw, ih = size
    d = ob.vcs_desc(iw, ih, iw // 2, ih // 2, method, in_fmt=ob.FMT[in_fmt], out_fmt=ob.FMT[out_fmt], site=2)
    frame = ob.i420_random_frame(iw, ih, seed=iw + method)
    want = ob.oracle_vcs_convert(d, frame)
    got = _convert(iw, ih, method, frame, 1, in_fmt=ob.FMT[in_fmt], out_fmt=ob.FMT[out_fmt])
    assert np.array_equal(got, want), _explain(got, want, iw // 2)
    el = g.CudaVideoConvertScale(add_borders=False, method=method)
    el.set_info(g.VideoInfo(ob.FMT[in_fmt], iw, ih).set_colorimetry(chroma_site=2), g.VideoInfo(ob.FMT[out_fmt], iw // 2, ih // 2))
    assert _lib.lib.b200_vcs_kernel_name(el._h) == b"vcs_lanczos2_v2_kernel"


def test_specialised_equals_generic_on_structured_content(cuda_device):
    iw, ih = 3840, 2160
    frame = ob.nv12_smpte_like_frame(iw, ih, 4)
    a = _convert(iw, ih, 3, frame, 0)
    b = _convert(iw, ih, 3, frame, 1)
    assert np.array_equal(a, b), _explain(b, a, iw // 2)


def test_extreme_values(cuda_device):
    """all-0 / all-255 / checkerboard inputs drive the FIR overshoot into both saturations"""
    iw, ih = 512, 288
    st = iw
    for pat in range(3):
        frame = np.zeros(st * ih * 3 // 2, dtype=np.uint8)
        if pat == 1:
            frame[:] = 255
        elif pat == 2:
            y = frame[: st * ih].reshape(ih, st)
            y[::2, ::2] = 255
            y[1::2, 1::2] = 255
            frame[st * ih:] = np.tile(np.array([0, 255, 255, 0], dtype=np.uint8), st * ih // 8)
        d = ob.vcs_desc(iw, ih, iw // 2, ih // 2, 3, site=2)
        want = ob.oracle_vcs_convert(d, frame)
        got = _convert(iw, ih, 3, frame, 1)
        assert np.array_equal(got, want), f"pattern {pat}: " + _explain(got, want, iw // 2)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("size", [(488, 264), (1928, 1096), (3840, 2160)], ids=lambda s: "%dx%d" % s)
def test_tile_shape_variants(cuda_device, monkeypatch, variant, size):
    """the tuning knob B200_L2_VARIANT picks tile shape / residency; every variant must be exact"""
    monkeypatch.setenv("B200_L2_VARIANT", str(variant))
    iw, ih = size
    d = ob.vcs_desc(iw, ih, iw // 2, ih // 2, 3, site=2)
    frame = ob.nv12_random_frame(iw, ih, seed=variant + iw)
    want = ob.oracle_vcs_convert(d, frame)
    got = _convert(iw, ih, 3, frame, 1)
    assert np.array_equal(got, want), _explain(got, want, iw // 2)
