"""The C-ABI shared library: loads, exports every symbol include/b200dsp.h declares, and fails
loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200dsp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import gstreamer_b200 as g
    lib = C.CDLL(g.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert not g.MISSING_SYMBOLS
    # the python binding table and the header agree
    assert sorted(g.EXPORTED_SYMBOLS) == declared


def test_version_and_strerror():
    import gstreamer_b200 as g
    assert g.lib.b200_version() >= 1
    assert b"no CPU fallback" in g.lib.b200_strerror(-3)
    assert g.lib.b200_strerror(0) == b"ok"


def test_no_cpu_fallback_without_device():
    """device = -1 builds the host plan only; every compute entry point must refuse"""
    import gstreamer_b200 as g
    from gstreamer_b200.audio import CudaAudioResample
    el = g.CudaVideoConvertScale(add_borders=False, method=3, cuda_device_id=-1)
    el.set_info(g.VideoInfo(23, 64, 48), g.VideoInfo(12, 32, 24))
    buf = np.zeros(64 * 48 * 2, dtype=np.uint8)
    with pytest.raises(g.B200Error) as e:
        el.transform_frame(buf.ctypes.data, buf.ctypes.data)
    assert e.value.status == -3
    with pytest.raises(g.B200Error) as e:
        el.transform_host_frames([buf.ctypes.data], [buf.ctypes.data])
    assert e.value.status == -3
    rs = CudaAudioResample(cuda_device_id=-1)
    rs.set_caps(48000, 44100, 2)
    with pytest.raises(g.B200Error) as e:
        rs.transform(buf.ctypes.data, 10, buf.ctypes.data, 10)
    assert e.value.status == -3


def test_argument_validation_matches_caps_ranges():
    """caps allow [1, 32767] (gstvideoconvertscale.c:168-169); unsupported formats are refused the
    way set_info() returning FALSE would be"""
    import gstreamer_b200 as g
    lib = g.lib
    from gstreamer_b200 import _lib
    ii, oi = _lib.VideoInfoC(), _lib.VideoInfoC()
    assert lib.b200_video_info_set_format(C.byref(ii), 23, 0, 10) == -1
    assert lib.b200_video_info_set_format(C.byref(ii), 2, 64, 48) == 0        # I420 is supported
    assert lib.b200_video_info_set_format(C.byref(ii), 16, 64, 48) == -2      # Y444: not on this path yet
    assert lib.b200_video_info_set_format(C.byref(ii), 23, 64, 48) == 0
    assert lib.b200_video_info_set_format(C.byref(oi), 12, 32, 24) == 0
    h = C.c_void_p()
    oi.width = 40000
    assert lib.b200_vcs_create(C.byref(ii), C.byref(oi), None, -1, C.byref(h)) == -1
    oi.width = 32
    assert lib.b200_video_info_set_format(C.byref(oi), 2, 32, 24) == 0        # NV12 -> I420: chain + chroma down-sampling
    assert lib.b200_vcs_create(C.byref(ii), C.byref(oi), None, -1, C.byref(h)) == 0
    lib.b200_vcs_destroy(h)
    oi.color_matrix = 3                                                       # ... but not with a matrix stage (bt601 -> bt709)
    assert lib.b200_vcs_create(C.byref(ii), C.byref(oi), None, -1, C.byref(h)) == -2
    assert lib.b200_video_info_set_format(C.byref(oi), 23, 32, 24) == 0       # NV12 -> NV12: plane scaling, accepted
    assert lib.b200_vcs_create(C.byref(ii), C.byref(oi), None, -1, C.byref(h)) == 0
    lib.b200_vcs_destroy(h)
    assert lib.b200_video_info_set_format(C.byref(oi), 12, 32, 24) == 0
    ii.stride[0] = 10                                                         # stride < width
    assert lib.b200_vcs_create(C.byref(ii), C.byref(oi), None, -1, C.byref(h)) == -1
    # compositor / resampler argument checks
    hc = C.c_void_p()
    assert lib.b200_comp_create(23, 64, 48, -1, C.byref(hc)) == 0          # NV12 output: b200_comp_blend_yuv
    lib.b200_comp_destroy(hc)
    assert lib.b200_comp_create(7, 64, 48, -1, C.byref(hc)) == -2           # RGBx: no alpha, unsupported
    assert lib.b200_comp_create(11, 0, 48, -1, C.byref(hc)) == -1
    cfg = _lib.ArsConfigC()
    cfg.in_rate, cfg.out_rate, cfg.channels, cfg.quality = 48000, 44100, 2, 11
    assert lib.b200_ars_create(C.byref(cfg), -1, C.byref(hc)) == -1


def test_default_layout_matches_gst_video_info():
    """b200_video_info_set_format == gst_video_info_set_format for the supported formats
    (video-info.c:1053-1063 NV12, :890-894 packed RGB)"""
    import gstreamer_b200 as g
    for (w, h) in [(3840, 2160), (1920, 1080), (641, 481), (1, 1), (33, 17)]:
        i = g.VideoInfo(g.VideoFormat.NV12, w, h)
        st = (w + 3) & ~3
        assert i.stride[:2] == [st, st]
        assert i.offset[:2] == [0, st * ((h + 1) & ~1)]
        assert i.size == st * ((h + 1) & ~1) * 3 // 2
        o = g.VideoInfo(g.VideoFormat.BGRA, w, h)
        assert o.stride[0] == 4 * w and o.size == 4 * w * h
