"""cudaaudioresample with the element's resample-method / sinc-filter-mode / sinc-filter-interpolation properties at
non-default values (blackman-nuttall window, forced full / interpolated tables, linear or no table interpolation, and the
nearest / linear / cubic methods, which have their own small kernel).  Only the
host side changes (tap tables and the mode decision — checked against the oracle bit for bit in tests/test_host_plan.py) and
the device kernels are the ones the default configuration runs — except for linear interpolation in the interpolated filter
mode, whose two-row blend is its own device code."""
import os

import numpy as np
import pytest

from oracle import bindings as ob

pytestmark = [pytest.mark.gpu]

M = {"nearest": 0, "linear": 1, "cubic": 2, "blackman-nuttall": 3, "kaiser": 4}
MO = {"interpolated": 0, "full": 1, "auto": 2}
I = {"none": 0, "linear": 1, "cubic": 2}


@pytest.mark.parametrize("fmt", ["F32", "S16", "S32", "F64"])
@pytest.mark.parametrize("method,mode,interp", [("blackman-nuttall", "auto", "cubic"), ("blackman-nuttall", "full", "none"),
                                                ("kaiser", "full", "cubic"), ("kaiser", "full", "none"),
                                                ("kaiser", "interpolated", "cubic"), ("kaiser", "interpolated", "none"),
                                                ("blackman-nuttall", "interpolated", "cubic"), ("kaiser", "auto", "none"),
                                                ("kaiser", "full", "linear"), ("kaiser", "interpolated", "linear"),
                                                ("blackman-nuttall", "auto", "linear"),
                                                ("nearest", "auto", "cubic"), ("linear", "auto", "cubic"), ("cubic", "auto", "cubic")])
def test_method_and_filter_mode_properties(cuda_device, fmt, method, mode, interp, monkeypatch):
    import torch
    from gstreamer_b200.audio import CudaAudioResample
    ofmt, gfmt, dt, _ = ob.AUDIO_FORMATS[fmt]
    tdt = {np.float32: torch.float32, np.int16: torch.int16, np.int32: torch.int32, np.float64: torch.float64}[dt]
    o = ob.oracle()
    for (a, b, ch, q) in [(48000, 44100, 2, 4), (44100, 48000, 1, 6), (8000, 16000, 2, 0), (96000, 44100, 3, 8), (101, 99, 1, 10),
                          (3, 2, 3, 5), (48000, 8000, 40, 1)]:
        ho = o.oracle_ars_new_opts(a, b, ch, q, ofmt, M[method], MO[mode], I[interp])
        rs = CudaAudioResample(quality=q, format=gfmt, resample_method=method, sinc_filter_mode=mode,
                               sinc_filter_interpolation=interp)
        rs.set_caps(a, b, ch)
        rng = np.random.default_rng(ch + a)
        for n in [480, 480, 100, 1, 2000, 37, None]:
            x = None
            if n is None:
                n = rs.max_latency
            else:
                x = ob.audio_test_signal(rng, n, ch, fmt)
            cap = int(n * b / a) + 64
            want = np.zeros((cap, ch), dtype=dt)
            assert rs.get_out_frames(n) == o.oracle_ars_get_out_frames(ho, n)
            nw = o.oracle_ars_process_any(ho, x.ctypes.data if x is not None else None, n, want.ctypes.data, cap)
            out = torch.full((cap * ch,), 7, dtype=tdt, device="cuda")
            ng = rs.transform(torch.from_numpy(x).cuda() if x is not None else None, n, out, cap)
            torch.cuda.synchronize()
            got = out.cpu().numpy().reshape(cap, ch)
            assert ng == nw and got[:ng].tobytes() == want[:nw].tobytes(), (a, b, ch, q, n)
            assert (got[ng:] == 7).all()
        o.oracle_ars_free(ho)


@pytest.mark.parametrize("fmt", ["F32", "S16", "F64"])
def test_nearest_decimation_skip_quirk(cuda_device, fmt, monkeypatch):
    """tests/test_oracle_vs_ref.py::test_audio_nearest_decimation_skip_quirk on the device"""
    import torch
    from gstreamer_b200.audio import CudaAudioResample
    ofmt, gfmt, dt, _ = ob.AUDIO_FORMATS[fmt]
    tdt = {np.float32: torch.float32, np.int16: torch.int16, np.int32: torch.int32, np.float64: torch.float64}[dt]
    o = ob.oracle()
    for (a, b, ch) in [(48000, 11025, 3), (96000, 8000, 1), (400, 3, 2)]:
        for seed in range(3):
            rng = np.random.default_rng(seed)
            ho = o.oracle_ars_new_opts(a, b, ch, 4, ofmt, 0, 2, 2)
            rs = CudaAudioResample(quality=4, format=gfmt, resample_method="nearest")
            rs.set_caps(a, b, ch)
            for n in [int(v) for v in rng.choice([1, 2, 7, 37, 100, 160, 480], 6)]:
                x = ob.audio_test_signal(rng, n, ch, fmt)
                cap = int(n * b / a) + 64
                want = np.full((cap, ch), 7, dtype=dt)
                nw = o.oracle_ars_process_any(ho, x.ctypes.data, n, want.ctypes.data, cap)
                out = torch.full((cap * ch,), 7, dtype=tdt, device="cuda")
                ng = rs.transform(torch.from_numpy(x).cuda(), n, out, cap)
                torch.cuda.synchronize()
                assert ng == nw and out.cpu().numpy().tobytes() == want.tobytes(), (a, b, ch, seed, n)
            o.oracle_ars_free(ho)
