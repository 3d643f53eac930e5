"""Host-side mirror of the reference audioresample interface over the C-ABI.

  CudaAudioResample ~ the `audioresample` element: `quality` property (gstaudioresample.c:68),
                      set_caps -> resampler setup (:398-460), transform (:885-957),
                      reset on flush/discont (:462, :907-915), drain (:590-662)
The arithmetic (filter design on the host, polyphase FIR on the device) lives in libb200dsp.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check, B200Error
from .video import _ptr, _stream


class AudioFormat:
    """GstAudioFormat values of the sample formats the element resamples natively (audio-converter.c:700-727)"""
    S16LE, S32LE, F32LE, F64LE = 4, 12, 28, 30
    DTYPE = {4: np.int16, 12: np.int32, 28: np.float32, 30: np.float64}


class CudaAudioResample:
    # the stock element's remaining properties (gstaudioresample.c:153-186) with their nicks.  Implemented: both
    # windowed-sinc methods, every filter mode, every table interpolation, and the nearest / linear / cubic methods
    # (device-verified: tests/test_ars_options_gpu.py).
    # sinc-filter-auto-threshold never reaches the reference's resampler (UINT stored, INT read: SURVEY A.10)
    REST_DEFAULTS = {"resample_method": "kaiser", "sinc_filter_mode": "auto", "sinc_filter_interpolation": "cubic"}
    METHODS = {"nearest": 1, "linear": 2, "cubic": 3, "blackman-nuttall": 4, "kaiser": 5}
    FILTER_MODES = {"interpolated": 1, "full": 2, "auto": 3}
    INTERPOLATIONS = {"none": 1, "linear": 2, "cubic": 3}

    def __init__(self, quality=4, cuda_device_id=0, format=AudioFormat.F32LE, sinc_filter_auto_threshold=1048576, **rest):
        unknown = set(rest) - set(self.REST_DEFAULTS)
        if unknown:
            raise TypeError(f"no such property: {sorted(unknown)}")
        self.rest = dict(self.REST_DEFAULTS, **rest)
        self.sinc_filter_auto_threshold = sinc_filter_auto_threshold
        self.quality = quality
        self.format = format
        self.cuda_device_id = cuda_device_id
        self._h = None
        self.passthrough = False
        self.in_rate = self.out_rate = self.channels = None

    # GstBaseTransformClass::set_caps (interleaved samples of self.format)
    def set_caps(self, in_rate, out_rate, channels):
        # gst_audio_resample_update_state (gstaudioresample.c:398-437): same format and channel count on a live resampler
        # = a rate change, the converter is UPDATED (history and phase survive); anything else builds a new one
        if self._h is not None and channels == getattr(self, "channels", None) and in_rate != out_rate:
            self.update_rates(in_rate, out_rate)
            self.in_rate, self.out_rate = in_rate, out_rate
            return True
        self._free()
        # gst_audio_resample_set_caps: equal rates put the base transform in pass-through mode (buffers are forwarded
        # untouched and transform() is never called)
        self.passthrough = in_rate == out_rate
        if self.passthrough:
            self.in_rate, self.out_rate, self.channels = in_rate, out_rate, channels
            return True
        method = self.METHODS[self.rest["resample_method"]]
        mode = self.FILTER_MODES[self.rest["sinc_filter_mode"]]
        interp = self.INTERPOLATIONS[self.rest["sinc_filter_interpolation"]]
        cfg = _lib.ArsConfigC()
        cfg.in_rate, cfg.out_rate, cfg.channels, cfg.quality = in_rate, out_rate, channels, self.quality
        cfg.format = int(self.format)
        cfg.resample_method, cfg.sinc_filter_mode, cfg.sinc_filter_interpolation = method, mode, interp
        h = C.c_void_p()
        check(lib.b200_ars_create(C.byref(cfg), self.cuda_device_id, C.byref(h)), "b200_ars_create")
        self._h = h
        self.in_rate, self.out_rate, self.channels = in_rate, out_rate, channels
        return True

    def get_out_frames(self, in_frames):
        return lib.b200_ars_get_out_frames(self._h, in_frames)

    def get_in_frames(self, out_frames):
        return lib.b200_ars_get_in_frames(self._h, out_frames)

    @property
    def max_latency(self):
        return lib.b200_ars_get_max_latency(self._h)

    # GstBaseTransformClass::transform — device buffers, F32 interleaved; inbuf None = drain zeros
    def transform(self, inbuf, in_frames, outbuf, out_capacity, stream=None):
        if self._h is None:
            raise B200Error(-1, "transform() in pass-through mode (equal rates): the buffer is forwarded as it is")
        n = C.c_size_t()
        check(lib.b200_ars_process(self._h, _ptr(inbuf), in_frames, _ptr(outbuf), out_capacity, C.byref(n),
                                   _stream(stream)), "b200_ars_process")
        return n.value

    # system-memory peers: host buffers in, host buffers out (pinned ring + side streams inside the library)
    def transform_host(self, in_ptr, in_frames, out_ptr, out_capacity, wait=True):
        """wait=False: queued (b200_ars_process_host_submit); host_wait(k) later returns when all but the k latest are done"""
        n = C.c_size_t()
        fn = lib.b200_ars_process_host if wait else lib.b200_ars_process_host_submit
        check(fn(self._h, in_ptr, in_frames, out_ptr, out_capacity, C.byref(n)), "b200_ars_process_host")
        return n.value

    def host_wait(self, keep_in_flight=0):
        check(lib.b200_ars_process_host_wait(self._h, keep_in_flight), "b200_ars_process_host_wait")

    def reset(self):
        check(lib.b200_ars_reset(self._h), "b200_ars_reset")

    def update_rates(self, in_rate, out_rate):
        """new caps with the same format / channels on a live stream (gst_audio_resample_update_state): keeps history"""
        check(lib.b200_ars_update(self._h, in_rate, out_rate), "b200_ars_update")

    def plan_info(self):
        info = _lib.ArsPlanInfoC()
        check(lib.b200_ars_get_plan_info(self._h, C.byref(info)))
        return info

    def phase_taps(self, phase):
        n = self.plan_info().n_taps
        t = np.zeros(n, dtype=np.float32)
        check(lib.b200_ars_get_phase_taps(self._h, phase, t.ctypes.data, n))
        return t

    def _free(self):
        if self._h:
            lib.b200_ars_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass
