"""Host-side mirror of the reference compositor's interface over the C-ABI.

  CudaCompositor          ~ the `compositor` element: `background` property (compositor.c:742),
                            request sink pads, GstVideoAggregatorClass::aggregate_frames (:1739)
  CudaCompositorPad       ~ GstCompositorPad: xpos / ypos / width / height / alpha / operator / sizing-policy
                            (compositor.c:190-196, :678-716)
"""
import ctypes as C
import enum

from . import _lib
from ._lib import lib, check
from .video import VideoFormat, _ptr, _stream


class Background(enum.IntEnum):
    CHECKER = 0
    BLACK = 1
    WHITE = 2
    TRANSPARENT = 3


class Operator(enum.IntEnum):
    SOURCE = 0
    OVER = 1
    ADD = 2


def pad_output_size(src_w, src_h, pad_w, pad_h, sizing_policy="none", zero_size_is_unscaled=True):
    """_mixer_pad_get_output_size (compositor.c:289-417) for pixel aspect ratio 1/1 everywhere: the size a pad's picture
    takes in the output and the offset that centres it.  pad_w / pad_h are the pad's width / height properties (unset:
    the source size), sizing_policy "none" or "keep-aspect-ratio".  Returns (width, height, x_offset, y_offset)."""
    from math import gcd
    if zero_size_is_unscaled:
        pad_w = src_w if pad_w <= 0 else pad_w
        pad_h = src_h if pad_h <= 0 else pad_h
    else:
        pad_w = src_w if pad_w < 0 else pad_w
        pad_h = src_h if pad_h < 0 else pad_h
    if pad_w == 0 or pad_h == 0:
        return 0, 0, 0, 0
    if sizing_policy in ("none", 0):
        # the display-ratio round trip (gst_video_calculate_display_ratio, :330-352) is the identity at 1/1
        return pad_w, pad_h, 0, 0
    g1, g2 = gcd(src_w, src_h), gcd(pad_w, pad_h)
    num, den = src_w // g1, src_h // g1
    if (num, den) == (pad_w // g2, pad_h // g2):
        return pad_w, pad_h, 0, 0
    sh = pad_w * den // num                     # gst_util_uint64_scale_int (pad_width, den, num)
    if sh == 0:
        return 0, 0, 0, 0
    # gst_video_center_rect (gstvideosink.c:122-164), scaling: doubles, then truncation to int
    src_ratio, dst_ratio = pad_w / sh, pad_w / pad_h
    if src_ratio > dst_ratio:
        w, h = pad_w, int(pad_w / src_ratio)
        return w, h, 0, (pad_h - h) // 2
    if src_ratio < dst_ratio:
        w, h = int(pad_h * src_ratio), pad_h
        return w, h, (pad_w - w) // 2, 0
    return pad_w, pad_h, 0, 0


class CudaCompositorPad:
    """GstCompositorPad.  `width`/`height` are the pad properties (compositor.c:190-196): the size the
    pad's picture takes in the output.  With `in_info` the pad is a GstVideoAggregatorConvertPad: its
    input frames (NV12/NV21/I420/YV12 of any size) are converted to the aggregator's format and scaled to
    width x height in prepare_frame (gstvideoaggregator.c:479-570, :782-830) by a GstVideoConverter with
    DEFAULT options — cubic b = c = 1/3 (video-converter.c:791, video-resampler.c:63-64), i.e. exactly
    cudavideoconvertscale method=mitchell — before blending."""

    def __init__(self, width, height, stride=None, xpos=0, ypos=0, alpha=1.0, operator=Operator.OVER, in_info=None,
                 sizing_policy="none"):
        # sizing-policy (compositor.c:714): "keep-aspect-ratio" shrinks a converting pad's picture inside width x height
        # so that the source's display aspect ratio survives, centred by x_offset / y_offset
        self.x_offset = self.y_offset = 0
        self.sizing_policy = sizing_policy
        if in_info is not None:
            width, height, self.x_offset, self.y_offset = pad_output_size(in_info.width, in_info.height, width, height,
                                                                          sizing_policy)
        self.width, self.height = width, height
        self.stride = stride or width * 4
        self.xpos, self.ypos, self.alpha, self.operator = xpos, ypos, alpha, Operator(operator)
        self.frame = None          # the pad's prepared frame (device memory)
        self.in_info = in_info
        self._conv = self._converted = None

    def set_frame(self, frame):
        self.frame = frame
        return self

    def _prepare_frame(self, out_format, device, stream):
        """prepare_frame of a convert pad: returns the frame to blend"""
        if self.in_info is None or self.frame is None:
            return self.frame
        import torch
        from .video import CudaVideoConvertScale, VideoInfo, VideoScaleMethod
        if self._conv is None:
            self._conv = CudaVideoConvertScale(add_borders=False, method=VideoScaleMethod.MITCHELL, cuda_device_id=device)
            self._out_info = VideoInfo(out_format, self.width, self.height)
            self._conv.set_info(self.in_info, self._out_info)
            self._converted = torch.empty(self._out_info.size, dtype=torch.uint8, device=f"cuda:{device}")
            self.stride = self._out_info.stride[0]
        self._conv.transform_frame(self.frame, self._converted, stream)
        return self._converted


class CudaCompositor:
    def __init__(self, out_format, width, height, background=Background.CHECKER, cuda_device_id=0):
        self.background = Background(background)
        self.width, self.height, self.format = width, height, VideoFormat(out_format)
        self.sinkpads = []
        self.device = cuda_device_id
        h = C.c_void_p()
        check(lib.b200_comp_create(int(out_format), width, height, cuda_device_id, C.byref(h)),
              "b200_comp_create")
        self._h = h

    def request_pad(self, *args, **kw):
        pad = CudaCompositorPad(*args, **kw)
        self.sinkpads.append(pad)
        return pad

    YUV_FORMATS = (VideoFormat.I420, VideoFormat.YV12, VideoFormat.NV12, VideoFormat.NV21, VideoFormat.Y444, VideoFormat.Y42B,
                   VideoFormat.I420_10LE, VideoFormat.I420_12LE, VideoFormat.I422_10LE, VideoFormat.I422_12LE,
                   VideoFormat.Y444_10LE, VideoFormat.Y444_12LE, VideoFormat.Y444_16LE)

    def _aggregate_yuv(self, outbuf, out_info, stream):
        """4:2:0 output: every pad frame has the output's format (blend.c PLANAR_YUV_BLEND / NV_YUV_BLEND);
        a pad's plane layout comes from pad.in_info, default layout otherwise"""
        from .video import VideoInfo
        out_info = out_info or VideoInfo(self.format, self.width, self.height)
        pads = [p for p in self.sinkpads if p.frame is not None]
        arr = (_lib.CompPadYuvC * max(len(pads), 1))()
        keep = []
        for i, p in enumerate(pads):
            info = p.in_info or VideoInfo(self.format, p.width, p.height)
            keep.append(info)
            arr[i].data = _ptr(p.frame)
            arr[i].info = info.c
            arr[i].xpos, arr[i].ypos, arr[i].alpha, arr[i].op = p.xpos + p.x_offset, p.ypos + p.y_offset, p.alpha, int(p.operator)
        check(lib.b200_comp_blend_yuv(self._h, _ptr(outbuf), C.byref(out_info.c), int(self.background), arr, len(pads),
                                      _stream(stream)), "b200_comp_blend_yuv")

    # GstVideoAggregatorClass::aggregate_frames (outbuf is device memory)
    def aggregate_frames(self, outbuf, out_stride=None, stream=None, out_info=None):
        if self.format in self.YUV_FORMATS:
            return self._aggregate_yuv(outbuf, out_info, stream)
        pads = [p for p in self.sinkpads if p.frame is not None]
        arr = (_lib.CompPadC * max(len(pads), 1))()
        for i, p in enumerate(pads):
            arr[i].data = _ptr(p._prepare_frame(self.format, self.device, stream))
            arr[i].width, arr[i].height, arr[i].stride = p.width, p.height, p.stride
            arr[i].xpos, arr[i].ypos, arr[i].alpha, arr[i].op = p.xpos + p.x_offset, p.ypos + p.y_offset, p.alpha, int(p.operator)
        check(lib.b200_comp_blend(self._h, _ptr(outbuf), out_stride or self.width * 4, int(self.background),
                                  arr, len(pads), _stream(stream)), "b200_comp_blend")

    # system-memory peers: pad frames and the output are HOST buffers (addresses); see b200_comp_blend_host_submit
    def aggregate_host_frames(self, out_ptr, pad_ptrs, out_stride=None, wait=True):
        pads = [p for p in self.sinkpads]
        arr = (_lib.CompPadC * max(len(pads), 1))()
        for i, (p, ptr) in enumerate(zip(pads, pad_ptrs)):
            arr[i].data = ptr
            arr[i].width, arr[i].height, arr[i].stride = p.width, p.height, p.stride
            arr[i].xpos, arr[i].ypos, arr[i].alpha, arr[i].op = p.xpos + p.x_offset, p.ypos + p.y_offset, p.alpha, int(p.operator)
        fn = lib.b200_comp_blend_host if wait else lib.b200_comp_blend_host_submit
        check(fn(self._h, out_ptr, out_stride or self.width * 4, int(self.background), arr, len(pads)), "b200_comp_blend_host")

    def aggregate_host_frames_yuv(self, out_ptr, pad_ptrs, out_info=None, wait=True):
        """system-memory peers of a planar / semi-planar YUV compositor (b200_comp_blend_yuv_host[_submit])"""
        from .video import VideoInfo
        out_info = out_info or VideoInfo(self.format, self.width, self.height)
        pads = [p for p in self.sinkpads]
        arr = (_lib.CompPadYuvC * max(len(pads), 1))()
        keep = []
        for i, (p, ptr) in enumerate(zip(pads, pad_ptrs)):
            info = p.in_info or VideoInfo(self.format, p.width, p.height)
            keep.append(info)
            arr[i].data = ptr
            arr[i].info = info.c
            arr[i].xpos, arr[i].ypos, arr[i].alpha, arr[i].op = p.xpos + p.x_offset, p.ypos + p.y_offset, p.alpha, int(p.operator)
        fn = lib.b200_comp_blend_yuv_host if wait else lib.b200_comp_blend_yuv_host_submit
        check(fn(self._h, out_ptr, C.byref(out_info.c), int(self.background), arr, len(pads)), "b200_comp_blend_yuv_host")

    def host_wait(self, keep_in_flight=0):
        check(lib.b200_comp_blend_host_wait(self._h, keep_in_flight), "b200_comp_blend_host_wait")

    def __del__(self):
        try:
            if self._h:
                lib.b200_comp_destroy(self._h)
                self._h = None
        except Exception:
            pass
