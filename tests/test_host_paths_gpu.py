"""System-memory entry points of the compositor and the audio resampler (b200_comp_blend_host*, b200_ars_process_host*):
pinned host buffers in and out, a ring of device slots and three side streams inside the library.  Same bytes as the
device-memory path / the oracle, with several submissions in flight."""
import numpy as np
import pytest

from oracle import bindings as ob

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", [23, 2, 20, 45, 88], ids=["NV12", "I420", "Y444", "I422_10LE", "Y444_16LE"])
def test_compositor_yuv_host_frames_pipelined(cuda_device, fmt):
    """b200_comp_blend_yuv_host_submit / _wait: planar / semi-planar YUV pads and destination in pinned host memory, more
    submissions than ring slots, the destination's stride padding comes back untouched"""
    import gstreamer_b200 as g
    from gstreamer_b200.compositor import CudaCompositor
    o = ob.oracle()
    W, H, bg = 322, 201, 2
    rng = np.random.default_rng(fmt)
    specs = [(161, 120, -10, 5, 1.0, 1), (200, 91, 100, 60, 0.5, 1), (64, 64, 250, 150, 0.7, 2), (50, 40, 10, 150, 0.0, 1)]
    comp = CudaCompositor(fmt, W, H, bg)
    for (w, h, x, y, a, op) in specs:
        comp.request_pad(w, h, xpos=x, ypos=y, alpha=a, operator=op)
    sz = o.oracle_compositor_yuv_size(fmt, W, H)
    outs, wants, keep = [], [], []
    for n in range(6):
        bufs = [g.PinnedBuffer(o.oracle_compositor_yuv_size(fmt, w, h)) for (w, h, *_r) in specs]
        opads = (ob.OraclePad * len(specs))()
        for k, ((w, h, x, y, a, op), b) in enumerate(zip(specs, bufs)):
            b.array[:] = rng.integers(0, 256, b.array.size, dtype=np.uint8)
            p = opads[k]
            p.data, p.width, p.height, p.stride, p.xpos, p.ypos, p.alpha, p.op = b.ptr, w, h, 0, x, y, a, op
        want = np.full(sz, 0x33, dtype=np.uint8)
        assert o.oracle_compositor_yuv(fmt, want.ctypes.data, W, H, bg, 1, opads, len(specs)) == 0
        out = g.PinnedBuffer(sz)
        out.array[:] = 0x33
        comp.aggregate_host_frames_yuv(out.ptr, [b.ptr for b in bufs], wait=False)
        comp.host_wait(keep_in_flight=1)
        if n > 0:
            assert np.array_equal(outs[-1].array, wants[-1]), f"frame {n - 1}"
        keep.append(bufs); outs.append(out); wants.append(want)
    comp.host_wait(0)
    assert np.array_equal(outs[-1].array, wants[-1])


def test_compositor_host_frames_pipelined(cuda_device):
    import gstreamer_b200 as g
    from gstreamer_b200.compositor import CudaCompositor
    W, H, fmt, bg = 320, 200, 11, 0
    rng = np.random.default_rng(3)
    specs = [(160, 120, -10, 5, 1.0, 1), (200, 90, 100, 60, 0.5, 1), (64, 64, 250, 150, 0.7, 2), (50, 40, 10, 150, 0.0, 1)]
    comp = CudaCompositor(fmt, W, H, bg)
    for (w, h, x, y, a, op) in specs:
        comp.request_pad(w, h, xpos=x, ypos=y, alpha=a, operator=op)
    frames, outs, wants = [], [], []
    for n in range(7):                                   # more submissions than ring slots
        bufs = [g.PinnedBuffer(w * h * 4) for (w, h, *_r) in specs]
        opads = (ob.OraclePad * len(specs))()
        for k, ((w, h, x, y, a, op), b) in enumerate(zip(specs, bufs)):
            b.array[:] = rng.integers(0, 256, w * h * 4, dtype=np.uint8)
            p = opads[k]
            p.data, p.width, p.height, p.stride, p.xpos, p.ypos, p.alpha, p.op = b.ptr, w, h, w * 4, x, y, a, op
        want = np.zeros((H, W, 4), dtype=np.uint8)
        assert ob.oracle().oracle_compositor(fmt, want.ctypes.data, W, H, W * 4, bg, opads, len(specs)) == 0
        out = g.PinnedBuffer(W * H * 4)
        out.array[:] = 0x33
        comp.aggregate_host_frames(out.ptr, [b.ptr for b in bufs], wait=False)
        comp.host_wait(keep_in_flight=1)                 # frame n-1 is complete while frame n is in flight
        if n > 0:
            assert np.array_equal(outs[-1].array.reshape(H, W, 4), wants[-1]), f"frame {n - 1}"
        frames.append(bufs); outs.append(out); wants.append(want)
    comp.host_wait(0)
    assert np.array_equal(outs[-1].array.reshape(H, W, 4), wants[-1])
    # the synchronous form
    comp.aggregate_host_frames(outs[0].ptr, [b.ptr for b in frames[3]])
    assert np.array_equal(outs[0].array.reshape(H, W, 4), wants[3])


@pytest.mark.parametrize("fmt_name,dtype", [("F32LE", np.float32), ("S16LE", np.int16)])
def test_audio_host_stream_of_buffers(cuda_device, fmt_name, dtype):
    """a stream cut into uneven buffers through the host path == the same stream through the device path"""
    import torch
    import gstreamer_b200 as g
    from gstreamer_b200.audio import CudaAudioResample, AudioFormat
    ch, a, b = 6, 48000, 44100
    rng = np.random.default_rng(11)
    total = 20000
    if dtype == np.float32:
        x = (rng.standard_normal((total, ch)) * 0.3).astype(np.float32)
    else:
        x = rng.integers(-20000, 20000, (total, ch)).astype(np.int16)
    fmt = getattr(AudioFormat, fmt_name)
    ref = CudaAudioResample(quality=4, format=fmt)
    ref.set_caps(a, b, ch)
    host = CudaAudioResample(quality=4, format=fmt)
    host.set_caps(a, b, ch)
    tdt = torch.float32 if dtype == np.float32 else torch.int16
    cuts = [0, 1000, 1001, 5000, 5003, 12000, 20000]
    got_all, want_all, keep = [], [], []
    for lo, hi in zip(cuts, cuts[1:]):
        n = hi - lo
        cap = int(n * b / a) + 16
        xin = torch.from_numpy(x[lo:hi].copy()).cuda().reshape(-1)
        out = torch.zeros(cap * ch, dtype=tdt, device="cuda")
        k = ref.transform(xin, n, out, cap)
        torch.cuda.synchronize()
        want_all.append(out.cpu().numpy().reshape(cap, ch)[:k])
        hin, hout = g.PinnedBuffer(n * ch * x.itemsize), g.PinnedBuffer(cap * ch * x.itemsize)
        hin.array[:] = x[lo:hi].reshape(-1).view(np.uint8)
        kh = host.transform_host(hin.ptr, n, hout.ptr, cap, wait=False)
        host.host_wait(keep_in_flight=1)
        assert kh == k
        keep.append((hin, hout, kh, cap))
    host.host_wait(0)
    for (hin, hout, kh, cap), want in zip(keep, want_all):
        got = hout.array.view(dtype).reshape(cap, ch)[:kh]
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    # drain through the host path: in_host == NULL feeds silence
    tail = g.PinnedBuffer(64 * ch * x.itemsize)
    kd = host.transform_host(None, 40, tail.ptr, 64)
    outd = torch.zeros(64 * ch, dtype=tdt, device="cuda")
    kr = ref.transform(None, 40, outd, 64)
    torch.cuda.synchronize()
    assert kd == kr
    assert np.array_equal(tail.array.view(dtype).reshape(64, ch)[:kd].view(np.uint8), outd.cpu().numpy().reshape(64, ch)[:kr].view(np.uint8))
