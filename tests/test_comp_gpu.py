"""cudacompositor (single-pass kernel, through the C-ABI) vs the CPU oracle: bit-exact."""
import numpy as np
import pytest

from oracle import bindings as ob

pytestmark = pytest.mark.gpu

FORMATS = {"RGBA": 11, "BGRA": 12, "ARGB": 13, "ABGR": 14}


def _run(fmt, W, H, bg, pad_specs, seed=0, opaque_some=True):
    """pad_specs: list of (w, h, xpos, ypos, alpha, op)"""
    import torch
    from gstreamer_b200.compositor import CudaCompositor
    rng = np.random.default_rng(seed)
    comp = CudaCompositor(fmt, W, H, bg)
    opads = (ob.OraclePad * max(len(pad_specs), 1))()
    keep = []
    for k, (w, h, x, y, a, op) in enumerate(pad_specs):
        src = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        if opaque_some and k % 3 == 0:
            src[..., 3 if fmt in (11, 12) else 0] = 255
        keep.append(src)
        comp.request_pad(w, h, xpos=x, ypos=y, alpha=a, operator=op).set_frame(torch.from_numpy(src).cuda())
        p = opads[k]
        p.data, p.width, p.height, p.stride, p.xpos, p.ypos, p.alpha, p.op = src.ctypes.data, w, h, w * 4, x, y, a, op
    want = np.full((H, W, 4), 0x11, dtype=np.uint8)
    assert ob.oracle().oracle_compositor(fmt, want.ctypes.data, W, H, W * 4, bg, opads, len(pad_specs)) == 0
    out = torch.full((H * W * 4,), 0x22, dtype=torch.uint8, device="cuda")
    comp.aggregate_frames(out)
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(H, W, 4), want


@pytest.mark.parametrize("fmt", list(FORMATS))
@pytest.mark.parametrize("bg", [0, 1, 2, 3])
def test_random_layouts(cuda_device, fmt, bg):
    rng = np.random.default_rng(FORMATS[fmt] * 10 + bg)
    for trial in range(6):
        W, H = int(rng.integers(1, 200)), int(rng.integers(1, 120))
        specs = []
        for _ in range(int(rng.integers(0, 7))):
            specs.append((int(rng.integers(1, 150)), int(rng.integers(1, 100)), int(rng.integers(-60, W + 10)),
                          int(rng.integers(-60, H + 10)), float(rng.choice([0.0, 0.004, 0.3, 0.5, 0.999, 1.0])),
                          int(rng.integers(0, 3))))
        got, want = _run(FORMATS[fmt], W, H, bg, specs, seed=trial)
        assert np.array_equal(got, want), f"trial {trial} {W}x{H} specs {specs}"


def test_no_pads_draws_background(cuda_device):
    for bg in range(4):
        got, want = _run(12, 70, 33, bg, [])
        assert np.array_equal(got, want)


def test_more_pads_than_one_launch_chunk(cuda_device):
    """> 32 pads are split across launches that continue from the destination"""
    specs = [(40, 30, (k * 7) % 120 - 10, (k * 5) % 70 - 8, 0.6 if k % 2 else 1.0, 1) for k in range(45)]
    for bg in (0, 3):
        got, want = _run(11, 128, 80, bg, specs, seed=9)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("bg", [0, 3])
def test_baseline_config_c4(cuda_device, bg):
    """BASELINE configs[3]: 16 x 1080p RGBA pads at (k%4*640, k/4*360) -> 3840x2160, alpha .5/1"""
    specs = [(1920, 1080, (k % 4) * 640, (k // 4) * 360, 0.5 if k % 2 else 1.0, 1) for k in range(16)]
    got, want = _run(11, 3840, 2160, bg, specs, seed=4, opaque_some=False)
    assert np.array_equal(got, want)


def test_zorder_and_obscuring(cuda_device):
    """an opaque full-size pad on top hides everything below it (compositor.c:519-601 culls such
    pads; the single-pass kernel must give the same bytes without culling)"""
    import torch  # noqa: F401
    specs = [(64, 48, 0, 0, 1.0, 1), (64, 48, 0, 0, 1.0, 1)]
    got, want = _run(12, 64, 48, 0, specs, seed=1, opaque_some=False)
    assert np.array_equal(got, want)


def test_convert_pads(cuda_device):
    """GstVideoAggregatorConvertPad: NV12 / I420 inputs of other sizes are converted (default converter =
    mitchell cubic) to the output format at the pad's width x height, then blended"""
    import torch
    import gstreamer_b200 as g
    from gstreamer_b200.compositor import CudaCompositor, Background
    W, H = 320, 200
    comp = CudaCompositor(g.VideoFormat.BGRA, W, H, Background.CHECKER)
    specs = [("NV12", 23, 160, 90, 200, 120, -20, 10, 1.0), ("I420", 2, 352, 288, 176, 144, 100, 40, 0.6),
             ("NV12", 23, 128, 72, 128, 72, 180, 120, 0.9)]
    opads = (ob.OraclePad * len(specs))()
    keep = []
    for k, (name, fmt, iw, ih, pw, ph, x, y, a) in enumerate(specs):
        frame = ob.i420_random_frame(iw, ih, k) if name == "I420" else ob.nv12_random_frame(iw, ih, k)
        pad = comp.request_pad(pw, ph, xpos=x, ypos=y, alpha=a, in_info=g.VideoInfo(fmt, iw, ih))
        pad.set_frame(torch.from_numpy(frame).cuda())
        conv = ob.oracle_vcs_convert(ob.vcs_desc(iw, ih, pw, ph, 9, in_fmt=fmt, out_fmt=12), frame)
        keep.append(conv)
        opads[k].data, opads[k].width, opads[k].height, opads[k].stride = conv.ctypes.data, pw, ph, pw * 4
        opads[k].xpos, opads[k].ypos, opads[k].alpha, opads[k].op = x, y, a, 1
    want = np.zeros((H, W, 4), dtype=np.uint8)
    ob.oracle().oracle_compositor(12, want.ctypes.data, W, H, W * 4, 0, opads, len(specs))
    out = torch.zeros(H * W * 4, dtype=torch.uint8, device="cuda")
    comp.aggregate_frames(out)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().reshape(H, W, 4), want)


@pytest.mark.parametrize("fmt", [2, 3, 23, 24, 20, 18, 43, 73, 45, 75, 47, 77, 88],
                         ids=["I420", "YV12", "NV12", "NV21", "Y444", "Y42B", "I420_10LE", "I420_12LE", "I422_10LE", "I422_12LE", "Y444_10LE",
                              "Y444_12LE", "Y444_16LE"])
def test_420_output_random_layouts(cuda_device, fmt):
    """planar / semi-planar YUV output and pads (4:2:0, 4:2:2, 4:4:4; 8 bits and little-endian 10 / 12 / 16 bits): every plane
    byte equals the oracle's (random sizes, positions, alphas, operators, backgrounds, both ranges; more pads than one launch
    chunk in some trials)"""
    import torch
    import gstreamer_b200 as g
    from gstreamer_b200.compositor import CudaCompositor
    o = ob.oracle()
    rng = np.random.default_rng(fmt)
    for trial in range(25):
        W, H, bg, rg = int(rng.integers(1, 200)), int(rng.integers(1, 150)), int(rng.integers(0, 4)), int(rng.integers(0, 2))
        n = int(rng.integers(0, 6)) if trial % 5 else 30
        comp = CudaCompositor(fmt, W, H, bg)
        out_info = g.VideoInfo(fmt, W, H)
        out_info.set_colorimetry(range=2 if rg else 1)
        pads = (ob.OraclePad * max(n, 1))()
        keep = []
        for i in range(n):
            w, h = int(rng.integers(1, 120)), int(rng.integers(1, 90))
            a = rng.integers(0, 256, o.oracle_compositor_yuv_size(fmt, w, h), dtype=np.uint8)
            keep.append(a)
            x, y = int(rng.integers(-40, W + 5)), int(rng.integers(-40, H + 5))
            al, op = float(rng.choice([0.0, 0.3, 0.5, 0.999, 1.0, 0.004])), int(rng.integers(0, 3))
            pads[i].data, pads[i].width, pads[i].height, pads[i].stride = a.ctypes.data, w, h, 0
            pads[i].xpos, pads[i].ypos, pads[i].alpha, pads[i].op = x, y, al, op
            comp.request_pad(w, h, xpos=x, ypos=y, alpha=al, operator=op).set_frame(torch.from_numpy(a).cuda())
        sz = o.oracle_compositor_yuv_size(fmt, W, H)
        assert sz == out_info.size
        want = np.zeros(sz, dtype=np.uint8)
        assert o.oracle_compositor_yuv(fmt, want.ctypes.data, W, H, bg, rg, pads, n) == 0
        out = torch.zeros(sz, dtype=torch.uint8, device="cuda")
        comp.aggregate_frames(out, out_info=out_info)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, f"trial {trial}: {bad.size} bytes differ, first at {bad[:6]}: got {got[bad[:6]]} want {want[bad[:6]]}"


def test_420_pitched_planes(cuda_device):
    """common-pitch GstCudaMemory layout for the destination and a pad"""
    import torch
    import gstreamer_b200 as g
    from gstreamer_b200.compositor import CudaCompositor
    o = ob.oracle()
    W, H, w, h = 100, 60, 64, 40
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, o.oracle_compositor_yuv_size(23, w, h), dtype=np.uint8)
    pads = (ob.OraclePad * 1)()
    pads[0].data, pads[0].width, pads[0].height, pads[0].stride = src.ctypes.data, w, h, 0
    pads[0].xpos, pads[0].ypos, pads[0].alpha, pads[0].op = 21, 9, 0.5, 1
    want = np.zeros(o.oracle_compositor_yuv_size(23, W, H), dtype=np.uint8)
    o.oracle_compositor_yuv(23, want.ctypes.data, W, H, 0, 1, pads, 1)
    # destination: pitch 256, UV plane after 64 rows; pad: pitch 128
    dpitch, spitch = 256, 128
    di = g.VideoInfo(23, W, H).set_layout([dpitch, dpitch], [0, dpitch * 64])
    si = g.VideoInfo(23, w, h).set_layout([spitch, spitch], [0, spitch * 40])
    sbuf = np.zeros(spitch * 60, dtype=np.uint8)
    sbuf[: spitch * 40].reshape(40, spitch)[:, :64] = src[: 64 * 40].reshape(40, 64)
    sbuf[spitch * 40:].reshape(20, spitch)[:, :64] = src[64 * 40:].reshape(20, 64)
    comp = CudaCompositor(23, W, H, 0)
    comp.request_pad(w, h, xpos=21, ypos=9, alpha=0.5, in_info=si).set_frame(torch.from_numpy(sbuf).cuda())
    out = torch.full((dpitch * 96,), 0x55, dtype=torch.uint8, device="cuda")
    comp.aggregate_frames(out, out_info=di)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.array_equal(got[: dpitch * 60].reshape(60, dpitch)[:, :100], want[: 100 * 60].reshape(60, 100))
    assert np.array_equal(got[dpitch * 64: dpitch * 94].reshape(30, dpitch)[:, :100], want[100 * 60:].reshape(30, 100))
    assert (got[: dpitch * 60].reshape(60, dpitch)[:, 100:] == 0x55).all()
