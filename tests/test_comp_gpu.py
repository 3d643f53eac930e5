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
