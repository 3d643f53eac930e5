#!/usr/bin/env python
"""tools/emu_fuzz_audio.py [seconds] [seed] — randomised differential run of the resampler: the reference build
(oracle/_ref, when present) against the oracle, and the oracle against the EMULATED product kernels (tests/cudaemu: the
kernel sources compiled for the host).  Random rate pairs (small, coprime-ish and odd ones included), channel counts,
sample formats, quality, every resample-method / sinc-filter-mode / sinc-filter-interpolation value, random buffer
sequences with tiny buffers and a final drain.  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "cudaemu"))

from oracle import bindings as ob   # noqa: E402

RATES = [8000, 11025, 16000, 22050, 32000, 44100, 48000, 88200, 96000, 192000]


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    import build_emu
    from gstreamer_b200 import _lib
    emu = C.CDLL(build_emu.build())
    for name, (res, args) in _lib._SIGS.items():
        if hasattr(emu, name):
            fn = getattr(emu, name)
            fn.restype, fn.argtypes = res, args
    o = ob.oracle()
    try:
        r = ob.ref()
    except Exception:
        r = None
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < budget:
        k = rng.random()
        if k < 0.5:
            a, b = (int(v) for v in rng.choice(RATES, 2))
        elif k < 0.8:
            a, b = (int(v) for v in rng.integers(1, 400, 2))
        else:
            a = int(rng.choice(RATES))
            b = a + int(rng.integers(-3, 4))
        if b <= 0:
            continue
        fmt = str(rng.choice(["F32", "S16", "S32", "F64"]))
        ofmt, gfmt, dt, _ = ob.AUDIO_FORMATS[fmt]
        ch = int(rng.choice([1, 2, 3, 6, 33, 66])) if rng.random() < 0.8 else int(rng.integers(1, 140))
        q = int(rng.integers(0, 11))
        method, mode, interp = int(rng.integers(0, 5)), int(rng.integers(0, 3)), int(rng.integers(0, 3))
        # keep the emulated run affordable: the tap count grows with the decimation ratio
        if a > 12 * b and q > 4:
            q = int(rng.integers(0, 4))
        ho = o.oracle_ars_new_opts(a, b, ch, q, ofmt, method, mode, interp)
        if not ho:
            continue
        info = [C.c_int() for _ in range(6)]
        o.oracle_ars_info(ho, *[C.byref(v) for v in info])
        if info[0].value > 2000 or info[0].value * ch > 60000:
            o.oracle_ars_free(ho)
            continue
        hr = r.ref_ars_new_opts(a, b, ch, q, gfmt, method, mode, interp) if r else None
        cfg = _lib.ArsConfigC()
        cfg.in_rate, cfg.out_rate, cfg.channels, cfg.quality, cfg.format = a, b, ch, q, gfmt
        cfg.resample_method, cfg.sinc_filter_mode, cfg.sinc_filter_interpolation = method + 1, mode + 1, interp + 1
        h = C.c_void_p()
        st = emu.b200_ars_create(C.byref(cfg), 0, C.byref(h))
        desc = (a, b, ch, q, fmt, method, mode, interp)
        if st != 0:
            print("CREATE", st, desc, flush=True)
            bad += 1
            o.oracle_ars_free(ho)
            continue
        bufs = [int(v) for v in rng.choice([1, 2, 7, 37, 100, 160, 480], int(rng.integers(2, 6)))] + [None]
        for nb in bufs:
            x = None
            if nb is None:
                nb = emu.b200_ars_get_max_latency(h)
            else:
                x = ob.audio_test_signal(rng, nb, ch, fmt)
            cap = int(nb * b / a) + 64
            # every buffer starts from the same fill: after a skip the reference can report frames it never writes
            want = np.full((cap, ch), 7, dtype=dt)
            px = x.ctypes.data if x is not None else None
            nw = o.oracle_ars_process_any(ho, px, nb, want.ctypes.data, cap)
            if hr:
                w2 = np.full((cap, ch), 7, dtype=dt)
                n2 = r.ref_ars_process(hr, px, nb, w2.ctypes.data, cap)
                if n2 != nw or w2.tobytes() != want.tobytes():
                    print("ORACLE != REF", desc, nb, flush=True)
                    bad += 1
                    break
            got = np.full((cap, ch), 7, dtype=dt)
            ng = C.c_size_t()
            st = emu.b200_ars_process(h, px, nb, got.ctypes.data, cap, C.byref(ng), None)
            if st != 0 or ng.value != nw or got.tobytes() != want.tobytes():
                print("EMU != ORACLE", st, desc, nb, ng.value, nw, flush=True)
                bad += 1
                break
        emu.b200_ars_destroy(h)
        o.oracle_ars_free(ho)
        if hr:
            r.ref_ars_free(hr)
        n += 1
    print(f"{n} configurations, {bad} mismatches, seed {seed}, reference {'present' if r else 'absent'}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
