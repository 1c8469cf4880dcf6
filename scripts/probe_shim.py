#!/usr/bin/env python
"""Real-time factor of the 18-call compat shim (Part 1 of the C ABI): one engine instance, host buffers, one block per
launch, driven quantum by quantum like the worklet (seek + process(0, 128)).  It is the literal drop-in, not the
throughput path; this says what it costs."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bauklank_audio_stretch_b200 as bs   # noqa: E402


def drive(e, clip, sr, n_out, rate, preset=None, block=None):
    ch = clip.shape[0]
    if block:
        e.configure(ch, *block); e.reset()
    elif preset == "cheaper":
        e.presetCheaper(ch, float(sr))
    else:
        e.presetDefault(ch, float(sr))
    buf = e.inputLatency() + e.outputLatency()
    e.setBuffers(ch, buf)
    out = np.zeros((ch, n_out), np.float32)
    pos = k = 0
    t0 = time.perf_counter()
    while pos < n_out:
        q = min(128, n_out - pos)
        e.setTransposeSemitones(3.0, 8000.0 / sr); e.setFormantSemitones(0.0, False); e.setFormantBase(0.0)
        end = int(np.floor(((k * 128) / sr + e.outputLatency() / sr) * rate * sr + e.inputLatency() + 0.5))
        ins, _ = e.io_views()
        ins[:] = 0
        lo, hi = max(end - buf, 0), min(end, clip.shape[1])
        if hi > lo:
            ins[:, lo - (end - buf):hi - (end - buf)] = clip[:, lo:hi]
        e.seek(buf, rate); e.process(0, q)
        _, outs = e.io_views()
        out[:, pos:pos + q] = outs[:, :q]
        pos += q; k += 1
    return out, time.perf_counter() - t0


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for name, sr, ch, kw in (("presetDefault 48 kHz stereo", 48000, 2, dict(preset="default")),
                             ("presetCheaper 48 kHz stereo", 48000, 2, dict(preset="cheaper")),
                             ("kiosk blockMs 200 (9600/2400, split)", 48000, 2, dict(block=(9600, 2400, 1))),
                             ("96 kHz 8-ch block 960/240 split", 96000, 8, dict(block=(960, 240, 1)))):
        clip = (0.2 * rng.standard_normal((ch, 6 * sr))).astype(np.float32)
        e = bs.StretchEngine(seed=1)
        drive(e, clip, sr, sr // 2, 0.8, **kw)                      # warm-up (library load, first launches)
        _, dt = drive(e, clip, sr, 5 * sr, 0.8, **kw)
        print("shim %-40s %6.2f x real-time (%.0f ms for 5 s of output, %d blocks, 8+ kernel launches and one D2H read per block)" % (
            name, 5.0 / dt, 1e3 * dt, (5 * sr) // e.intervalSamples()))
