"""Probe: relayed chain wavefront.  (a) one long presetDefault stream, (b) the headline batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bauklank_audio_stretch_b200 as bs
sr = 48000
g = torch.Generator(device="cuda").manual_seed(1)
def timeit(eng, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); eng.run(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
for D, chunk in ((120.0, 256), (120.0, 0)):
    clip = (0.1 * torch.randn((2, int(D * sr)), device="cuda", generator=g)).contiguous()
    eng = bs.BatchStretch(2, sr, preset="default")
    outs = eng.plan([clip], [bs.KioskDrive(int(D * sr), [bs.segment(rate=1.0, semitones=3.0)])], chunk_blocks=chunk)
    ms = timeit(eng)
    print("1 x %.0f s presetDefault chunk=%d: %.1f ms -> %.0f x real-time, launches %d, checksum %.6f" % (D, chunk, ms, D / (ms / 1e3), eng.launch_count(), float(outs[0].double().abs().sum())), flush=True)
    eng.close()
