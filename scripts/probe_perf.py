"""Throughput probe: S streams x D seconds, presetDefault, kiosk drive, mixed rate/transpose (BASELINE config 3 shape)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bauklank_audio_stretch_b200 as bs

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
preset = sys.argv[3] if len(sys.argv) > 3 else "default"
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sr = 48000
rng = np.random.default_rng(1)
rates = np.exp(rng.uniform(np.log(0.5), np.log(2.0), S))
sts = rng.integers(-12, 13, S)
n_in = int(D * sr)
g = torch.Generator(device="cuda").manual_seed(1)
clips = [(0.1 * torch.randn((2, n_in), device="cuda", generator=g)).contiguous() for _ in range(S)]
drives = [bs.KioskDrive(int(n_in / rates[i]), [bs.segment(rate=float(rates[i]), semitones=float(sts[i]))]) for i in range(S)]
eng = bs.BatchStretch(2, sr, preset=preset)
t0 = time.time(); outs = eng.plan(clips, drives, chunk_blocks=chunk); t1 = time.time()
print("plan %.3fs blocks=%d chunk=%d" % (t1 - t0, eng.total_blocks(), eng.chunk_blocks()))
out_sec = sum(d.n_out for d in drives) / sr
for it in range(3):
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); eng.run(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("run %d: %.1f ms -> %.0f x realtime (out-sec %.0f) launches=%d" % (it, ms, out_sec / (ms / 1e3), out_sec, eng.launch_count()), eng.kernel_stats())
