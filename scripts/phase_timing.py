"""Debug: per-phase cycle split of the spectral kernel (needs the -DBS_PHASE_TIMING build in /tmp/bs_timing.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bauklank_audio_stretch_b200 as bs
lib = bs.load_library(sys.argv[1])
S, D = 64, 2.0
sr = 48000; n_in = int(D*sr)
clips = [(0.1*torch.randn((2, n_in), device="cuda")).contiguous() for _ in range(S)]
drives = [bs.KioskDrive(n_in, [bs.segment(rate=1.0, semitones=3.0)]) for _ in range(S)]
eng = bs.BatchStretch(2, sr, lib=lib); eng.plan(clips, drives); eng.run(); torch.cuda.synchronize()
import cuda.bindings.runtime as rt
names = ["rot+energy","esum","smooth","peaks","map","formants","S5","terms","chain"]
# read the device symbol through a helper exported by the lib
buf = (ctypes.c_ulonglong*16)()
lib.bs_debug_phase_cycles(buf)
tot = sum(buf[:9]); nb = eng.stream_blocks(0)
for n, v in zip(names, buf): print("%-12s %10.0f cycles/block  %5.1f%%" % (n, v/nb, 100*v/tot))
print("total %.0f cycles/block" % (tot/nb))
