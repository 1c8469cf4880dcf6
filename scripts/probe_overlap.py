"""Probe: the headline workload (256 x 60 s) with chunk pipelining off/on and several chunk lengths.
usage: python scripts/probe_overlap.py [streams] [seconds] [variants: e.g. 0:0,1:0,0:512]   (overlap:chunk)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bauklank_audio_stretch_b200 as bs

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
variants = sys.argv[3] if len(sys.argv) > 3 else "0:0,1:0"
sr = 48000
rng = np.random.default_rng(1)
rates = np.exp(rng.uniform(np.log(0.5), np.log(2.0), S))
sts = rng.integers(-12, 13, S)
n_in = int(D * sr)
g = torch.Generator(device="cuda").manual_seed(1)
clips = [(0.1 * torch.randn((2, n_in), device="cuda", generator=g)).contiguous() for _ in range(S)]
drives = [bs.KioskDrive(int(n_in / rates[i]), [bs.segment(rate=float(rates[i]), semitones=float(sts[i]))]) for i in range(S)]
out_sec = sum(d.n_out for d in drives) / sr
ref = None
for v in variants.split(","):
    ov, chunk = (int(x) for x in v.split(":"))
    eng = bs.BatchStretch(2, sr, preset="default", lib=bs.load_library(os.environ["BSLIB"]) if os.environ.get("BSLIB") else None)
    eng.set_overlap(bool(ov))
    outs = eng.plan(clips, drives, chunk_blocks=chunk)
    best = 1e9
    for it in range(4):
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); eng.run(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    cs = [float(o.double().abs().sum()) for o in outs[:8]]
    if ref is None: ref = cs
    print("overlap=%d chunk=%d (%d): %.1f ms -> %.0f x realtime, launches=%d, same=%s" % (
        ov, chunk, eng.chunk_blocks(), best, out_sec / (best / 1e3), eng.launch_count(), cs == ref), flush=True)
    if not ov:
        eng.set_profiling(True); eng.run(); torch.cuda.synchronize()
        print("   ", {k: round(x["ms"], 1) for k, x in eng.kernel_stats().items()}, flush=True)
    eng.close(); del outs, eng
    torch.cuda.empty_cache()
