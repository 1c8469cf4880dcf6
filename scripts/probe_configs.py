"""Throughput probes for the other BASELINE configurations (not the headline): presetCheaper batch, the kiosk's shipped
200 ms configuration, and a single long 96 kHz 8-channel stream with formant shift (configs[4] shape, shortened)."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bauklank_audio_stretch_b200 as bs


def run(name, eng, clips, drives, sr):
    outs = eng.plan(clips, drives)
    out_sec = sum(o.shape[1] for o in outs) / sr
    for _ in range(2):
        eng.run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.set_profiling(True)
    a.record(); eng.run(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    st = eng.kernel_stats()
    print("%-28s %8.1f ms  %9.0f x real-time  (%d blocks, chunk %d)  %s" % (
        name, ms, out_sec / (ms / 1e3), eng.total_blocks(), eng.chunk_blocks(), {k: round(v["ms"], 1) for k, v in st.items()}), flush=True)
    eng.close()


g = torch.Generator(device="cuda").manual_seed(1)
rng = np.random.default_rng(1)
S, D, sr = 256, 10.0, 48000
rates = np.exp(rng.uniform(math.log(0.5), math.log(2.0), S)); sts = rng.integers(-12, 13, S)
n_in = int(D * sr)
clips = [(0.1 * torch.randn((2, n_in), device="cuda", generator=g)).contiguous() for _ in range(S)]
drv = lambda: [bs.KioskDrive(int(n_in / rates[i]), [bs.segment(rate=float(rates[i]), semitones=float(sts[i]))]) for i in range(S)]
run("256x10s presetCheaper", bs.BatchStretch(2, sr, preset="cheaper"), clips, drv(), sr)
run("256x10s kiosk blockMs=200", bs.BatchStretch(2, sr, block_samples=9600, interval_samples=2400, split_computation=True), clips, drv(), sr)
drvf = [bs.KioskDrive(int(n_in / rates[i]), [bs.segment(rate=float(rates[i]), semitones=float(sts[i]), formant_semitones=3.0,
                                                         formant_compensation=True, formant_base_hz=0.0)]) for i in range(S)]
run("256x10s default + formants", bs.BatchStretch(2, sr, preset="default"), clips, drvf, sr)
del clips
sr8, D8 = 96000, 60.0
clip8 = (0.1 * torch.randn((8, int(D8 * sr8)), device="cuda", generator=g)).contiguous()
run("1 x 60s 96k 8ch lowlat fmt", bs.BatchStretch(8, sr8, block_samples=960, interval_samples=240, split_computation=True), [clip8],
    [bs.KioskDrive(int(D8 * sr8), [bs.segment(rate=1.0, semitones=2.0, formant_semitones=3.0, formant_compensation=True, formant_base_hz=0.0)])], sr8)
run("1 x 60s 96k 8ch presetDefault", bs.BatchStretch(8, sr8, preset="default"), [clip8],
    [bs.KioskDrive(int(D8 * sr8), [bs.segment(rate=1.0, semitones=2.0)])], sr8)
