"""A/B of the wide chain kernel's relay parameters on a single long 8-channel stream (BASELINE configs[4] shape, 60 s):
every library under bauklank-audio-stretch_b200/_variants/ (built with -DBS_WIDE_TILE / -DBS_WIDE_AHEAD / -DBS_WIDE_WARPS_MIN)."""
import glob, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bauklank_audio_stretch_b200 as bs

g = torch.Generator(device="cuda").manual_seed(1)
sr8, D8 = 96000, 60.0
clip8 = (0.1 * torch.randn((8, int(D8 * sr8)), device="cuda", generator=g)).contiguous()
ref = {}
for path in sorted(glob.glob(os.path.join(os.path.dirname(bs.__file__), "_variants", "*.so"))):
    lib = bs.load_library(path)
    for name, mk, seg in (("lowlat", lambda: bs.BatchStretch(8, sr8, block_samples=960, interval_samples=240, split_computation=True, lib=lib),
                           bs.segment(rate=1.0, semitones=2.0, formant_semitones=3.0, formant_compensation=True, formant_base_hz=0.0)),
                          ("default", lambda: bs.BatchStretch(8, sr8, preset="default", lib=lib), bs.segment(rate=1.0, semitones=2.0))):
        eng = mk()
        outs = eng.plan([clip8], [bs.KioskDrive(int(D8 * sr8), [seg])])
        for _ in range(2):
            eng.run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.run(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        chk = outs[0].double().abs().sum().item()
        same = ref.setdefault(name, chk) == chk
        print("%-24s %-8s %8.1f ms %7.0f x real-time  checksum %s" % (os.path.basename(path), name, ms, D8 / (ms / 1e3), "same" if same else "DIFFERENT"), flush=True)
        eng.close()
