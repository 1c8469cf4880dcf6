"""One run of a single 60 s 96 kHz 8-channel low-latency stream (BASELINE configs[4] shape): the ncu target for chain_wide_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bauklank_audio_stretch_b200 as bs
g = torch.Generator(device="cuda").manual_seed(1)
sr8, D8 = 96000, float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
clip8 = (0.1 * torch.randn((8, int(D8 * sr8)), device="cuda", generator=g)).contiguous()
eng = bs.BatchStretch(8, sr8, block_samples=960, interval_samples=240, split_computation=True)
outs = eng.plan([clip8], [bs.KioskDrive(int(D8 * sr8), [bs.segment(rate=1.0, semitones=2.0, formant_semitones=3.0, formant_compensation=True, formant_base_hz=0.0)])])
eng.run(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.run(); b.record(); torch.cuda.synchronize()
print("%.1f ms, %.0f x real-time, chunk %d blocks" % (a.elapsed_time(b), D8 / (a.elapsed_time(b) / 1e3), eng.chunk_blocks()))
