"""The FMA-contracted STFT kernels (bsb_set_fft_fma) against the default bit-identical path: error over every golden case
that has a specialised geometry, and the time of both on bench.py's default workload.  Run on the GPU box:
    python scripts/measure_fft_fma.py > gpurun_out/fft_fma.txt"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bauklank_audio_stretch_b200 as bs  # noqa: E402
import cases  # noqa: E402


def main():
    print("# case, bit_identical, max|err|, min SNR over channels [dB]   (FMA build vs default build, same inputs)")
    worst_err, worst_snr = 0.0, float("inf")
    for name, case in cases.CASES.items():
        clip = cases.make_clip(case["clip"])
        outs = []
        for fma in (False, True):
            eng = cases.make_batch(bs, case, clip.shape[0])
            if not eng.fast_fft_active():
                eng.close(); outs = None; break
            eng.set_fft_fma(fma)
            o = eng.plan([torch.from_numpy(clip).cuda()], [cases.batch_drive(bs, case, clip.shape[1])])
            eng.run(); torch.cuda.synchronize()
            outs.append(o[0].cpu().numpy()); eng.close()
        if outs is None:
            print("%-32s (generic geometry: mode has no effect)" % name); continue
        same, err, snr = cases.compare(outs[1], outs[0])
        worst_err, worst_snr = max(worst_err, err), min(worst_snr, snr)
        print("%-32s %-5s %10.3g %8.1f" % (name, same, err, snr))
    print("# worst: max|err| %.3g, SNR %.1f dB  (BASELINE tolerance: 1e-4, 90 dB)" % (worst_err, worst_snr))


if __name__ == "__main__":
    main()
