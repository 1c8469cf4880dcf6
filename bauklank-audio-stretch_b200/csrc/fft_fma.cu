// The preset-geometry STFT kernels once more, in a translation unit compiled with -fmad=true: nvcc contracts the
// butterflies' a*b+c into single FFMAs.  Opt-in (bsb_set_fft_fma): the output then differs from the reference in the
// last bits of every spectrum and is only held to BASELINE's tolerance (max|err| <= 1e-4, SNR >= 90 dB); the default path
// (engine.cu, -fmad=false) stays bit-identical.  Same source (fft_fast.cuh), different rounding.
#include "fft_kernels.cuh"

namespace bs {

bool launch_analysis_fast_fma(const DevGeom &g, const DevTables &T, int S, int nSlots, cudaStream_t q, const StreamDev *streams, const BlockRec *blocks,
                              const Window *windows, long long slot0, cf *specIn) {
  return launch_analysis_fast<1>(g, T, S, nSlots, q, streams, blocks, windows, slot0, specIn);
}
bool launch_isynth_fast_fma(const DevGeom &g, const DevTables &T, int S, int nSlots, cudaStream_t q, const StreamDev *streams, long long slot0,
                            const cf *specOut, const StateDev &st) {
  return launch_isynth_fast<1>(g, T, S, nSlots, q, streams, slot0, specOut, st);
}
bool fast_kernels_set_smem_fma() { return fast_kernels_set_smem<1>(); }

}  // namespace bs
