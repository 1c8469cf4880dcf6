// The STFT kernels of the preset geometries (fft_fast.cuh) and their launchers.
//
// The template parameter FMA only names the instantiation: engine.cu (compiled with -fmad=false, the reference's arithmetic)
// instantiates FMA = 0; fft_fma.cu includes this header once more and instantiates FMA = 1 in a translation unit compiled
// with -fmad=true, where nvcc contracts a*b+c into one FFMA.  That second set is the opt-in tolerance mode
// (bsb_set_fft_fma): not bit-identical to the reference any more, measured against BASELINE's tolerance in DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include "kernels.cuh"
#include "fft_fast.cuh"

namespace bs {

// grid (slot, stream, {cur,prev} x channel) -- no index division
template <int LG, int OUTER, int FMA>
__global__ void __launch_bounds__(kFastNT, (FastOcc<LG, OUTER>::ctas)) analysis_fast_kernel(DevGeom g, DevTables T, const StreamDev *streams, const BlockRec *blocks,
                                                                  const Window *windows, long long slot0, int nSlots, cf *specIn) {
  extern __shared__ __align__(16) float sm[];
  const int slot = blockIdx.x, s = blockIdx.y, which = blockIdx.z & 1, c = blockIdx.z >> 1;
  const StreamDev sd = streams[s];
  const long long m = slot0 + slot;
  if (m >= sd.nBlocks) return;
  if (!(blocks[sd.blockBase + m].flags & kNew)) return;
  const Window w = windows[2 * (sd.blockBase + m) + which];
  cf *X = specIn + ((((size_t)s * nSlots + slot) * 2 + which) * g.C + c) * guard_pitch(g.B) + kGuard;
  fast_analyse<LG, OUTER>(g, T, sd.clip + (size_t)c * sd.clipLen, w, X, (cf *)sm, which == 1);
}
template <int LG, int OUTER, int FMA>
__global__ void __launch_bounds__(kFastNT, (FastOcc<LG, OUTER>::ctas)) isynth_fast_kernel(DevGeom g, DevTables T, const StreamDev *streams, long long slot0, int nSlots,
                                                                const cf *specOut, StateDev st) {
  extern __shared__ __align__(16) float sm[];
  const int slot = blockIdx.x, s = blockIdx.y, c = blockIdx.z;
  const StreamDev sd = streams[s];
  if (slot0 + slot >= sd.nBlocks) return;
  const size_t blk = (size_t)s * nSlots + slot;
  fast_synth<LG, OUTER>(g, T, specOut + (blk * g.C + c) * g.B, st.frames + (blk * g.C + c) * g.L, (cf *)sm);
}

// launchers: false = no specialised kernel for this geometry (or more streams than a grid dimension holds)
template <int FMA>
static bool launch_analysis_fast(const DevGeom &g, const DevTables &T, int S, int nSlots, cudaStream_t q, const StreamDev *streams, const BlockRec *blocks,
                                 const Window *windows, long long slot0, cf *specIn) {
  if (!fast_ok(g) || S > 65535) return false;
#define X_(LG, OUTER) if (g.inner == (1 << LG) && g.outer == OUTER) { \
    analysis_fast_kernel<LG, OUTER, FMA><<<dim3((unsigned)nSlots, (unsigned)S, (unsigned)(2 * g.C)), kFastNT, fast_smem_bytes<LG, OUTER>(), q>>>(g, T, streams, blocks, windows, slot0, nSlots, specIn); return true; }
  BS_FAST_GEOMS(X_)
#undef X_
  return false;
}
template <int FMA>
static bool launch_isynth_fast(const DevGeom &g, const DevTables &T, int S, int nSlots, cudaStream_t q, const StreamDev *streams, long long slot0,
                               const cf *specOut, const StateDev &st) {
  if (!fast_ok(g) || S > 65535) return false;
#define X_(LG, OUTER) if (g.inner == (1 << LG) && g.outer == OUTER) { \
    isynth_fast_kernel<LG, OUTER, FMA><<<dim3((unsigned)nSlots, (unsigned)S, (unsigned)g.C), kFastNT, fast_smem_bytes<LG, OUTER>(), q>>>(g, T, streams, slot0, nSlots, specOut, st); return true; }
  BS_FAST_GEOMS(X_)
#undef X_
  return false;
}
// (every instantiation always asks for the same, compile-time, amount: setting the limit again is harmless)
template <int FMA>
static bool fast_kernels_set_smem() {
  bool ok = true;
#define X_(LG, OUTER) ok = ok && \
    cudaFuncSetAttribute(analysis_fast_kernel<LG, OUTER, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes<LG, OUTER>()) == cudaSuccess && \
    cudaFuncSetAttribute(isynth_fast_kernel<LG, OUTER, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes<LG, OUTER>()) == cudaSuccess;
  BS_FAST_GEOMS(X_)
#undef X_
  return ok;
}

// the FMA-contracted set, defined in fft_fma.cu
bool launch_analysis_fast_fma(const DevGeom &g, const DevTables &T, int S, int nSlots, cudaStream_t q, const StreamDev *streams, const BlockRec *blocks,
                              const Window *windows, long long slot0, cf *specIn);
bool launch_isynth_fast_fma(const DevGeom &g, const DevTables &T, int S, int nSlots, cudaStream_t q, const StreamDev *streams, long long slot0,
                            const cf *specOut, const StateDev &st);
bool fast_kernels_set_smem_fma();

}  // namespace bs
