// Host side of libbauklank_stretch.so: batched engine (Part 2 of include/bauklank_stretch.h) and kernel launchers.
// Built by nvcc for sm_100a (product) or, with -DBS_HOSTEMU, by g++ as the serial test emulation of the same code.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bauklank_stretch.h"
#include "kernels.cuh"
#include "fft_fast.cuh"

#ifndef BS_HOSTEMU
#include <cuda_runtime.h>
#include "chain.cuh"
#include "chain_wide.cuh"
#endif

namespace bs {

// the specialised transforms (fft_fast.cuh), by geometry
#define BS_FAST_GEOMS(X) X(10, 3) X(9, 5) X(10, 5) X(11, 3) X(9, 1)
#ifdef BS_HOSTEMU
static bool fast_analyse_any(const DevGeom &g, const DevTables &T, const float *x, Window w, cf *X, float *sm, bool rotate, float *E) {
  if (!fast_ok(g)) return false;
#define X_(LG, OUTER) if (g.inner == (1 << LG) && g.outer == OUTER) { fast_analyse<LG, OUTER>(g, T, x, w, X, (cf *)sm, rotate, E); return true; }
  BS_FAST_GEOMS(X_)
#undef X_
  return false;
}
static bool fast_synth_any(const DevGeom &g, const DevTables &T, const cf *X, float *frame, float *sm) {
  if (!fast_ok(g)) return false;
#define X_(LG, OUTER) if (g.inner == (1 << LG) && g.outer == OUTER) { fast_synth<LG, OUTER>(g, T, X, frame, (cf *)sm); return true; }
  BS_FAST_GEOMS(X_)
#undef X_
  return false;
}
static size_t fast_smem_floats_max() {
  size_t m = 0;
#define X_(LG, OUTER) m = std::max(m, fast_smem_bytes<LG, OUTER>() / sizeof(float));
  BS_FAST_GEOMS(X_)
#undef X_
  return m;
}
#endif

// ---------------------------------------------------------------------------------- memory / launch abstraction
#ifdef BS_HOSTEMU
typedef void *stream_t;
static bool dev_ok(std::string &) { return true; }
static void *dmalloc(size_t n) { return std::calloc(1, n ? n : 1); }
static void dfree(void *p) { std::free(p); }
static void dzero(void *p, size_t n, stream_t) { std::memset(p, 0, n); }
static void h2d(void *d, const void *h, size_t n, stream_t) { std::memcpy(d, h, n); }
#else
typedef cudaStream_t stream_t;
// A kernel's dynamic shared-memory limit is process-wide state: engines of different geometries live side by side (a batch of
// mixed presets = two engines), so the limit is only ever raised, never set to what the newest engine happens to need.
template <class K>
static cudaError_t raise_smem_limit(K kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<const void *, size_t> limit;
  std::lock_guard<std::mutex> lock(mu);
  size_t &cur = limit[(const void *)kernel];
  if (bytes <= cur) return cudaSuccess;
  const cudaError_t ce = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (ce == cudaSuccess) cur = bytes;
  return ce;
}
static bool dev_ok(std::string &err) {
  int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { err = std::string("no CUDA device: ") + cudaGetErrorString(e); return false; }
  return true;
}
static void *dmalloc(size_t n) {
  void *p = nullptr;
  if (cudaMalloc(&p, n ? n : 1) != cudaSuccess) return nullptr;
  return p;
}
static void dfree(void *p) { if (p) cudaFree(p); }
static void dzero(void *p, size_t n, stream_t s) { cudaMemsetAsync(p, 0, n, s); }
static void h2d(void *d, const void *h, size_t n, stream_t s) { cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s); }

__global__ void __launch_bounds__(256, 3) analysis_kernel(DevGeom g, DevTables T, const StreamDev *streams, const BlockRec *blocks,
                                                       const Window *windows, long long slot0, int nSlots, cf *specIn, float *inEnergy) {
  extern __shared__ __align__(16) float sm[];
  int idx = blockIdx.x;
  const int c = idx % g.C; idx /= g.C;
  const int which = idx & 1; idx >>= 1;
  const int slot = idx % nSlots; const int s = idx / nSlots;
  const StreamDev sd = streams[s];
  const long long m = slot0 + slot;
  if (m >= sd.nBlocks) return;
  if (!(blocks[sd.blockBase + m].flags & kNew)) return;
  const Window w = windows[2 * (sd.blockBase + m) + which];
  cf *X = specIn + ((((size_t)s * nSlots + slot) * 2 + which) * g.C + c) * guard_pitch(g.B) + kGuard;
  float *E = which == 0 ? inEnergy + (((size_t)s * nSlots + slot) * g.C + c) * guard_pitch(g.B) + kGuard : nullptr;   // (map_energy's layout)
  analyse_window(g, T, sd.clip + (size_t)c * sd.clipLen, w, X, sm, threadIdx.x, blockDim.x, which == 1, E);
}

// the same two kernels for the preset geometries (fft_fast.cuh): grid (slot, stream, {cur,prev} x channel) -- no index division
template <int LG, int OUTER>
__global__ void __launch_bounds__(kFastNT, (FastOcc<LG, OUTER>::ctas)) analysis_fast_kernel(DevGeom g, DevTables T, const StreamDev *streams, const BlockRec *blocks,
                                                                  const Window *windows, long long slot0, int nSlots, cf *specIn, float *inEnergy) {
  extern __shared__ __align__(16) float sm[];
  const int slot = blockIdx.x, s = blockIdx.y, which = blockIdx.z & 1, c = blockIdx.z >> 1;
  const StreamDev sd = streams[s];
  const long long m = slot0 + slot;
  if (m >= sd.nBlocks) return;
  if (!(blocks[sd.blockBase + m].flags & kNew)) return;
  const Window w = windows[2 * (sd.blockBase + m) + which];
  cf *X = specIn + ((((size_t)s * nSlots + slot) * 2 + which) * g.C + c) * guard_pitch(g.B) + kGuard;
  float *E = which == 0 ? inEnergy + (((size_t)s * nSlots + slot) * g.C + c) * guard_pitch(g.B) + kGuard : nullptr;
  fast_analyse<LG, OUTER>(g, T, sd.clip + (size_t)c * sd.clipLen, w, X, (cf *)sm, which == 1, E);
}
template <int LG, int OUTER>
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
static bool launch_analysis_fast(const DevGeom &g, const DevTables &T, int S, int nSlots, cudaStream_t q, const StreamDev *streams, const BlockRec *blocks,
                                 const Window *windows, long long slot0, cf *specIn, float *inEnergy) {
  if (!fast_ok(g) || S > 65535) return false;
#define X_(LG, OUTER) if (g.inner == (1 << LG) && g.outer == OUTER) { \
    analysis_fast_kernel<LG, OUTER><<<dim3((unsigned)nSlots, (unsigned)S, (unsigned)(2 * g.C)), kFastNT, fast_smem_bytes<LG, OUTER>(), q>>>(g, T, streams, blocks, windows, slot0, nSlots, specIn, inEnergy); return true; }
  BS_FAST_GEOMS(X_)
#undef X_
  return false;
}
static bool launch_isynth_fast(const DevGeom &g, const DevTables &T, int S, int nSlots, cudaStream_t q, const StreamDev *streams, long long slot0,
                               const cf *specOut, const StateDev &st) {
  if (!fast_ok(g) || S > 65535) return false;
#define X_(LG, OUTER) if (g.inner == (1 << LG) && g.outer == OUTER) { \
    isynth_fast_kernel<LG, OUTER><<<dim3((unsigned)nSlots, (unsigned)S, (unsigned)g.C), kFastNT, fast_smem_bytes<LG, OUTER>(), q>>>(g, T, streams, slot0, nSlots, specOut, st); return true; }
  BS_FAST_GEOMS(X_)
#undef X_
  return false;
}
static bool fast_set_smem() {
  bool ok = true;
#define X_(LG, OUTER) ok = ok && raise_smem_limit(analysis_fast_kernel<LG, OUTER>, (size_t)fast_smem_bytes<LG, OUTER>()) == cudaSuccess && \
                       raise_smem_limit(isynth_fast_kernel<LG, OUTER>, (size_t)fast_smem_bytes<LG, OUTER>()) == cudaSuccess;
  BS_FAST_GEOMS(X_)
#undef X_
  return ok;
}

// ---- map stage kernels (see kernels.cuh "map stage")
struct SlotCtx { StreamDev sd; BlockRec rec; BlockRec2 rec2; size_t slot; bool valid; };
__device__ __forceinline__ SlotCtx slot_ctx(const StreamDev *streams, const BlockRec *blocks, const BlockRec2 *blocks2, long long slot0,
                                            int nSlots, int s, int t) {
  SlotCtx c; c.sd = streams[s]; c.slot = (size_t)s * nSlots + t;
  c.valid = slot0 + t < c.sd.nBlocks;
  if (c.valid) { c.rec = blocks[c.sd.blockBase + slot0 + t]; c.rec2 = blocks2[c.sd.blockBase + slot0 + t]; }
  return c;
}

__global__ void __launch_bounds__(256) map_energy_kernel(DevGeom g, const StreamDev *streams, const BlockRec *blocks, const BlockRec2 *blocks2,
                                                         long long slot0, int nSlots, const cf *specIn, StateDev st) {
  const int t = blockIdx.x % nSlots, s = blockIdx.x / nSlots;
  const SlotCtx c = slot_ctx(streams, blocks, blocks2, slot0, nSlots, s, t);
  if (!c.valid) return;
  const size_t CBg = (size_t)g.C * guard_pitch(g.B);
  const cf *inp = block_input(g, c.rec2, s, slot0, nSlots, specIn, st.lastInput);
  float *inE = st.inEnergy + c.slot * CBg + kGuard, *en = st.energy + c.slot * g.B, *sm = st.smoothed + c.slot * g.B;
  float *fm = st.fm + c.slot * fm_pitch(g.B), *mp = st.map + c.slot * g.B * 2;
  if (g.C == 2) map_energy<2>(g, c.rec, inp, inE, en, sm, fm, mp, threadIdx.x, blockDim.x);
  else if (g.C == 1) map_energy<1>(g, c.rec, inp, inE, en, sm, fm, mp, threadIdx.x, blockDim.x);
  else map_energy<0>(g, c.rec, inp, inE, en, sm, fm, mp, threadIdx.x, blockDim.x);
}

// ---- the one-pole smoothers.  One lane per block (the recurrence s <- s + (v[i]-s)*slew is strictly serial over the
// bins), 32 blocks per warp in lock step.  A lane reading its own row straight from global memory costs the load/store
// unit one line per lane per instruction, and that -- not HBM and not the arithmetic -- was what the kernel waited on.
// So a warp moves a tile (32 rows x 128 bytes) with coalesced 16-byte accesses (four whole lines per instruction) and
// transposes it through an XOR-swizzled shared-memory stage; each lane then picks up its own 32 samples with
// conflict-free LDS.128.  Four sweeps (backward, forward, backward, forward: smoothEnergy steps 1 and 2 / the two
// formant-envelope passes, W#48 8420-8520), the carry running through all of them.  Same operations in the same order
// as smooth_pass_g (kernels.cuh), which stays for odd row lengths and for the serial emulation.
constexpr int kSmoothWarps = 4;
__device__ __forceinline__ void smooth4_warp(float *v /* this lane's row, nullptr = none */, int n, float slew, float4 *stage /* [32][8] */) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, sub = lane >> 3, cc = lane & 7, nt = n >> 5;
  if (!__any_sync(full, v != nullptr)) return;
  float4 *rp[8];                       // the eight rows this lane moves in the coalesced phases: rows i*4 + sub, chunk cc
#pragma unroll
  for (int i = 0; i < 8; ++i) rp[i] = (float4 *)__shfl_sync(full, (unsigned long long)v, i * 4 + sub);
  int co[8], ow[8];                    // stage positions: coalesced phase (row i*4+sub, chunk cc), own row (chunk j)
#pragma unroll
  for (int i = 0; i < 8; ++i) { const int r = i * 4 + sub; co[i] = r * 8 + (cc ^ (r & 7)); ow[i] = lane * 8 + (i ^ (lane & 7)); }
  float s = 0.f;
  float4 nxt[8];
  for (int sweep = 0; sweep < 4; ++sweep) {
    const bool back = !(sweep & 1);
    const int step = back ? -1 : 1;
    int tile = back ? nt - 1 : 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) if (rp[i]) nxt[i] = rp[i][tile * 8 + cc];
    for (int it = 0; it < nt; ++it, tile += step) {
#pragma unroll
      for (int i = 0; i < 8; ++i) stage[co[i]] = nxt[i];
      __syncwarp();
      float x[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float4 q = stage[ow[j]]; x[4 * j] = q.x; x[4 * j + 1] = q.y; x[4 * j + 2] = q.z; x[4 * j + 3] = q.w; }
      if (it + 1 < nt) {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (rp[i]) nxt[i] = rp[i][(tile + step) * 8 + cc];
      }
      if (back) {
#pragma unroll
        for (int j = 31; j >= 0; --j) { s = ((x[j] - s) * slew) + s; x[j] = s; }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) { s = ((x[j] - s) * slew) + s; x[j] = s; }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) stage[ow[j]] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float4 q = stage[co[i]]; if (rp[i]) rp[i][tile * 8 + cc] = q; }
      __syncwarp();
    }
  }
}

// one lane per (stream, block), consecutive blocks of a stream side by side: which = 0 band-energy smoothing (mapped
// blocks), 1 formant-envelope smoothing, 2 formant auto-detect pick (each strictly serial over the bins of a block).
__global__ void __launch_bounds__(32 * kSmoothWarps) map_smooth_kernel(DevGeom g, const StreamDev *streams, const BlockRec *blocks,
                                                                       const BlockRec2 *blocks2, long long slot0, int nSlots, int S, int t0,
                                                                       int nT, StateDev st, int which) {
  __shared__ float4 stageAll[kSmoothWarps][32 * 8];
  const int nTp = (nT + 31) & ~31;                       // whole warps per stream
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = i / nTp, t = t0 + i % nTp;
  SlotCtx c; c.valid = false;
  if (s < S && t < t0 + nT && t < nSlots) c = slot_ctx(streams, blocks, blocks2, slot0, nSlots, s, t);
  float4 *stage = stageAll[threadIdx.x >> 5];
  const bool tiled = (g.B % 32) == 0;                    // rows are whole, aligned 128-byte lines
  if (which == 0) {
    const bool act = c.valid && (c.rec.flags & kMapped);
    const float fN = (float)(uint32_t)g.N, fH = (float)(uint32_t)g.H, ratio = fN / fH;
    const float slew = 1.0f / ((ratio * 0.5f) + 1.0f);
    float *v = act ? st.smoothed + c.slot * g.B : nullptr;
    if (tiled) smooth4_warp(v, g.B, slew, stage);
    else if (act) {
      float carry = smooth_pass_g(v, g.B, slew, 0.f);   // smoothEnergy steps 1,2: the carry runs through both
      smooth_pass_g(v, g.B, slew, carry);
    }
  } else if (which == 1) {
    const bool act = c.valid && (c.rec.flags & kFormants);
    if (tiled) smooth4_warp(act ? st.fm + c.slot * fm_pitch(g.B) : nullptr, g.B, act ? fm_slew(g, c.rec, fm_auto(c.rec) ? st.fmBase[c.slot] : 0.f) : 0.f, stage);
    else if (act) fm_smooth(g, c.rec, fm_auto(c.rec) ? st.fmBase[c.slot] : 0.f, st.fm + c.slot * fm_pitch(g.B));
  } else {
    if (!c.valid || !fm_auto(c.rec)) return;
    fm_auto_pick(g, st.energy + c.slot * g.B, st.fmAuto + 2 * c.slot);
  }
}

__global__ void __launch_bounds__(256) map_peaks_kernel(DevGeom g, const StreamDev *streams, const BlockRec *blocks, const BlockRec2 *blocks2,
                                                        long long slot0, int nSlots, StateDev st) {
  extern __shared__ float4 sm4[];
  const int t = blockIdx.x % nSlots, s = blockIdx.x / nSlots;
  const SlotCtx c = slot_ctx(streams, blocks, blocks2, slot0, nSlots, s, t);
  if (!c.valid || !(c.rec.flags & kMapped)) return;
  map_peaks(g, c.rec, st.energy + c.slot * g.B, st.smoothed + c.slot * g.B, st.map + c.slot * g.B * 2, st.fmAuto + 2 * c.slot, (float *)sm4,
            threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(256) map_fmapply_kernel(DevGeom g, const StreamDev *streams, const BlockRec *blocks, const BlockRec2 *blocks2,
                                                          long long slot0, int nSlots, StateDev st) {
  const int t = blockIdx.x % nSlots, s = blockIdx.x / nSlots;
  const SlotCtx c = slot_ctx(streams, blocks, blocks2, slot0, nSlots, s, t);
  if (!c.valid || !(c.rec.flags & kFormants)) return;
  float *inE = st.inEnergy + c.slot * (size_t)g.C * guard_pitch(g.B) + kGuard; const float *fm = st.fm + c.slot * fm_pitch(g.B);
  if (g.C == 2) fm_apply<2>(g, c.rec, c.rec2, fm, inE, threadIdx.x, blockDim.x);
  else if (g.C == 1) fm_apply<1>(g, c.rec, c.rec2, fm, inE, threadIdx.x, blockDim.x);
  else fm_apply<0>(g, c.rec, c.rec2, fm, inE, threadIdx.x, blockDim.x);
}

// the formant base estimate is a two-tap leaky average over the blocks of a stream: one thread per stream
__global__ void freqest_kernel(int S, const StreamDev *streams, const BlockRec *blocks, long long slot0, int nSlots, StateDev st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const StreamDev sd = streams[s];
  for (int t = 0; t < nSlots && slot0 + t < sd.nBlocks; ++t)
    if (fm_auto(blocks[sd.blockBase + slot0 + t])) {
      const size_t slot = (size_t)s * nSlots + t;
      st.fmBase[slot] = freqest_step(st.freqEst + 2 * s, st.fmAuto + 2 * slot);
    }
}

// preterms: one CTA per (stream, block): the per-bin coefficient records of the phase prediction
__global__ void __launch_bounds__(kTermTile, BS_TERM_CTAS) preterms_kernel(DevGeom g, DevTables T, const StreamDev *streams, const BlockRec *blocks,
                                                       const BlockRec2 *blocks2, long long slot0, int nSlots, const cf *specIn, StateDev st) {
  const int t = blockIdx.x % nSlots, s = blockIdx.x / nSlots;
  const StreamDev sd = streams[s];
  const long long m = slot0 + t;
  if (m >= sd.nBlocks) return;
  const BlockRec rec = blocks[sd.blockBase + m];
  const BlockRec2 rec2 = blocks2[sd.blockBase + m];
  const size_t CB = (size_t)g.C * g.B, CBg = (size_t)g.C * guard_pitch(g.B), slot = (size_t)s * nSlots + t;
  const cf *inp = block_input(g, rec2, s, slot0, nSlots, specIn, st.lastInput);
  const cf *prev = (rec.flags & kNew) ? specIn + (slot * 2 + 1) * CBg + kGuard : nullptr;
  const float *inE = st.inEnergy + slot * CBg + kGuard, *mp = st.map + slot * g.B * 2;
  const bool last = (m + 1 == sd.nBlocks) || (t + 1 == nSlots);
  const uint32_t rng0 = minstd_jump(st.seeds[s], (uint32_t)(((unsigned long long)rec2.rngSkip * (unsigned long long)(2 * g.B - 2)) % 2147483646ull));
  float *rr = st.rec + (size_t)s * ((nSlots + 31) / 32) * rec_group_floats(g.B, g.longStep, g.C) + rec_slot_offset(t, g.B, g.longStep, g.C);
  const float *pE = st.predE[st.parity] + (size_t)s * CB;
  float *pEo = last ? st.predE[st.parity ^ 1] + (size_t)s * CB : nullptr;
  const float *pInE = t > 0 ? inE - CBg : nullptr, *pMap = t > 0 ? mp - (size_t)g.B * 2 : nullptr;
  const int tid = threadIdx.x, nt = blockDim.x;
  extern __shared__ float4 sm4[];
  float *sm = (float *)sm4;
  if (g.C == 2) preterms_block<2>(g, T, rec, rng0, inp, prev, inE, mp, pInE, pMap, pE, pEo, rr, sm, tid, nt);
  else if (g.C == 1) preterms_block<1>(g, T, rec, rng0, inp, prev, inE, mp, pInE, pMap, pE, pEo, rr, sm, tid, nt);
  else preterms_block<0>(g, T, rec, rng0, inp, prev, inE, mp, pInE, pMap, pE, pEo, rr, sm, tid, nt);
}

// carry: the last analysed spectrum of a stream survives the chunk if the next chunk starts with a block that has no
// new spectrum of its own (streaming drive with inputInterval 0).  Runs after the term stage, in the same CUDA stream.
__global__ void __launch_bounds__(256) carry_kernel(DevGeom g, const StreamDev *streams, const BlockRec *blocks, const BlockRec2 *blocks2,
                                                    long long slot0, int nSlots, const cf *specIn, StateDev st) {
  const int s = blockIdx.x;
  const StreamDev sd = streams[s];
  long long nv = sd.nBlocks - slot0; if (nv > nSlots) nv = nSlots;
  if (nv <= 0) return;
  const long long mLast = slot0 + nv - 1;
  if (!(g.incremental || (mLast + 1 < sd.nBlocks && !(blocks[sd.blockBase + mLast + 1].flags & kNew)))) return;
  const BlockRec2 r2 = blocks2[sd.blockBase + mLast];
  if (r2.lastNew < slot0) return;
  const size_t CBg = (size_t)g.C * guard_pitch(g.B);   // (guards and all: they are zero on both sides)
  const cf *src = block_input(g, r2, s, slot0, nSlots, specIn, st.lastInput) - kGuard;
  cf *dst = st.lastInput + (size_t)s * CBg;
  for (int i = threadIdx.x; i < (int)CBg; i += blockDim.x) dst[i] = src[i];
}

// warps per chain CTA: as many as the chunk can use, the kernel was compiled for, and shared memory holds
static int chain_warps(int C, int longStep, int nSlots) {
  int w = std::max(1, std::min(kChainWarpsUsed, (nSlots + 31) / 32));
  while (w > 1 && chain_smem_bytes(C, longStep, w) > (size_t)200 * 1024) --w;
  return w;
}

// one or two channels: a lane per block (chain.cuh); three and more: the channels of a block over lanes (chain_wide.cuh)
template <int C>
static void launch_chain(int S, int warps, size_t smem, cudaStream_t q, const DevGeom &g, const DevTables &T, const StreamDev *streams, const BlockRec *blocks,
                         const BlockRec2 *blocks2, long long slot0, int nSlots, const cf *specIn, cf *specOut, const StateDev &st, int ctas, int *prog, int *err) {
  if constexpr (C <= 2) {
    if (g.incremental) chain_kernel<C, true><<<S * ctas, 32 * warps, smem, q>>>(g, T, streams, blocks, blocks2, slot0, nSlots, specIn, specOut, st, ctas, prog, err);
    else chain_kernel<C, false><<<S * ctas, 32 * warps, smem, q>>>(g, T, streams, blocks, blocks2, slot0, nSlots, specIn, specOut, st, ctas, prog, err);
  }
  else if (wide_split(warps)) chain_wide_kernel<C, true><<<S * ctas, wide_threads(warps), smem, q>>>(g, T, streams, blocks, blocks2, slot0, nSlots, specIn, specOut, st, ctas, prog, err);
  else chain_wide_kernel<C, false><<<S * ctas, wide_threads(warps), smem, q>>>(g, T, streams, blocks, blocks2, slot0, nSlots, specIn, specOut, st, ctas, prog, err);
}
typedef void (*chain_launch_fn)(int, int, size_t, cudaStream_t, const DevGeom &, const DevTables &, const StreamDev *, const BlockRec *, const BlockRec2 *,
                                long long, int, const cf *, cf *, const StateDev &, int, int *, int *);
static const chain_launch_fn kChainLaunch[8] = {launch_chain<1>, launch_chain<2>, launch_chain<3>, launch_chain<4>,
                                                launch_chain<5>, launch_chain<6>, launch_chain<7>, launch_chain<8>};
template <int C> static cudaError_t chain_attr(size_t smem) {
  if constexpr (C <= 2) {
    const cudaError_t ce = raise_smem_limit(chain_kernel<C, false>, (size_t)smem);
    return ce != cudaSuccess ? ce : raise_smem_limit(chain_kernel<C, true>, (size_t)smem);
  }
  else {
    const cudaError_t ce = raise_smem_limit(chain_wide_kernel<C, true>, (size_t)smem);
    return ce != cudaSuccess ? ce : raise_smem_limit(chain_wide_kernel<C, false>, (size_t)smem);
  }
}
template <int C> static int chain_occ(int threads, size_t smem, bool wideSplit) {
  int n = 0;
  cudaError_t ce;
  if constexpr (C <= 2) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, chain_kernel<C, false>, threads, smem);
  else if (wideSplit) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, chain_wide_kernel<C, true>, threads, smem);
  else ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, chain_wide_kernel<C, false>, threads, smem);
  if (ce != cudaSuccess) n = 1;
  return n < 1 ? 1 : n;
}
// blocks of a stream one chain CTA walks in a pass, and the shared memory it needs
// (wideWarps: warps per CTA of the wide kernel, chosen per batch -- wide_warps_for)
static int chain_pass_blocks(int C, int longStep, int nSlots, int wideWarps) { return C <= 2 ? 32 * chain_warps(C, longStep, nSlots) : wide_pass_blocks(C, wideWarps); }
static size_t chain_cta_smem(int C, int longStep, int nSlots, int wideWarps) {
  return C <= 2 ? chain_smem_bytes(C, longStep, chain_warps(C, longStep, nSlots)) : wide_smem_bytes(C, longStep, wideWarps);
}
static int chain_resident_ctas(int C, int threads, size_t smem, bool wideSplit) {   // chain CTAs the whole GPU holds at once
  int dev = 0, sms = 1; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int occ;
  switch (C) {
    case 1: occ = chain_occ<1>(threads, smem, wideSplit); break; case 2: occ = chain_occ<2>(threads, smem, wideSplit); break; case 3: occ = chain_occ<3>(threads, smem, wideSplit); break;
    case 4: occ = chain_occ<4>(threads, smem, wideSplit); break; case 5: occ = chain_occ<5>(threads, smem, wideSplit); break; case 6: occ = chain_occ<6>(threads, smem, wideSplit); break;
    case 7: occ = chain_occ<7>(threads, smem, wideSplit); break; default: occ = chain_occ<8>(threads, smem, wideSplit); break;
  }
  return sms * occ;
}
static cudaError_t chain_set_smem(int C, size_t smem) {
  switch (C) {
    case 1: return chain_attr<1>(smem); case 2: return chain_attr<2>(smem); case 3: return chain_attr<3>(smem); case 4: return chain_attr<4>(smem);
    case 5: return chain_attr<5>(smem); case 6: return chain_attr<6>(smem); case 7: return chain_attr<7>(smem); default: return chain_attr<8>(smem);
  }
}

__global__ void __launch_bounds__(256, 3) isynth_kernel(DevGeom g, DevTables T, const StreamDev *streams, long long slot0, int nSlots,
                                                     const cf *specOut, StateDev st) {
  extern __shared__ __align__(16) float sm[];
  int idx = blockIdx.x;
  const int c = idx % g.C; idx /= g.C;
  const int slot = idx % nSlots; const int s = idx / nSlots;
  const StreamDev sd = streams[s];
  if (slot0 + slot >= sd.nBlocks) return;
  const size_t blk = (size_t)s * nSlots + slot;
  synth_frame(g, T, specOut + (blk * g.C + c) * g.B, st.frames + (blk * g.C + c) * g.L, sm, threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(256) ola_kernel(DevGeom g, DevTables T, const StreamDev *streams, long long slot0, int nSlots, int mode,
                                                  StateDev st, int quad) {
  const int s = blockIdx.y / g.C, c = blockIdx.y % g.C;
  const StreamDev sd = streams[s];
  long long nvl = sd.nBlocks - slot0; if (nvl > nSlots) nvl = nSlots;
  if (nvl <= 0) return;
  const OlaGeom o = ola_geom(g, slot0, (int)nvl, mode);
  const size_t rc = ((size_t)s * g.C + c) * g.L;
  const float *frames = st.frames + (size_t)s * nSlots * g.C * g.L;
  if (quad) {
    const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (x >= o.xE1 + g.L) return;
    ola_quad(g, T, sd, c, x, o, frames, st.ring[st.ringPar] + rc, st.ring[st.ringPar ^ 1] + rc);
  } else {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= o.xE1 + g.L) return;
    ola_sample(g, T, sd, c, x, o, frames, st.ring[st.ringPar] + rc, st.ring[st.ringPar ^ 1] + rc);
  }
}
#endif

struct Stream {
  const float *clip = nullptr; float *out = nullptr; long long clipLen = 0; uint32_t seed = 1;
  StreamPlan plan; bool planned = false;
  int gateIn = 0; long long gateCalls = 0;   // streaming drive: process(gateIn, .) x gateCalls (0 = kiosk drive, no gate)
};

// ---- silence-gate watch for streaming drives.  process() stops running blocks once 2L consecutive input samples were
// silent (sum of squares of a call's input below 1e-15, W#48 7838-7943); the batched path plans every block ahead of the
// data and does not follow the reference through that branch, so it at least says when the branch would have been
// taken: one thread per call sums its input exactly like the reference (channel by channel, sample by sample), one
// thread per stream then runs the counter.  bsb_gate_events() returns the number of calls the reference would have gated.
struct GateDev { const float *clip; long long clipLen, nCalls, callBase; int nIn, pad; };
BS_HD uint8_t gate_call_loud(const GateDev &gd, int C, long long k) {
  float total = 0.f;
  for (int c = 0; c < C; ++c) { const float *x = gd.clip + (size_t)c * gd.clipLen + k * gd.nIn; for (int i = 0; i < gd.nIn; ++i) total = (x[i] * x[i]) + total; }
  return total >= 1e-15f ? 1 : 0;
}
BS_HD int gate_count(const GateDev &gd, int L, const uint8_t *loud) {
  unsigned counter = 0; int fired = 0;
  for (long long k = 0; k < gd.nCalls; ++k) {
    if (loud[gd.callBase + k]) counter = 0;
    else if (counter >= ((unsigned)L << 1)) ++fired;
    else counter += (unsigned)gd.nIn;
  }
  return fired;
}
// A seek the plan assumed loud (control.hpp KioskPlanner): W#49's energy sum over the buffer's clip samples, one accumulator
// through all channels; the zeros around them add nothing.  1 = the assumption failed (the reference's counter kept running).
struct SeekDev { const float *clip; long long clipLen, start; int count, pad; };
BS_HD int seek_watch_failed(const SeekDev &w, int C) {
  float energy = 0.f;
  for (int c = 0; c < C; ++c) { const float *x = w.clip + (size_t)c * w.clipLen + w.start; for (int i = 0; i < w.count; ++i) energy = (x[i] * x[i]) + energy; }
  return energy >= 1e-15f ? 0 : 1;
}
#ifndef BS_HOSTEMU
__global__ void seek_watch_kernel(const SeekDev *ws, int n, int C, int *failed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) failed[i] = seek_watch_failed(ws[i], C);
}
__global__ void gate_energy_kernel(const GateDev *gds, int C, uint8_t *loud) {
  const GateDev gd = gds[blockIdx.y];
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < gd.nCalls) loud[gd.callBase + k] = gate_call_loud(gd, C, k);
}
__global__ void gate_count_kernel(const GateDev *gds, int n, int L, const uint8_t *loud, int *fired) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) fired[s] = gate_count(gds[s], L, loud);
}
#endif

}  // namespace bs

using namespace bs;

struct bsb_engine {
  Geometry g; Tables T; double sampleRate = 48000.0;
  DevGeom dg{}; DevTables dt{};
  std::vector<void *> owned;      // device allocations of the tables
  std::vector<void *> batchOwned; // device allocations of the current batch
  std::vector<Stream> streams;
  std::vector<StreamDev> hs;
  StreamDev *dStreams = nullptr; BlockRec *dBlocks = nullptr; BlockRec2 *dBlocks2 = nullptr; Window *dWindows = nullptr;
  uint32_t *dSeeds = nullptr;
  StateDev st{}; cf *specIn = nullptr, *specOut = nullptr;
  int chunk = 0; long long maxBlocks = 0, totalBlocks = 0, launches = 0;
  // the run as a list of time chunks.  All streams advance together; they are kept longest first, so the streams with
  // blocks left are always a prefix.  Once few enough are left, a chunk covers `ctas` x 256 blocks per stream and the
  // chain wavefront of a stream is relayed across `ctas` CTAs (chain.cuh): the SMs the finished streams freed shorten the
  // remaining streams' critical path instead of idling.
  struct Chunk { long long slot0; int nSlots, nLive, ctas; };
  std::vector<Chunk> chunks, chunksHost;     // device-resident audio / host audio (short first chunk)
  std::vector<long long> needEndHost;
  std::vector<int> order, posOf;   // hs[pos] describes streams[order[pos]]
  int wideWarps = 8;                                   // warps per CTA of chain_wide_kernel for this batch
  int *dChainProg = nullptr, *dChainErr = nullptr;   // relay of the chain wavefront: ticket + progress words; time-out flag
  int *hChainErr = nullptr;                           // pinned copy of the flag, refreshed behind every run
  std::vector<long long> blockBase;
  std::vector<BlockRec> hostBlocks;
  std::vector<long long> needEnd;   // [chunk][stream]: clip samples (per channel) the chunk's analysis windows reach
  int nChunks = 0;
  bool committed = false;
  // per-kernel accounting of the last bsb_run: launches and units always; device time when profiling is on
  // (one CUDA event pair per launch, read back lazily so the run itself is never serialised)
  struct KStat { const char *name; double ms; long long launches, units; };
  std::vector<KStat> kstat;
  struct LaunchRec { int k; double ms; long long units; };
  std::vector<LaunchRec> launchRecs;   // of the last run, in launch order (ms only while profiling)
  bool profiling = false;
  bool fastFft = true;             // specialised STFT kernels where the geometry has them (bsb_set_fast_fft; off = the run-time-geometry path)
  bool overlap = true;             // run the chain/synthesis of chunk i beside the analysis/map/terms of chunk i+1 (two CUDA
                                   // streams).  Gains nothing while every stream is live (each kernel fills the GPU alone), but
                                   // once the shorter streams of a batch have ended the chain runs on a fraction of the SMs and
                                   // the next chunk's front half fills the rest: 334 -> 301 ms on 256 x 60 s of mixed rates
  float *recBuf[2] = {nullptr, nullptr};
  std::vector<GateDev> gate; std::vector<int> gateStream; GateDev *dGate = nullptr; uint8_t *dLoud = nullptr; int *dFired = nullptr;
  std::vector<SeekDev> seekWatch; std::vector<int> seekStream; SeekDev *dSeek = nullptr; int *dSeekFailed = nullptr;
  stream_t lastRunStream = 0;     // bsb_gate_events reads its counters behind whatever the last run queued there
  long long gateCallsTotal = 0, gateMaxCalls = 0;
#ifndef BS_HOSTEMU
  cudaStream_t sFront = nullptr, sBack = nullptr, sIn = nullptr, sOut = nullptr;
  cudaEvent_t evFront[2] = {nullptr, nullptr}, evBack[2] = {nullptr, nullptr}, evFork = nullptr, evJoin[2] = {nullptr, nullptr};
  bool backUsed[2] = {false, false};
  struct Span { int k; cudaEvent_t a, b; size_t rec; };
  std::vector<Span> spans; std::vector<cudaEvent_t> evPool; size_t evUsed = 0;
  cudaEvent_t get_event() {
    if (evUsed == evPool.size()) { cudaEvent_t ev; cudaEventCreate(&ev); evPool.push_back(ev); }
    return evPool[evUsed++];
  }
#endif
  long long maxBlocksOr1(long long slot0) const {   // blocks of the longest stream from slot0 on (at least 1)
    long long n = 1;
    for (const StreamDev &d : hs) n = std::max<long long>(n, d.nBlocks - slot0);
    return n;
  }
  int kidx(const char *name) {
    for (size_t i = 0; i < kstat.size(); ++i) if (!std::strcmp(kstat[i].name, name)) return (int)i;
    kstat.push_back(KStat{name, 0.0, 0, 0}); return (int)kstat.size() - 1;
  }
  std::string err;
  int fail(const char *fmt, ...) {
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    err = buf; return -1;
  }
};

template <class T>
static T *upload(bsb_engine *e, const std::vector<T> &v, std::vector<void *> &owner) {
  T *d = (T *)dmalloc(v.size() * sizeof(T));
  if (!d) return nullptr;
  owner.push_back(d);
  h2d(d, v.data(), v.size() * sizeof(T), 0);
  return d;
}
template <class T>
static T *dalloc(size_t n, std::vector<void *> &owner) {
  T *d = (T *)dmalloc(n * sizeof(T));
  if (d) owner.push_back(d);
  return d;
}

static void free_batch(bsb_engine *e) {
  for (void *p : e->batchOwned) dfree(p);
  e->batchOwned.clear(); e->committed = false;
}

static void reset_state(bsb_engine *e, stream_t q) {
  const Geometry &g = e->g;
  const int S = (int)e->hs.size();
  const size_t CB = (size_t)g.C * g.B;
  StateDev &st = e->st;
  // reset(): zero phase state, rings, maps (the RNG is re-derived from the per-stream seed and the block table)
  dzero(st.outSpec, S * CB * sizeof(cf), q); dzero(st.predE[0], S * CB * sizeof(float), q); dzero(st.predE[1], S * CB * sizeof(float), q);
  dzero(st.lastInput, S * (size_t)g.C * guard_pitch(g.B) * sizeof(cf), q); dzero(st.freqEst, 2 * (size_t)S * sizeof(float), q);
  dzero(st.ring[0], (size_t)S * g.C * g.L * sizeof(float), q); dzero(st.ring[1], (size_t)S * g.C * g.L * sizeof(float), q);
  st.ringPar = 0;
  st.parity = 0;
}

// one time-chunk: stages bit0 = analysis + map + terms + chain, bit1 = synthesis (with `synthMode`)
// Front half (analysis, map, terms, carry) goes to qF, back half (chain, synthesis) to qB; with two different streams
// the back half of chunk i runs beside the front half of chunk i+1, handing the term records over in buffer `buf`.
static int launch_chunk(bsb_engine *e, long long slot0, int nSlots, int nLive, int ctas, stream_t qF, stream_t qB, int buf, int stages, int synthMode) {
  const Geometry &g = e->g;
  const int S = nLive;   // streams are kept longest first: the ones with blocks left at slot0 are the first nLive
#ifdef BS_HOSTEMU
  const size_t CB = (size_t)g.C * g.B, CBg = (size_t)g.C * guard_pitch(g.B);
#endif
  e->st.rec = e->recBuf[buf];
  StateDev &st = e->st;
  const int nt = 256;
  (void)nt; (void)qF; (void)qB;
  // units: analyses (window x channel) actually computed; channel-blocks for the other stages
  long long nNew = 0, nBlk = 0; bool anyAuto = false, anyMapped = false, anyFormants = false;
  for (int s = 0; s < S; ++s) {
    const long long n = std::min<long long>(nSlots, e->hs[s].nBlocks - slot0);
    if (n > 0) nBlk += n;
  }
  if (stages & 1) {
    if (e->dg.incremental) {
      const BlockRec &r = e->hostBlocks[0];
      nNew = (r.flags & kNew) ? 1 : 0; anyAuto = fm_auto(r); anyMapped = r.flags & kMapped; anyFormants = r.flags & kFormants;
    }
    else for (int s = 0; s < S; ++s)
      for (long long m = slot0; m < slot0 + nSlots && m < e->hs[s].nBlocks; ++m) {
        const BlockRec &r = e->hostBlocks[e->hs[s].blockBase + m];
        nNew += (r.flags & kNew) ? 1 : 0; anyAuto = anyAuto || fm_auto(r);
        anyMapped = anyMapped || (r.flags & kMapped); anyFormants = anyFormants || (r.flags & kFormants);
      }
  }
  bool needCarry = e->dg.incremental != 0;
  if ((stages & 1) && !needCarry)
    for (int s = 0; s < S && !needCarry; ++s) {
      const long long mNext = slot0 + nSlots;
      if (mNext < e->hs[s].nBlocks && !(e->hostBlocks[e->hs[s].blockBase + mNext].flags & kNew)) needCarry = true;
    }
  (void)needCarry;
  auto account = [&](const char *name, long long units) {
    const int k = e->kidx(name);
    e->kstat[k].launches += 1; e->kstat[k].units += units; e->launches += 1;
    e->launchRecs.push_back({k, 0.0, units});
    return k;
  };
#ifdef BS_HOSTEMU
  std::vector<f4> smv((map_smem_floats(g.B) + 4 * (size_t)fft_pitch(g.M) + g.L + preterms_smem_floats(g.C, g.longStep) + fast_smem_floats_max() + 64) / 4 + 1);
  float *sm = (float *)smv.data();
  const size_t recPerStream = (size_t)((nSlots + 31) / 32) * rec_group_floats(g.B, g.longStep, g.C);
  auto mapE = g.C == 2 ? map_energy<2> : (g.C == 1 ? map_energy<1> : map_energy<0>);
  auto fmA = g.C == 2 ? fm_apply<2> : (g.C == 1 ? fm_apply<1> : fm_apply<0>);
  auto termFn = g.C == 2 ? preterms_block<2> : (g.C == 1 ? preterms_block<1> : preterms_block<0>);
  if (stages & 1) {
    account("analysis_kernel", nNew * 2 * g.C);
    for (int s = 0; s < S; ++s) {
      const StreamDev &sd = e->hs[s];
      for (int t = 0; t < nSlots && slot0 + t < sd.nBlocks; ++t) {
        long long m = slot0 + t;
        if (!(e->dBlocks[sd.blockBase + m].flags & kNew)) continue;
        for (int which = 0; which < 2; ++which)
          for (int c = 0; c < g.C; ++c) {
            cf *X = e->specIn + ((((size_t)s * nSlots + t) * 2 + which) * g.C + c) * guard_pitch(g.B) + kGuard;
            const float *x = sd.clip + (size_t)c * sd.clipLen; const Window w = e->dWindows[2 * (sd.blockBase + m) + which];
            float *E = which == 0 ? st.inEnergy + (((size_t)s * nSlots + t) * g.C + c) * guard_pitch(g.B) + kGuard : nullptr;
            if (!e->fastFft || !fast_analyse_any(e->dg, e->dt, x, w, X, sm, which == 1, E)) analyse_window(e->dg, e->dt, x, w, X, sm, 0, 1, which == 1, E);
          }
      }
    }
    account("map_energy_kernel", nBlk * g.C);
    if (anyMapped) account("map_smooth_kernel", nBlk);
    if (anyMapped) account("map_peaks_kernel", nBlk);
    if (anyAuto) { account("map_smooth_kernel", 0); account("freqest_kernel", nBlk); }
    if (anyFormants) { account("map_smooth_kernel", 0); account("map_fmapply_kernel", nBlk * g.C); }
    for (int s = 0; s < S; ++s) {   // the map-stage kernels, per stream in block order
      const StreamDev &sd = e->hs[s];
      for (int t = 0; t < nSlots && slot0 + t < sd.nBlocks; ++t) {
        long long m = slot0 + t; const size_t slot = (size_t)s * nSlots + t;
        const BlockRec rec = e->dBlocks[sd.blockBase + m];
        const BlockRec2 rec2 = e->dBlocks2[sd.blockBase + m];
        float *inE = st.inEnergy + slot * CBg + kGuard, *en = st.energy + slot * g.B, *smo = st.smoothed + slot * g.B, *fm = st.fm + slot * fm_pitch(g.B);
        float *mp = st.map + slot * g.B * 2;
        mapE(e->dg, rec, block_input(e->dg, rec2, s, slot0, nSlots, e->specIn, st.lastInput), inE, en, smo, fm, mp, 0, 1);
        if (rec.flags & kMapped) {
          const float fN = (float)(uint32_t)g.N, fH = (float)(uint32_t)g.H, ratio = fN / fH, slew = 1.0f / ((ratio * 0.5f) + 1.0f);
          float carry = smooth_pass_g(smo, g.B, slew, 0.f);
          smooth_pass_g(smo, g.B, slew, carry);
        }
        if (rec.flags & kMapped) map_peaks(e->dg, rec, en, smo, mp, st.fmAuto + 2 * slot, sm, 0, 1);
        float base = 0.f;
        if (fm_auto(rec)) fm_auto_pick(e->dg, en, st.fmAuto + 2 * slot);
        if (fm_auto(rec)) base = st.fmBase[slot] = freqest_step(st.freqEst + 2 * s, st.fmAuto + 2 * slot);
        if (rec.flags & kFormants) { fm_smooth(e->dg, rec, base, fm); fmA(e->dg, rec, rec2, fm, inE, 0, 1); }
      }
    }
    account("preterms_kernel", nBlk * g.C);
    for (int s = 0; s < S; ++s) {
      const StreamDev &sd = e->hs[s];
      for (int t = 0; t < nSlots && slot0 + t < sd.nBlocks; ++t) {
        long long m = slot0 + t; const size_t slot = (size_t)s * nSlots + t;
        const BlockRec rec = e->dBlocks[sd.blockBase + m];
        const BlockRec2 rec2 = e->dBlocks2[sd.blockBase + m];
        const float *inE = st.inEnergy + slot * CBg + kGuard, *mp = st.map + slot * g.B * 2;
        const bool last = (m + 1 == sd.nBlocks) || (t + 1 == nSlots);
        const uint32_t rng0 = minstd_jump(st.seeds[s], (uint32_t)(((unsigned long long)rec2.rngSkip * (unsigned long long)(2 * g.B - 2)) % 2147483646ull));
        termFn(e->dg, e->dt, rec, rng0, block_input(e->dg, rec2, s, slot0, nSlots, e->specIn, st.lastInput),
               (rec.flags & kNew) ? e->specIn + (slot * 2 + 1) * CBg + kGuard : nullptr, inE, mp, t > 0 ? inE - CBg : nullptr,
               t > 0 ? mp - (size_t)g.B * 2 : nullptr, st.predE[st.parity] + (size_t)s * CB,
               last ? st.predE[st.parity ^ 1] + (size_t)s * CB : nullptr,
               st.rec + (size_t)s * recPerStream + rec_slot_offset(t, g.B, g.longStep, g.C), sm, 0, 1);
      }
    }
    account("chain_kernel", nBlk * g.C);
    for (int s = 0; s < S; ++s) {
      const StreamDev &sd = e->hs[s];
      long long nv = std::min<long long>(nSlots, sd.nBlocks - slot0);
      if (nv <= 0) continue;
      const size_t slot = (size_t)s * nSlots;
      const BlockRec *bl = e->dBlocks + sd.blockBase + slot0;
      const BlockRec2 *bl2 = e->dBlocks2 + sd.blockBase + slot0;
      const float *rr = st.rec + (size_t)s * recPerStream;
      cf *so = e->specOut + slot * CB, *state = st.outSpec + (size_t)s * CB;
      switch (g.C) {
        case 1: chain_host<1>(e->dg, e->dt, bl, (int)nv, rr, so, state, bl2); break; case 2: chain_host<2>(e->dg, e->dt, bl, (int)nv, rr, so, state, bl2); break;
        case 3: chain_host<3>(e->dg, e->dt, bl, (int)nv, rr, so, state, bl2); break; case 4: chain_host<4>(e->dg, e->dt, bl, (int)nv, rr, so, state, bl2); break;
        case 5: chain_host<5>(e->dg, e->dt, bl, (int)nv, rr, so, state, bl2); break; case 6: chain_host<6>(e->dg, e->dt, bl, (int)nv, rr, so, state, bl2); break;
        case 7: chain_host<7>(e->dg, e->dt, bl, (int)nv, rr, so, state, bl2); break; default: chain_host<8>(e->dg, e->dt, bl, (int)nv, rr, so, state, bl2); break;
      }
      const long long mLast = slot0 + nv - 1;
      if (e->dg.incremental || (mLast + 1 < sd.nBlocks && !(e->dBlocks[sd.blockBase + mLast + 1].flags & kNew))) {
        const BlockRec2 r2 = e->dBlocks2[sd.blockBase + mLast];
        if (r2.lastNew >= slot0)
          std::memcpy(st.lastInput + (size_t)s * CBg, block_input(e->dg, r2, s, slot0, nSlots, e->specIn, st.lastInput) - kGuard, CBg * sizeof(cf));
      }
    }
    st.parity ^= 1;
  }
  if (stages & 2) {
    if (synthMode & (kSynthAdd | kSynthFrames)) {
      account("isynth_kernel", nBlk * g.C);
      for (int s = 0; s < S; ++s)
        for (int t = 0; t < nSlots && slot0 + t < e->hs[s].nBlocks; ++t)
          for (int c = 0; c < g.C; ++c) {
            const size_t blk = (size_t)s * nSlots + t;
            if (!e->fastFft || !fast_synth_any(e->dg, e->dt, e->specOut + (blk * g.C + c) * g.B, st.frames + (blk * g.C + c) * g.L, sm))
              synth_frame(e->dg, e->dt, e->specOut + (blk * g.C + c) * g.B, st.frames + (blk * g.C + c) * g.L, sm, 0, 1);
          }
    }
    if (!(synthMode & kSynthFrames)) account("ola_kernel", nBlk * g.C);
    for (int s = 0; s < S && !(synthMode & kSynthFrames); ++s) {
      const StreamDev &sd = e->hs[s];
      long long nvl = std::min<long long>(nSlots, sd.nBlocks - slot0);
      if (nvl <= 0) continue;
      const OlaGeom o = ola_geom(e->dg, slot0, (int)nvl, synthMode);
      for (int c = 0; c < g.C; ++c) {
        const size_t rc = ((size_t)s * g.C + c) * g.L;
        if (ola_quad_ok(e->dg))
          for (int x = 0; x < o.xE1 + g.L; x += 4)
            ola_quad(e->dg, e->dt, sd, c, x, o, st.frames + (size_t)s * nSlots * g.C * g.L, st.ring[st.ringPar] + rc, st.ring[st.ringPar ^ 1] + rc);
        else
          for (int x = 0; x < o.xE1 + g.L; ++x)
            ola_sample(e->dg, e->dt, sd, c, x, o, st.frames + (size_t)s * nSlots * g.C * g.L, st.ring[st.ringPar] + rc, st.ring[st.ringPar ^ 1] + rc);
      }
    }
    if (!(synthMode & kSynthFrames)) st.ringPar ^= 1;
  }
#else
  const size_t smA = 4 * (size_t)fft_pitch(g.M) * sizeof(float);
  const int chainWarps = g.C <= 2 ? chain_warps(g.C, g.longStep, ctas > 1 ? 1 << 20 : nSlots) : e->wideWarps;
  const size_t smT = preterms_smem_floats(g.C, g.longStep) * sizeof(float);
  const size_t smM = map_smem_floats(g.B) * sizeof(float), smC = chain_cta_smem(g.C, g.longStep, ctas > 1 ? 1 << 20 : nSlots, e->wideWarps);
  const bool twoStreams = (qF != qB);
  stream_t q = qF;
  bool launchFailed = false;
  auto span = [&](const char *name, long long units, auto &&launch) {
    const int k = account(name, units);
    if (e->profiling) {
      cudaEvent_t a = e->get_event(), b = e->get_event();
      cudaEventRecord(a, q); launch(); cudaEventRecord(b, q);
      e->spans.push_back({k, a, b, e->launchRecs.size() - 1});
    } else launch();
    const cudaError_t ce = cudaGetLastError();   // (a host-side query: says which launch was refused, and why)
    if (ce != cudaSuccess && !launchFailed) { launchFailed = true; e->fail("launch of %s failed (%d streams x %d slots): %s", name, S, nSlots, cudaGetErrorString(ce)); }
  };
  if (stages & 1) {
    const unsigned nCta = (unsigned)((size_t)S * nSlots);
    if (twoStreams && e->backUsed[buf]) cudaStreamWaitEvent(qF, e->evBack[buf], 0);   // the chain that read this record buffer is done
    span("analysis_kernel", nNew * 2 * g.C, [&] {
      if (!e->fastFft || !launch_analysis_fast(e->dg, e->dt, S, nSlots, q, e->dStreams, e->dBlocks, e->dWindows, slot0, e->specIn, st.inEnergy))
        analysis_kernel<<<nCta * 2 * g.C, nt, smA, q>>>(e->dg, e->dt, e->dStreams, e->dBlocks, e->dWindows, slot0, nSlots, e->specIn, st.inEnergy); });
    span("map_energy_kernel", nBlk * g.C, [&] {
      map_energy_kernel<<<nCta, nt, 0, q>>>(e->dg, e->dStreams, e->dBlocks, e->dBlocks2, slot0, nSlots, e->specIn, st); });
    // the smoothers: one launch over the whole chunk, one lane per (stream, block) (slicing the chunk so that the arrays stay
    // L2 resident across the four sweeps was measured slower: too few threads left in flight)
    auto smooth_all = [&](int which) {
      const size_t nThr = (size_t)S * ((nSlots + 31) & ~31);
      map_smooth_kernel<<<(unsigned)((nThr + 32 * kSmoothWarps - 1) / (32 * kSmoothWarps)), 32 * kSmoothWarps, 0, q>>>(e->dg, e->dStreams, e->dBlocks, e->dBlocks2, slot0, nSlots, S, 0, nSlots, st, which);
    };
    if (anyMapped) span("map_smooth_kernel", nBlk, [&] { smooth_all(0); });
    if (anyMapped) span("map_peaks_kernel", nBlk, [&] {
      map_peaks_kernel<<<nCta, 256, smM, q>>>(e->dg, e->dStreams, e->dBlocks, e->dBlocks2, slot0, nSlots, st); });
    if (anyAuto) span("map_smooth_kernel", 0, [&] { smooth_all(2); });
    if (anyAuto) span("freqest_kernel", nBlk, [&] { freqest_kernel<<<(S + 63) / 64, 64, 0, q>>>(S, e->dStreams, e->dBlocks, slot0, nSlots, st); });
    if (anyFormants) {
      span("map_smooth_kernel", 0, [&] { smooth_all(1); });
      span("map_fmapply_kernel", nBlk * g.C, [&] {
        map_fmapply_kernel<<<nCta, nt, 0, q>>>(e->dg, e->dStreams, e->dBlocks, e->dBlocks2, slot0, nSlots, st); });
    }
    span("preterms_kernel", nBlk * g.C, [&] {
      preterms_kernel<<<nCta, kTermTile, smT, q>>>(e->dg, e->dt, e->dStreams, e->dBlocks, e->dBlocks2, slot0, nSlots, e->specIn, st); });
    if (needCarry) span("carry_kernel", 0, [&] {
      carry_kernel<<<S, 256, 0, q>>>(e->dg, e->dStreams, e->dBlocks, e->dBlocks2, slot0, nSlots, e->specIn, st); });
    if (twoStreams) { cudaEventRecord(e->evFront[buf], qF); cudaStreamWaitEvent(qB, e->evFront[buf], 0); }
    q = qB;
    if (ctas > 1) cudaMemsetAsync(e->dChainProg, 0, ((size_t)S * ctas + 1) * sizeof(int), q);   // the ticket counter + progress words
    span("chain_kernel", nBlk * g.C, [&] {
      kChainLaunch[g.C - 1](S, chainWarps, smC, q, e->dg, e->dt, e->dStreams, e->dBlocks, e->dBlocks2, slot0, nSlots, e->specIn, e->specOut, st, ctas, e->dChainProg, e->dChainErr); });
    if (twoStreams) { cudaEventRecord(e->evBack[buf], qB); e->backUsed[buf] = true; }
    st.parity ^= 1;
  }
  q = qB;
  if (stages & 2) {
    if (synthMode & (kSynthAdd | kSynthFrames))
      span("isynth_kernel", nBlk * g.C, [&] {
        if (!e->fastFft || !launch_isynth_fast(e->dg, e->dt, S, nSlots, q, e->dStreams, slot0, e->specOut, st))
          isynth_kernel<<<(unsigned)((size_t)S * nSlots * g.C), nt, smA, q>>>(e->dg, e->dt, e->dStreams, slot0, nSlots, e->specOut, st); });
    const long long span_n = (long long)std::min<long long>(nSlots, e->maxBlocksOr1(slot0)) * g.H + g.L;
    if (!(synthMode & kSynthFrames)) {
      span("ola_kernel", nBlk * g.C, [&] {
        const int quad = ola_quad_ok(e->dg) ? 1 : 0;
        const long long nThr = quad ? (span_n + 3) / 4 : span_n;
        ola_kernel<<<dim3((unsigned)((nThr + 255) / 256), (unsigned)(S * g.C)), 256, 0, q>>>(e->dg, e->dt, e->dStreams, slot0, nSlots, synthMode, st, quad); });
      st.ringPar ^= 1;
    }
  }
  if (launchFailed) return -1;
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return e->fail("kernel launch failed: %s", cudaGetErrorString(ce));
#endif
  return 0;
}

extern "C" {

bsb_engine *bsb_create(int channels, int block, int interval, int split, double sampleRate) {
  std::string err;
  if (channels < 1 || channels > 8 || block < 8 || interval < 1 || interval > block) return nullptr;
  if (!dev_ok(err)) { std::fprintf(stderr, "bauklank_stretch: %s\n", err.c_str()); return nullptr; }
  bsb_engine *e = new bsb_engine();
  e->sampleRate = sampleRate;
  e->g = make_geometry(channels, block, interval, split);
  if (e->g.longStep < 1 || e->g.longStep > 32) { delete e; return nullptr; }
  make_tables(e->g, e->T);
  const Geometry &g = e->g;
  e->dg = DevGeom{g.C, g.L, g.H, g.N, g.B, g.M, g.inner, g.outer, g.split, g.longStep, g.L >> 1, (int)e->T.wpStart.size(), 0, 1u, 0, 0};
  {   // multiply-shift division by `outer`, verified for every index it will see
    bool ok = false;
    for (int sh = 0; sh <= 20 && !ok; ++sh) {
      const unsigned long long mg = (((unsigned long long)1 << sh) + g.outer - 1) / g.outer;
      if (mg * (unsigned long long)(g.M > 0 ? g.M - 1 : 0) >= ((unsigned long long)1 << 32)) break;
      ok = true;
      for (int j = 0; j < g.M && ok; ++j) ok = (int)(((unsigned)j * (unsigned)mg) >> sh) == j / g.outer;
      if (ok) { e->dg.divMagic = (unsigned)mg; e->dg.divShift = sh; }
    }
    if (!ok) { delete e; return nullptr; }
  }
  e->dt.win = upload(e, e->T.win, e->owned); e->dt.tw = upload(e, e->T.tw, e->owned);
  e->dt.otr = upload(e, e->T.otr, e->owned); e->dt.oti = upload(e, e->T.oti, e->owned);
  e->dt.untangle = upload(e, e->T.untangle, e->owned); e->dt.rot = upload(e, e->T.rot, e->owned);
  e->dt.specRot = upload(e, e->T.specRot, e->owned);
  e->dt.wpStart = upload(e, e->T.wpStart, e->owned); e->dt.wpSteady = upload(e, e->T.wpSteady, e->owned);
  e->dt.otw = upload(e, e->T.otw, e->owned);
  e->dg.packTabOk = e->T.packTab.empty() ? 0 : 1;
  e->dt.packTab = e->dg.packTabOk ? upload(e, e->T.packTab, e->owned) : nullptr;
#ifndef BS_HOSTEMU
  cudaStreamSynchronize(0);   // table uploads done before any run can be queued on another stream
  cudaStreamCreateWithFlags(&e->sFront, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&e->sBack, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&e->sIn, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&e->sOut, cudaStreamNonBlocking);
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&e->evFront[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&e->evBack[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->evJoin[i], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&e->evFork, cudaEventDisableTiming);
  if (cudaHostAlloc((void **)&e->hChainErr, sizeof(int), cudaHostAllocDefault) == cudaSuccess) *e->hChainErr = 0; else e->hChainErr = nullptr;
  const size_t smA = 4 * (size_t)fft_pitch(g.M) * sizeof(float);
  if (raise_smem_limit(analysis_kernel, (size_t)smA) != cudaSuccess ||
      raise_smem_limit(isynth_kernel, (size_t)smA) != cudaSuccess || !fast_set_smem() ||
      raise_smem_limit(preterms_kernel, (size_t)(preterms_smem_floats(g.C, g.longStep) * sizeof(float))) != cudaSuccess ||
      chain_set_smem(g.C, chain_cta_smem(g.C, g.longStep, 1 << 20, kChainWarps)) != cudaSuccess ||
      raise_smem_limit(map_peaks_kernel, (size_t)(map_smem_floats(g.B) * sizeof(float))) != cudaSuccess) {
    std::fprintf(stderr, "bauklank_stretch: block %d / interval %d needs more shared memory than one SM has\n", block, interval);
    bsb_destroy(e); return nullptr;
  }
#endif
  return e;
}

bsb_engine *bsb_create_preset(int channels, double sampleRate, int cheaper) {
  // W#58 presetDefault / W#57 presetCheaper: f64 multiply then truncate; the wasm ABI passes sampleRate as f32
  double d = (double)(float)sampleRate;
  if (cheaper) return bsb_create(channels, (int)(d * 0.1), (int)(d * 0.04), 1, sampleRate);
  return bsb_create(channels, (int)(d * 0.12), (int)(d * 0.03), 0, sampleRate);
}

void bsb_destroy(bsb_engine *e) {
  if (!e) return;
  free_batch(e);
  for (void *p : e->owned) dfree(p);
#ifndef BS_HOSTEMU
  for (auto &ev : e->evPool) cudaEventDestroy(ev);
  for (int i = 0; i < 2; ++i) { if (e->evFront[i]) cudaEventDestroy(e->evFront[i]); if (e->evBack[i]) cudaEventDestroy(e->evBack[i]); if (e->evJoin[i]) cudaEventDestroy(e->evJoin[i]); }
  if (e->evFork) cudaEventDestroy(e->evFork);
  if (e->hChainErr) cudaFreeHost(e->hChainErr);
  if (e->sFront) cudaStreamDestroy(e->sFront);
  if (e->sBack) cudaStreamDestroy(e->sBack);
  if (e->sIn) cudaStreamDestroy(e->sIn);
  if (e->sOut) cudaStreamDestroy(e->sOut);
#endif
  delete e;
}
int bsb_block_samples(const bsb_engine *e) { return e->g.L; }
int bsb_interval_samples(const bsb_engine *e) { return e->g.H; }
int bsb_input_latency(const bsb_engine *e) { return e->g.inLat; }
int bsb_output_latency(const bsb_engine *e) { return e->g.outLat; }
int bsb_fft_samples(const bsb_engine *e) { return e->g.N; }
int bsb_bands(const bsb_engine *e) { return e->g.B; }
const char *bsb_last_error(const bsb_engine *e) { return e->err.c_str(); }
long long bsb_total_blocks(const bsb_engine *e) { return e->totalBlocks; }
long long bsb_stream_blocks(const bsb_engine *e, int s) { return (s >= 0 && s < (int)e->streams.size()) ? (long long)e->streams[s].plan.blocks.size() : -1; }
int bsb_chunk_blocks(const bsb_engine *e) { return e->chunk; }
long long bsb_launch_count(const bsb_engine *e) { return e->launches; }
long long bsb_gate_events(bsb_engine *e) {
  if (!e->committed) return -1;
  if (e->gate.empty() && e->seekWatch.empty()) return 0;
  std::vector<int> fired(e->gate.size()), failed(e->seekWatch.size());
#ifdef BS_HOSTEMU
  if (!fired.empty()) std::memcpy(fired.data(), e->dFired, fired.size() * sizeof(int));
  if (!failed.empty()) std::memcpy(failed.data(), e->dSeekFailed, failed.size() * sizeof(int));
#else
  // the counters are written by kernels queued on the stream of the last run (possibly a non-blocking one): read behind them
  if (!fired.empty() && cudaMemcpyAsync(fired.data(), e->dFired, fired.size() * sizeof(int), cudaMemcpyDeviceToHost, e->lastRunStream) != cudaSuccess) return -1;
  if (!failed.empty() && cudaMemcpyAsync(failed.data(), e->dSeekFailed, failed.size() * sizeof(int), cudaMemcpyDeviceToHost, e->lastRunStream) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(e->lastRunStream) != cudaSuccess) return -1;
#endif
  long long n = 0;
  for (int v : fired) n += v;
  for (int v : failed) n += v;
  return n;
}
void bsb_set_profiling(bsb_engine *e, int on) { e->profiling = on != 0; }
int bsb_kernel_count(const bsb_engine *e) { return (int)e->kstat.size(); }
int bsb_kernel_stat(bsb_engine *e, int i, const char **name, double *ms, long long *launches, long long *units) {
  if (i < 0 || i >= (int)e->kstat.size()) return -1;
#ifndef BS_HOSTEMU
  if (!e->spans.empty()) {   // fold the recorded event pairs into the per-kernel totals (blocks until they completed)
    for (auto &sp : e->spans) {
      cudaEventSynchronize(sp.b);
      float ms1 = 0.f; cudaEventElapsedTime(&ms1, sp.a, sp.b); e->kstat[sp.k].ms += ms1;
      if (sp.rec < e->launchRecs.size()) e->launchRecs[sp.rec].ms = ms1;
    }
    e->spans.clear(); e->evUsed = 0;
  }
#endif
  const bsb_engine::KStat &k = e->kstat[i];
  *name = k.name; *ms = k.ms; *launches = k.launches; *units = k.units;
  return 0;
}

int bsb_kernel_launches(bsb_engine *e, int i, double *ms, long long *units, int max) {
  const char *name; double t; long long a, b;
  if (bsb_kernel_stat(e, i, &name, &t, &a, &b) != 0) return -1;   // (folds the pending event pairs)
  int n = 0;
  for (const auto &r : e->launchRecs)
    if (r.k == i) { if (n < max) { if (ms) ms[n] = r.ms; if (units) units[n] = r.units; } ++n; }
  return n;
}

int bsb_begin(bsb_engine *e, int n) {
  if (n < 1) return e->fail("n_streams must be >= 1");
  free_batch(e);
  e->streams.assign(n, Stream());
  return 0;
}

static int check_segments(bsb_engine *e, const bsb_segment *segs, int n, bool allowInactive) {
  if (!segs || n < 1) return e->fail("at least one time-map segment is required");
  for (int i = 0; i < n; ++i) {
    if (!segs[i].active && !allowInactive) return e->fail("inactive time-map segments are not supported by the streaming drive (its silence gate depends on the audio)");
    if (i && segs[i].output < segs[i - 1].output) return e->fail("time-map segments must be ordered by output time");
  }
  return 0;
}
static std::vector<Segment> to_segments(const bsb_segment *s, int n) {
  std::vector<Segment> v(n);
  for (int i = 0; i < n; ++i)
    v[i] = Segment{s[i].output, s[i].input, s[i].rate, s[i].semitones, s[i].tonality_hz, s[i].formant_semitones,
                   s[i].formant_base_hz, s[i].loop_start, s[i].loop_end, s[i].active, s[i].formant_compensation,
                   s[i].transpose_factor, s[i].formant_factor};
  return v;
}

int bsb_add_kiosk(bsb_engine *e, int si, const float *dClip, long long clipLen, float *dOut, long long nOut, int quantum,
                  const bsb_segment *segs, int nSegs, uint32_t seed) {
  if (si < 0 || si >= (int)e->streams.size()) return e->fail("stream index out of range");
  if (quantum < 1 || nOut < 0 || clipLen < 0) return e->fail("bad sizes");
  if (check_segments(e, segs, nSegs, true)) return -1;
  Stream &s = e->streams[si];
  s = Stream(); s.clip = dClip; s.out = dOut; s.clipLen = clipLen; s.seed = seed;
  auto v = to_segments(segs, nSegs);
  plan_kiosk(e->g, e->sampleRate, quantum, nOut, clipLen, v.data(), nSegs, s.plan);
  if (s.plan.error) return e->fail("stream %d: %s", si, s.plan.error);
  s.planned = true; e->committed = false;
  return 0;
}

int bsb_add_kiosk_table(bsb_engine *e, int si, const float *dClip, long long clipLen, float *dOut, long long nOut, int quantum,
                        const bsb_quantum *table, long long nQuanta, uint32_t seed) {
  if (si < 0 || si >= (int)e->streams.size()) return e->fail("stream index out of range");
  if (quantum < 1 || nOut < 0 || clipLen < 0 || !table) return e->fail("bad sizes");
  if (nQuanta * quantum < nOut) return e->fail("the quantum table is shorter than the requested output");
  std::vector<Quantum> qs((size_t)nQuanta);
  for (long long k = 0; k < nQuanta; ++k) {
    const bsb_quantum &t = table[k];
    if (t.valid_start < 0 || t.valid_end > clipLen || t.valid_end < t.valid_start) return e->fail("quantum %lld: valid range outside the clip", k);
    qs[k] = Quantum{t.rate, t.input_samples_end, t.valid_start, t.valid_end, t.semitones, t.tonality_limit, t.formant_semitones,
                    t.formant_base, t.formant_compensation, t.active, t.transpose_factor, t.formant_factor};
  }
  Stream &s = e->streams[si];
  s = Stream(); s.clip = dClip; s.out = dOut; s.clipLen = clipLen; s.seed = seed;
  plan_kiosk_table(e->g, quantum, nOut, qs.data(), nQuanta, s.plan);
  if (s.plan.error) return e->fail("stream %d: %s", si, s.plan.error);
  s.planned = true; e->committed = false;
  return 0;
}

int bsb_add_kiosk_trace(bsb_engine *e, int si, const float *dClip, long long clipLen, float *dOut, long long nOut, int quantum,
                        const bsb_trace_event *events, long long nEvents, uint32_t seed) {
  if (si < 0 || si >= (int)e->streams.size()) return e->fail("stream index out of range");
  if (quantum < 1 || nOut < 0 || clipLen < 0 || nEvents < 0 || (nEvents > 0 && !events)) return e->fail("bad sizes");
  std::vector<TraceEvent> ev((size_t)nEvents);
  for (long long i = 0; i < nEvents; ++i) {
    const bsb_trace_event &t = events[i];
    if (i && t.quantum < events[i - 1].quantum) return e->fail("trace events must be ordered by quantum");
    ev[i] = TraceEvent{t.quantum, t.output_time, t.input, t.rate, t.semitones, t.loop_start, t.loop_end, t.tonality_hz, t.formant_semitones,
                       t.formant_base_hz, t.active, t.formant_compensation, t.transpose_factor, t.formant_factor};
  }
  Stream &s = e->streams[si];
  s = Stream(); s.clip = dClip; s.out = dOut; s.clipLen = clipLen; s.seed = seed;
  plan_kiosk_trace(e->g, e->sampleRate, quantum, nOut, clipLen, ev.data(), nEvents, s.plan);
  if (s.plan.error) return e->fail("stream %d: %s", si, s.plan.error);
  s.planned = true; e->committed = false;
  return 0;
}

int bsb_query_geometry(int block, int interval, int split, int out[6]) {
  if (block < 8 || interval < 1 || interval > block) return -1;
  const Geometry g = make_geometry(1, block, interval, split);
  out[0] = g.N; out[1] = g.B; out[2] = g.inLat; out[3] = g.outLat; out[4] = g.longStep; out[5] = g.inner * 16 + g.outer;
  return 0;
}

int bsb_add_streaming(bsb_engine *e, int si, const float *dClip, long long clipLen, float *dOut, int nIn, int nOut, long long nCalls,
                      const bsb_segment *segs, int nSegs, uint32_t seed) {
  if (si < 0 || si >= (int)e->streams.size()) return e->fail("stream index out of range");
  if (nIn < 1 || nOut < 1 || nCalls < 0 || nCalls * nIn > clipLen) return e->fail("bad sizes (n_calls*n_in must fit the clip)");
  if (check_segments(e, segs, nSegs, false)) return -1;
  Stream &s = e->streams[si];
  s = Stream(); s.clip = dClip; s.out = dOut; s.clipLen = clipLen; s.seed = seed;
  auto v = to_segments(segs, nSegs);
  plan_stream(e->g, e->sampleRate, nIn, nOut, nCalls, clipLen, v.data(), nSegs, s.plan);
  s.gateIn = nIn; s.gateCalls = nCalls;
  s.planned = true; e->committed = false;
  return 0;
}

int bsb_rebind(bsb_engine *e, int si, const float *dClip, float *dOut) {
  if (!e->committed || si < 0 || si >= (int)e->streams.size()) return e->fail("rebind needs a committed batch and a valid stream");
  e->streams[si].clip = dClip; e->streams[si].out = dOut;
  const int pos = e->posOf[si];
  e->hs[pos].clip = dClip; e->hs[pos].out = dOut;
  h2d(e->dStreams + pos, &e->hs[pos], sizeof(StreamDev), 0);
  for (size_t i = 0; i < e->gate.size(); ++i)
    if (e->gateStream[i] == si) { e->gate[i].clip = dClip; h2d(e->dGate + i, &e->gate[i], sizeof(GateDev), 0); }
  for (size_t i = 0; i < e->seekWatch.size(); ++i)
    if (e->seekStream[i] == si) { e->seekWatch[i].clip = dClip; h2d(e->dSeek + i, &e->seekWatch[i], sizeof(SeekDev), 0); }
#ifndef BS_HOSTEMU
  cudaStreamSynchronize(0);   // (see bsb_commit: the next run may be queued on a stream that does not order against this one)
#endif
  return 0;
}

#ifdef BS_HOSTEMU
static int chain_pass_blocks(int C, int, int nSlots, int wideWarps) { return C <= 2 ? 32 * std::max(1, std::min(8, (nSlots + 31) / 32)) : wideWarps * (C <= 4 ? 8 : 4); }   // (planning only)
#endif
static int chain_capacity(bsb_engine *e) {   // chain CTAs resident at once on this GPU
#ifdef BS_HOSTEMU
  (void)e; return 296;
#else
  const int threads = e->g.C <= 2 ? chain_pass_blocks(e->g.C, e->g.longStep, 1 << 20, e->wideWarps) : wide_threads(e->wideWarps);   // (a lane per block / per channel)
  return std::max(1, chain_resident_ctas(e->g.C, threads, chain_cta_smem(e->g.C, e->g.longStep, 1 << 20, e->wideWarps), e->g.C > 2 && wide_split(e->wideWarps)));
#endif
}

int bsb_commit(bsb_engine *e, int chunkBlocks) {
  const Geometry &g = e->g;
  const int S = (int)e->streams.size();
  if (S < 1) return e->fail("no batch: call bsb_begin first");
  free_batch(e);
  std::vector<BlockRec> blocks; std::vector<BlockRec2> blocks2; std::vector<Window> windows; std::vector<uint32_t> seeds(S);
  e->hs.assign(S, StreamDev{}); e->blockBase.assign(S, 0); e->maxBlocks = 0;
  for (int s = 0; s < S; ++s) if (!e->streams[s].planned) return e->fail("stream %d was not added", s);
  // longest stream first (stable): the streams that still have blocks at any point of the run are a prefix
  e->order.resize(S); e->posOf.resize(S);
  for (int s = 0; s < S; ++s) e->order[s] = s;
  std::stable_sort(e->order.begin(), e->order.end(), [&](int a, int b) { return e->streams[a].plan.blocks.size() > e->streams[b].plan.blocks.size(); });
  for (int s = 0; s < S; ++s) {
    Stream &st = e->streams[e->order[s]];
    e->posOf[e->order[s]] = s;
    e->blockBase[s] = (long long)blocks.size();
    blocks.insert(blocks.end(), st.plan.blocks.begin(), st.plan.blocks.end());
    blocks2.insert(blocks2.end(), st.plan.blocks2.begin(), st.plan.blocks2.end());
    windows.insert(windows.end(), st.plan.windows.begin(), st.plan.windows.end());
    StreamDev &d = e->hs[s];
    d.clip = st.clip; d.out = st.out; d.clipLen = st.clipLen; d.nOut = st.plan.nOut; d.blockBase = e->blockBase[s];
    d.nBlocks = (long long)st.plan.blocks.size(); d.outStride = st.plan.nOut; d.outBase = 0;
    d.nLive = st.plan.nLive >= 0 ? std::min(st.plan.nLive, st.plan.nOut) : st.plan.nOut;
    e->maxBlocks = std::max<long long>(e->maxBlocks, d.nBlocks);
    uint32_t sd = st.seed % 2147483647u; seeds[s] = sd <= 1u ? 1u : sd;   // W#26: minstd_rand seeding
  }
  e->totalBlocks = (long long)blocks.size();
  e->hostBlocks = blocks;
  if (blocks.empty()) { blocks.push_back(BlockRec{}); blocks2.push_back(BlockRec2{}); windows.resize(2); }
  const size_t CB = (size_t)g.C * g.B, CBg = (size_t)g.C * guard_pitch(g.B);
  // scratch per chunk slot of one stream: specIn (cur+prev, guarded) + specOut + frames + inEnergy (guarded) + map/energy/smoothed/fm;
  // the term records come in groups of 32 slots per stream (one chain warp), two buffers with the chunk pipelining
  const size_t slotBytes = 2 * CBg * sizeof(cf) + CB * sizeof(cf) + (size_t)g.C * g.L * 4 + CBg * 4 + (size_t)g.B * 16 + (size_t)fm_pitch(g.B) * 4 + 16;
  const size_t groupBytes = (e->overlap ? 2 : 1) * rec_group_floats(g.B, g.longStep, g.C) * 4;
  auto scratch_for = [&](size_t chunk) { return (size_t)S * (chunk * slotBytes + ((chunk + 31) / 32) * groupBytes); };
  const size_t perSlot = (size_t)S * (slotBytes + groupBytes / 32);
  const bool autoChunk = chunkBlocks <= 0;
  // scratch budget: 56 GB of the 180, less on a device that has less to give (other engines of the process, other tenants)
  size_t budget = (size_t)56 << 30;
#ifndef BS_HOSTEMU
  { size_t freeB = 0, totalB = 0; if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) budget = std::min(budget, freeB > ((size_t)6 << 30) ? freeB - ((size_t)6 << 30) : freeB / 2); }
#endif
  if (autoChunk) {   // whole warps of the chain wavefront (32 blocks per stream) while they fit, down to one block per stream
    chunkBlocks = 256;
    while (chunkBlocks > 32 && scratch_for(chunkBlocks) > budget) chunkBlocks -= 32;
    while (chunkBlocks > 1 && scratch_for(chunkBlocks) > budget) chunkBlocks /= 2;
    if (scratch_for(chunkBlocks) > budget)
      return e->fail("not enough device memory for %d streams: %.1f GB of scratch needed for one block per stream, %.1f GB free", S,
                     scratch_for(chunkBlocks) / 1073741824.0, budget / 1073741824.0);
  }
  if (e->maxBlocks > 0 && chunkBlocks > e->maxBlocks) chunkBlocks = (int)e->maxBlocks;
  e->chunk = chunkBlocks;
  // slots allocated in all: S x chunk, or -- a small batch of long streams -- as many as the chain kernel can work on at
  // once when every stream's wavefront is relayed across several CTAs
  e->wideWarps = e->dg.incremental ? 1 : wide_warps_for(S);   // (the shim walks one block at a time)
  const int perPass = chain_pass_blocks(g.C, g.longStep, 1 << 20, e->wideWarps);
#ifndef BS_CHAIN_CTA_CAP
#define BS_CHAIN_CTA_CAP 1024
#endif
  const int cap = std::min(chain_capacity(e), BS_CHAIN_CTA_CAP);   // (more CTAs in flight than that only lengthen the queue of waiting ones)
  size_t allocSlots = (size_t)S * chunkBlocks;
  const bool relayOk = autoChunk && chunkBlocks >= perPass && !e->dg.incremental;
  if (relayOk) {
    const size_t want = (size_t)S * (size_t)((e->maxBlocks + perPass - 1) / perPass) * perPass;
    const size_t room = std::min<size_t>(budget / (perSlot / S), (size_t)cap * perPass);
    allocSlots = std::max(allocSlots, std::min(want, room));
  }
  // `lead` > 0: a short first chunk, for bsb_run_host -- nothing can be computed before the first chunk's clip samples
  // have crossed the bus, so the first chunk is kept small and the rest of the upload hides behind its kernels
  auto build_chunks = [&](int lead, std::vector<bsb_engine::Chunk> &list, std::vector<long long> &need) {
    list.clear();
    for (long long pos = 0; pos < std::max<long long>(e->maxBlocks, 1);) {
      int nLive = 0;
      while (nLive < S && e->hs[nLive].nBlocks > pos) ++nLive;
      if (nLive < 1) nLive = 1;
      bsb_engine::Chunk c{pos, chunkBlocks, nLive, 1};
      if (pos == 0 && lead > 0 && lead < chunkBlocks) c.nSlots = lead;
      else if (relayOk) {
        long long k = std::min<long long>(cap / nLive, (long long)(allocSlots / ((size_t)nLive * perPass)));
        k = std::min<long long>(k, (e->maxBlocks - pos + perPass - 1) / perPass);
        if (k > 1) { c.ctas = (int)k; c.nSlots = (int)k * perPass; }
      }
      list.push_back(c);
      pos += c.nSlots;
    }
    need.assign(std::max<size_t>(1, list.size()) * S, 0);
    for (int s = 0; s < S; ++s) {   // running maximum of the analysis windows' reach, chunk by chunk (for bsb_run_host)
      long long reach = 0;
      for (size_t i = 0; i < list.size(); ++i) {
        const long long m0 = list[i].slot0, m1 = std::min<long long>(m0 + list[i].nSlots, e->hs[s].nBlocks);
        for (long long m = m0; m < m1; ++m)
          for (int w = 0; w < 2; ++w) {
            const Window &x = windows[2 * (e->blockBase[s] + m) + w];
            if (x.hi > x.lo) reach = std::max(reach, x.start + x.hi);
          }
        need[i * S + s] = std::min<long long>(reach, e->hs[s].clipLen);
      }
    }
  };
  build_chunks(0, e->chunks, e->needEnd);
  // (half a chunk, in whole chain warps: measured on 256 x 60 s -- lead 16 / 32 / 64 / 96 / 128 / 192 / none: 232.0 / 231.0 / 229.9 /
  //  228.4 / 226.4 / 228.3 / 229.7 ms per end-to-end step -- a shorter lead waits less for its samples but runs the chain's fill and
  //  drain for little work)
  build_chunks(autoChunk && !e->dg.incremental ? std::max(32, (chunkBlocks / 2) & ~31) : 0, e->chunksHost, e->needEndHost);
  e->nChunks = (int)e->chunks.size();
  auto &own = e->batchOwned;
  e->dStreams = upload(e, e->hs, own); e->dBlocks = upload(e, blocks, own); e->dBlocks2 = upload(e, blocks2, own);
  e->dWindows = upload(e, windows, own); e->dSeeds = upload(e, seeds, own);
  e->specIn = dalloc<cf>(allocSlots * 2 * CBg, own);
  e->specOut = dalloc<cf>(allocSlots * CB, own);
  e->dChainProg = dalloc<int>((size_t)std::max(cap, S) + 64, own);
  e->dChainErr = dalloc<int>(1, own);
  if (e->dChainErr) dzero(e->dChainErr, sizeof(int), 0);
  StateDev &st = e->st;
  const size_t nSlotTot = allocSlots;
  st.outSpec = dalloc<cf>(S * CB, own); st.predE[0] = dalloc<float>(S * CB, own); st.predE[1] = dalloc<float>(S * CB, own);
  st.lastInput = dalloc<cf>(S * CBg, own); st.freqEst = dalloc<float>(2 * (size_t)S, own);
  st.ring[0] = dalloc<float>((size_t)S * g.C * g.L, own); st.ring[1] = dalloc<float>((size_t)S * g.C * g.L, own);
  st.frames = dalloc<float>(nSlotTot * g.C * g.L, own); st.ringPar = 0;
  st.inEnergy = dalloc<float>(nSlotTot * CBg, own); st.map = dalloc<float>(nSlotTot * g.B * 2, own);
  st.fmAuto = dalloc<float>(nSlotTot * 2, own); st.fmBase = dalloc<float>(nSlotTot, own);
  st.energy = dalloc<float>(nSlotTot * g.B, own); st.smoothed = dalloc<float>(nSlotTot * g.B, own); st.fm = dalloc<float>(nSlotTot * fm_pitch(g.B), own);
  size_t recGroups = 1;   // record groups (32 slots of one stream) the fullest chunk of either list uses
  for (const auto *list : {&e->chunks, &e->chunksHost})
    for (const bsb_engine::Chunk &c : *list) recGroups = std::max(recGroups, (size_t)c.nLive * (size_t)((c.nSlots + 31) / 32));
  const size_t recFloats = recGroups * rec_group_floats(g.B, g.longStep, g.C);
  e->recBuf[0] = dalloc<float>(recFloats, own);
  e->recBuf[1] = (e->overlap && e->chunksHost.size() > 1) ? dalloc<float>(recFloats, own) : nullptr;   // only for chunk pipelining
  st.rec = e->recBuf[0];
  st.seeds = e->dSeeds; st.parity = 0;
  if (!e->dStreams || !e->dBlocks || !e->dBlocks2 || !e->dWindows || !e->dSeeds || !e->specIn || !e->specOut || !e->dChainProg || !e->dChainErr || !st.outSpec ||
      !st.predE[0] || !st.predE[1] || !st.lastInput || !st.freqEst || !st.ring[0] || !st.ring[1] || !st.frames || !st.inEnergy || !st.map || !st.fmAuto || !st.fmBase || !st.energy || !st.smoothed || !st.fm ||
      !st.rec) {
    free_batch(e);
    return e->fail("device allocation failed (streams=%d, chunk=%d)", S, chunkBlocks);
  }
  dzero(e->specIn, allocSlots * 2 * CBg * sizeof(cf), 0); dzero(st.inEnergy, nSlotTot * CBg * sizeof(float), 0);   // the guard zeros (kGuard)
  e->gate.clear(); e->gateStream.clear(); e->gateCallsTotal = 0; e->gateMaxCalls = 0;
  for (int s = 0; s < S; ++s) {
    const Stream &x = e->streams[s];
    if (x.gateIn < 1 || x.gateCalls < 1) continue;
    e->gate.push_back(GateDev{x.clip, x.clipLen, x.gateCalls, e->gateCallsTotal, x.gateIn, 0}); e->gateStream.push_back(s);
    e->gateCallsTotal += x.gateCalls; e->gateMaxCalls = std::max(e->gateMaxCalls, x.gateCalls);
  }
  e->seekWatch.clear(); e->seekStream.clear(); e->dSeek = nullptr; e->dSeekFailed = nullptr;
  for (int s = 0; s < S; ++s)
    for (const SeekWatch &w : e->streams[s].plan.watch) {
      e->seekWatch.push_back(SeekDev{e->streams[s].clip, e->streams[s].clipLen, w.start, w.count, 0}); e->seekStream.push_back(s);
    }
  if (!e->seekWatch.empty()) {
    e->dSeek = upload(e, e->seekWatch, own); e->dSeekFailed = dalloc<int>(e->seekWatch.size(), own);
    if (!e->dSeek || !e->dSeekFailed) { free_batch(e); return e->fail("device allocation failed (seek watch)"); }
    dzero(e->dSeekFailed, e->seekWatch.size() * sizeof(int), 0);
  }
  e->dGate = nullptr; e->dLoud = nullptr; e->dFired = nullptr;
  if (!e->gate.empty()) {
    e->dGate = upload(e, e->gate, own); e->dLoud = dalloc<uint8_t>((size_t)e->gateCallsTotal, own); e->dFired = dalloc<int>(e->gate.size(), own);
    if (!e->dGate || !e->dLoud || !e->dFired) { free_batch(e); return e->fail("device allocation failed (gate watch)"); }
    dzero(e->dFired, e->gate.size() * sizeof(int), 0);
  }
#ifndef BS_HOSTEMU
  // the uploads above went through the legacy default stream from pageable memory; the run may be queued on any stream,
  // including non-blocking ones that do not order against it
  if (cudaStreamSynchronize(0) != cudaSuccess) { free_batch(e); return e->fail("device error during the table upload"); }
  if (e->hChainErr) *e->hChainErr = 0;
#endif
  e->committed = true;
  return 0;
}

// Shared by bsb_run (device-resident audio) and bsb_run_host (host audio, copies pipelined chunk by chunk on two
// extra streams: the clip samples chunk i+1 needs go up while chunk i computes, the output samples of chunk i-1 come
// down at the same time).
static int run_impl(bsb_engine *e, stream_t q, const float *const *hClips, float *const *hOuts) {
  if (!e->committed) return e->fail("bsb_run needs a committed batch");
#ifndef BS_HOSTEMU
  if (e->hChainErr && *e->hChainErr) return e->fail("chain relay timed out in an earlier run of this batch (results invalid); re-commit to clear");
#endif
  const Geometry &g = e->g;
  const int S = (int)e->streams.size();
  reset_state(e, q);
  e->launches = 0;
  e->lastRunStream = q;
  for (auto &k : e->kstat) { k.ms = 0.0; k.launches = 0; k.units = 0; }
  e->launchRecs.clear();
  // streams whose silence gate closed for good (inactive time-map segments): the output from there on is zeros
  for (int s = 0; s < S; ++s) {
    const StreamDev &d = e->hs[s];
    if (d.nLive >= d.nOut) continue;
#ifdef BS_HOSTEMU
    for (int c = 0; c < g.C; ++c) std::memset(d.out + (size_t)c * d.outStride + d.nLive, 0, (size_t)(d.nOut - d.nLive) * sizeof(float));
#else
    cudaMemset2DAsync(d.out + d.nLive, (size_t)d.outStride * sizeof(float), 0, (size_t)(d.nOut - d.nLive) * sizeof(float), (size_t)g.C, q);
#endif
  }
#ifdef BS_HOSTEMU
  for (int s = 0; s < S && hClips; ++s)
    std::memcpy((void *)e->streams[s].clip, hClips[s], (size_t)g.C * e->streams[s].clipLen * sizeof(float));
  for (const bsb_engine::Chunk &c : e->chunks)
    if (e->maxBlocks > 0 && launch_chunk(e, c.slot0, c.nSlots, c.nLive, c.ctas, q, q, 0, 3, kSynthEmit | kSynthAdd)) return -1;
  for (int s = 0; s < S && hClips && hOuts; ++s)
    std::memcpy(hOuts[s], e->streams[s].out, (size_t)g.C * e->streams[s].plan.nOut * sizeof(float));
  for (size_t i = 0; i < e->gate.size(); ++i) {
    for (long long k = 0; k < e->gate[i].nCalls; ++k) e->dLoud[e->gate[i].callBase + k] = gate_call_loud(e->gate[i], g.C, k);
    e->dFired[i] = gate_count(e->gate[i], g.L, e->dLoud);
  }
  for (size_t i = 0; i < e->seekWatch.size(); ++i) e->dSeekFailed[i] = seek_watch_failed(e->seekWatch[i], g.C);
#else
  e->spans.clear(); e->evUsed = 0;
  const bool host = hClips != nullptr;
  const std::vector<bsb_engine::Chunk> &chunks = host ? e->chunksHost : e->chunks;
  const std::vector<long long> &needEnd = host ? e->needEndHost : e->needEnd;
  const bool two = e->overlap && e->recBuf[1] != nullptr && chunks.size() > 1;
  if (two || host) {   // fork: the internal streams start after everything already queued on the caller's stream
    cudaEventRecord(e->evFork, q);
    if (two) { cudaStreamWaitEvent(e->sFront, e->evFork, 0); cudaStreamWaitEvent(e->sBack, e->evFork, 0); }
    if (host) { cudaStreamWaitEvent(e->sIn, e->evFork, 0); cudaStreamWaitEvent(e->sOut, e->evFork, 0); }
    e->backUsed[0] = e->backUsed[1] = false;
  }
  std::vector<long long> copied(host ? S : 0, 0), fetched(host ? S : 0, 0);
  // Host audio: streams whose buffers lie back to back with one pitch, on the device and on the host alike (a batch cut out of
  // one big allocation -- thousands of streams then move with one strided copy per time chunk instead of one per stream; the
  // per-call cost of cudaMemcpy2DAsync, not the bus, was what a 4096-stream job waited for).  inRun[s] / outRun[s]: streams
  // in the run that starts at position s (0: s belongs to an earlier run).
  std::vector<int> inRun(host ? S : 0, 1), outRun(host ? S : 0, 1);
  if (host) {
    for (int s = S - 2; s >= 0; --s) {
      const StreamDev &a = e->hs[s], &b = e->hs[s + 1];
      const size_t ci = (size_t)g.C * a.clipLen;
      if (a.clipLen == b.clipLen && a.clipLen > 0 && b.clip == a.clip + ci && hClips[e->order[s + 1]] == hClips[e->order[s]] + ci && inRun[s + 1] < 16384) {
        inRun[s] = inRun[s + 1] + 1; inRun[s + 1] = 0;
      }
      const size_t co = (size_t)g.C * a.nOut;
      if (hOuts && a.nOut == b.nOut && a.nOut > 0 && a.nBlocks == b.nBlocks && a.outStride == a.nOut && b.outStride == b.nOut && b.out == a.out + co &&
          hOuts[e->order[s + 1]] == hOuts[e->order[s]] + co && outRun[s + 1] < 16384) {
        outRun[s] = outRun[s + 1] + 1; outRun[s + 1] = 0;
      }
    }
  }
  long long i = 0;
  for (const bsb_engine::Chunk &ck : chunks) {
    if (e->maxBlocks <= 0) break;
    const long long slot0 = ck.slot0;
    stream_t qF = two ? e->sFront : q, qB = two ? e->sBack : q;
    if (host) {   // clip samples first needed by this chunk, all channels of a run of streams in one strided copy
      for (int s = 0; s < S; ++s) {
        if (!inRun[s]) continue;
        long long need = 0, have = e->hs[s].clipLen;
        for (int r = 0; r < inRun[s]; ++r) { need = std::max(need, needEnd[(size_t)i * S + s + r]); have = std::min(have, copied[s + r]); }
        if (need > have) {
          const StreamDev &d = e->hs[s];
          cudaMemcpy2DAsync((void *)(d.clip + have), (size_t)d.clipLen * sizeof(float), hClips[e->order[s]] + have, (size_t)d.clipLen * sizeof(float),
                            (size_t)(need - have) * sizeof(float), (size_t)g.C * inRun[s], cudaMemcpyHostToDevice, e->sIn);
          for (int r = 0; r < inRun[s]; ++r) copied[s + r] = need;
        }
      }
      cudaEvent_t ev = e->get_event(); cudaEventRecord(ev, e->sIn); cudaStreamWaitEvent(qF, ev, 0);
    }
    if (launch_chunk(e, slot0, ck.nSlots, ck.nLive, ck.ctas, qF, qB, two ? (int)(i & 1) : 0, 3, kSynthEmit | kSynthAdd)) return -1;
    if (host && hOuts) {   // the output samples this chunk emitted
      cudaEvent_t ev = e->get_event(); cudaEventRecord(ev, qB); cudaStreamWaitEvent(e->sOut, ev, 0);
      for (int s = 0; s < S; ++s) {   // (a run: same length, same block count -- the same sample range for all of it)
        if (!outRun[s]) continue;
        const StreamDev &d = e->hs[s];
        const long long n0 = std::min<long long>(slot0 * g.H, d.nOut);
        const long long n1 = std::min<long long>(std::min<long long>(slot0 + ck.nSlots, d.nBlocks) * (long long)g.H, d.nOut);
        if (n1 > n0)
          cudaMemcpy2DAsync(hOuts[e->order[s]] + n0, (size_t)d.nOut * sizeof(float), d.out + n0, (size_t)d.outStride * sizeof(float),
                            (size_t)(n1 - n0) * sizeof(float), (size_t)g.C * outRun[s], cudaMemcpyDeviceToHost, e->sOut);
        for (int r = 0; r < outRun[s]; ++r) fetched[s + r] = std::max(fetched[s + r], n1);
      }
    }
    ++i;
  }
  if (host && hOuts)   // what no block covers: the zeros behind a closed silence gate (written on `cudaStream` before the fork)
    for (int s = 0; s < S; ++s) {
      const StreamDev &d = e->hs[s];
      if (d.nOut > fetched[s])
        cudaMemcpy2DAsync(hOuts[e->order[s]] + fetched[s], (size_t)d.nOut * sizeof(float), d.out + fetched[s], (size_t)d.outStride * sizeof(float),
                          (size_t)(d.nOut - fetched[s]) * sizeof(float), (size_t)g.C, cudaMemcpyDeviceToHost, e->sOut);
    }
  if (two) {   // join
    cudaEventRecord(e->evJoin[0], e->sFront); cudaEventRecord(e->evJoin[1], e->sBack);
    cudaStreamWaitEvent(q, e->evJoin[0], 0); cudaStreamWaitEvent(q, e->evJoin[1], 0);
  }
  if (host) {
    cudaEvent_t a = e->get_event(), b = e->get_event();
    cudaEventRecord(a, e->sIn); cudaEventRecord(b, e->sOut);
    cudaStreamWaitEvent(q, a, 0); cudaStreamWaitEvent(q, b, 0);
  }
  if (!e->gate.empty()) {   // every clip sample is on the device by now
    gate_energy_kernel<<<dim3((unsigned)((e->gateMaxCalls + 127) / 128), (unsigned)e->gate.size()), 128, 0, q>>>(e->dGate, g.C, e->dLoud);
    gate_count_kernel<<<(unsigned)((e->gate.size() + 63) / 64), 64, 0, q>>>(e->dGate, (int)e->gate.size(), g.L, e->dLoud, e->dFired);
  }
  if (e->hChainErr) cudaMemcpyAsync(e->hChainErr, e->dChainErr, sizeof(int), cudaMemcpyDeviceToHost, q);
  if (!e->seekWatch.empty())
    seek_watch_kernel<<<(unsigned)((e->seekWatch.size() + 63) / 64), 64, 0, q>>>(e->dSeek, (int)e->seekWatch.size(), g.C, e->dSeekFailed);
  if (cudaGetLastError() != cudaSuccess) return e->fail("copy or launch failed");
#endif
  return 0;
}

int bsb_run(bsb_engine *e, void *cudaStream) { return run_impl(e, (stream_t)cudaStream, nullptr, nullptr); }

int bsb_run_host(bsb_engine *e, const float *const *hClips, float *const *hOuts, void *cudaStream) {
  if (!hClips) return e->fail("bsb_run_host needs host clip pointers");
  return run_impl(e, (stream_t)cudaStream, hClips, hOuts);
}

void bsb_set_overlap(bsb_engine *e, int on) { e->overlap = on != 0; }
void bsb_set_fast_fft(bsb_engine *e, int on) { e->fastFft = on != 0; }
int bsb_fast_fft_active(const bsb_engine *e) { return (e->fastFft && fast_ok(e->dg) && e->streams.size() <= 65535) ? 1 : 0; }

int bsb_synchronize(bsb_engine *e) {
#ifndef BS_HOSTEMU
  if (cudaStreamSynchronize(e->lastRunStream) != cudaSuccess) return e->fail("device error: %s", cudaGetErrorString(cudaGetLastError()));
  if (e->hChainErr && *e->hChainErr) return e->fail("chain relay timed out: a CTA waited for its predecessor's spectrum for seconds (results of the last run are invalid)");
#endif
  return 0;
}

int bsb_selftest_arith(const float *dX, const float *dD, float *dQ, float *dR, int *dFlags, int n) {
#ifdef BS_HOSTEMU
  for (int i = 0; i < n; ++i) { dQ[i] = div_pos(dX[i], dD[i]); dR[i] = sqrt_z(dX[i]); dFlags[i] = 0; }
  return 0;
#else
  arith_selftest_kernel<<<(n + 255) / 256, 256>>>(dX, dD, dQ, dR, dFlags, n);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
#endif
}

int bsb_block_info(const bsb_engine *e, int s, long long b, long long out[8]) {
  if (s < 0 || s >= (int)e->streams.size()) return -1;
  const StreamPlan &p = e->streams[s].plan;
  if (b < 0 || b >= (long long)p.blocks.size()) return -1;
  out[0] = p.blocks[b].flags; uint32_t u; std::memcpy(&u, &p.blocks[b].timeFactor, 4); out[1] = u;
  for (int w = 0; w < 2; ++w) { const Window &x = p.windows[2 * b + w]; out[2 + 3 * w] = x.start; out[3 + 3 * w] = x.lo; out[4 + 3 * w] = x.hi; }
  return 0;
}

}  // extern "C"

// ---- the reference's STFT output ring (W#22/23, the synthesis steps and the per-sample read of W#48), kept as such
// for the compat shim: ring [C][L] + window products [L] in device memory, the position on the host.  Blocks may start
// at any ring position (after the silence gate re-arms mid-interval they do), which the batched path's index-space
// overlap-add does not model.
BS_HD void shim_ring_add_one(int C /* channels to add */, int L, int i, int pos, float fN, const float *win, const float *frames, float *ring, float *wp) {
  const int p = (pos + i) % L;
  wp[p] = ((win[i] * fN) * win[i]) + wp[p];                               // synthesis step 0: addWindowProduct
  for (int c = 0; c < C; ++c) ring[(size_t)c * L + p] = ring[(size_t)c * L + p] + frames[(size_t)c * L + i];
}
BS_HD void shim_ring_read_one(int C, int L, int i, int pos, int outStride, float *ring, float *wp, float *out, bool consume) {
  const int p = (pos + i) % L;
  const float w = wp[p];
  for (int c = 0; c < C; ++c) {
    if (out) out[(size_t)c * outStride + i] = ring[(size_t)c * L + p] / w;
    if (consume) ring[(size_t)c * L + p] = 0.f;
  }
  if (consume) wp[p] = 1e-30f;                                            // moveOutput
}
#ifndef BS_HOSTEMU
__global__ void shim_ring_add_kernel(int C, int L, int pos, float fN, const float *win, const float *frames, float *ring, float *wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L) shim_ring_add_one(C, L, i, pos, fN, win, frames, ring, wp);
}
__global__ void shim_ring_read_kernel(int C, int L, int n, int pos, int outStride, float *ring, float *wp, float *out, int consume) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) shim_ring_read_one(C, L, i, pos, outStride, ring, wp, out, consume != 0);
}
#endif
static void shim_ring_add(int C, int L, int pos, float fN, const float *win, const float *frames, float *ring, float *wp) {
#ifdef BS_HOSTEMU
  for (int i = 0; i < L; ++i) shim_ring_add_one(C, L, i, pos, fN, win, frames, ring, wp);
#else
  shim_ring_add_kernel<<<(L + 255) / 256, 256>>>(C, L, pos, fN, win, frames, ring, wp);
#endif
}
// out == nullptr: moveOutput(n) only
static void shim_ring_read(int C, int L, int n, int pos, int outStride, float *ring, float *wp, float *out, bool consume) {
  if (n <= 0) return;
#ifdef BS_HOSTEMU
  for (int i = 0; i < n; ++i) shim_ring_read_one(C, L, i, pos, outStride, ring, wp, out, consume);
#else
  shim_ring_read_kernel<<<(n + 255) / 256, 256>>>(C, L, n, pos, outStride, ring, wp, out, consume ? 1 : 0);
#endif
}

// =====================================================================================================================
// Part 1 of the header: the reference's own 18 entry points for one "current" engine instance.
// Host side keeps what the reference keeps in wasm linear memory (I/O buffer, input ring) and runs the same control
// flow as W#48/W#49; every block's DSP runs on the GPU through the kernels above, one block at a time.
namespace {

#ifdef BS_HOSTEMU
void d2h_sync(void *h, const void *d, size_t n) { std::memcpy(h, d, n); }
#else
void d2h_sync(void *h, const void *d, size_t n) {
  if (cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost) != cudaSuccess) bs::die("device error while reading output samples");
}
#endif

#ifdef BS_HOSTEMU
void d2d(void *dst, const void *src, size_t n) { std::memcpy(dst, src, n); }
#else
void d2d(void *dst, const void *src, size_t n) { cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, 0); }
#endif

struct Compat {
  bsb_engine *e = nullptr;
  Geometry g{};
  std::unique_ptr<Control> ctl;
  Params p;
  uint32_t seed = 1, rngState = 1;
  std::vector<float> io; int ioCh = 0, ioLen = 0;
  std::vector<float> ring; int ringPos = 0, inLen = 0;   // stft input ring [C][L+H+1]
  std::vector<float> lastCur, stage;                     // [C][L], [C][2L]
  float *dStage = nullptr, *dOut = nullptr;
  // the STFT output ring and its stashed copy (split computation reads the copy): [0] = live, [1] = stash
  float *dRing[2] = {nullptr, nullptr}, *dWp[2] = {nullptr, nullptr};
  int outPos[2] = {0, 0};
  StreamPlan plan;
  long long m = -1;            // block whose interval is being emitted
  bool deferred = false;       // split mode: block m's spectral work not launched yet
  Window dWin[2]{};
  bool curLaunched = false;
  uint32_t silenceCounter = 0; bool silenceFirst = true;
  int prevCopied = 0;
};
Compat *cc() { static Compat c; return &c; }

void compat_free(Compat *c) {
  if (c->e) { bsb_destroy(c->e); c->e = nullptr; }
  if (c->dStage) dfree(c->dStage);
  if (c->dOut) dfree(c->dOut);
  c->dStage = c->dOut = nullptr;
  for (int i = 0; i < 2; ++i) { if (c->dRing[i]) dfree(c->dRing[i]); if (c->dWp[i]) dfree(c->dWp[i]); c->dRing[i] = c->dWp[i] = nullptr; }
}

// W#22 stft.reset(weight) as far as the output ring goes: zero ring, window products of one frame folded over the
// interval, weighted, then moveOutput(H)
void compat_ring_reset(Compat *c, float weight) {
  const Geometry &g = c->g; const int L = g.L, H = g.H;
  const std::vector<float> &win = c->e->T.win;
  std::vector<float> wp(L, 0.f);
  const float fN = (float)(uint32_t)g.N;
  for (int i = 0; i < L; ++i) wp[i] = ((win[i] * fN) * win[i]) + wp[i];
  for (int i = L - H - 1; i >= 0; --i) wp[i] = wp[i] + wp[i + H];
  for (int i = 0; i < L; ++i) wp[i] = (wp[i] * weight) + 1e-30f;
  for (int i = 0; i < H && i < L; ++i) wp[i] = 1e-30f;   // moveOutput(H) from position 0 (the ring itself is all zero)
  dzero(c->dRing[0], (size_t)g.C * L * sizeof(float), 0);
  h2d(c->dWp[0], wp.data(), (size_t)L * sizeof(float), 0);
#ifndef BS_HOSTEMU
  cudaStreamSynchronize(0);   // wp is a local
#endif
  c->outPos[0] = H % L;
}
void compat_ring_stash(Compat *c) {   // stashedOutput = stft output (ring, window products, position)
  const Geometry &g = c->g;
  d2d(c->dRing[1], c->dRing[0], (size_t)g.C * g.L * sizeof(float)); d2d(c->dWp[1], c->dWp[0], (size_t)g.L * sizeof(float));
  c->outPos[1] = c->outPos[0];
}

// stft.reset + stretch reset (W#22, W#59): host ring, device state, control
void compat_reset_state(Compat *c) {
  const Geometry &g = c->g;
  std::fill(c->ring.begin(), c->ring.end(), 0.f); c->ringPos = g.L;
  std::fill(c->lastCur.begin(), c->lastCur.end(), 0.f);
  reset_state(c->e, 0);
  compat_ring_reset(c, 0.1f); compat_ring_stash(c);
  c->ctl->resetAll();
  c->m = -1; c->deferred = false; c->curLaunched = false;
  c->plan = StreamPlan();
  c->dWin[0] = Window{0, 0, g.L}; c->dWin[1] = Window{g.L, 0, g.L};
  c->silenceCounter = 0;
}

void compat_configure(int ch, int L, int H, int split) {
  Compat *c = cc();
  compat_free(c);
  c->e = bsb_create(ch, L, H, split, 48000.0);
  if (!c->e) bs::die("configure failed: no CUDA device or unsupported block/interval (there is no CPU fallback)");
  bsb_engine *e = c->e;
  c->g = e->g; const Geometry &g = c->g;
  e->dg.incremental = 1;
  c->inLen = L + H + 1;
  c->ring.assign((size_t)ch * c->inLen, 0.f);
  c->lastCur.assign((size_t)ch * L, 0.f); c->stage.assign((size_t)ch * 2 * L, 0.f);
  c->dStage = (float *)dmalloc((size_t)ch * 2 * L * sizeof(float)); c->dOut = (float *)dmalloc((size_t)ch * H * sizeof(float));
  for (int i = 0; i < 2; ++i) { c->dRing[i] = (float *)dmalloc((size_t)ch * L * sizeof(float)); c->dWp[i] = (float *)dmalloc((size_t)L * sizeof(float)); }
  // a one-stream, one-slot batch whose single table entry is rewritten for every block
  e->streams.assign(1, Stream()); e->streams[0].clip = c->dStage; e->streams[0].out = c->dOut; e->streams[0].clipLen = 2 * L;
  e->streams[0].seed = c->seed; e->streams[0].planned = true;
  e->streams[0].plan.blocks.assign(1, BlockRec{}); e->streams[0].plan.blocks2.assign(1, BlockRec2{}); e->streams[0].plan.windows.assign(2, Window{0, 0, 0});
  e->streams[0].plan.nOut = H;
  if (!c->dStage || !c->dOut || !c->dRing[0] || !c->dRing[1] || !c->dWp[0] || !c->dWp[1] || bsb_commit(e, 1) != 0) bs::die("configure: device allocation failed");
  e->hs[0].nBlocks = (long long)1 << 60; e->hs[0].nOut = (long long)1 << 60; e->hs[0].outStride = H;
  c->ctl.reset(new Control(g));
  c->silenceFirst = true;
  compat_reset_state(c);
}

void compat_launch(Compat *c, long long m, const BlockRec &rec, BlockRec2 rec2, const Window *win, int stages, int mode) {
  bsb_engine *e = c->e;
  rec2.lastNew = (rec.flags & kNew) ? (int)m : -1;
  rec2.rngSkip = 0;   // the shim tracks the generator state itself (it outlives configure(), like the reference's)
  e->hs[0].blockBase = -m; e->hs[0].outBase = m * c->g.H;
  h2d(e->dStreams, &e->hs[0], sizeof(StreamDev), 0);
  if (stages & 1) {
    e->hostBlocks.assign(1, rec);
    h2d(e->dBlocks, &rec, sizeof(BlockRec), 0); h2d(e->dBlocks2, &rec2, sizeof(BlockRec2), 0);
    h2d(e->dWindows, win, 2 * sizeof(Window), 0);
    h2d(e->dSeeds, &c->rngState, sizeof(uint32_t), 0);
  }
  if (launch_chunk(e, m, 1, 1, 1, 0, 0, 0, stages, mode)) bs::die(e->err.c_str());
  if ((stages & 1) && !((rec.timeFactor < 0.5f ? 0.5f : rec.timeFactor) <= 2.0f) && c->g.B >= 2)
    c->rngState = minstd_jump(c->rngState, (uint32_t)(2 * c->g.B - 2));
}

// W#24 copyInput(toIndex)
void compat_copy_input(Compat *c, int toIndex) {
  const Geometry &g = c->g;
  int length = toIndex - c->prevCopied, cap = g.L + g.H;
  if (length > cap) length = cap;
  if (length > 0) {
    int offset = toIndex - length;
    for (int ch = 0; ch < g.C; ++ch) {
      const float *src = c->io.data() + (size_t)c->ioLen * ch;
      float *ring = c->ring.data() + (size_t)ch * c->inLen;
      for (int i = 0; i < length; ++i) ring[(c->ringPos + i) % c->inLen] = src[offset + i];
    }
  }
  c->ringPos = (int)(((uint32_t)length + (uint32_t)c->ringPos) % (uint32_t)c->inLen);
  c->prevCopied = toIndex;
}

// the L samples the analysis at `samplesInPast` would read (W#35 3897-3902)
void compat_window(Compat *c, int samplesInPast, float *dst /* [C][2L] stride */, int slot) {
  const Geometry &g = c->g; const int len = c->inLen, L = g.L;
  int start = (int)(((uint32_t)c->ringPos + ((uint32_t)len << 1) - (uint32_t)(samplesInPast + L)) % (uint32_t)len);
  for (int ch = 0; ch < g.C; ++ch) {
    const float *ring = c->ring.data() + (size_t)ch * len; float *d = dst + (size_t)ch * 2 * L + (size_t)slot * L;
    for (int i = 0; i < L; ++i) d[i] = ring[(start + i) % len];
  }
}

// run every step of block m (analysis .. inverse FFT) and add its frames to the live output ring at the ring position
void compat_block(Compat *c, long long m, int nSynth = -1 /* channels whose synthesis step ran; default all */) {
  const Geometry &g = c->g;
  compat_launch(c, m, c->plan.blocks[m], c->plan.blocks2[m], c->dWin, 1 | 2, kSynthFrames);
  if (nSynth != 0)   // synthesis step 0 adds the window products, step c the frame of channel c
    shim_ring_add(nSynth < 0 ? g.C : nSynth, g.L, c->outPos[0], (float)(uint32_t)g.N, c->e->dt.win, c->e->st.frames, c->dRing[0], c->dWp[0]);
}
// The silence gate re-arms blockProcess (W#48 7842-7845) while, with split computation, the block under way has only
// run some of its steps.  What survives of it (every Band is cleared right after): the frames of the channels whose
// synthesis step has run, and the random draws of the vertical-prediction steps that have run.
void compat_abandon_block(Compat *c) {
  if (!c->deferred) return;
  const Geometry &g = c->g;
  int nS6 = 0, nSyn = 0;
  c->ctl->progress(nS6, nSyn);
  if (nSyn > 0) compat_block(c, c->m, nSyn);
  else {
    const float tf = c->plan.blocks[c->m].timeFactor;
    const uint32_t k1 = ((uint32_t)g.B * (uint32_t)nS6) >> 3;
    if (!((tf < 0.5f ? 0.5f : tf) <= 2.0f) && k1 >= 1) c->rngState = minstd_jump(c->rngState, k1 < (uint32_t)g.B ? 2 * k1 - 1 : 2 * k1 - 2);
  }
  c->deferred = false;
}

void compat_process(int nIn, int nOut) {
  Compat *c = cc();
  if (!c->e) bs::die("process() before configure()/presetDefault()/presetCheaper()");
  const Geometry &g = c->g; const int C = g.C, L = g.L, H = g.H;
  if (std::max(nIn, nOut) > c->ioLen) bs::die("process(): sample count exceeds the setBuffers() length");
  if (c->ioCh < g.C) bs::die("process(): setBuffers() was called with fewer channels than the engine is configured for");
  c->prevCopied = 0;
  float total = 0.f; bool loud = false;
  if (C > 0 && nIn > 0) {
    for (int ch = 0; ch < C; ++ch) { const float *x = c->io.data() + (size_t)c->ioLen * ch; for (int i = 0; i < nIn; ++i) total = (x[i] * x[i]) + total; }
    loud = total >= 1e-15f;
  }
  float *outs = c->io.data() + (size_t)c->ioLen * C;
  if (!loud) {   // silence gate, W#48 7838-7943
    if (c->silenceCounter >= ((uint32_t)L << 1)) {
      if (c->silenceFirst) {
        c->silenceFirst = false;
        compat_abandon_block(c);   // blockProcess = {}
        c->ctl->resetBlockProcess();
        const size_t CB = (size_t)C * g.B;
        // every Band cleared: input, prevInput, output (inputEnergy is rewritten by each block)
        dzero(c->e->st.outSpec, CB * sizeof(cf), 0); dzero(c->e->st.lastInput, (size_t)C * guard_pitch(g.B) * sizeof(cf), 0);
        std::fill(c->lastCur.begin(), c->lastCur.end(), 0.f);
      }
      if (nIn > 0) {
        for (int i = 0, j = 0; i < nOut; ++i) { for (int ch = 0; ch < C; ++ch) outs[(size_t)c->ioLen * ch + i] = c->io[(size_t)c->ioLen * ch + j]; j = (j + 1 != nIn) ? j + 1 : 0; }
      } else {
        for (int ch = 0; ch < C; ++ch) for (int i = 0; i < nOut; ++i) outs[(size_t)c->ioLen * ch + i] = 0.f;
      }
      compat_copy_input(c, nIn);
      return;
    }
    c->silenceCounter += (uint32_t)nIn;
  } else { c->silenceFirst = true; c->silenceCounter = 0; }

  c->ctl->p = c->p;
  // c->plan holds the record of the block under way only (plan.blocks[c->m], c->m == 0 once a second block has started):
  // an always-on kiosk must not grow a table
  auto onStart = [&](int, int inputOffset, int, bool rean, bool isNew, long long bi) {
    compat_copy_input(c, inputOffset);
    if (c->deferred) {   // split mode: every step of the previous block has run by now, its records are final
      compat_block(c, c->m);
      c->deferred = false;
    }
    (void)bi;
    c->ctl->keepOnlyCurrent(c->plan);
    c->m = (long long)c->plan.blocks.size() - 1; c->curLaunched = false;
    if (isNew) {
      compat_window(c, 0, c->stage.data(), 0);
      if (rean) compat_window(c, H, c->stage.data(), 1);
      else for (int ch = 0; ch < C; ++ch) std::memcpy(c->stage.data() + (size_t)ch * 2 * L + L, c->lastCur.data() + (size_t)ch * L, L * sizeof(float));
      for (int ch = 0; ch < C; ++ch) std::memcpy(c->lastCur.data() + (size_t)ch * L, c->stage.data() + (size_t)ch * 2 * L, L * sizeof(float));
      h2d(c->dStage, c->stage.data(), c->stage.size() * sizeof(float), 0);
    }
    if (g.split) {   // stash the output ring (this interval is read from the copy), make room for this block's frames
      compat_ring_stash(c);
      shim_ring_read(C, L, H, c->outPos[0], 0, c->dRing[0], c->dWp[0], nullptr, true);
      c->outPos[0] = (c->outPos[0] + H) % L;
      c->deferred = true; c->curLaunched = true;
    }
  };
  std::vector<float> got((size_t)C * H);
  auto onSpan = [&](int idx, int span, uint32_t) {
    if (!c->curLaunched) {   // non-split: every step of the block ran at its first sample, with the current parameters
      compat_block(c, c->m);
      c->curLaunched = true;
    }
    const int r = g.split ? 1 : 0;   // the per-sample read: ring / window products, then moveOutput(1)
    for (int done = 0; done < span;) {   // (a span never exceeds H; the ring read is chunked to the H-sample staging buffer)
      const int n = std::min(span - done, H);
      shim_ring_read(C, L, n, c->outPos[r], H, c->dRing[r], c->dWp[r], c->dOut, true);
      d2h_sync(got.data(), c->dOut, got.size() * sizeof(float));
      for (int ch = 0; ch < C; ++ch) std::memcpy(outs + (size_t)c->ioLen * ch + idx + done, got.data() + (size_t)ch * H, (size_t)n * sizeof(float));
      c->outPos[r] = (c->outPos[r] + n) % L;
      done += n;
    }
  };
  c->ctl->run(c->plan, 0, nOut, nIn, nOut, onStart, onSpan);
  compat_copy_input(c, nIn);
  c->ctl->endCall(nIn);
}

// W#46 flush(outputSamples) -- exported by the reference, never called by its JS.  Control operation on the output ring
// (no DSP): the ring and its window products come down, the arithmetic below is the bytecode's, the state goes back up.
// With split computation the block under way has run only some of its steps when flush() arrives and the reference runs
// the rest afterwards, on the state flush() cleared.  Reproduced: flush before the vertical prediction has started
// (the whole block then runs on the cleared state) and flush after it has finished (synthesised channels are in the old
// ring, the rest add nothing but their window products to the new one).  A flush landing inside the eight
// vertical-prediction steps is handled like one right after them (the reference would still synthesise the bins
// predicted after the flush).
void compat_flush(int nOut) {
  Compat *c = cc();
  if (!c->e) bs::die("flush() before configure()");
  const Geometry &g = c->g; const int C = g.C, L = g.L;
  if (nOut > c->ioLen) bs::die("flush(): sample count exceeds the setBuffers() length");
  if (c->ioCh < g.C) bs::die("flush(): setBuffers() was called with fewer channels than the engine is configured for");
  bool keepPrevInput = false, wpAfter = false;
  if (c->deferred) {
    int nS6 = 0, nSyn = 0;
    c->ctl->progress(nS6, nSyn);
    const int step = c->ctl->stepsDone();
    if (nS6 >= 1 && nS6 <= 7) {
      // inside the eight vertical-prediction steps: the bins predicted so far are cleared with everything else, the rest
      // are predicted after the flush from cleared neighbours and cleared preliminary predictions -- the whole block run
      // on Band.output = 0, with the output of the bins below the step boundary forced to 0
      c->plan.blocks2[c->m].zeroBelow = ((uint32_t)g.B * (uint32_t)nS6) >> 3;
      keepPrevInput = c->ctl->curIsNew();   // its last spectral step sets prevInput = input again
    } else if (step <= c->ctl->stepS6()) {   // runs after the flush, on Band.output = 0 and (unless its copy is still to come) prevInput = 0
      // stft.reset also clears the spectrum scratch: a channel analysed before the flush whose copy into the Bands comes
      // after it arrives as zeros; prevInput copied before the flush is cleared by the flush itself
      const bool isNew = c->ctl->curIsNew(), rean = c->ctl->curReanalysesPrev();
      const int aCur = (isNew && rean) ? C + 1 : 0;
      for (int ch = 0; ch < C; ++ch) {
        const bool prevOk = isNew && rean && step <= ch;
        const bool curLost = isNew && step > aCur + ch && step <= aCur + C;
        float *w = c->stage.data() + (size_t)ch * 2 * L;
        if (!prevOk) std::fill(w + L, w + 2 * L, 0.f);
        if (curLost) { std::fill(w, w + L, 0.f); std::fill(c->lastCur.begin() + (size_t)ch * L, c->lastCur.begin() + (size_t)(ch + 1) * L, 0.f); }
      }
      h2d(c->dStage, c->stage.data(), c->stage.size() * sizeof(float), 0);
#ifndef BS_HOSTEMU
      cudaStreamSynchronize(0);
#endif
      keepPrevInput = c->ctl->curIsNew();   // its last spectral step sets prevInput = input again
    } else {
      compat_block(c, c->m, nSyn);
      c->deferred = false; c->curLaunched = true;
      wpAfter = (nSyn == 0);              // synthesis step 0 is still to come: it adds the window products to the new ring
      keepPrevInput = c->ctl->curIsNew() && step <= c->ctl->stepFinal();
    }
  }
  std::vector<float> ring((size_t)C * L), wp(L);
  d2h_sync(ring.data(), c->dRing[0], ring.size() * sizeof(float)); d2h_sync(wp.data(), c->dWp[0], wp.size() * sizeof(float));
  const int pos = c->outPos[0];
  float m = 0.f;
  for (int i = 0; i < L; ++i) { const int p = (pos + i) % L; const float x = wp[p]; m = (m > x) ? m : x; wp[p] = m; }
  const int n1 = (L < nOut) ? L : nOut, rest = L - n1, n2 = (rest < nOut) ? rest : nOut;
  for (int ch = 0; ch < C; ++ch) {
    const float *r = ring.data() + (size_t)ch * L; float *out = c->io.data() + (size_t)c->ioLen * (C + ch);
    for (int i = 0; i < n1; ++i) { const int p = (pos + i) % L; out[i] = r[p] / wp[p]; }
    for (int i = 0; i < n2; ++i) { const int p = (pos + n1 + i) % L; out[nOut - 1 - i] = out[nOut - 1 - i] - (r[p] / wp[p]); }
  }
  // stft.reset(0.1): input ring, output ring (the stashed copies stay); then Band.prevInput = Band.output = 0
  std::fill(c->ring.begin(), c->ring.end(), 0.f); c->ringPos = L;
  compat_ring_reset(c, 0.1f);
  if (wpAfter) shim_ring_add(0, L, c->outPos[0], (float)(uint32_t)g.N, c->e->dt.win, c->e->st.frames, c->dRing[0], c->dWp[0]);
  dzero(c->e->st.outSpec, (size_t)C * g.B * sizeof(cf), 0);
  if (!keepPrevInput) std::fill(c->lastCur.begin(), c->lastCur.end(), 0.f);
}

void compat_seek(int n, double rate) {   // W#49
  Compat *c = cc();
  if (!c->e) bs::die("seek() before configure()");
  const Geometry &g = c->g; const int cap = g.L + g.H;
  if (n > c->ioLen) bs::die("seek(): sample count exceeds the setBuffers() length");
  if (c->ioCh < g.C) bs::die("seek(): setBuffers() was called with fewer channels than the engine is configured for");
  std::vector<float> tmp(cap, 0.f);
  int start = n - cap; if (start < 0) start = 0;
  float energy = 0.f;
  for (int ch = 0; ch < g.C; ++ch) {
    const float *x = c->io.data() + (size_t)c->ioLen * ch;
    for (int i = start; i < n; ++i) { float v = x[i]; tmp[i + (cap - n)] = v; energy = (v * v) + energy; }
    float *ring = c->ring.data() + (size_t)ch * c->inLen;
    for (int i = 0; i < cap; ++i) ring[(c->ringPos + i) % c->inLen] = tmp[i];
  }
  c->ringPos = (c->ringPos + cap) % c->inLen;
  if (energy >= 1e-15f) { c->silenceCounter = 0; c->silenceFirst = true; }
  c->ctl->seek(rate);
}

}  // namespace

extern "C" {
float *setBuffers(int channels, int length) {
  Compat *c = cc();
  c->io.assign((size_t)2 * channels * length, 0.f); c->ioCh = channels; c->ioLen = length;
  return c->io.data();
}
int blockSamples(void) { return cc()->g.L; }
int intervalSamples(void) { return cc()->g.H; }
int inputLatency(void) { return cc()->g.inLat; }
int outputLatency(void) { return cc()->g.outLat; }
void reset(void) { Compat *c = cc(); if (!c->e) bs::die("reset() before configure()"); compat_reset_state(c); }
void presetDefault(int channels, float sampleRate) { double d = (double)sampleRate; compat_configure(channels, (int)(d * 0.12), (int)(d * 0.03), 0); }
void presetCheaper(int channels, float sampleRate) { double d = (double)sampleRate; compat_configure(channels, (int)(d * 0.1), (int)(d * 0.04), 1); }
void configure(int channels, int block, int interval, int split) { compat_configure(channels, block, interval, split); }
void setTransposeFactor(float m, float tl) { cc()->p.setTransposeFactor(m, tl); }
void setTransposeSemitones(float st, float tl) { cc()->p.setTransposeSemitones(st, tl); }
void setFormantFactor(float m, int comp) { cc()->p.setFormantFactor(m, comp); }
void setFormantSemitones(float st, int comp) { cc()->p.setFormantSemitones(st, comp); }
void setFormantBase(float f) { cc()->p.setFormantBase(f); }
void seek(int n, double rate) { compat_seek(n, rate); }
void process(int nIn, int nOut) { compat_process(nIn, nOut); }
void flush(int n) { compat_flush(n); }
int stretch_main(int, char **) { return 0; }
void stretch_set_seed(uint32_t seed) {
  Compat *c = cc(); c->seed = seed;
  uint32_t sd = seed % 2147483647u;   // W#26: minstd_rand(seed): state = seed % (2^31-1), 0 and 1 map to 1
  c->rngState = sd <= 1u ? 1u : sd;
}
}
