// Device code of the stretch engine: the kernels of one time-chunk of blocks
//
//   analysis_kernel   (stream, block, {cur,prev}, channel)  window -> half-bin-shifted real FFT -> spectrum in HBM (+ the current
//                     window's input energies); analysis_fast_kernel<LG,OUTER> (fft_fast.cuh) for the preset geometries
//   map stage         energy / smooth / peaks (/ freqest / fmsmooth / fmapply) kernels, see "map stage" below
//   preterms_kernel   (stream, block)  per-bin coefficient records of the phase prediction (everything state-free)
//   chain_kernel<1,2> (chain.cuh)  the bin-to-bin / block-to-block phase recurrence as a wavefront over consecutive blocks of a
//                     stream: a lane per block, 8 warps per CTA, relayed from CTA to CTA for long streams;
//   chain_wide_kernel<3..8> (chain_wide.cuh)  the same with the channels of a block spread over lanes
//   isynth_kernel     (stream, block, channel)  inverse FFT -> synthesis window -> frame in HBM; isynth_fast_kernel<LG,OUTER>
//   ola_kernel        (stream, channel, output sample)  overlap-add of the frames in block order -> normalised output
//
// All arithmetic is f32 in the reference's operation order (compile with -fmad=false; IEEE div/sqrt), see the
// citations on each routine (W#n = wasm function n of the blob at app/SignalsmithStretch.mjs:265).
//
// Every routine is written as a (tid, nthreads) work loop so that the identical source can be executed serially on
// the host by the test-only emulation build (tests/hostemu, -DBS_HOSTEMU); the product build is CUDA only.
#pragma once
#include <cstdint>
#include <cmath>

#include "control.hpp"

#ifdef BS_HOSTEMU
#include <cstring>
#define BS_HD inline
#define BS_HHD inline
#define BS_SYNC() ((void)0)
#define BS_WARPSYNC() ((void)0)
inline int __float_as_int_hd(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline float __int_as_float_hd(int i) { float f; std::memcpy(&f, &i, 4); return f; }
#else
#define BS_HD __device__ __forceinline__
#define BS_HHD __host__ __device__ __forceinline__
#define BS_SYNC() __syncthreads()
#define BS_WARPSYNC() __syncwarp()
#define __float_as_int_hd(x) __float_as_int(x)
#define __int_as_float_hd(x) __int_as_float(x)
#endif

namespace bs {

struct DevGeom {
  int C, L, H, N, B, M, inner, outer, split, longStep, off;  // off = L>>1 (analysis/synthesis offset)
  int wpStartLen;
  int incremental;  // blocks arrive one at a time (compat shim): always carry the input spectrum forward
  unsigned divMagic; int divShift;   // j / outer == (j * divMagic) >> divShift for every j < M (checked at create time)
  int packTabOk;                     // DevTables::packTab exists (window halves on pair boundaries)
};
enum : int { kSynthEmit = 1, kSynthAdd = 2, kSynthFrames = 4 };   // kSynthFrames: windowed frames only, no overlap-add (the compat shim keeps the reference's own output ring)
struct DevTables {
  const float *win; const cf *tw; const float *otr, *oti; const cf *untangle, *rot, *specRot;
  const float *wpStart, *wpSteady;
  const float *packTab; const cf *otw;   // fft_fast.cuh
};
// per-stream device record
struct StreamDev {
  const float *clip;      // planar [C][clipLen]
  float *out;             // planar [C][nOut]
  long long clipLen, nOut;
  long long blockBase;    // index of this stream's first block in the global block/window arrays
  long long outStride;    // channel stride of `out`
  long long outBase;      // output sample n is stored at out[c*outStride + n - outBase]
  long long nBlocks;
  long long nLive;        // output samples [0, nLive) come from the blocks; [nLive, nOut) is where the silence gate had closed (zeros)
};
// persistent per-stream state + per-chunk scratch (all device pointers, stream-major)
struct StateDev {
  cf *outSpec;        // [S][C][B]     Band.output carried between wavefront passes and chunks
  float *predE[2];    // [S][C][B] x2  Prediction.energy of a chunk's last block (read [parity], write [parity^1])
  cf *lastInput;      // [S][C][B]     last analysed spectrum (only used by blocks without a new spectrum)
  float *freqEst;     // [S][2]        freqEstimateWeighted, freqEstimateWeight
  float *ring[2];     // [S][C][L] x2  overlap-add partial sums carried between chunks (read [ringPar], write [ringPar^1])
  float *frames;      // [S][T][C][L]  windowed synthesis frames of the chunk
  int ringPar;
  // per chunk slot
  float *inEnergy;    // [S][T][C][B]
  float *map;         // [S][T][B][2]  {inputBin, freqGrad}
  float *energy, *smoothed;  // [S][T][B]  channel-summed band energy and its smoothed copy
  float *fm;          // [S][T][fm_pitch(B)]  formant metric (sqrt of the band energy), smoothed in place
  float *fmAuto;      // [S][T][2]     formant auto-detect: spectral peak (top, index)
  float *fmBase;      // [S][T]        formant base bin after the leaky averages
  float *rec;         // [S][T][B+longStep+1][NR]  term records (one row per wavefront step)
  const uint32_t *seeds;  // [S]       minstd_rand state at the start of the stream
  int parity;
};

BS_HD int trunc_i32(float x) { return fabsf(x) < 2147483648.0f ? (int)x : INT32_MIN; }
// Arrays the term stage interpolates in at data-dependent positions (analysis spectra, the carried input spectrum, input
// energies) carry two zero elements on either side of every channel: a position outside [0, B) reads as 0 (W#48 9314-9455)
// without a bounds check -- the index is clamped to [-2, B] and both neighbours then fall on guard zeros.  Channel pitch
// B + 4; pointers to such arrays point at element 0.  The guards are zeroed once when the arrays are allocated and never written.
constexpr int kGuard = 2;
BS_HHD int guard_pitch(int B) { return B + 2 * kGuard; }
struct alignas(16) f4 { float x, y, z, w; };
struct alignas(8) f2 { float x, y; };

// ------------------------------------------------------------------------------------------------------------
// One radix-4 butterfly of the reference's decimation-in-time pass (W#34 forward / W#33 inverse), in registers, in
// place: (A, B, C, D) <- outputs 0..3.  Operation order is the reference's; twiddles are applied even when trivial.
template <bool INV>
BS_HD void bfly4(float &Ar, float &Ai, float &Br, float &Bi, float &Cr, float &Ci, float &Dr, float &Di, const cf tB, const cf tC, const cf tD) {
  float dRe, bRe, cRe, dIm, bIm, cIm;
  if (!INV) {
    dRe = (Dr * tD.re) - (Di * tD.im); bRe = (Br * tB.re) - (Bi * tB.im); cRe = (Cr * tC.re) - (Ci * tC.im);
    dIm = (Di * tD.re) + (Dr * tD.im); bIm = (Bi * tB.re) + (Br * tB.im); cIm = (Ci * tC.re) + (Cr * tC.im);
  } else {
    dRe = (Di * tD.im) + (Dr * tD.re); bRe = (Bi * tB.im) + (Br * tB.re); cRe = (Ci * tC.im) + (Cr * tC.re);
    dIm = (Di * tD.re) - (Dr * tD.im); bIm = (Bi * tB.re) - (Br * tB.im); cIm = (Ci * tC.re) - (Cr * tC.im);
  }
  const float bdRe = dRe + bRe, acRe = cRe + Ar, bdIm = dIm + bIm, acIm = Ai + cIm;
  const float x = INV ? (dIm - bIm) : (bIm - dIm), y = Ar - cRe;
  const float z = INV ? (bRe - dRe) : (dRe - bRe), w = Ai - cIm;
  Ar = bdRe + acRe; Ai = bdIm + acIm;
  Br = x + y;       Bi = z + w;
  Cr = acRe - bdRe; Ci = acIm - bdIm;
  Dr = y - x;       Di = w - z;
}

// The reference's radix-4 DIT passes on split arrays (W#21/W#34 forward, W#20/W#33 inverse), `outer` sub-transforms of
// length `inner` (a power of two) side by side.  Consecutive passes are fused two at a time: a thread takes the 16
// inputs of four first-level butterflies, keeps their outputs in registers and feeds them straight into the four
// second-level butterflies they belong to -- the same butterflies on the same values as two separate passes, one
// shared-memory round trip and one barrier fewer.  All index arithmetic is shifts and masks; for the common
// geometries (LG = log2(inner), OUTER, NT = threads compile-time constants) every stride is an immediate and the pass
// loop unrolls completely.  LG == 0: run-time geometry.
// Data starts in (ar, ai); returns 0 if the result is in (ar, ai), 1 if in (br, bi).
template <bool INV, int LG, int OUTER, int NT>
BS_HD int pow2_ffts_t(const DevGeom &g, const cf *tw, float *ar, float *ai, float *br, float *bi, int tid, int ntRun) {
  const int inner = LG ? (1 << LG) : g.inner, outer = LG ? OUTER : g.outer, nt = NT ? NT : ntRun;
  if (inner <= 1) return 0;
  int lgRun = 0;
  if (!LG) while ((1 << lgRun) < inner) ++lgRun;
  const int lg = LG ? LG : lgRun;
  float *sr = ar, *si = ai, *dr = br, *di = bi;
  int which = 0, lgSize = 0;
  if (lg & 1) {
    const int stride = inner >> 1, total = outer * stride;
    for (int idx = tid; idx < total; idx += nt) {
      int sub = idx >> (lg - 1), s = idx & (stride - 1), p0 = sub * inner + s, p1 = p0 + stride;
      float a_i = si[p0], b_i = si[p1], b_r = sr[p1], a_r = sr[p0];
      dr[p0] = b_r + a_r; di[p0] = b_i + a_i; dr[p1] = a_r - b_r; di[p1] = a_i - b_i;
    }
    BS_SYNC();
    float *t; t = sr; sr = dr; dr = t; t = si; si = di; di = t; which ^= 1; lgSize = 1;
  }
#pragma unroll
  for (int stage = 0; stage < 8; ++stage) {
    if (lgSize >= lg) break;
    if (lg - lgSize >= 4) {
      // fused passes A (size 4*s0) and B (size 16*s0)
      const int lgQA = lgSize, lgStrideA = lg - lgSize - 2, lgStrideB = lgStrideA - 2, lgQB = lgSize + 2;
      const int strideB = 1 << lgStrideB;
      const int lgPer = lg - 4, total = outer << lgPer;            // items per sub-transform: qA * strideB = inner/16
      for (int idx = tid; idx < total; idx += nt) {
        const int sub = idx >> lgPer, r = idx & ((1 << lgPer) - 1), iA = r >> lgStrideB, sB = r & (strideB - 1);
        // padIn: this pass reads with a 4-element inner stride -- eight of a warp's lanes per bank unless every 64-float
        // group of the array is shifted by 4 floats; the pass before (padOut) wrote it that way.  Offsets within a group
        // and whole-group offsets are compile-time constants either way.
        constexpr bool kSpec = LG != 0;
        const bool padIn = kSpec && lgStrideB == 2, padOut = kSpec && (lg - lgSize == 10);
        int base = sub * inner + (iA << (lgStrideA + 2)) + sB;
        if (padIn) base += (base >> 6) << 2;
        float vr[4][4], vi[4][4];   // [a][j]
        if (LG && lgStrideB == 0) {   // the four `a` inputs of a `j` are contiguous: one 16-byte load each (and no bank conflicts)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const f4 qr = *(const f4 *)(sr + base + (j << lgStrideA)), qi = *(const f4 *)(si + base + (j << lgStrideA));
            vr[0][j] = qr.x; vr[1][j] = qr.y; vr[2][j] = qr.z; vr[3][j] = qr.w;
            vi[0][j] = qi.x; vi[1][j] = qi.y; vi[2][j] = qi.z; vi[3][j] = qi.w;
          }
        } else {
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j) { const int p = base + (j << lgStrideA) + (a << lgStrideB); vr[a][j] = sr[p]; vi[a][j] = si[p]; }
        }
        {
          const cf tB = tw[iA << lgStrideA], tC = tw[(2 * iA) << lgStrideA], tD = tw[(3 * iA) << lgStrideA];
#pragma unroll
          for (int a = 0; a < 4; ++a) bfly4<INV>(vr[a][0], vi[a][0], vr[a][1], vi[a][1], vr[a][2], vi[a][2], vr[a][3], vi[a][3], tB, tC, tD);
        }
        int obase = sub * inner + (iA << lgStrideB) + sB;
        if (padOut) obase += (obase >> 6) << 2;      // (lgStrideB == 6 here: sB < 64, the offsets below are whole groups)
#pragma unroll
        for (int jA = 0; jA < 4; ++jA) {
          const int iB = iA + (jA << lgQA);
          const cf tB = tw[iB << lgStrideB], tC = tw[(2 * iB) << lgStrideB], tD = tw[(3 * iB) << lgStrideB];
          bfly4<INV>(vr[0][jA], vi[0][jA], vr[1][jA], vi[1][jA], vr[2][jA], vi[2][jA], vr[3][jA], vi[3][jA], tB, tC, tD);
#pragma unroll
          for (int jB = 0; jB < 4; ++jB) {
            const int o = (jA << (lgQA + lgStrideB)) + (jB << (lgQB + lgStrideB));
            const int p = obase + o + (padOut ? (o >> 6) << 2 : 0);
            dr[p] = vr[jB][jA]; di[p] = vi[jB][jA];
          }
        }
      }
      lgSize += 4;
    } else {
      // single pass (size 4*s0)
      const int lgStride = lg - lgSize - 2, stride = 1 << lgStride, lgQ = lgSize, lgPer = lg - 2, total = outer << lgPer;
      for (int idx = tid; idx < total; idx += nt) {
        const int sub = idx >> lgPer, r = idx & ((1 << lgPer) - 1), i = r >> lgStride, s = r & (stride - 1), base = sub * inner;
        const cf tB = tw[i << lgStride], tC = tw[(2 * i) << lgStride], tD = tw[(3 * i) << lgStride];
        const int pa = base + ((4 * i) << lgStride) + s;
        float Ar, Ai, Br, Bi, Cr, Ci, Dr, Di;
        if (LG && lgStride == 0) {   // the last pass: four contiguous inputs, one 16-byte load per array
          const f4 qr = *(const f4 *)(sr + pa), qi = *(const f4 *)(si + pa);
          Ar = qr.x; Br = qr.y; Cr = qr.z; Dr = qr.w; Ai = qi.x; Bi = qi.y; Ci = qi.z; Di = qi.w;
        } else {
          Ar = sr[pa]; Ai = si[pa]; Br = sr[pa + stride]; Bi = si[pa + stride];
          Cr = sr[pa + 2 * stride]; Ci = si[pa + 2 * stride]; Dr = sr[pa + 3 * stride]; Di = si[pa + 3 * stride];
        }
        bfly4<INV>(Ar, Ai, Br, Bi, Cr, Ci, Dr, Di, tB, tC, tD);
        const int po = base + (i << lgStride) + s, qs = 1 << (lgQ + lgStride);
        dr[po] = Ar;          di[po] = Ai;
        dr[po + qs] = Br;     di[po + qs] = Bi;
        dr[po + 2 * qs] = Cr; di[po + 2 * qs] = Ci;
        dr[po + 3 * qs] = Dr; di[po + 3 * qs] = Di;
      }
      lgSize += 2;
    }
    BS_SYNC();
    float *t; t = sr; sr = dr; dr = t; t = si; si = di; di = t; which ^= 1;
  }
  return which;
}
// dispatch on the geometry: the presets and the kiosk's shipped configuration get fully specialised code
template <bool INV>
BS_HD int pow2_ffts(const DevGeom &g, const cf *tw, float *ar, float *ai, float *br, float *bi, int tid, int nt) {
#ifndef BS_HOSTEMU
  if (nt == 256) {
    if (g.inner == 1024 && g.outer == 3) return pow2_ffts_t<INV, 10, 3, 256>(g, tw, ar, ai, br, bi, tid, nt);   // 48 kHz presetDefault
    if (g.inner == 512 && g.outer == 5) return pow2_ffts_t<INV, 9, 5, 256>(g, tw, ar, ai, br, bi, tid, nt);     // 48 kHz presetCheaper
    if (g.inner == 1024 && g.outer == 5) return pow2_ffts_t<INV, 10, 5, 256>(g, tw, ar, ai, br, bi, tid, nt);   // kiosk blockMs 200
    if (g.inner == 2048 && g.outer == 3) return pow2_ffts_t<INV, 11, 3, 256>(g, tw, ar, ai, br, bi, tid, nt);   // 96 kHz presetDefault
  }
#endif
  return pow2_ffts_t<INV, 0, 0, 0>(g, tw, ar, ai, br, bi, tid, nt);
}

// outer twiddles + final DFT-3 / DFT-5 across the sub-transforms, in place (plan steps 8 and 10/12; W#35 4404-4625
// forward, W#48 10408-10628 inverse)
template <bool INV, int OUTER>
BS_HD void outer_stage_t(const DevGeom &g, const DevTables &T, float *dr, float *di, int tid, int nt) {
  const int inner = g.inner, outer = OUTER ? OUTER : g.outer;   // OUTER fixed: the per-sub-transform arrays stay in registers
  if (outer < 2) return;
  constexpr int OB = OUTER == 3 ? 4 : (OUTER == 5 ? 2 : 1);   // bins per trip whose twiddles are fetched together
  for (int i0 = tid; i0 < inner; i0 += nt * OB) {
   float twr[OB][4], twi[OB][4];
#pragma unroll
   for (int b = 0; b < OB; ++b) {
     const int i = i0 + b * nt;
#pragma unroll
     for (int s = 1; s < (OUTER ? OUTER : 5); ++s) {
       if (s >= outer || i >= inner) break;
       twr[b][s - 1] = T.otr[i + inner * (s - 1)]; twi[b][s - 1] = T.oti[i + inner * (s - 1)];
     }
   }
#pragma unroll
   for (int b = 0; b < OB; ++b) {
    const int i = i0 + b * nt;
    if (i >= inner) break;
    float xr[5], xi[5];
    xr[0] = dr[i]; xi[0] = di[i];
#pragma unroll
    for (int s = 1; s < (OUTER ? OUTER : 5); ++s) {
      if (s >= outer) break;
      float vr = dr[i + s * inner], vi = di[i + s * inner];
      float wr = twr[b][s - 1], wi = twi[b][s - 1];
      if (!INV) { xr[s] = (wr * vr) - (wi * vi); xi[s] = (wi * vr) + (vi * wr); }
      else { xr[s] = (vi * wi) + (vr * wr); xi[s] = (vi * wr) - (wi * vr); }
    }
    if (outer == 3) {
      const float h = INV ? 0x1.bb67aep-1f : -0x1.bb67aep-1f;
      float ar = xr[0], br = xr[1], cr = xr[2], ai = xi[0], bi = xi[1], ci = xi[2];
      dr[i] = (br + ar) + cr; di[i] = ci + (bi + ai);
      float p = ar + (br * -0.5f), q = bi * h, r = cr * -0.5f, t = ci * h;
      float u = ai + (bi * -0.5f), v = br * h, x = cr * h, y = ci * -0.5f;
      dr[i + inner] = ((p - q) + r) + t;     di[i + inner] = ((u + v) - x) + y;
      dr[i + 2 * inner] = ((p + q) + r) - t; di[i + 2 * inner] = ((u - v) + x) + y;
    } else {
      const float c1 = 0x1.3c6ef4p-2f, c2 = 0x1.9e377ap-1f, s1 = 0x1.e6f0e2p-1f, s2 = 0x1.2cf23p-1f;
      float ar = xr[0], br = xr[1], cr = xr[2], d_r = xr[3], er = xr[4];
      float ai = xi[0], bi = xi[1], ci = xi[2], d_i = xi[3], ei = xi[4];
      float dcR = d_r + cr, ebR = er + br, dcI = d_i + ci, ebI = ei + bi;
      dr[i] = (dcR + ar) + ebR; di[i] = (ai + dcI) + ebI;
      float p1r = ar + ((ebR * c1) - (dcR * c2)), p1i = ai + ((ebI * c1) - (dcI * c2));
      float p2r = ar + ((dcR * c1) - (ebR * c2)), p2i = ai + ((dcI * c1) - (ebI * c2));
      float q1r, q1i, q2r, q2i;
      if (!INV) {
        float a = d_i - ci, b = ei - bi, c = cr - d_r, d = br - er;
        q1r = (a * -s2) - (b * s1); q1i = (c * -s2) - (d * s1);
        q2r = (b * -s2) + (a * s1); q2i = (d * -s2) + (c * s1);
      } else {
        float a = ei - bi, b = d_i - ci, c = br - er, d = cr - d_r;
        q1r = (a * s1) + (b * s2); q1i = (c * s1) + (d * s2);
        q2r = (a * s2) - (b * s1); q2i = (c * s2) - (d * s1);
      }
      dr[i + inner] = p1r + q1r;     di[i + inner] = p1i + q1i;
      dr[i + 2 * inner] = p2r + q2r; di[i + 2 * inner] = p2i + q2i;
      dr[i + 3 * inner] = p2r - q2r; di[i + 3 * inner] = p2i - q2i;
      dr[i + 4 * inner] = p1r - q1r; di[i + 4 * inner] = p1i - q1i;
    }
   }
  }
  BS_SYNC();
}

template <bool INV>
BS_HD void outer_stage(const DevGeom &g, const DevTables &T, float *dr, float *di, int tid, int nt) {
  if (g.outer == 3) outer_stage_t<INV, 3>(g, T, dr, di, tid, nt);
  else if (g.outer == 5) outer_stage_t<INV, 5>(g, T, dr, di, tid, nt);
  else outer_stage_t<INV, 0>(g, T, dr, di, tid, nt);
}

// Pitch of the four FFT work arrays in shared memory: room for 4 padding floats per 64 (pow2_ffts_t pads the one array
// whose next reader would otherwise hit eight-way bank conflicts), rounded to whole 16-byte pieces.
BS_HHD int fft_pitch(int M) { return (M + (M >> 4) + 3) & ~3; }

// position of packed sample j after the interleave step of the split FFT (plan types 1-5): j = i*outer + s -> s*inner + i
BS_HD int deint(const DevGeom &g, int j) {
  const int q = (int)(((unsigned)j * g.divMagic) >> g.divShift);   // j / outer without a division or a branch
  return (j - q * g.outer) * g.inner + q;
}

// S1 (W#48 8238-8300): prevInput *= rot, the per-bin rotation table
BS_HD cf rot_prev(cf v, cf r) { cf o; o.re = (v.re * r.re) - (v.im * r.im); o.im = (v.im * r.re) + (v.re * r.im); return o; }

// ------------------------------------------------------------------------------------------------------------
// analysis of one window of one channel (W#35): window, zero-phase rotate, zero-pad, modified real FFT.
// smem: 4*fft_pitch(M) floats.  `x` = channel base of the clip.
BS_HD void analyse_window(const DevGeom &g, const DevTables &T, const float *x, Window w, cf *X, float *sm, int tid, int nt,
                          bool rotate = false /* store X[k] * specRot[k]: the S1 rotation of a "previous" spectrum, applied once here */,
                          float *E = nullptr /* [B]: also store |X[k]|^2, the block's input energies (map_energy then only sums them) */) {
  const int M = g.M, N = g.N, L = g.L, off = g.off, MP = fft_pitch(M);
  float *ar = sm, *ai = sm + MP, *br = sm + 2 * MP, *bi = sm + 3 * MP;
  // window * sample for the packed pair (2j, 2j+1): second half of the window first (zero-phase rotation), the first
  // half at the end with the half-bin shift's sign flip, zeros between.
  const int nA = L - off, cStart = N - off;
  if (((nA | cStart | off) & 1) == 0) {
    // common case: the region boundaries fall between pairs -- the two window coefficients of a pair come as one 8-byte
    // load.  Samples outside [lo, hi) read as zero: the clip's ends, and the first H samples of every "previous" window
    // of presetDefault, whose pre-roll is one interval short.
    const float *xs = x + w.start;
    const int lo = w.lo, hi = w.hi;
    const int jA = nA >> 1, jC = cStart >> 1;
    constexpr int UN = 6;   // six pairs per trip: all their loads are issued before the first one is used
    for (int j0 = tid; j0 < M; j0 += nt * UN) {
      float x0[UN], x1[UN]; f2 wv[UN]; cf rt[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int j = j0 + u * nt;
        const bool inA = j < jA, live = j < M && (inA || j >= jC);
        const int i = live ? (inA ? 2 * j + off : 2 * j - cStart) : 0;
        x0[u] = (live && i >= lo && i < hi) ? xs[i] : 0.f; x1[u] = (live && i + 1 >= lo && i + 1 < hi) ? xs[i + 1] : 0.f;
        wv[u] = *(const f2 *)(T.win + i);
        rt[u] = T.rot[j < M ? j : 0];
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int j = j0 + u * nt;
        if (j < M) {
          const bool inA = j < jA, live = inA || j >= jC;
          const float t0 = live ? x0[u] * (inA ? wv[u].x : -wv[u].x) : 0.f, t1 = live ? x1[u] * (inA ? wv[u].y : -wv[u].y) : 0.f;
          const cf r = rt[u];
          const int d = deint(g, j);
          ar[d] = (r.re * t0) - (r.im * t1); ai[d] = (r.im * t0) + (r.re * t1);
        }
      }
    }
  } else {
    for (int j = tid; j < M; j += nt) {
      float t[2];
      for (int e = 0; e < 2; ++e) {
        const int n = 2 * j + e;
        const bool inA = n < nA, inC = n >= cStart;
        if (!(inA || inC)) { t[e] = 0.f; continue; }
        const int i = inA ? n + off : n - cStart;
        const float xv = (i >= w.lo && i < w.hi) ? x[w.start + i] : 0.f;
        const float wv = T.win[i];
        t[e] = xv * (inA ? wv : -wv);
      }
      const cf r = T.rot[j];
      const int d = deint(g, j);
      ar[d] = (r.re * t[0]) - (r.im * t[1]); ai[d] = (r.im * t[0]) + (r.re * t[1]);
    }
  }
  BS_SYNC();
  int which = pow2_ffts<false>(g, T.tw, ar, ai, br, bi, tid, nt);
  float *rr = which ? br : ar, *ri = which ? bi : ai;
  outer_stage<false>(g, T, rr, ri, tid, nt);
  const int half = M >> 1;
  constexpr int UT = 4;   // four bin pairs per trip, their twiddle loads issued together
  for (int i0 = tid; i0 <= half; i0 += nt * UT) {
    cf uu[UT];
#pragma unroll
    for (int t = 0; t < UT; ++t) { const int i = i0 + t * nt; if (i <= half) uu[t] = T.untangle[i]; }
#pragma unroll
    for (int t = 0; t < UT; ++t) {
      const int i = i0 + t * nt;
      if (i > half || (i == half - 1 && !(M & 1))) continue;  // iteration i = M/2 rewrites pair half-1 last in the reference loop
      const int j = M - 1 - i; const cf u = uu[t];
      float sI = (ri[j] + ri[i]) * 0.5f, dR = (rr[i] - rr[j]) * 0.5f;
      float p = (sI * u.re) + (dR * u.im), dI = (ri[i] - ri[j]) * 0.5f;
      float q = (dR * u.re) - (sI * u.im), sR = (rr[j] + rr[i]) * 0.5f;
      cf xi_, xj_;
      xi_.im = p + dI; xi_.re = q + sR; xj_.im = p - dI; xj_.re = sR - q;
      if (rotate) { xi_ = rot_prev(xi_, T.specRot[i]); xj_ = rot_prev(xj_, T.specRot[j]); }
      X[i] = xi_; X[j] = xj_;
      if (E) { E[i] = (xi_.im * xi_.im) + (xi_.re * xi_.re); E[j] = (xj_.im * xj_.im) + (xj_.re * xj_.re); }
    }
  }
  BS_SYNC();
}

// inverse modified real FFT of one channel's output spectrum, synthesis window applied: the block's contribution to
// the output, frame[i] for output sample frameStart + i (W#48 9986-10932; the first half of the window carries the
// half-bin shift's sign flip, `ring -= t*w` there = adding -(t*w)).  smem: 4*fft_pitch(M) floats.
BS_HD void synth_frame(const DevGeom &g, const DevTables &T, const cf *X, float *frame, float *sm, int tid, int nt) {
  const int M = g.M, N = g.N, L = g.L, off = g.off, MP = fft_pitch(M);
  float *ar = sm, *ai = sm + MP, *br = sm + 2 * MP, *bi = sm + 3 * MP;
  const int half = M >> 1;
  constexpr int UN = 4;   // four bin pairs per trip, their global loads issued together
  for (int i0 = tid; i0 <= half; i0 += nt * UN) {
    cf uu[UN], xa[UN], xb[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int i = i0 + u * nt;
      if (i <= half) { uu[u] = T.untangle[i]; xa[u] = X[i]; xb[u] = X[M - 1 - i]; }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int i = i0 + u * nt;
      if (i > half || (i == half - 1 && !(M & 1))) continue;   // pair half-1 is rewritten by pair half in the reference loop
      const int j = M - 1 - i; const cf un = uu[u], xi_ = xa[u], xj_ = xb[u];
      float sI = xj_.im + xi_.im, dR = xi_.re - xj_.re;
      float p = (sI * un.im) + (dR * un.re), sR = xj_.re + xi_.re;
      float q = (sI * un.re) - (dR * un.im), dI = xi_.im - xj_.im;
      int di_ = deint(g, i), dj = deint(g, j);
      ar[di_] = p + sR; ai[di_] = q + dI; ar[dj] = sR - p; ai[dj] = q - dI;
    }
  }
  BS_SYNC();
  int which = pow2_ffts<true>(g, T.tw, ar, ai, br, bi, tid, nt);
  float *rr = which ? br : ar, *ri = which ? bi : ai;
  outer_stage<true>(g, T, rr, ri, tid, nt);
  const int nA = L - off, cStart = N - off;
  const bool evenCuts = ((nA | cStart | off) & 1) == 0;
  if (evenCuts) {
    // both samples of a pair fall in the same half of the window (one 8-byte coefficient load, one 8-byte store); four
    // pairs per trip, all their table loads issued before the first one is used
    constexpr int UO = 4;
    const int jA = nA >> 1, jC = cStart >> 1;
    for (int j0 = tid; j0 < M; j0 += nt * UO) {
      cf rt[UO]; f2 wv[UO]; int ii[UO];
#pragma unroll
      for (int u = 0; u < UO; ++u) {
        const int j = j0 + u * nt;
        const bool inA = j < jA, live = j < M && (inA || j >= jC);
        ii[u] = live ? (inA ? 2 * j + off : 2 * j - cStart) : -1;
        if (live) { rt[u] = T.rot[j]; wv[u] = *(const f2 *)(T.win + ii[u]); }
      }
#pragma unroll
      for (int u = 0; u < UO; ++u) {
        const int j = j0 + u * nt;
        if (ii[u] < 0) continue;
        const cf r = rt[u];
        const float t1 = (r.re * ri[j]) - (r.im * rr[j]);
        const float t0 = (r.im * ri[j]) + (r.re * rr[j]);
        f2 o;
        if (j < jA) { o.x = t0 * wv[u].x; o.y = t1 * wv[u].y; }
        else { o.x = -(t0 * wv[u].x); o.y = -(t1 * wv[u].y); }
        *(f2 *)(frame + ii[u]) = o;
      }
    }
    BS_SYNC();
    return;
  }
  for (int j = tid; j < M; j += nt) {
    const cf r = T.rot[j];
    const float t1 = (r.re * ri[j]) - (r.im * rr[j]);
    const float t0 = (r.im * ri[j]) + (r.re * rr[j]);
    {
      const float tv[2] = {t0, t1};
      for (int e = 0; e < 2; ++e) {
        const int n = 2 * j + e;
        if (n < nA) { const int i = n + off; frame[i] = tv[e] * T.win[i]; }
        else if (n >= cStart) { const int i = n - cStart; frame[i] = -(tv[e] * T.win[i]); }
      }
    }
  }
  BS_SYNC();
}

// Overlap-add of one output sample (W#48 10700-10932 + the per-sample read 7990-8080): the partial sum carried from
// earlier chunks, plus this chunk's frames that cover the sample, added in block order exactly like the reference's ring
// (whose slot starts from 0.0 each time it is re-used), then either emitted (divided by the window-product sum) or
// kept as the partial sum for the next chunk.  All positions are relative to ringBase (the first sample the carried
// ring describes), so the per-sample arithmetic is 32-bit:
//   x        sample index - ringBase;  emit range [0, xE1);  frame of chunk slot t starts at (t + fsOff) * H
//   rbModL   ringBase % L;  wpPhase  (ringBase - wpStartLen) mod H  (only used once n >= wpStartLen)
struct OlaGeom { long long ringBase; int xE1, fsOff, rbModL, wpPhase, nv, addFrames; };
BS_HHD OlaGeom ola_geom(const DevGeom &g, long long slot0, int nv, int mode) {
  OlaGeom o;
  const bool emit = (mode & kSynthEmit) != 0;
  o.ringBase = (emit ? slot0 : slot0 + g.split) * (long long)g.H;
  o.xE1 = emit ? nv * g.H : 0;
  o.fsOff = emit ? g.split : 0;
  o.rbModL = (int)(o.ringBase % g.L);
  long long ph = (o.ringBase - g.wpStartLen) % g.H; if (ph < 0) ph += g.H;
  o.wpPhase = (int)ph;
  o.nv = nv; o.addFrames = (mode & kSynthAdd) ? 1 : 0;
  return o;
}
// the per-sample read of four finished samples x..x+3: divided by the window-product sum, stored while the stream is live
BS_HD void ola_emit_quad(const DevGeom &g, const DevTables &T, const StreamDev &sd, int c, int x, const OlaGeom &o, const f4 acc) {
  const int H = g.H;
  const long long n = o.ringBase + x;
  if (n < sd.nLive) {
    f4 wp;
    if (n + 3 < g.wpStartLen) wp = *(const f4 *)(T.wpStart + n);
    else if (n >= g.wpStartLen) { int q = o.wpPhase + x; q -= (q / H) * H; wp = *(const f4 *)(T.wpSteady + q); }
    else {   // the quad straddles the end of the start-up table
      float w[4];
      for (int i = 0; i < 4; ++i) { const long long ni = n + i; int q = o.wpPhase + x + i; q -= (q / H) * H; w[i] = ni < g.wpStartLen ? T.wpStart[ni] : T.wpSteady[q]; }
      wp.x = w[0]; wp.y = w[1]; wp.z = w[2]; wp.w = w[3];
    }
    float *dst = sd.out + (size_t)c * sd.outStride + (n - sd.outBase);
    const float r0 = acc.x / wp.x, r1 = acc.y / wp.y, r2 = acc.z / wp.z, r3 = acc.w / wp.w;
    dst[0] = r0;   // (scalar stores: the channel stride of `out` is the stream's length, any alignment)
    if (n + 1 < sd.nLive) dst[1] = r1;
    if (n + 2 < sd.nLive) dst[2] = r2;
    if (n + 3 < sd.nLive) dst[3] = r3;
  }
}
// four consecutive samples at once (x, L, H, the ring base and the frame starts all multiples of 4): same additions in
// the same order per sample, 16-byte loads, the index divisions shared by the four
BS_HD void ola_quad(const DevGeom &g, const DevTables &T, const StreamDev &sd, int c, int x, const OlaGeom &o,
                    const float *frames, const float *ringOld, float *ringNew) {
  const int L = g.L, H = g.H;
  int p = o.rbModL + x; p -= (p / L) * L;
  f4 acc = {0.f, 0.f, 0.f, 0.f};
  if (x < L) acc = *(const f4 *)(ringOld + p);
  if (o.addFrames) {
    int tLo = (x - L >= 0 ? (x - L) / H + 1 : 0) - o.fsOff, tHi = x / H - o.fsOff;
    if (tLo < 0) tLo = 0;
    if (tHi > o.nv - 1) tHi = o.nv - 1;
    for (int t = tLo; t <= tHi; ++t) {
      const f4 v = *(const f4 *)(frames + ((size_t)t * g.C + c) * L + (x - (t + o.fsOff) * H));
      acc.x = acc.x + v.x; acc.y = acc.y + v.y; acc.z = acc.z + v.z; acc.w = acc.w + v.w;
    }
  }
  if (x < o.xE1) ola_emit_quad(g, T, sd, c, x, o, acc);
  else *(f4 *)(ringNew + p) = acc;
}
BS_HHD bool ola_quad_ok(const DevGeom &g) { return (g.L % 4 == 0) && (g.H % 4 == 0) && (g.wpStartLen % 4 == 0); }

BS_HD void ola_sample(const DevGeom &g, const DevTables &T, const StreamDev &sd, int c, int x, const OlaGeom &o,
                      const float *frames /* [slot][C][L] of this stream */, const float *ringOld, float *ringNew /* [L] */) {
  const int L = g.L, H = g.H;
  int p = o.rbModL + x; p -= (p / L) * L;
  float acc = (x < L) ? ringOld[p] : 0.f;
  if (o.addFrames) {
    int tLo = (x - L >= 0 ? (x - L) / H + 1 : 0) - o.fsOff, tHi = x / H - o.fsOff;
    if (tLo < 0) tLo = 0;
    if (tHi > o.nv - 1) tHi = o.nv - 1;
    for (int t = tLo; t <= tHi; ++t) acc = acc + frames[((size_t)t * g.C + c) * L + (x - (t + o.fsOff) * H)];
  }
  if (x < o.xE1) {
    const long long n = o.ringBase + x;
    if (n < sd.nLive) {
      float wp;
      if (n < g.wpStartLen) wp = T.wpStart[n];
      else { int q = o.wpPhase + x; q -= (q / H) * H; wp = T.wpSteady[q]; }
      sd.out[(size_t)c * sd.outStride + (n - sd.outBase)] = acc / wp;
    }
  } else ringNew[p] = acc;
}

// ------------------------------------------------------------------------------------------------------------
// spectral stage helpers
BS_HD float map_freq(float f, float mult, float limit) {
  if (!(f <= limit)) return ((mult + -1.0f) * limit) + f;
  return mult * f;
}
// One-pole smoother, backward then forward (W#48 8420-8520); the recurrence s <- s + (v[i]-s)*slew is strictly serial, so
// the loop keeps 16 samples in registers at a time: 16 independent loads, 16 dependent updates, 16 stores.
BS_HD float smooth_pass(float *v, int n, float slew, float s) {
  constexpr int U = 16;
  int i = n;
  for (; i >= U; i -= U) {
    float x[U];
#pragma unroll
    for (int j = 0; j < U; ++j) x[j] = v[i - U + j];
#pragma unroll
    for (int j = U - 1; j >= 0; --j) { s = ((x[j] - s) * slew) + s; x[j] = s; }
#pragma unroll
    for (int j = 0; j < U; ++j) v[i - U + j] = x[j];
  }
  for (--i; i >= 0; --i) { s = ((v[i] - s) * slew) + s; v[i] = s; }
  i = 0;
  for (; i + U <= n; i += U) {
    float x[U];
#pragma unroll
    for (int j = 0; j < U; ++j) x[j] = v[i + j];
#pragma unroll
    for (int j = 0; j < U; ++j) { s = ((x[j] - s) * slew) + s; x[j] = s; }
#pragma unroll
    for (int j = 0; j < U; ++j) v[i + j] = x[j];
  }
  for (; i < n; ++i) { s = ((v[i] - s) * slew) + s; v[i] = s; }
  return s;
}
// interpolation in a guarded array (see kGuard): elements outside [0, B) read as zero
BS_HD int guard_index(int low, int B) { low = low < -kGuard ? -kGuard : low; return low > B ? B : low; }
BS_HD cf lerp_c(const cf *a, int B, int low, float fr) {
  const int l = guard_index(low, B);
  const cf lo = a[l], hi = a[l + 1];
  cf r;
  r.re = ((hi.re - lo.re) * fr) + lo.re; r.im = ((hi.im - lo.im) * fr) + lo.im;
  return r;
}
BS_HD float lerp_f(const float *a, int B, int low, float fr) {
  const int l = guard_index(low, B);
  const float lo = a[l], hi = a[l + 1];
  return ((hi - lo) * fr) + lo;
}
// minstd_rand: state after n steps from x (x_{k+1} = 48271 x_k mod 2^31-1), by square-and-multiply
BS_HHD uint32_t minstd_jump(uint32_t x, uint32_t n) {
  unsigned long long a = 48271ull, r = 1ull; const unsigned long long m = 2147483647ull;
  while (n) { if (n & 1u) r = (r * a) % m; a = (a * a) % m; n >>= 1; }
  return (uint32_t)((r * (unsigned long long)x) % m);
}
// x / d and sqrt(x) where d is a positive normal number.  A zero x passes through unchanged, which is what IEEE division
// and square root give (+-0 / d = +-0, sqrt(+-0) = +-0); it is singled out only so that the hardware divide / sqrt
// sequences never see a zero operand, which would send the whole warp through their slow path.  Silent bins (pitched
// down above the input's Nyquist, or digital silence) are full of such zeros.
#ifdef BS_HOSTEMU
#define BS_OPAQUE(v) ((void)0)
#else
#define BS_OPAQUE(v) asm volatile("" : "+f"(v))   // keeps the compiler from folding the substitute operand back in
#endif
BS_HD float div_pos(float x, float d) { const bool z = (x == 0.f); float xs = z ? 1.0f : x; BS_OPAQUE(xs); const float q = xs / d; return z ? x : q; }
BS_HD float sqrt_z(float x) { const bool z = (x == 0.f); float xs = z ? 1.0f : x; BS_OPAQUE(xs); const float q = sqrtf(xs); return z ? x : q; }
BS_HD void make_output(float energy, cf fallback, float re, float im, float &ore, float &oim) {
  float n2 = (im * im) + (re * re), div;
  if (n2 > 1e-15f) div = n2;
  else { re = fallback.re; im = fallback.im; div = ((re * re) + 1e-15f) + (im * im); }
  float s = sqrt_z(div_pos(energy, div));
  oim = s * im; ore = s * re;
}

// ------------------------------------------------------------------------------------------------------------
// Spectral stage.  Split by what depends on the carried phase state (Band.output):
//
//   map stage      (premap_kernel)    input energies, energy smoothing, peak picking, output frequency map, formant
//                                     envelope (W#48 8238-9312).  State-free except the formant auto-detect, whose two
//                                     leaky averages are advanced by freqest_kernel between phase A and phase B.
//   term stage     (preterms_kernel)  every coefficient of the preliminary prediction (S5) and of the vertical chain
//                                     (S6) that does not involve Band.output: one record per (block, bin).
//   chain stage    (chain_kernel)     the recurrence itself: bin k of block m needs bins k-1, k-longStep of block m and
//                                     bins k+1, k+longStep of block m-1, so consecutive blocks of one stream run as a
//                                     wavefront -- lane j of a warp walks block m0+j, `lag` bins behind lane j-1.
//
// The map stage runs as a short pipeline of kernels per chunk; the strictly serial parts (the one-pole smoothers) run
// with ONE THREAD PER BLOCK -- 32 blocks per warp in lock step -- instead of one thread per CTA:
//   energy   (bins in parallel)      inputEnergy, band energy sums, smoother input, sqrt metric for the formants
//   smooth   (one thread per block)  smoothEnergy steps 1,2 on the band energies
//   peaks    (CTA per block)         findPeaks + updateOutputMap; formant auto-detect peak pick
//   freqest  (one thread per stream) leaky averages of the formant base over the blocks
//   fmsmooth (one thread per block)  formant envelope smoothing
//   fmapply  (bins in parallel)      formant envelope applied to the input energies
BS_HHD int fm_pitch(int B) { return (B + 2 + 31) & ~31; }   // whole 128-byte lines: the smoother's tiled path needs aligned rows
BS_HHD bool fm_auto(const BlockRec &rec) { return (rec.flags & kFormants) && !(rec.fmBaseFreq > 0.f); }

// energy: per bin.  energy[k] = sum over channels in channel order, exactly as the reference accumulates it.
template <int CT>
BS_HD void map_energy(const DevGeom &g, const BlockRec rec, const cf *inp, float *inEnergy, float *energy, float *smoothed, float *fm,
                      float *mapv, int tid, int nt) {
  const int C = CT > 0 ? CT : g.C, B = g.B;
  const bool mapped = rec.flags & kMapped, formants = rec.flags & kFormants;
  const bool have = rec.flags & kNew;   // the block was analysed in this chunk: its input energies are in place already
  for (int k = tid; k < B; k += nt) {
    float e = 0.f;
    for (int c = 0; c < C; ++c) {
      float en;
      if (have) en = inEnergy[(size_t)c * guard_pitch(B) + k];   // stored by the analysis of this block, same expression
      else {
        const cf v = inp[(size_t)c * guard_pitch(B) + k];
        en = (v.im * v.im) + (v.re * v.re);
        inEnergy[(size_t)c * guard_pitch(B) + k] = en;
      }
      e = e + en;
    }
    if (mapped || formants) energy[k] = e;
    if (mapped) smoothed[k] = e;
    else { mapv[2 * k] = (float)(uint32_t)k; mapv[2 * k + 1] = 1.0f; }
    if (formants) fm[k] = sqrtf(e);
  }
  if (formants) for (int k = B + tid; k < fm_pitch(B); k += nt) fm[k] = 0.f;
}

// One-pole smoother over an array in global memory, one thread per array: backward then forward (W#48 8420-8520).  The
// next 16 samples are requested before the current 16 are processed, so a thread always has a load in flight.
BS_HD float smooth_pass_g(float *v, int n, float slew, float s) {
  constexpr int U = 32, Q = U / 4;   // one whole 128-byte line per tile
  if (n % U != 0 || (((size_t)v) & 127) != 0) {   // odd sizes: plain loop
    for (int i = n - 1; i >= 0; --i) { s = ((v[i] - s) * slew) + s; v[i] = s; }
    for (int i = 0; i < n; ++i) { s = ((v[i] - s) * slew) + s; v[i] = s; }
    return s;
  }
  f4 *v4 = (f4 *)v;
  const int nc = n / U;
  f4 cur[Q], nxt[Q];
  for (int dir = 0; dir < 2; ++dir) {
    int ch = dir == 0 ? nc - 1 : 0;
    const int step = dir == 0 ? -1 : 1;
#pragma unroll
    for (int j = 0; j < Q; ++j) cur[j] = v4[ch * Q + j];
    for (int it = 0; it < nc; ++it, ch += step) {
      const int chn = ch + step;
      if (it + 1 < nc) {
#pragma unroll
        for (int j = 0; j < Q; ++j) nxt[j] = v4[chn * Q + j];
      }
      float *x = (float *)cur;
      if (dir == 0) {
#pragma unroll
        for (int j = U - 1; j >= 0; --j) { s = ((x[j] - s) * slew) + s; x[j] = s; }
      } else {
#pragma unroll
        for (int j = 0; j < U; ++j) { s = ((x[j] - s) * slew) + s; x[j] = s; }
      }
#pragma unroll
      for (int j = 0; j < Q; ++j) v4[ch * Q + j] = cur[j];
#pragma unroll
      for (int j = 0; j < Q; ++j) cur[j] = nxt[j];
    }
  }
  return s;
}

// smem of the peaks stage: energy[B] f32 | cpk[B/2+2] i32 | peaks[B] f32 | misc[16] i32 | wordCnt[W+1] i32 | mask[W+1] u32,
// W = ceil(B/32)
BS_HHD size_t map_smem_floats(int B) { return (((size_t)B + (B / 2 + 2) + B + 16 + 2 * ((B + 31) / 32 + 2)) + 3) & ~(size_t)3; }
BS_HD int popc_hd(uint32_t x) {
#ifdef BS_HOSTEMU
  return __builtin_popcount(x);
#else
  return __popc(x);
#endif
}
BS_HD int ffs_hd(uint32_t x) {   // index of the lowest set bit (x != 0)
#ifdef BS_HOSTEMU
  return __builtin_ctz(x);
#else
  return __ffs((int)x) - 1;
#endif
}

// peaks: findPeaks + updateOutputMap from the band energies and their smoothed copy (both in global memory); for
// auto-detect formant blocks also the spectral-peak pick feeding the base estimate
BS_HD void map_peaks(const DevGeom &g, const BlockRec rec, const float *energy, const float *smoothed, float *mapv,
                     float *fmAuto /* [2]: top, i1 (int bits) */, float *sm, int tid, int nt) {
  const int B = g.B;
  const bool mapped = rec.flags & kMapped;
  const float fN = (float)(uint32_t)g.N;
  const int nWords = (B + 31) / 32;
  float *smE = sm;
  int *cpk = (int *)(smE + B);
  float *peaksG = (float *)(cpk + (B / 2 + 2));
  int *misc = (int *)(peaksG + B);   // [0] nPeaks, [1] cpk non-decreasing, [2] cpk strictly increasing, [3] peaks with cpk < 0
  int *wordCnt = misc + 16;          // run starts per 32-bin word, then their exclusive prefix sum
  uint32_t *mask = (uint32_t *)(wordCnt + (nWords + 2));   // bit k&31 of word k>>5: energy[k] > smoothed[k]
  if (mapped) {
    // findPeaks (W#48 8560-8700): a peak = a maximal run of bins with energy > smoothed.  Runs are independent: the
    // comparison flags are packed into bit words (coalesced reads, warp ballot) and the energies staged in shared
    // memory; every thread then owns the runs that START in its 32-bin word and follows them past the word's end; the
    // peak index is the number of run starts before it (prefix sum over the words).
#ifdef BS_HOSTEMU
    for (int w = tid; w < nWords; w += nt) {
      uint32_t m = 0;
      for (int j = 0; j < 32 && 32 * w + j < B; ++j) { const int k = 32 * w + j; smE[k] = energy[k]; if (!(energy[k] <= smoothed[k])) m |= 1u << j; }
      mask[w] = m;
    }
#else
    for (int k0 = (tid & ~31); k0 < nWords * 32; k0 += nt) {   // nt is a multiple of 32: a warp covers one word per trip
      const int k = k0 + (tid & 31);
      bool ab = false;
      if (k < B) { const float e = energy[k]; smE[k] = e; ab = !(e <= smoothed[k]); }
      const uint32_t m = __ballot_sync(0xffffffffu, ab);
      if ((tid & 31) == 0) mask[k0 >> 5] = m;
    }
#endif
    BS_SYNC();
    for (int w = tid; w < nWords; w += nt) {
      const uint32_t m = mask[w], carry = w > 0 ? (mask[w - 1] >> 31) : 0u;
      wordCnt[w] = popc_hd(m & ~((m << 1) | carry));
    }
    BS_SYNC();
    if (tid == 0) { int acc = 0; for (int w = 0; w < nWords; ++w) { int c = wordCnt[w]; wordCnt[w] = acc; acc += c; } misc[0] = acc; misc[1] = 1; misc[2] = 1; misc[3] = 0; }
    BS_SYNC();
    for (int w = tid; w < nWords; w += nt) {
      const uint32_t m = mask[w], carry = w > 0 ? (mask[w - 1] >> 31) : 0u;
      uint32_t starts = m & ~((m << 1) | carry);
      int nP = wordCnt[w];
      while (starts) {
        const int k = 32 * w + ffs_hd(starts);
        starts &= starts - 1;
        float sum = 0.f, wsum = 0.f;
        for (int kk = k; kk < B && ((mask[kk >> 5] >> (kk & 31)) & 1u); ++kk) { const float en = smE[kk]; sum = en + sum; wsum = (en * (float)kk) + wsum; }
        float avg = wsum / sum;
        float f = (avg + 0.5f) / fN;
        float o = (map_freq(f, rec.pkMult, rec.pkLimit) * fN) + -0.5f;
        peaksG[2 * nP] = avg; peaksG[2 * nP + 1] = o;
        cpk[nP] = trunc_i32(ceilf(o));
        ++nP;
      }
    }
    BS_SYNC();
    // Section lookup for the (usual) strictly increasing cpk: the peaks' bins are marked in a bit array (the run
    // mask's storage, free by now) and "first p with cpk[p] > k" becomes a prefix count + a popcount.  Equal or
    // descending neighbours fall back to the searches below.
    const int nAll = misc[0];
    for (int w = tid; w <= nWords; w += nt) mask[w] = 0u;
    for (int i = tid + 1; i < nAll; i += nt) {   // benign races: every writer stores the same value
      if (cpk[i] < cpk[i - 1]) misc[1] = 0;
      if (cpk[i] <= cpk[i - 1]) misc[2] = 0;
    }
    BS_SYNC();
    const bool marked = misc[2] != 0;
    if (marked) {
      for (int i = tid; i < nAll; i += nt) {
        const int c = cpk[i];
        if (c < 0) {
#ifdef BS_HOSTEMU
          misc[3] += 1;
#else
          atomicAdd(&misc[3], 1);
#endif
        } else if (c < B) {
#ifdef BS_HOSTEMU
          mask[c >> 5] |= 1u << (c & 31);
#else
          atomicOr(&mask[c >> 5], 1u << (c & 31));
#endif
        }
      }
    }
    BS_SYNC();
    if (marked && tid == 0) { int acc = misc[3]; for (int w = 0; w < nWords; ++w) { wordCnt[w] = acc; acc += popc_hd(mask[w]); } }
    BS_SYNC();
    // updateOutputMap: every bin finds the LAST section (in the reference's write order) that covers it
    const int nP = misc[0], mono = misc[1];
    for (int k = tid; k < B; k += nt) {
      float ib, gr = 1.0f;
      if (nP == 0) { ib = (float)(uint32_t)k; }
      else {
        float lIn = peaksG[2 * (nP - 1)], lOut = peaksG[2 * (nP - 1) + 1];
        int loLast = trunc_i32(lOut); if (loLast < 0) loLast = 0;
        int sec = -2;  // -2: no writer, -1: last section, 0: first section, p>=1: middle section p
        if (k >= loLast) sec = -1;
        else if (marked) {
          const int lo = wordCnt[k >> 5] + popc_hd(mask[k >> 5] & (0xffffffffu >> (31 - (k & 31))));  // peaks with cpk <= k
          if (lo < nP) sec = lo;
        } else if (mono) {
          int lo = 0, hi = nP;  // first p with cpk[p] > k
          while (lo < hi) { int mid = (lo + hi) >> 1; if (cpk[mid] > k) hi = mid; else lo = mid + 1; }
          if (lo < nP) sec = lo;  // lo == 0 -> first section (k < ceil(out0)); lo >= 1 -> between lo-1 and lo
        } else {
          for (int p = nP - 1; p >= 1 && sec == -2; --p) {
            int hi = cpk[p] > B ? B : cpk[p], lo = cpk[p - 1] < 0 ? 0 : cpk[p - 1];
            if (k >= lo && k < hi) sec = p;
          }
          if (sec == -2 && k < (cpk[0] > B ? B : cpk[0])) sec = 0;
        }
        if (sec == -2) continue;  // stale value stays, as in the reference
        float kf = (float)(uint32_t)k;
        if (sec == -1) ib = (lIn - lOut) + kf;
        else if (sec == 0) ib = (peaksG[0] - peaksG[1]) + kf;
        else {
          float pIn = peaksG[2 * (sec - 1)], pOut = peaksG[2 * (sec - 1) + 1], nIn = peaksG[2 * sec], nOut = peaksG[2 * sec + 1];
          float offs = pIn - pOut, inv = 1.0f / (nOut - pOut);
          float delta = (pOut - (nOut + pIn)) + nIn;
          float g6 = (inv * delta) * 6.0f;
          float rr = (kf - pOut) * inv;
          gr = ((g6 * rr) * (1.0f - rr)) + 1.0f;
          ib = (offs + kf) + (((rr * rr) * delta) * (3.0f - (rr + rr)));
        }
      }
      mapv[2 * k] = ib; mapv[2 * k + 1] = gr;
    }
  }
}

// formant auto-detect: the three largest local maxima of the channel-summed energy in scan order, then the harmonic
// fix-ups (W#48 8820-8900).  Strictly sequential over the bins; one thread per block.
BS_HD void fm_auto_pick(const DevGeom &g, const float *fm /* channel-summed energy [B] */, float *fmAuto /* [2]: top, i1 (int bits) */) {
  const int B = g.B;
  int i1 = 0, i2 = 0, i3 = 0;
  float f1 = fm[0], f2 = f1, f3 = f1;          // fm[i1], fm[i2], fm[i3] kept in registers
  float prev = fm[0], cur = B > 1 ? fm[1] : 0.f;
  for (int i = 1; i <= B - 2; ++i) {
    const float nxt = fm[i + 1], v = cur;
    const float below = prev;
    prev = cur; cur = nxt;
    if (v < below) continue;
    if (v <= nxt) continue;
    if (v <= f3) continue;
    if (f2 >= v) { i3 = i; f3 = v; continue; }
    if (f1 < v) { i3 = i2; f3 = f2; i2 = i1; f2 = f1; i1 = i; f1 = v; continue; }
    i3 = i2; f3 = f2; i2 = i; f2 = v;
  }
  float top = f1; double dtop = (double)top;
  if ((double)f2 > (dtop * 0.1)) {
    int d = i1 - i2; if (d < 0) d = -d;
    if (d > i1 / 8 && d < (i1 * 7) / 8) i1 = i1 % d;
    if (!((double)f3 <= (dtop * 0.01))) {
      d = i1 - i3; if (d < 0) d = -d;
      if (d > i1 / 8 && d < (i1 * 7) / 8) i1 = i1 % d;
    }
  }
  fmAuto[0] = top; fmAuto[1] = __int_as_float_hd(i1);
}

// the two leaky averages of the formant base estimate, one block (W#48 8900-8960); returns the base bin
BS_HD float freqest_step(float *freqEst, const float *fmAuto) {
  const float top = fmAuto[0]; const int i1 = __float_as_int_hd(fmAuto[1]);
  float w = freqEst[1];
  float nw = (float)(((double)(top - w) * 0.25) + (double)w);
  freqEst[1] = nw;
  float ww = freqEst[0];
  ww = (float)(((double)((top * (float)i1) - ww) * 0.25) + (double)ww);
  freqEst[0] = ww;
  return ww / (nw + 1e-30f);
}

// formant envelope smoothing, one thread per block (W#48 8962-9040): fm = sqrt(channel-summed energy), written by the
// energy stage; slew from the base bin (fixed, or the auto-detected estimate)
BS_HD float fm_slew(const DevGeom &g, const BlockRec rec, float baseBinAuto) {
  const float fN = (float)(uint32_t)g.N, base = rec.fmBaseFreq;
  const float baseBin = (base > 0.f) ? ((base * fN) + -0.5f) : baseBinAuto;
  return (float)(1.0 / (((double)baseBin * 0.5) + 1.0));
}
BS_HD void fm_smooth(const DevGeom &g, const BlockRec rec, float baseBinAuto, float *fm) {
  const float slew = fm_slew(g, rec, baseBinAuto);
  const float st = smooth_pass_g(fm, g.B, slew, 0.f);
  smooth_pass_g(fm, g.B, slew, st);
}
// formant envelope applied to the input energies (W#48 9042-9312), per bin
template <int CT>
BS_HD void fm_apply(const DevGeom &g, const BlockRec rec, const BlockRec2 rec2, const float *fm, float *inEnergy, int tid, int nt) {
  const int C = CT > 0 ? CT : g.C, B = g.B;
  const float fN = (float)(uint32_t)g.N;
  const bool comp = rec.flags & kFormantComp;
  for (int k = tid; k < B; k += nt) {
    float f = ((float)k + 0.5f) / fN;
    if (comp) f = map_freq(f, rec.fmFreqMult, rec2.fmLimit);
    float metric = fm[k], lim = rec2.fmLimit;
    float gq = rec.fmInv * f;
    float fi = (gq > lim) ? (((1.0f - rec.fmMult) * lim) + f) : gq;
    float pos = (fi * fN) + -0.5f, env = 0.f;
    if (!(pos < 0.f)) {
      float fB = (float)B, x = (fB < pos) ? fB : pos, fl = floorf(x), fr = x - fl;
      int idx = trunc_i32(fl);
      float lo = fm[idx];
      env = (fr * (fm[idx + 1] - lo)) + lo;
    }
    float g2 = env / (metric + 1e-30f); g2 = g2 * g2;
    for (int c = 0; c < C; ++c) { size_t o = (size_t)c * guard_pitch(B) + k; inEnergy[o] = g2 * inEnergy[o]; }
  }
}

// ---- term stage: per-(block, bin) records.  One row of NR floats per wavefront step of a block: row r holds the chain
// coefficients of bin k = r - R0 (R0 = longStep + 1) and the S5 coefficients of bin q = r, because that is the pair of
// bins a lane of the chain kernel works on in the same step.  Rows 0 .. B-1+R0.
//   0..3 up1.re up1.im upLong.re upLong.im | 4..7 down1.re down1.im downLong.re downLong.im | 8 maxChannel (int bits)
//   9+5c: energy, predIn.re, predIn.im, chanTwist.re, chanTwist.im            (bin k)
//   9+5C+3c: S5 twist.re, twist.im, divisor                                   (bin q)
BS_HHD constexpr int nr_floats(int C) { return (9 + 8 * C + 3) & ~3; }
BS_HHD int rec_rows(int B, int longStep) { return B + longStep + 1; }
// Storage is "wavefront-major": the slots of a chunk are grouped by 32 (one warp of the chain kernel), and inside a
// group the rows of the 32 blocks are interleaved along diagonals u = row + lane*lag, so that the 32 rows a warp
// consumes in one step are ONE contiguous run (32 row pitches) instead of 32 cursors in 32 different blocks' records.
// The row pitch is a whole number of 128-byte lines, so a row is always written and read as full lines.
// Stereo rows are stored compactly in 24 floats = 96 bytes (three whole 32-byte sectors): the logical row has 25 fields;
// the max-channel index travels in the sign bit of channel 1's energy (energies are never negative) and the fields
// behind it move up by one.  Other channel counts keep the logical row, padded to whole 128-byte lines.
// Three and more channels (chain_wide.cuh reads single fields of the rows of several blocks at once): the smallest pitch that
// is 8 modulo 32 floats, so that the same field of different blocks' rows falls into different shared-memory banks.
BS_HHD constexpr int nr_pitch(int C) { return C == 2 ? 24 : (C == 1 ? 32 : ((nr_floats(C) - 8 + 31) & ~31) + 8); }
// row stride of preterms' staging rows.  Stereo rows are staged in their stored (packed) form, 24 floats in a 28-float
// pitch: 16-byte stores by consecutive threads then fall into eight different bank groups.
BS_HHD int nr_stage(int C) { return C == 2 ? 28 : nr_pitch(C) + 4; }
BS_HD float pack2_field(const float *logical, int f) {                 // physical field f of a stereo row
  if (f < 8) return logical[f];
  float v = logical[f + 1];
  if (f == 13 && __float_as_int_hd(logical[8]) == 1) v = -v;            // logical[14] = energy of channel 1 (>= +0)
  return v;
}
BS_HD void unpack2_row(const float *phys, float *logical) {            // 24 physical -> 25 logical fields (28 allocated)
#pragma unroll
  for (int f = 0; f < 8; ++f) logical[f] = phys[f];
#pragma unroll
  for (int f = 8; f < 24; ++f) logical[f + 1] = phys[f];
  logical[8] = __int_as_float_hd((int)((unsigned)__float_as_int_hd(phys[13]) >> 31));
  logical[14] = __int_as_float_hd(__float_as_int_hd(phys[13]) & 0x7fffffff);
}
BS_HHD int chain_lag(int longStep) { return longStep + 2; }
// Three and more channels (chain_wide.cuh): warps of blocks per chain CTA, by how many streams there are to fill the GPU with.
// Every warp of a CTA executes every step, so a step of a full CTA costs 8 warps' instructions on one SM; with few streams the
// CTAs are smaller, and their S5 / S6 halves run on separate warps (wide_split).
#ifndef BS_WIDE_WARPS_MIN
#define BS_WIDE_WARPS_MIN 4
#endif
BS_HHD int wide_warps_for(int streams) { return streams >= 64 ? 8 : (streams >= 16 ? 4 : BS_WIDE_WARPS_MIN); }
BS_HHD size_t rec_group_floats(int B, int longStep, int C) { return (size_t)(rec_rows(B, longStep) + 31 * chain_lag(longStep)) * 32 * nr_pitch(C); }
BS_HHD size_t rec_row_stride(int C) { return (size_t)32 * nr_pitch(C); }
// row 0 of chunk slot `slot` of a stream whose chunk records start at `base`; row r is rec_row_stride floats further per row
BS_HHD size_t rec_slot_offset(int slot, int B, int longStep, int C) {
  const int lane = slot & 31, grp = slot >> 5;
  return (size_t)grp * rec_group_floats(B, longStep, C) + ((size_t)lane * chain_lag(longStep) * 32 + lane) * nr_pitch(C);
}

// Rows are produced in tiles of kTermTile bins staged in shared memory, so that every record row leaves the SM as part
// of one contiguous, 16-byte-vectorised burst (a row mixes two bins R0 apart, hence the R0 rows carried tile to tile).
#ifndef BS_TERM_TILE
#define BS_TERM_TILE 128
#endif
#ifndef BS_TERM_CTAS
#define BS_TERM_CTAS 10
#endif
constexpr int kTermTile = BS_TERM_TILE;
BS_HHD size_t preterms_smem_floats(int C, int longStep) { return (size_t)(kTermTile + longStep + 1) * nr_stage(C); }

template <int CT>
BS_HD void preterms_block(const DevGeom &g, const DevTables &T, const BlockRec rec, uint32_t rng0, const cf *inp,
                          const cf *inPrev /* [C][B]; nullptr: no new spectrum */, const float *inEnergy, const float *mapv,
                          const float *prevInE, const float *prevMap /* previous block of this stream in the chunk, or nullptr */,
                          const float *prevEState /* Prediction.energy carried from the previous chunk */,
                          float *predEOut /* nullptr unless this is the stream's last block of the chunk */,
                          float *recRows /* row 0 of this block; rows are rec_row_stride apart */, float *sm, int tid, int nt) {
  const int C = CT > 0 ? CT : g.C, B = g.B, R0 = g.longStep + 1, SO = 9 + 5 * C, TB = kTermTile;
  const int NRP = nr_pitch(C), NR = nr_stage(C);   // NR: row stride of the staging rows (padded against bank conflicts)
  const size_t rowStride = rec_row_stride(C);
  const bool isNew = rec.flags & kNew;
  const cf *prv = isNew ? inPrev : inp;   // (a new block's previous spectrum was rotated (S1) by the analysis kernel when it was stored)
  const int longStep = g.longStep, BP = guard_pitch(B);
  const float tf = rec.timeFactor < 0.5f ? 0.5f : rec.timeFactor;
  const float rlo = ((tf > 2.0f) ? 4.0f : 0.0f) - tf, rscale = (tf - rlo) * 0x1p-31f, fLong = (float)longStep;
  const bool randomTF = !(tf <= 2.0f);
  for (int k0 = 0; k0 < B; k0 += TB) {
    const int nb = (B - k0 < TB) ? B - k0 : TB;
    for (int i = tid; i < nb; i += nt) {
      const int k = k0 + i;
      float *ra = sm + (size_t)(i + R0) * NR, *rb = sm + (size_t)i * NR + SO;   // local rows: chain part R0 rows further down
      // The loads below form three dependent levels (map entries -> per-channel gathers -> gathers of the maximum
      // channel); everything inside a level is issued together, boundary cases are selected away afterwards.
      const int kN = (k + 1 < B) ? k + 1 : k, kL = (k + longStep < B) ? k + longStep : k;
      const float ib = mapv[2 * k], grad = mapv[2 * k + 1], ib1 = mapv[2 * kN], ibL = mapv[2 * kL];
      const float fl = floorf(ib);
      const int low = trunc_i32(fl); const float fr = ib - fl;
      const float gpos = grad > 0.f ? grad : 0.f;
      int lowP = 0; float frP = 0.f, gP = 0.f;
      if (prevMap) {
        const float ibP = prevMap[2 * k], flP = floorf(ibP), grP = prevMap[2 * k + 1];
        lowP = trunc_i32(flP); frP = ibP - flP; gP = grP > 0.f ? grP : 0.f;
      }
      // S5 coefficients (W#48 9314-9455) and the maximum-energy channel
      int mc = 0; float me = 0.f, pRe = 0.f, pIm = 0.f;
      float enC[2], inRe[2], inIm[2], s5[2][3];   // stereo: the row stays in registers
      for (int c = 0; c < C; ++c) {
        const float en = lerp_f(inEnergy + (size_t)c * BP, B, low, fr) * gpos;
        const cf in = lerp_c(inp + (size_t)c * BP, B, low, fr);
        const cf pv = lerp_c(prv + (size_t)c * BP, B, low, fr);
        const float prevE = prevMap ? (lerp_f(prevInE + (size_t)c * BP, B, lowP, frP) * gP) : prevEState[(size_t)c * B + k];
        if (predEOut) predEOut[(size_t)c * B + k] = en;
        const float tIm = (pv.re * in.im) - (pv.im * in.re), tRe = (pv.im * in.im) + (pv.re * in.re);
        const float dv = ((en > prevE) ? en : prevE) + 1e-15f;
        if (CT == 2) { enC[c & 1] = en; inRe[c & 1] = in.re; inIm[c & 1] = in.im; s5[c & 1][0] = tRe; s5[c & 1][1] = tIm; s5[c & 1][2] = dv; }
        else {
          rb[3 * c] = tRe; rb[3 * c + 1] = tIm; rb[3 * c + 2] = dv;
          ra[9 + 5 * c] = en; ra[9 + 5 * c + 1] = in.re; ra[9 + 5 * c + 2] = in.im;
        }
        if (c == 0 || en > me) { me = en; mc = c; pRe = in.re; pIm = in.im; }
      }
      float twRe[2], twIm[2];
      for (int c = 0; c < C; ++c) {   // channel twists: predIn[c] * conj(predIn[mc])
        const float cRe = CT == 2 ? inRe[c & 1] : ra[9 + 5 * c + 1], cIm = CT == 2 ? inIm[c & 1] : ra[9 + 5 * c + 2];
        const float wRe = (pIm * cIm) + (pRe * cRe), wIm = (pRe * cIm) - (pIm * cRe);
        if (CT == 2) { twRe[c & 1] = wRe; twIm[c & 1] = wIm; }
        else { ra[9 + 5 * c + 3] = wRe; ra[9 + 5 * c + 4] = wIm; }
      }
      if (CT != 2) ra[8] = __int_as_float_hd(mc);
      const cf *ic = inp + (size_t)mc * BP;
      // S6 terms of the maximum channel (W#48 9458-9873): upward neighbours k-1, k-longStep and downward neighbours k+1,
      // k+longStep (whose predIn is re-interpolated here for channel mc).  With random time factors (timeFactor > 2)
      // the "up" pair and the "down" pair each draw one value, in bin order.
      float btfU = tf, btfD = tf;
      if (randomTF) {
        if (k > 0) { uint32_t x = minstd_jump(rng0, (uint32_t)(2 * k)); btfU = (rscale * (float)(uint32_t)(x - 1u)) + rlo; }
        if (k < B - 1) { uint32_t x = minstd_jump(rng0, (uint32_t)(2 * k + 1)); btfD = (rscale * (float)(uint32_t)(x - 1u)) + rlo; }
      }
      const float xU1 = ib - btfU, xUL = ib - (btfU * fLong), xD1 = ib1 - btfD, xDL = ibL - (btfD * fLong);
      const float fl1 = floorf(ib1), flL = floorf(ibL);
      const int lU1 = trunc_i32(floorf(xU1)), lUL = trunc_i32(floorf(xUL)), lD1 = trunc_i32(floorf(xD1)), lDL = trunc_i32(floorf(xDL));
      const cf dU1 = lerp_c(ic, B, lU1, xU1 - (float)lU1), dUL = lerp_c(ic, B, lUL, xUL - (float)lUL);
      const cf uN = lerp_c(ic, B, trunc_i32(fl1), ib1 - fl1), dD1 = lerp_c(ic, B, lD1, xD1 - (float)lD1);
      const cf uL = lerp_c(ic, B, trunc_i32(flL), ibL - flL), dDL = lerp_c(ic, B, lDL, xDL - (float)lDL);
      float u0 = 0.f, u1 = 0.f, u2 = 0.f, u3 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      if (k > 0) {
        u1 = (dU1.re * pIm) - (dU1.im * pRe); u0 = (dU1.im * pIm) + (dU1.re * pRe);
        if (k >= longStep) { u2 = (dUL.im * pIm) + (dUL.re * pRe); u3 = (dUL.re * pIm) - (dUL.im * pRe); }
      }
      if (k < B - 1) {
        d0 = (dD1.im * uN.im) + (dD1.re * uN.re); d1 = (dD1.re * uN.im) - (dD1.im * uN.re);
        if (k < B - longStep) { d2 = (dDL.im * uL.im) + (dDL.re * uL.re); d3 = (dDL.re * uL.im) - (dDL.im * uL.re); }
      }
      if (CT == 2) {   // the stored form (see unpack2_row): 24 floats, the max-channel index in the sign of channel 1's energy
        f4 *pa = (f4 *)ra, *pb = (f4 *)(sm + (size_t)i * NR);
        f4 v;
        v.x = u0; v.y = u1; v.z = u2; v.w = u3; pa[0] = v;
        v.x = d0; v.y = d1; v.z = d2; v.w = d3; pa[1] = v;
        v.x = enC[0]; v.y = inRe[0]; v.z = inIm[0]; v.w = twRe[0]; pa[2] = v;
        v.x = twIm[0]; v.y = (mc == 1) ? -enC[1] : enC[1]; v.z = inRe[1]; v.w = inIm[1]; pa[3] = v;
        f2 w; w.x = twRe[1]; w.y = twIm[1]; *(f2 *)(ra + 16) = w;
        w.x = s5[0][0]; w.y = s5[0][1]; *(f2 *)(sm + (size_t)i * NR + 18) = w;
        v.x = s5[0][2]; v.y = s5[1][0]; v.z = s5[1][1]; v.w = s5[1][2]; pb[5] = v;
      } else {
        ra[0] = u0; ra[1] = u1; ra[2] = u2; ra[3] = u3; ra[4] = d0; ra[5] = d1; ra[6] = d2; ra[7] = d3;
      }
    }
    BS_SYNC();
    // local rows [0, nOut) are complete: global rows k0 .. k0+nOut-1 (the last tile also flushes the R0 trailing rows)
    const bool lastTile = k0 + TB >= B;
    const int nOut = lastTile ? nb + R0 : TB;
    // whole sectors / lines, one 16-byte piece per thread.  The threads walk the staging area piece by piece INCLUDING the
    // padding piece(s) of every row (which they skip): consecutive threads then read consecutive pieces -- conflict-free --
    // where a walk over the stored pieces only (6 of every 7) had two threads of every eight on one bank group.
    {
      const int P = NR / 4, dr = nt / P, df = nt - dr * P;   // (row, piece) of a thread's next element: no division per piece
      int r = tid / P, f = tid - r * P;
      for (int i = tid; i < nOut * P; i += nt) {
        if (f < NRP / 4) {
          const f4 v = ((const f4 *)sm)[i];   // staged in stored form: a plain copy
          ((f4 *)(recRows + (size_t)(k0 + r) * rowStride))[f] = v;
        }
        r += dr; f += df;
        if (f >= P) { f -= P; ++r; }
      }
    }
    BS_SYNC();
    const int SC = (CT == 2) ? SO - 1 : SO;   // floats of a staged row that belong to the chain part
    if (!lastTile)   // chain parts of this tile's last R0 bins belong to the first R0 rows of the next tile
      for (int i = tid; i < R0 * SC; i += nt) { const int r = i / SC, f = i - r * SC; sm[(size_t)r * NR + f] = sm[(size_t)(TB + r) * NR + f]; }
    BS_SYNC();
  }
}

// ---- chain stage arithmetic (shared by the CUDA wavefront kernel and the serial test emulation)
// S1 rotate + S5: Band.output of the previous block at one bin -> the preliminary prediction of this block
BS_HD cf s5_bin(cf o, bool isNew, cf r, float tRe, float tIm, float div) {
  if (isNew) { cf n; n.im = (o.im * r.re) + (o.re * r.im); n.re = (o.re * r.re) - (o.im * r.im); o = n; }
  cf n;
  n.im = div_pos((tIm * o.re) + (tRe * o.im), div);
  n.re = div_pos((tRe * o.re) - (tIm * o.im), div);
  return n;
}
// S6 at bin k.  ra: the bin's recA record.  oPrev/oLong: this block's new output of channel mc at k-1 / k-longStep;
// n1/nL: this block's S5 prediction of channel mc at k+1 / k+longStep.
template <int C>
BS_HD void chain_bin(const float *ra, int mc, int k, int B, int ls, cf oPrev, cf oLong, cf n1, cf nL, cf *out) {
  float phRe = 0.f, phIm = 0.f;
  if (k > 0) {
    phIm = (ra[1] * oPrev.re) + (ra[0] * oPrev.im); phRe = (ra[0] * oPrev.re) - (ra[1] * oPrev.im);
    if (k >= ls) {
      phIm = ((ra[2] * oLong.im) + phIm) + (ra[3] * oLong.re);
      phRe = ((ra[2] * oLong.re) + phRe) - (oLong.im * ra[3]);
    }
  }
  if (k < B - 1) {
    const float t4 = ra[4] * n1.re, t5 = ra[5] * n1.im, t8 = (ra[4] * n1.im) - (ra[5] * n1.re);
    phIm = t8 + phIm; phRe = (t4 + phRe) + t5;
    if (k < B - ls) {
      const float t6 = ra[6] * nL.re, t7 = ra[7] * nL.im, t9 = ra[6] * nL.im, t10 = nL.re * ra[7];
      phIm = (t9 + phIm) - t10; phRe = (t6 + phRe) + t7;
    }
  }
  float eMc = 0.f; cf fbMc = {0.f, 0.f};
#pragma unroll
  for (int c = 0; c < C; ++c) if (c == mc) { eMc = ra[9 + 5 * c]; fbMc.re = ra[9 + 5 * c + 1]; fbMc.im = ra[9 + 5 * c + 2]; }
  cf om;
  make_output(eMc, fbMc, phRe, phIm, om.re, om.im);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    cf o = om;
    if (c != mc) {
      const float tRe = ra[9 + 5 * c + 3], tIm = ra[9 + 5 * c + 4];
      const float qIm = (tIm * om.re) + (tRe * om.im), qRe = (tRe * om.re) - (tIm * om.im);
      cf fb; fb.re = ra[9 + 5 * c + 1]; fb.im = ra[9 + 5 * c + 2];
      make_output(ra[9 + 5 * c], fb, qRe, qIm, o.re, o.im);
    }
    out[c] = o;
  }
}

// input spectrum of block m: the "current" analysis of the most recent block that had a new spectrum
BS_HD const cf *block_input(const DevGeom &g, const BlockRec2 &r2, int s, long long slot0, int nSlots, const cf *specIn,
                                                 const cf *lastInput) {
  const size_t CB = (size_t)g.C * guard_pitch(g.B);   // guarded arrays: the pointer returned is channel 0's element 0
  if (r2.lastNew >= slot0) return specIn + (((size_t)s * nSlots + (r2.lastNew - slot0)) * 2 + 0) * CB + kGuard;
  return lastInput + (size_t)s * CB + kGuard;
}

#ifdef BS_HOSTEMU
// serial emulation of the chain stage for the blocks [0, nValid) of one stream's chunk: same per-bin functions, one
// block after the other (the wavefront order of the CUDA kernel computes exactly the same values)
template <int C>
inline void chain_host(const DevGeom &g, const DevTables &T, const BlockRec *blocks /* of this stream, at slot0 */, int nValid,
                       const float *rec /* the stream's chunk records */, cf *specOut /* [slot][C][B] */, cf *stateOut /* [C][B] */,
                       const BlockRec2 *blocks2 = nullptr) {
  const int B = g.B, ls = g.longStep, R0 = ls + 1, SO = 9 + 5 * C;
  const size_t NR = rec_row_stride(C);
  std::vector<cf> o5((size_t)C * B);
  for (int t = 0; t < nValid; ++t) {
    const bool isNew = blocks[t].flags & kNew;
    const float *rr = rec + rec_slot_offset(t, B, ls, C);
    cf *so = specOut + (size_t)t * C * B;
    float lg[(9 + 8 * C + 3) & ~3];
    auto load_row = [&](size_t row) -> const float * {   // logical view of a stored row
      const float *p = rr + row * NR;
      if (C == 2) { unpack2_row(p, lg); return lg; }
      return p;
    };
    for (int q = 1; q < B; ++q) {
      const float *row = load_row((size_t)q);
      for (int c = 0; c < C; ++c) {
        const float *b = row + SO + 3 * c;
        o5[(size_t)c * B + q] = s5_bin(stateOut[(size_t)c * B + q], isNew, T.specRot[q], b[0], b[1], b[2]);
      }
    }
    for (int k = 0; k < B; ++k) {
      const float *r = load_row((size_t)(k + R0));
      const int mc = __float_as_int_hd(r[8]);
      cf z = {0.f, 0.f}, out[C];
      chain_bin<C>(r, mc, k, B, ls, k > 0 ? so[(size_t)mc * B + k - 1] : z, k >= ls ? so[(size_t)mc * B + k - ls] : z,
                   k < B - 1 ? o5[(size_t)mc * B + k + 1] : z, k < B - ls ? o5[(size_t)mc * B + k + ls] : z, out);
      if (g.incremental && blocks2 && k < (int)blocks2[t].zeroBelow) for (int c = 0; c < C; ++c) out[c].re = out[c].im = 0.f;
      for (int c = 0; c < C; ++c) so[(size_t)c * B + k] = out[c];
    }
    for (size_t i = 0; i < (size_t)C * B; ++i) stateOut[i] = so[i];
  }
}
#endif

}  // namespace bs
