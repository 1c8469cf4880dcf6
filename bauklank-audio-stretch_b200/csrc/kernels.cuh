// Device code of the stretch engine: three kernels per time-chunk of blocks
//
//   analysis_kernel   (stream, block, {cur,prev}, channel)  window -> half-bin-shifted real FFT -> spectrum in HBM
//   spectral_kernel   one persistent CTA per stream, walks the chunk's blocks in order; phase state
//                     (Band.output, Prediction.energy, RNG, formant estimates) stays resident per stream
//   synthesis_kernel  one persistent CTA per (stream, channel): inverse FFT -> window -> overlap-add ring in shared
//                     memory -> normalised output samples
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

#if defined(BS_PHASE_TIMING) && !defined(BS_HOSTEMU)
__device__ unsigned long long g_phase_cycles[16];
#define BS_MARK(i) do { if (tid == 0 && blockIdx.x == 0) { unsigned long long t_ = clock64(); atomicAdd(&g_phase_cycles[i], t_ - t_mark_); t_mark_ = t_; } } while (0)
#define BS_MARK_INIT() unsigned long long t_mark_ = clock64()
#else
#define BS_MARK(i) ((void)0)
#define BS_MARK_INIT() ((void)0)
#endif

namespace bs {

struct DevGeom {
  int C, L, H, N, B, M, inner, outer, split, longStep, off;  // off = L>>1 (analysis/synthesis offset)
  int wpStartLen;
  int incremental;  // blocks arrive one at a time (compat shim): always carry the input spectrum forward
};
enum : int { kSynthEmit = 1, kSynthAdd = 2 };
struct DevTables {
  const float *win; const cf *tw; const float *otr, *oti; const cf *untangle, *rot, *specRot;
  const float *wpStart, *wpSteady;
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
};
// persistent per-stream state + scratch (all device pointers, stream-major)
struct StateDev {
  cf *outSpec;        // [S][C][B]   Band.output
  float *predE;       // [S][C][B]   Prediction.energy
  cf *lastInput;      // [S][C][B]   last analysed spectrum (only used by blocks without a new spectrum)
  uint32_t *rng;      // [S]
  float *freqEst;     // [S][2]      freqEstimateWeighted, freqEstimateWeight
  float *ring;        // [S][C][L]   overlap-add ring between chunks
  // per chunk slot, written by the premap kernel
  float *inEnergy;    // [S][T][C][B]
  float *map;         // [S][T][B][2]   {inputBin, freqGrad}
  // scratch
  cf *predIn;         // [S][C][B]
  float *terms;       // [S][B][NT]
};

BS_HD int trunc_i32(float x) { return fabsf(x) < 2147483648.0f ? (int)x : INT32_MIN; }
struct alignas(16) f4 { float x, y, z, w; };
BS_HHD int nterms(int C) { return (16 + 3 * C + 3) & ~3; }

// ------------------------------------------------------------------------------------------------------------
// radix-4 DIT passes on split arrays (W#21/W#34 forward, W#20/W#33 inverse), `outer` sub-transforms of length
// `inner` side by side.  Data starts in (ar, ai); returns 0 if the result is in (ar, ai), 1 if in (br, bi).
template <bool INV>
BS_HD int pow2_ffts(const DevGeom &g, const cf *tw, float *ar, float *ai, float *br, float *bi, int tid, int nt) {
  const int inner = g.inner, outer = g.outer;
  if (inner <= 1) return 0;
  int lg = 0; while ((1 << lg) < inner) ++lg;
  float *sr = ar, *si = ai, *dr = br, *di = bi;
  int which = 0, size = 1;
  if (lg & 1) {
    const int stride = inner >> 1, total = outer * stride;
    for (int idx = tid; idx < total; idx += nt) {
      int sub = idx / stride, s = idx - sub * stride, p0 = sub * inner + s, p1 = p0 + stride;
      float a_i = si[p0], b_i = si[p1], b_r = sr[p1], a_r = sr[p0];
      dr[p0] = b_r + a_r; di[p0] = b_i + a_i; dr[p1] = a_r - b_r; di[p1] = a_i - b_i;
    }
    BS_SYNC();
    float *t; t = sr; sr = dr; dr = t; t = si; si = di; di = t; which ^= 1; size = 2;
  }
  while (size < inner) {
    size <<= 2;
    const int stride = inner / size, q = size >> 2, step = inner / size, per = inner >> 2, total = outer * per;
    for (int idx = tid; idx < total; idx += nt) {
      int sub = idx / per, r = idx - sub * per, i = r / stride, s = r - i * stride, base = sub * inner;
      cf tB = tw[i * step], tC = tw[2 * i * step], tD = tw[3 * i * step];
      int pa = base + (4 * i) * stride + s;
      float Ar = sr[pa], Ai = si[pa], Br = sr[pa + stride], Bi = si[pa + stride];
      float Cr = sr[pa + 2 * stride], Ci = si[pa + 2 * stride], Dr = sr[pa + 3 * stride], Di = si[pa + 3 * stride];
      float dRe, bRe, cRe, dIm, bIm, cIm;
      if (!INV) {
        dRe = (Dr * tD.re) - (Di * tD.im); bRe = (Br * tB.re) - (Bi * tB.im); cRe = (Cr * tC.re) - (Ci * tC.im);
        dIm = (Di * tD.re) + (Dr * tD.im); bIm = (Bi * tB.re) + (Br * tB.im); cIm = (Ci * tC.re) + (Cr * tC.im);
      } else {
        dRe = (Di * tD.im) + (Dr * tD.re); bRe = (Bi * tB.im) + (Br * tB.re); cRe = (Ci * tC.im) + (Cr * tC.re);
        dIm = (Di * tD.re) - (Dr * tD.im); bIm = (Bi * tB.re) - (Br * tB.im); cIm = (Ci * tC.re) - (Cr * tC.im);
      }
      float bdRe = dRe + bRe, acRe = cRe + Ar, bdIm = dIm + bIm, acIm = Ai + cIm;
      float x = INV ? (dIm - bIm) : (bIm - dIm), y = Ar - cRe;
      float z = INV ? (bRe - dRe) : (dRe - bRe), w = Ai - cIm;
      int po = base + i * stride + s, qs = q * stride;
      dr[po] = bdRe + acRe;          di[po] = bdIm + acIm;
      dr[po + qs] = x + y;           di[po + qs] = z + w;
      dr[po + 2 * qs] = acRe - bdRe; di[po + 2 * qs] = acIm - bdIm;
      dr[po + 3 * qs] = y - x;       di[po + 3 * qs] = w - z;
    }
    BS_SYNC();
    float *t; t = sr; sr = dr; dr = t; t = si; si = di; di = t; which ^= 1;
  }
  return which;
}

// outer twiddles + final DFT-3 / DFT-5 across the sub-transforms, in place (plan steps 8 and 10/12; W#35 4404-4625
// forward, W#48 10408-10628 inverse)
template <bool INV>
BS_HD void outer_stage(const DevGeom &g, const DevTables &T, float *dr, float *di, int tid, int nt) {
  const int inner = g.inner, outer = g.outer;
  if (outer < 2) return;
  for (int i = tid; i < inner; i += nt) {
    float xr[5], xi[5];
    xr[0] = dr[i]; xi[0] = di[i];
    for (int s = 1; s < outer; ++s) {
      float vr = dr[i + s * inner], vi = di[i + s * inner];
      float wr = T.otr[i + inner * (s - 1)], wi = T.oti[i + inner * (s - 1)];
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
  BS_SYNC();
}

// position of packed sample j after the interleave step of the split FFT (plan types 1-5): j = i*outer + s -> s*inner + i
BS_HD int deint(const DevGeom &g, int j) { return g.outer < 2 ? j : (j % g.outer) * g.inner + j / g.outer; }

// ------------------------------------------------------------------------------------------------------------
// analysis of one window of one channel (W#35): window, zero-phase rotate, zero-pad, modified real FFT.
// smem: 4*M floats.  `x` = channel base of the clip.
BS_HD void analyse_window(const DevGeom &g, const DevTables &T, const float *x, Window w, cf *X, float *sm, int tid, int nt) {
  const int M = g.M, N = g.N, L = g.L, off = g.off;
  float *ar = sm, *ai = sm + M, *br = sm + 2 * M, *bi = sm + 3 * M;
  for (int j = tid; j < M; j += nt) {
    float t[2];
    for (int e = 0; e < 2; ++e) {
      int n = 2 * j + e, i; float sgn;
      if (n < L - off) { i = n + off; sgn = 1.f; }
      else if (n >= N - off) { i = n - (N - off); sgn = -1.f; }
      else { t[e] = 0.f; continue; }
      float xv = (i >= w.lo && i < w.hi) ? x[w.start + i] : 0.f;
      float wv = T.win[i];
      t[e] = xv * (sgn < 0.f ? -wv : wv);
    }
    cf r = T.rot[j];
    int d = deint(g, j);
    ar[d] = (r.re * t[0]) - (r.im * t[1]); ai[d] = (r.im * t[0]) + (r.re * t[1]);
  }
  BS_SYNC();
  int which = pow2_ffts<false>(g, T.tw, ar, ai, br, bi, tid, nt);
  float *rr = which ? br : ar, *ri = which ? bi : ai;
  outer_stage<false>(g, T, rr, ri, tid, nt);
  const int half = M >> 1;
  for (int i = tid; i <= half; i += nt) {
    if (i == half - 1 && !(M & 1)) continue;  // iteration i = M/2 rewrites this pair last in the reference loop
    int j = M - 1 - i; cf u = T.untangle[i];
    float sI = (ri[j] + ri[i]) * 0.5f, dR = (rr[i] - rr[j]) * 0.5f;
    float p = (sI * u.re) + (dR * u.im), dI = (ri[i] - ri[j]) * 0.5f;
    float q = (dR * u.re) - (sI * u.im), sR = (rr[j] + rr[i]) * 0.5f;
    cf xi_, xj_;
    xi_.im = p + dI; xi_.re = q + sR; xj_.im = p - dI; xj_.re = sR - q;
    X[i] = xi_; X[j] = xj_;
  }
  BS_SYNC();
}

// inverse modified real FFT of one channel's output spectrum + windowed overlap-add into `ring` (smem, [L]) at `pos`
// (W#48 9986-10932).  smem: 4*M floats.
BS_HD void synth_frame(const DevGeom &g, const DevTables &T, const cf *X, float *ring, int pos, float *sm, int tid, int nt) {
  const int M = g.M, N = g.N, L = g.L, off = g.off;
  float *ar = sm, *ai = sm + M, *br = sm + 2 * M, *bi = sm + 3 * M;
  const int half = M >> 1;
  for (int i = tid; i <= half; i += nt) {
    if (i == half - 1 && !(M & 1)) continue;
    int j = M - 1 - i; cf u = T.untangle[i];
    cf xi_ = X[i], xj_ = X[j];
    float sI = xj_.im + xi_.im, dR = xi_.re - xj_.re;
    float p = (sI * u.im) + (dR * u.re), sR = xj_.re + xi_.re;
    float q = (sI * u.re) - (dR * u.im), dI = xi_.im - xj_.im;
    int di_ = deint(g, i), dj = deint(g, j);
    ar[di_] = p + sR; ai[di_] = q + dI; ar[dj] = sR - p; ai[dj] = q - dI;
  }
  BS_SYNC();
  int which = pow2_ffts<true>(g, T.tw, ar, ai, br, bi, tid, nt);
  float *rr = which ? br : ar, *ri = which ? bi : ai;
  outer_stage<true>(g, T, rr, ri, tid, nt);
  for (int j = tid; j < M; j += nt) {
    cf r = T.rot[j];
    float t1 = (r.re * ri[j]) - (r.im * rr[j]);
    float t0 = (r.im * ri[j]) + (r.re * rr[j]);
    float tv[2] = {t0, t1};
    for (int e = 0; e < 2; ++e) {
      int n = 2 * j + e;
      if (n < L - off) { int i = n + off; int p = pos + i; if (p >= L) p -= L; ring[p] = ring[p] + (tv[e] * T.win[i]); }
      else if (n >= N - off) { int i = n - (N - off); int p = pos + i; if (p >= L) p -= L; ring[p] = ring[p] - (tv[e] * T.win[i]); }
    }
  }
  BS_SYNC();
}

BS_HD float wp_at(const DevGeom &g, const DevTables &T, long long n) {
  return n < g.wpStartLen ? T.wpStart[n] : T.wpSteady[(int)((n - g.wpStartLen) % g.H)];
}

// one chunk of one (stream, channel): ring persists in global memory between chunks
BS_HD void synth_stream(const DevGeom &g, const DevTables &T, const StreamDev &sd, int c, long long slot0, int nSlots, int mode,
                        const cf *specOut /* this stream's [slot][C][B] */, float *ringG, float *sm, float *ring, int tid, int nt) {
  const int L = g.L, H = g.H;
  for (int i = tid; i < L; i += nt) ring[i] = ringG[i];
  BS_SYNC();
  for (int t = 0; t < nSlots; ++t) {
    long long m = slot0 + t;
    if (m >= sd.nBlocks) break;
    int pos = (int)((m * H) % L);
    const cf *X = specOut + ((size_t)t * g.C + c) * g.B;
    if (!g.split && (mode & kSynthAdd)) synth_frame(g, T, X, ring, pos, sm, tid, nt);
    if (mode & kSynthEmit) {
      long long n0 = m * H;
      float *outc = sd.out + (size_t)c * sd.outStride - sd.outBase;
      for (int j = tid; j < H; j += nt) {
        long long n = n0 + j;
        int p = pos + j; if (p >= L) p -= L;
        if (n < sd.nOut) outc[n] = ring[p] / wp_at(g, T, n);
        ring[p] = 0.f;
      }
      BS_SYNC();
    }
    if (g.split && (mode & kSynthAdd)) { int p2 = pos + H; if (p2 >= L) p2 -= L; synth_frame(g, T, X, ring, p2, sm, tid, nt); }
  }
  for (int i = tid; i < L; i += nt) ringG[i] = ring[i];
}

// ------------------------------------------------------------------------------------------------------------
// spectral stage helpers
BS_HD float map_freq(float f, float mult, float limit) {
  if (!(f <= limit)) return ((mult + -1.0f) * limit) + f;
  return mult * f;
}
BS_HD float smooth_pass(float *v, int n, float slew, float s) {  // backward then forward one-pole pass
  for (int i = n - 1; i >= 0; --i) { s = ((v[i] - s) * slew) + s; v[i] = s; }
  for (int i = 0; i < n; ++i) { s = ((v[i] - s) * slew) + s; v[i] = s; }
  return s;
}
BS_HD cf lerp_c(const cf *a, int B, int low, float fr) {
  cf lo = {0.f, 0.f}, hi = {0.f, 0.f}, r;
  if (low >= 0 && low < B) lo = a[low];
  if (low + 1 >= 0 && low + 1 < B) hi = a[low + 1];
  r.re = ((hi.re - lo.re) * fr) + lo.re; r.im = ((hi.im - lo.im) * fr) + lo.im;
  return r;
}
BS_HD cf rot_prev(cf v, cf r) { cf o; o.re = (v.re * r.re) - (v.im * r.im); o.im = (v.im * r.re) + (v.re * r.im); return o; }
BS_HD cf lerp_prev(const cf *a, const cf *rot, int B, int low, float fr) {  // rot == nullptr: no rotation
  cf lo = {0.f, 0.f}, hi = {0.f, 0.f}, r;
  if (low >= 0 && low < B) lo = rot ? rot_prev(a[low], rot[low]) : a[low];
  if (low + 1 >= 0 && low + 1 < B) hi = rot ? rot_prev(a[low + 1], rot[low + 1]) : a[low + 1];
  r.re = ((hi.re - lo.re) * fr) + lo.re; r.im = ((hi.im - lo.im) * fr) + lo.im;
  return r;
}
BS_HD float lerp_f(const float *a, int B, int low, float fr) {
  float lo = (low >= 0 && low < B) ? a[low] : 0.f, hi = (low + 1 >= 0 && low + 1 < B) ? a[low + 1] : 0.f;
  return ((hi - lo) * fr) + lo;
}
// minstd_rand: state after n steps from x (x_{k+1} = 48271 x_k mod 2^31-1), by square-and-multiply
BS_HD uint32_t minstd_jump(uint32_t x, uint32_t n) {
  unsigned long long a = 48271ull, r = 1ull; const unsigned long long m = 2147483647ull;
  while (n) { if (n & 1u) r = (r * a) % m; a = (a * a) % m; n >>= 1; }
  return (uint32_t)((r * (unsigned long long)x) % m);
}
BS_HD void make_output(float energy, cf fallback, float re, float im, float &ore, float &oim) {
  float n2 = (im * im) + (re * re), div;
  if (n2 > 1e-15f) div = n2;
  else { re = fallback.re; im = fallback.im; div = ((re * re) + 1e-15f) + (im * im); }
  float s = sqrtf(energy / div);
  oim = s * im; ore = s * re;
}

// State-independent part of the spectral stage of one block: input energies, energy smoothing, peak picking,
// the output frequency map and the formant envelope (W#48 8238-9312).  It depends only on the block's input spectrum
// and parameters (except the formant auto-detect, which carries freqEst across blocks), so the premap kernel runs it
// for all blocks of a chunk in parallel.  Sequential recurrences run on tid 0.
// smem (floats): energy[B+2] | smoothed[B] | cpk[B/2+2] ints | peaks[B] | misc[16]
BS_HHD size_t map_smem_floats(int B) { return (((size_t)(B + 2) + B + (B / 2 + 2) + B + 16) + 3) & ~(size_t)3; }
template <int CT>
BS_HD void map_stage(const DevGeom &g, const DevTables &T, const BlockRec rec, const BlockRec2 rec2, const cf *inp,
                     float *freqEst, float *inEnergy, float *mapv, float *sm, int tid, int nt) {
  const int C = CT > 0 ? CT : g.C, B = g.B;
  const bool mapped = rec.flags & kMapped, formants = rec.flags & kFormants;
  const float fN = (float)(uint32_t)g.N, fH = (float)(uint32_t)g.H, ratio = fN / fH;
  float *energy = sm, *smoothed = sm + (B + 2);
  int *cpk = (int *)(smoothed + B);
  float *peaksG = (float *)(cpk + (B / 2 + 2));
  int *misc = (int *)(peaksG + B);   // [0] nPeaks, [1] monotone flag, [2] formant base bin
  BS_MARK_INIT();
  for (int idx = tid; idx < C * B; idx += nt) {
    cf v = inp[idx];
    inEnergy[idx] = (v.im * v.im) + (v.re * v.re);
  }
  BS_SYNC();
  BS_MARK(0);
  if (mapped) {
    for (int k = tid; k < B; k += nt) {
      float e = 0.f;
      for (int c = 0; c < C; ++c) e = e + inEnergy[(size_t)c * B + k];
      energy[k] = e; smoothed[k] = e;
    }
    BS_SYNC();
    if (tid == 0) {
      // smoothEnergy steps 1,2 (one-pole, carry kept across both) then findPeaks
      float slew = 1.0f / ((ratio * 0.5f) + 1.0f), carry = 0.f;
      BS_MARK(1);
      carry = smooth_pass(smoothed, B, slew, carry);
      carry = smooth_pass(smoothed, B, slew, carry);
      BS_MARK(2);
      int nP = 0, k = 0, mono = 1, prevC = INT32_MIN;
      while (k < B) {
        if (!(energy[k] <= smoothed[k])) {
          float sum = 0.f, wsum = 0.f;
          while (k < B) {
            float en = energy[k];
            if (en <= smoothed[k]) break;
            sum = en + sum; wsum = (en * (float)k) + wsum; ++k;
          }
          float avg = wsum / sum;
          float f = (avg + 0.5f) / fN;
          float o = (map_freq(f, rec.pkMult, rec.pkLimit) * fN) + -0.5f;
          peaksG[2 * nP] = avg; peaksG[2 * nP + 1] = o;
          int cc = trunc_i32(ceilf(o));
          if (cc < prevC) mono = 0;
          prevC = cc; cpk[nP] = cc;
          ++nP;
        }
        ++k;
      }
      misc[0] = nP; misc[1] = mono;
      BS_MARK(3);
    }
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
        else if (mono) {
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
  } else {
    for (int k = tid; k < B; k += nt) { mapv[2 * k] = (float)(uint32_t)k; mapv[2 * k + 1] = 1.0f; }
  }
  BS_SYNC();
  BS_MARK(4);
  if (formants) {
    float *fm = energy;  // [B+2]
    for (int k = tid; k < B + 2; k += nt) {
      float e = 0.f;
      if (k < B) for (int c = 0; c < C; ++c) e = e + inEnergy[(size_t)c * B + k];
      fm[k] = e;
    }
    BS_SYNC();
    if (tid == 0) {
      float base = rec.fmBaseFreq, baseBin = (base * fN) + -0.5f;
      if (!(base > 0.f)) {
        int i1 = 0, i2 = 0, i3 = 0;
        for (int i = 1; i <= B - 2; ++i) {
          float v = fm[i];
          if (v < fm[i - 1]) continue;
          if (v <= fm[i + 1]) continue;
          if (v <= fm[i3]) continue;
          if (fm[i2] >= v) { i3 = i; continue; }
          if (fm[i1] < v) { i3 = i2; i2 = i1; i1 = i; continue; }
          i3 = i2; i2 = i;
        }
        float top = fm[i1]; double dtop = (double)top;
        if ((double)fm[i2] > (dtop * 0.1)) {
          int d = i1 - i2; if (d < 0) d = -d;
          if (d > i1 / 8 && d < (i1 * 7) / 8) i1 = i1 % d;
          if (!((double)fm[i3] <= (dtop * 0.01))) {
            d = i1 - i3; if (d < 0) d = -d;
            if (d > i1 / 8 && d < (i1 * 7) / 8) i1 = i1 % d;
          }
        }
        float w = freqEst[1];
        float nw = (float)(((double)(top - w) * 0.25) + (double)w);
        freqEst[1] = nw;
        float ww = freqEst[0];
        ww = (float)(((double)((top * (float)i1) - ww) * 0.25) + (double)ww);
        freqEst[0] = ww;
        baseBin = ww / (nw + 1e-30f);
      }
      misc[2] = __float_as_int_hd(baseBin);
    }
    BS_SYNC();
    for (int k = tid; k < B; k += nt) fm[k] = sqrtf(fm[k]);
    BS_SYNC();
    if (tid == 0) {
      float baseBin = __int_as_float_hd(misc[2]);
      float slew = (float)(1.0 / (((double)baseBin * 0.5) + 1.0));
      float st = smooth_pass(fm, B, slew, 0.f);
      smooth_pass(fm, B, slew, st);
    }
    BS_SYNC();
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
      for (int c = 0; c < C; ++c) { size_t o = (size_t)c * B + k; inEnergy[o] = g2 * inEnergy[o]; }
    }
    BS_SYNC();
  }
}

// smem of the spectral kernel (floats): [map_stage area] | term tiles [2][32][NT] | chain rings
BS_HHD size_t spectral_smem_head(int B) { return map_smem_floats(B); }
BS_HHD size_t spectral_smem_floats(int B, int C) { return spectral_smem_head(B) + 64 * (size_t)nterms(C) + 192 + 128 * (size_t)C + 16; }

// State-dependent part of one block of one stream (W#48 8193-8229 rotate, 9314-9455 preliminary prediction,
// 9458-9873 vertical prediction chain).  `doMap`: the block's map stage was not precomputed (formant auto-detect).
template <int CT>
BS_HD void spectral_block(const DevGeom &g, const DevTables &T, const BlockRec rec, const BlockRec2 rec2,
                          const cf *inp, const cf *inPrev /* [C][B] each; inPrev == nullptr: no new spectrum */, cf *outSpec,
                          float *predE, uint32_t *rngp, float *freqEst, float *inEnergy, float *mapv, cf *predIn,
                          float *terms, cf *specOut, bool doMap, float *sm, int tid, int nt) {
  const int C = CT > 0 ? CT : g.C, B = g.B, NT = nterms(C);
  const bool isNew = rec.flags & kNew;
  float *tileT = sm + spectral_smem_head(B);            // [2][32][NT] two 32-bin term tiles (16-byte aligned)
  float *tileO = tileT + 64 * (size_t)NT;               // chain output rings: [64] cf, [64] int, [64][C] cf
  const cf *prv = isNew ? inPrev : inp;
  const cf *prvRot = isNew ? T.specRot : nullptr;
  BS_MARK_INIT();
  // S1 rotate Band.output (prevInput is rotated on the fly when read)
  if (isNew) {
    for (int idx = tid; idx < C * B; idx += nt) {
      int k = idx % B;
      cf o = outSpec[idx], r = T.specRot[k], n;
      n.im = (o.im * r.re) + (o.re * r.im); n.re = (o.re * r.re) - (o.im * r.im);
      outSpec[idx] = n;
    }
  }
  if (doMap) map_stage<CT>(g, T, rec, rec2, inp, freqEst, inEnergy, mapv, sm, tid, nt);
  BS_SYNC();
  BS_MARK(5);
  // S5 preliminary prediction, all (channel, bin) in parallel
  for (int idx = tid; idx < C * B; idx += nt) {
    int c = idx / B, k = idx - c * B;
    const cf *ic = inp + (size_t)c * B, *pc = prv + (size_t)c * B;
    float ib = mapv[2 * k], fl = floorf(ib);
    int low = trunc_i32(fl); float fr = ib - fl;
    float prevE = predE[idx], grad = mapv[2 * k + 1];
    float en = lerp_f(inEnergy + (size_t)c * B, B, low, fr) * (grad > 0.f ? grad : 0.f);
    predE[idx] = en;
    cf in = lerp_c(ic, B, low, fr); predIn[idx] = in;
    cf pv = lerp_prev(pc, prvRot, B, low, fr);
    float tIm = (pv.re * in.im) - (pv.im * in.re), tRe = (pv.im * in.im) + (pv.re * in.re);
    cf o = outSpec[idx];
    float div = ((en > prevE) ? en : prevE) + 1e-15f;
    cf n;
    n.im = ((tIm * o.re) + (tRe * o.im)) / div;
    n.re = ((tRe * o.re) - (tIm * o.im)) / div;
    outSpec[idx] = n;
  }
  BS_SYNC();
  BS_MARK(6);
  // S6 part 1: everything that does not depend on the chain, per bin, in parallel -> terms[k][NT]
  const int longStep = g.longStep;
  const float tf = rec.timeFactor < 0.5f ? 0.5f : rec.timeFactor;
  const float rlo = ((tf > 2.0f) ? 4.0f : 0.0f) - tf, rscale = (tf - rlo) * 0x1p-31f, fLong = (float)longStep;
  const bool randomTF = !(tf <= 2.0f);
  const uint32_t rng0 = *rngp;
  for (int k = tid; k < B; k += nt) {
    float *tt = terms + (size_t)k * NT;
    int mc = 0; float me = predE[k];
    for (int c = 1; c < C; ++c) { float en = predE[(size_t)c * B + k]; if (en > me) { me = en; mc = c; } }
    const cf *ic = inp + (size_t)mc * B; const cf *pi = predIn + (size_t)mc * B; const cf *oc = outSpec + (size_t)mc * B;
    float pRe = pi[k].re, pIm = pi[k].im;
    tt[11] = __int_as_float_hd(mc);
    if (k > 0) {
      float ib = mapv[2 * k], btf = tf;
      if (randomTF) { uint32_t x = minstd_jump(rng0, (uint32_t)(2 * k)); btf = (rscale * (float)(uint32_t)(x - 1u)) + rlo; }
      float x = ib - btf; int low = trunc_i32(floorf(x)); float fr = x - (float)low;
      cf d = lerp_c(ic, B, low, fr);
      tt[1] = (d.re * pIm) - (d.im * pRe); tt[0] = (d.im * pIm) + (d.re * pRe);
      if (k >= longStep) {
        x = ib - (btf * fLong); low = trunc_i32(floorf(x)); fr = x - (float)low;
        d = lerp_c(ic, B, low, fr);
        tt[2] = (d.im * pIm) + (d.re * pRe); tt[3] = (d.re * pIm) - (d.im * pRe);
      }
    }
    if (k < B - 1) {
      float btf = tf;
      if (randomTF) { uint32_t x = minstd_jump(rng0, (uint32_t)(2 * k + 1)); btf = (rscale * (float)(uint32_t)(x - 1u)) + rlo; }
      float x = mapv[2 * (k + 1)] - btf; int low = trunc_i32(floorf(x)); float fr = x - (float)low;
      cf d = lerp_c(ic, B, low, fr);
      float uIm = pi[k + 1].im, uRe = pi[k + 1].re;
      float tRe = (d.im * uIm) + (d.re * uRe), tIm = (d.re * uIm) - (d.im * uRe);
      float oIm = oc[k + 1].im, oRe = oc[k + 1].re;
      tt[4] = tRe * oRe; tt[5] = tIm * oIm;                 // phRe = ((tt4 + phRe) + tt5)
      tt[8] = (tRe * oIm) - (tIm * oRe);                    // phIm = tt8 + phIm
      if (k < B - longStep) {
        int kk = k + longStep;
        x = mapv[2 * kk] - (btf * fLong); low = trunc_i32(floorf(x)); fr = x - (float)low;
        d = lerp_c(ic, B, low, fr);
        uIm = pi[kk].im; uRe = pi[kk].re;
        tRe = (d.im * uIm) + (d.re * uRe); tIm = (d.re * uIm) - (d.im * uRe);
        oIm = oc[kk].im; oRe = oc[kk].re;
        tt[6] = tRe * oRe; tt[7] = tIm * oIm;               // phRe = ((tt6 + phRe) + tt7)
        tt[9] = tRe * oIm; tt[10] = oRe * tIm;              // phIm = ((tt9 + phIm) - tt10)
      }
    }
    tt[12] = me; tt[13] = pRe; tt[14] = pIm;                   // energy and fallback input of the max channel
    for (int c = 0; c < C; ++c) {
      tt[16 + c] = predE[(size_t)c * B + k];
      cf cp = predIn[(size_t)c * B + k];
      tt[16 + C + 2 * c] = (pIm * cp.im) + (pRe * cp.re);      // channel twist re
      tt[16 + C + 2 * c + 1] = (pRe * cp.im) - (pIm * cp.re);  // channel twist im
    }
  }
  BS_SYNC();
  BS_MARK(7);
  // S6 part 2: the bin-to-bin chain (first warp only).  32-bin tiles of terms are staged through shared memory.
  // Lane 0 walks the bins computing only the maximum-energy channel (the one the recurrence runs through); the other
  // channels ("followers": out[c] = makeOutput(out[mc] * channelTwist[c])) do not feed the recurrence unless the
  // maximum channel changes, so they are filled in for a whole tile at once by all lanes afterwards, and computed on
  // demand by lane 0 in the rare bins where it needs one early.
#ifdef BS_HOSTEMU
  const int lanes = 1, lane = 0; const bool inChain = true;
#else
  const int lanes = 32, lane = tid & 31; const bool inChain = tid < 32;
#endif
  if (inChain) {
    cf *ringMc = (cf *)tileO;                  // [64] output of the max channel per bin
    int *ringIdx = (int *)(tileO + 128);       // [64] which channel that was
    cf *ringFull = (cf *)(tileO + 192);        // [64][C] all channels (complete for finished tiles)
    const int nTiles = (B + 31) / 32;
    for (int i = lane; i < 32 * NT && i < B * NT; i += lanes) tileT[i] = terms[i];
    BS_WARPSYNC();
    float pr = 0.f, pi_ = 0.f; int mcPrev = -1;
    for (int tIdx = 0; tIdx < nTiles; ++tIdx) {
      const int k0 = tIdx * 32, k1 = (k0 + 32 < B) ? k0 + 32 : B;
      float *cur = tileT + (size_t)(tIdx & 1) * 32 * NT, *nxt = tileT + (size_t)((tIdx + 1) & 1) * 32 * NT;
      if (tIdx + 1 < nTiles) {  // prefetch next tile while lane 0 works
        int nEl = ((k1 + 32 < B) ? 32 : (B - k1)) * NT;
        const float *src = terms + (size_t)k1 * NT;
        for (int i = lane; i < nEl; i += lanes) nxt[i] = src[i];
      }
      // out[c][j] for a channel c that was not the maximum at bin j
      auto follower = [&](int j, int c, float &re, float &im) {
        if (j < k0) { cf v = ringFull[(size_t)(j & 63) * C + c]; re = v.re; im = v.im; return; }
        const float *tj = cur + (size_t)(j - k0) * NT;
        cf om = ringMc[j & 63];
        float tRe = tj[16 + C + 2 * c], tIm = tj[16 + C + 2 * c + 1];
        float qIm = (tIm * om.re) + (tRe * om.im), qRe = (tRe * om.re) - (tIm * om.im);
        float n2 = (qIm * qIm) + (qRe * qRe);
        cf fb = {0.f, 0.f};
        if (!(n2 > 1e-15f)) fb = predIn[(size_t)c * B + j];
        make_output(tj[16 + c], fb, qRe, qIm, re, im);
      };
      if (lane == 0) {
        if (k0 >= longStep && k1 <= B - longStep && longStep >= 2) {
          // interior tile: no edge cases; next bin's terms and long-step history are fetched one bin ahead
          const float *tt = cur;
          f4 ta = *(const f4 *)tt, tb = *(const f4 *)(tt + 4), tc = *(const f4 *)(tt + 8), td = *(const f4 *)(tt + 12);
          cf oL = ringMc[(k0 - longStep) & 63]; int iL = ringIdx[(k0 - longStep) & 63];
          for (int k = k0; k < k1; ++k) {
            const float *tn = tt + NT;
            const f4 na = *(const f4 *)tn, nb = *(const f4 *)(tn + 4), nc = *(const f4 *)(tn + 8), nd = *(const f4 *)(tn + 12);
            const cf noL = ringMc[(k + 1 - longStep) & 63]; const int niL = ringIdx[(k + 1 - longStep) & 63];
            const int mc = __float_as_int_hd(tc.w);
            float oRe = pr, oIm = pi_, lRe = oL.re, lIm = oL.im;
            if ((mc != mcPrev) | (iL != mc)) {
              if (mc != mcPrev) follower(k - 1, mc, oRe, oIm);
              if (iL != mc) follower(k - longStep, mc, lRe, lIm);
            }
            float phIm = (ta.y * oRe) + (ta.x * oIm), phRe = (ta.x * oRe) - (ta.y * oIm);
            phIm = ((ta.z * lIm) + phIm) + (ta.w * lRe);
            phRe = ((ta.z * lRe) + phRe) - (lIm * ta.w);
            phIm = tc.x + phIm; phRe = (tb.x + phRe) + tb.y;
            phIm = (tc.y + phIm) - tc.z; phRe = (tb.z + phRe) + tb.w;
            const float n2 = (phIm * phIm) + (phRe * phRe);
            float div = n2;
            if (!(n2 > 1e-15f)) { phRe = td.y; phIm = td.z; div = ((phRe * phRe) + 1e-15f) + (phIm * phIm); }
            const float sc = sqrtf(td.x / div);
            pi_ = sc * phIm; pr = sc * phRe; mcPrev = mc;
            cf o; o.re = pr; o.im = pi_;
            ringMc[k & 63] = o; ringIdx[k & 63] = mc;
            ta = na; tb = nb; tc = nc; td = nd; oL = noL; iL = niL; tt = tn;
          }
        } else {
          for (int k = k0; k < k1; ++k) {
            const float *tt = cur + (size_t)(k - k0) * NT;
            const f4 ta = *(const f4 *)tt, tb = *(const f4 *)(tt + 4), tc = *(const f4 *)(tt + 8), td = *(const f4 *)(tt + 12);
            const int mc = __float_as_int_hd(tc.w);
            float phRe = 0.f, phIm = 0.f, oRe, oIm;
            if (k > 0) {
              if (mc == mcPrev) { oRe = pr; oIm = pi_; } else follower(k - 1, mc, oRe, oIm);
              phIm = (ta.y * oRe) + (ta.x * oIm); phRe = (ta.x * oRe) - (ta.y * oIm);
              if (k >= longStep) {
                const int j = k - longStep;
                if (ringIdx[j & 63] == mc) { cf v = ringMc[j & 63]; oRe = v.re; oIm = v.im; } else follower(j, mc, oRe, oIm);
                phIm = ((ta.z * oIm) + phIm) + (ta.w * oRe);
                phRe = ((ta.z * oRe) + phRe) - (oIm * ta.w);
              }
            }
            if (k < B - 1) {
              phIm = tc.x + phIm;
              phRe = (tb.x + phRe) + tb.y;
              if (k < B - longStep) {
                phIm = (tc.y + phIm) - tc.z;
                phRe = (tb.z + phRe) + tb.w;
              }
            }
            cf fb; fb.re = td.y; fb.im = td.z;
            make_output(td.x, fb, phRe, phIm, oRe, oIm);
            pr = oRe; pi_ = oIm; mcPrev = mc;
            cf o; o.re = oRe; o.im = oIm;
            ringMc[k & 63] = o; ringIdx[k & 63] = mc;
          }
        }
      }
      BS_WARPSYNC();
      for (int j = lane; j < k1 - k0; j += lanes) {   // followers of this tile + coalesced write-back
        const int kk = k0 + j;
        const int mc = ringIdx[kk & 63];
        const cf om = ringMc[kk & 63];
        for (int c = 0; c < C; ++c) {
          cf o = om;
          if (c != mc) follower(kk, c, o.re, o.im);
          ringFull[(size_t)(kk & 63) * C + c] = o;
          outSpec[(size_t)c * B + kk] = o;
          specOut[(size_t)c * B + kk] = o;
        }
      }
      BS_WARPSYNC();
    }
    if (lane == 0 && randomTF && B >= 2) *rngp = minstd_jump(rng0, (uint32_t)(2 * B - 2));
  }
  BS_SYNC();
  BS_MARK(8);
}

// input spectrum of block m: the "current" analysis of the most recent block that had a new spectrum
BS_HD const cf *block_input(const DevGeom &g, const BlockRec2 &r2, int s, long long slot0, int nSlots, const cf *specIn,
                                                 const cf *lastInput) {
  const size_t CB = (size_t)g.C * g.B;
  if (r2.lastNew >= slot0) return specIn + (((size_t)s * nSlots + (r2.lastNew - slot0)) * 2 + 0) * CB;
  return lastInput + (size_t)s * CB;
}
BS_HD bool needs_inline_map(const BlockRec &rec) { return (rec.flags & kFormants) && !(rec.fmBaseFreq > 0.f); }


}  // namespace bs
