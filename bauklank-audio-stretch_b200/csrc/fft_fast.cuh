// Specialised STFT kernels for the preset geometries: the same arithmetic as analyse_window / synth_frame (kernels.cuh)
// -- the reference's radix-4 decimation-in-time butterflies on the same values in the same order (W#21/W#34 forward,
// W#20/W#33 inverse, outer DFT-3/5 of W#35 4404-4625 and W#48 10408-10628, untangle, window) -- with the data movement
// rebuilt around it.  The run-time-geometry path spends more instructions on indices, predicates and shared-memory
// traffic than on the butterflies (ncu: 19 400 warp instructions per 3072-point analysis, 7 500 of them FADD/FMUL); here
// every stride is a compile-time constant and a transform is a fixed sequence of stages:
//
//   forward   pack (window, zero-phase rotate, half-bin rotation; a leading radix-2 pass if log2(inner) is odd)
//             -> [radix-16 = two fused radix-4 passes in registers] x n -> [radix-4] -> radix-4 + outer twiddles + DFT-3/5
//             in registers -> untangle (pairs k, M-1-k) -> spectrum in HBM
//   inverse   untangle^-1 from HBM (+ radix-2) -> the same passes, inverse butterflies -> last radix-4 + outer DFT
//             + half-bin rotation + synthesis window in registers -> frame in HBM
//
// Complex values travel between stages as interleaved {re, im} pairs (one 64-bit shared-memory access per value), two
// buffers used alternately; every buffer carries the padding that makes its reader conflict-free (its writer always
// stores consecutive elements from consecutive lanes).  Stage bodies are (tid) work loops over NT logical threads so
// that the serial test emulation (BS_HOSTEMU) can run the identical code.
#pragma once
#include "kernels.cuh"
#ifdef BS_HOSTEMU
#include <algorithm>
using std::max;
using std::min;
#endif

namespace bs {

constexpr int kFastNT = 256;

// geometry of a specialised transform
template <int LG, int OUTER>
struct FastGeom {
  static constexpr int inner = 1 << LG, M = inner * OUTER, NT = kFastNT;
  static constexpr int lg0 = LG & 1;                       // lgSize after the radix-2 pass that the pack stage absorbs
  static constexpr int rem = LG - lg0 - 2;                 // bits left for the stages before the final radix-4 (+ outer DFT)
  static constexpr int nR16 = rem / 4;
  static constexpr bool hasR4 = (rem % 4) == 2;
  static constexpr int nStages = nR16 + (hasR4 ? 1 : 0) + 1;   // shared-memory stages after the pack
  static_assert(rem >= 0 && (rem % 2) == 0, "unsupported inner size");
};

// buffer layout: element e of sub-transform `sub` lives at sub * PITCH + e + (e >> SHIFT) * ADD   (ADD == 0: no padding)
// A stage addresses e = e0 + d with a thread-dependent e0 and compile-time offsets d that never carry into each other
// below bit SHIFT (see the stages), so the padding separates: at(sub, e0) once per work item, plus the constant pad(d).
template <int SHIFT, int ADD, int PITCH>
struct Lay {
  static constexpr int pitch = PITCH;
  BS_HHD static int at(int sub, int e) { return sub * PITCH + e + (ADD ? ((e >> SHIFT) * ADD) : 0); }
  BS_HHD static constexpr int pad(int d) { return d + (ADD ? ((d >> SHIFT) * ADD) : 0); }
};
constexpr int lay_padded(int n, int shift, int add) { return n + (add ? ((n >> shift) * add) : 0); }
constexpr int ilog2c(int x) { return x <= 1 ? 0 : 1 + ilog2c(x >> 1); }
// A stage reads runs of G consecutive elements per lane group, the groups P elements apart; with G < 16 the groups of
// a half warp would share banks unless every P elements are followed by G elements of padding.
//   radix-16 at lgSize: G = strideB = inner >> (lgSize + 4), P = 16 G;   radix-4 at lgSize: G = inner >> (lgSize + 2), P = 4 G
//   the final radix-4 (stride 1) reads its four inputs as two 128-bit words: 2 elements of padding per 16
template <int LG, int OUTER, int K> struct StageOf {   // stage K (0-based) after the pack
  using F = FastGeom<LG, OUTER>;
  static constexpr bool isR16 = K < F::nR16;
  static constexpr bool isLast = K == F::nStages - 1;
  static constexpr int lgSize = isR16 ? F::lg0 + 4 * K : (isLast ? LG - 2 : F::lg0 + 4 * F::nR16);
  static constexpr int G = isR16 ? (F::inner >> (lgSize + 4)) : (F::inner >> (lgSize + 2));
  static constexpr int P = isR16 ? 16 * G : 4 * G;
  static constexpr int shift = isLast ? 4 : ilog2c(P);
  static constexpr int add = isLast ? 2 : (G >= 16 ? 0 : G);
  // The pack stage (and untangle^-1) stores element (j % OUTER, j / OUTER) from lane j: 16 consecutive j are conflict-free
  // exactly when the sub-transform pitch is 11 (OUTER 3) or 13 (OUTER 5) modulo 16 elements -- found by enumeration.
  static constexpr int packMod = OUTER == 3 ? 11 : (OUTER == 5 ? 13 : 0);
  static constexpr int need = lay_padded(F::inner, shift, add);
  static constexpr int pitch = (K == 0 && packMod) ? need + ((packMod - need % 16 + 16) % 16) : ((need + 1) & ~1);
  static_assert(!(K == 0 && isLast), "a single-stage transform would need an even pitch here");
  using In = Lay<shift, add, pitch>;
};
template <int LG, int OUTER> struct NaturalLay { using L = Lay<0, 0, (1 << LG)>; };   // spectrum order k = i + s * inner
// shared-memory floats of one CTA: two buffers, each large enough for any stage's input
template <int LG, int OUTER, int K = 0> struct FastSmem {
  using F = FastGeom<LG, OUTER>;
  static constexpr int here = K < F::nStages ? StageOf<LG, OUTER, K>::pitch * OUTER : 0;
  static constexpr int rest = FastSmem<LG, OUTER, K + 1>::value;
  static constexpr int value = here > rest ? here : rest;
};
template <int LG, int OUTER> struct FastSmem<LG, OUTER, 8> { static constexpr int value = (1 << LG) * OUTER; };
template <int LG, int OUTER> struct FastBuf { static constexpr int elems = (FastSmem<LG, OUTER>::value + 1) & ~1; };
// The second radix-4 pass of a radix-16 stage at sub-transform size 2^LGS takes 12 twiddles per work item that depend only on
// the item's position iA in [0, 2^LGS): with 16 or 32 positions, a warp's lanes gather them from 8 or 16 places of the table --
// one sector each, 12 loads per item -- although the whole stage uses 12 * 2^LGS values.  That stage keeps them in shared
// memory ([jA][m][iA]: the positions of a row side by side, conflict-free, lanes of equal iA broadcast).
template <int LG, int OUTER, int K = 0> struct TwCacheStage {   // the radix-16 stage (0-based, after the pack) that has the cache, or -1
  using F = FastGeom<LG, OUTER>;
  static constexpr bool here = K < F::nR16 && (F::lg0 + 4 * K) >= 4;
  static constexpr int value = here ? K : TwCacheStage<LG, OUTER, K + 1>::value;
};
template <int LG, int OUTER> struct TwCacheStage<LG, OUTER, 4> { static constexpr int value = -1; };
template <int LG, int OUTER> struct TwCache {
  static constexpr int stage = TwCacheStage<LG, OUTER>::value;
  static constexpr int lgs = stage < 0 ? 0 : FastGeom<LG, OUTER>::lg0 + 4 * stage, sets = 1 << lgs;
  static constexpr int elems = stage < 0 ? 0 : 12 * sets;          // cf
};
template <int LG, int OUTER> constexpr size_t fast_smem_bytes() { return (2 * (size_t)FastBuf<LG, OUTER>::elems + TwCache<LG, OUTER>::elems) * sizeof(cf); }
// CTAs per SM the kernels are compiled for: three where shared memory allows (80 registers), else what fits of 227 KB
template <int LG, int OUTER> struct FastOcc {
  static constexpr int bySmem = (int)(232448 / ((2 * (size_t)FastBuf<LG, OUTER>::elems + TwCache<LG, OUTER>::elems) * sizeof(cf) + 1024));
  static constexpr int ctas = bySmem >= 4 ? 4 : (bySmem < 1 ? 1 : bySmem);
};

// ---- the outer stage of one bin: twiddles on sub-transforms 1.., then the DFT across them (outer_stage_t, kernels.cuh)
template <bool INV, int OUTER>
BS_HD void outer_point(float *xr, float *xi, const cf *w /* [OUTER-1] */) {
  if (OUTER < 2) return;
#pragma unroll
  for (int s = 1; s < OUTER; ++s) {
    const float vr = xr[s], vi = xi[s], wr = w[s - 1].re, wi = w[s - 1].im;
    if (!INV) { xr[s] = (wr * vr) - (wi * vi); xi[s] = (wi * vr) + (vi * wr); }
    else { xr[s] = (vi * wi) + (vr * wr); xi[s] = (vi * wr) - (wi * vr); }
  }
  if (OUTER == 3) {
    const float h = INV ? 0x1.bb67aep-1f : -0x1.bb67aep-1f;
    const float ar = xr[0], br = xr[1], cr = xr[2], ai = xi[0], bi = xi[1], ci = xi[2];
    xr[0] = (br + ar) + cr; xi[0] = ci + (bi + ai);
    const float p = ar + (br * -0.5f), q = bi * h, r = cr * -0.5f, t = ci * h;
    const float u = ai + (bi * -0.5f), v = br * h, x = cr * h, y = ci * -0.5f;
    xr[1] = ((p - q) + r) + t; xi[1] = ((u + v) - x) + y;
    xr[2] = ((p + q) + r) - t; xi[2] = ((u - v) + x) + y;
  } else if (OUTER == 5) {
    const float c1 = 0x1.3c6ef4p-2f, c2 = 0x1.9e377ap-1f, s1 = 0x1.e6f0e2p-1f, s2 = 0x1.2cf23p-1f;
    const float ar = xr[0], br = xr[1], cr = xr[2], d_r = xr[3], er = xr[4];
    const float ai = xi[0], bi = xi[1], ci = xi[2], d_i = xi[3], ei = xi[4];
    const float dcR = d_r + cr, ebR = er + br, dcI = d_i + ci, ebI = ei + bi;
    xr[0] = (dcR + ar) + ebR; xi[0] = (ai + dcI) + ebI;
    const float p1r = ar + ((ebR * c1) - (dcR * c2)), p1i = ai + ((ebI * c1) - (dcI * c2));
    const float p2r = ar + ((dcR * c1) - (ebR * c2)), p2i = ai + ((dcI * c1) - (ebI * c2));
    float q1r, q1i, q2r, q2i;
    if (!INV) {
      const float a = d_i - ci, b = ei - bi, c = cr - d_r, d = br - er;
      q1r = (a * -s2) - (b * s1); q1i = (c * -s2) - (d * s1);
      q2r = (b * -s2) + (a * s1); q2i = (d * -s2) + (c * s1);
    } else {
      const float a = ei - bi, b = d_i - ci, c = br - er, d = cr - d_r;
      q1r = (a * s1) + (b * s2); q1i = (c * s1) + (d * s2);
      q2r = (a * s2) - (b * s1); q2i = (c * s2) - (d * s1);
    }
    xr[1] = p1r + q1r; xi[1] = p1i + q1i;
    xr[2] = p2r + q2r; xi[2] = p2i + q2i;
    xr[3] = p2r - q2r; xi[3] = p2i - q2i;
    xr[4] = p1r - q1r; xi[4] = p1i - q1i;
  }
}

// ---- radix-16 stage: radix-4 passes at sub-transform sizes 2^LGS and 2^(LGS+2), fused in registers (pow2_ffts_t)
template <int LG, int OUTER, bool INV, int LGS, class LIn, class LOut, bool CACHED = false>
BS_HD void fast_r16(const cf *tw, const cf *src, cf *dst, int tid, const cf *twc = nullptr /* CACHED: TwCache, filled by fast_fill_twcache */) {
  constexpr int lgStrideA = LG - LGS - 2, lgStrideB = lgStrideA - 2, lgPer = LG - 4, nItems = OUTER << lgPer, strideB = 1 << lgStrideB;
  static_assert(lgStrideB >= 0, "radix-16 stage does not fit");
  for (int idx = tid; idx < nItems; idx += kFastNT) {
    const int sub = idx >> lgPer, r = idx & ((1 << lgPer) - 1), iA = r >> lgStrideB, sB = r & (strideB - 1);
    const int base = (iA << (lgStrideA + 2)) + sB;
    float vr[4][4], vi[4][4];   // [a][j]
    const cf *ps = src + LIn::at(sub, base);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) { const cf v = ps[LIn::pad((j << lgStrideA) + (a << lgStrideB))]; vr[a][j] = v.re; vi[a][j] = v.im; }
    {
      const cf *pt = tw + (iA << lgStrideA);
      const cf tB = pt[0], tC = pt[iA << lgStrideA], tD = pt[(2 * iA) << lgStrideA];
#pragma unroll
      for (int a = 0; a < 4; ++a) bfly4<INV>(vr[a][0], vi[a][0], vr[a][1], vi[a][1], vr[a][2], vi[a][2], vr[a][3], vi[a][3], tB, tC, tD);
    }
    cf *pd = dst + LOut::at(sub, (iA << lgStrideB) + sB);   // = element r of the sub-transform: consecutive lanes, consecutive elements
    const cf *pt = tw + (iA << lgStrideB);
#pragma unroll
    for (int jA = 0; jA < 4; ++jA) {
      // iB = iA + (jA << LGS): tw[iB << lgStrideB], tw[2 iB << lgStrideB], tw[3 iB << lgStrideB] = constant offsets from tw[k iA << lgStrideB]
      constexpr int one = 1 << (LGS + lgStrideB);
      cf tB, tC, tD;
      if constexpr (CACHED) { const cf *pc = twc + (jA * 3 << LGS) + iA; tB = pc[0]; tC = pc[1 << LGS]; tD = pc[2 << LGS]; }
      else { tB = pt[jA * one]; tC = pt[(iA << lgStrideB) + 2 * jA * one]; tD = pt[((2 * iA) << lgStrideB) + 3 * jA * one]; }
      bfly4<INV>(vr[0][jA], vi[0][jA], vr[1][jA], vi[1][jA], vr[2][jA], vi[2][jA], vr[3][jA], vi[3][jA], tB, tC, tD);
#pragma unroll
      for (int jB = 0; jB < 4; ++jB) {
        cf o; o.re = vr[jB][jA]; o.im = vi[jB][jA];
        pd[LOut::pad((jA << (LGS + lgStrideB)) + (jB << (LGS + 2 + lgStrideB)))] = o;
      }
    }
  }
}

// ---- a single radix-4 pass at sub-transform size 2^LGS (not the last one)
template <int LG, int OUTER, bool INV, int LGS, class LIn, class LOut>
BS_HD void fast_r4(const cf *tw, const cf *src, cf *dst, int tid) {
  constexpr int lgStride = LG - LGS - 2, stride = 1 << lgStride, lgPer = LG - 2, nItems = OUTER << lgPer;
  for (int idx = tid; idx < nItems; idx += kFastNT) {
    const int sub = idx >> lgPer, r = idx & ((1 << lgPer) - 1), i = r >> lgStride, s = r & (stride - 1);
    const cf tB = tw[i << lgStride], tC = tw[(2 * i) << lgStride], tD = tw[(3 * i) << lgStride];
    const cf *ps = src + LIn::at(sub, ((4 * i) << lgStride) + s);
    cf a = ps[0], b = ps[LIn::pad(stride)], c = ps[LIn::pad(2 * stride)], d = ps[LIn::pad(3 * stride)];
    bfly4<INV>(a.re, a.im, b.re, b.im, c.re, c.im, d.re, d.im, tB, tC, tD);
    constexpr int qs = 1 << (LGS + lgStride);
    cf *pd = dst + LOut::at(sub, (i << lgStride) + s);
    pd[0] = a; pd[LOut::pad(qs)] = b; pd[LOut::pad(2 * qs)] = c; pd[LOut::pad(3 * qs)] = d;
  }
}

// ---- the last radix-4 pass (stride 1) of every sub-transform of one bin quadruple, then the outer stage of the four bins
// it produces, all in registers.  emit(k, re, im) receives bin k = i + m * inner/4 + s * inner of the finished transform.
template <int LG, int OUTER, bool INV, class LIn, class Emit>
BS_HD void fast_last(const cf *tw, const cf *otw, const cf *src, int tid, Emit &&emit) {
  constexpr int inner = 1 << LG, quarter = inner >> 2;
  for (int i = tid; i < quarter; i += kFastNT) {
    float vr[OUTER][4], vi[OUTER][4];
    const cf tB = tw[i], tC = tw[2 * i], tD = tw[3 * i];
#pragma unroll
    for (int s = 0; s < OUTER; ++s) {
      const f4 *p = (const f4 *)(src + LIn::at(s, 4 * i));   // four consecutive elements, 16-byte aligned (padding is 2 per 16)
      const f4 lo = p[0], hi = p[1];
      vr[s][0] = lo.x; vi[s][0] = lo.y; vr[s][1] = lo.z; vi[s][1] = lo.w; vr[s][2] = hi.x; vi[s][2] = hi.y; vr[s][3] = hi.z; vi[s][3] = hi.w;
      bfly4<INV>(vr[s][0], vi[s][0], vr[s][1], vi[s][1], vr[s][2], vi[s][2], vr[s][3], vi[s][3], tB, tC, tD);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = i + m * quarter;
      float xr[OUTER], xi[OUTER];
      cf w[OUTER > 1 ? OUTER - 1 : 1];
      if constexpr (OUTER > 1 && ((OUTER - 1) & 1) == 0) {   // the bin's outer twiddles lie side by side: 16 bytes at a time
#pragma unroll
        for (int h = 0; h < (OUTER - 1) / 2; ++h) {
          const f4 v = ((const f4 *)otw)[(size_t)k * ((OUTER - 1) / 2) + h];
          w[2 * h].re = v.x; w[2 * h].im = v.y; w[2 * h + 1].re = v.z; w[2 * h + 1].im = v.w;
        }
      } else {
#pragma unroll
        for (int s = 1; s < OUTER; ++s) w[s - 1] = otw[(size_t)k * (OUTER - 1) + (s - 1)];
      }
#pragma unroll
      for (int s = 0; s < OUTER; ++s) { xr[s] = vr[s][m]; xi[s] = vi[s][m]; }
      outer_point<INV, OUTER>(xr, xi, w);
#pragma unroll
      for (int s = 0; s < OUTER; ++s) emit(k + s * inner, xr[s], xi[s]);
    }
  }
}

// The pack stage of the forward transform (analyse_window's pack loop): packed pair j = window samples i, i+1 with
// i = 2j + off (second half of the window, j < jA) or 2j - cStart (first half, j >= jC; the sign of the half-bin shift is
// in the table), zeros between; samples outside [lo, hi) -- beyond the clip, or the short pre-roll of a "previous"
// window -- read as zero.  PAIR: both samples of a pair are valid together and 8-byte aligned (one 64-bit load).
// Loads of U pairs are issued before the first one is used.
struct PackCtx { const float *xs; const f4 *tab; int lo, span, jA, jC, off, cStart; bool none; };
// the clip pointer comes out of a structure in global memory: say that the samples are in global memory too (LDG, not a generic LD)
#ifdef BS_HOSTEMU
BS_HD float pack_ld(const float *p) { return *p; }
BS_HD f2 pack_ld2(const float *p) { return *(const f2 *)p; }
#else
BS_HD float pack_ld(const float *p) { return __ldg(p); }
BS_HD f2 pack_ld2(const float *p) { const float2 v = __ldg((const float2 *)p); f2 r; r.x = v.x; r.y = v.y; return r; }
#endif
template <bool PAIR>
BS_HD void fast_pack_load(const PackCtx &c, int j, float &x0, float &x1, f4 &t, bool &live) {
  const bool inA = j < c.jA;
  live = inA || j >= c.jC;
  const int i = 2 * j + (inA ? c.off : -c.cStart);
  const bool v0 = live && !c.none && (unsigned)(i - c.lo) < (unsigned)c.span;
  t = c.tab[j];
  // no branch around the loads (the U pairs of a trip must be in flight together): a sample that is not there is read from
  // the nearest position that is (the caller has made sure there is one: span > 0) and replaced by zero
  const int hiM = c.lo + c.span - (PAIR ? 2 : 1);
  if (PAIR) { const int ic = min(max(i, c.lo), hiM); const f2 v = pack_ld2(c.xs + ic); x0 = v0 ? v.x : 0.f; x1 = v0 ? v.y : 0.f; }
  else {
    const bool v1 = live && !c.none && (unsigned)(i + 1 - c.lo) < (unsigned)c.span;
    const float a = pack_ld(c.xs + min(max(i, c.lo), hiM)), b = pack_ld(c.xs + min(max(i + 1, c.lo), hiM));
    x0 = v0 ? a : 0.f; x1 = v1 ? b : 0.f;
  }
}
BS_HD cf fast_pack_finish(float x0, float x1, const f4 t, bool live) {
  const float t0 = live ? x0 * t.x : 0.f, t1 = live ? x1 * t.y : 0.f;
  cf z; z.re = (t.z * t0) - (t.w * t1); z.im = (t.w * t0) + (t.z * t1);
  return z;
}
template <int LG, int OUTER, class LOut, bool PAIR>
BS_HD void fast_fwd_pack_t(const PackCtx &c, cf *dst, int tid) {
  using F = FastGeom<LG, OUTER>;
  constexpr int M = F::M, inner = F::inner;
  if (!F::lg0) {
    constexpr int U = 4;
    constexpr bool whole = (M % (kFastNT * U)) == 0;   // every trip is full: no bounds checks between the loads
    for (int j0 = tid; j0 < M; j0 += kFastNT * U) {
      float x0[U], x1[U]; f4 t[U]; bool live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { const int j = j0 + u * kFastNT; if (whole || j < M) fast_pack_load<PAIR>(c, j, x0[u], x1[u], t[u], live[u]); }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * kFastNT;
        if (whole || j < M) { const int q = j / OUTER, sub = j - q * OUTER; dst[LOut::at(sub, q)] = fast_pack_finish(x0[u], x1[u], t[u], live[u]); }
      }
    }
  } else {   // the radix-2 pass on elements q and q + inner/2 of a sub-transform = packed pairs j and j + M/2
    constexpr int U = (M / 2) % (2 * kFastNT) == 0 ? 2 : 1;   // full trips only
    static_assert((M / 2) % (kFastNT * U) == 0 || M / 2 < kFastNT, "pack trips");
    for (int j0 = tid; j0 < M / 2; j0 += kFastNT * U) {
      float x0[2 * U], x1[2 * U]; f4 t[2 * U]; bool live[2 * U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * kFastNT;
        if (j < M / 2) { fast_pack_load<PAIR>(c, j, x0[2 * u], x1[2 * u], t[2 * u], live[2 * u]); fast_pack_load<PAIR>(c, j + M / 2, x0[2 * u + 1], x1[2 * u + 1], t[2 * u + 1], live[2 * u + 1]); }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * kFastNT;
        if (j < M / 2) {
          const int q = j / OUTER, sub = j - q * OUTER;
          const cf a = fast_pack_finish(x0[2 * u], x1[2 * u], t[2 * u], live[2 * u]), b = fast_pack_finish(x0[2 * u + 1], x1[2 * u + 1], t[2 * u + 1], live[2 * u + 1]);
          cf lo, hi; lo.re = b.re + a.re; lo.im = b.im + a.im; hi.re = a.re - b.re; hi.im = a.im - b.im;
          cf *pd = dst + LOut::at(sub, q);
          pd[0] = lo; pd[LOut::pad(inner / 2)] = hi;
        }
      }
    }
  }
}
template <int LG, int OUTER, class LOut>
BS_HD void fast_fwd_pack(const DevGeom &g, const DevTables &T, const float *x, Window w, cf *dst, int tid) {
  PackCtx c;
  c.xs = x + w.start; c.tab = (const f4 *)T.packTab; c.lo = w.lo; c.span = w.hi - w.lo; c.none = false;
  c.off = g.off; c.cStart = g.N - g.off; c.jA = (g.L - g.off) >> 1; c.jC = c.cStart >> 1;
  // pairs start at even window positions: whole pairs are valid or not, and 8-byte aligned, if lo, hi and the start allow
  const bool pair = ((((size_t)c.xs) & 7) == 0) && ((w.lo | w.hi) & 1) == 0;
  if (c.span <= 0) {   // nothing of the clip in this window: every sample reads as zero (loads go to the table, results are dropped)
    c.xs = (const float *)T.packTab; c.lo = 0; c.span = 1; c.none = true;
    fast_fwd_pack_t<LG, OUTER, LOut, false>(c, dst, tid);
  } else if (pair) fast_fwd_pack_t<LG, OUTER, LOut, true>(c, dst, tid);
  else fast_fwd_pack_t<LG, OUTER, LOut, false>(c, dst, tid);
}

// untangle of the forward transform: bins i and M-1-i from the complex transform's outputs (analyse_window's last loop)
template <int LG, int OUTER>
BS_HD void fast_fwd_untangle(const DevTables &T, const cf *Y /* natural order */, cf *X, bool rotate, float *E /* |X|^2 too, or nullptr */, int tid) {
  constexpr int M = FastGeom<LG, OUTER>::M, half = M >> 1, U = 3;
  constexpr bool whole = (half % (kFastNT * U)) == 0;   // every trip is full: nothing between the loads
  for (int p0 = tid; p0 < half; p0 += kFastNT * U) {
    cf u[U], a[U], b[U], ri[U], rj[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int p = p0 + k * kFastNT;
      if (whole || p < half) {
        const int i = (p < half - 1) ? p : half, j = M - 1 - i;   // the reference's loop leaves pair (half, half-1) as written by its last iteration
        u[k] = T.untangle[i]; a[k] = Y[i]; b[k] = Y[j];
        if (rotate) { ri[k] = T.specRot[i]; rj[k] = T.specRot[j]; }
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int p = p0 + k * kFastNT;
      if (whole || p < half) {
        const int i = (p < half - 1) ? p : half, j = M - 1 - i;
        const float sI = (b[k].im + a[k].im) * 0.5f, dR = (a[k].re - b[k].re) * 0.5f;
        const float pp = (sI * u[k].re) + (dR * u[k].im), dI = (a[k].im - b[k].im) * 0.5f;
        const float qq = (dR * u[k].re) - (sI * u[k].im), sR = (b[k].re + a[k].re) * 0.5f;
        cf xi_, xj_;
        xi_.im = pp + dI; xi_.re = qq + sR; xj_.im = pp - dI; xj_.re = sR - qq;
        if (rotate) { xi_ = rot_prev(xi_, ri[k]); xj_ = rot_prev(xj_, rj[k]); }
        X[i] = xi_; X[j] = xj_;
        if (E) { E[i] = (xi_.im * xi_.im) + (xi_.re * xi_.re); E[j] = (xj_.im * xj_.im) + (xj_.re * xj_.re); }
      }
    }
  }
}

// untangle^-1 of the inverse transform (synth_frame's first loop): spectrum pairs -> the complex transform's inputs
BS_HD void fast_inv_pair(const cf un, const cf xi_, const cf xj_, cf &oi, cf &oj) {
  const float sI = xj_.im + xi_.im, dR = xi_.re - xj_.re;
  const float p = (sI * un.im) + (dR * un.re), sR = xj_.re + xi_.re;
  const float q = (sI * un.re) - (dR * un.im), dI = xi_.im - xj_.im;
  oi.re = p + sR; oi.im = q + dI; oj.re = sR - p; oj.im = q - dI;
}
template <int LG, int OUTER, class LOut>
BS_HD void fast_inv_untangle(const DevTables &T, const cf *X, cf *dst, int tid) {
  using F = FastGeom<LG, OUTER>;
  constexpr int M = F::M, half = M >> 1, inner = F::inner;
  auto put = [&](int k, cf v) { const int q = k / OUTER, sub = k - q * OUTER; dst[LOut::at(sub, q)] = v; };
  if (!F::lg0) {
    constexpr int U = 3;   // loads of U pairs in flight before the first is used
    for (int p0 = tid; p0 < half; p0 += kFastNT * U) {
      cf un[U], xi_[U], xj_[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int p = p0 + k * kFastNT;
        if (p < half) { const int i = (p < half - 1) ? p : half; un[k] = T.untangle[i]; xi_[k] = X[i]; xj_[k] = X[M - 1 - i]; }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int p = p0 + k * kFastNT;
        if (p < half) {
          const int i = (p < half - 1) ? p : half, j = M - 1 - i;
          cf oi, oj;
          fast_inv_pair(un[k], xi_[k], xj_[k], oi, oj);
          put(i, oi); put(j, oj);
        }
      }
    }
  } else {
    // with the radix-2 pass: elements k and k + M/2 meet, so a work item takes the pairs (i, M-1-i) and (M/2-1-i, M/2+i)
    constexpr int quarter = M >> 2;
    for (int i = tid; i < quarter; i += kFastNT) {
      cf a0, a1, b0, b1;   // a0 = element i, a1 = element M-1-i, b0 = element M/2+i, b1 = element M/2-1-i
      fast_inv_pair(T.untangle[i], X[i], X[M - 1 - i], a0, a1);
      if (i == 0) fast_inv_pair(T.untangle[half], X[half], X[half - 1], b0, b1);          // the pair the reference's loop writes last
      else fast_inv_pair(T.untangle[half - 1 - i], X[half - 1 - i], X[half + i], b1, b0);
      auto r2 = [&](int k, cf a, cf b) {   // a at k, b at k + M/2
        const int q = k / OUTER, sub = k - q * OUTER;
        cf lo, hi; lo.re = b.re + a.re; lo.im = b.im + a.im; hi.re = a.re - b.re; hi.im = a.im - b.im;
        dst[LOut::at(sub, q)] = lo; dst[LOut::at(sub, q + inner / 2)] = hi;
      };
      r2(i, a0, b0);
      r2(half - 1 - i, b1, a1);
    }
  }
  (void)put;
}

// the twiddle cache of TwCache's stage: entry (jA, m, iA) = tw[(m + 1) * ((iA << lgStrideB) + (jA << (LG - 4)))], the value fast_r16 reads
template <int LG, int OUTER>
BS_HD void fast_fill_twcache(const cf *tw, cf *twc, int tid) {
  using C = TwCache<LG, OUTER>;
  if constexpr (C::stage >= 0) {
    constexpr int lgStrideB = LG - C::lgs - 4;
    for (int idx = tid; idx < C::elems; idx += kFastNT) {
      const int iA = idx & (C::sets - 1), jm = idx >> C::lgs, jA = jm / 3, m = jm - 3 * jA;
      twc[idx] = tw[(m + 1) * ((iA << lgStrideB) + (jA << (LG - 4)))];
    }
  }
}

// ---- stage sequencing: stage K reads buffer (K & 1), writes the other one
template <int LG, int OUTER, bool INV, int K>
BS_HD void fast_mid_stage(const cf *tw, cf *buf0, cf *buf1, int tid) {
  using S = StageOf<LG, OUTER, K>;
  using Nx = StageOf<LG, OUTER, K + 1>;
  const cf *src = (K & 1) ? buf1 : buf0; cf *dst = (K & 1) ? buf0 : buf1;
  if constexpr (S::isR16 && TwCache<LG, OUTER>::stage == K)
    fast_r16<LG, OUTER, INV, S::lgSize, typename S::In, typename Nx::In, true>(tw, src, dst, tid, buf0 + 2 * FastBuf<LG, OUTER>::elems);   // (buf1 = buf0 + elems)
  else if constexpr (S::isR16) fast_r16<LG, OUTER, INV, S::lgSize, typename S::In, typename Nx::In>(tw, src, dst, tid);
  else fast_r4<LG, OUTER, INV, S::lgSize, typename S::In, typename Nx::In>(tw, src, dst, tid);
}

// Runs the stages between the pack and the last one; SYNC() between them.  Returns nothing: after it the input of the last
// stage is in buffer ((nStages - 1) & 1).
#ifdef BS_HOSTEMU
#define BS_FAST_FORALL(stmt) for (int tid = 0; tid < kFastNT; ++tid) { stmt; }
#else
#define BS_FAST_FORALL(stmt) { const int tid = threadIdx.x; stmt; } __syncthreads();
#endif
template <int LG, int OUTER, bool INV>
BS_HD void fast_mid_stages(const cf *tw, cf *buf0, cf *buf1) {
  using F = FastGeom<LG, OUTER>;
  if constexpr (F::nStages > 1) { BS_FAST_FORALL((fast_mid_stage<LG, OUTER, INV, 0>(tw, buf0, buf1, tid))) }
  if constexpr (F::nStages > 2) { BS_FAST_FORALL((fast_mid_stage<LG, OUTER, INV, 1>(tw, buf0, buf1, tid))) }
  if constexpr (F::nStages > 3) { BS_FAST_FORALL((fast_mid_stage<LG, OUTER, INV, 2>(tw, buf0, buf1, tid))) }
  static_assert(F::nStages <= 4, "add a stage");
}

// whole forward transform of one window of one channel; sm = 2 * fast_buf_elems cf (16-byte aligned)
template <int LG, int OUTER>
BS_HD void fast_analyse(const DevGeom &g, const DevTables &T, const float *x, Window w, cf *X, cf *sm, bool rotate, float *E = nullptr) {
  using F = FastGeom<LG, OUTER>;
  constexpr int NB = FastBuf<LG, OUTER>::elems;
  cf *buf0 = sm, *buf1 = sm + NB;
  using L0 = typename StageOf<LG, OUTER, 0>::In;
  BS_FAST_FORALL((fast_fill_twcache<LG, OUTER>(T.tw, buf0 + 2 * NB, tid), fast_fwd_pack<LG, OUTER, L0>(g, T, x, w, buf0, tid)))
  fast_mid_stages<LG, OUTER, false>(T.tw, buf0, buf1);
  constexpr int KL = F::nStages - 1;
  using LL = typename StageOf<LG, OUTER, KL>::In;
  cf *src = (KL & 1) ? buf1 : buf0, *Y = (KL & 1) ? buf0 : buf1;
  BS_FAST_FORALL((fast_last<LG, OUTER, false, LL>(T.tw, T.otw, src, tid, [&](int k, float re, float im) { cf v; v.re = re; v.im = im; Y[k] = v; })))
  BS_FAST_FORALL((fast_fwd_untangle<LG, OUTER>(T, Y, X, rotate, E, tid)))
}

// whole inverse transform of one channel's output spectrum, synthesis window applied (synth_frame)
template <int LG, int OUTER>
BS_HD void fast_synth(const DevGeom &g, const DevTables &T, const cf *X, float *frame, cf *sm) {
  using F = FastGeom<LG, OUTER>;
  constexpr int NB = FastBuf<LG, OUTER>::elems;
  cf *buf0 = sm, *buf1 = sm + NB;
  using L0 = typename StageOf<LG, OUTER, 0>::In;
  BS_FAST_FORALL((fast_fill_twcache<LG, OUTER>(T.tw, buf0 + 2 * NB, tid), fast_inv_untangle<LG, OUTER, L0>(T, X, buf0, tid)))
  fast_mid_stages<LG, OUTER, true>(T.tw, buf0, buf1);
  constexpr int KL = F::nStages - 1;
  using LL = typename StageOf<LG, OUTER, KL>::In;
  const cf *src = (KL & 1) ? buf1 : buf0;
  const f4 *tab = (const f4 *)T.packTab;
  const int jA = (g.L - g.off) >> 1, jC = (g.N - g.off) >> 1, off = g.off, cStart = g.N - g.off;
  // the last pass hands every finished time-domain pair straight to the half-bin rotation and the synthesis window; the
  // table's window coefficients carry the sign of the first half (ring -= t*w there = adding -(t*w) = t*(-w))
  BS_FAST_FORALL((fast_last<LG, OUTER, true, LL>(T.tw, T.otw, src, tid, [&](int j, float re, float im) {
    const bool inA = j < jA;
    const f4 t = tab[j];
    const float t1 = (t.z * im) - (t.w * re), t0 = (t.w * im) + (t.z * re);
    f2 o; o.x = t0 * t.x; o.y = t1 * t.y;
    if (inA || j >= jC) *(f2 *)(frame + (inA ? 2 * j + off : 2 * j - cStart)) = o;
  })))
}

// geometries with a specialised path (every other one takes analyse_window / synth_frame)
BS_HHD bool fast_geometry(int inner, int outer) {
  return (inner == 1024 && outer == 3) || (inner == 512 && outer == 5) || (inner == 1024 && outer == 5) || (inner == 2048 && outer == 3) ||
         (inner == 512 && outer == 1);
}
// ... provided the window halves fall on pair boundaries (they do for every even block size)
BS_HHD bool fast_ok(const DevGeom &g) {
  return g.packTabOk && fast_geometry(g.inner, g.outer) && (((g.L - g.off) | (g.N - g.off) | g.off) & 1) == 0;
}

}  // namespace bs
