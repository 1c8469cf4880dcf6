/*
 * stretch_oracle.c -- CPU restatement of the reference's time-stretch / pitch-shift engine.
 *
 * TEST INFRASTRUCTURE ONLY ("kind = port").  Nothing in the product path may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * The reference (hanskerkhof/BAUKLANK-audio-stretch) ships its DSP only as a compiled WebAssembly blob
 * embedded at app/SignalsmithStretch.mjs:265 (sha256 83869197...2d8ca3); the source is the third-party
 * Signalsmith Stretch C++ library (version unrecorded by the reference).  Every function below restates
 * the arithmetic of one wasm function of that blob -- W#n = wasm function index n, as enumerated in
 * SURVEY.md section 8a -- in the SAME f32/f64 operation order, so that with -ffp-contract=off the results
 * are bit-identical to the blob.  PARITY PIN: tests/test_oracle.py compares this port bit-for-bit with
 * oracle/_ref (the blob translated mechanically by oracle/wasm2c.py) and with the known answers
 * KA1..KA6 of SURVEY.md section 8c, committed as fixtures under tests/golden/.
 *
 * Build: gcc -O2 -fno-fast-math -ffp-contract=off -fPIC -shared -o libstretch_oracle.so stretch_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float re, im; } c32;
/* Band record of the blob: 28 bytes {input, prevInput, output, inputEnergy} */
typedef struct { c32 input, prevInput, output; float inputEnergy; } Band;
typedef struct { float energy; c32 input; } Pred;
typedef struct { float inputBin, freqGrad; } MapPoint;
typedef struct { float input, output; } Peak;

/* ------------------------------------------------------------------------------------------------
 * musl sinf/cosf as compiled into the blob (W#9 sinf, W#10 cosf, W#11 __sindf, W#12 __cosdf).
 * Only the |x| <= 9pi/4 branches are restated; every twiddle angle the engine forms lies inside. */
static float k_sindf(double x) {
  double z = x * x, s = x * z;
  return (float)(((s * (z * z)) * ((z * 0x1.6cd878c3b46a7p-19) + -0x1.a00f9e2cae774p-13)) +
                 ((s * ((z * 0x1.11110896efbb2p-7) + -0x1.5555554cbac77p-3)) + x));
}
static float k_cosdf(double x) {
  double z = x * x, w = z * z;
  return (float)(((z * w) * ((z * 0x1.99342e0ee5069p-16) + -0x1.6c087e80f1e27p-10)) +
                 ((w * 0x1.55553e1053a42p-5) + ((z * -0x1.ffffffd0c5e81p-2) + 1.0)));
}
static const double PIO2 = 0x1.921fb54442d18p+0, PI_ = 0x1.921fb54442d18p+1, PI3O2 = 0x1.2d97c7f3321d2p+2,
                    PI2 = 0x1.921fb54442d18p+2;
static uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float m_sinf(float x) {
  uint32_t b = fbits(x), ix = b & 0x7fffffffu; int neg = (int32_t)b < 0;
  if (ix <= 0x3f490fdau) { if (ix < 0x39800000u) return x; return k_sindf((double)x); }
  if (ix <= 0x407b53d1u) {
    if (ix <= 0x4016cbe3u) { if (neg) return -k_cosdf((double)x + PIO2); return k_cosdf((double)x + -PIO2); }
    return k_sindf(-((neg ? PI_ : -PI_) + (double)x));
  }
  if (ix <= 0x40e231d5u) {
    if (ix <= 0x40afeddfu) { if (neg) return k_cosdf((double)x + PI3O2); return -k_cosdf((double)x + -PI3O2); }
    return k_sindf((neg ? PI2 : -PI2) + (double)x);
  }
  fprintf(stderr, "stretch_oracle: sinf argument out of restated range\n"); abort();
}
static float m_cosf(float x) {
  uint32_t b = fbits(x), ix = b & 0x7fffffffu; int neg = (int32_t)b < 0;
  if (ix <= 0x3f490fdau) { if (ix < 0x39800000u) return 1.0f; return k_cosdf((double)x); }
  if (ix <= 0x407b53d1u) {
    if (ix >= 0x4016cbe4u) return -k_cosdf((neg ? PI_ : -PI_) + (double)x);
    if (neg) return k_sindf((double)x + PIO2);
    return k_sindf(PIO2 - (double)x);
  }
  if (ix <= 0x40e231d5u) {
    if (ix >= 0x40afede0u) return k_cosdf((neg ? PI2 : -PI2) + (double)x);
    if (neg) return k_sindf(-PI3O2 - (double)x);
    return k_sindf((double)x + -PI3O2);
  }
  fprintf(stderr, "stretch_oracle: cosf argument out of restated range\n"); abort();
}
/* wasm i32.trunc_f32_s with the blob's guard (|x| < 2^31 else INT_MIN) */
static int32_t trunc_i32(float x) { return fabsf(x) < 2147483648.0f ? (int32_t)x : INT32_MIN; }

/* ------------------------------------------------------------------------------------------------
 * FFT: half-bin-shifted real FFT over a split-complex FFT of M = N/2 = inner(pow2) x outer(odd)
 * tables: W#38 (5181-5947 of the translation); passes: W#21/W#34 forward, W#20/W#33 inverse. */
typedef struct {
  int N, M, inner, outer;
  c32 *tw;            /* pow2 twiddles [3*inner/4] */
  float *otr, *oti;   /* outer twiddles, split [inner*(outer-1)] */
  c32 *untangle;      /* [N/4+1] */
  c32 *rot;           /* half-bin rotations [M] */
  float *w1, *w2, *tmp, *work; /* 2M each (re then im); work = pow2 ping-pong [2*inner] */
} FFT;

static void fft_free(FFT *f) {
  free(f->tw); free(f->otr); free(f->oti); free(f->untangle); free(f->rot);
  free(f->w1); free(f->w2); free(f->tmp); free(f->work);
  memset(f, 0, sizeof(*f));
}

static void fft_setup(FFT *f, int N) {
  fft_free(f);
  int M = N >> 1, inner = 1, outer = M;
  while (outer > 1 && !(outer & 1)) { outer >>= 1; inner <<= 1; }   /* W#38: strip all factors of two */
  f->N = N; f->M = M; f->inner = inner; f->outer = outer;
  if (outer != 1 && outer != 3 && outer != 5) {
    fprintf(stderr, "stretch_oracle: outer factor %d not restated (engine sizes give 1, 3 or 5)\n", outer); abort();
  }
  int ntw = (3 * inner) >> 2;
  f->tw = (c32 *)calloc(ntw > 0 ? ntw : 1, sizeof(c32));
  double rinner = 1.0 / (double)inner;
  for (int i = 0; i < ntw; ++i) {
    float a = (float)(((double)i * -PI2) * rinner);
    f->tw[i].im = m_sinf(a); f->tw[i].re = m_cosf(a);
  }
  int not_ = inner * (outer - 1);
  f->otr = (float *)calloc(not_ > 0 ? not_ : 1, 4); f->oti = (float *)calloc(not_ > 0 ? not_ : 1, 4);
  for (int i = 0; i < inner && outer >= 2; ++i) {
    double a0 = (double)i * -PI2;
    for (int s = 1; s < outer; ++s) {
      float a = (float)((a0 * (double)s) / ((double)inner * (double)outer));
      f->oti[i + inner * (s - 1)] = m_sinf(a); f->otr[i + inner * (s - 1)] = m_cosf(a);
    }
  }
  int nun = (N >> 2) + 1;
  f->untangle = (c32 *)calloc(nun, sizeof(c32));
  double rN = 1.0 / (double)N;
  for (int i = 0; i < nun; ++i) {
    float a = (float)(((((double)i * -PI2) + -PI_) * rN) + -PIO2);
    f->untangle[i].im = m_sinf(a); f->untangle[i].re = m_cosf(a);
  }
  f->rot = (c32 *)calloc(M, sizeof(c32));
  for (int i = 0; i < M; ++i) {
    /* pairs use *(1/N); a trailing odd element (only when N&2) uses /N -- W#38 5911-5942 */
    float a = ((N & 2) && i == M - 1) ? (float)(((double)i * -PI2) / (double)N) : (float)(((double)i * -PI2) * rN);
    f->rot[i].im = m_sinf(a); f->rot[i].re = m_cosf(a);
  }
  f->w1 = (float *)calloc(2 * M, 4); f->w2 = (float *)calloc(2 * M, 4); f->tmp = (float *)calloc(2 * M, 4);
  f->work = (float *)calloc(2 * inner, 4);
}

/* combine4, W#34 (forward) / W#33 (inverse): radix-4 DIT on split arrays, blocks of `stride` */
static void combine4(const FFT *f, int inverse, int size, int stride, const float *ir, const float *ii, float *or_,
                     float *oi) {
  int step = f->inner / size, q = size >> 2;
  for (int i = 0; i < q; ++i) {
    c32 tB = f->tw[i * step], tC = f->tw[i * 2 * step], tD = f->tw[i * 3 * step];
    const float *Ar = ir + (4 * i) * stride, *Ai = ii + (4 * i) * stride;
    const float *Br = ir + (4 * i + 1) * stride, *Bi = ii + (4 * i + 1) * stride;
    const float *Cr = ir + (4 * i + 2) * stride, *Ci = ii + (4 * i + 2) * stride;
    const float *Dr = ir + (4 * i + 3) * stride, *Di = ii + (4 * i + 3) * stride;
    float *oAr = or_ + i * stride, *oAi = oi + i * stride;
    float *oBr = or_ + (i + q) * stride, *oBi = oi + (i + q) * stride;
    float *oCr = or_ + (i + 2 * q) * stride, *oCi = oi + (i + 2 * q) * stride;
    float *oDr = or_ + (i + 3 * q) * stride, *oDi = oi + (i + 3 * q) * stride;
    for (int s = 0; s < stride; ++s) {
      float dRe, bRe, cRe, dIm, bIm, cIm;
      if (!inverse) {
        dRe = (Dr[s] * tD.re) - (Di[s] * tD.im); bRe = (Br[s] * tB.re) - (Bi[s] * tB.im);
        cRe = (Cr[s] * tC.re) - (Ci[s] * tC.im);
        dIm = (Di[s] * tD.re) + (Dr[s] * tD.im); bIm = (Bi[s] * tB.re) + (Br[s] * tB.im);
        cIm = (Ci[s] * tC.re) + (Cr[s] * tC.im);
      } else {
        dRe = (Di[s] * tD.im) + (Dr[s] * tD.re); bRe = (Bi[s] * tB.im) + (Br[s] * tB.re);
        cRe = (Ci[s] * tC.im) + (Cr[s] * tC.re);
        dIm = (Di[s] * tD.re) - (Dr[s] * tD.im); bIm = (Bi[s] * tB.re) - (Br[s] * tB.im);
        cIm = (Ci[s] * tC.re) - (Cr[s] * tC.im);
      }
      float bdRe = dRe + bRe, acRe = cRe + Ar[s];
      float bdIm = dIm + bIm, acIm = Ai[s] + cIm;
      float x = inverse ? (dIm - bIm) : (bIm - dIm), y = Ar[s] - cRe;
      float z = inverse ? (bRe - dRe) : (dRe - bRe), w = Ai[s] - cIm;
      oAr[s] = bdRe + acRe; oAi[s] = bdIm + acIm;
      oBr[s] = x + y;       oBi[s] = z + w;
      oCr[s] = acRe - bdRe; oCi[s] = acIm - bdIm;
      oDr[s] = y - x;       oDi[s] = w - z;
    }
  }
}
/* fftPass, W#21/W#20 */
static void fft_pass(const FFT *f, int inverse, int size, int stride, const float *ir, const float *ii, float *or_,
                     float *oi, float *wr, float *wi) {
  if (size > 7) {
    fft_pass(f, inverse, size >> 2, stride << 2, ir, ii, wr, wi, or_, oi);
    combine4(f, inverse, size, stride, wr, wi, or_, oi);
  } else if (size == 4) {
    combine4(f, inverse, 4, stride, ir, ii, or_, oi);
  } else {
    for (int s = 0; s < stride; ++s) {
      float ai = ii[s], bi = ii[s + stride], br = ir[s + stride], ar = ir[s];
      or_[s] = br + ar; oi[s] = bi + ai; or_[s + stride] = ar - br; oi[s + stride] = ai - bi;
    }
  }
}
static void pow2_fft(const FFT *f, int inverse, const float *ir, const float *ii, float *or_, float *oi) {
  if (f->inner <= 1) { *or_ = *ir; *oi = *ii; return; }
  fft_pass(f, inverse, f->inner, 1, ir, ii, or_, oi, f->work, f->work + f->inner);
}
/* complex split FFT of length M: src (re, im) -> dst (re, im); plan interpreter inlined in W#35 / W#48 */
static void split_fft(FFT *f, int inverse, const float *sr, const float *si, float *dr, float *di) {
  int M = f->M, inner = f->inner, outer = f->outer;
  if (outer < 2) { pow2_fft(f, inverse, sr, si, dr, di); return; }
  float *tr = f->tmp, *ti = f->tmp + M;
  for (int i = 0; i < inner; ++i)
    for (int s = 0; s < outer; ++s) { tr[s * inner + i] = sr[i * outer + s]; ti[s * inner + i] = si[i * outer + s]; }
  for (int s = 0; s < outer; ++s) pow2_fft(f, inverse, tr + s * inner, ti + s * inner, dr + s * inner, di + s * inner);
  for (int k = 0; k < inner * (outer - 1); ++k) {
    float xr = dr[inner + k], xi = di[inner + k], wr = f->otr[k], wi = f->oti[k];
    if (!inverse) { dr[inner + k] = (wr * xr) - (wi * xi); di[inner + k] = (wi * xr) + (xi * wr); }
    else { dr[inner + k] = (xi * wi) + (xr * wr); di[inner + k] = (xi * wr) - (wi * xr); }
  }
  if (outer == 3) {
    const float h = inverse ? 0x1.bb67aep-1f : -0x1.bb67aep-1f; /* +-0.8660254 (0x3f5db3d7) */
    for (int i = 0; i < inner; ++i) {
      float ar = dr[i], br = dr[i + inner], cr = dr[i + 2 * inner];
      float ai = di[i], bi = di[i + inner], ci = di[i + 2 * inner];
      dr[i] = (br + ar) + cr; di[i] = ci + (bi + ai);
      float p = ar + (br * -0.5f), q = bi * h, r = cr * -0.5f, t = ci * h;
      float u = ai + (bi * -0.5f), v = br * h, x = cr * h, y = ci * -0.5f;
      dr[i + inner] = ((p - q) + r) + t;     di[i + inner] = ((u + v) - x) + y;
      dr[i + 2 * inner] = ((p + q) + r) - t; di[i + 2 * inner] = ((u - v) + x) + y;
    }
  } else { /* outer == 5 */
    const float c1 = 0x1.3c6ef4p-2f, c2 = 0x1.9e377ap-1f, s1 = 0x1.e6f0e2p-1f, s2 = 0x1.2cf23p-1f;
    for (int i = 0; i < inner; ++i) {
      float ar = dr[i], br = dr[i + inner], cr = dr[i + 2 * inner], er = dr[i + 4 * inner], d_r = dr[i + 3 * inner];
      float ai = di[i], bi = di[i + inner], ci = di[i + 2 * inner], ei = di[i + 4 * inner], d_i = di[i + 3 * inner];
      float dcR = d_r + cr, ebR = er + br, dcI = d_i + ci, ebI = ei + bi;
      dr[i] = (dcR + ar) + ebR; di[i] = (ai + dcI) + ebI;
      float p1r = ar + ((ebR * c1) - (dcR * c2)), p1i = ai + ((ebI * c1) - (dcI * c2));
      float p2r = ar + ((dcR * c1) - (ebR * c2)), p2i = ai + ((dcI * c1) - (ebI * c2));
      float q1r, q1i, q2r, q2i;
      if (!inverse) {
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
/* forward modified real FFT: t[N] -> X[M], X[k] = sum t[n] e^{-2 pi i (k+1/2) n / N}; W#35 4016-4833 */
static void rfft_forward(FFT *f, const float *t, c32 *X) {
  int M = f->M;
  float *ar = f->w2, *ai = f->w2 + M, *br = f->w1, *bi = f->w1 + M;
  for (int j = 0; j < M; ++j) {
    c32 r = f->rot[j]; float x = t[2 * j], y = t[2 * j + 1];
    ar[j] = (r.re * x) - (r.im * y); ai[j] = (r.im * x) + (r.re * y);
  }
  split_fft(f, 0, ar, ai, br, bi);
  for (int i = 0; i <= (M >> 1); ++i) {
    int j = M - 1 - i; c32 w = f->untangle[i];
    float sI = (bi[j] + bi[i]) * 0.5f, dR = (br[i] - br[j]) * 0.5f;
    float p = (sI * w.re) + (dR * w.im), dI = (bi[i] - bi[j]) * 0.5f;
    float q = (dR * w.re) - (sI * w.im), sR = (br[j] + br[i]) * 0.5f;
    X[i].im = p + dI; X[i].re = q + sR; X[j].im = p - dI; X[j].re = sR - q;
  }
}
/* inverse: X[M] -> t[N] (unnormalised); W#48 9986-10807 */
static void rfft_inverse(FFT *f, const c32 *X, float *t) {
  int M = f->M;
  float *ar = f->w1, *ai = f->w1 + M, *br = f->w2, *bi = f->w2 + M;
  for (int i = 0; i <= (M >> 1); ++i) {
    int j = M - 1 - i; c32 w = f->untangle[i];
    float xjI = X[j].im, xiI = X[i].im, xiR = X[i].re, xjR = X[j].re;
    float sI = xjI + xiI, dR = xiR - xjR;
    float p = (sI * w.im) + (dR * w.re), sR = xjR + xiR;
    float q = (sI * w.re) - (dR * w.im), dI = xiI - xjI;
    ar[i] = p + sR; ai[i] = q + dI; ar[j] = sR - p; ai[j] = q - dI;
  }
  split_fft(f, 1, ar, ai, br, bi);
  for (int j = 0; j < M; ++j) {
    c32 r = f->rot[j];
    t[2 * j + 1] = (r.re * bi[j]) - (r.im * br[j]);
    t[2 * j] = (r.im * bi[j]) + (r.re * br[j]);
  }
}

/* ------------------------------------------------------------------------------------------------ */
typedef struct { int pos; float *ring; float *wp; int sinceSynth; } OutState; /* stft.output (+ stash) */
typedef struct { int pos; float *ring; } InState;

typedef struct Engine {
  int channels, L, H, N, B, split, inLen; /* inLen = L+H+1 */
  FFT fft;
  float *win;                 /* analysis == synthesis window [L] */
  InState in, stashIn;
  OutState out, stashOut;
  c32 *spectrum;              /* [C*B] */
  float *timeBuf;             /* [N] */
  float *tmpBuf;              /* [L+H] */
  Band *bands; Pred *preds; MapPoint *map; Peak *peaks; int nPeaks;
  float *energy, *smoothed, *formantMetric;
  /* blockProcess */
  uint32_t samplesSinceLast; int steps, step; int newSpectrum, reanalysePrev, mapped, formants; float timeFactor;
  int spectrumSteps;
  uint32_t silenceCounter; int silenceFirst;
  float freqMultiplier, freqTonalityLimit, formantMultiplier, invFormantMultiplier, formantBaseFreq, formantBaseBin;
  int formantCompensation;
  float freqEstimateWeighted, freqEstimateWeight, smoothCarry;
  int prevInputOffset, didSeek; float seekTimeFactor;
  uint32_t rng;
  float *buffers; int bufCh, bufLen;
  int prevCopiedInput;
} Engine;

static __thread Engine *cur;

/* W#23 moveOutput */
static void move_output(Engine *e, OutState *o, int n) {
  int L = e->L;
  for (int c = 0; c < e->channels; ++c)
    for (int i = 0; i < n; ++i) o->ring[c * L + (o->pos + i) % L] = 0.0f;
  for (int i = 0; i < n; ++i) o->wp[(o->pos + i) % L] = 1e-30f;
  o->pos = (o->pos + n) % L;
  o->sinceSynth += n;
}
/* W#22 stft.reset(weight) */
static void stft_reset(Engine *e, float weight) {
  int L = e->L, H = e->H;
  e->out.pos = 0; e->in.pos = L;
  memset(e->in.ring, 0, sizeof(float) * e->channels * e->inLen);
  memset(e->out.ring, 0, sizeof(float) * e->channels * L);
  memset(e->spectrum, 0, sizeof(c32) * e->channels * e->B);
  memset(e->out.wp, 0, sizeof(float) * L);
  e->out.sinceSynth = 0;
  float fN = (float)(uint32_t)e->N;
  for (int i = 0; i < L; ++i) e->out.wp[i] = ((e->win[i] * fN) * e->win[i]) + e->out.wp[i];
  for (int i = L - H - 1; i >= 0; --i) e->out.wp[i] = e->out.wp[i] + e->out.wp[i + H];
  for (int i = 0; i < L; ++i) e->out.wp[i] = (e->out.wp[i] * weight) + 1e-30f;
  move_output(e, &e->out, H);
}
static void stash_all(Engine *e) {
  e->stashIn.pos = e->in.pos; memcpy(e->stashIn.ring, e->in.ring, sizeof(float) * e->channels * e->inLen);
  e->stashOut.pos = e->out.pos; e->stashOut.sinceSynth = e->out.sinceSynth;
  memcpy(e->stashOut.ring, e->out.ring, sizeof(float) * e->channels * e->L);
  memcpy(e->stashOut.wp, e->out.wp, sizeof(float) * e->L);
}
/* W#36 setInterval(H, kaiser) with forcePerfectReconstruction */
static void make_window(Engine *e) {
  int L = e->L, H = e->H;
  double dL = (double)L, bw = dL / (double)H, t = bw + 3.0;
  double heur = (8.0 / (t * t)) + bw, rem = 3.0 - bw;
  bw = heur + ((rem < 0.0 ? 0.0 : rem) * 0.25);
  bw = bw < 2.0 ? 2.0 : bw;
  double beta = sqrt(((bw * bw) * 0.25) + -1.0) * PI_, b2 = beta * beta;
  double term = 1.0, sum = 0.0, k = 0.0;
  do { sum = sum + term; k = k + 1.0; term = (b2 * term) / ((k * k) * 4.0); } while (term > 1e-4);
  double invI0 = 1.0 / sum, invL = 1.0 / dL;
  for (int i = 0; i < L; ++i) {
    double r = ((double)(uint32_t)((i << 1) | 1) * invL) + -1.0;
    double arg = sqrt(1.0 - (r * r)) * beta, a2 = arg * arg;
    term = 1.0; sum = 0.0; k = 0.0;
    do { sum = sum + term; k = k + 1.0; term = (a2 * term) / ((k * k) * 4.0); } while (term > 1e-4);
    e->win[i] = (float)(sum * invI0);
  }
  for (int i = 0; i < H; ++i) {
    double s = 0.0;
    if (i >= L) continue;
    for (int j = i; j < L; j += H) { float w = e->win[j]; s = s + (double)(w * w); }
    double f = 1.0 / sqrt(s);
    for (int j = i; j < L; j += H) e->win[j] = (float)((double)e->win[j] * f);
  }
}
static void engine_free_bufs(Engine *e) {
  fft_free(&e->fft);
  free(e->win); free(e->in.ring); free(e->stashIn.ring); free(e->out.ring); free(e->out.wp);
  free(e->stashOut.ring); free(e->stashOut.wp); free(e->spectrum); free(e->timeBuf); free(e->tmpBuf);
  free(e->bands); free(e->preds); free(e->map); free(e->peaks); free(e->energy); free(e->smoothed);
  free(e->formantMetric);
}
static void reset_block_process(Engine *e) {
  e->samplesSinceLast = 0xffffffffu; e->steps = 0; e->step = 0;
  e->newSpectrum = e->reanalysePrev = e->mapped = e->formants = 0; e->timeFactor = 0.0f;
}
/* W#25 configure */
static void engine_configure(Engine *e, int ch, int L, int H, int split) {
  engine_free_bufs(e);
  e->channels = ch; e->L = L; e->H = H; e->split = split & 1;
  uint32_t x = ((((uint32_t)L + 1u) >> 1) + 1u) >> 1, lim = x >= 16u ? 16u : x, p = 1u, q;
  do { q = p; p = q << 1; } while (q < lim);      /* q: first power of two >= min(x,16) */
  do { p = q; q = p << 1; } while ((p << 3) < x); /* p: grow until 8p >= x */
  uint32_t m = (p + x - 1u) / p;
  e->N = (int)((p * (m == 7u ? 8u : m)) << 2);
  e->B = e->N >> 1; e->inLen = L + H + 1;
  fft_setup(&e->fft, e->N);
  int C = ch, B = e->B;
  e->win = (float *)calloc(L, 4);
  e->in.ring = (float *)calloc((size_t)C * e->inLen, 4); e->stashIn.ring = (float *)calloc((size_t)C * e->inLen, 4);
  e->out.ring = (float *)calloc((size_t)C * L, 4); e->out.wp = (float *)calloc(L, 4);
  e->stashOut.ring = (float *)calloc((size_t)C * L, 4); e->stashOut.wp = (float *)calloc(L, 4);
  e->spectrum = (c32 *)calloc((size_t)C * B, sizeof(c32));
  e->timeBuf = (float *)calloc(e->N, 4); e->tmpBuf = (float *)calloc(L + H, 4);
  e->bands = (Band *)calloc((size_t)C * B, sizeof(Band)); e->preds = (Pred *)calloc((size_t)C * B, sizeof(Pred));
  e->map = (MapPoint *)calloc(B, sizeof(MapPoint)); e->peaks = (Peak *)calloc(B, sizeof(Peak)); e->nPeaks = 0;
  e->energy = (float *)calloc(B, 4); e->smoothed = (float *)calloc(B, 4);
  e->formantMetric = (float *)calloc(B + 2, 4);
  make_window(e);
  stft_reset(e, 0.1f);
  stash_all(e);
  reset_block_process(e);
}

/* W#24 copyInput(toIndex) (lambda inside process) */
static void copy_input(Engine *e, int toIndex) {
  int length = toIndex - e->prevCopiedInput, cap = e->L + e->H;
  if (length > cap) length = cap;
  if (length > 0) {
    int offset = toIndex - length;
    for (int c = 0; c < e->channels; ++c) {
      const float *src = e->buffers + (size_t)e->bufLen * c;
      float *ring = e->in.ring + (size_t)c * e->inLen;
      for (int i = 0; i < length; ++i) ring[(e->in.pos + i) % e->inLen] = src[offset + i];
    }
  }
  e->in.pos = (int)(((uint32_t)length + (uint32_t)e->in.pos) % (uint32_t)e->inLen);
  e->prevCopiedInput = toIndex;
}
/* W#35 analyse(channel, samplesInPast) reading the ring `in` */
static void analyse(Engine *e, const InState *in, int c, int samplesInPast) {
  int L = e->L, N = e->N, len = e->inLen, off = L >> 1;
  int start = (int)(((uint32_t)in->pos + ((uint32_t)len << 1) - (uint32_t)(samplesInPast + L)) % (uint32_t)len);
  const float *ring = in->ring + (size_t)c * len;
  float *t = e->timeBuf;
  for (int i = 0; i < off; ++i) t[N - off + i] = ring[(start + i) % len] * (-e->win[i]);
  for (int i = off; i < L; ++i) t[i - off] = ring[(start + i) % len] * e->win[i];
  for (int i = L - off; i < N - off; ++i) t[i] = 0.0f;
  rfft_forward(&e->fft, t, e->spectrum + (size_t)c * e->B);
}

static float map_freq(const Engine *e, float f) {
  if (!(f <= e->freqTonalityLimit)) return ((e->freqMultiplier + -1.0f) * e->freqTonalityLimit) + f;
  return e->freqMultiplier * f;
}
/* one-pole smoothing pass pair used by smoothEnergy and the formant envelope */
static float smooth_back_fwd(float *v, int n, float slew, float s) {
  for (int i = n - 1; i >= 0; --i) { s = ((v[i] - s) * slew) + s; v[i] = s; }
  for (int i = 0; i < n; ++i) { s = ((v[i] - s) * slew) + s; v[i] = s; }
  return s;
}
/* getFractional<&Band::X>: lerp with out-of-range bins reading as zero (W#48 9343-9398 etc.) */
static float frac_energy(const Band *b, int B, int low, float fr) {
  float lo = (low >= 0 && low < B) ? b[low].inputEnergy : 0.0f;
  float hi = (low + 1 >= 0 && low + 1 < B) ? b[low + 1].inputEnergy : 0.0f;
  return ((hi - lo) * fr) + lo;
}
static c32 frac_c(const Band *b, int B, int low, float fr, int which) {
  c32 lo = {0, 0}, hi = {0, 0}, r;
  if (low >= 0 && low < B) lo = which ? b[low].prevInput : b[low].input;
  if (low + 1 >= 0 && low + 1 < B) hi = which ? b[low + 1].prevInput : b[low + 1].input;
  r.re = ((hi.re - lo.re) * fr) + lo.re; r.im = ((hi.im - lo.im) * fr) + lo.im;
  return r;
}
static uint32_t minstd_next(uint32_t x) { /* Schrage form as compiled, W#48 9549-9555 */
  uint32_t q = x / 44488u, t = (x - q * 44488u) * 48271u, u = q * 3399u;
  return (t < u ? 2147483647u : 0u) + (t - u);
}
static c32 make_output(const Pred *p, float re, float im) {
  float n2 = (im * im) + (re * re), div;
  if (n2 > 1e-15f) div = n2;
  else { re = p->input.re; im = p->input.im; div = ((re * re) + 1e-15f) + (im * im); }
  float s = sqrtf(p->energy / div);
  c32 o; o.im = s * im; o.re = s * re; return o;
}

/* one spectral step; W#48 8170-9873 */
static void spectrum_step(Engine *e, int s) {
  int C = e->channels, B = e->B;
  float timeFactor = e->timeFactor;
  float fN = (float)(uint32_t)e->N, fH = (float)(uint32_t)e->H, ratio = fN / fH;
  int longStep = trunc_i32(roundf(ratio));
  if (e->newSpectrum) {
    if (s < C) { /* rotate: S1 */
      Band *b = e->bands + (size_t)s * B;
      float twoPiH = fH * 0x1.921fb6p+2f, half = 0.5f / fN;
      float stepA = twoPiH * ((1.5f / fN) - half);
      float sS = m_sinf(stepA), cS = m_cosf(stepA);
      float a0 = twoPiH * half;
      float cr = m_cosf(a0), sr = m_sinf(a0);
      for (int k = 0; k < B; ++k) {
        float oi = b[k].output.im, orr = b[k].output.re;
        b[k].output.im = (oi * cr) + (orr * sr); b[k].output.re = (orr * cr) - (oi * sr);
        float pr = b[k].prevInput.re, pi = b[k].prevInput.im;
        b[k].prevInput.re = (pr * cr) - (pi * sr); b[k].prevInput.im = (pi * cr) + (pr * sr);
        float t = cr * sS;
        cr = (cr * cS) - (sr * sS); sr = t + (sr * cS);
      }
      return;
    }
    s -= C;
  }
  int r; /* index into [formants x3][prelim x C][vertical x8][prevInput] */
  if (e->mapped) {
    if (s <= 2) {
      if (s == 0) {
        memset(e->energy, 0, sizeof(float) * B);
        for (int c = 0; c < C; ++c) {
          Band *b = e->bands + (size_t)c * B;
          for (int k = 0; k < B; ++k) {
            float en = (b[k].input.im * b[k].input.im) + (b[k].input.re * b[k].input.re);
            b[k].inputEnergy = en; e->energy[k] = e->energy[k] + en;
          }
        }
        memcpy(e->smoothed, e->energy, sizeof(float) * B);
        e->smoothCarry = 0.0f;
      } else {
        float slew = 1.0f / ((ratio * 0.5f) + 1.0f);
        e->smoothCarry = smooth_back_fwd(e->smoothed, B, slew, e->smoothCarry);
      }
      return;
    }
    if (s == 3) { /* findPeaks */
      e->nPeaks = 0;
      int k = 0;
      while (k < B) {
        if (!(e->energy[k] <= e->smoothed[k])) {
          float sum = 0.0f, wsum = 0.0f;
          while (k < B) {
            float en = e->energy[k];
            if (en <= e->smoothed[k]) break;
            sum = en + sum; wsum = (en * (float)k) + wsum; ++k;
          }
          float avg = wsum / sum;
          float f = (avg + 0.5f) / fN;
          float o = (map_freq(e, f) * fN) + -0.5f;
          e->peaks[e->nPeaks].input = avg; e->peaks[e->nPeaks].output = o; e->nPeaks++;
        }
        ++k;
      }
      return;
    }
    if (s == 4) { /* updateOutputMap */
      MapPoint *m = e->map;
      if (e->nPeaks == 0) { for (int k = 0; k < B; ++k) { m[k].inputBin = (float)(uint32_t)k; m[k].freqGrad = 1.0f; } return; }
      const Peak *P = e->peaks; int nP = e->nPeaks;
      { int top = trunc_i32(ceilf(P[0].output)); if (top > B) top = B;
        float offs = P[0].input - P[0].output;
        for (int k = 0; k < top; ++k) { m[k].inputBin = offs + (float)(uint32_t)k; m[k].freqGrad = 1.0f; } }
      for (int p = 1; p < nP; ++p) {
        float nOut = P[p].output, pOut = P[p - 1].output;
        int hi = trunc_i32(ceilf(nOut)); if (hi > B) hi = B;
        int lo = trunc_i32(ceilf(pOut)); if (lo < 0) lo = 0;
        if (hi > lo) {
          float pIn = P[p - 1].input;
          float offs = pIn - pOut, inv = 1.0f / (nOut - pOut);
          float delta = (pOut - (nOut + pIn)) + P[p].input;
          float g6 = (inv * delta) * 6.0f;
          for (int k = lo; k < hi; ++k) {
            float kf = (float)(uint32_t)k, rr = (kf - pOut) * inv;
            m[k].freqGrad = ((g6 * rr) * (1.0f - rr)) + 1.0f;
            m[k].inputBin = (offs + kf) + (((rr * rr) * delta) * (3.0f - (rr + rr)));
          }
        }
      }
      { float lIn = P[nP - 1].input, lOut = P[nP - 1].output;
        int lo = trunc_i32(lOut); if (lo < 0) lo = 0;
        float offs = lIn - lOut;
        for (int k = lo; k < B; ++k) { m[k].inputBin = offs + (float)(uint32_t)k; m[k].freqGrad = 1.0f; } }
      return;
    }
    r = s - 5;
  } else {
    if (s == 0) {
      for (int c = 0; c < C; ++c) {
        Band *b = e->bands + (size_t)c * B;
        for (int k = 0; k < B; ++k) b[k].inputEnergy = (b[k].input.im * b[k].input.im) + (b[k].input.re * b[k].input.re);
      }
      for (int k = 0; k < B; ++k) { e->map[k].inputBin = (float)(uint32_t)k; e->map[k].freqGrad = 1.0f; }
      return;
    }
    r = s - 1;
  }
  if (e->formants) {
    if (r < 3) {
      float *fm = e->formantMetric;
      if (r == 0) {
        memset(fm, 0, sizeof(float) * (B + 2));
        for (int c = 0; c < C; ++c) {
          const Band *b = e->bands + (size_t)c * B;
          for (int k = 0; k < B; ++k) fm[k] = fm[k] + b[k].inputEnergy;
        }
        float base = e->formantBaseFreq;
        e->formantBaseBin = (base * fN) + -0.5f;
        if (!(base > 0.0f)) { /* auto-detect: top three local maxima + harmonic fix-ups */
          int i1 = 0, i2 = 0, i3 = 0; /* i1 strongest */
          for (int i = 1; i <= B - 2 && B >= 3; ++i) {
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
            int d = abs(i1 - i2);
            if (d > i1 / 8 && d < (i1 * 7) / 8) i1 = i1 % d;
            if (!((double)fm[i3] <= (dtop * 0.01))) {
              d = abs(i1 - i3);
              if (d > i1 / 8 && d < (i1 * 7) / 8) i1 = i1 % d;
            }
          }
          float w = e->freqEstimateWeight;
          float nw = (float)(((double)(top - w) * 0.25) + (double)w);
          e->freqEstimateWeight = nw;
          float ww = e->freqEstimateWeighted;
          ww = (float)(((double)((top * (float)i1) - ww) * 0.25) + (double)ww);
          e->freqEstimateWeighted = ww;
          e->formantBaseBin = ww / (nw + 1e-30f);
        }
        for (int k = 0; k < B; ++k) fm[k] = sqrtf(fm[k]);
      } else if (r == 1) {
        float slew = (float)(1.0 / (((double)e->formantBaseBin * 0.5) + 1.0));
        float st = smooth_back_fwd(fm, B, slew, 0.0f);
        smooth_back_fwd(fm, B, slew, st);
      } else {
        for (int k = 0; k < B; ++k) {
          float f = ((float)k + 0.5f) / fN;
          if (e->formantCompensation == 1) f = map_freq(e, f);
          float metric = fm[k], lim = e->freqTonalityLimit;
          float g = e->invFormantMultiplier * f;
          float fi = (g > lim) ? (((1.0f - e->formantMultiplier) * lim) + f) : g;
          float pos = (fi * fN) + -0.5f, env = 0.0f;
          if (!(pos < 0.0f)) {
            float fB = (float)B, x = (fB < pos) ? fB : pos, fl = floorf(x), fr = x - fl;
            int idx = trunc_i32(fl);
            float lo = fm[idx];
            env = (fr * (fm[idx + 1] - lo)) + lo;
          }
          float g2 = env / (metric + 1e-30f); g2 = g2 * g2;
          for (int c = 0; c < C; ++c) { Band *b = e->bands + (size_t)c * B + k; b->inputEnergy = g2 * b->inputEnergy; }
        }
      }
      return;
    }
    r -= 3;
  }
  if (r < C) { /* preliminary prediction for channel r: S5 */
    Band *b = e->bands + (size_t)r * B; Pred *pr = e->preds + (size_t)r * B;
    for (int k = 0; k < B; ++k) {
      float ib = e->map[k].inputBin, fl = floorf(ib);
      int low = trunc_i32(fl); float fr = ib - fl;
      float prevE = pr[k].energy, grad = e->map[k].freqGrad;
      float en = frac_energy(b, B, low, fr) * (grad > 0.0f ? grad : 0.0f);
      pr[k].energy = en;
      c32 in = frac_c(b, B, low, fr, 0); pr[k].input = in;
      c32 pv = frac_c(b, B, low, fr, 1);
      float tIm = (pv.re * in.im) - (pv.im * in.re), tRe = (pv.im * in.im) + (pv.re * in.re);
      float oRe = b[k].output.re, oIm = b[k].output.im;
      float div = ((en > prevE) ? en : prevE) + 1e-15f;
      b[k].output.im = ((tIm * oRe) + (tRe * oIm)) / div;
      b[k].output.re = ((tRe * oRe) - (tIm * oIm)) / div;
    }
    return;
  }
  r -= C;
  if (r <= 7) { /* main vertical prediction over bins [B*r/8, B*(r+1)/8): S6 */
    int k0 = (int)(((uint32_t)B * (uint32_t)r) >> 3), k1 = (int)(((uint32_t)B * (uint32_t)(r + 1)) >> 3);
    float tf = timeFactor < 0.5f ? 0.5f : timeFactor;
    float rlo = ((tf > 2.0f) ? 4.0f : 0.0f) - tf, rscale = (tf - rlo) * 0x1p-31f;
    int randomTF = !(tf <= 2.0f);
    float fLong = (float)longStep;
    for (int k = k0; k < k1; ++k) {
      int mc = 0; float me = e->preds[k].energy;
      for (int c = 1; c < C; ++c) { float en = e->preds[(size_t)c * B + k].energy; if (en > me) { me = en; mc = c; } }
      Band *b = e->bands + (size_t)mc * B; Pred *pr = e->preds + (size_t)mc * B;
      float pRe = pr[k].input.re, pIm = pr[k].input.im;
      float phRe = 0.0f, phIm = 0.0f;
      if (k > 0) {
        float ib = e->map[k].inputBin, btf = tf;
        if (randomTF) { e->rng = minstd_next(e->rng); btf = (rscale * (float)(uint32_t)(e->rng - 1u)) + rlo; }
        float x = ib - btf; int low = trunc_i32(floorf(x)); float fr = x - (float)low;
        c32 d = frac_c(b, B, low, fr, 0);
        float tIm = (d.re * pIm) - (d.im * pRe), tRe = (d.im * pIm) + (d.re * pRe);
        float oRe = b[k - 1].output.re, oIm = b[k - 1].output.im;
        phIm = (tIm * oRe) + (tRe * oIm); phRe = (tRe * oRe) - (tIm * oIm);
        if (k >= longStep) {
          x = ib - (btf * fLong); low = trunc_i32(floorf(x)); fr = x - (float)low;
          d = frac_c(b, B, low, fr, 0);
          tRe = (d.im * pIm) + (d.re * pRe); tIm = (d.re * pIm) - (d.im * pRe);
          oIm = b[k - longStep].output.im; oRe = b[k - longStep].output.re;
          phIm = ((tRe * oIm) + phIm) + (tIm * oRe);
          phRe = ((tRe * oRe) + phRe) - (oIm * tIm);
        }
      }
      if (k < B - 1) {
        float btf = tf;
        if (randomTF) { e->rng = minstd_next(e->rng); btf = (rscale * (float)(uint32_t)(e->rng - 1u)) + rlo; }
        float x = e->map[k + 1].inputBin - btf; int low = trunc_i32(floorf(x)); float fr = x - (float)low;
        c32 d = frac_c(b, B, low, fr, 0);
        float uIm = pr[k + 1].input.im, uRe = pr[k + 1].input.re;
        float tRe = (d.im * uIm) + (d.re * uRe), tIm = (d.re * uIm) - (d.im * uRe);
        float oIm = b[k + 1].output.im, oRe = b[k + 1].output.re;
        phIm = ((tRe * oIm) - (tIm * oRe)) + phIm;
        phRe = ((tRe * oRe) + phRe) + (tIm * oIm);
        if (k < B - longStep) {
          int kk = k + longStep;
          x = e->map[kk].inputBin - (btf * fLong); low = trunc_i32(floorf(x)); fr = x - (float)low;
          d = frac_c(b, B, low, fr, 0);
          uIm = pr[kk].input.im; uRe = pr[kk].input.re;
          tRe = (d.im * uIm) + (d.re * uRe); tIm = (d.re * uIm) - (d.im * uRe);
          oIm = b[kk].output.im; oRe = b[kk].output.re;
          phIm = ((tRe * oIm) + phIm) - (oRe * tIm);
          phRe = ((tRe * oRe) + phRe) + (tIm * oIm);
        }
      }
      c32 o = make_output(&pr[k], phRe, phIm);
      b[k].output = o;
      for (int c = 0; c < C; ++c) {
        if (c == mc) continue;
        Pred *cp = e->preds + (size_t)c * B + k;
        float tIm = (pRe * cp->input.im) - (pIm * cp->input.re), tRe = (pIm * cp->input.im) + (pRe * cp->input.re);
        float qIm = (tIm * o.re) + (tRe * o.im), qRe = (tRe * o.re) - (tIm * o.im);
        e->bands[(size_t)c * B + k].output = make_output(cp, qRe, qIm);
      }
    }
    return;
  }
  if (r == 8 && e->newSpectrum) {
    for (size_t i = 0; i < (size_t)C * B; ++i) e->bands[i].prevInput = e->bands[i].input;
  }
}

/* synthesis step for channel c (c == 0 also adds the window product); W#48 9906-10932 */
static void synth_step(Engine *e, int c) {
  int L = e->L, N = e->N, off = L >> 1;
  OutState *o = &e->out;
  if (c == 0) {
    o->sinceSynth = 0;
    float fN = (float)(uint32_t)N;
    for (int i = 0; i < L; ++i) { int p = (o->pos + i) % L; o->wp[p] = ((e->win[i] * fN) * e->win[i]) + o->wp[p]; }
  }
  rfft_inverse(&e->fft, e->spectrum + (size_t)c * e->B, e->timeBuf);
  float *ring = o->ring + (size_t)c * L; const float *t = e->timeBuf;
  for (int i = 0; i < off; ++i) { int p = (o->pos + i) % L; ring[p] = ring[p] - (t[N - off + i] * e->win[i]); }
  for (int i = off; i < L; ++i) { int p = (o->pos + i) % L; ring[p] = ring[p] + (t[i - off] * e->win[i]); }
}

static void run_step(Engine *e, int step) {
  int C = e->channels, B = e->B, s = step;
  if (e->newSpectrum) {
    if (e->reanalysePrev) {
      if (s < C) { analyse(e, &e->stashIn, s, e->H); return; }
      if (s == C) {
        for (int c = 0; c < C; ++c) for (int k = 0; k < B; ++k) e->bands[(size_t)c * B + k].prevInput = e->spectrum[(size_t)c * B + k];
        return;
      }
      s -= C + 1;
    }
    if (s < C) { analyse(e, &e->stashIn, s, 0); return; }
    if (s == C) {
      for (int c = 0; c < C; ++c) for (int k = 0; k < B; ++k) e->bands[(size_t)c * B + k].input = e->spectrum[(size_t)c * B + k];
      return;
    }
    s -= C + 1;
  }
  if (s < e->spectrumSteps) { spectrum_step(e, s); return; }
  if (s == e->spectrumSteps) {
    for (int c = 0; c < C; ++c) for (int k = 0; k < B; ++k) e->spectrum[(size_t)c * B + k] = e->bands[(size_t)c * B + k].output;
    return;
  }
  s -= e->spectrumSteps + 1;
  if (s < C) synth_step(e, s);
}

/* W#48 process(inputSamples, outputSamples) */
static void engine_process(Engine *e, int nIn, int nOut) {
  int C = e->channels, L = e->L, H = e->H;
  e->prevCopiedInput = 0;
  float total = 0.0f;
  int loud = 0;
  if (C > 0 && nIn > 0) {
    for (int c = 0; c < C; ++c) { const float *x = e->buffers + (size_t)e->bufLen * c; for (int i = 0; i < nIn; ++i) total = (x[i] * x[i]) + total; }
    loud = total >= 1e-15f;
  }
  if (!loud) {
    if (e->silenceCounter >= ((uint32_t)L << 1)) {
      if (e->silenceFirst) {
        e->silenceFirst = 0; reset_block_process(e);
        memset(e->bands, 0, sizeof(Band) * (size_t)C * e->B);
      }
      float *outs = e->buffers + (size_t)e->bufLen * C;
      if (nIn > 0) {
        for (int i = 0, j = 0; i < nOut; ++i) { for (int c = 0; c < C; ++c) outs[(size_t)e->bufLen * c + i] = e->buffers[(size_t)e->bufLen * c + j]; j = (j + 1 != nIn) ? j + 1 : 0; }
      } else {
        for (int c = 0; c < C; ++c) for (int i = 0; i < nOut; ++i) outs[(size_t)e->bufLen * c + i] = 0.0f;
      }
      copy_input(e, nIn);
      return;
    }
    e->silenceCounter += (uint32_t)nIn;
  } else { e->silenceFirst = 1; e->silenceCounter = 0; }

  float invOut = 1.0f / (float)(uint32_t)nOut, fIn = (float)nIn;
  for (int idx = 0; idx < nOut; ++idx) {
    int toStep = 0;
    if (e->samplesSinceLast >= (uint32_t)H) {
      e->step = 0; e->samplesSinceLast = 0; e->steps = 0;
      int inputOffset = trunc_i32(roundf(((float)(uint32_t)idx * fIn) * invOut));
      int prev = e->prevInputOffset; e->prevInputOffset = inputOffset;
      copy_input(e, inputOffset);
      e->stashIn.pos = e->in.pos; memcpy(e->stashIn.ring, e->in.ring, sizeof(float) * C * e->inLen);
      if (e->split) {
        e->stashOut.pos = e->out.pos; e->stashOut.sinceSynth = e->out.sinceSynth;
        memcpy(e->stashOut.ring, e->out.ring, sizeof(float) * C * L); memcpy(e->stashOut.wp, e->out.wp, sizeof(float) * L);
        move_output(e, &e->out, H);
      }
      int inputInterval = inputOffset - prev;
      e->newSpectrum = e->didSeek || inputInterval > 0;
      e->mapped = e->freqMultiplier != 1.0f;
      if (e->newSpectrum) {
        e->reanalysePrev = e->didSeek || abs(inputInterval - H) > 1;
        if (e->reanalysePrev) e->steps += C + 1;
        e->steps += C + 1;
      }
      e->formants = (e->formantMultiplier == 1.0f) ? (e->formantCompensation && e->mapped) : 1;
      if (e->didSeek) e->timeFactor = e->seekTimeFactor;
      else { float fi = (float)inputInterval; e->timeFactor = (float)(uint32_t)H / (fi > 1.0f ? fi : 1.0f); }
      e->didSeek = 0;
      e->spectrumSteps = C + (e->newSpectrum ? 10 : 9) + (e->newSpectrum ? C : 0) + (e->mapped ? 4 : 0) + (e->formants ? 3 : 0);
      e->steps = C + (e->spectrumSteps + e->steps) + 1;
      toStep = e->steps;
    }
    if (e->split) {
      float pr = (((float)(uint32_t)e->steps + 0.999f) * (float)(uint32_t)(e->samplesSinceLast + 1u)) / (float)(uint32_t)H;
      uint32_t lim = (pr < 4294967296.0f && pr >= 0.0f) ? (uint32_t)pr : 0u;
      toStep = lim > (uint32_t)e->steps ? e->steps : (int)lim;
    }
    while (e->step < toStep) { int st = e->step++; run_step(e, st); }
    e->samplesSinceLast += 1u;
    if (e->split) { OutState t = e->out; e->out = e->stashOut; e->stashOut = t; }
    {
      OutState *o = &e->out; int p = o->pos % L;
      float *outs = e->buffers + (size_t)e->bufLen * C;
      for (int c = 0; c < C; ++c) outs[(size_t)e->bufLen * c + idx] = o->ring[(size_t)c * L + p] / o->wp[p];
      move_output(e, o, 1);
    }
    if (e->split) { OutState t = e->out; e->out = e->stashOut; e->stashOut = t; }
  }
  copy_input(e, nIn);
  e->prevInputOffset -= nIn;
}

/* W#49 seek(inputSamples, playbackRate) */
static void engine_seek(Engine *e, int n, double rate) {
  int cap = e->L + e->H, C = e->channels;
  memset(e->tmpBuf, 0, sizeof(float) * cap);
  int start = n - cap; if (start < 0) start = 0;
  float energy = 0.0f;
  for (int c = 0; c < C; ++c) {
    const float *x = e->buffers + (size_t)e->bufLen * c;
    for (int i = start; i < n; ++i) { float v = x[i]; e->tmpBuf[i + (cap - n)] = v; energy = (v * v) + energy; }
    float *ring = e->in.ring + (size_t)c * e->inLen;
    for (int i = 0; i < cap; ++i) ring[(e->in.pos + i) % e->inLen] = e->tmpBuf[i];
  }
  e->in.pos = (e->in.pos + cap) % e->inLen;
  if (energy >= 1e-15f) { e->silenceCounter = 0; e->silenceFirst = 1; }
  e->didSeek = 1;
  double dH = (double)(uint32_t)e->H;
  e->seekTimeFactor = (float)(((rate * dH) > 1.0) ? (1.0 / rate) : dH);
}

/* ---------------------------------------------------------------------------------------------- C ABI */
#define API __attribute__((visibility("default")))
API Engine *so_new(uint32_t seed) {
  Engine *e = (Engine *)calloc(1, sizeof(Engine));
  e->freqMultiplier = 1.0f; e->freqTonalityLimit = 0.5f; e->formantMultiplier = 1.0f; e->invFormantMultiplier = 1.0f;
  uint32_t s = seed % 2147483647u; e->rng = s <= 1u ? 1u : s; /* W#26 ctor: minstd_rand(random_device()) */
  e->prevInputOffset = -1; e->silenceFirst = 1; /* note: the blob zero-inits silenceFirst? see test */
  cur = e; return e;
}
API void so_free(Engine *e) { if (cur == e) cur = 0; engine_free_bufs(e); free(e->buffers); free(e); }
API void so_select(Engine *e) { cur = e; }
API float *so_buffers(Engine *e) { return e->buffers; }
API float *so_setBuffers(int ch, int len) { /* W#40 */
  free(cur->buffers); cur->buffers = (float *)calloc((size_t)2 * ch * len, 4); cur->bufCh = ch; cur->bufLen = len;
  return cur->buffers;
}
API int so_blockSamples(void) { return cur->L; }
API int so_intervalSamples(void) { return cur->H; }
API int so_inputLatency(void) { return cur->L - (cur->L >> 1); }
API int so_outputLatency(void) { return (cur->L >> 1) + cur->H * cur->split; }
API void so_configure(int ch, int L, int H, int split) { engine_configure(cur, ch, L, H, split); }
API void so_presetDefault(int ch, float sr) { double d = (double)sr; engine_configure(cur, ch, (int)(d * 0.12), (int)(d * 0.03), 0); }
API void so_presetCheaper(int ch, float sr) { double d = (double)sr; engine_configure(cur, ch, (int)(d * 0.1), (int)(d * 0.04), 1); }
API void so_reset(void) { /* W#59 */
  Engine *e = cur;
  stft_reset(e, 0.1f); stash_all(e);
  e->prevInputOffset = -1;
  memset(e->bands, 0, sizeof(Band) * (size_t)e->channels * e->B);
  reset_block_process(e);
  e->freqEstimateWeighted = e->freqEstimateWeight = 0.0f; e->didSeek = 0; e->silenceCounter = 0;
}
API void so_setTransposeFactor(float m, float tl) { /* W#54 */
  cur->freqMultiplier = m; cur->freqTonalityLimit = (tl <= 0.0f) ? 1.0f : (tl / sqrtf(m));
}
API void so_setTransposeSemitones(float st, float tl) { /* W#53: exp2 in f64 (musl W#30), demoted */
  so_setTransposeFactor((float)exp2((double)(st * 0x1.555556p-4f)), tl);
}
API void so_setFormantFactor(float m, int comp) { cur->formantMultiplier = m; cur->formantCompensation = comp & 0xff; cur->invFormantMultiplier = 1.0f / m; }
API void so_setFormantSemitones(float st, int comp) { so_setFormantFactor((float)exp2((double)(st * 0x1.555556p-4f)), comp); }
API void so_setFormantBase(float f) { cur->formantBaseFreq = f; }
API void so_seek(int n, double rate) { engine_seek(cur, n, rate); }
API void so_process(int nIn, int nOut) { engine_process(cur, nIn, nOut); }
/* W#46 flush(outputSamples).  Exported, never called by the reference JS (app/SignalsmithStretch.mjs:479 is its only
 * mention); restated from the bytecode: the window-product ring becomes its own running maximum (in ring order from the
 * read position), up to L samples are read WITHOUT consuming them, what is left of the ring after those (at most
 * outputSamples more) is folded back reversed and subtracted from the end of the output, then stft.reset(0.1) and every
 * Band's prevInput and output are cleared (input and inputEnergy stay; blockProcess and prevInputOffset too). */
static void engine_flush(Engine *e, int nOut) {
  int C = e->channels, L = e->L;
  OutState *o = &e->out;
  float m = 0.0f;
  for (int i = 0; i < L; ++i) { int p = (o->pos + i) % L; float x = o->wp[p]; m = (m > x) ? m : x; o->wp[p] = m; }
  if (C > 0) {
    int n1 = (L < nOut) ? L : nOut, rest = L - n1, n2 = (rest < nOut) ? rest : nOut;
    for (int c = 0; c < C; ++c) {
      const float *ring = o->ring + (size_t)c * L;
      float *out = e->buffers + (size_t)e->bufLen * (C + c);
      for (int i = 0; i < n1; ++i) { int p = (o->pos + i) % L; out[i] = ring[p] / o->wp[p]; }
      for (int i = 0; i < n2; ++i) { int p = (o->pos + n1 + i) % L; out[nOut - 1 - i] = out[nOut - 1 - i] - (ring[p] / o->wp[p]); }
    }
  }
  stft_reset(e, 0.1f);
  for (size_t i = 0; i < (size_t)C * e->B; ++i) { e->bands[i].prevInput.re = e->bands[i].prevInput.im = 0.0f; e->bands[i].output.re = e->bands[i].output.im = 0.0f; }
}
API void so_flush(int nOut) { engine_flush(cur, nOut); }
/* inspection helpers for the stage-by-stage tests */
API const float *so_window(void) { return cur->win; }
API const c32 *so_spectrum(void) { return cur->spectrum; }
API const Band *so_bands(void) { return cur->bands; }
API int so_fftSize(void) { return cur->N; }
