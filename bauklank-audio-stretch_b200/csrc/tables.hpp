// Host-side constant tables for one engine configuration (window, FFT twiddles, spectral rotation,
// window-product normalisation).  Product code: computed here from first principles, never read from oracle/.
//
// Every formula follows the arithmetic of the reference's WASM engine (blob at app/SignalsmithStretch.mjs:265;
// W#n = wasm function index) so that the f32 constants are bit-identical to the ones the
// reference builds at configure() time:
//   fast-size rule + buffer sizes  W#25 configure          Kaiser window + perfect reconstruction  W#36
//   FFT twiddles / plan            W#38                     window products                         W#22, W#23
//   twiddle sin/cos                musl sinf/cosf as linked into the blob (W#9-W#12), |x| <= 9*pi/4 branches
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace bs {

struct alignas(8) cf { float re, im; };   // 8-byte aligned: one 64-bit load/store per complex value

// ---- musl single-precision sin/cos kernels (polynomials evaluated in double, rounded once) ----
inline float sindf_(double x) {
  double z = x * x, s = x * z;
  return (float)(((s * (z * z)) * ((z * 0x1.6cd878c3b46a7p-19) + -0x1.a00f9e2cae774p-13)) +
                 ((s * ((z * 0x1.11110896efbb2p-7) + -0x1.5555554cbac77p-3)) + x));
}
inline float cosdf_(double x) {
  double z = x * x, w = z * z;
  return (float)(((z * w) * ((z * 0x1.99342e0ee5069p-16) + -0x1.6c087e80f1e27p-10)) +
                 ((w * 0x1.55553e1053a42p-5) + ((z * -0x1.ffffffd0c5e81p-2) + 1.0)));
}
constexpr double kPio2 = 0x1.921fb54442d18p+0, kPi = 0x1.921fb54442d18p+1, kPi3o2 = 0x1.2d97c7f3321d2p+2,
                 kPi2 = 0x1.921fb54442d18p+2;
inline uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
[[noreturn]] inline void die(const char *what) { std::fprintf(stderr, "bauklank_stretch: %s\n", what); std::abort(); }
inline float sinf_ref(float x) {
  uint32_t b = bits(x), ix = b & 0x7fffffffu; bool neg = (int32_t)b < 0;
  if (ix <= 0x3f490fdau) return ix < 0x39800000u ? x : sindf_((double)x);
  if (ix <= 0x407b53d1u) {
    if (ix <= 0x4016cbe3u) return neg ? -cosdf_((double)x + kPio2) : cosdf_((double)x + -kPio2);
    return sindf_(-((neg ? kPi : -kPi) + (double)x));
  }
  if (ix <= 0x40e231d5u) {
    if (ix <= 0x40afeddfu) return neg ? cosdf_((double)x + kPi3o2) : -cosdf_((double)x + -kPi3o2);
    return sindf_((neg ? kPi2 : -kPi2) + (double)x);
  }
  die("sin argument outside the table range");
}
inline float cosf_ref(float x) {
  uint32_t b = bits(x), ix = b & 0x7fffffffu; bool neg = (int32_t)b < 0;
  if (ix <= 0x3f490fdau) return ix < 0x39800000u ? 1.0f : cosdf_((double)x);
  if (ix <= 0x407b53d1u) {
    if (ix >= 0x4016cbe4u) return -cosdf_((neg ? kPi : -kPi) + (double)x);
    return neg ? sindf_((double)x + kPio2) : sindf_(kPio2 - (double)x);
  }
  if (ix <= 0x40e231d5u) {
    if (ix >= 0x40afede0u) return cosdf_((neg ? kPi2 : -kPi2) + (double)x);
    return neg ? sindf_(-kPi3o2 - (double)x) : sindf_((double)x + -kPi3o2);
  }
  die("cos argument outside the table range");
}
inline cf polar1(float a) { cf c; c.im = sinf_ref(a); c.re = cosf_ref(a); return c; }

// ---- configuration (W#25) ----
struct Geometry {
  int C = 0, L = 0, H = 0, N = 0, B = 0, M = 0, inner = 0, outer = 0, split = 0, longStep = 0;
  int inLat = 0, outLat = 0;
};
inline int fft_size_for_block(int L) {
  uint32_t x = ((((uint32_t)L + 1u) >> 1) + 1u) >> 1, lim = x >= 16u ? 16u : x, p = 1u, q;
  do { q = p; p = q << 1; } while (q < lim);
  do { p = q; q = p << 1; } while ((p << 3) < x);
  uint32_t m = (p + x - 1u) / p;
  return (int)((p * (m == 7u ? 8u : m)) << 2);
}
inline Geometry make_geometry(int C, int L, int H, int split) {
  Geometry g; g.C = C; g.L = L; g.H = H; g.split = split ? 1 : 0;
  g.N = fft_size_for_block(L); g.B = g.M = g.N >> 1;
  int inner = 1, outer = g.M;
  while (outer > 1 && !(outer & 1)) { outer >>= 1; inner <<= 1; }
  g.inner = inner; g.outer = outer;
  float r = roundf((float)(uint32_t)g.N / (float)(uint32_t)H);   // longVerticalStep, W#48 8179-8191
  g.longStep = std::fabs(r) < 2147483648.0f ? (int)r : INT32_MIN;
  g.inLat = L - (L >> 1); g.outLat = (L >> 1) + H * g.split;      // W#61 / W#60
  return g;
}

struct Tables {
  std::vector<float> win;        // [L] analysis == synthesis window
  std::vector<cf> tw;            // [3*inner/4] radix-4 twiddles
  std::vector<float> otr, oti;   // [inner*(outer-1)] outer twiddles, split
  std::vector<cf> untangle;      // [N/4+1]
  std::vector<cf> rot;           // [M] half-bin rotations
  std::vector<cf> specRot;       // [B] running rotation applied to output/prevInput on a new spectrum (S1)
  std::vector<float> packTab;    // [M][4] per packed sample pair: signed window coefficients of its two samples, half-bin rotation
                                 //        (empty unless the window halves fall on pair boundaries) -- fft_fast.cuh
  std::vector<cf> otw;           // [inner][outer-1] the outer twiddles again, the ones of a bin side by side
  std::vector<float> wpStart;    // window-product denominators for the first wpStart.size() output samples
  std::vector<float> wpSteady;   // [H] periodic part after that
  std::vector<float> winProd;    // [L] (w*N)*w
};

// Kaiser window with the heuristic bandwidth and forced perfect reconstruction, W#36 (all f64, stored f32)
inline void make_window(const Geometry &g, std::vector<float> &w) {
  int L = g.L, H = g.H;
  w.assign(L, 0.f);
  double dL = (double)L, bw = dL / (double)H, t = bw + 3.0;
  double heur = (8.0 / (t * t)) + bw, rem = 3.0 - bw;
  bw = heur + ((rem < 0.0 ? 0.0 : rem) * 0.25);
  bw = bw < 2.0 ? 2.0 : bw;
  double beta = std::sqrt(((bw * bw) * 0.25) + -1.0) * kPi, b2 = beta * beta;
  double term = 1.0, sum = 0.0, k = 0.0;
  do { sum = sum + term; k = k + 1.0; term = (b2 * term) / ((k * k) * 4.0); } while (term > 1e-4);
  double invI0 = 1.0 / sum, invL = 1.0 / dL;
  for (int i = 0; i < L; ++i) {
    double r = ((double)(uint32_t)((i << 1) | 1) * invL) + -1.0;
    double arg = std::sqrt(1.0 - (r * r)) * beta, a2 = arg * arg;
    term = 1.0; sum = 0.0; k = 0.0;
    do { sum = sum + term; k = k + 1.0; term = (a2 * term) / ((k * k) * 4.0); } while (term > 1e-4);
    w[i] = (float)(sum * invI0);
  }
  for (int i = 0; i < H && i < L; ++i) {
    double s = 0.0;
    for (int j = i; j < L; j += H) { float v = w[j]; s = s + (double)(v * v); }
    double f = 1.0 / std::sqrt(s);
    for (int j = i; j < L; j += H) w[j] = (float)((double)w[j] * f);
  }
}

inline void make_fft_tables(const Geometry &g, Tables &T) {
  int N = g.N, M = g.M, inner = g.inner, outer = g.outer;
  if (outer != 1 && outer != 3 && outer != 5) die("unsupported FFT factorisation (outer factor must be 1, 3 or 5)");
  int ntw = (3 * inner) >> 2;
  T.tw.assign(ntw > 0 ? ntw : 1, cf{1.f, 0.f});
  double rinner = 1.0 / (double)inner;
  for (int i = 0; i < ntw; ++i) T.tw[i] = polar1((float)(((double)i * -kPi2) * rinner));
  int no = inner * (outer - 1);
  T.otr.assign(no > 0 ? no : 1, 0.f); T.oti.assign(no > 0 ? no : 1, 0.f);
  for (int i = 0; i < inner && outer >= 2; ++i) {
    double a0 = (double)i * -kPi2;
    for (int s = 1; s < outer; ++s) {
      cf c = polar1((float)((a0 * (double)s) / ((double)inner * (double)outer)));
      T.otr[i + inner * (s - 1)] = c.re; T.oti[i + inner * (s - 1)] = c.im;
    }
  }
  double rN = 1.0 / (double)N;
  T.untangle.resize((N >> 2) + 1);
  for (size_t i = 0; i < T.untangle.size(); ++i) T.untangle[i] = polar1((float)(((((double)i * -kPi2) + -kPi) * rN) + -kPio2));
  T.rot.resize(M);
  for (int i = 0; i < M; ++i)
    T.rot[i] = polar1(((N & 2) && i == M - 1) ? (float)(((double)i * -kPi2) / (double)N) : (float)(((double)i * -kPi2) * rN));
  T.otw.assign(no > 0 ? no : 1, cf{0.f, 0.f});
  for (int i = 0; i < inner && outer >= 2; ++i)
    for (int s = 1; s < outer; ++s) T.otw[(size_t)i * (outer - 1) + (s - 1)] = cf{T.otr[i + inner * (s - 1)], T.oti[i + inner * (s - 1)]};
}
// The window and the half-bin rotation as the specialised kernels consume them: packed pair j holds window samples i, i+1
// with i = 2j + off (second half of the window, j < jA) or 2j - cStart (first half, sign flipped: the half-bin shift), and
// is multiplied by rot[j].  x * (-w) == -(x * w) exactly, so the sign lives in the table.  (W#35 / W#48 9986-10932)
inline void make_pack_table(const Geometry &g, Tables &T) {
  const int L = g.L, N = g.N, M = g.M, off = L >> 1, nA = L - off, cStart = N - off;
  T.packTab.clear();
  if (((nA | cStart | off) & 1) != 0) return;
  T.packTab.assign((size_t)4 * M, 0.f);
  for (int j = 0; j < M; ++j) {
    const int n = 2 * j;
    const bool inA = n < nA, inC = n >= cStart;
    float w0 = 0.f, w1 = 0.f;
    if (inA) { w0 = T.win[n + off]; w1 = T.win[n + off + 1]; }
    else if (inC) { w0 = -T.win[n - cStart]; w1 = -T.win[n - cStart + 1]; }
    T.packTab[4 * j] = w0; T.packTab[4 * j + 1] = w1; T.packTab[4 * j + 2] = T.rot[j].re; T.packTab[4 * j + 3] = T.rot[j].im;
  }
}

// S1: rot = polar(2pi*H*0.5/N), rotStep = polar(2pi*H*(1.5/N - 0.5/N)); running f32 product over bins (W#48 8193-8229)
inline void make_spec_rot(const Geometry &g, std::vector<cf> &sr) {
  float fN = (float)(uint32_t)g.N, fH = (float)(uint32_t)g.H;
  float twoPiH = fH * 0x1.921fb6p+2f, half = 0.5f / fN;
  float stepA = twoPiH * ((1.5f / fN) - half);
  float sS = sinf_ref(stepA), cS = cosf_ref(stepA);
  float a0 = twoPiH * half;
  float c = cosf_ref(a0), s = sinf_ref(a0);
  sr.resize(g.B);
  for (int k = 0; k < g.B; ++k) {
    sr[k].re = c; sr[k].im = s;
    float t = c * sS;
    c = (c * cS) - (s * sS); s = t + (s * cS);
  }
}

// Window-product ring (W#22 reset(0.1), W#23 moveOutput, addWindowProduct in W#48 9913-9981), simulated exactly.
// Output sample n is divided by wp(n): wpStart[n] for n < wpStart.size(), else wpSteady[(n - wpStart.size()) % H].
inline void make_window_products(const Geometry &g, Tables &T) {
  int L = g.L, H = g.H;
  float fN = (float)(uint32_t)g.N;
  T.winProd.resize(L);
  for (int i = 0; i < L; ++i) T.winProd[i] = (T.win[i] * fN) * T.win[i];
  std::vector<float> wp(L, 0.f);
  auto move = [&](std::vector<float> &v, int &pos, int n) { for (int i = 0; i < n; ++i) v[(pos + i) % L] = 1e-30f; pos = (pos + n) % L; };
  for (int i = 0; i < L; ++i) wp[i] = T.winProd[i] + wp[i];
  for (int i = L - H - 1; i >= 0; --i) wp[i] = wp[i] + wp[i + H];
  for (int i = 0; i < L; ++i) wp[i] = (wp[i] * 0.1f) + 1e-30f;
  int pos = 0; move(wp, pos, H);
  // run the per-block protocol until the denominators are H-periodic
  int nStart = ((3 * L + 2 * H) / H + 2) * H, nTotal = nStart + 2 * H;
  std::vector<float> den(nTotal);
  std::vector<float> stash; int spos = 0;
  for (int n = 0; n < nTotal; ++n) {
    if (n % H == 0) {
      if (g.split) { stash = wp; spos = pos; move(wp, pos, H); }
      for (int i = 0; i < L; ++i) { int p = (pos + i) % L; wp[p] = T.winProd[i] + wp[p]; }
    }
    if (g.split) { den[n] = stash[spos % L]; move(stash, spos, 1); }
    else { den[n] = wp[pos % L]; move(wp, pos, 1); }
  }
  for (int j = 0; j < H; ++j)
    if (bits(den[nStart + j]) != bits(den[nStart + H + j]) || bits(den[nStart + j]) != bits(den[nStart - H + j]))
      die("window products did not become periodic");
  T.wpStart.assign(den.begin(), den.begin() + nStart);
  T.wpSteady.assign(den.begin() + nStart, den.begin() + nStart + H);
}

inline void make_tables(const Geometry &g, Tables &T) {
  make_window(g, T.win);
  make_fft_tables(g, T);
  make_pack_table(g, T);
  make_spec_rot(g, T.specRot);
  make_window_products(g, T);
}

}  // namespace bs
