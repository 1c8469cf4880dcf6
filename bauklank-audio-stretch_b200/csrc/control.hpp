// Host-side control plane of one stream: an exact, data-free emulation of the reference engine's block scheduling
// (W#48 process control, W#49 seek, setters W#50-54) and of the worklet's time map arithmetic
// (app/SignalsmithStretch.mjs:826-954).  It produces, for every analysis/synthesis block, a BlockRec (flags,
// timeFactor, the parameter values each spectral step will see) and two Window descriptors saying where in the
// stream's clip the "current" and "previous" analysis windows come from.  That table is what makes frame/hop
// indexing bit-exact on the GPU: all integer and rounding rules live here, in f32/f64 exactly as the reference.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "tables.hpp"

namespace bs {

enum : uint32_t { kNew = 1u, kMapped = 2u, kFormants = 4u, kFormantComp = 8u, kValid = 16u };

// One block of one stream, as consumed by the spectral kernel (32 bytes).
struct BlockRec {
  uint32_t flags;
  float timeFactor;
  float pkMult, pkLimit;    // freqMultiplier / freqTonalityLimit seen by findPeaks (mapped step 3)
  float fmBaseFreq;         // formantBaseFreq seen by formant step 0
  float fmMult, fmInv;      // formantMultiplier / invFormantMultiplier seen by formant step 2
  float fmFreqMult;         // freqMultiplier seen by formant step 2 (fmLimit below)
};
struct BlockRec2 {
  float fmLimit;
  int lastNew;        // stream-relative index of the block whose "current" spectrum is this block's input (-1: none yet)
  uint32_t rngSkip;   // blocks before this one that drew from the RNG (timeFactor > 2): each consumed 2*bands-2 draws
  uint32_t zeroBelow; // compat shim only (flush inside the vertical-prediction steps): Band.output of the bins below reads as 0
};
// sample i (0 <= i < L) of a window = (lo <= i < hi) ? clip[ch][start + i] : 0
struct Window { long long start; int lo, hi; };

struct Params {  // what the setters store (W#50-54)
  float freqMult = 1.f, tonLimit = 0.5f, formantMult = 1.f, formantInv = 1.f, formantBase = 0.f;
  int formantComp = 0;
  void setTransposeFactor(float m, float tl) { freqMult = m; tonLimit = (tl <= 0.f) ? 1.f : (tl / sqrtf(m)); }
  void setTransposeSemitones(float st, float tl) { setTransposeFactor((float)exp2((double)(st * 0x1.555556p-4f)), tl); }
  void setFormantFactor(float m, int comp) { formantMult = m; formantComp = comp & 0xff; formantInv = 1.f / m; }
  void setFormantSemitones(float st, int comp) { setFormantFactor((float)exp2((double)(st * 0x1.555556p-4f)), comp); }
  void setFormantBase(float f) { formantBase = f; }
};

// Mirrors one timeMap entry of the worklet (app/SignalsmithStretch.mjs:587-600), all JS numbers = double.
struct Segment {
  double output, input, rate;
  double semitones, tonalityHz, formantSemitones, formantBaseHz, loopStart, loopEnd;
  int active, formantCompensation;
  // not NaN: the driver calls setTransposeFactor / setFormantFactor (exports "q" / "s", app/SignalsmithStretch.mjs:472,474)
  // with this multiplier instead of the semitone setters
  double transposeFactor = NAN, formantFactor = NAN;
};

// A seek whose loudness the plan had to assume (see KioskPlanner): sum of squares over clip[c][start, start+count) of every
// channel, accumulated like W#49 does, must be >= 1e-15 -- checked on the device after the run (bsb_gate_events).
struct SeekWatch { long long start; int count; };
struct StreamPlan {
  std::vector<BlockRec> blocks;
  std::vector<BlockRec2> blocks2;
  std::vector<Window> windows;  // 2 per block: cur, prev
  long long nOut = 0;
  long long nLive = -1;         // output samples before the silence gate closed for good (the rest is zeros); -1: all of them
  std::vector<SeekWatch> watch;
  const char *error = nullptr;  // a drive the batched path cannot express
};

class Control {
 public:
  explicit Control(const Geometry &g) : g_(g) {}
  Params p;

  // W#49 seek: only its control effects (the ring provenance is the caller's business)
  void seek(double rate) {
    didSeek_ = true;
    double dH = (double)(uint32_t)g_.H;
    seekTF_ = (float)(((rate * dH) > 1.0) ? (1.0 / rate) : dH);
  }

  // Emulates `count` output samples of one process(nIn, nOut) call starting at output index idx0, during which the
  // parameters `p` do not change.  onStart(idx, inputOffset, inputInterval, rec index) is called when a block starts.
  struct NoSpan { void operator()(int, int, uint32_t) const {} };
  template <class StartFn>
  void run(StreamPlan &plan, int idx0, int count, int nIn, int nOut, StartFn &&onStart) { run(plan, idx0, count, nIn, nOut, onStart, NoSpan()); }
  // onSpan(idx, span, since): `span` output samples starting at call index idx, `since` samples after the block start
  template <class StartFn, class SpanFn>
  void run(StreamPlan &plan, int idx0, int count, int nIn, int nOut, StartFn &&onStart, SpanFn &&onSpan) {
    const int H = g_.H, C = g_.C;
    float invOut = 1.0f / (float)(uint32_t)nOut, fIn = (float)nIn;
    int idx = idx0, end = idx0 + count;
    while (idx < end) {
      if (since_ >= (uint32_t)H) {
        step_ = 0; since_ = 0; steps_ = 0;
        float r = roundf(((float)(uint32_t)idx * fIn) * invOut);
        int inputOffset = std::fabs(r) < 2147483648.0f ? (int)r : INT32_MIN;
        int interval = inputOffset - prevInputOffset_;
        prevInputOffset_ = inputOffset;
        bool isNew = didSeek_ || interval > 0;
        bool mapped = p.freqMult != 1.0f;
        bool rean = false;
        if (isNew) {
          rean = didSeek_ || std::abs(interval - H) > 1;
          if (rean) steps_ += C + 1;
          steps_ += C + 1;
        }
        bool formants = (p.formantMult == 1.0f) ? (p.formantComp && mapped) : true;
        float tf;
        if (didSeek_) tf = seekTF_;
        else { float fi = (float)interval; tf = (float)(uint32_t)H / (fi > 1.0f ? fi : 1.0f); }
        didSeek_ = false;
        int specSteps = C + (isNew ? 10 : 9) + (isNew ? C : 0) + (mapped ? 4 : 0) + (formants ? 3 : 0);
        int analysisSteps = steps_;
        steps_ = C + (specSteps + steps_) + 1;
        // absolute indices of the parameter-reading steps
        int s0 = analysisSteps + (isNew ? C : 0);
        stepPeaks_ = mapped ? s0 + 3 : -1;
        int f0 = s0 + (mapped ? 5 : 1);
        stepFm0_ = formants ? f0 : -1;
        stepFm2_ = formants ? f0 + 2 : -1;
        stepS6_ = f0 + (formants ? 3 : 0) + C;   // after the C preliminary-prediction steps
        stepSyn_ = analysisSteps + specSteps + 1;
        curNew_ = isNew; curRean_ = rean;
        BlockRec rec{};
        rec.flags = kValid | (isNew ? kNew : 0u) | (mapped ? kMapped : 0u) | (formants ? kFormants : 0u);
        rec.timeFactor = tf;
        plan.blocks.push_back(rec);
        if (isNew) lastNew_ = (int)plan.blocks.size() - 1;
        BlockRec2 r2{}; r2.lastNew = lastNew_; r2.rngSkip = rngBlocks_;
        if (!((tf < 0.5f ? 0.5f : tf) <= 2.0f)) ++rngBlocks_;   // W#48 9470: random time factors only when tf > 2
        plan.blocks2.push_back(r2);
        plan.windows.push_back(Window{0, 0, 0});
        plan.windows.push_back(Window{0, 0, 0});
        cur_ = (long long)plan.blocks.size() - 1;
        onStart(idx, inputOffset, interval, rean, isNew, cur_);
      }
      // samples until the next block start or the end of this span
      int span = std::min<long long>(end - idx, (long long)H - (long long)since_);
      int toStep;
      if (g_.split) {
        float pr = (((float)(uint32_t)steps_ + 0.999f) * (float)(uint32_t)(since_ + (uint32_t)span)) / (float)(uint32_t)H;
        uint32_t lim = (pr < 4294967296.0f && pr >= 0.0f) ? (uint32_t)pr : 0u;
        toStep = lim > (uint32_t)steps_ ? steps_ : (int)lim;
      } else {
        toStep = steps_;
      }
      if (cur_ >= 0 && toStep > step_) snapshot(plan, step_, toStep);
      if (toStep > step_) step_ = toStep;
      onSpan(idx, span, since_);
      since_ += (uint32_t)span;
      idx += span;
    }
  }
  void endCall(int nIn) { prevInputOffset_ -= nIn; }
  // How far the block under way has got (split computation spreads its steps over the interval): vertical-prediction
  // steps (of 8) and synthesis steps (of C) already executed.  Without split computation every step ran at once.
  void progress(int &nS6, int &nSyn) const {
    nS6 = std::min(8, std::max(0, step_ - stepS6_)); nSyn = std::min(g_.C, std::max(0, step_ - stepSyn_));
  }
  int stepsDone() const { return step_; }
  int stepS6() const { return stepS6_; }
  int stepFinal() const { return stepSyn_ - 2; }   // "prevInput = input" (only blocks with a new spectrum have it)
  bool curIsNew() const { return curNew_; }
  bool curReanalysesPrev() const { return curRean_; }
  // `blockProcess = {}` of the silence gate (W#48 7842-7845) and of reset() (W#59)
  void resetBlockProcess() { since_ = 0xffffffffu; steps_ = 0; step_ = 0; cur_ = -1; }
  void resetAll() { resetBlockProcess(); prevInputOffset_ = -1; didSeek_ = false; }
  // One-block-at-a-time callers (the 18-call shim runs for as long as the kiosk is up): drop the records of the blocks before
  // the one under way; it becomes block 0 of the plan.  Call from onStart, where the current block is the plan's last.
  void keepOnlyCurrent(StreamPlan &plan) {
    if (cur_ <= 0) return;
    plan.blocks.erase(plan.blocks.begin(), plan.blocks.begin() + cur_);
    plan.blocks2.erase(plan.blocks2.begin(), plan.blocks2.begin() + cur_);
    plan.windows.erase(plan.windows.begin(), plan.windows.begin() + 2 * cur_);
    lastNew_ = lastNew_ >= cur_ ? (int)(lastNew_ - cur_) : -1;
    cur_ = 0;
  }

 private:
  void snapshot(StreamPlan &plan, int from, int to) {
    BlockRec &r = plan.blocks[cur_];
    if (stepPeaks_ >= from && stepPeaks_ < to) { r.pkMult = p.freqMult; r.pkLimit = p.tonLimit; }
    if (stepFm0_ >= from && stepFm0_ < to) r.fmBaseFreq = p.formantBase;
    if (stepFm2_ >= from && stepFm2_ < to) {
      r.fmMult = p.formantMult; r.fmInv = p.formantInv; r.fmFreqMult = p.freqMult;
      plan.blocks2[cur_].fmLimit = p.tonLimit;
      if (p.formantComp == 1) r.flags |= kFormantComp;
    }
  }
  Geometry g_;
  uint32_t since_ = 0xffffffffu;
  int steps_ = 0, step_ = 0;
  int prevInputOffset_ = -1;
  bool didSeek_ = false;
  float seekTF_ = 0.f;
  int stepPeaks_ = -1, stepFm0_ = -1, stepFm2_ = -1, stepS6_ = 0, stepSyn_ = 0;
  bool curNew_ = false, curRean_ = false;
  long long cur_ = -1;
  int lastNew_ = -1;
  uint32_t rngBlocks_ = 0;
};

inline void apply_segment_params(Params &p, const Segment &s, double sampleRate) {
  // WasmProcessor.process, app/SignalsmithStretch.mjs:847-849 (JS doubles -> f32 at the wasm call boundary)
  if (std::isnan(s.transposeFactor)) p.setTransposeSemitones((float)s.semitones, (float)(s.tonalityHz / sampleRate));
  else p.setTransposeFactor((float)s.transposeFactor, (float)(s.tonalityHz / sampleRate));
  if (std::isnan(s.formantFactor)) p.setFormantSemitones((float)s.formantSemitones, s.formantCompensation ? 1 : 0);
  else p.setFormantFactor((float)s.formantFactor, s.formantCompensation ? 1 : 0);
  p.setFormantBase((float)(s.formantBaseHz / sampleRate));
}

inline Window clip_window(long long start, int L, int validFrom, long long clipLen, long long clipFrom = 0, int validTo = 1 << 30) {
  Window w; w.start = start;
  long long lo = std::max<long long>(validFrom, clipFrom - start), hi = std::min<long long>(std::min(L, validTo), clipLen - start);
  if (lo < 0) lo = 0;
  if (hi < lo) hi = lo;
  w.lo = (int)lo; w.hi = (int)hi;
  return w;
}

// One stream under the worklet's drive, quantum by quantum (app/SignalsmithStretch.mjs:840-943).  Two kinds of quantum:
//   active    buffer playback (:883-943): `_seek(bufferLength, rate)` then `_process(0, q)`;
//   inactive  (:861-869): the input buffer is zeroed and `_process(q, q)` runs on it.
// The engine's STFT input ring is modelled by provenance: the most recent seek left clip[end-n, end) (zeros before it),
// every inactive quantum appends zeros, and an analysis window is the L ring samples ending 0 (current) or H (previous)
// samples before the write position -- always "a run of clip samples with zeros on either side", which is what a Window
// says.  The silence gate of process() (W#48 7838-7943) is followed too: `silenceCounter` grows by q per inactive
// quantum, a seek over audio resets it (W#49), and once it has reached 2L every further call without loud input returns
// zeros without touching the output ring.  The batched path follows the reference up to that point and, if no seek re-opens the gate,
// to the end (all zeros); a stream that is started again after the gate has closed re-arms its block phase at an
// arbitrary output position and is left to the 18-call shim (`error` is set).
// Data dependence: whether a seek is "loud" depends on the audio.  A seek whose buffer holds clip samples is planned as
// loud; where that matters (the counter was not zero) the seek is put on the watch list and checked on the device.
class KioskPlanner {
 public:
  KioskPlanner(const Geometry &g, StreamPlan &plan) : ctl(g), g_(g), plan_(plan) {}
  Control ctl;   // the caller applies the quantum's setter values to ctl.p before each call of quantum()

  // active quantum: `end` = Math.round(inputTime * sampleRate), the stored audio covers clip samples [validStart, validEnd)
  // inactive quantum: rate / end / valid range unused
  void quantum(int q, bool active, double rate, long long end, long long validStart, long long validEnd) {
    if (plan_.error) return;
    const int L = g_.L, H = g_.H, cap = L + H, bufLen = g_.inLat + g_.outLat;
    if (active) {
      const int n = std::min(bufLen, cap);   // samples of pre-roll that seek() keeps
      const long long a0 = std::max<long long>(end - n, validStart), a1 = std::min<long long>(end, validEnd);
      if (a1 > a0) {   // the buffer holds audio: planned as loud (silenceCounter = 0, W#49)
        if (counter_ > 0) plan_.watch.push_back(SeekWatch{a0, (int)(a1 - a0)});
        counter_ = 0;
      }
      ringEnd_ = end; ringN_ = n; ringLo_ = validStart; ringHi_ = validEnd; zeros_ = 0;
      ctl.seek(rate);
    }
    const bool gated = counter_ >= ((uint32_t)L << 1);   // process(): no loud input (there is none, or it is zeros)
    if (gated) {
      if (plan_.nLive < 0) plan_.nLive = pos_;
      if (!active) { pos_ += q; return; }
      // (reached only if a silent seek follows a closed gate: stays closed)
      pos_ += q; return;
    }
    if (plan_.nLive >= 0) { plan_.error = "a stream that plays again after the silence gate has closed (more than 2 blocks of inactive output) re-arms its block phase: drive it through the 18-call engine"; return; }
    const int nIn = active ? 0 : q;
    ctl.run(plan_, 0, q, nIn, q, [&](int idx, int, int, bool rean, bool isNew, long long bi) {
      if (!isNew) return;  // spectra are carried over (flag kNew clear)
      // zeros appended since the seek: the whole inactive quanta so far, plus copyInput(inputOffset) of this call (W#24)
      const long long z = std::min<long long>(zeros_ + (active ? 0 : idx), (long long)cap + 1);
      // a window ending p samples before the write position: sample i is ring history position i - p - L (0 = write position);
      // history [-z, 0) is zeros, [-z-n, -z) is clip[ringEnd - n, ringEnd), older is zeros
      auto win = [&](int p) { return clip_window(ringEnd_ + z - p - L, L, (int)std::max<long long>(0, (long long)p + L - z - ringN_), ringHi_, ringLo_, (int)std::max<long long>(0, (long long)p + L - z)); };
      const Window cur = win(0);
      plan_.windows[2 * bi + 0] = cur;
      plan_.windows[2 * bi + 1] = rean ? win(H) : lastCur_;
      lastCur_ = cur;
    });
    ctl.endCall(nIn);
    if (!active) { counter_ += (uint32_t)q; zeros_ += q; }
    pos_ += q;
  }

 private:
  Geometry g_;
  StreamPlan &plan_;
  long long pos_ = 0;
  uint32_t counter_ = 0;                 // silenceCounter
  long long ringEnd_ = 0, ringLo_ = 0, ringHi_ = 0, zeros_ = 0; int ringN_ = 0;
  Window lastCur_{0, 0, 0};
};

// Buffer-playback drive of the worklet (app/SignalsmithStretch.mjs:883-943): every quantum `_seek(bufferLength, rate)`
// then `_process(0, q)`; inactive segments run `_process(q, q)` on silence (:861-869).  `segs` is the (already ordered)
// time map; currentTime = k*quantum/sampleRate.
inline void plan_kiosk(const Geometry &g, double sampleRate, int quantum, long long nOut, long long clipLen,
                       const Segment *segs, int nSegs, StreamPlan &plan) {
  KioskPlanner kp(g, plan);
  const double inLatS = (double)g.inLat / sampleRate, outLatS = (double)g.outLat / sampleRate;
  std::vector<Segment> tm(segs, segs + nSegs);
  size_t si = 0;
  plan.nOut = nOut;
  long long pos = 0;
  for (long long k = 0; pos < nOut; ++k) {
    int q = (int)std::min<long long>(quantum, nOut - pos);
    double currentTime = (double)(k * quantum) / sampleRate;
    double outputTime = currentTime + outLatS;
    while (si + 1 < tm.size() && tm[si + 1].output <= outputTime) ++si;
    Segment &seg = tm[si];
    apply_segment_params(kp.ctl.p, seg, sampleRate);
    long long end = 0;
    if (seg.active) {
      double inputTime = seg.input + (outputTime - seg.output) * seg.rate;
      double loopLength = seg.loopEnd - seg.loopStart;
      if (loopLength > 0 && inputTime >= seg.loopEnd) { seg.input -= loopLength; inputTime -= loopLength; }
      inputTime += inLatS;
      end = (long long)std::floor(inputTime * sampleRate + 0.5);  // Math.round
    }
    kp.quantum(q, seg.active != 0, seg.rate, end, 0, clipLen);
    pos += q;
  }
}

// The same drive from a per-quantum table (what a host-side mirror of the worklet's time map resolves a control trace
// into): for quantum k the f32 values the three setters receive (app/SignalsmithStretch.mjs:847-849), the segment's rate
// and Math.round(inputTime * sampleRate) (:897), and the range of clip samples the buffer store holds at that moment.
struct Quantum {
  double rate; long long inputSamplesEnd, validStart, validEnd;
  float semitones, tonalityLimit, formantSemitones, formantBase;
  int formantCompensation, active;
  float transposeFactor = NAN, formantFactor = NAN;   // not NaN: setTransposeFactor / setFormantFactor instead of the semitone setters
};
inline void plan_kiosk_table(const Geometry &g, int quantum, long long nOut, const Quantum *qs, long long nQ, StreamPlan &plan) {
  KioskPlanner kp(g, plan);
  plan.nOut = nOut;
  long long pos = 0;
  for (long long k = 0; pos < nOut && k < nQ; ++k) {
    const Quantum &qq = qs[k];
    const int q = (int)std::min<long long>(quantum, nOut - pos);
    if (std::isnan(qq.transposeFactor)) kp.ctl.p.setTransposeSemitones(qq.semitones, qq.tonalityLimit);
    else kp.ctl.p.setTransposeFactor(qq.transposeFactor, qq.tonalityLimit);
    if (std::isnan(qq.formantFactor)) kp.ctl.p.setFormantSemitones(qq.formantSemitones, qq.formantCompensation ? 1 : 0);
    else kp.ctl.p.setFormantFactor(qq.formantFactor, qq.formantCompensation ? 1 : 0);
    kp.ctl.p.setFormantBase(qq.formantBase);
    kp.quantum(q, qq.active != 0, qq.rate, qq.inputSamplesEnd, qq.validStart, qq.validEnd);
    pos += q;
  }
}

// The worklet's control plane in front of the same drive: a trace of schedule() calls, each made
// between two render quanta, edits the time map exactly like remoteMethods.schedule (app/SignalsmithStretch.mjs:656-701:
// later segments popped, active / rate / semitones / loop bounds inherited from the latest one, input extrapolated at its
// rate -- 0 if it was inactive --, then every segment whose successor starts by the NEW segment's output time shifted out,
// so a segment scheduled ahead is current at once); every quantum then looks its segment up like process() does (:840-844).
// All JS numbers are doubles.  One stored audio buffer covering the whole clip, as the kiosk app adds it (app/multi/app.mjs:369-376).
struct TraceEvent {
  long long quantum;      // applied before this render quantum (currentTime = quantum * q / sampleRate)
  double outputTime;      // NaN: currentTime
  double input;           // NaN: extrapolated
  double rate, semitones, loopStart, loopEnd;   // NaN: inherited
  double tonalityHz, formantSemitones, formantBaseHz;   // not inherited by the reference: must be given
  int active;             // < 0: inherited
  int formantCompensation;
  double transposeFactor = NAN, formantFactor = NAN;
};
inline void plan_kiosk_trace(const Geometry &g, double sampleRate, int quantum, long long nOut, long long clipLen,
                             const TraceEvent *ev, long long nEv, StreamPlan &plan) {
  KioskPlanner kp(g, plan);
  const double inLatS = (double)g.inLat / sampleRate, outLatS = (double)g.outLat / sampleRate;
  std::vector<Segment> tm;
  {   // the worklet's initial segment (:587-600)
    Segment s0{}; s0.output = 0; s0.input = 0; s0.rate = 1; s0.semitones = 0; s0.tonalityHz = 8000; s0.formantSemitones = 0; s0.formantBaseHz = 0;
    s0.loopStart = 0; s0.loopEnd = 0; s0.active = 0; s0.formantCompensation = 0;
    tm.push_back(s0);
  }
  plan.nOut = nOut;
  long long pos = 0, ei = 0;
  for (long long k = 0; pos < nOut; ++k) {
    const double currentTime = (double)(k * quantum) / sampleRate;
    for (; ei < nEv && ev[ei].quantum <= k; ++ei) {
      const TraceEvent &e = ev[ei];
      if (std::isnan(e.tonalityHz) || std::isnan(e.formantSemitones) || std::isnan(e.formantBaseHz) || e.formantCompensation < 0) {
        plan.error = "schedule() without tonalityHz / formantSemitones / formantCompensation / formantBaseHz: the reference does not inherit them and would hand NaN to the engine";
        return;
      }
      const double outputTime = std::isnan(e.outputTime) ? currentTime : e.outputTime;
      Segment latest = tm.back();
      while (!tm.empty() && tm.back().output >= outputTime) { latest = tm.back(); tm.pop_back(); }
      Segment s{};
      s.active = e.active < 0 ? latest.active : (e.active ? 1 : 0);
      s.output = outputTime;
      s.rate = std::isnan(e.rate) ? latest.rate : e.rate;
      s.semitones = std::isnan(e.semitones) ? latest.semitones : e.semitones;
      s.loopStart = std::isnan(e.loopStart) ? latest.loopStart : e.loopStart;
      s.loopEnd = std::isnan(e.loopEnd) ? latest.loopEnd : e.loopEnd;
      s.tonalityHz = e.tonalityHz; s.formantSemitones = e.formantSemitones; s.formantBaseHz = e.formantBaseHz;
      s.formantCompensation = e.formantCompensation ? 1 : 0;
      s.transposeFactor = e.transposeFactor; s.formantFactor = e.formantFactor;
      s.input = std::isnan(e.input) ? latest.input + (s.output - latest.output) * (latest.active ? latest.rate : 0.0) : e.input;
      tm.push_back(s);
      size_t drop = 0;
      while (tm.size() - drop > 1 && tm[drop + 1].output <= outputTime) ++drop;
      if (drop) tm.erase(tm.begin(), tm.begin() + drop);
    }
    const int q = (int)std::min<long long>(quantum, nOut - pos);
    const double outputTime = currentTime + outLatS;
    {
      size_t drop = 0;
      while (tm.size() - drop > 1 && tm[drop + 1].output <= outputTime) ++drop;
      if (drop) tm.erase(tm.begin(), tm.begin() + drop);
    }
    Segment &seg = tm[0];
    apply_segment_params(kp.ctl.p, seg, sampleRate);
    long long end = 0;
    if (seg.active) {
      double inputTime = seg.input + (outputTime - seg.output) * seg.rate;
      const double loopLength = seg.loopEnd - seg.loopStart;
      if (loopLength > 0 && inputTime >= seg.loopEnd) { seg.input -= loopLength; inputTime -= loopLength; }
      inputTime += inLatS;
      end = (long long)std::floor(inputTime * sampleRate + 0.5);
    }
    kp.quantum(q, seg.active != 0, seg.rate, end, 0, clipLen);
    if (plan.error) return;
    pos += q;
  }
}

// Streaming drive (live-input branch generalised): process(nIn, nOut) over a contiguous input, no seek.
// Parameter segments are keyed by the output time of the first sample of a call.
inline void plan_stream(const Geometry &g, double sampleRate, int nIn, int nOut, long long nCalls, long long clipLen,
                        const Segment *segs, int nSegs, StreamPlan &plan) {
  Control ctl(g);
  const int L = g.L, H = g.H;
  size_t si = 0;
  plan.nOut = nCalls * nOut;
  long long lastEnd = 0;   // stream position of the most recently analysed "current" window end
  bool haveLast = false;
  for (long long k = 0; k < nCalls; ++k) {
    double t = (double)(k * nOut) / sampleRate;
    while ((int)si + 1 < nSegs && segs[si + 1].output <= t) ++si;
    apply_segment_params(ctl.p, segs[si], sampleRate);
    long long base = k * nIn;
    ctl.run(plan, 0, nOut, nIn, nOut, [&](int, int inputOffset, int, bool rean, bool isNew, long long bi) {
      if (!isNew) return;
      long long P = base + inputOffset;
      long long prevEnd = (rean || !haveLast) ? P - H : lastEnd;
      plan.windows[2 * bi + 0] = clip_window(P - L, L, 0, clipLen);
      plan.windows[2 * bi + 1] = clip_window(prevEnd - L, L, 0, clipLen);
      lastEnd = P; haveLast = true;
    });
    ctl.endCall(nIn);
  }
}

}  // namespace bs
