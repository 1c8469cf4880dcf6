"""Shared parity-case table and drivers (test infrastructure).

A *case* is a dict describing a seeded synthetic clip, an engine configuration and a drive.  The same case runs on
* any engine with the reference's 18-call surface (oracle/_ref blob, oracle C port, the GPU compat shim
  ``StretchEngine``) through ``run_case`` -- a re-enactment of ``WasmProcessor.process``
  (app/SignalsmithStretch.mjs:826-954: time-map lookup :840-844, setters :847-849, buffer fill :883-931,
  ``_seek`` + ``_process`` :935-936, live input :870-882);
* the batched device path (``BatchStretch``) through ``run_case_batch``.
"""
import math

import numpy as np

from oracle import refdrive

T8 = 8000.0  # worklet default tonalityHz (app/SignalsmithStretch.mjs:594)


def seg(output=0.0, input=0.0, rate=1.0, semitones=0.0, tonality_hz=T8, formant_semitones=0.0,
        formant_compensation=False, formant_base_hz=0.0, loop_start=0.0, loop_end=0.0, active=True,
        transpose_factor=None, formant_factor=None):
    """One time-map entry.  ``transpose_factor`` / ``formant_factor``: the driver uses the engine's setTransposeFactor /
    setFormantFactor exports (app/SignalsmithStretch.mjs:472,474) instead of the semitone setters."""
    return dict(output=output, input=input, rate=rate, semitones=semitones, tonality_hz=tonality_hz,
                formant_semitones=formant_semitones, formant_compensation=formant_compensation,
                formant_base_hz=formant_base_hz, loop_start=loop_start, loop_end=loop_end, active=active,
                transpose_factor=transpose_factor, formant_factor=formant_factor)


def schedule(segments, output, **changes):
    """remoteMethods.schedule (app/SignalsmithStretch.mjs:656-701) for a segment appended at ``output``: fields are
    inherited from the latest segment and ``input`` is derived by linear extrapolation unless given."""
    last = segments[-1]
    new = dict(last)
    new.update(changes)
    new["output"] = output
    if "input" not in changes:     # (:679-682: an inactive segment's input stands still)
        new["input"] = last["input"] + (output - last["output"]) * (last["rate"] if last.get("active", True) else 0.0)
    segments.append(new)
    return segments


def make_clip(spec):
    kind = spec[0]
    if kind == "survey":
        return refdrive.survey_clip(spec[1])
    if kind == "gated":     # (kind, n, ((start, stop), ...)): the survey clip with digital silence over the given spans
        x = refdrive.survey_clip(spec[1]).copy()
        for a, b in spec[2]:
            x[:, a:b] = 0.0
        return x
    if kind == "sweep":     # (kind, seconds, sr, channels)
        return refdrive.sweep_clip(spec[1], spec[2], spec[3])
    if kind == "noise":     # (kind, n, channels, seed, amplitude)
        rng = np.random.default_rng(spec[3])
        return (spec[4] * rng.standard_normal((spec[2], spec[1]))).astype(np.float32)
    if kind == "tones":     # (kind, n, channels, sr): a few steady partials per channel + a little noise
        n, ch, sr = spec[1], spec[2], spec[3]
        t = np.arange(n) / sr
        rng = np.random.default_rng(7)
        out = np.zeros((ch, n), np.float32)
        for c in range(ch):
            f0 = 110.0 * (1.0 + 0.31 * c)
            y = sum(0.3 / (h + 1) * np.sin(2 * math.pi * f0 * (h + 1) * t + c) for h in range(6))
            out[c] = (y + 0.01 * rng.standard_normal(n)).astype(np.float32)
        return out
    raise ValueError(kind)


def _setup(engine, case, channels):
    blk = case.get("block")
    if blk:
        return refdrive.setup(engine, channels, case["sr"], block=blk[0], interval=blk[1], split=blk[2])
    return refdrive.setup(engine, channels, case["sr"], preset=case.get("preset", "default"))


def _apply(engine, s, sr):
    if s.get("transpose_factor") is None:
        engine.setTransposeSemitones(s["semitones"], s["tonality_hz"] / sr)
    else:
        engine.setTransposeFactor(s["transpose_factor"], s["tonality_hz"] / sr)
    if s.get("formant_factor") is None:
        engine.setFormantSemitones(s["formant_semitones"], bool(s["formant_compensation"]))
    else:
        engine.setFormantFactor(s["formant_factor"], bool(s["formant_compensation"]))
    engine.setFormantBase(s["formant_base_hz"] / sr)


def kiosk_run(engine, clip, case):
    """Buffer playback (:883-943) over a time map of segments."""
    sr, n_out, quantum = case["sr"], case["n_out"], case.get("quantum", 128)
    ch = clip.shape[0]
    buf_len = _setup(engine, case, ch)
    in_lat_s, out_lat_s = engine.inputLatency() / sr, engine.outputLatency() / sr
    tm = [dict(s) for s in case["segments"]]
    out = np.zeros((ch, n_out), np.float32)
    pos = k = 0
    while pos < n_out:
        q = min(quantum, n_out - pos)
        output_time = (k * quantum) / sr + out_lat_s
        while len(tm) > 1 and tm[1]["output"] <= output_time:
            tm.pop(0)
        s = tm[0]
        _apply(engine, s, sr)
        if not s.get("active", True):        # :861-869: zeroed input, process(q, q)
            ins, _ = engine.io_views()
            ins[:, :q] = 0
            engine.process(q, q)
            _, outs = engine.io_views()
            out[:, pos:pos + q] = outs[:, :q]
            pos += q
            k += 1
            continue
        input_time = s["input"] + (output_time - s["output"]) * s["rate"]
        loop_len = s["loop_end"] - s["loop_start"]
        if loop_len > 0 and input_time >= s["loop_end"]:
            s["input"] -= loop_len
            input_time -= loop_len
        input_time += in_lat_s
        end = refdrive.js_round(input_time * sr)
        ins, _ = engine.io_views()
        refdrive.kiosk_fill(ins, clip, end)
        engine.seek(buf_len, s["rate"])
        engine.process(0, q)
        _, outs = engine.io_views()
        out[:, pos:pos + q] = outs[:, :q]
        pos += q
        k += 1
    return out


def stream_run(engine, clip, case):
    """Live-input branch (:870-882) generalised to process(n_in, n_out); segments keyed by the call's output time."""
    sr, n_in, n_out = case["sr"], case["n_in"], case["n_out"]
    ch = clip.shape[0]
    _setup(engine, case, ch)
    calls = case.get("n_calls", clip.shape[1] // n_in)
    flush_at, n_flush = case.get("flush", (None, 0))        # flush(n_flush) after that many calls; its samples are kept
    engine.setBuffers(ch, max(n_in, n_out, n_flush))
    segs = case["segments"]
    si = 0
    parts = []
    for k in range(calls):
        if k == flush_at:
            _, outs = engine.io_views()
            outs[:] = 0.25                                      # flush subtracts from what the buffer holds
            engine.flush(n_flush)
            _, outs = engine.io_views()
            parts.append(outs[:, :n_flush].copy())
        t = (k * n_out) / sr
        while si + 1 < len(segs) and segs[si + 1]["output"] <= t:
            si += 1
        _apply(engine, segs[si], sr)
        ins, _ = engine.io_views()
        ins[:, :n_in] = clip[:, k * n_in:(k + 1) * n_in]
        engine.process(n_in, n_out)
        _, outs = engine.io_views()
        parts.append(outs[:, :n_out].copy())
    return np.concatenate(parts, axis=1)


def run_case(engine, case, clip=None):
    clip = make_clip(case["clip"]) if clip is None else clip
    return kiosk_run(engine, clip, case) if case["drive"] == "kiosk" else stream_run(engine, clip, case)


def batch_drive(bs, case, clip_len):
    segs = [bs.segment(**s) for s in case["segments"]]
    seed = case.get("seed", 1)
    if case["drive"] == "kiosk":
        return bs.KioskDrive(case["n_out"], segs, quantum=case.get("quantum", 128), seed=seed)
    calls = case.get("n_calls", clip_len // case["n_in"])
    return bs.StreamingDrive(case["n_in"], case["n_out"], calls, segs, seed=seed)


def make_batch(bs, case, channels, lib=None):
    blk = case.get("block")
    if blk:
        return bs.BatchStretch(channels, case["sr"], block_samples=blk[0], interval_samples=blk[1],
                               split_computation=bool(blk[2]), lib=lib)
    return bs.BatchStretch(channels, case["sr"], preset=case.get("preset", "default"), lib=lib)


def run_cases_batch(bs, cases, lib=None, device=None, chunk_blocks=0):
    """Run several cases that share one configuration as ONE batch.  ``device`` None = numpy host arrays (only
    valid with the host-emulation build); else a torch device."""
    clips = [make_clip(c["clip"]) for c in cases]
    eng = make_batch(bs, cases[0], clips[0].shape[0], lib=lib)
    if device is not None:
        import torch
        dclips = [torch.from_numpy(x).to(device).contiguous() for x in clips]
    else:
        dclips = [np.ascontiguousarray(x) for x in clips]
    outs = eng.plan(dclips, [batch_drive(bs, c, x.shape[1]) for c, x in zip(cases, clips)], chunk_blocks=chunk_blocks)
    eng.run()
    if device is not None:
        import torch
        torch.cuda.synchronize()
        res = [o.cpu().numpy() for o in outs]
    else:
        res = [np.array(o) for o in outs]
    eng.close()
    return res


def expected_gate_events(clip, n_in, n_calls, block_samples):
    """How many process(n_in, .) calls the reference short-circuits (W#48 7838-7943): a call is silent when the f32 sum
    of squares of its input, accumulated channel by channel in sample order, is below 1e-15; once 2L silent samples
    have been counted every further silent call is gated."""
    counter = fired = 0
    for k in range(n_calls):
        total = np.float32(0.0)
        for c in range(clip.shape[0]):
            x = clip[c, k * n_in:(k + 1) * n_in]
            if np.any(x):
                for v in x:
                    total = np.float32(np.float32(v * v) + total)
        if total >= np.float32(1e-15):
            counter = 0
        elif counter >= 2 * block_samples:
            fired += 1
        else:
            counter += n_in
    return fired


def config_key(case):
    return (case.get("preset", "default"), tuple(case["block"]) if case.get("block") else None, case["sr"],
            case["clip"][2] if case["clip"][0] in ("noise", "tones") else (case["clip"][3] if case["clip"][0] == "sweep" else 2))


def compare(a, b):
    """(bit_identical, max_abs_err, min_snr_db over channels)"""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    same = a.shape == b.shape and bool((a.view(np.uint32) == b.view(np.uint32)).all())
    d = a.astype(np.float64) - b.astype(np.float64)
    err = float(np.abs(d).max()) if d.size else 0.0
    snr = float("inf")
    for c in range(a.shape[0]):
        num = float((b[c].astype(np.float64) ** 2).sum())
        den = float((d[c] ** 2).sum())
        if den > 0:
            snr = min(snr, 10 * math.log10(max(num, 1e-300) / den))
    return same, err, snr


# ------------------------------------------------------------------------------------------------ the table
def _sweep_segments(seconds_out, sr, n_steps, rate0=0.5, rate1=2.0, st0=-12, st1=12, **kw):
    """BASELINE config 2: rate geometric rate0 -> rate1, transpose st0 -> st1 in integer steps, rescheduled n_steps
    times over the output duration (each = one schedule() call)."""
    segs = [seg(rate=rate0, semitones=float(st0), **kw)]
    for i in range(1, n_steps):
        u = i / (n_steps - 1)
        schedule(segs, seconds_out * i / n_steps, rate=rate0 * (rate1 / rate0) ** u,
                 semitones=float(round(st0 + (st1 - st0) * u)))
    return segs


SURVEY = ("survey", 96000)
CASES = {
    # SURVEY.md section 8c known answers (transpose set every quantum with tonalityLimit 8000/48000)
    "KA1": dict(drive="stream", clip=SURVEY, sr=48000, n_in=512, n_out=512, preset="default", segments=[seg()]),
    "KA2": dict(drive="stream", clip=SURVEY, sr=48000, n_in=512, n_out=512, preset="default", segments=[seg(semitones=7.0)]),
    "KA3": dict(drive="kiosk", clip=SURVEY, sr=48000, n_out=96000, preset="cheaper", segments=[seg()]),
    "KA4": dict(drive="kiosk", clip=SURVEY, sr=48000, n_out=128000, preset="cheaper", segments=[seg(rate=0.75, semitones=5.0)]),
    "KA5": dict(drive="kiosk", clip=SURVEY, sr=48000, n_out=48000, preset="default", segments=[seg(rate=2.0, semitones=-12.0)]),
    "KA6": dict(drive="kiosk", clip=SURVEY, sr=48000, n_out=192000, preset="default",
                segments=[seg(rate=0.5, semitones=3.0, formant_semitones=4.0, formant_compensation=True, formant_base_hz=200.0)]),
    # more drives, minted with the same translated blob (tests/golden/make_golden.py)
    "rng_low_rate": dict(drive="kiosk", clip=("survey", 30000), sr=48000, n_out=40000, preset="default", seed=5,
                         segments=[seg(rate=0.3, semitones=2.0)]),
    "stream_480_512_cheaper": dict(drive="stream", clip=("survey", 30000), sr=48000, n_in=480, n_out=512, preset="cheaper",
                                   segments=[seg(semitones=1.0)]),
    "stream_100_900": dict(drive="stream", clip=("survey", 30000), sr=48000, n_in=100, n_out=900, preset="default",
                           segments=[seg()]),
    "sweep_default": dict(drive="kiosk", clip=("sweep", 4.0, 48000, 2), sr=48000, n_out=160000, preset="default",
                          segments=_sweep_segments(160000 / 48000, 48000, 25, tonality_hz=8000.0)),
    "sweep_cheaper_formant": dict(drive="kiosk", clip=("sweep", 3.0, 48000, 2), sr=48000, n_out=100000, preset="cheaper",
                                  segments=_sweep_segments(100000 / 48000, 48000, 40, rate0=0.6, rate1=1.9, st0=-7, st1=9,
                                                           tonality_hz=16000.0, formant_semitones=-3.0,
                                                           formant_compensation=True)),
    "loop": dict(drive="kiosk", clip=("survey", 40000), sr=48000, n_out=70000, preset="cheaper",
                 segments=[seg(rate=1.25, semitones=-4.0, loop_start=0.2, loop_end=0.55)]),
    "mono_custom_block": dict(drive="kiosk", clip=("noise", 30000, 1, 3, 0.2), sr=44100, n_out=33000, block=(2048, 512, 0),
                              segments=[seg(rate=0.9, semitones=2.0, formant_semitones=2.0)]),
    "lowlat_8ch_formant_auto": dict(drive="kiosk", clip=("tones", 48000, 8, 96000), sr=96000, n_out=40000, block=(960, 240, 1),
                                    segments=[seg(rate=1.1, semitones=-2.0, formant_semitones=3.0, formant_compensation=True,
                                                  formant_base_hz=0.0)]),
    "kiosk_shipped_200ms": dict(drive="kiosk", clip=("survey", 50000), sr=48000, n_out=50000, block=(9600, 2400, 1),
                                segments=[seg(rate=0.8, semitones=-5.0, tonality_hz=16000.0)]),
    "stream_transpose_only_q96": dict(drive="kiosk", clip=("survey", 30000), sr=48000, n_out=30000, preset="default", quantum=96,
                                      segments=[seg(rate=1.0, semitones=12.0)]),
    # setTransposeFactor / setFormantFactor (exports "q" / "s"): multipliers instead of semitones
    "factor_setters": dict(drive="kiosk", clip=("survey", 40000), sr=48000, n_out=40000, preset="default",
                           segments=[seg(rate=0.8, transpose_factor=1.3348399, formant_factor=0.84, formant_compensation=True)]),
    "factor_setters_stream": dict(drive="stream", clip=("survey", 30000), sr=48000, n_in=480, n_out=512, preset="cheaper",
                                  segments=[seg(transpose_factor=0.75, tonality_hz=0.0, formant_factor=1.2)]),
    # the kiosk's real operating range: rate 0.001 by default (app/multi/app.mjs:113), clamp floor 1e-5 (:483); rate*H <= 1
    # takes the other branch of seekTimeFactor (W#49) and every block draws random time factors
    "rate_1e-5": dict(drive="kiosk", clip=("survey", 20000), sr=48000, n_out=30000, preset="default", seed=7,
                      segments=[seg(rate=1e-5, input=0.2, semitones=-3.0)]),
    "rate_5e-4_cheaper": dict(drive="kiosk", clip=("survey", 20000), sr=48000, n_out=30000, preset="cheaper", seed=8,
                              segments=[seg(rate=5e-4, input=0.1, semitones=4.0, tonality_hz=16000.0)]),
    "rate_1e-3_shipped": dict(drive="kiosk", clip=("survey", 30000), sr=48000, n_out=40000, block=(9600, 2400, 1), seed=9,
                              segments=[seg(rate=1e-3, input=0.3, semitones=0.0, tonality_hz=16000.0, formant_base_hz=200.0)]),
    # stop() / start(): inactive segments run process(q, q) on a zeroed buffer (:861-869).  A short pause (the gate stays
    # open), and a stop that lasts: after 2 * blockSamples of it the silence gate closes and the output is zeros
    "pause_and_resume": dict(drive="kiosk", clip=("survey", 40000), sr=48000, n_out=40000, preset="default",
                             segments=[seg(rate=1.2, semitones=2.0), seg(output=0.2, input=0.24, rate=1.2, semitones=2.0, active=False),
                                       seg(output=0.33, input=0.24, rate=0.9, semitones=-1.0)]),
    "stop_for_good_cheaper": dict(drive="kiosk", clip=("survey", 40000), sr=48000, n_out=50000, preset="cheaper",
                                  segments=[seg(rate=1.0, semitones=3.0), seg(output=0.31, input=0.31, rate=1.0, semitones=3.0, active=False)]),
}

# Cases for the 18-call surface only (oracle engines and the compat shim): the silence gate of process() (W#48
# 7838-7943; fires after 2L silent input samples, re-arms the block phase wherever it stands) and flush() (W#46).  The
# batched path plans whole drives ahead of the data and does not model either.
_GAPS = ((6000, 30000), (40000, 55000))
SHIM_CASES = {
    "gate_default": dict(drive="stream", clip=("gated", 60000, _GAPS), sr=48000, n_in=512, n_out=512, preset="default",
                         segments=[seg(semitones=2.0)]),
    "gate_cheaper_split": dict(drive="stream", clip=("gated", 60000, _GAPS), sr=48000, n_in=128, n_out=128, preset="cheaper",
                               segments=[seg(semitones=2.0)]),
    "gate_cheaper_random_tf": dict(drive="stream", clip=("gated", 60000, _GAPS), sr=48000, n_in=300, n_out=700, preset="cheaper",
                                   seed=5, segments=[seg(semitones=2.0)]),
    "flush_default": dict(drive="stream", clip=("survey", 24000), sr=48000, n_in=300, n_out=700, preset="default", seed=5,
                          flush=(30, 3000), segments=[seg(semitones=3.0)]),
    "flush_cheaper_before_prediction": dict(drive="stream", clip=("survey", 24000), sr=48000, n_in=480, n_out=512, preset="cheaper",
                                            flush=(27, 7000), segments=[seg(semitones=3.0)]),
    "flush_cheaper_inside_prediction": dict(drive="stream", clip=("survey", 24000), sr=48000, n_in=480, n_out=512, preset="cheaper",
                                            flush=(25, 2000), segments=[seg(semitones=3.0)]),
    "flush_cheaper_during_synthesis": dict(drive="stream", clip=("survey", 24000), sr=48000, n_in=480, n_out=512, preset="cheaper",
                                           flush=(22, 1440), segments=[seg(semitones=3.0)]),
}

# small subset used by the CPU suite for the (slower) serial emulation of the kernels
FAST = ["KA5", "rng_low_rate", "stream_480_512_cheaper", "stream_100_900", "loop", "mono_custom_block",
        "lowlat_8ch_formant_auto", "stream_transpose_only_q96", "factor_setters", "factor_setters_stream", "rate_1e-5",
        "rate_5e-4_cheaper", "rate_1e-3_shipped", "pause_and_resume", "stop_for_good_cheaper"]


def sha_of(y):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(y, np.float32).tobytes()).hexdigest()
