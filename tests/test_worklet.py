"""CPU: the host mirror of the worklet's control plane (remoteMethods + per-quantum bookkeeping) and the controller
message mapping.  The mirror drives the CPU oracle quantum by quantum like ``WasmProcessor.process``; the same control
trace resolved into a per-quantum table and run through the batched engine (host-emulation build here, CUDA in the
gpu tests) must give the same bits."""
import json

import numpy as np
import pytest

import bauklank_audio_stretch_b200 as bs
import cases
from conftest import HOSTEMU
from oracle import refdrive

FULL = dict(tonalityHz=8000.0, formantSemitones=0.0, formantCompensation=False, formantBaseHz=0.0)


@pytest.fixture(scope="module")
def emu():
    return bs.load_library(HOSTEMU)


def test_schedule_semantics(emu):
    """remoteMethods.schedule (app/SignalsmithStretch.mjs:656-701): later segments are popped, fields inherited from
    the latest one, input extrapolated at the latest segment's rate (0 if it was inactive), adjustPrevious re-rates."""
    t = bs.WorkletTimeline(48000.0, lib=emu)
    assert t.latency() == pytest.approx(5760 / 48000.0) and t.buffer_length == 5760
    a = t.schedule(dict(active=True, input=1.0, rate=0.5, outputTime=2.0, **FULL))
    assert a["input"] == 1.0 and a["output"] == 2.0
    b = t.schedule(dict(rate=2.0, outputTime=4.0, **FULL))
    assert b["input"] == pytest.approx(1.0 + 2.0 * 0.5) and b["active"] is True and b["semitones"] == 0.0
    c = t.schedule(dict(outputTime=3.0, semitones=5, **FULL))          # pops b (output 4 >= 3), inherits from it
    # ... and the trailing loop of schedule() (:690-694) SHIFTS every segment whose successor starts by the new output
    # time -- so a segment scheduled ahead becomes the current one at once and is extrapolated backwards until its time
    assert [s["output"] for s in t.time_map] == [3.0]
    assert c["rate"] == 2.0 and c["input"] == pytest.approx(2.0 + (3.0 - 4.0) * 2.0)
    d = t.schedule(dict(outputTime=5.0, input=10.0, **FULL), adjust_previous=True)
    assert c["rate"] == pytest.approx((10.0 - c["input"]) / (5.0 - 3.0)) and d["input"] == 10.0 and d["rate"] == 2.0
    s = t.stop(6.0)
    assert s["active"] is False and s["input"] == pytest.approx(10.0 + 1.0 * d["rate"])
    e = t.schedule(dict(outputTime=7.0, **FULL))                        # after an inactive segment the input stands still
    assert e["input"] == s["input"] and t.time_map == [e]
    t2 = bs.WorkletTimeline(48000.0, config=dict(blockMs=200, splitComputation=True), lib=emu)   # the kiosk's shipped config
    assert (t2.block_samples, t2.interval_samples, t2.buffer_length) == (9600, 2400, 12000)
    with pytest.raises(ValueError):                                      # the NaN quirk is refused, not propagated
        t3 = bs.WorkletTimeline(48000.0, lib=emu); t3.schedule(dict(active=True, outputTime=0.0)); t3.quantum()


def _trace():
    """A control trace with everything the worklet supports in buffer playback: start, re-rates with and without
    explicit input (seek), transpose / formant changes, a loop, schedule-ahead and a segment that pops a future one."""
    ev = [(0, "schedule", (dict(active=True, input=0.0, rate=1.0, semitones=0, outputTime=0.0, **FULL),)),
          (40, "schedule", (dict(rate=0.7, semitones=3, outputTime=40 * 128 / 48000 + 0.1, **FULL),)),
          (90, "schedule", (dict(rate=1.6, semitones=-5, tonalityHz=16000.0, formantSemitones=2.0, formantCompensation=True,
                                 formantBaseHz=180.0, outputTime=0.5),)),
          (95, "schedule", (dict(rate=1.1, outputTime=0.3, tonalityHz=8000.0, formantSemitones=0.0, formantCompensation=False,
                                 formantBaseHz=0.0),)),                  # pops the 0.5 segment
          (150, "schedule", (dict(input=0.1, rate=0.9, semitones=7, loopStart=0.1, loopEnd=0.35, outputTime=150 * 128 / 48000, **FULL),))]
    return ev


@pytest.mark.parametrize("preset", ["default", "cheaper"])
def test_trace_through_table_equals_quantum_by_quantum_drive(preset, emu):
    clip = refdrive.survey_clip(30000)
    n_out = 40000
    tl = bs.WorkletTimeline(48000.0, config=dict(preset=preset), lib=emu)
    ref = tl.render(refdrive.PortEngine(), n_out, events=_trace(), clip=clip)
    tl2 = bs.WorkletTimeline(48000.0, config=dict(preset=preset), lib=emu)
    tl2.addBuffers(clip)
    recs = tl2.resolve(n_out, events=_trace())
    assert any(r["rate"] == 0.7 for r in recs) and any(r["input_samples_end"] < recs[i - 1]["input_samples_end"] for i, r in enumerate(recs) if i)
    eng = bs.BatchStretch(2, 48000.0, preset=preset, lib=emu)
    outs = eng.plan([np.ascontiguousarray(clip)], [bs.TableDrive(n_out, bs.WorkletTimeline.table(recs))])
    eng.run()
    assert cases.compare(np.array(outs[0]), ref)[0]
    eng.close()


def test_table_equals_segment_drive_on_a_golden_case(emu, golden):
    """One static segment: the table path must reproduce the golden vector of the segment path."""
    from conftest import assert_matches_golden
    case = cases.CASES["loop"]
    clip = cases.make_clip(case["clip"])
    s = case["segments"][0]
    tl = bs.WorkletTimeline(48000.0, config=dict(preset="cheaper"), lib=emu)
    tl.addBuffers(clip)
    tl.time_map = [dict(active=True, input=s["input"], output=s["output"], rate=s["rate"], semitones=s["semitones"], tonalityHz=s["tonality_hz"],
                        formantSemitones=s["formant_semitones"], formantCompensation=s["formant_compensation"],
                        formantBaseHz=s["formant_base_hz"], loopStart=s["loop_start"], loopEnd=s["loop_end"])]
    recs = tl.resolve(case["n_out"])
    eng = bs.BatchStretch(2, 48000.0, preset="cheaper", lib=emu)
    outs = eng.plan([np.ascontiguousarray(clip)], [bs.TableDrive(case["n_out"], bs.WorkletTimeline.table(recs))])
    eng.run()
    assert_matches_golden("loop", np.array(outs[0]), golden)
    eng.close()


def test_controller_messages_to_schedule(emu):
    """server-multi.py:47-48 wire format -> app/multi/app.mjs:537-616 mapping -> schedule() calls."""
    m = bs.ControllerMapper(audio_duration=0.6, channel="A")
    assert m.normalize(dict(type="set", key="tone", value="7"))["value"] == 7
    assert m.normalize(dict(type="set", key="rate", value="0.5"))["value"] == 0.5
    lines = [(0.00, json.dumps(dict(type="set", channel="A", key="rate", value=1.0))),
             (0.05, json.dumps(dict(type="set", channel="B", key="rate", value=2.0))),      # other engine: ignored
             (0.10, json.dumps(dict(type="set", channel="A", key="tone", value=30))),        # clamped to +24, integer
             (0.20, json.dumps(dict(type="set", channel="A", key="volume", value=55))),      # still re-schedules (controlsChanged)
             (0.30, json.dumps(dict(type="set", channel="A", key="rate", value=5.0))),       # clamped to 2 by controlsChanged
             (0.35, "not json"),
             (0.40, json.dumps(dict(type="set", channel="A", key="tone", value="x")))]       # not a number: ignored
    ev = m.trace_to_events(lines)
    assert len(ev) == 4 and ev[1][2][0]["semitones"] == 24 and ev[3][2][0]["rate"] == 2 and m.values["volume"] == 0.55
    assert all(e[1] == "schedule" and e[2][0]["outputTime"] == pytest.approx(e[0] * 128 / 48000.0 + 0.1) for e in ev)
    assert ev[0][2][0]["tonalityHz"] == 16000 and ev[0][2][0]["formantBaseHz"] == 200 and ev[0][2][0]["loopStart"] == 0.6   # clamped to the audio
    # and the trace runs: mirror-driven oracle == table-driven engine
    clip = refdrive.survey_clip(28800)
    n_out = 30000
    ref = bs.WorkletTimeline(48000.0, config=dict(preset="cheaper"), lib=emu).render(refdrive.PortEngine(), n_out, events=ev, clip=clip)
    tl = bs.WorkletTimeline(48000.0, config=dict(preset="cheaper"), lib=emu); tl.addBuffers(clip)
    recs = tl.resolve(n_out, events=ev)
    assert recs[0]["active"] and recs[0]["rate"] == 1.0      # scheduled 0.1 s ahead, yet current at once (see test_schedule_semantics)
    eng = bs.BatchStretch(2, 48000.0, preset="cheaper", lib=emu)
    outs = eng.plan([np.ascontiguousarray(clip)], [bs.TableDrive(n_out, bs.WorkletTimeline.table(recs))])
    eng.run()
    assert cases.compare(np.array(outs[0]), ref)[0] and np.abs(ref).max() > 1e-3
    eng.close()
    # an idle worklet (nothing scheduled: the initial segment is inactive) is process(q,q) on silence: zeros
    tl = bs.WorkletTimeline(48000.0, lib=emu); tl.addBuffers(clip)
    eng = bs.BatchStretch(2, 48000.0, lib=emu)
    outs = eng.plan([np.ascontiguousarray(clip)], [bs.TableDrive(1000, bs.WorkletTimeline.table(tl.resolve(1000)))])
    eng.run()
    assert np.array(outs[0]).shape == (2, 1000) and not np.array(outs[0]).any()
    eng.close()


def _stop_start_trace(kind):
    """play, stop, play again -- what the kiosk's `active` control produces: controlsChanged re-schedules the whole parameter set
    with active toggled (app/multi/app.mjs:495-507).  (remoteMethods.stop() itself, :625-630, schedules a segment without
    tonalityHz & co. and hands NaN to the engine -- the quirk the mirror refuses, see worklet.py.)"""
    def stop(k):
        return (k, "schedule", (dict(active=False, outputTime=k * 128 / 48000, **FULL),))
    ev = [(3, "schedule", (dict(active=True, input=0.05, rate=1.1, semitones=2, outputTime=3 * 128 / 48000, **FULL),))]
    if kind == "pause":            # 0.1 s of inactivity: the silence gate stays open (2 * 5760 samples = 0.24 s)
        ev += [stop(60), (98, "schedule", (dict(active=True, rate=0.8, semitones=-3, outputTime=98 * 128 / 48000, **FULL),))]
    elif kind == "late_start":     # idle for 0.2 s before anything plays, then a stop that lasts to the end
        ev = [(75, "schedule", (dict(active=True, input=0.0, rate=1.0, semitones=0, outputTime=75 * 128 / 48000, **FULL),)), stop(170)]
    elif kind == "stop":           # stop for good
        ev += [stop(90)]
    elif kind == "restart":        # ... and start again long after the gate has closed
        ev += [stop(60), (260, "schedule", (dict(active=True, rate=1.0, semitones=0, outputTime=260 * 128 / 48000, **FULL),))]
    return ev


@pytest.mark.parametrize("preset", ["default", "cheaper"])
@pytest.mark.parametrize("kind", ["pause", "late_start", "stop"])
def test_stop_and_start_run_on_the_batched_path(kind, preset, emu):
    """Inactive segments are process(q, q) on a zeroed buffer (app/SignalsmithStretch.mjs:861-869), silence gate included:
    the table-driven engine follows the oracle driven quantum by quantum, bit for bit."""
    clip = refdrive.survey_clip(30000)
    n_out = 45000
    ref = bs.WorkletTimeline(48000.0, config=dict(preset=preset), lib=emu).render(refdrive.PortEngine(), n_out, events=_stop_start_trace(kind), clip=clip)
    tl = bs.WorkletTimeline(48000.0, config=dict(preset=preset), lib=emu); tl.addBuffers(clip)
    recs = tl.resolve(n_out, events=_stop_start_trace(kind))
    assert any(not r["active"] for r in recs) and any(r["active"] for r in recs)
    eng = bs.BatchStretch(2, 48000.0, preset=preset, lib=emu)
    clips = [np.ascontiguousarray(clip)]
    outs = eng.plan(clips, [bs.TableDrive(n_out, bs.WorkletTimeline.table(recs))], chunk_blocks=7)
    eng.run()
    assert cases.compare(np.array(outs[0]), ref)[0] and np.abs(ref).max() > 1e-2
    assert eng.gate_events() == 0
    if kind != "pause":
        assert not ref[:, -4000:].any()                      # the gate closed: zeros to the end
    host = [np.zeros_like(ref)]                                # ... and through the host-audio entry point (tail included)
    eng.run_host(clips, host)
    assert cases.compare(host[0], ref)[0]
    eng.close()


def test_restart_after_the_gate_closed_is_refused_and_a_silent_seek_is_reported(emu):
    clip = refdrive.survey_clip(30000)
    tl = bs.WorkletTimeline(48000.0, lib=emu); tl.addBuffers(clip)
    recs = tl.resolve(45000, events=_stop_start_trace("restart"))
    eng = bs.BatchStretch(2, 48000.0, lib=emu)
    with pytest.raises(RuntimeError, match="silence gate"):
        eng.plan([np.ascontiguousarray(clip)], [bs.TableDrive(45000, bs.WorkletTimeline.table(recs))])
    # a pause whose first seek afterwards lands on digital silence inside the clip: the plan took it for loud, the run says so
    quiet = clip.copy(); quiet[:, 4000:26000] = 0.0
    tl = bs.WorkletTimeline(48000.0, lib=emu); tl.addBuffers(quiet)
    recs = tl.resolve(30000, events=_stop_start_trace("pause"))
    outs = eng.plan([np.ascontiguousarray(quiet)], [bs.TableDrive(30000, bs.WorkletTimeline.table(recs))])
    eng.run()
    assert eng.gate_events() == 1
    eng.close()


@pytest.mark.parametrize("preset", ["default", "cheaper"])
def test_native_trace_drive_equals_the_python_mirror(preset, emu):
    """bsb_add_kiosk_trace keeps the worklet's time map inside the library (schedule() semantics in C++): same bits as the
    Python mirror's per-quantum table, for the full-feature trace, a controller trace, and a pause."""
    clip = refdrive.survey_clip(30000)
    n_out = 40000
    m = bs.ControllerMapper(audio_duration=30000 / 48000.0, channel="B")
    rng = np.random.default_rng(4)
    lines, t = [], 0.0
    while t < 0.8:
        key = "rate" if rng.random() < 0.5 else "tone"
        val = float(np.exp(rng.uniform(np.log(0.5), np.log(2.0)))) if key == "rate" else int(rng.integers(-12, 13))
        lines.append((t, json.dumps(dict(type="set", channel="B", key=key, value=val))))
        t += float(rng.uniform(0.02, 0.1))
    traces = [_trace(), m.trace_to_events(lines), _stop_start_trace("pause")]
    for ev in traces:
        tl = bs.WorkletTimeline(48000.0, config=dict(preset=preset), lib=emu); tl.addBuffers(clip)
        recs = tl.resolve(n_out, events=ev)
        eng = bs.BatchStretch(2, 48000.0, preset=preset, lib=emu)
        c = np.ascontiguousarray(clip)
        outs = eng.plan([c, c], [bs.TableDrive(n_out, bs.WorkletTimeline.table(recs)), bs.TraceDrive(n_out, bs.trace_events(ev))])
        eng.run()
        assert cases.compare(np.array(outs[0]), np.array(outs[1]))[0] and np.abs(np.array(outs[1])).max() > 1e-2
        for b in range(eng.stream_blocks(0)):
            assert eng.block_info(0, b) == eng.block_info(1, b)
        eng.close()
    with pytest.raises(RuntimeError, match="NaN"):        # the quirk the mirror refuses is refused here too
        eng = bs.BatchStretch(2, 48000.0, lib=emu)
        eng.plan([np.ascontiguousarray(clip)], [bs.TraceDrive(1000, bs.trace_events([(0, "schedule", (dict(active=True),))]))])
