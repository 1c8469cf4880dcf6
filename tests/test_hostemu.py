"""CPU: the kernels' own source, executed serially on the host by the TEST-ONLY emulation build
(tests/hostemu, -DBS_HOSTEMU: every (tid, nthreads) work loop runs with one thread), must reproduce the reference
bit-for-bit.  This checks the arithmetic and the host block-schedule compiler without a GPU; the `-m gpu` tests run
the same cases through the real CUDA build."""
import numpy as np
import pytest

import bauklank_audio_stretch_b200 as bs
import cases
from conftest import HOSTEMU, assert_matches_golden
from oracle import refdrive


@pytest.fixture(scope="module")
def emu():
    return bs.load_library(HOSTEMU)


@pytest.mark.parametrize("name", cases.FAST)
def test_batched_path_bit_exact(name, emu, golden):
    y = cases.run_cases_batch(bs, [cases.CASES[name]], lib=emu)[0]
    assert_matches_golden(name, y, golden)


def test_batch_of_mixed_streams_and_chunk_invariance(emu, golden):
    """Several presetDefault streams in ONE batch, different lengths/drives; result must not depend on the chunking."""
    names = ["KA5", "rng_low_rate", "stream_100_900", "stream_transpose_only_q96"]
    cs = [cases.CASES[n] for n in names]
    a = cases.run_cases_batch(bs, cs, lib=emu, chunk_blocks=5)
    for n, y in zip(names, a):
        assert_matches_golden(n, y, golden)
    b = cases.run_cases_batch(bs, cs[:2], lib=emu, chunk_blocks=64)
    for n, y in zip(names[:2], b):
        assert_matches_golden(n, y, golden)


def test_block_schedule_matches_the_reference_drive(emu):
    """Frame/hop indexing: the host-compiled block table (start sample of each analysis window, timeFactor bits,
    flags) equals what the reference engine does under the worklet's drive.  Checked through the oracle's seek
    positions: for a kiosk drive the 'current' window of the block that starts in quantum k ends at
    inputSamplesEnd(k) (app/SignalsmithStretch.mjs:897)."""
    case = dict(cases.CASES["KA4"]); case["n_out"] = 30000
    clip = cases.make_clip(case["clip"])
    eng = cases.make_batch(bs, case, 2, lib=emu)
    eng.plan([clip], [cases.batch_drive(bs, case, clip.shape[1])])
    L, H = eng.blockSamples(), eng.intervalSamples()
    sr = 48000.0
    in_lat, out_lat = eng.inputLatency(), eng.outputLatency()
    nb = eng.stream_blocks(0)
    assert nb == (30000 + H - 1) // H
    for m in range(nb):
        k = (m * H) // 128                                   # quantum in which output sample m*H falls
        out_t = (k * 128) / sr + out_lat / sr
        end = refdrive.js_round((0.0 + out_t * 0.75 + in_lat / sr) * sr)
        info = eng.block_info(0, m)
        assert info["flags"] & 1                             # seek => new spectrum every block
        assert info["cur"][0] == end - L
        assert info["prev"][0] == end - L - H
        assert np.float32(info["timeFactor"]) == np.float32(1.0 / 0.75)
    eng.close()


@pytest.mark.parametrize("name", ["KA5", "stream_480_512_cheaper", "rng_low_rate"])
def test_compat_shim_bit_exact(name, emu, golden):
    """The reference's own 18 entry points (Part 1 of the header) driven exactly like the worklet drives the wasm."""
    e = bs.StretchEngine(seed=cases.CASES[name].get("seed", 1), lib=emu)
    y = cases.run_case(e, cases.CASES[name])
    assert_matches_golden(name, y, golden)


@pytest.mark.parametrize("name", list(cases.SHIM_CASES))
def test_compat_shim_gate_and_flush(name, emu, golden):
    """process()'s silence gate (re-arming the block phase mid-interval, with and without split computation, with random
    time factors) and flush(): the shim keeps the reference's own output ring, so both are bit-exact."""
    e = bs.StretchEngine(seed=cases.SHIM_CASES[name].get("seed", 1), lib=emu)
    y = cases.run_case(e, cases.SHIM_CASES[name])
    assert_matches_golden(name, y, golden)


def test_shim_parameter_changes_every_quantum(emu):
    """Q4: with splitComputation the steps of a block see the parameters current when they run."""
    x = refdrive.survey_clip(20000)
    def pf(k, t):
        return dict(semitones=float((k // 7) % 13 - 6), rate=0.5 + (k % 50) / 40.0, formant_semitones=float((k // 11) % 5 - 2),
                    formant_comp=bool(k % 2))
    for preset in ("cheaper", "default"):
        a, _ = refdrive.kiosk_drive(refdrive.PortEngine(5), x, 48000, 25000, 1.0, preset=preset, params=dict(tonality_hz=8000.0), param_fn=pf)
        b, _ = refdrive.kiosk_drive(bs.StretchEngine(seed=5, lib=emu), x, 48000, 25000, 1.0, preset=preset, params=dict(tonality_hz=8000.0), param_fn=pf)
        assert cases.compare(a, b)[0], preset


def test_error_behaviour(emu):
    eng = bs.BatchStretch(2, 48000.0, lib=emu)
    x = np.zeros((2, 1000), np.float32)
    with pytest.raises(RuntimeError):      # the streaming drive's silence gate depends on the audio: no inactive segments there
        eng.plan([x], [bs.StreamingDrive(100, 100, 5, [bs.segment(active=False)])])
    with pytest.raises(RuntimeError):      # n_calls * n_in must fit the clip
        eng.plan([x], [bs.StreamingDrive(512, 512, 10)])
    # playing again after the silence gate has closed re-arms the block phase mid-interval: refused, with a reason
    y = np.ones((2, 60000), np.float32)
    with pytest.raises(RuntimeError, match="silence gate"):
        eng.plan([y], [bs.KioskDrive(60000, [bs.segment(), bs.segment(output=0.1, input=0.1, active=False), bs.segment(output=0.6, input=0.1)])])
    eng.close()
    with pytest.raises(RuntimeError):      # block < 8
        bs.BatchStretch(2, 48000.0, block_samples=4, interval_samples=1, lib=emu)


def test_run_host_entry_point(emu, golden):
    case = cases.CASES["KA5"]
    clip = cases.make_clip(case["clip"])
    eng = cases.make_batch(bs, case, 2, lib=emu)
    outs = eng.plan([np.zeros_like(clip)], [cases.batch_drive(bs, case, clip.shape[1])], chunk_blocks=5)
    ho = [np.zeros_like(outs[0])]
    eng.run_host([clip], ho)
    assert_matches_golden("KA5", ho[0], golden)
    eng.close()


def test_edge_cases_empty_tiny_and_ragged(emu):
    """Empty output, one output sample, a clip shorter than the block, output ending mid-interval and mid-quantum, and a
    stream that plays far past the end of its clip -- in one batch next to a normal stream; each equals the oracle."""
    rng = np.random.default_rng(5)
    specs = [(0, 3000, 1.0, 0.0), (1, 3000, 1.0, 3.0), (777, 100, 0.5, -2.0), (1441, 9000, 2.0, 5.0), (12345, 2000, 1.3, 0.0),
             (20000, 20000, 0.9, 7.0)]
    clips, drives, refs = [], [], []
    for n_out, n_in, rate, st in specs:
        clip = (0.2 * rng.standard_normal((2, n_in))).astype(np.float32)
        clips.append(clip)
        drives.append(bs.KioskDrive(n_out, [bs.segment(rate=rate, semitones=st)]))
        e = refdrive.PortEngine()
        case = dict(drive="kiosk", sr=48000, n_out=n_out, preset="default", segments=[cases.seg(rate=rate, semitones=st)])
        refs.append(cases.run_case(e, case, clip=clip)); e.close()
    eng = bs.BatchStretch(2, 48000.0, lib=emu)
    outs = eng.plan(clips, drives, chunk_blocks=3)
    eng.run()
    for (n_out, *_), o, r in zip(specs, outs, refs):
        assert o.shape == (2, n_out) and cases.compare(np.array(o), r)[0], n_out
    assert eng.stream_blocks(0) == 0 and eng.stream_blocks(1) == 1 and eng.stream_blocks(3) == 2
    eng.close()
    # a batch whose every stream is empty is legal and does nothing
    eng = bs.BatchStretch(2, 48000.0, lib=emu)
    outs = eng.plan([clips[0]], [bs.KioskDrive(0, [bs.segment()])]); eng.run(); eng.close()
    assert outs[0].shape == (2, 0)


def test_batch_reports_when_the_silence_gate_would_fire(emu):
    """The batched path does not model process()'s silence gate; it counts the calls the reference would have gated."""
    case = cases.SHIM_CASES["gate_default"]
    loud = dict(case); loud["clip"] = ("survey", 20000)
    for c, want_some in ((case, True), (loud, False)):
        clip = cases.make_clip(c["clip"])
        eng = cases.make_batch(bs, c, 2, lib=emu)
        eng.plan([np.ascontiguousarray(clip)], [cases.batch_drive(bs, c, clip.shape[1])])
        eng.run()
        want = cases.expected_gate_events(clip, c["n_in"], clip.shape[1] // c["n_in"], eng.blockSamples())
        assert eng.gate_events() == want and (want > 0) == want_some
        eng.close()


def test_longest_first_ordering_relay_chunks_and_host_audio(emu):
    """Streams are kept longest first inside the engine and the tail of the run uses chunks of several hundred blocks per
    stream (the relayed chain wavefront on the GPU); results and the stream <-> buffer association must not change.
    Three mono streams of very different lengths in a small-block geometry, given shortest first; device-style run and
    host-audio run (short first chunk) against the oracle."""
    specs = [dict(n=9000, rate=1.4, st=5.0, seed=7), dict(n=60000, rate=0.55, st=-2.0, seed=8), dict(n=30000, rate=0.9, st=3.0, seed=9)]
    cs = [dict(drive="kiosk", clip=("noise", sp["n"], 1, sp["seed"], 0.2), sr=44100, n_out=int(sp["n"] / sp["rate"]), block=(512, 128, 0),
               seed=sp["seed"], segments=[cases.seg(rate=sp["rate"], semitones=sp["st"])]) for sp in specs]
    refs = []
    for c in cs:
        eng = refdrive.PortEngine(seed=c["seed"]); refs.append(cases.run_case(eng, c)); eng.close()
    got = cases.run_cases_batch(bs, cs, lib=emu)
    for y, r in zip(got, refs):
        assert cases.compare(y, r)[0]
    clips = [np.ascontiguousarray(cases.make_clip(c["clip"])) for c in cs]
    eng = cases.make_batch(bs, cs[0], 1, lib=emu)
    outs = eng.plan([np.zeros_like(x) for x in clips], [cases.batch_drive(bs, c, x.shape[1]) for c, x in zip(cs, clips)])
    assert eng.stream_blocks(1) > 256 > eng.stream_blocks(0)          # long enough for a relay chunk; given shortest first
    host_outs = [np.zeros_like(o) for o in outs]
    eng.run_host(clips, host_outs)
    for y, r in zip(host_outs, refs):
        assert cases.compare(y, r)[0]
    eng.close()


def test_shim_flush_at_many_positions_of_the_block_cycle(emu):
    """flush() while a block's steps are spread over the interval (split computation): before, inside and after the
    vertical prediction, during synthesis, on interval boundaries -- always the reference's bits, also for the
    processing that follows."""
    bad = []
    for n_in, n_out, positions in ((480, 512, range(20, 32)), (300, 700, range(20, 30, 2)), (64, 64, range(30, 50, 3))):
        for calls in positions:
            case = dict(drive="stream", clip=("survey", 30000), sr=48000, n_in=n_in, n_out=n_out, preset="cheaper", seed=5,
                        n_calls=calls + 14, flush=(calls, 2000), segments=[cases.seg(semitones=3.0)])
            a = refdrive.PortEngine(seed=5); ya = cases.run_case(a, case); a.close()
            yb = cases.run_case(bs.StretchEngine(seed=5, lib=emu), case)
            if not cases.compare(ya, yb)[0]:
                bad.append((n_in, n_out, calls))
    assert not bad, bad


@pytest.mark.parametrize("channels,block,interval,split,sr", [
    (2, 5760, 1440, 0, 48000), (2, 4800, 1920, 1, 48000), (2, 9600, 2400, 1, 48000), (1, 11520, 2880, 0, 96000), (3, 960, 240, 1, 96000)])
def test_specialised_stft_kernels_equal_the_generic_ones_and_the_oracle(channels, block, interval, split, sr, emu):
    """fft_fast.cuh (the preset geometries: inner x outer = 1024x3, 512x5, 1024x5, 2048x3, 512x1) against the run-time-geometry
    path of kernels.cuh and against the CPU oracle, through the serial emulation of the same source; a window that starts
    before the clip, the short pre-roll of the non-split presets and an odd clip offset are all in there."""
    rng = np.random.default_rng(block)
    n_in = int(0.5 * sr)
    clip = (0.2 * rng.standard_normal((channels, n_in))).astype(np.float32)
    case = dict(drive="kiosk", sr=sr, n_out=int(0.35 * sr), block=(block, interval, split), seed=3,
                segments=[cases.seg(rate=0.83, input=0.0131, semitones=3.0, formant_semitones=-2.0, formant_compensation=True)])
    e = refdrive.PortEngine(seed=3)
    ref = cases.run_case(e, case, clip=clip); e.close()
    outs = []
    for fast in (True, False):
        eng = bs.BatchStretch(channels, sr, block_samples=block, interval_samples=interval, split_computation=bool(split), lib=emu)
        eng.set_fast_fft(fast)
        assert eng.fast_fft_active() == fast
        o = eng.plan([np.ascontiguousarray(clip)], [cases.batch_drive(bs, case, n_in)], chunk_blocks=9)
        eng.run()
        outs.append(np.array(o[0])); eng.close()
    assert cases.compare(outs[0], outs[1])[0] and cases.compare(outs[0], ref)[0]
