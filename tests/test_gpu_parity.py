"""GPU (B200): the CUDA path, called through the C ABI, against (a) the golden vectors minted from the reference's
own blob, (b) the CPU oracle run live on fresh seeded inputs, (c) size-independent properties at batch sizes the
oracle could not cover in seconds.

Bar: integer/index work (block schedule) bit-exact; audio within max|err| <= 1e-4 and SNR >= 90 dB per channel
(BASELINE.json north_star).  The kernels follow the reference's f32 operation order, so in practice the audio is
bit-identical and the tests assert that first and report the tolerance numbers if it ever is not."""
import numpy as np
import pytest

import cases
from conftest import assert_matches_golden
from oracle import refdrive

pytestmark = pytest.mark.gpu

MAX_ABS_ERR = 1e-4
MIN_SNR_DB = 90.0


@pytest.fixture(scope="module")
def bs():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    import bauklank_audio_stretch_b200 as m
    m.load_library()          # the in-tree CUDA build; raises if missing (no fallback)
    return m


@pytest.mark.parametrize("name", list(cases.CASES))
def test_batched_path_matches_reference_golden(name, bs, golden):
    y = cases.run_cases_batch(bs, [cases.CASES[name]], device="cuda:0")[0]
    assert_matches_golden(name, y, golden)


def test_mixed_batch_and_chunk_invariance(bs, golden):
    names = ["KA1", "KA2", "KA5", "KA6", "rng_low_rate", "stream_100_900", "sweep_default", "stream_transpose_only_q96"]
    cs = [cases.CASES[n] for n in names]
    for chunk in (0, 7, 33):
        ys = cases.run_cases_batch(bs, cs, device="cuda:0", chunk_blocks=chunk)
        for n, y in zip(names, ys):
            assert_matches_golden(n, y, golden)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_fresh_random_cases_against_live_oracle(seed, bs):
    """Inputs nobody minted a vector for: random rate/transpose/formant schedule, checked against the CPU oracle."""
    rng = np.random.default_rng(seed)
    preset = "cheaper" if seed % 2 else "default"
    segs = [cases.seg(rate=float(rng.uniform(0.5, 2.0)), semitones=float(rng.integers(-12, 13)),
                      tonality_hz=float(rng.choice([8000.0, 16000.0])))]
    for i in range(1, 12):
        cases.schedule(segs, 0.12 * i, rate=float(np.exp(rng.uniform(np.log(0.5), np.log(2.0)))), semitones=float(rng.integers(-12, 13)),
                       formant_semitones=float(rng.integers(-3, 4)), formant_compensation=bool(rng.integers(0, 2)),
                       formant_base_hz=float(rng.choice([0.0, 150.0])))
    case = dict(drive="kiosk", clip=("noise", 70000, 2, seed, 0.2), sr=48000, n_out=72000, preset=preset, segments=segs, seed=seed)
    clip = cases.make_clip(case["clip"])
    eng = refdrive.PortEngine(seed=seed)
    ref = cases.run_case(eng, case, clip=clip); eng.close()
    got = cases.run_cases_batch(bs, [case], device="cuda:0")[0]
    same, err, snr = cases.compare(got, ref)
    assert err <= MAX_ABS_ERR and snr >= MIN_SNR_DB, (same, err, snr)
    assert same, "within tolerance but not bit-identical: err %.3g snr %.1f dB" % (err, snr)


def test_relayed_wavefront_for_long_streams(bs):
    """Few long streams: a chunk covers several hundred blocks per stream and each stream's chain wavefront is relayed
    across several CTAs (chain.cuh).  Three streams of different lengths (the live prefix shrinks along the run), one of
    them with random time factors; against the CPU oracle, and against the same batch run in plain 96-block chunks."""
    specs = [dict(n=150000, rate=0.6, st=4.0, seed=3), dict(n=90000, rate=0.45, st=-3.0, seed=4), dict(n=60000, rate=1.3, st=7.0, seed=5)]
    cs = []
    for sp in specs:
        cs.append(dict(drive="kiosk", clip=("noise", sp["n"], 2, sp["seed"], 0.2), sr=48000, n_out=int(sp["n"] / sp["rate"]), preset="default",
                       seed=sp["seed"], segments=[cases.seg(rate=sp["rate"], semitones=sp["st"])]))
    got = cases.run_cases_batch(bs, cs, device="cuda:0")
    plain = cases.run_cases_batch(bs, cs, device="cuda:0", chunk_blocks=96)
    for c, y, z in zip(cs, got, plain):
        assert cases.compare(y, z)[0], "relayed and plain chunking differ"
        eng = refdrive.PortEngine(seed=c["seed"])
        ref = cases.run_case(eng, c); eng.close()
        same, err, snr = cases.compare(y, ref)
        assert same, "not bit-identical to the oracle: err %.3g snr %.1f dB" % (err, snr)


@pytest.mark.parametrize("name", ["KA5", "stream_480_512_cheaper", "lowlat_8ch_formant_auto"])
def test_compat_shim_on_gpu(name, bs, golden):
    """The reference's 18 entry points, host buffers in / host buffers out, one block per launch."""
    e = bs.StretchEngine(seed=cases.CASES[name].get("seed", 1))
    y = cases.run_case(e, cases.CASES[name])
    assert_matches_golden(name, y, golden)


@pytest.mark.parametrize("name", list(cases.SHIM_CASES))
def test_compat_shim_gate_and_flush_on_gpu(name, bs, golden):
    """The silence gate of process() and flush() through the reference's own entry points (W#48 7838-7943, W#46)."""
    e = bs.StretchEngine(seed=cases.SHIM_CASES[name].get("seed", 1))
    y = cases.run_case(e, cases.SHIM_CASES[name])
    assert_matches_golden(name, y, golden)


def test_batch_reports_when_the_silence_gate_would_fire(bs):
    import torch
    case = cases.SHIM_CASES["gate_default"]
    loud = dict(case); loud["clip"] = ("survey", 20000)
    for c in (case, loud):
        clip = cases.make_clip(c["clip"])
        eng = cases.make_batch(bs, c, 2)
        eng.plan([torch.from_numpy(clip).cuda().contiguous()], [cases.batch_drive(bs, c, clip.shape[1])])
        eng.run()
        want = cases.expected_gate_events(clip, c["n_in"], clip.shape[1] // c["n_in"], eng.blockSamples())
        assert eng.gate_events() == want
        eng.close()


def test_block_schedule_bit_exact(bs):
    """frame/hop indexing: the device-side block table equals the one derived from the worklet arithmetic."""
    import torch
    case = dict(cases.CASES["KA5"])
    clip = torch.from_numpy(cases.make_clip(case["clip"])).cuda()
    eng = cases.make_batch(bs, case, 2)
    eng.plan([clip], [cases.batch_drive(bs, case, clip.shape[1])])
    L, H = eng.blockSamples(), eng.intervalSamples()
    for m in range(eng.stream_blocks(0)):
        k = (m * H) // 128
        out_t = (k * 128) / 48000.0 + eng.outputLatency() / 48000.0
        end = refdrive.js_round((out_t * 2.0 + eng.inputLatency() / 48000.0) * 48000.0)
        info = eng.block_info(0, m)
        assert info["cur"][0] == end - L and info["prev"][0] == end - L - H
        assert np.float32(info["timeFactor"]) == np.float32(0.5)
    eng.close()


def test_properties_at_batch_scale(bs):
    """256 streams x 8 s (BASELINE config 3 shape, shortened): (a) streams with identical drive and input give
    identical output wherever they sit in the batch, (b) a sample of streams equals the CPU oracle, (c) rate-1
    presetCheaper streams reproduce their input (Q1: 1920 % 128 == 0), (d) every output is finite and non-trivial."""
    import torch
    S, sr, n_in = 256, 48000, 8 * 48000
    rng = np.random.default_rng(1)
    rates = np.exp(rng.uniform(np.log(0.5), np.log(2.0), S)); sts = rng.integers(-12, 13, S)
    rates[7], sts[7] = rates[200], sts[200]
    g = torch.Generator(device="cuda").manual_seed(1)
    base = (0.1 * torch.randn((S, 2, n_in), device="cuda", generator=g))
    base[7] = base[200]
    clips = [base[i].contiguous() for i in range(S)]
    drives = [bs.KioskDrive(int(n_in / rates[i]), [bs.segment(rate=float(rates[i]), semitones=float(sts[i]))]) for i in range(S)]
    eng = bs.BatchStretch(2, sr, preset="default")
    outs = eng.plan(clips, drives); eng.run(); torch.cuda.synchronize()
    assert torch.equal(outs[7], outs[200])
    for i in (0, 7, 131, 255):
        e = refdrive.PortEngine()
        case = dict(drive="kiosk", sr=sr, n_out=drives[i].n_out, preset="default", segments=[cases.seg(rate=float(rates[i]), semitones=float(sts[i]))])
        ref = cases.run_case(e, case, clip=clips[i].cpu().numpy()); e.close()
        same, err, snr = cases.compare(outs[i].cpu().numpy(), ref)
        assert err <= MAX_ABS_ERR and snr >= MIN_SNR_DB and same, (i, same, err, snr)
    for o in outs:
        assert bool(torch.isfinite(o).all()) and float(o.abs().max()) > 1e-3
    eng.close()
    eng = bs.BatchStretch(2, sr, preset="cheaper")
    sub = clips[:32]
    outs = eng.plan(sub, [bs.KioskDrive(n_in, [bs.segment(rate=1.0, semitones=0.0)]) for _ in sub]); eng.run(); torch.cuda.synchronize()
    for x, y in zip(sub, outs):
        assert float((x[:, 8000:-8000] - y[:, 8000:-8000]).abs().max()) <= 1e-6
    eng.close()


def test_branch_free_div_sqrt_are_ieee(bs):
    """The chain kernel's hot loop divides and takes square roots without the compiler's guarded slow path; inside the
    helpers' safe range the results must be the correctly rounded IEEE ones (numpy float32), bit for bit."""
    import torch
    lib = bs.load_library()
    n = 1 << 22
    g = torch.Generator(device="cuda").manual_seed(3)
    ex = torch.randint(-60, 40, (n,), device="cuda", generator=g).float()
    x = (torch.rand(n, device="cuda", generator=g) + 1.0) * torch.exp2(ex) * torch.where(torch.rand(n, device="cuda", generator=g) < 0.5, -1.0, 1.0)
    d = (torch.rand(n, device="cuda", generator=g) + 1.0) * torch.exp2(torch.randint(-50, 40, (n,), device="cuda", generator=g).float())
    x[::1001] = 0.0; x[5::4001] = -0.0
    x[7::5003] = 1e-44; d[11::6007] = 3e38       # outside the safe range: must be flagged
    q = torch.empty_like(x); r = torch.empty_like(x); fl = torch.zeros(n, dtype=torch.int32, device="cuda")
    assert lib.bsb_selftest_arith(x.data_ptr(), d.data_ptr(), q.data_ptr(), r.data_ptr(), fl.data_ptr(), n) == 0
    xn, dn, qn, rn, fn = (t.cpu().numpy() for t in (x, d, q, r, fl))
    with np.errstate(all="ignore"):
        qe = (xn / dn).astype(np.float32)
        re_ = np.sqrt(xn).astype(np.float32)
    ok_d = (fn & 1) == 0
    assert ok_d.mean() > 0.99 and bool((fn[7::5003] & 1).all()) and bool((fn[11::6007] & 1).all())
    assert (qn.view(np.uint32)[ok_d] == qe.view(np.uint32)[ok_d]).all()
    pos = (xn >= 0) & ((fn & 2) == 0)
    assert pos.mean() > 0.45
    assert (rn.view(np.uint32)[pos] == re_.view(np.uint32)[pos]).all()


def test_overlap_switch_does_not_change_results(bs, golden):
    """Chunk pipelining on two CUDA streams vs strictly serial kernels: identical output."""
    import torch
    case = cases.CASES["KA6"]
    clip = torch.from_numpy(cases.make_clip(case["clip"])).cuda()
    for on in (True, False):
        eng = cases.make_batch(bs, case, 2)
        eng.set_overlap(on)          # before plan(): the second record buffer is allocated at commit time
        outs = eng.plan([clip], [cases.batch_drive(bs, case, clip.shape[1])], chunk_blocks=32)
        eng.run(); eng.run()
        torch.cuda.synchronize()
        assert_matches_golden("KA6", outs[0].cpu().numpy(), golden)
        eng.close()


def test_run_host_pipelined_copies(bs, golden):
    """Host audio in, host audio out (pinned), copies pipelined with the kernels chunk by chunk: same bits."""
    import torch
    names = ["KA6", "stream_100_900", "sweep_default"]
    cs = [cases.CASES[n] for n in names]
    clips = [cases.make_clip(c["clip"]) for c in cs]
    for chunk in (24, 0):          # explicit small chunks; automatic (short first chunk for the upload, longest stream first)
        eng = cases.make_batch(bs, cs[0], 2)
        dclips = [torch.zeros(x.shape, dtype=torch.float32, device="cuda") for x in clips]
        outs = eng.plan(dclips, [cases.batch_drive(bs, c, x.shape[1]) for c, x in zip(cs, clips)], chunk_blocks=chunk)
        hin = [torch.from_numpy(x).pin_memory() for x in clips]
        hout = [torch.zeros(tuple(o.shape), dtype=torch.float32).pin_memory() for o in outs]
        for _ in range(2):
            eng.run_host(hin, hout)
            torch.cuda.synchronize()
            for n, y in zip(names, hout):
                assert_matches_golden(n, y.numpy(), golden)
        eng.close()


def test_control_trace_through_the_quantum_table(bs):
    """A worklet control trace (schedule calls with seeks, re-rates, a loop, formant changes) resolved on the host into
    the per-quantum table and run on the GPU == the oracle driven quantum by quantum like WasmProcessor.process."""
    import torch
    import test_worklet
    clip = refdrive.survey_clip(30000)
    n_out = 40000
    ref = bs.WorkletTimeline(48000.0).render(refdrive.PortEngine(), n_out, events=test_worklet._trace(), clip=clip)
    tl = bs.WorkletTimeline(48000.0); tl.addBuffers(clip)
    recs = tl.resolve(n_out, events=test_worklet._trace())
    eng = bs.BatchStretch(2, 48000.0)
    outs = eng.plan([torch.from_numpy(clip).cuda()], [bs.TableDrive(n_out, bs.WorkletTimeline.table(recs))])
    eng.run(); torch.cuda.synchronize()
    same, err, snr = cases.compare(outs[0].cpu().numpy(), ref)
    assert err <= MAX_ABS_ERR and snr >= MIN_SNR_DB and same, (same, err, snr)
    eng.close()


@pytest.mark.parametrize("kind", ["pause", "late_start", "stop"])
def test_stop_and_start_on_gpu(kind, bs):
    """Inactive time-map segments (process(q, q) on a zeroed buffer, app/SignalsmithStretch.mjs:861-869) with the silence
    gate, device path and host-audio path, against the oracle driven quantum by quantum."""
    import torch
    import test_worklet
    clip = refdrive.survey_clip(30000)
    n_out = 45000
    ref = bs.WorkletTimeline(48000.0).render(refdrive.PortEngine(), n_out, events=test_worklet._stop_start_trace(kind), clip=clip)
    tl = bs.WorkletTimeline(48000.0); tl.addBuffers(clip)
    recs = tl.resolve(n_out, events=test_worklet._stop_start_trace(kind))
    eng = bs.BatchStretch(2, 48000.0)
    dclip = torch.from_numpy(clip).cuda()
    outs = eng.plan([dclip], [bs.TableDrive(n_out, bs.WorkletTimeline.table(recs))], chunk_blocks=9)
    outs[0].fill_(7.0)                       # stale device memory must not show through behind a closed gate
    eng.run(); torch.cuda.synchronize()
    assert cases.compare(outs[0].cpu().numpy(), ref)[0] and eng.gate_events() == 0
    hin, hout = [torch.from_numpy(clip).pin_memory()], [torch.full((2, n_out), 7.0).pin_memory()]
    eng.run_host(hin, hout)
    assert cases.compare(hout[0].numpy(), ref)[0]
    eng.close()


def test_edge_cases_empty_tiny_and_ragged_on_gpu(bs):
    """Empty output, one output sample, a clip shorter than the block, ragged ends, playing past the clip's end --
    all in one batch, each equal to the oracle."""
    import torch
    rng = np.random.default_rng(5)
    specs = [(0, 3000, 1.0, 0.0), (1, 3000, 1.0, 3.0), (777, 100, 0.5, -2.0), (1441, 9000, 2.0, 5.0), (12345, 2000, 1.3, 0.0),
             (20000, 20000, 0.9, 7.0)]
    clips, drives, refs = [], [], []
    for n_out, n_in, rate, st in specs:
        clip = (0.2 * rng.standard_normal((2, n_in))).astype(np.float32)
        clips.append(torch.from_numpy(clip).cuda())
        drives.append(bs.KioskDrive(n_out, [bs.segment(rate=rate, semitones=st)]))
        e = refdrive.PortEngine()
        case = dict(drive="kiosk", sr=48000, n_out=n_out, preset="default", segments=[cases.seg(rate=rate, semitones=st)])
        refs.append(cases.run_case(e, case, clip=clip)); e.close()
    eng = bs.BatchStretch(2, 48000.0)
    outs = eng.plan(clips, drives, chunk_blocks=3)
    eng.run(); torch.cuda.synchronize()
    for (n_out, *_), o, r in zip(specs, outs, refs):
        assert tuple(o.shape) == (2, n_out) and cases.compare(o.cpu().numpy(), r)[0], n_out
    eng.close()


@pytest.mark.parametrize("channels,block,interval,split,sr", [
    (1, 1000, 250, 0, 32000), (3, 1536, 384, 1, 48000), (4, 2500, 700, 0, 44100), (5, 960, 240, 1, 96000),
    (6, 3000, 1000, 0, 48000), (7, 514, 130, 1, 22050), (8, 4800, 1920, 1, 48000), (2, 5762, 1441, 0, 48000),
    (2, 1111, 277, 1, 48000), (2, 14000, 3500, 0, 96000),
    (8, 960, 720, 0, 48000), (3, 1024, 1000, 1, 48000), (2, 960, 720, 1, 48000)])   # interval > block / 2: longStep 1, the smallest rings
def test_other_channel_counts_and_geometries_against_live_oracle(channels, block, interval, split, sr, bs):
    """Every chain_kernel<C> instantiation and the run-time FFT geometries (outer factors 1..8, odd block sizes that
    take the generic pack / overlap-add paths), each with transpose + formant shift, against the CPU oracle."""
    import torch
    rng = np.random.default_rng(channels * 1000 + block)
    n_in = int(0.9 * sr)
    clip = (0.15 * rng.standard_normal((channels, n_in))).astype(np.float32)
    t = np.arange(n_in) / sr
    for c in range(channels):
        clip[c] += (0.3 * np.sin(2 * np.pi * (180.0 + 40 * c) * t)).astype(np.float32)
    rate = float(rng.choice([0.45, 0.8, 1.0, 1.7]))
    n_out = int(0.8 * sr / max(rate, 0.6))
    segs = [cases.seg(rate=rate, semitones=float(rng.integers(-7, 8)), tonality_hz=6000.0,
                      formant_semitones=float(rng.integers(-2, 3)), formant_compensation=bool(channels % 2), formant_base_hz=0.0 if channels % 3 else 150.0)]
    case = dict(drive="kiosk", sr=sr, n_out=n_out, block=(block, interval, split), segments=segs, seed=3)
    e = refdrive.PortEngine(seed=3)
    ref = cases.run_case(e, case, clip=clip); e.close()
    eng = bs.BatchStretch(channels, sr, block_samples=block, interval_samples=interval, split_computation=bool(split))
    outs = eng.plan([torch.from_numpy(clip).cuda()], [cases.batch_drive(bs, case, n_in)], chunk_blocks=40)
    eng.run(); torch.cuda.synchronize()
    same, err, snr = cases.compare(outs[0].cpu().numpy(), ref)
    assert err <= MAX_ABS_ERR and snr >= MIN_SNR_DB and same, (same, err, snr)
    eng.close()


@pytest.mark.parametrize("channels,block,interval,split,sr", [
    (2, 5760, 1440, 0, 48000), (2, 4800, 1920, 1, 48000), (2, 9600, 2400, 1, 48000), (1, 11520, 2880, 0, 96000), (3, 960, 240, 1, 96000)])
def test_specialised_stft_kernels_on_gpu(channels, block, interval, split, sr, bs):
    """fft_fast.cuh on the device against the run-time-geometry kernels and the CPU oracle (see the CPU twin of this test)."""
    import torch
    rng = np.random.default_rng(block)
    n_in = int(0.5 * sr)
    clip = (0.2 * rng.standard_normal((channels, n_in))).astype(np.float32)
    case = dict(drive="kiosk", sr=sr, n_out=int(0.35 * sr), block=(block, interval, split), seed=3,
                segments=[cases.seg(rate=0.83, input=0.0131, semitones=3.0, formant_semitones=-2.0, formant_compensation=True)])
    e = refdrive.PortEngine(seed=3)
    ref = cases.run_case(e, case, clip=clip); e.close()
    outs = []
    for fast in (True, False):
        eng = bs.BatchStretch(channels, sr, block_samples=block, interval_samples=interval, split_computation=bool(split))
        eng.set_fast_fft(fast)
        assert eng.fast_fft_active() == fast
        o = eng.plan([torch.from_numpy(clip).cuda()], [cases.batch_drive(bs, case, n_in)], chunk_blocks=9)
        eng.run(); torch.cuda.synchronize()
        outs.append(o[0].cpu().numpy()); eng.close()
    assert cases.compare(outs[0], outs[1])[0] and cases.compare(outs[0], ref)[0]


def test_engines_of_different_geometries_side_by_side(bs, golden):
    """Two engines in one process (a batch of mixed presets = one engine per preset): a later engine with smaller shared-memory
    needs must not lower the limits the earlier one launches with."""
    import torch
    a_case, b_case = cases.CASES["KA5"], cases.CASES["lowlat_8ch_formant_auto"]
    ca, cb = cases.make_clip(a_case["clip"]), cases.make_clip(b_case["clip"])
    ea = cases.make_batch(bs, a_case, 2)
    oa = ea.plan([torch.from_numpy(ca).cuda()], [cases.batch_drive(bs, a_case, ca.shape[1])])
    eb = cases.make_batch(bs, b_case, 8)            # created after ea: block 960, much smaller kernels' shared memory
    ob = eb.plan([torch.from_numpy(cb).cuda()], [cases.batch_drive(bs, b_case, cb.shape[1])])
    ea.run(); eb.run(); ea.run(); torch.cuda.synchronize()
    assert_matches_golden("KA5", oa[0].cpu().numpy(), golden)
    assert_matches_golden("lowlat_8ch_formant_auto", ob[0].cpu().numpy(), golden)
    ea.close(); eb.close()


def test_run_host_with_streams_cut_from_one_allocation(bs):
    """Host audio for a batch whose streams lie back to back (device and host alike, equal lengths): the per-chunk copies of a
    run of streams are merged into one strided copy.  Same bits as the device-resident run; a stream of another length in the
    middle of the batch breaks the run and is copied on its own."""
    import torch
    rng = np.random.default_rng(21)
    n_in, n_out, S = 30000, 26000, 7
    lens = [(n_in, n_out)] * 3 + [(20000, 9000)] + [(n_in, n_out)] * 3
    host = torch.from_numpy((0.2 * rng.standard_normal(sum(2 * a for a, _ in lens))).astype(np.float32)).pin_memory()
    dev_in = torch.zeros_like(host, device="cuda")
    dev_out = torch.zeros(sum(2 * b for _, b in lens), device="cuda")
    hout = torch.zeros(dev_out.shape).pin_memory()
    hc, dc, do, ho, oi, oo = [], [], [], [], 0, 0
    for a, b in lens:
        hc.append(host[oi:oi + 2 * a].view(2, a)); dc.append(dev_in[oi:oi + 2 * a].view(2, a)); oi += 2 * a
        do.append(dev_out[oo:oo + 2 * b].view(2, b)); ho.append(hout[oo:oo + 2 * b].view(2, b)); oo += 2 * b
    drives = [bs.KioskDrive(b, [bs.segment(rate=0.7 + 0.1 * i, semitones=float(i - 3))]) for i, (_, b) in enumerate(lens)]
    eng = bs.BatchStretch(2, 48000.0)
    eng.plan(dc, drives, outputs=do, chunk_blocks=5)
    eng.run_host(hc, ho)
    got = hout.clone()
    dev_in.copy_(host.cuda()); dev_out.zero_()
    eng.run(); torch.cuda.synchronize()
    assert torch.equal(got, dev_out.cpu()) and float(got.abs().max()) > 1e-3
    eng.close()


@pytest.mark.parametrize("channels,n_streams", [(8, 3), (6, 5), (3, 20), (4, 70)])
def test_wide_chain_relay_on_ragged_batches(channels, n_streams, bs):
    """chain_wide_kernel with the wavefront of a stream relayed across CTAs (automatic chunking), for the batch sizes that pick
    its three shapes -- S5/S6 on separate warps with 4 warps of blocks (< 64 streams), one warp doing both with 8 (>= 64) -- on
    streams of different lengths: against the same batch run in fixed 40-block chunks (no relay; that path is pinned to the
    oracle by the tests above), and two of the streams against the CPU oracle directly."""
    import torch
    rng = np.random.default_rng(100 * channels + n_streams)
    sr, blk = 48000, (960, 240, 1)
    clips, cs = [], []
    for i in range(n_streams):
        n_in = int((0.5 + 0.9 * rng.random()) * sr)
        x = (0.15 * rng.standard_normal((channels, n_in))).astype(np.float32)
        x[i % channels] *= 3.0                                    # the maximum channel differs from stream to stream
        rate = float(rng.choice([0.6, 0.9, 1.0, 1.5]))
        clips.append(x)
        cs.append(dict(drive="kiosk", sr=sr, n_out=int(0.9 * n_in / rate), block=blk, seed=3 + i,
                       segments=[cases.seg(rate=rate, semitones=float(rng.integers(-5, 6)), formant_semitones=float(rng.integers(-2, 3)),
                                           formant_compensation=bool(i % 2), formant_base_hz=0.0 if i % 3 else 200.0)]))
    runs = []
    for chunk in (0, 40):
        eng = bs.BatchStretch(channels, sr, block_samples=blk[0], interval_samples=blk[1], split_computation=True)
        o = eng.plan([torch.from_numpy(x).cuda() for x in clips], [cases.batch_drive(bs, c, x.shape[1]) for c, x in zip(cs, clips)],
                     chunk_blocks=chunk)
        eng.run(); torch.cuda.synchronize()
        runs.append([t.cpu().numpy() for t in o]); eng.close()
    for a, b in zip(*runs):
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for i in (0, n_streams - 1):
        e = refdrive.PortEngine(seed=cs[i]["seed"])
        ref = cases.run_case(e, cs[i], clip=clips[i]); e.close()
        same, err, snr = cases.compare(runs[0][i], ref)
        assert same, (i, err, snr)
