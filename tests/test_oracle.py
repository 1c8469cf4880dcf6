"""CPU: pins the oracle.  (1) the readable C port (oracle/stretch_oracle.c) reproduces every golden vector minted
from the reference's own blob (tests/golden/make_golden.py) bit-for-bit; (2) where the translated blob is present
(oracle/_ref, built from /root/reference in the build container) the two agree live, including on internal state;
(3) analytic anchors that do not depend on either translation."""
import os

import numpy as np
import pytest

import cases
from conftest import assert_matches_golden, sha
from oracle import refdrive

HAVE_REF = os.path.exists(refdrive.REF_SO)


def test_survey_clip_is_the_documented_input(golden):
    meta, _ = golden
    x = refdrive.survey_clip()
    assert sha(x) == meta["_survey_clip"]["sha256"]
    assert sha(x).startswith("3ed33f6d36b7efbc")
    np.testing.assert_allclose(x[0, :4], [0.12500143, 0.27961421, 0.18285495, 0.26621723], rtol=0, atol=1e-8)


@pytest.mark.parametrize("name", list(cases.CASES))
def test_port_reproduces_golden(name, golden):
    eng = refdrive.PortEngine(seed=cases.CASES[name].get("seed", 1))
    y = cases.run_case(eng, cases.CASES[name])
    eng.close()
    assert_matches_golden(name, y, golden)


@pytest.mark.parametrize("name", list(cases.SHIM_CASES))
def test_port_reproduces_golden_gate_and_flush(name, golden):
    """Silence gate (W#48 7838-7943) and flush (W#46, restated from the bytecode) against vectors minted from the blob."""
    eng = refdrive.PortEngine(seed=cases.SHIM_CASES[name].get("seed", 1))
    y = cases.run_case(eng, cases.SHIM_CASES[name])
    eng.close()
    assert_matches_golden(name, y, golden)


@pytest.mark.skipif(not HAVE_REF, reason="translated reference blob not built (oracle/_ref)")
def test_port_flush_equals_translated_blob_live():
    """flush() at many positions of the block cycle, then more processing: port == blob bit for bit."""
    for preset, n_in, n_out in (("default", 300, 700), ("cheaper", 200, 200)):
        for calls in (20, 23, 26, 29):
            case = dict(drive="stream", clip=("survey", 16000), sr=48000, n_in=n_in, n_out=n_out, preset=preset, seed=5,
                        flush=(calls, 2000), segments=[cases.seg(semitones=3.0)])
            a = refdrive.RefEngine(seed=5); b = refdrive.PortEngine(seed=5)
            ya, yb = cases.run_case(a, case), cases.run_case(b, case)
            assert cases.compare(ya, yb)[0], (preset, calls)
            a.close(); b.close()


SURVEY_PREFIX = dict(KA1="83b7dd548f69c080", KA2="4b020352d89f76a1", KA3="75865524063b07dc", KA4="2d2d72427453da4e",
                     KA5="18e7b06e2a8eb64e", KA6="05a59d3d8ea0ae29")


def test_golden_file_holds_the_survey_known_answers(golden):
    meta, _ = golden
    for k, pre in SURVEY_PREFIX.items():
        assert meta[k]["sha256"].startswith(pre)


def test_analytic_anchors():
    """KA1 = the input delayed by inputLatency+outputLatency (5760) to 5e-7; KA3 = the input, undelayed, to 1e-6."""
    x = refdrive.survey_clip()
    e = refdrive.PortEngine()
    y1 = cases.run_case(e, cases.CASES["KA1"], clip=x)
    lat = e.inputLatency() + e.outputLatency()
    assert lat == 5760
    assert np.abs(y1[:, lat:] - x[:, :y1.shape[1] - lat]).max() <= 5e-7
    y3 = cases.run_case(e, cases.CASES["KA3"], clip=x)
    n = y3.shape[1]
    assert np.abs(y3[:, 8000:n - 8000] - x[:, 8000:n - 8000]).max() <= 1e-6
    e.close()


def test_geometry_table():
    """SURVEY.md section 8 size table (W#25 arithmetic)."""
    e = refdrive.PortEngine()
    e.presetDefault(2, 48000.0)
    assert (e.blockSamples(), e.intervalSamples(), e.inputLatency(), e.outputLatency()) == (5760, 1440, 2880, 2880)
    e.presetCheaper(2, 48000.0)
    assert (e.blockSamples(), e.intervalSamples(), e.inputLatency(), e.outputLatency()) == (4800, 1920, 2400, 4320)
    e.presetDefault(8, 96000.0)
    assert (e.blockSamples(), e.intervalSamples()) == (11520, 2880)
    e.configure(8, 960, 240, 1)
    assert (e.inputLatency(), e.outputLatency()) == (480, 720)
    e.close()


@pytest.mark.skipif(not HAVE_REF, reason="translated reference blob not built (oracle/_ref)")
@pytest.mark.parametrize("name", ["KA4", "rng_low_rate", "lowlat_8ch_formant_auto", "stream_480_512_cheaper"])
def test_port_equals_translated_blob_live(name):
    case = cases.CASES[name]
    a = refdrive.RefEngine(seed=case.get("seed", 1))
    b = refdrive.PortEngine(seed=case.get("seed", 1))
    ya, yb = cases.run_case(a, case), cases.run_case(b, case)
    assert cases.compare(ya, yb)[0]
    a.close(); b.close()


@pytest.mark.skipif(not HAVE_REF, reason="translated reference blob not built (oracle/_ref)")
def test_port_window_matches_blob_memory():
    """Kaiser window with perfect-reconstruction scaling (W#36), read from the blob's linear memory (Appendix A)."""
    a = refdrive.RefEngine(); b = refdrive.PortEngine()
    for eng in (a, b):
        eng.presetCheaper(2, 48000.0)
    L = a.blockSamples()
    wa = a.vec(6460, np.float32, L)
    import ctypes as C
    b._select()
    b.lib.so_window.restype = C.POINTER(C.c_float)
    wb = np.ctypeslib.as_array(b.lib.so_window(), (L,))
    assert (wa.view(np.uint32) == wb.view(np.uint32)).all()
    a.close(); b.close()


def test_seed_matters_only_below_half_rate():
    """Q5: the RNG is consulted only when timeFactor > 2."""
    x = refdrive.survey_clip(20000)
    def run(rate, seed):
        e = refdrive.PortEngine(seed=seed)
        c = dict(drive="kiosk", sr=48000, n_out=20000, preset="default", segments=[cases.seg(rate=rate, semitones=2.0)])
        y = cases.run_case(e, c, clip=x); e.close(); return y
    assert cases.compare(run(0.5, 1), run(0.5, 99))[0]
    assert not cases.compare(run(0.4, 1), run(0.4, 99))[0]
