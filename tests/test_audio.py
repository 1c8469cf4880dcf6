"""CPU: audio file decoding / resampling (the kiosk's decodeAudioData step; no bit-level reference --
Web Audio leaves the resampler unspecified -- so these are property tests)."""
import io

import numpy as np
import pytest

import bauklank_audio_stretch_b200 as bs
from bauklank_audio_stretch_b200 import audio


def _tone(sr, f, n, ch=2):
    t = np.arange(n) / sr
    return np.stack([0.5 * np.sin(2 * np.pi * f * (1 + 0.5 * c) * t) for c in range(ch)]).astype(np.float32)


@pytest.mark.parametrize("float32", [True, False])
def test_wav_round_trip(float32):
    x = _tone(44100, 440.0, 5000)
    f = io.BytesIO()
    audio.write_wav(f, x, 44100, float32=float32)
    y, sr = audio.read_wav(f.getvalue())
    assert sr == 44100 and y.shape == x.shape
    assert np.abs(y - x).max() <= (0 if float32 else 1.0 / 32768 + 1e-7)


def test_wav_24_bit_and_extensible_header():
    import struct
    x = np.array([[0.5, -0.25, 0.999, -1.0]], np.float32)
    v = np.round(x[0] * 8388608.0).clip(-8388608, 8388607).astype(np.int32)
    body = b"".join(struct.pack("<i", int(s))[:3] for s in v)
    fmt = struct.pack("<HHIIHH", 0xFFFE, 1, 48000, 48000 * 3, 3, 24) + struct.pack("<HHI", 22, 24, 4) + struct.pack("<H", 1) + b"\0" * 14
    data = b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + 8 + len(body)) + b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(body)) + body
    y, sr = audio.read_wav(data)
    assert sr == 48000 and np.abs(y - x).max() <= 1.0 / 8388608 + 1e-7


@pytest.mark.parametrize("sr_in,sr_out", [(44100, 48000), (48000, 44100), (96000, 48000), (22050, 48000)])
def test_resample_keeps_a_tone_and_its_length(sr_in, sr_out):
    f0, n = 1000.0, sr_in // 10
    x = _tone(sr_in, f0, n, ch=1)
    y = audio.resample(x, sr_in, sr_out)
    assert abs(y.shape[1] - n * sr_out / sr_in) <= 1
    t = np.arange(y.shape[1]) / sr_out
    want = 0.5 * np.sin(2 * np.pi * f0 * t)
    mid = slice(400, y.shape[1] - 400)
    assert np.abs(y[0, mid] - want[mid]).max() < 2e-4            # same tone, same phase (zero-phase filter), unity gain


def test_resample_rejects_what_would_alias():
    sr_in, sr_out = 96000, 48000
    x = _tone(sr_in, 30000.0, 6000, ch=1)                        # above the new Nyquist
    y = audio.resample(x, sr_in, sr_out)
    assert np.abs(y[0, 500:-500]).max() < 1e-3


def test_decode_audio_end_to_end(tmp_path):
    x = _tone(44100, 300.0, 44100 // 10, ch=1)
    p = tmp_path / "clip.wav"
    audio.write_wav(str(p), x, 44100)
    y = bs.decode_audio(str(p), 48000, channels=2)
    assert y.dtype == np.float32 and y.shape[0] == 2 and abs(y.shape[1] - 4800) <= 1 and y.flags["C_CONTIGUOUS"]
    assert np.array_equal(y[0], y[1])
    with pytest.raises(ValueError):
        bs.decode_audio(b"ID3\x03junk", 48000)
