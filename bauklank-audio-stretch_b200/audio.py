"""Audio file -> the planar float32 channel buffers the engine takes.

The kiosk hands the browser a file and gets ``AudioBuffer`` channels back from ``audioContext.decodeAudioData`` -- decoded
and resampled to the context's sample rate -- which go to ``addBuffers`` as they are (app/multi/app.mjs:365-384).  This is
the host-side counterpart: ``decode_audio(path_or_bytes, sample_rate)`` returns ``[channels, n]`` float32 at the engine's
rate.  Web Audio does not specify its resampler, so this step has no bit-level reference; it is a band-limited polyphase
resampler (Kaiser-windowed sinc), and everything downstream of it is the parity-checked path.

Containers: RIFF/WAVE (PCM 8/16/24/32-bit, IEEE float 32/64, WAVE_FORMAT_EXTENSIBLE), ``.npy`` ([channels, n] or [n]) and
headerless float32.  Compressed formats (the kiosk's MP3s) need a decoder this package does not carry: decode them to WAV first.
"""
import io
import math
import struct

import numpy as np


def read_wav(data):
    """RIFF/WAVE bytes -> (float32 [channels, n] in [-1, 1), sample_rate)."""
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and len(body) >= 26:          # WAVE_FORMAT_EXTENSIBLE: the real tag is the sub-format's first word
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError("WAVE file without fmt/data chunk")
    tag, ch, sr, bits = fmt
    if tag == 1:       # integer PCM
        if bits == 8:
            x = (np.frombuffer(pcm, np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(pcm, "<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(pcm[:len(pcm) // 3 * 3], np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = ((v ^ 0x800000) - 0x800000).astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(pcm, "<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise ValueError("unsupported PCM width %d" % bits)
    elif tag == 3:     # IEEE float
        x = np.frombuffer(pcm, "<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError("unsupported WAVE format tag %d (decode compressed audio to PCM first)" % tag)
    n = len(x) // ch
    return np.ascontiguousarray(x[:n * ch].reshape(n, ch).T), int(sr)


def write_wav(path_or_file, audio, sample_rate, float32=True):
    """[channels, n] float32 -> RIFF/WAVE (IEEE float 32, or 16-bit PCM)."""
    a = np.asarray(audio, np.float32)
    if a.ndim == 1:
        a = a[None, :]
    ch, n = a.shape
    inter = np.ascontiguousarray(a.T)
    if float32:
        tag, bits, body = 3, 32, inter.astype("<f4").tobytes()
    else:
        tag, bits, body = 1, 16, np.clip(np.round(inter * 32768.0), -32768, 32767).astype("<i2").tobytes()
    hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + len(body), b"WAVE", b"fmt ", 16, tag, ch, int(sample_rate),
                      int(sample_rate) * ch * bits // 8, ch * bits // 8, bits, b"data", len(body))
    if hasattr(path_or_file, "write"):
        path_or_file.write(hdr + body)
    else:
        with open(path_or_file, "wb") as f:
            f.write(hdr + body)


def resample(audio, rate_in, rate_out, taps_per_phase=32, beta=9.0):
    """Band-limited resampling [channels, n] rate_in -> rate_out: polyphase Kaiser-windowed sinc, cutoff at the lower Nyquist,
    unity pass-band gain, zero phase (output sample m sits at input time m * rate_in / rate_out)."""
    a = np.asarray(audio, np.float32)
    if a.ndim == 1:
        a = a[None, :]
    rate_in, rate_out = int(rate_in), int(rate_out)
    if rate_in == rate_out:
        return np.ascontiguousarray(a)
    g = math.gcd(rate_in, rate_out)
    up, down = rate_out // g, rate_in // g
    cutoff = 1.0 / max(up, down)                       # in units of the up-sampled Nyquist
    half = taps_per_phase * max(up, down)
    t = np.arange(-half, half + 1, dtype=np.float64)
    h = cutoff * np.sinc(cutoff * t) * np.kaiser(2 * half + 1, beta) * up
    n_out = (a.shape[1] * up + down - 1) // down
    out = np.zeros((a.shape[0], n_out), np.float32)
    # output m = sum_k h[m*down - k*up] x[k]; for each phase p = (m*down) % up the taps form a fixed filter
    m = np.arange(n_out, dtype=np.int64)
    pos = m * down
    k0 = pos // up                                     # newest input sample at or before the output instant
    phase = pos - k0 * up
    span = (half // up) + 1
    ks = np.arange(-span, span + 1, dtype=np.int64)
    for c in range(a.shape[0]):
        x = np.concatenate([np.zeros(span + 1, np.float64), a[c].astype(np.float64), np.zeros(span + 2, np.float64)])
        acc = np.zeros(n_out, np.float64)
        for dk in ks:
            tap = phase - dk * up + half               # index into h of (pos - (k0 + dk) * up)
            ok = (tap >= 0) & (tap <= 2 * half)
            acc += np.where(ok, h[np.clip(tap, 0, 2 * half)], 0.0) * x[k0 + dk + span + 1]
        out[c] = acc.astype(np.float32)
    return out


def decode_audio(source, sample_rate, channels=None):
    """File path / bytes / file object -> float32 [channels, n] at ``sample_rate`` (what decodeAudioData + getChannelData give
    the kiosk).  ``channels``: up- or down-mix like an AudioBuffer played into a node of that width (mono <-> stereo only)."""
    if isinstance(source, (bytes, bytearray)):
        data = bytes(source)
    elif hasattr(source, "read"):
        data = source.read()
    else:
        if str(source).endswith(".npy"):
            a = np.load(source).astype(np.float32)
            return _mix(np.ascontiguousarray(a[None, :] if a.ndim == 1 else a), channels)
        with open(source, "rb") as f:
            data = f.read()
    if data[:4] == b"RIFF":
        a, sr = read_wav(data)
        a = resample(a, sr, sample_rate)
    elif data[:6] == b"\x93NUMPY":
        a = np.load(io.BytesIO(data)).astype(np.float32)
        a = a[None, :] if a.ndim == 1 else a
    else:
        raise ValueError("unknown container (WAVE and .npy are read here; decode compressed audio to WAVE first)")
    return _mix(np.ascontiguousarray(a), channels)


def _mix(a, channels):
    if channels is None or channels == a.shape[0]:
        return a
    if a.shape[0] == 1:
        return np.ascontiguousarray(np.repeat(a, channels, axis=0))
    if channels == 1:
        return np.ascontiguousarray(a.mean(axis=0, keepdims=True).astype(np.float32))
    raise ValueError("cannot mix %d channels to %d" % (a.shape[0], channels))
