"""ctypes binding + drivers for the CPU checkers under ``oracle/``.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  The product package never does.

Two engines share one Python surface (the 18-name C ABI of ``app/SignalsmithStretch.mjs:462-479``):

* ``RefEngine``    -- ``oracle/_ref/libstretch_ref.so``: the reference's own WASM blob, mechanically translated
                      to C by ``oracle/wasm2c.py`` (bit-faithful: "kind = reference").
* ``PortEngine``   -- ``oracle/libstretch_oracle.so``: the readable C restatement in ``oracle/stretch_oracle.c``
                      ("kind = port"), pinned against RefEngine by ``tests/test_oracle.py``.

Drivers re-enact ``WasmProcessor.process`` (``app/SignalsmithStretch.mjs:826-954``):
``kiosk_drive`` = buffer playback branch (:883-943, ``_seek`` + ``_process(0, q)`` per quantum);
``stream_drive`` = live-input branch (:870-882, ``_process(q, q)``), generalised to (in, out) per call.
"""
import ctypes as C
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libstretch_ref.so")
PORT_SO = os.path.join(HERE, "libstretch_oracle.so")


def js_round(x):
    """JavaScript Math.round: floor(x + 0.5) (app/SignalsmithStretch.mjs:897 ``Math.round(inputTime*sampleRate)``)."""
    return int(math.floor(x + 0.5))


class _Engine:
    """Common surface; subclasses provide self.lib, self.h and a name prefix."""
    prefix = ""

    def _bind(self, lib):
        p = self.prefix
        sig = {
            "setBuffers": (C.c_uint32, [C.c_uint32, C.c_uint32]),
            "blockSamples": (C.c_uint32, []), "intervalSamples": (C.c_uint32, []),
            "inputLatency": (C.c_uint32, []), "outputLatency": (C.c_uint32, []),
            "reset": (None, []),
            "presetDefault": (None, [C.c_uint32, C.c_float]), "presetCheaper": (None, [C.c_uint32, C.c_float]),
            "configure": (None, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
            "setTransposeFactor": (None, [C.c_float, C.c_float]),
            "setTransposeSemitones": (None, [C.c_float, C.c_float]),
            "setFormantFactor": (None, [C.c_float, C.c_uint32]),
            "setFormantSemitones": (None, [C.c_float, C.c_uint32]),
            "setFormantBase": (None, [C.c_float]),
            "seek": (None, [C.c_uint32, C.c_double]),
            "process": (None, [C.c_uint32, C.c_uint32]),
            "flush": (None, [C.c_uint32]),
        }
        self._fn = {}
        for name, (res, args) in sig.items():
            f = getattr(lib, p + name)
            f.restype = res
            f.argtypes = args
            self._fn[name] = f

    def _call(self, name, *a):
        self._select()
        return self._fn[name](*a)

    # -- the reference's operator surface
    def presetDefault(self, ch, sr): self._call("presetDefault", ch, sr); self._cfg(ch)
    def presetCheaper(self, ch, sr): self._call("presetCheaper", ch, sr); self._cfg(ch)
    def configure(self, ch, block, interval, split=0): self._call("configure", ch, block, interval, int(split)); self._cfg(ch)
    def reset(self): self._call("reset")
    def blockSamples(self): return self._call("blockSamples")
    def intervalSamples(self): return self._call("intervalSamples")
    def inputLatency(self): return self._call("inputLatency")
    def outputLatency(self): return self._call("outputLatency")
    def setTransposeFactor(self, m, tl=0.0): self._call("setTransposeFactor", m, tl)
    def setTransposeSemitones(self, st, tl=0.0): self._call("setTransposeSemitones", st, tl)
    def setFormantFactor(self, m, comp=False): self._call("setFormantFactor", m, int(comp))
    def setFormantSemitones(self, st, comp=False): self._call("setFormantSemitones", st, int(comp))
    def setFormantBase(self, f): self._call("setFormantBase", f)
    def seek(self, n, rate): self._call("seek", n, rate)
    def process(self, n_in, n_out): self._call("process", n_in, n_out)
    def flush(self, n_out): self._call("flush", n_out)

    def _cfg(self, ch):
        self.channels = ch
        self.buf_len = 0

    def setBuffers(self, ch, length):
        """WasmProcessor.updateBuffers (app/SignalsmithStretch.mjs:803-816): in[c] at base+len*c, out[c] at base+len*(c+ch)."""
        self.channels = ch
        self.buf_len = length
        self.buf_base = self._call("setBuffers", ch, length)
        return self.buf_base

    def io_views(self):
        """(inputs, outputs) as float32 views [ch, buf_len] over the engine-owned buffer."""
        raise NotImplementedError


class RefEngine(_Engine):
    prefix = "ref_"
    _lib = None

    def __init__(self, seed=1):
        if RefEngine._lib is None:
            lib = C.CDLL(REF_SO)
            lib.ref_new.restype = C.c_void_p
            lib.ref_new.argtypes = [C.c_uint32]
            lib.ref_free.argtypes = [C.c_void_p]
            lib.ref_select.argtypes = [C.c_void_p]
            lib.ref_memory.restype = C.c_void_p
            lib.ref_memory.argtypes = [C.c_void_p]
            lib.ref_memory_bytes.restype = C.c_uint32
            lib.ref_memory_bytes.argtypes = [C.c_void_p]
            RefEngine._lib = lib
        self.lib = RefEngine._lib
        self._bind(self.lib)
        self.h = self.lib.ref_new(seed)

    def _select(self):
        self.lib.ref_select(self.h)

    def close(self):
        if self.h:
            self.lib.ref_free(self.h)
            self.h = None

    def memory(self):
        """Whole wasm linear memory as a uint8 numpy view (SURVEY.md Appendix A gives the map)."""
        n = self.lib.ref_memory_bytes(self.h)
        base = self.lib.ref_memory(self.h)
        return np.ctypeslib.as_array((C.c_uint8 * n).from_address(base))

    def io_views(self):
        mem = self.memory()
        ch, n = self.channels, self.buf_len
        arr = mem[self.buf_base:self.buf_base + 8 * ch * n].view(np.float32).reshape(2 * ch, n)
        return arr[:ch], arr[ch:]

    # helpers to read engine internals
    def u32(self, addr):
        return int(self.memory()[addr:addr + 4].view(np.uint32)[0])

    def vec(self, addr, dtype, count=None):
        """Read a libc++ std::vector {begin,end,cap} stored at ``addr``."""
        mem = self.memory()
        b, e = (int(v) for v in mem[addr:addr + 8].view(np.uint32))
        raw = mem[b:e]
        out = raw.view(dtype)
        return out[:count] if count is not None else out


class PortEngine(_Engine):
    """The C restatement: handle-based ``so_*`` functions, same names."""
    prefix = "so_"
    _lib = None

    def __init__(self, seed=1):
        if PortEngine._lib is None:
            lib = C.CDLL(PORT_SO)
            lib.so_new.restype = C.c_void_p
            lib.so_new.argtypes = [C.c_uint32]
            lib.so_free.argtypes = [C.c_void_p]
            lib.so_select.argtypes = [C.c_void_p]
            lib.so_buffers.restype = C.c_void_p
            lib.so_buffers.argtypes = [C.c_void_p]
            PortEngine._lib = lib
        self.lib = PortEngine._lib
        self._bind(self.lib)
        self.h = self.lib.so_new(seed)

    def _select(self):
        self.lib.so_select(self.h)

    def close(self):
        if self.h:
            self.lib.so_free(self.h)
            self.h = None

    def io_views(self):
        ch, n = self.channels, self.buf_len
        base = self.lib.so_buffers(self.h)
        arr = np.ctypeslib.as_array((C.c_float * (2 * ch * n)).from_address(base)).reshape(2 * ch, n)
        return arr[:ch], arr[ch:]


# ------------------------------------------------------------------------------------------------
# drivers
def setup(engine, channels, sample_rate, preset="default", block=None, interval=None, split=0):
    """WasmProcessor.configure + updateBuffers (app/SignalsmithStretch.mjs:786-816)."""
    if block is not None:
        engine.configure(channels, block, interval if interval is not None else int(round(block * 0.25)), split)
        engine.reset()
    elif preset == "cheaper":
        engine.presetCheaper(channels, sample_rate)
    else:
        engine.presetDefault(channels, sample_rate)
    buf_len = engine.inputLatency() + engine.outputLatency()
    engine.setBuffers(channels, buf_len)
    return buf_len


def apply_params(engine, sr, semitones=0.0, tonality_hz=8000.0, formant_semitones=0.0, formant_comp=False,
                 formant_base_hz=0.0):
    """The three setter calls made every quantum (app/SignalsmithStretch.mjs:847-849)."""
    engine.setTransposeSemitones(semitones, tonality_hz / sr)
    engine.setFormantSemitones(formant_semitones, formant_comp)
    engine.setFormantBase(formant_base_hz / sr)


def kiosk_fill(buf, clip, end):
    """Fill buf[ch, n] with the n samples of clip ending at ``end``, zero outside the clip
    (app/SignalsmithStretch.mjs:900-931, single stored audio buffer starting at sample 0)."""
    n = buf.shape[1]
    start = end - n
    buf[:] = 0
    lo = max(start, 0)
    hi = min(end, clip.shape[1])
    if hi > lo:
        buf[:, lo - start:hi - start] = clip[:, lo:hi]


def kiosk_drive(engine, clip, sr, n_out, rate=1.0, quantum=128, params=None, preset="default", block=None,
                interval=None, split=0, param_fn=None, seg_input=0.0, seg_output=0.0):
    """Buffer-playback drive: one time-map segment {input, output, rate}; per quantum seek + process(0, q).

    ``param_fn(k, t_out)`` may return a dict overriding rate/semitones/... for quantum k, emulating a new
    ``schedule()`` call whose segment starts exactly at that quantum (input re-derived by extrapolation, as
    ``remoteMethods.schedule`` does at app/SignalsmithStretch.mjs:656-701).
    """
    ch = clip.shape[0]
    buf_len = setup(engine, ch, sr, preset, block, interval, split)
    in_lat_s = engine.inputLatency() / sr
    out_lat_s = engine.outputLatency() / sr
    params = dict(params or {})
    seg = dict(input=seg_input, output=seg_output, rate=rate)
    out = np.zeros((ch, n_out), np.float32)
    ends = []
    k = 0
    pos = 0
    while pos < n_out:
        q = min(quantum, n_out - pos)
        current_time = (k * quantum) / sr
        output_time = current_time + out_lat_s
        if param_fn is not None:
            upd = param_fn(k, output_time)
            if upd:
                upd = dict(upd)
                if "rate" in upd and upd["rate"] != seg["rate"]:
                    new_in = seg["input"] + (output_time - seg["output"]) * seg["rate"]
                    seg = dict(input=new_in, output=output_time, rate=upd.pop("rate"))
                else:
                    upd.pop("rate", None)
                params.update(upd)
        apply_params(engine, sr, **params)
        input_time = seg["input"] + (output_time - seg["output"]) * seg["rate"]
        input_time += in_lat_s
        end = js_round(input_time * sr)
        ends.append(end)
        ins, outs = engine.io_views()
        kiosk_fill(ins, clip, end)
        engine.seek(buf_len, seg["rate"])
        engine.process(0, q)
        ins, outs = engine.io_views()
        out[:, pos:pos + q] = outs[:, :q]
        pos += q
        k += 1
    return out, ends


def stream_drive(engine, clip, sr, n_in=512, n_out=512, params=None, preset="default", block=None, interval=None,
                 split=0, param_fn=None):
    """Streaming drive: process(n_in, n_out) over the clip (whole calls only)."""
    ch = clip.shape[0]
    setup(engine, ch, sr, preset, block, interval, split)
    engine.setBuffers(ch, max(n_in, n_out))
    params = dict(params or {})
    calls = clip.shape[1] // n_in
    out = np.zeros((ch, calls * n_out), np.float32)
    for k in range(calls):
        if param_fn is not None:
            upd = param_fn(k, k * n_out / sr)
            if upd:
                params.update(upd)
        apply_params(engine, sr, **params)
        ins, outs = engine.io_views()
        ins[:, :n_in] = clip[:, k * n_in:(k + 1) * n_in]
        engine.process(n_in, n_out)
        ins, outs = engine.io_views()
        out[:, k * n_out:(k + 1) * n_out] = outs[:, :n_out]
    return out


# ------------------------------------------------------------------------------------------------
# synthetic inputs
def survey_clip(n=96000):
    """The SURVEY.md section 8c known-answer input: 0.25*LCG noise + 0.5*triangle, two channels."""
    def chan(seed, period):
        s = seed
        noise = np.empty(n, np.float64)
        for i in range(n):
            noise[i] = s / 2147483648.0 - 0.5
            s = (1103515245 * s + 12345) % 2147483648
        idx = np.arange(n)
        tri = 2.0 * np.abs(((idx % period) / period) - 0.5) - 0.5
        return (0.25 * noise + 0.5 * tri).astype(np.float32)
    return np.stack([chan(12345, 109), chan(54321, 173)])


def sweep_clip(seconds=30.0, sr=48000, channels=2, seed=1234, noise=0.05):
    """BASELINE config 1: 0.5*sin log sweep 50 Hz -> 16 kHz (reversed on odd channels) + 0.05*N(0,1)."""
    n = int(round(seconds * sr))
    t = np.arange(n) / sr
    f0, f1 = 50.0, 16000.0
    k = math.log(f1 / f0) / seconds
    phase = 2 * math.pi * f0 * (np.exp(k * t) - 1.0) / k
    up = 0.5 * np.sin(phase)
    rng = np.random.default_rng(seed)
    out = np.empty((channels, n), np.float32)
    for c in range(channels):
        base = up if c % 2 == 0 else up[::-1]
        out[c] = (base + noise * rng.standard_normal(n)).astype(np.float32)
    return out
