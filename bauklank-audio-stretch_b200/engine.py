"""The reference's single-engine operator surface (Part 1 of the C ABI) as a Python object.

Same 18 calls, same meaning and order as the wasm exports the worklet drives (app/SignalsmithStretch.mjs:462-479),
so the drivers written for the reference (tests use ``oracle/refdrive.py``) run unchanged against the GPU engine.
One engine instance per process, like one wasm module instance per AudioWorkletNode.
"""
import ctypes as C

import numpy as np

from . import _capi


class StretchEngine:
    def __init__(self, seed=None, lib=None):
        self.lib = lib or _capi.load_library()
        self.channels = 0
        self.buf_len = 0
        self.buf_base = None
        if seed is not None:
            self.lib.stretch_set_seed(int(seed) & 0xFFFFFFFF)
        self._seed = seed

    # --- the reference's exports
    def presetDefault(self, ch, sr): self.lib.presetDefault(ch, sr); self._after_configure(ch)
    def presetCheaper(self, ch, sr): self.lib.presetCheaper(ch, sr); self._after_configure(ch)
    def configure(self, ch, block, interval, split=0): self.lib.configure(ch, block, interval, int(split)); self._after_configure(ch)
    def reset(self): self.lib.reset()
    def blockSamples(self): return self.lib.blockSamples()
    def intervalSamples(self): return self.lib.intervalSamples()
    def inputLatency(self): return self.lib.inputLatency()
    def outputLatency(self): return self.lib.outputLatency()
    def setTransposeFactor(self, m, tl=0.0): self.lib.setTransposeFactor(m, tl)
    def setTransposeSemitones(self, st, tl=0.0): self.lib.setTransposeSemitones(st, tl)
    def setFormantFactor(self, m, comp=False): self.lib.setFormantFactor(m, int(comp))
    def setFormantSemitones(self, st, comp=False): self.lib.setFormantSemitones(st, int(comp))
    def setFormantBase(self, f): self.lib.setFormantBase(f)
    def seek(self, n, rate): self.lib.seek(n, rate)
    def process(self, n_in, n_out): self.lib.process(n_in, n_out)
    def flush(self, n_out): self.lib.flush(n_out)

    def _after_configure(self, ch):
        self.channels = ch
        self.buf_len = 0
        if self._seed is not None:
            self.lib.stretch_set_seed(int(self._seed) & 0xFFFFFFFF)

    def setBuffers(self, ch, length):
        self.channels, self.buf_len = ch, length
        self.buf_base = self.lib.setBuffers(ch, length)
        return self.buf_base

    def io_views(self):
        """(inputs, outputs): float32 views [channels, length] over the engine-owned host buffer
        (WasmProcessor.updateBuffers, app/SignalsmithStretch.mjs:803-816)."""
        ch, n = self.channels, self.buf_len
        arr = np.ctypeslib.as_array((C.c_float * (2 * ch * n)).from_address(self.buf_base)).reshape(2 * ch, n)
        return arr[:ch], arr[ch:]

    def close(self):
        pass
