"""Batched stretch engine: many independent streams with device-resident audio (Part 2 of the C ABI).

Host code is plumbing only -- it describes each stream's *drive* the way the reference's worklet would perform it
(``WasmProcessor.process``, app/SignalsmithStretch.mjs:826-954) and hands device pointers to the C ABI; the block
schedule is compiled in C++ (``csrc/control.hpp``) and executed by the CUDA kernels (``csrc/kernels.cuh``).
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _capi


def segment(output=0.0, input=0.0, rate=1.0, semitones=0.0, tonality_hz=8000.0, formant_semitones=0.0,
            formant_compensation=False, formant_base_hz=0.0, loop_start=0.0, loop_end=0.0, active=True,
            transpose_factor=None, formant_factor=None):
    """One time-map entry; defaults are the worklet's initial segment (app/SignalsmithStretch.mjs:587-600),
    except ``active`` (a playing stream).  ``transpose_factor`` / ``formant_factor``: drive the engine's
    ``setTransposeFactor`` / ``setFormantFactor`` exports (:472, :474) with a frequency multiplier instead of the
    semitone setters the worklet calls."""
    nan = float("nan")
    return _capi.Segment(float(output), float(input), float(rate), float(semitones), float(tonality_hz),
                         float(formant_semitones), float(formant_base_hz), float(loop_start), float(loop_end),
                         1 if active else 0, 1 if formant_compensation else 0,
                         nan if transpose_factor is None else float(transpose_factor),
                         nan if formant_factor is None else float(formant_factor))


@dataclass
class KioskDrive:
    """Buffer playback: per render quantum ``seek(bufferLength, rate); process(0, quantum)`` (:883-943)."""
    n_out: int
    segments: Sequence = field(default_factory=lambda: [segment()])
    quantum: int = 128
    seed: int = 1


@dataclass
class StreamingDrive:
    """``process(n_in, n_out)`` ``n_calls`` times over a contiguous input (:870-882 generalised)."""
    n_in: int
    n_out: int
    n_calls: int
    segments: Sequence = field(default_factory=lambda: [segment()])
    seed: int = 1

    @property
    def total_out(self):
        return self.n_out * self.n_calls


@dataclass
class TableDrive:
    """Buffer playback resolved quantum by quantum (``worklet.WorkletTimeline.resolve`` -> ``table``)."""
    n_out: int
    table: object            # ctypes array of _capi.Quantum
    quantum: int = 128
    seed: int = 1


@dataclass
class TraceDrive:
    """Buffer playback under a control trace: ``events`` = ctypes array of ``bsb_trace_event`` (``trace_events``), each a
    ``schedule()`` call made before a render quantum; the time map is kept inside the library exactly like the worklet keeps
    it (``bsb_add_kiosk_trace``)."""
    n_out: int
    events: object
    quantum: int = 128
    seed: int = 1


def trace_events(events):
    """[(quantum_index, "schedule", (args_dict,))] -- what ``ControllerMapper.trace_to_events`` produces and
    ``WorkletTimeline.resolve`` consumes -- as a ctypes array of ``bsb_trace_event``.  Keys of args_dict are the worklet's
    (active input rate semitones tonalityHz formantSemitones formantCompensation formantBaseHz loopStart loopEnd outputTime);
    a missing key is "not in the call's object" (inherited or derived, see ``bsb_trace_event``)."""
    nan = float("nan")
    events = sorted(events, key=lambda e: e[0])
    for _, method, args in events:
        if method != "schedule" or len(args) != 1:
            raise ValueError("only schedule(obj) calls can be compiled into a trace (use TableDrive for the rest)")
    objs = [e[2][0] for e in events]
    a = np.zeros(len(events), dtype=_TRACE_DTYPE)       # same layout as bsb_trace_event (checked at import)
    a["quantum"] = [int(e[0]) for e in events]
    for field, key in _TRACE_KEYS:
        a[field] = [nan if o.get(key) is None else float(o[key]) for o in objs]
    a["active"] = [-1 if "active" not in o else int(bool(o["active"])) for o in objs]
    a["formant_compensation"] = [-1 if "formantCompensation" not in o else int(bool(o["formantCompensation"])) for o in objs]
    arr = (_capi.TraceEvent * len(events))()
    if len(events):
        C.memmove(arr, a.ctypes.data, a.nbytes)
    return arr


_TRACE_KEYS = (("output_time", "outputTime"), ("input", "input"), ("rate", "rate"), ("semitones", "semitones"), ("loop_start", "loopStart"),
               ("loop_end", "loopEnd"), ("tonality_hz", "tonalityHz"), ("formant_semitones", "formantSemitones"),
               ("formant_base_hz", "formantBaseHz"), ("transpose_factor", "transposeFactor"), ("formant_factor", "formantFactor"))
_TRACE_DTYPE = np.dtype([(n, {C.c_longlong: "<i8", C.c_double: "<f8", C.c_int32: "<i4"}[t]) for n, t in _capi.TraceEvent._fields_], align=True)
assert _TRACE_DTYPE.itemsize == C.sizeof(_capi.TraceEvent) and all(_TRACE_DTYPE.fields[n][1] == getattr(_capi.TraceEvent, n).offset for n, _ in _capi.TraceEvent._fields_)


def _ptr(a):
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


def _alloc_like(clip, channels, n):
    if hasattr(clip, "data_ptr"):
        import torch
        return torch.zeros((channels, n), dtype=torch.float32, device=clip.device)
    return np.zeros((channels, n), np.float32)


class BatchStretch:
    """N independent streams sharing one engine configuration (presetDefault / presetCheaper / configure)."""

    def __init__(self, channels, sample_rate=48000.0, preset="default", block_samples=None, interval_samples=None,
                 split_computation=False, lib=None):
        self.lib = lib or _capi.load_library()
        self.channels = int(channels)
        self.sample_rate = float(sample_rate)
        if block_samples is not None:
            if interval_samples is None:
                interval_samples = int(round(block_samples * 0.25))
            self.h = self.lib.bsb_create(channels, int(block_samples), int(interval_samples),
                                         1 if split_computation else 0, self.sample_rate)
        else:
            self.h = self.lib.bsb_create_preset(channels, self.sample_rate, 1 if preset == "cheaper" else 0)
        if not self.h:
            raise RuntimeError("bsb_create failed (no CUDA device, or unsupported configuration)")
        self._keep = []
        self.outputs = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.bsb_destroy(self.h)
            self.h = None

    __del__ = close

    # the reference's getters
    def blockSamples(self): return self.lib.bsb_block_samples(self.h)
    def intervalSamples(self): return self.lib.bsb_interval_samples(self.h)
    def inputLatency(self): return self.lib.bsb_input_latency(self.h)
    def outputLatency(self): return self.lib.bsb_output_latency(self.h)
    def fftSamples(self): return self.lib.bsb_fft_samples(self.h)
    def bands(self): return self.lib.bsb_bands(self.h)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("bauklank_stretch: " + self.lib.bsb_last_error(self.h).decode())

    def plan(self, clips, drives, chunk_blocks=0, outputs=None):
        """clips[i]: float32 [channels, len] in device memory (torch CUDA tensor); drives[i]: Kiosk/StreamingDrive.
        Returns the list of output tensors [channels, n_out] (allocated here unless given)."""
        assert len(clips) == len(drives) and len(clips) > 0
        self._check(self.lib.bsb_begin(self.h, len(clips)))
        self._keep = []
        outs = []
        for i, (clip, d) in enumerate(zip(clips, drives)):
            assert tuple(clip.shape)[0] == self.channels and len(clip.shape) == 2
            if hasattr(clip, "is_contiguous"):
                assert clip.is_contiguous() and str(clip.dtype) == "torch.float32"
            else:
                assert clip.flags["C_CONTIGUOUS"] and clip.dtype == np.float32
            clip_len = int(clip.shape[1])
            if isinstance(d, TraceDrive):
                n_out = int(d.n_out)
                out = outputs[i] if outputs is not None else _alloc_like(clip, self.channels, n_out)
                self._check(self.lib.bsb_add_kiosk_trace(self.h, i, _ptr(clip), clip_len, _ptr(out), n_out, int(d.quantum),
                                                         d.events, len(d.events), int(d.seed) & 0xFFFFFFFF))
                self._keep.append((clip, out, d.events))
                outs.append(out)
                continue
            if isinstance(d, TableDrive):
                n_out = int(d.n_out)
                segs = d.table
                out = outputs[i] if outputs is not None else _alloc_like(clip, self.channels, n_out)
                self._check(self.lib.bsb_add_kiosk_table(self.h, i, _ptr(clip), clip_len, _ptr(out), n_out, int(d.quantum),
                                                         d.table, len(d.table), int(d.seed) & 0xFFFFFFFF))
                self._keep.append((clip, out, segs))
                outs.append(out)
                continue
            segs = (_capi.Segment * len(d.segments))(*d.segments)
            if isinstance(d, KioskDrive):
                n_out = int(d.n_out)
                out = outputs[i] if outputs is not None else _alloc_like(clip, self.channels, n_out)
                self._check(self.lib.bsb_add_kiosk(self.h, i, _ptr(clip), clip_len, _ptr(out), n_out, int(d.quantum),
                                                   segs, len(d.segments), int(d.seed) & 0xFFFFFFFF))
            else:
                n_out = int(d.total_out)
                out = outputs[i] if outputs is not None else _alloc_like(clip, self.channels, n_out)
                self._check(self.lib.bsb_add_streaming(self.h, i, _ptr(clip), clip_len, _ptr(out), int(d.n_in),
                                                       int(d.n_out), int(d.n_calls), segs, len(d.segments),
                                                       int(d.seed) & 0xFFFFFFFF))
            self._keep.append((clip, out, segs))
            outs.append(out)
        self._check(self.lib.bsb_commit(self.h, int(chunk_blocks)))
        self.outputs = outs
        return outs

    def rebind(self, i, clip, out):
        self._check(self.lib.bsb_rebind(self.h, i, _ptr(clip), _ptr(out)))
        self._keep[i] = (clip, out, self._keep[i][2])
        self.outputs[i] = out

    def run(self, cuda_stream=None):
        """Execute every block of every stream.  ``cuda_stream``: a raw cudaStream_t (int); default = torch's
        current stream when torch tensors are in use, else the null stream."""
        if cuda_stream is None:
            cuda_stream = 0
            if self._keep and hasattr(self._keep[0][0], "data_ptr"):
                import torch
                cuda_stream = torch.cuda.current_stream().cuda_stream
        self._check(self.lib.bsb_run(self.h, C.c_void_p(cuda_stream)))
        return self.outputs

    def run_host(self, host_clips, host_outs, cuda_stream=None, sync=True):
        """Like run(), for audio in host memory: host_clips[i] / host_outs[i] are float32 [channels, n] arrays (pinned
        torch CPU tensors or numpy) matching the planned shapes.  The planned device tensors serve as staging; copies
        and kernels are pipelined chunk by chunk inside the library.  ``sync`` (default): return once the host outputs
        are complete; with ``sync=False`` the copies may still be in flight -- call ``synchronize()`` (or synchronise
        the stream) before reading ``host_outs``."""
        n = len(self._keep)
        assert len(host_clips) == n and (host_outs is None or len(host_outs) == n)
        if cuda_stream is None:
            cuda_stream = 0
            if self._keep and hasattr(self._keep[0][0], "data_ptr"):
                import torch
                cuda_stream = torch.cuda.current_stream().cuda_stream
        for i, ((clip, out, _), hc) in enumerate(zip(self._keep, host_clips)):
            assert tuple(hc.shape) == tuple(clip.shape) and (host_outs is None or tuple(host_outs[i].shape) == tuple(out.shape))
        ins = (C.c_void_p * n)(*[_ptr(x) for x in host_clips])
        outs = (C.c_void_p * n)(*[_ptr(x) for x in host_outs]) if host_outs is not None else None   # None: outputs stay on the device
        self._check(self.lib.bsb_run_host(self.h, ins, outs, C.c_void_p(cuda_stream)))
        if sync:
            self.synchronize()
        return host_outs

    def synchronize(self):
        """Block until the last run()/run_host() has finished on the device."""
        self._check(self.lib.bsb_synchronize(self.h))

    def total_blocks(self): return self.lib.bsb_total_blocks(self.h)
    def stream_blocks(self, i): return self.lib.bsb_stream_blocks(self.h, i)
    def chunk_blocks(self): return self.lib.bsb_chunk_blocks(self.h)
    def launch_count(self): return self.lib.bsb_launch_count(self.h)

    def gate_events(self):
        """Streaming drives: number of process() calls of the last run that the reference's silence gate would have
        short-circuited (2 * blockSamples of silent input, app-side ``_process(q, q)`` on silence).  The batched path
        does not model that branch; a non-zero count means the result differs from the reference there -- drive such
        material through ``StretchEngine`` instead.  Synchronises with the device."""
        return int(self.lib.bsb_gate_events(self.h))

    def set_profiling(self, on=True):
        """Bracket every kernel launch of the following runs with CUDA events (on the run's stream)."""
        self.lib.bsb_set_profiling(self.h, 1 if on else 0)

    def set_overlap(self, on=True):
        """Chunk pipelining on two internal CUDA streams (default on); off = strictly serial kernels."""
        self.lib.bsb_set_overlap(self.h, 1 if on else 0)

    def set_fast_fft(self, on=True):
        """STFT kernels specialised for the preset geometries (default on); off = the run-time-geometry kernels."""
        self.lib.bsb_set_fast_fft(self.h, 1 if on else 0)

    def fast_fft_active(self):
        return bool(self.lib.bsb_fast_fft_active(self.h))

    def kernel_stats(self):
        """{kernel name: dict(ms, launches, units)} of the last run; ``ms`` is 0 unless profiling was on."""
        out = {}
        for i in range(self.lib.bsb_kernel_count(self.h)):
            name, ms, n, u = C.c_char_p(), C.c_double(), C.c_longlong(), C.c_longlong()
            self.lib.bsb_kernel_stat(self.h, i, C.byref(name), C.byref(ms), C.byref(n), C.byref(u))
            out[name.value.decode()] = dict(ms=ms.value, launches=n.value, units=u.value)
        return out

    def kernel_launches(self, name):
        """[(ms, units)] of every launch of kernel ``name`` in the last run, in launch order (ms needs profiling on)."""
        for i in range(self.lib.bsb_kernel_count(self.h)):
            nm, ms, n, u = C.c_char_p(), C.c_double(), C.c_longlong(), C.c_longlong()
            self.lib.bsb_kernel_stat(self.h, i, C.byref(nm), C.byref(ms), C.byref(n), C.byref(u))
            if nm.value.decode() == name:
                cnt = int(n.value)
                a, b = (C.c_double * max(cnt, 1))(), (C.c_longlong * max(cnt, 1))()
                got = self.lib.bsb_kernel_launches(self.h, i, a, b, cnt)
                return [(a[j], b[j]) for j in range(min(got, cnt))]
        return []

    def block_info(self, stream, block):
        a = (C.c_longlong * 8)()
        if self.lib.bsb_block_info(self.h, stream, block, a) != 0:
            raise IndexError("no such block")
        tf = np.array([a[1]], np.uint32).view(np.float32)[0]
        return dict(flags=int(a[0]), timeFactor=float(tf), cur=(int(a[2]), int(a[3]), int(a[4])),
                    prev=(int(a[5]), int(a[6]), int(a[7])))
