"""ctypes binding of ``libbauklank_stretch.so`` (declared in ``include/bauklank_stretch.h``).

The product loads the CUDA build that sits next to this file and fails loudly if it is missing or no GPU is present:
there is no CPU fallback.  (``tests/`` may bind the serial host-emulation build of the same sources by passing an
explicit path; nothing in the package does.)
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "libbauklank_stretch.so")


class Segment(C.Structure):
    """``bsb_segment``: one entry of the worklet's time map (app/SignalsmithStretch.mjs:587-600)."""
    _fields_ = [("output", C.c_double), ("input", C.c_double), ("rate", C.c_double),
                ("semitones", C.c_double), ("tonality_hz", C.c_double), ("formant_semitones", C.c_double),
                ("formant_base_hz", C.c_double), ("loop_start", C.c_double), ("loop_end", C.c_double),
                ("active", C.c_int32), ("formant_compensation", C.c_int32),
                ("transpose_factor", C.c_double), ("formant_factor", C.c_double)]


class Quantum(C.Structure):
    """``bsb_quantum``: one render quantum of a resolved control trace."""
    _fields_ = [("rate", C.c_double), ("input_samples_end", C.c_longlong), ("valid_start", C.c_longlong),
                ("valid_end", C.c_longlong), ("semitones", C.c_float), ("tonality_limit", C.c_float),
                ("formant_semitones", C.c_float), ("formant_base", C.c_float), ("formant_compensation", C.c_int32),
                ("active", C.c_int32), ("transpose_factor", C.c_float), ("formant_factor", C.c_float)]


class TraceEvent(C.Structure):
    """``bsb_trace_event``: one ``schedule()`` call of a control trace, applied before render quantum ``quantum``."""
    _fields_ = [("quantum", C.c_longlong), ("output_time", C.c_double), ("input", C.c_double), ("rate", C.c_double),
                ("semitones", C.c_double), ("loop_start", C.c_double), ("loop_end", C.c_double), ("tonality_hz", C.c_double),
                ("formant_semitones", C.c_double), ("formant_base_hz", C.c_double), ("active", C.c_int32),
                ("formant_compensation", C.c_int32), ("transpose_factor", C.c_double), ("formant_factor", C.c_double)]


_BATCH_SIG = {
    "bsb_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "bsb_create_preset": (C.c_void_p, [C.c_int, C.c_double, C.c_int]),
    "bsb_destroy": (None, [C.c_void_p]),
    "bsb_block_samples": (C.c_int, [C.c_void_p]),
    "bsb_interval_samples": (C.c_int, [C.c_void_p]),
    "bsb_input_latency": (C.c_int, [C.c_void_p]),
    "bsb_output_latency": (C.c_int, [C.c_void_p]),
    "bsb_fft_samples": (C.c_int, [C.c_void_p]),
    "bsb_bands": (C.c_int, [C.c_void_p]),
    "bsb_last_error": (C.c_char_p, [C.c_void_p]),
    "bsb_begin": (C.c_int, [C.c_void_p, C.c_int]),
    "bsb_add_kiosk": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int,
                                C.POINTER(Segment), C.c_int, C.c_uint32]),
    "bsb_add_kiosk_table": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int,
                                      C.POINTER(Quantum), C.c_longlong, C.c_uint32]),
    "bsb_add_kiosk_trace": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int,
                                      C.POINTER(TraceEvent), C.c_longlong, C.c_uint32]),
    "bsb_query_geometry": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "bsb_add_streaming": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int,
                                    C.c_longlong, C.POINTER(Segment), C.c_int, C.c_uint32]),
    "bsb_commit": (C.c_int, [C.c_void_p, C.c_int]),
    "bsb_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bsb_run_host": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    "bsb_rebind": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "bsb_synchronize": (C.c_int, [C.c_void_p]),
    "bsb_total_blocks": (C.c_longlong, [C.c_void_p]),
    "bsb_stream_blocks": (C.c_longlong, [C.c_void_p, C.c_int]),
    "bsb_chunk_blocks": (C.c_int, [C.c_void_p]),
    "bsb_launch_count": (C.c_longlong, [C.c_void_p]),
    "bsb_gate_events": (C.c_longlong, [C.c_void_p]),
    "bsb_block_info": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.POINTER(C.c_longlong)]),
    "bsb_selftest_arith": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "bsb_set_profiling": (None, [C.c_void_p, C.c_int]),
    "bsb_set_overlap": (None, [C.c_void_p, C.c_int]),
    "bsb_set_fast_fft": (None, [C.c_void_p, C.c_int]),
    "bsb_fast_fft_active": (C.c_int, [C.c_void_p]),
    "bsb_kernel_count": (C.c_int, [C.c_void_p]),
    "bsb_kernel_launches": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]),
    "bsb_kernel_stat": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double),
                                  C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
}

# Part 1 of the header: the reference's own 18 names (app/SignalsmithStretch.mjs:462-479)
_COMPAT_SIG = {
    "setBuffers": (C.c_void_p, [C.c_int, C.c_int]),
    "blockSamples": (C.c_int, []), "intervalSamples": (C.c_int, []),
    "inputLatency": (C.c_int, []), "outputLatency": (C.c_int, []),
    "reset": (None, []),
    "presetDefault": (None, [C.c_int, C.c_float]), "presetCheaper": (None, [C.c_int, C.c_float]),
    "configure": (None, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "setTransposeFactor": (None, [C.c_float, C.c_float]),
    "setTransposeSemitones": (None, [C.c_float, C.c_float]),
    "setFormantFactor": (None, [C.c_float, C.c_int]),
    "setFormantSemitones": (None, [C.c_float, C.c_int]),
    "setFormantBase": (None, [C.c_float]),
    "seek": (None, [C.c_int, C.c_double]),
    "process": (None, [C.c_int, C.c_int]),
    "flush": (None, [C.c_int]),
    "stretch_main": (C.c_int, [C.c_int, C.c_void_p]),
    "stretch_set_seed": (None, [C.c_uint32]),
}

EXPORTS = tuple(_BATCH_SIG) + tuple(_COMPAT_SIG)

_cache = {}


def load_library(path=None):
    """dlopen the engine and attach prototypes.  Raises if the shared object is missing (no fallback)."""
    path = os.path.abspath(path or os.environ.get("BAUKLANK_STRETCH_LIB") or DEFAULT_LIB)   # (the variable: A/B builds, scripts/probe_wide.py)
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise RuntimeError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % path)
    lib = C.CDLL(path)
    for table in (_BATCH_SIG, _COMPAT_SIG):
        for name, (res, args) in table.items():
            fn = getattr(lib, name, None)
            if fn is None:
                if table is _COMPAT_SIG and os.environ.get("BS_ALLOW_PARTIAL"):
                    continue
                raise RuntimeError("%s does not export %s" % (path, name))
            fn.restype = res
            fn.argtypes = args
    _cache[path] = lib
    return lib
