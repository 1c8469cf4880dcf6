"""CPU: the C-ABI shared library (the CUDA build) loads without a GPU and exports every symbol that
include/bauklank_stretch.h declares; compute entry points fail loudly instead of falling back."""
import ctypes
import os
import re
import subprocess

import pytest

import bauklank_audio_stretch_b200 as bs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bauklank_stretch.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^[A-Za-z_][\w \*]*?[ \*]([A-Za-z_]\w*)\s*\(", src, flags=re.M)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_header_lists_the_reference_names():
    names = declared_symbols()
    for n in ("setBuffers", "blockSamples", "intervalSamples", "inputLatency", "outputLatency", "reset", "presetDefault",
              "presetCheaper", "configure", "setTransposeFactor", "setTransposeSemitones", "setFormantFactor",
              "setFormantSemitones", "setFormantBase", "seek", "process", "flush"):   # app/SignalsmithStretch.mjs:462-479
        assert n in names
    assert len(names) >= 40


def test_library_exports_every_declared_symbol():
    lib_path = os.path.join(ROOT, "bauklank-audio-stretch_b200", "libbauklank_stretch.so")
    assert os.path.exists(lib_path), "CUDA build missing: run __graft_entry__.build()"
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if " T " in line)
    missing = [n for n in declared_symbols() if n not in exported]
    assert not missing, missing
    lib = bs.load_library()      # dlopen + prototypes for every name in the Python binding
    for n in bs.EXPORTS:
        assert hasattr(lib, n)
    assert set(bs.EXPORTS) >= set(declared_symbols())


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        bs.BatchStretch(2, 48000.0)          # bsb_create returns NULL: no CUDA device


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "bauklank-audio-stretch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", txt, flags=re.M), f
                assert not re.search(r"#include[^\n]*oracle", txt), f
                assert "libstretch_oracle" not in txt and "libstretch_ref" not in txt and "/root/reference" not in txt, f
