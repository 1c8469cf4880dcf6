"""Generates tests/golden/*.json|*.npy by running the REFERENCE's own engine (its WASM blob translated mechanically
to C under oracle/_ref, built from /root/reference by `make -C oracle ref`) on seeded synthetic inputs.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The fixtures are what travels to the GPU box; this script never runs there.

Each case stores: shape, sha256 of the raw little-endian f32 output bytes, sum|y|, two probe samples and a strided
subsample (every 97th sample per channel) so that a mismatch can be localised without the full signal.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refdrive  # noqa: E402
import cases  # noqa: E402  (tests/cases.py: the shared case table)


def describe(y):
    y = np.ascontiguousarray(y, np.float32)
    return dict(shape=list(y.shape), sha256=hashlib.sha256(y.tobytes()).hexdigest(),
                sum_abs=float(np.abs(y.astype(np.float64)).sum()),
                probe=[float(y[0, min(20000, y.shape[1] - 1)]), float(y[-1, min(40000, y.shape[1] - 1)])])


def main():
    assert os.path.exists(refdrive.REF_SO), "build oracle/_ref first: make -C oracle ref"
    meta = {}
    sub = {}
    for name, case in list(cases.CASES.items()) + list(cases.SHIM_CASES.items()):
        eng = refdrive.RefEngine(seed=case.get("seed", 1))
        y = cases.run_case(eng, case)
        eng.close()
        meta[name] = describe(y)
        sub[name] = np.ascontiguousarray(y[:, ::97])
        print(name, meta[name]["shape"], meta[name]["sha256"][:16], meta[name]["probe"], meta[name]["sum_abs"], flush=True)
    x = refdrive.survey_clip()
    meta["_survey_clip"] = dict(sha256=hashlib.sha256(x.tobytes()).hexdigest(), first=[float(v) for v in x[0, :4]])
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "subsamples.npz"), **sub)


if __name__ == "__main__":
    main()
