"""CPU, world_size 2 over gloo: the N>1 path = disjoint stream ranges per rank, no data-path collective; only the
counters are combined.  Each rank runs its shard through the host-emulation build; the union must equal the
single-process result bit-for-bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bauklank_audio_stretch_b200 as bs
import cases
from conftest import HOSTEMU


def test_partition_properties():
    p = bs.shard.partition_streams
    assert p([10] * 8, 2) == [(0, 4), (4, 8)]
    assert p([5, 5], 4)[-1][1] == 2
    rng = np.random.default_rng(0)
    for ws in (1, 2, 3, 4, 8):
        costs = [int(c) for c in rng.integers(1, 4000, 257)]
        r = p(costs, ws)
        assert r[0][0] == 0 and r[-1][1] == len(costs) and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        loads = [sum(costs[a:b]) for a, b in r]
        assert max(loads) <= sum(costs) / ws + max(costs)
    assert bs.shard.estimated_blocks(1441, 1440) == 2 and bs.shard.estimated_blocks(0, 1440) == 0


NAMES = ["KA5", "rng_low_rate", "stream_transpose_only_q96", "stream_100_900"]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = bs.load_library(HOSTEMU)
    cs = [cases.CASES[n] for n in NAMES]
    costs = [bs.shard.estimated_blocks(c.get("n_out", 0) if c["drive"] == "kiosk" else 30000 // c["n_in"] * c["n_out"], 1440) for c in cs]
    lo, hi = bs.shard.my_range(costs, rank, world)
    outs = cases.run_cases_batch(bs, cs[lo:hi], lib=lib) if hi > lo else []
    dist.barrier()
    secs = sum(o.shape[1] for o in outs) / 48000.0
    sums, maxima = bs.shard.combine_counters(([secs, float(hi - lo)], [float(rank + 1)]))
    q.put((rank, lo, hi, [cases.sha_of(o) for o in outs], sums, maxima))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_union_equals_single_process(golden):
    meta, _ = golden
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60); assert p.exitcode == 0
    covered = []
    for rank, lo, hi, shas, sums, maxima in res:
        covered += list(range(lo, hi))
        for n, h in zip(NAMES[lo:hi], shas):
            assert h == meta[n]["sha256"], n
        assert sums[1] == len(NAMES) and maxima == [2.0]
        assert abs(sums[0] - sum(meta[n]["shape"][1] for n in NAMES) / 48000.0) < 1e-9
    assert covered == list(range(len(NAMES)))
