#!/usr/bin/env python
"""Headline benchmark: aggregate audio-seconds processed per wall-second (48 kHz stereo), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA engine
    python bench.py --impl reference [...]                          # the reference's CPU engine on the host cores

Workload (config.workload): BASELINE.json configs[2] -- 256 independent 60 s 48 kHz stereo streams per GPU,
presetDefault, kiosk drive (seek + process(0,128) per render quantum, app/SignalsmithStretch.mjs:883-943), per-stream
constant rate (log-uniform 0.5..2) and transpose (integer -12..+12 st), tonality limit 8 kHz.  configs[1] is a single
stream and cannot express an aggregate; configs[0] is the CPU-runnable case.  One *step* = every block of every
stream of the batch, start to finish.  Audio-seconds are OUTPUT seconds (input seconds are reported beside them).
N > 1: weak scaling, each rank owns its own 256-stream shard, no data-path collective.

value  = device-resident inputs and outputs (HBM), CUDA-event time of K steps, max over ranks.
e2e    = the same K steps through the public API with HOST (pinned) buffers: H2D of every clip and D2H of every output
         inside the timed region.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 48000
METRIC = "aggregate audio-sec processed per wall-sec (48 kHz stereo)"
UNIT = "audio-s/s"


def workload(streams, seconds, rank):
    """Per-stream (rate, semitones) of SURVEY.md section 8d config 3, seeded; rank-disjoint for N > 1."""
    import numpy as np
    rng = np.random.default_rng(1 + rank)
    rates = np.exp(rng.uniform(math.log(0.5), math.log(2.0), streams))
    sts = rng.integers(-12, 13, streams)
    n_in = int(seconds * SR)
    n_out = [int(n_in / r) for r in rates]
    return rates, sts, n_in, n_out


def config_dict(args, preset="default"):
    return {"workload": "BASELINE configs[2]: %d x %g s 48 kHz stereo streams per GPU, preset%s, kiosk drive quantum 128, "
                        "rate log-uniform 0.5-2, transpose -12..+12 st, tonality 8 kHz" % (args.streams, args.seconds, preset.capitalize()),
            "streams_per_gpu": args.streams, "seconds_in_per_stream": args.seconds, "sample_rate": SR, "channels": 2,
            "preset": preset, "sharding": "streams, no collective",
            "l2": "not flushed: per-step inputs+outputs (%.1f GB) exceed the 126 MB L2" % (args.streams * args.seconds * SR * 2 * 4 * 2.08 / 1e9)}


# ------------------------------------------------------------------------------------------------ CPU reference leg
def _cpu_worker(job):
    """One stream through the reference's CPU engine (TEST INFRASTRUCTURE under oracle/, used here only as the
    timed baseline, never as the product path)."""
    idx, kind, seconds, rate, st = job
    import numpy as np
    from oracle import refdrive
    rng = np.random.default_rng(1000 + idx)
    n_in = int(seconds * SR)
    t = np.arange(n_in) / SR
    clip = np.stack([0.3 * np.sin(2 * math.pi * (110.0 + 7 * idx + 40 * c) * t) + 0.1 * rng.standard_normal(n_in) for c in range(2)]).astype(np.float32)
    eng = refdrive.RefEngine() if kind == "reference" else refdrive.PortEngine()
    n_out = int(n_in / rate)
    t0 = time.perf_counter()
    refdrive.kiosk_drive(eng, clip, SR, n_out, rate=rate, params=dict(semitones=float(st), tonality_hz=8000.0))
    dt = time.perf_counter() - t0
    eng.close()
    return n_out / SR, n_in / SR, dt


def cpu_reference(args, steps=1, warmup=0, sample_seconds=None):
    """Times the reference engine on the host cores: P processes, one stream each per step (the engine is
    single-threaded by construction).  Returns dict for the cpu_baseline object."""
    import multiprocessing as mp
    from oracle import refdrive
    kind = "reference" if os.path.exists(refdrive.REF_SO) else "port"
    cores = max(1, min(len(os.sched_getaffinity(0)), args.cpu_procs))
    sample_seconds = sample_seconds or args.cpu_sample_seconds
    rates, sts, _, _ = workload(args.streams, args.seconds, 0)
    jobs = [(i, kind, sample_seconds, float(rates[i % args.streams]), int(sts[i % args.streams])) for i in range(cores)]
    ctx = mp.get_context("spawn")
    out_s = in_s = 0.0
    wall = 0.0
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(i, kind, 0.5, 1.0, 0) for i in range(cores)])      # load libraries, page in
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            if it >= warmup:
                wall += dt
                out_s += sum(r[0] for r in res); in_s += sum(r[1] for r in res)
    per_core = [r[0] / r[2] for r in res]
    return dict(value=out_s / wall, unit=UNIT, cores=cores, kind=kind,
                sample="%d streams (one per core) x %g s input of the workload's first streams, kiosk drive, same rate/transpose draw; "
                       "engine = %s; per-core x real-time %.1f..%.1f; input-s/s %.1f" % (
                           cores, sample_seconds,
                           "the reference's wasm blob translated to C (oracle/wasm2c.py), gcc -O2" if kind == "reference" else "C port oracle/stretch_oracle.c, gcc -O2",
                           min(per_core), max(per_core), in_s / wall),
                ms_per_step=1e3 * wall / steps)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cb = cpu_reference(args, steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(args),
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v == "Active":
                    reasons.add(n)
        if not sm:                       # a timed region shorter than the sampling period: take what was seen around it
            for t, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0])); mx = float(f[1])
                except (ValueError, IndexError):
                    continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU leg
def stage_bytes(name, g):
    """Algorithmic HBM bytes per unit of each kernel (SURVEY.md section 8d stage model)."""
    L, H, B = g["L"], g["H"], g["B"]
    table = {
        "analysis_kernel": 4 * L + 8 * B,             # per (window, channel): L samples in, B complex bins out
        "map_energy_kernel": 8 * B + 4 * B + 4 * B,   # per channel-block: spectrum in, input energy out, band energy + smoother input (8B per block / 2 ch)
        "map_smooth_kernel": 4 * B * 8,               # per block: 4 sweeps over the smoothed array, read + write each
        "map_peaks_kernel": 8 * B + 8 * B,            # per block: energy + smoothed in, map out
        "map_fmapply_kernel": 4 * B + 8 * B,          # per channel-block: envelope in, input energy read + write
        "preterms_kernel": 16 * B + 8 * B + 48 * B,   # per channel-block: cur+prev spectra, energy+map in, 96-byte record row / 2 channels out
        "chain_kernel": 48 * B + 8 * B,               # per channel-block: record rows in, output spectrum out (phase state stays on chip)
        "isynth_kernel": 8 * B + 4 * L,               # per channel-block: output spectrum in, windowed frame out
        "ola_kernel": 4 * L + 4 * H,                  # per channel-block: frame in, H output samples out
    }
    return table.get(name)


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host buffers of the e2e leg
    are allocated next to it (first touch).  Best effort: returns the node or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def main_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import bauklank_audio_stretch_b200 as bs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU engine)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)                  # pinned host buffers on the GPU's own memory node (best effort)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"            # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    bs.load_library()                                    # in-tree CUDA build; raises if missing

    S = args.streams
    rates, sts, n_in, n_out = workload(S, args.seconds, rank)
    out_sec = sum(n_out) / SR
    in_sec = S * n_in / SR
    # synthetic clips, generated on the device: a few steady partials + noise per channel, distinct per stream
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    t = torch.arange(n_in, device=dev, dtype=torch.float32) / SR
    clips_all = torch.empty((S, 2, n_in), device=dev, dtype=torch.float32)
    for i in range(S):
        for c in range(2):
            f0 = 80.0 + 3.0 * i + 35.0 * c
            clips_all[i, c] = 0.25 * torch.sin(2 * math.pi * f0 * t) + 0.15 * torch.sin(2 * math.pi * 3.1 * f0 * t + 1.0)
        clips_all[i] += 0.08 * torch.randn((2, n_in), device=dev, generator=g)
    outs_all = torch.zeros((2 * sum(n_out),), device=dev, dtype=torch.float32)
    clips = [clips_all[i] for i in range(S)]
    outs, off = [], 0
    for n in n_out:
        outs.append(outs_all[off:off + 2 * n].view(2, n)); off += 2 * n
    drives = [bs.KioskDrive(n_out[i], [bs.segment(rate=float(rates[i]), semitones=float(sts[i]), tonality_hz=8000.0)]) for i in range(S)]
    eng = bs.BatchStretch(2, SR, preset=args.preset)
    if args.no_fast_fft:
        eng.set_fast_fft(False)
    t0 = time.perf_counter()
    eng.plan(clips, drives, outputs=outs)
    plan_s = time.perf_counter() - t0
    geom = dict(L=eng.blockSamples(), H=eng.intervalSamples(), B=eng.bands())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(args.warmup):
        eng.run()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    eng.set_profiling(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    a.record()
    for _ in range(args.steps):
        eng.run()
    b.record()
    barrier()
    w1 = time.perf_counter()
    ms = a.elapsed_time(b)
    launches = eng.launch_count() * args.steps
    stats = eng.kernel_stats()                            # of the last timed run (kernels of adjacent chunks overlap)
    clocks = sampler.stop(w0, w1) if sampler else None
    # one more, untimed, run with chunk pipelining off: every kernel alone on the GPU, for the per-kernel table
    eng.set_overlap(False)
    eng.run()
    torch.cuda.synchronize()
    stats_iso = eng.kernel_stats()
    eng.set_overlap(True)
    eng.set_profiling(False)
    chk = float(outs_all[::4097].abs().sum().item())      # the result is read (and must be finite)
    assert math.isfinite(chk) and chk > 0.0

    # ---- end to end: host buffers in, host buffers out, same steps
    h_in = torch.empty(clips_all.shape, dtype=torch.float32, pin_memory=True)
    h_in.copy_(clips_all)
    h_out = torch.empty(outs_all.shape, dtype=torch.float32, pin_memory=True)
    clips_all.zero_()
    e2e_steps = args.steps
    h_clips = [h_in[i] for i in range(S)]
    h_outs, off = [], 0
    for n in n_out:
        h_outs.append(h_out[off:off + 2 * n].view(2, n)); off += 2 * n

    def e2e_step():
        eng.run_host(h_clips, h_outs)      # H2D of every clip, all kernels, D2H of every output (pipelined per time chunk)
    e2e_step()
    barrier()
    a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a2.record()
    for _ in range(e2e_steps):
        e2e_step()
    b2.record()
    barrier()
    ms2 = a2.elapsed_time(b2)
    assert math.isfinite(float(h_out[::4097].abs().sum()))

    tmax = torch.tensor([ms, ms2], dtype=torch.float64, device=dev)
    tot = torch.tensor([out_sec, in_sec, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms2 = tmax.tolist()
    out_tot, in_tot, launches_tot = tot.tolist()

    if rank == 0:
        value = out_tot * args.steps / (ms / 1e3)
        e2e = out_tot * e2e_steps / (ms2 / 1e3)
        # dominant kernel by device time (last timed run), against the measured HBM peak
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        timed = {k: v for k, v in stats.items() if v["ms"] > 0}
        timed_iso = {k: v for k, v in stats_iso.items() if v["ms"] > 0}
        dom = max(timed_iso or timed, key=lambda k: (timed_iso or timed)[k]["ms"])   # dominant = most device time when run alone
        kernels = {}
        for k, v in stats.items():
            bpu = stage_bytes(k, geom)
            gbs = (bpu * v["units"] / (v["ms"] * 1e-3) / 1e9) if (bpu and v["ms"] > 0) else None
            kernels[k] = {"ms_per_step": round(v["ms"], 3), "launches": v["launches"], "units": v["units"], "bytes_per_unit": bpu,
                          "achieved_gbs": round(gbs, 1) if gbs else None, "frac": round(gbs / peak, 4) if gbs else None}
        d = kernels[dom]
        iso = {}
        for k, v in stats_iso.items():
            bpu = stage_bytes(k, geom)
            gbs = (bpu * v["units"] / (v["ms"] * 1e-3) / 1e9) if (bpu and v["ms"] > 0) else None
            iso[k] = {"ms_per_step": round(v["ms"], 3), "achieved_gbs": round(gbs, 1) if gbs else None, "frac": round(gbs / peak, 4) if gbs else None}
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the committed ncu --set full capture
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": d["frac"],
                    "achieved_isolated": iso.get(dom, {}).get("achieved_gbs"), "frac_isolated": iso.get(dom, {}).get("frac"),
                    "traffic": traffic, "peak_source": peak_src,
                    "bytes_per_launch": d["bytes_per_unit"] * d["units"] / max(1, d["launches"]) if d["bytes_per_unit"] else None,
                    "avg_launch_ms": d["ms_per_step"] / max(1, d["launches"]), "kernels": kernels, "kernels_isolated": iso,
                    "note": "achieved/kernels: per-kernel CUDA-event time inside the timed region (last timed step), each on the "
                            "stream it is launched on; the chain+synthesis of one time chunk run beside the analysis/map/terms of the "
                            "next, so these durations include sharing the GPU.  kernels_isolated: one extra untimed run with the "
                            "pipelining off (every kernel alone on the GPU)",
                    "nominal_hbm_gbs": 7700.0}
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference(args)
            cb.pop("ms_per_step", None)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config_dict(args, args.preset),
                "audio_seconds_out_per_step": out_tot, "audio_seconds_in_per_step": in_tot, "input_audio_s_per_s": in_tot * args.steps / (ms / 1e3),
                "x_realtime_per_gpu": value / world, "plan_seconds": plan_s,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h_in.numel() * 4) * world, "d2h_bytes_per_step": int(h_out.numel() * 4) * world,
                        "ms_per_step": ms2 / e2e_steps},
                "gpu_launches": int(launches_tot), "numa_node_rank0": numa, "roofline": roofline, "cpu_baseline": cb, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=256, help="streams per GPU (BASELINE configs[2]: 256)")
    ap.add_argument("--seconds", type=float, default=60.0, help="input seconds per stream (BASELINE configs[2]: 60)")
    ap.add_argument("--cpu-procs", type=int, default=32)
    ap.add_argument("--cpu-sample-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast-fft", action="store_true", help="A/B: the run-time-geometry STFT kernels instead of the specialised ones")
    ap.add_argument("--preset", default="default", choices=["default", "cheaper"], help="engine preset (the headline number uses default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
