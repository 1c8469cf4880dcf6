#!/usr/bin/env python
"""Headline benchmark: aggregate audio-seconds processed per wall-second (48 kHz stereo), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C]     # this repo's CUDA engine
    python bench.py --impl reference [...]                                # the reference's CPU engine on the host cores

Workloads = BASELINE.json `configs` (SURVEY.md section 8d restates them as concrete synthetic inputs):

  --config 0  one 30 s 48 kHz stereo sweep+noise clip, presetDefault, rate 1, 0 st, kiosk drive
  --config 1  the same clip, parameters re-scheduled EVERY render quantum along the rate 0.5->2 / transpose -12..+12 st
              curve (tonality limit on), one stream on one GPU
  --config 2  256 independent 60 s streams per GPU, presetDefault, per-stream constant rate / transpose   [default at N = 1]
  --config 3  ONE job of 4096 streams -- the controller mix: the 5 (controller, channel) pairs of the reference's topology
              replicated, every stream driven by its own `set rate` / `set tone` message list (10-50 msgs/s) through the
              kiosk app's mapping and schedule(); half presetDefault, half presetCheaper -- partitioned over the N GPUs by
              block count (strong scaling, no collective on the data path)                         [default at N >= 2]
  --config 4  one 1 h 96 kHz 8-channel stream, configure(8, 960, 240, split), formant shift with auto base (replicas only)

BASELINE.json quotes 256 streams "on 1 B200" and 4096 streams "sharded across 2/4/8 B200": the defaults follow that.
One *step* = every block of every stream of the job, start to finish.  Audio-seconds are OUTPUT seconds.

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
# reference's controller -> channel topology (time_pitch_mapping.py:43-49): the 5 (controller, channel) pairs
TOPOLOGY = [("BKTP_CTL_01", "A"), ("BKTP_CTL_01", "B"), ("BKTP_CTL_02", "A"), ("BKTP_CTL_02", "B"), ("BKTP_CTL_03", "A")]


# ------------------------------------------------------------------------------------------------ workloads
def sweep_clip(seconds, sr, channels, seed=1234, noise=0.05):
    """BASELINE configs[0]: 0.5*sin log sweep 50 Hz -> 16 kHz (reversed on odd channels) + 0.05*N(0,1)."""
    import numpy as np
    n = int(round(seconds * sr))
    t = np.arange(n) / sr
    k = math.log(16000.0 / 50.0) / seconds
    up = 0.5 * np.sin(2 * math.pi * 50.0 * (np.exp(k * t) - 1.0) / k)
    rng = np.random.default_rng(seed)
    return np.stack([((up if c % 2 == 0 else up[::-1]) + noise * rng.standard_normal(n)).astype(np.float32) for c in range(channels)])


def tone_clip(idx, n, sr, channels, seed):
    """A few steady partials + noise per channel, distinct per stream (host twin of the device generator below)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    return np.stack([0.25 * np.sin(2 * math.pi * (80.0 + 3.0 * (idx % 200) + 35.0 * c) * t) +
                     0.15 * np.sin(2 * math.pi * 3.1 * (80.0 + 3.0 * (idx % 200) + 35.0 * c) * t + 1.0) +
                     0.08 * rng.standard_normal(n) for c in range(channels)]).astype(np.float32)


def controller_messages(i, seconds):
    """The `set` messages one hardware channel sends during `seconds` (wire format server-multi.py:47-48): piece-wise
    constant rate (log-uniform 0.5-2) and tone (integer -12..12), 10-50 messages per second, seeded per stream."""
    import numpy as np
    rng = np.random.default_rng(7000 + i)
    ctl, ch = TOPOLOGY[i % len(TOPOLOGY)]
    per_s = float(rng.uniform(10.0, 50.0))
    msgs = [(0.0, dict(type="set", controller=ctl, channel=ch, key="rate", value=float(np.exp(rng.uniform(math.log(0.5), math.log(2.0)))))),
            (0.0, dict(type="set", controller=ctl, channel=ch, key="tone", value=int(rng.integers(-12, 13))))]
    n = int(seconds * per_s * 1.3) + 8
    gaps = rng.exponential(1.0 / per_s, n)
    times = np.cumsum(gaps)
    is_rate = rng.random(n) < 0.5
    rates = np.exp(rng.uniform(math.log(0.5), math.log(2.0), n))
    tones = rng.integers(-12, 13, n)
    for t, r, rv, tv in zip(times.tolist(), is_rate.tolist(), rates.tolist(), tones.tolist()):
        if t >= seconds:
            break
        msgs.append((t, dict(type="set", controller=ctl, channel=ch, key="rate", value=rv) if r else
                     dict(type="set", controller=ctl, channel=ch, key="tone", value=tv)))
    return ch, msgs


def sweep_events(seconds_out, sr, quantum=128):
    """configs[1]: one schedule() call per render quantum along the curve rate 0.5 -> 2.0 (geometric over the output) and
    transpose -12 -> +12 st in integer steps (the hardware `tone` semantics, app/multi/app.mjs:568-575), tonality 8 kHz."""
    nq = int(math.ceil(seconds_out * sr / quantum))
    ev = []
    for k in range(nq):
        u = k / max(1, nq - 1)
        ev.append((k, "schedule", (dict(active=True, rate=0.5 * 4.0 ** u, semitones=float(round(-12 + 24 * u)), tonalityHz=8000.0,
                                        formantSemitones=0.0, formantCompensation=False, formantBaseHz=0.0, outputTime=k * quantum / sr),)))
    return ev


class Stream:
    """One stream of a job: how to make its clip, its drive, and the engine group it belongs to."""
    __slots__ = ("idx", "group", "n_in", "n_out", "clip", "drive_kind", "rate", "st", "events", "blocks")


GROUPS = {   # engine configurations a job may mix (one BatchStretch each per rank)
    "default48": dict(channels=2, sr=48000, kw=dict(preset="default"), H=1440),
    "cheaper48": dict(channels=2, sr=48000, kw=dict(preset="cheaper"), H=1920),
    "lowlat96": dict(channels=8, sr=96000, kw=dict(block_samples=960, interval_samples=240, split_computation=True), H=240),
}


def build_job(args):
    """The whole job (all ranks): list of Stream descriptions, cheap to make (no audio yet)."""
    import numpy as np
    cfg = args.config
    streams = []

    def add(idx, group, n_in, n_out, clip, kind, rate=1.0, st=0.0, events=None):
        s = Stream()
        s.idx, s.group, s.n_in, s.n_out, s.clip, s.drive_kind, s.rate, s.st, s.events = idx, group, n_in, n_out, clip, kind, rate, st, events
        s.blocks = (n_out + GROUPS[group]["H"] - 1) // GROUPS[group]["H"]
        streams.append(s)

    if cfg == 0:
        add(0, "default48", int(args.seconds * SR), int(args.seconds * SR), ("sweep", args.seconds), "static", 1.0, 0.0)
    elif cfg == 1:
        n_out = int(args.seconds * SR / 1.08)      # the curve's mean rate is 1.08: the whole clip is played about once
        add(0, "default48", int(args.seconds * SR), n_out, ("sweep", args.seconds), "trace")
    elif cfg == 2:
        n_in = int(args.seconds * SR)
        rng = np.random.default_rng(1)                       # SURVEY 8d config 3; the same draw on every rank (weak scaling)
        rates = np.exp(rng.uniform(math.log(0.5), math.log(2.0), args.streams))
        sts = rng.integers(-12, 13, args.streams)
        for r in range(args.world):
            for i in range(args.streams):
                add(r * args.streams + i, "default48", n_in, int(n_in / rates[i]), ("tones", i, 1234 + i), "static", float(rates[i]), float(sts[i]))
    elif cfg == 3:
        n_out = int(args.seconds * SR)                       # every stream plays `seconds` of output; the clip is 1.5 x that
        n_in = int(1.5 * args.seconds * SR)
        for i in range(args.streams):
            add(i, "default48" if i % 2 == 0 else "cheaper48", n_in, n_out, ("tones", i, 1234 + i), "controller")
    elif cfg == 4:
        n = int(args.seconds * 96000)
        add(0, "lowlat96", n, n, ("tones96", 0, 99), "static_formant", 1.0, 0.0)
    else:
        raise SystemExit("unknown --config")
    return streams


def shard_job(args, streams, rank, world):
    """This rank's streams.  configs[2]: its own replica of the 256 streams (weak scaling).  configs[3]: a contiguous range
    of the one job, balanced by block count.  Single-stream configs: replicas."""
    import bauklank_audio_stretch_b200 as bs
    if args.config == 2:
        return [s for s in streams if s.idx // args.streams == rank]
    if args.config == 3:
        lo, hi = bs.shard.my_range([s.blocks for s in streams], rank, world)
        return streams[lo:hi]
    return streams                                            # replicas only


def defaults_for(args):
    if args.config is None:
        args.config = 2 if args.world == 1 else 3
    if args.streams is None:
        args.streams = {2: 256, 3: 4096}.get(args.config, 1)
    if args.seconds is None:
        args.seconds = {0: 30.0, 1: 30.0, 2: 60.0, 3: 20.0, 4: 3600.0}[args.config]
    if args.steps is None:
        args.steps = 1 if args.config == 4 else 3
    if args.warmup is None:
        args.warmup = 1 if args.config == 4 else 3
    return args


def config_dict(args, nbytes=None):
    names = {0: "BASELINE configs[0]: one %g s 48 kHz stereo sweep+noise clip, presetDefault, rate 1, 0 st, kiosk drive quantum 128" % args.seconds,
             1: "BASELINE configs[1]: the same %g s clip, schedule() every render quantum along rate 0.5->2 (geometric) / transpose -12..+12 st (integer), "
                "tonality 8 kHz, presetDefault, one stream" % args.seconds,
             2: "BASELINE configs[2]: %d x %g s 48 kHz stereo streams per GPU, presetDefault, kiosk drive quantum 128, rate log-uniform 0.5-2, "
                "transpose -12..+12 st, tonality 8 kHz" % (args.streams, args.seconds),
             3: "BASELINE configs[3]: ONE job of %d streams x %g s output, controller mix (5 topology pairs replicated, per-stream set rate / set tone "
                "messages at 10-50 /s through the app mapping and schedule()), half presetDefault half presetCheaper, sharded by block count" % (args.streams, args.seconds),
             4: "BASELINE configs[4]: one %g s 96 kHz 8-channel stream, configure(8, 960, 240, split), formant +3 st with compensation, base auto" % args.seconds}
    d = {"workload": names[args.config], "config_index": args.config, "streams": args.streams, "seconds_per_stream": args.seconds,
         "sample_rate": 96000 if args.config == 4 else SR, "channels": 8 if args.config == 4 else 2,
         "sharding": {2: "weak: every rank its own replica of the batch, no collective", 3: "strong: contiguous stream ranges balanced by block count, no collective"}.get(args.config, "replicas only (a single stream cannot be split)")}
    if nbytes is not None:
        d["l2"] = "not flushed: per-step inputs+outputs (%.2f GB per GPU) %s the 126 MB L2" % (nbytes / 1e9, "exceed" if nbytes > 126e6 else "FIT IN")
    return d


# ------------------------------------------------------------------------------------------------ CPU reference leg
def host_clip(s):
    kind = s.clip[0]
    if kind == "sweep":
        return sweep_clip(s.clip[1], SR, 2)
    if kind == "tones":
        return tone_clip(s.clip[1], s.n_in, SR, 2, s.clip[2])
    return tone_clip(s.clip[1], s.n_in, 96000, 8, s.clip[2])


def stream_events(s):
    """The schedule() calls that drive stream `s` (trace / controller kinds), as WorkletTimeline events."""
    import bauklank_audio_stretch_b200 as bs
    if s.drive_kind == "trace":
        return sweep_events(s.n_out / SR, SR)
    ch, msgs = controller_messages(s.idx, s.n_out / SR)
    m = bs.ControllerMapper(audio_duration=s.n_in / SR, channel=ch)
    return m.messages_to_events(msgs, SR)


def oracle_stream(bs, s, n_out, clip, kind=None):
    """Stream `s` through the CPU engine (TEST INFRASTRUCTURE under oracle/: the checker and the timed baseline, never the
    product path), driven like the worklet drives the wasm: first n_out output samples."""
    from oracle import refdrive
    kind = kind or ("reference" if os.path.exists(refdrive.REF_SO) else "port")
    eng = refdrive.RefEngine() if kind == "reference" else refdrive.PortEngine()
    grp = GROUPS[s.group]
    if s.drive_kind in ("trace", "controller"):
        tl = bs.WorkletTimeline(float(grp["sr"]), channels=grp["channels"], config=dict(preset="cheaper") if s.group == "cheaper48" else None)
        y = tl.render(eng, n_out, events=stream_events(s), clip=clip)
    elif s.drive_kind == "static_formant":
        y, _ = refdrive.kiosk_drive(eng, clip, grp["sr"], n_out, rate=1.0, block=960, interval=240, split=1,
                                    params=dict(semitones=0.0, tonality_hz=8000.0, formant_semitones=3.0, formant_comp=True, formant_base_hz=0.0))
    else:
        y, _ = refdrive.kiosk_drive(eng, clip, SR, n_out, rate=s.rate, preset="cheaper" if s.group == "cheaper48" else "default",
                                    params=dict(semitones=float(s.st), tonality_hz=8000.0))
    eng.close()
    return y, kind


_JOBS = {}
_CLIPS = {}


def _cpu_worker(job):
    """A bounded sample of one stream through the reference's CPU engine; returns (output seconds, busy seconds)."""
    import copy
    import bauklank_audio_stretch_b200 as bs          # host-side mirror of the worklet only (geometry query); no device needed
    argv, idx, out_seconds, kind = job
    if tuple(argv) not in _JOBS:
        _JOBS[tuple(argv)] = build_job(_parse(argv))
    streams = _JOBS[tuple(argv)]
    s = copy.copy(streams[idx % len(streams)])
    sr = GROUPS[s.group]["sr"]
    n_out = int(min(out_seconds * sr, s.n_out))
    s.n_in = min(s.n_in, int(n_out * 2.0) + 4 * sr // 10)   # the sample only ever reads the start of the clip (rate <= 2)
    key = (tuple(argv), s.idx, s.n_in)
    if key not in _CLIPS:
        if len(_CLIPS) > 64:
            _CLIPS.clear()
        _CLIPS[key] = host_clip(s)
    clip = _CLIPS[key]
    t0 = time.perf_counter()
    oracle_stream(bs, s, n_out, clip, kind)
    return n_out / sr, time.perf_counter() - t0


def cpu_reference(args, steps=1, warmup=1, jobs_per_core=8, job_seconds=8.0):
    """Times the reference engine on the host cores: P worker processes fed `jobs_per_core` x P equal-sized jobs per step
    through imap_unordered (dynamic balancing: wall is not the slowest stream), each job a bounded sample -- the first
    `job_seconds` of output -- of one stream of the workload.  The engine is single-threaded by construction (SURVEY.md
    section 5), so the aggregate is P independent instances, like P kiosk processes."""
    import multiprocessing as mp
    from oracle import refdrive
    kind = "reference" if os.path.exists(refdrive.REF_SO) else "port"
    cores = max(1, min(len(os.sched_getaffinity(0)), args.cpu_procs))
    argv = _argv_of(args)
    n_streams = len(build_job(args))
    stride = max(1, n_streams // (cores * jobs_per_core))
    jobs = [(argv, (j * stride) % max(1, n_streams), job_seconds, kind) for j in range(cores * jobs_per_core)]
    ctx = mp.get_context("spawn")
    out_s = busy_s = wall = 0.0
    with ctx.Pool(cores) as pool:
        list(pool.imap_unordered(_cpu_worker, [(argv, j, 0.25, kind) for j in range(cores)]))      # load libraries, page in
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            res = list(pool.imap_unordered(_cpu_worker, jobs, chunksize=1))
            dt = time.perf_counter() - t0
            if it >= warmup:
                wall += dt
                out_s += sum(r[0] for r in res); busy_s += sum(r[1] for r in res)
    per_core = out_s / busy_s
    return dict(value=out_s / wall, unit=UNIT, cores=cores, kind=kind, per_core_x_realtime=per_core, balanced_aggregate=per_core * cores,
                sample="%d jobs per step (%d per core, imap_unordered) = the first %g s of output of every %d-th stream of the workload, same drive "
                       "as the GPU arm; engine = %s; per-core %.1f x real-time (engine busy time), %d cores x that = %.0f; value = output seconds / wall "
                       "(includes making the clips and the pool's hand-over)" % (
                           len(jobs), jobs_per_core, job_seconds, stride,
                           "the reference's wasm blob translated to C (oracle/wasm2c.py), gcc -O2" if kind == "reference" else "C port oracle/stretch_oracle.c, gcc -O2",
                           per_core, cores, per_core * cores),
                ms_per_step=1e3 * wall / steps)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cb = cpu_reference(args, steps=args.steps, warmup=max(1, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb.pop("ms_per_step"), "higher_is_better": True,
            "scaling": "strong" if args.config == 3 else "weak", "vs_baseline": None,
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
def own_bytes(name, g):
    """HBM bytes per unit as THIS implementation moves them.  The term records written by
    preterms_kernel and read back by chain_kernel are an artefact of the design, not algorithmic traffic: see stage_model."""
    L, H, B, C = g["L"], g["H"], g["B"], g.get("C", 2)
    table = {
        "analysis_kernel": 4 * L + 8 * B + 2 * B,     # per (window, channel): L samples in, B complex bins out; the current window (one of
                                                      # the two) also stores its B input energies
        "map_energy_kernel": 4 * B + 8 * B // C,      # per channel-block: input energy in; per block: band energy + smoother input out
        "map_smooth_kernel": 4 * B * 8,               # per block: 4 sweeps over the smoothed array, read + write each
        "map_peaks_kernel": 8 * B + 8 * B,            # per block: energy + smoothed in, map out
        "map_fmapply_kernel": 4 * B + 8 * B,          # per channel-block: envelope in, input energy read + write
        "preterms_kernel": 16 * B + 8 * B + 48 * B,   # per channel-block: cur+prev spectra, energy+map in, 96-byte record row / 2 channels out
        "chain_kernel": 48 * B + 8 * B,               # per channel-block: record rows in, output spectrum out
        "isynth_kernel": 8 * B + 4 * L,               # per channel-block: output spectrum in, windowed frame out
        "ola_kernel": 4 * L + 4 * H,                  # per channel-block: frame in, H output samples out
    }
    return table.get(name)


def stage_model(g, A=2):
    """SURVEY.md section 8d stage-model bytes per unit (one block of one channel): each array crosses HBM once per stage."""
    L, H, B = g["L"], g["H"], g["B"]
    return {"analysis": A * (4 * L + 8 * B), "spectral": 8 * B * A + 24 * B, "synthesis_ola": 8 * B + 8 * L + 4 * H,
            "total": A * (4 * L + 16 * B) + 32 * B + 8 * L + 4 * H}


STAGE_OF = {"analysis_kernel": "analysis", "isynth_kernel": "synthesis_ola", "ola_kernel": "synthesis_ola"}   # everything else: spectral


def numa_of_gpu(index):
    """NUMA node of the GPU's PCI function (-1: the platform exposes a single memory node) and, if there is a real node and
    the cpuset allows, binds this process's CPUs and memory policy to it so that pinned buffers are local."""
    info = {"node": None, "bound": False, "nodes_online": None}
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        info["bdf"] = bdf
        info["nodes_online"] = open("/sys/devices/system/node/online").read().strip()
        info["node"] = int(open("/sys/bus/pci/devices/" + bdf + "/numa_node").read())
        if info["node"] >= 0:
            cpus = set()
            for part in open("/sys/bus/pci/devices/" + bdf + "/local_cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound"] = True
            try:                                   # memory policy: prefer the GPU's node even if the CPUs could not move
                import ctypes
                libc = ctypes.CDLL(None, use_errno=True)
                mask = ctypes.c_ulong(1 << info["node"])
                if libc.syscall(238, 1, ctypes.byref(mask), 64) == 0:      # set_mempolicy(MPOL_PREFERRED, ...)
                    info["bound"] = True
            except Exception:
                pass
    except Exception as e:
        info["error"] = repr(e)[:120]
    return info


def device_clip(s, dev, gen):
    """The stream's clip on the device (same formula as tone_clip, device RNG for the noise -- the parity check below copies
    the device clips of its sample streams to the host, so both engines see the same bits)."""
    import torch
    kind = s.clip[0]
    if kind == "sweep":
        return torch.from_numpy(sweep_clip(s.clip[1], SR, 2)).to(dev)
    sr, ch = (96000, 8) if kind == "tones96" else (SR, 2)
    t = torch.arange(s.n_in, device=dev, dtype=torch.float32) / sr
    out = torch.empty((ch, s.n_in), device=dev, dtype=torch.float32)
    for c in range(ch):
        f0 = 80.0 + 3.0 * (s.clip[1] % 200) + 35.0 * c
        out[c] = 0.25 * torch.sin(2 * math.pi * f0 * t) + 0.15 * torch.sin(2 * math.pi * 3.1 * f0 * t + 1.0)
    out += 0.08 * torch.randn((ch, s.n_in), device=dev, generator=gen)
    return out


def make_drive(bs, s):
    if s.drive_kind == "static":
        return bs.KioskDrive(s.n_out, [bs.segment(rate=s.rate, semitones=s.st, tonality_hz=8000.0)])
    if s.drive_kind == "static_formant":
        return bs.KioskDrive(s.n_out, [bs.segment(rate=1.0, semitones=0.0, tonality_hz=8000.0, formant_semitones=3.0, formant_compensation=True,
                                                  formant_base_hz=0.0)])
    return bs.TraceDrive(s.n_out, bs.trace_events(stream_events(s)))


def compare(got, ref):
    import numpy as np
    same = bool((got.view(np.uint32) == ref.view(np.uint32)).all())
    d = got.astype(np.float64) - ref.astype(np.float64)
    err = float(np.abs(d).max()) if d.size else 0.0
    den = float((d ** 2).sum())
    snr = None if den == 0 else 10 * math.log10(max(float((ref.astype(np.float64) ** 2).sum()), 1e-300) / den)
    return same, err, snr


def main_gpu(args):
    import torch
    import torch.distributed as dist
    import bauklank_audio_stretch_b200 as bs

    world, rank = args.world, int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU engine)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = numa_of_gpu(local)                            # pinned host buffers next to the GPU where the platform has nodes
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"            # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    bs.load_library()                                    # in-tree CUDA build; raises if missing

    t0 = time.perf_counter()
    job = build_job(args)
    mine = shard_job(args, job, rank, world)
    gen = torch.Generator(device=dev).manual_seed(1234 + (rank if args.config != 2 else 0))
    groups = {}
    for s in mine:
        groups.setdefault(s.group, []).append(s)
    engines = []                                         # (group name, engine, streams, clips, outs, flat in tensor, flat out tensor)
    for name, ss in groups.items():
        grp = GROUPS[name]
        ch = grp["channels"]
        clips_all = torch.empty((sum(s.n_in for s in ss) * ch,), device=dev, dtype=torch.float32)
        outs_all = torch.zeros((sum(s.n_out for s in ss) * ch,), device=dev, dtype=torch.float32)
        clips, outs, oi, oo = [], [], 0, 0
        for s in ss:
            c = clips_all[oi:oi + ch * s.n_in].view(ch, s.n_in); oi += ch * s.n_in
            c.copy_(device_clip(s, dev, gen))
            clips.append(c)
            outs.append(outs_all[oo:oo + ch * s.n_out].view(ch, s.n_out)); oo += ch * s.n_out
        eng = bs.BatchStretch(ch, grp["sr"], **grp["kw"])
        if args.no_fast_fft:
            eng.set_fast_fft(False)
        eng.plan(clips, [make_drive(bs, s) for s in ss], outputs=outs)
        engines.append((name, eng, ss, clips, outs, clips_all, outs_all))
    plan_s = time.perf_counter() - t0
    out_sec = sum(s.n_out / GROUPS[s.group]["sr"] for s in mine)
    in_sec = sum(s.n_in / GROUPS[s.group]["sr"] for s in mine)
    my_blocks = sum(s.blocks for s in mine)
    io_bytes = sum(int(e[5].numel() + e[6].numel()) * 4 for e in engines)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # end to end, one CUDA stream per engine (a batch of mixed presets = one engine per preset): the second engine's uploads run beside
    # the first one's kernels (+6 % on the 4096-stream job; --serial-engines: one after the other)
    side = [torch.cuda.Stream() for _ in engines] if len(engines) > 1 and not args.serial_engines else None

    def on_engines(fn):
        if side is None:
            for i, e in enumerate(engines):
                fn(i, e, None)
            return
        cur = torch.cuda.current_stream()
        for i, (e, st) in enumerate(zip(engines, side)):
            st.wait_stream(cur)
            fn(i, e, st.cuda_stream)
        for st in side:
            cur.wait_stream(st)

    def step():   # (device-resident: one engine after the other measured 1 % faster than interleaved; end to end it is the other way round)
        for e in engines:
            e[1].run()

    # ---- device-resident timing
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    for e in engines:
        e[1].set_profiling(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    barrier()
    w1 = time.perf_counter()
    ms = a.elapsed_time(b)
    launches = sum(e[1].launch_count() for e in engines) * args.steps

    def gather_stats():
        tot = {}
        for e in engines:
            for k, v in e[1].kernel_stats().items():
                t = tot.setdefault(k, dict(ms=0.0, launches=0, units=0))
                t["ms"] += v["ms"]; t["launches"] += v["launches"]; t["units"] += v["units"]
        return tot
    stats = gather_stats()                                # of the last timed run (kernels of adjacent chunks overlap)
    clocks = sampler.stop(w0, w1) if sampler else None
    # one more, untimed, run with chunk pipelining off: every kernel alone on the GPU, for the per-kernel table
    for e in engines:
        e[1].set_overlap(False)
    for e in engines:                                     # (one engine after the other: nothing else on the GPU)
        e[1].run()
    torch.cuda.synchronize()
    stats_iso = gather_stats()
    first = {}                                            # the first launch of every kernel of the biggest group (first time chunk: the fullest launch)
    big = max(engines, key=lambda e: len(e[2]))
    for k in stats_iso:
        ls = big[1].kernel_launches(k)
        if ls:
            first[k] = ls[0]
    for e in engines:
        e[1].set_overlap(True)
        e[1].set_profiling(False)
    for e in engines:
        chk = float(e[6][::4097].abs().sum().item())      # the result is read (and must be finite)
        assert math.isfinite(chk) and (chk > 0.0 or e[6].numel() < 4097)

    # ---- parity of the timed batch: sample streams against the CPU engine on the same clip bits (outside the timed region)
    parity = None
    if rank == 0 and not args.no_parity_check:
        cands = sorted(((s.n_out, gi, si) for gi, e in enumerate(engines) for si, s in enumerate(e[2])), key=lambda t: t[0])
        picks = [cands[0], cands[len(cands) // 2], cands[-1]] if len(cands) >= 3 else cands
        res, kind = [], None
        for _, gi, si in dict.fromkeys(picks):
            e = engines[gi]; s = e[2][si]
            n_check = min(s.n_out, int(args.parity_seconds * GROUPS[s.group]["sr"]))
            ref, kind = oracle_stream(bs, s, n_check, e[3][si].cpu().numpy())
            same, err, snr = compare(e[4][si][:, :n_check].cpu().numpy(), ref)
            res.append(dict(stream=s.idx, group=s.group, samples=n_check, bit_identical=same, max_abs_err=err, snr_db=snr))
        parity = dict(checker=kind, tolerance="max|err| <= 1e-4 and SNR >= 90 dB per stream (BASELINE north_star)", streams=res,
                      ok=all(r["max_abs_err"] <= 1e-4 and (r["snr_db"] is None or r["snr_db"] >= 90.0) for r in res))
        assert parity["ok"], parity

    # ---- end to end: host buffers in, host buffers out, same steps
    hosts = []
    for e in engines:
        h_in = torch.empty(e[5].shape, dtype=torch.float32, pin_memory=True)
        h_in.copy_(e[5])
        h_out = torch.empty(e[6].shape, dtype=torch.float32, pin_memory=True)
        e[5].zero_()
        ch = GROUPS[e[0]]["channels"]
        hc, ho, oi, oo = [], [], 0, 0
        for s in e[2]:
            hc.append(h_in[oi:oi + ch * s.n_in].view(ch, s.n_in)); oi += ch * s.n_in
            ho.append(h_out[oo:oo + ch * s.n_out].view(ch, s.n_out)); oo += ch * s.n_out
        hosts.append((hc, ho, h_in, h_out))
    e2e_steps = args.steps

    def e2e_step(outs=True):   # H2D of every clip, all kernels, D2H of every output (pipelined per time chunk)
        on_engines(lambda i, e, q: e[1].run_host(hosts[i][0], hosts[i][1] if outs else None, cuda_stream=q, sync=False))
    e2e_step()
    barrier()
    a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a2.record()
    for _ in range(e2e_steps):
        e2e_step()
    b2.record()
    barrier()
    ms2 = a2.elapsed_time(b2)
    for h in hosts:
        assert math.isfinite(float(h[3][::4097].abs().sum()))
    # ... and with the outputs left on the device (inputs still cross the bus every step): says how much of the e2e cost is the
    # device-to-host direction of the host's memory system
    n3 = max(1, min(e2e_steps, 3))
    e2e_step(False)
    barrier()
    a3, b3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a3.record()
    for _ in range(n3):
        e2e_step(False)
    b3.record()
    barrier()
    ms3 = a3.elapsed_time(b3) * e2e_steps / n3

    tmax = torch.tensor([ms, ms2, ms3, float(my_blocks), -float(my_blocks)], dtype=torch.float64, device=dev)
    tot = torch.tensor([out_sec, in_sec, float(launches), float(sum(int(h[2].numel()) for h in hosts) * 4), float(sum(int(h[3].numel()) for h in hosts) * 4)],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms2, ms3, blocks_max, neg_blocks_min = tmax.tolist()
    out_tot, in_tot, launches_tot, h2d_tot, d2h_tot = tot.tolist()

    if rank == 0:
        value = out_tot * args.steps / (ms / 1e3)
        e2e = out_tot * e2e_steps / (ms2 / 1e3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        eng0 = big[1]
        geom = dict(L=eng0.blockSamples(), H=eng0.intervalSamples(), B=eng0.bands(), C=GROUPS[engines[0][0]]["channels"])
        single_geom = len(engines) == 1
        sm = stage_model(geom)

        def table(st):
            out = {}
            for k, v in st.items():
                bpu = own_bytes(k, geom) if single_geom else None
                gbs = (bpu * v["units"] / (v["ms"] * 1e-3) / 1e9) if (bpu and v["ms"] > 0) else None
                out[k] = {"ms_per_step": round(v["ms"], 3), "launches": v["launches"], "units": v["units"], "bytes_per_unit": bpu,
                          "achieved_gbs": round(gbs, 1) if gbs else None, "frac": round(gbs / peak, 4) if gbs else None}
            return out
        kernels, iso = table(stats), table(stats_iso)
        timed_iso = {k: v for k, v in stats_iso.items() if v["ms"] > 0}
        dom = max(timed_iso, key=lambda k: timed_iso[k]["ms"]) if timed_iso else "chain_kernel"   # dominant = most device time when run alone
        # roofline of the dominant kernel: its FIRST launch (the first time chunk: every stream live, every slot used -- the
        # launch the committed ncu capture describes), timed alone on the GPU in this run
        f_ms, f_units = first.get(dom, (0.0, 0))
        bpu = own_bytes(dom, geom)
        f_bytes = bpu * f_units if bpu else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes of that same launch shape from the committed ncu --set full capture
        if os.path.exists(tp) and args.config == 2 and args.streams == 256 and args.seconds >= 60:
            traffic = json.load(open(tp)).get(dom)
        ach = (f_bytes / (f_ms * 1e-3) / 1e9) if (f_bytes and f_ms > 0) else None
        # stage model: bytes every stage must move if each array crossed HBM once, over the device time the
        # stage's kernels took alone -- the yardstick that does not credit the term-record round trip
        stage_ms = {"analysis": 0.0, "spectral": 0.0, "synthesis_ola": 0.0}
        for k, v in stats_iso.items():
            stage_ms[STAGE_OF.get(k, "spectral")] += v["ms"]
        units = stats_iso.get("chain_kernel", {}).get("units", 0)       # channel-blocks of a step
        stages = {}
        for st_name, t_ms in stage_ms.items():
            if single_geom and t_ms > 0 and units:
                gbs = sm[st_name] * units / (t_ms * 1e-3) / 1e9
                stages[st_name] = {"bytes_per_unit": sm[st_name], "ms_isolated": round(t_ms, 3), "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        whole = (sm["total"] * units * args.steps / (ms * 1e-3) / 1e9) if (single_geom and units) else None
        roofline = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1) if ach else None, "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4) if ach else None, "traffic": traffic, "peak_source": peak_src,
                    "launch": "first launch of the kernel in the isolated run (first time chunk: the fullest launch; the one profiles/*_ncu.txt captures)",
                    "bytes_per_launch": f_bytes, "launch_ms": round(f_ms, 4), "launch_units": f_units,
                    "bytes_model": "own bytes (what this implementation moves: includes the 96-byte term records); see stage_model for SURVEY 8d's",
                    "stage_model": {"bytes_per_unit": sm, "stages_isolated": stages,
                                    "whole_step_gbs": round(whole, 1) if whole else None, "whole_step_frac": round(whole / peak, 4) if whole else None,
                                    "note": "SURVEY.md section 8d: analysis A(4L+8B), spectral 8BA+24B, synthesis+OLA 8B+8L+4H per channel-block, A=2"},
                    "kernels": kernels, "kernels_isolated": iso,
                    "note": "kernels: per-kernel CUDA-event time inside the timed region (last timed step), each on the stream it is launched "
                            "on; the chain+synthesis of one time chunk run beside the analysis/map/terms of the next.  kernels_isolated: one "
                            "extra untimed run with the pipelining off (every kernel alone on the GPU)",
                    "nominal_hbm_gbs": 7700.0}
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference(args)
            cb.pop("ms_per_step", None)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.config == 3 else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, io_bytes),
                "audio_seconds_out_per_step": out_tot, "audio_seconds_in_per_step": in_tot,
                "x_realtime_per_gpu": value / world, "plan_seconds": plan_s, "fast_fft": bool(eng0.fast_fft_active()),
                "blocks_per_rank": {"min": -neg_blocks_min, "max": blocks_max},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d_tot), "d2h_bytes_per_step": int(d2h_tot), "ms_per_step": ms2 / e2e_steps,
                        "engines": ("%d engines (one per preset), each on its own CUDA stream" % len(engines)) if side is not None else
                                   ("%d engine%s, one after the other" % (len(engines), "" if len(engines) == 1 else "s")),
                        "outputs_left_on_device": {"value": out_tot * e2e_steps / (ms3 / 1e3), "ms_per_step": ms3 / e2e_steps,
                                                   "note": "same steps, inputs from pinned host memory every step, outputs not copied back"}},
                "gpu_launches": int(launches_tot), "numa": numa, "parity_check": parity, "roofline": roofline, "cpu_baseline": cb, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=None, choices=[0, 1, 2, 3, 4], help="BASELINE.json configs index (default: 2 on one GPU, 3 on several)")
    ap.add_argument("--streams", type=int, default=None, help="streams per GPU (config 2: 256) / of the whole job (config 3: 4096)")
    ap.add_argument("--seconds", type=float, default=None, help="seconds per stream (config 2: 60 s input; config 3: 20 s output; config 4: 3600)")
    ap.add_argument("--cpu-procs", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--parity-seconds", type=float, default=30.0, help="output seconds per sample stream checked against the CPU engine")
    ap.add_argument("--serial-engines", action="store_true", help="A/B: the engines of a mixed-preset batch one after the other on one stream")
    ap.add_argument("--no-fast-fft", action="store_true", help="A/B: the run-time-geometry STFT kernels instead of the specialised ones")
    return ap


def _parse(argv):
    args = _parser().parse_args(argv)
    args.world = int(os.environ.get("WORLD_SIZE", "1")) if args.impl == "b200" else max(1, args.gpus)
    return defaults_for(args)


def _argv_of(args):
    return ["--impl", "reference", "--gpus", str(args.world), "--config", str(args.config), "--streams", str(args.streams), "--seconds", str(args.seconds)]


def main():
    args = _parse(sys.argv[1:])
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
