#!/usr/bin/env python
"""Turns gpurun_out/<tag>_launches.csv (ncu launch list) and <tag>_prof.ncu-rep (one --set full capture) into the
small text summaries committed under profiles/.   usage: scripts/summarise_ncu.py <tag> [<out-tag>]"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else tag
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

path = os.path.join(G, tag + "_launches.csv")
if os.path.exists(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += v_us; a[2] = min(a[2], v_us); a[3] = max(a[3], v_us)
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(os.path.join(P, out + "_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: gpurun_out/%s_launches.csv, %d launches\n" % (tag, sum(a[0] for a in agg.values())))
        f.write("%-28s %8s %12s %8s %10s %10s\n" % ("kernel", "launches", "total_us", "share", "min_us", "max_us"))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-28s %8d %12.1f %7.1f%% %10.1f %10.1f\n" % (k, a[0], a[1], 100 * a[1] / tot, a[2], a[3]))
    print(open(os.path.join(P, out + "_launches_summary.txt")).read())

rep = os.path.join(G, tag + "_prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
            "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
            "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
            "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    traffic = {}
    with open(os.path.join(P, out + "_top_kernel_ncu.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none, source: gpurun_out/%s_prof.ncu-rep (values per launch)\n" % tag)
        for r in data:
            f.write("\n")
            for w, i in idx:
                f.write("%-72s %s %s\n" % (w, r[i], units[i]))
            try:
                name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0].strip()
                def b(col):
                    v = float(r[hdr.index(col)].replace(",", "")); u = units[hdr.index(col)]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                traffic.setdefault(name, []).append(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))
            except Exception as e:      # noqa
                pass
    print(open(os.path.join(P, out + "_top_kernel_ncu.txt")).read()[:1500])
    tp = os.path.join(P, "traffic.json")
    cur = json.load(open(tp)) if os.path.exists(tp) else {}
    for k, v in traffic.items():
        cur[k] = max(v)          # the fullest captured launch (a whole chunk)
    cur["_note"] = "dram__bytes_read.sum + dram__bytes_write.sum of the FIRST launch of each kernel (first time chunk of bench.py's default workload: 256 streams x 256 blocks, every slot used) from the latest ncu --set full capture, see profiles/*_top_kernel_ncu.txt; bench.py pairs it with the event-timed duration and own-model bytes of that same launch"
    cur["_source"] = tag
    json.dump(cur, open(tp, "w"), indent=1, sort_keys=True)
