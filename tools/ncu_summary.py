#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): one line per profiled launch with the metrics the
roofline discussion uses.  usage: tools/ncu_summary.py <rep> [extra metric substrings...]"""
import csv, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum"] + extra
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    parts = []
    for w in want:
        for i, h in enumerate(hdr):
            if h == w or (w in extra and w in h):
                parts.append(f"{h.split('.')[0].split('__')[-1] if w not in extra else h}={r[i]}{rows[1][i]}")
    print(name[:90], "|", " ".join(parts))
