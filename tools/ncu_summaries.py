#!/usr/bin/env python
"""Text summaries of ncu output for profiles/:

    python tools/ncu_summaries.py launches gpurun_out/x_launches.csv "header line" > profiles/rNN_ncu_launches_..._summary.txt
    python tools/ncu_summaries.py full gpurun_out/x.ncu-rep "header line" > profiles/rNN_ncu_full_....txt
    python tools/ncu_summaries.py traffic gpurun_out/x.ncu-rep > profiles/ncu_traffic.json   (dram bytes per launch, by stage)
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("pcs::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    return name.strip()


def launches(path, header):
    rows = list(csv.reader(open(path)))
    hdr = [r for r in rows if len(r) > 5 and r[0] == "ID"][0]
    agg = collections.OrderedDict()
    n = 0
    for r in rows:
        if len(r) != len(hdr) or r[0] == "ID":
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = d.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)         # -> us
        key = (short(d["Kernel Name"]), d.get("Grid Size", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print(header)
    print(f"{n} launches, {tot / 1e3:.3f} ms of kernel time (serialised, cold cache: compare SHARES)\n")
    print(f"{'kernel':<78} {'grid':>16} {'launches':>8} {'total ms':>10} {'avg us':>10} {'share':>7}")
    for (k, g), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:78]:<78} {g:>16} {c:>8} {t / 1e3:>10.3f} {t / c:>10.1f} {100 * t / tot:>6.1f}%")


FULL = [("time_duration", "gpu__time_duration.sum"), ("pipe_tensor_cycles_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("sm_throughput", "sm__throughput.avg.pct_of_peak_sustained_elapsed"), ("dram_bytes_read", "dram__bytes_read.sum"),
        ("dram_bytes_write", "dram__bytes_write.sum"), ("dram_throughput", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2_throughput", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("registers", "launch__registers_per_thread"),
        ("warps_active", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("issue_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("tc_smem_wavefronts", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("fp64_pipe", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), ("inst_executed", "smsp__inst_executed.sum")]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def full(rep, header):
    hdr, units, rows = raw(rep)
    u = dict(zip(hdr, units))
    print(header)
    print("One line per profiled launch: " + ", ".join(k for k, _ in FULL) + " (ncu units).")
    for r in rows:
        d = dict(zip(hdr, r))
        parts = []
        for k, m in FULL:
            if m not in d:
                continue
            parts.append(f"{k}={d[m]}{u.get(m, '')}")
        print(short(d["Kernel Name"])[:90] + " | " + " ".join(parts))


def traffic(rep):
    hdr, units, rows = raw(rep)
    u = dict(zip(hdr, units))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = collections.OrderedDict()
    for r in rows:
        d = dict(zip(hdr, r))
        b = sum(float(d[m].replace(",", "")) * scale.get(u[m], 1) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out.setdefault(short(d["Kernel Name"]), []).append(int(b))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": lambda: launches(sys.argv[2], sys.argv[3]), "full": lambda: full(sys.argv[2], sys.argv[3]),
     "traffic": lambda: traffic(sys.argv[2])}[sys.argv[1]]()
