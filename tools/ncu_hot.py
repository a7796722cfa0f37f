#!/usr/bin/env python
"""Summarise an ncu report: per kernel the headline metrics, and (with --src REGEX) the hottest SASS ranges of one kernel.

    python tools/ncu_hot.py gpurun_out/x.ncu-rep [--src kernel_regex] [--chunk 40]
"""
import csv
import io
import subprocess
import sys


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr = rows[0]
    want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(d.get("Kernel Name", "?")[:90])
        print("   " + "  ".join(f"{k.split('.')[0].replace('smsp__', '').replace('sm__', '')}={d[k]}" for k in want if k in d))
    if "--src" in sys.argv:
        kre = sys.argv[sys.argv.index("--src") + 1]
        chunk = int(sys.argv[sys.argv.index("--chunk") + 1]) if "--chunk" in sys.argv else 40
        rows = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre]))))
        h = [i for i, r in enumerate(rows) if len(r) > 5 and r[1] == "Source"][0]
        hdr = rows[h]
        ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
        data = []
        for r in rows[h + 1:]:
            try:
                data.append((int(r[ia]), r[isrc].strip(), int(r[isamp])))
            except (ValueError, IndexError):
                break
        tot = sum(d[0] for d in data) or 1
        tsm = sum(d[2] for d in data) or 1
        print(f"total warp instructions {tot}, SASS lines {len(data)}")
        for i in range(0, len(data), chunk):
            ch = data[i:i + chunk]
            c = sum(x[0] for x in ch)
            sm = sum(x[2] for x in ch)
            if c * 100 >= tot or sm * 100 >= tsm:
                ops = " ".join(x[1].split()[0] if not x[1].startswith("@") else x[1].split()[1] for x in ch[:12])
                print(f"  [{i:5d}] inst {100 * c / tot:5.1f}%  samples {100 * sm / tsm:5.1f}%  {ops}")


if __name__ == "__main__":
    main()
