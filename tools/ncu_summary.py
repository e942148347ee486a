#!/usr/bin/env python
"""Summarise ncu output for profiles/:  ncu_summary.py launches <csv>  |  ncu_summary.py full <ncu-rep>"""
import csv
import subprocess
import sys
from collections import defaultdict

FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "smsp__cycles_active.avg"]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = r[ik].split("(")[0].replace("void ", "").replace("f9::<unnamed>::", "").replace("<unnamed>::", "")
        tot[name] += v; cnt[name] += 1
    total = sum(tot.values())
    print(f"# {len(rows) - 1} launches, {total:.3f} ms of kernel time (ncu-serialised, cold cache: compare shares)")
    print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{k[:60]:60s} {cnt[k]:8d} {tot[k]:10.3f} {tot[k] / cnt[k]:9.4f} {100 * tot[k] / total:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for m in FULL:
            if m in hdr:
                print(f"  {m:70s} {r[hdr.index(m)]:>18s} {units[hdr.index(m)]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
