#!/usr/bin/env python
"""Turn ncu outputs into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
    python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep  > profiles/rNN_gemm_full.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    tot = 0.0
    for x in rows:
        name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "")
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"# ncu launch list: {len(rows)} launches, {tot / 1e3:.2f} ms total (gpu__time_duration.sum, --clock-control none;")
    print("# cold-cache, serialised: compare SHARES with the live CUDA-event numbers, not absolutes)\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:90]}` | {n} | {t / 1e3:.3f} | {100 * t / tot:.1f}% | {t / n:.1f} |")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "launch__cluster_dim_x", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem"]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of {path} ({len(rows) - 2} captured launches)\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## `{name[:100]}`  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
        print("| metric | value |\n|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
