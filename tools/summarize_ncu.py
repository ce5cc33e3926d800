#!/usr/bin/env python
"""Turns the raw ncu artefacts a gpurun call brought back (gpurun_out/) into the committed summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
    python tools/summarize_ncu.py kernel gpurun_out/prof_conv_r1.ncu-rep profiles/r1_conv_v1_full.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "smsp__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "smsp__inst_executed.sum",
    "sm__memory_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
]


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0][:70]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
        tot[k] += v
        cnt[k] += 1
    total = sum(tot.values())
    with open(dst, "w") as out:
        out.write(f"# ncu launch list ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares)\n\n")
        out.write("| kernel | launches | avg us | share of captured time |\n|---|---:|---:|---:|\n")
        for k in sorted(tot, key=tot.get, reverse=True):
            out.write(f"| `{k}` | {cnt[k]} | {tot[k] / cnt[k]:.1f} | {100 * tot[k] / total:.1f}% |\n")


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as out:
        out.write(f"# ncu --set full summary ({src})\n\n")
        names = [r[hdr.index("Kernel Name")][:60] for r in data]
        out.write("launches captured: " + "; ".join(f"`{n}`" for n in names) + "\n\n| metric | unit | " + " | ".join(f"#{i}" for i in range(len(data))) + " |\n")
        out.write("|---|---|" + "---:|" * len(data) + "\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.write(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in data) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
