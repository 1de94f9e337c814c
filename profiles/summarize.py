#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed here.
usage: python profiles/summarize.py launches <launches.csv> <out.txt>
       python profiles/summarize.py kernel <report.ncu-rep> <out.txt>"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max.per_second", "sm__inst_executed.sum.per_cycle_active",
        "smsp__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum"]


def launches(src, dst):
    lines = [l for l in open(src) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for row in r:
        v = float(row[vi].replace(",", ""))
        v = v / 1000.0 if row[ui] in ("ns", "nsecond") else v
        seq.append((row[ki].split("(")[0].replace("void ", ""), v))
    agg = collections.OrderedDict()
    for n, v in seq:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in seq)
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; {len(seq)} launches, {tot/1000:.2f} ms total\n")
        f.write("# per-kernel totals (cold-cache, serialised: compare SHARES, not absolutes)\n")
        for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{n:60s} launches={c:4d} total_us={v:12.1f} avg_us={v/c:10.1f} share={100*v/tot:5.1f}%\n")
        f.write("\n# launch sequence (us)\n")
        for n, v in seq:
            f.write(f"{n:60s} {v:10.1f}\n")


def kernel(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, source {src}\n")
        for row in rows:
            f.write(f"\n== {row[ki][:120]}\n")
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write(f"{h:75s} {row[i]:>18s} {units[i]}\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
