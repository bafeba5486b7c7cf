"""Turns the raw ncu outputs brought back in gpurun_out/ into the small tracked summaries in profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1b.csv profiles/r1_launches_scene10m.txt
    python profiles/summarize.py kernel   gpurun_out/prof_feat_r1b.ncu-rep profiles/r1_feature_kernel_v2_ncu.txt

`launches` aggregates the `--metrics gpu__time_duration.sum` launch list per kernel (cold-cache,
serialised times: compare SHARES, not absolutes).  `kernel` extracts the metrics the roofline
discussion in DESIGN.md quotes from one `ncu --set full` capture.
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__warps_eligible.avg.per_cycle_active",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki].split("(")[0][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# source: %s (ncu --metrics gpu__time_duration.sum --clock-control none)\n# total %.3f ms over %d launches\n" % (src, tot, sum(a[0] for a in agg.values())))
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-92s n=%4d %12.3f ms %6.2f%%\n" % (k, c, t, 100 * t / tot))


def kernel(src, dst):
    out = subprocess.check_output(["ncu", "-i", src, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# source: %s (ncu --set full --clock-control none --import-source on)\n" % src)
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write("## %s  grid %s block %s\n" % (d.get("Kernel Name", "?"), d.get("Grid Size", d.get("launch__grid_size", "?")), d.get("Block Size", "?")))
            for k in KEYS:
                if k in d:
                    f.write("%-90s %-12s %s\n" % (k, units[hdr.index(k)], d[k]))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
