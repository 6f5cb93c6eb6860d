#!/usr/bin/env python3
"""Turns gpurun_out/<launches>.csv (ncu --metrics gpu__time_duration.sum) and <prof>.ncu-rep (ncu --set full)
into the text summaries committed under profiles/.   usage: summarise.py launches.csv prof.ncu-rep > out.txt"""
import csv
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct"]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    t = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        t[r[ki][:90]][0] += 1
        t[r[ki][:90]][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in t.values())
    print("== launch list (%s): per-kernel device time, cold-cache serialised; compare SHARES ==" % path)
    for n, (c, s) in sorted(t.items(), key=lambda x: -x[1][1]):
        print("%-92s n=%4d total_ns=%14.0f avg_ns=%12.0f share=%.4f" % (n, c, s, s / c, s / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    print("\n== ncu --set full (%s) ==" % path)
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:140])
        for w in WANT:
            if w in hdr:
                print("  %-78s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
        # warp-state (PC sampling) shares: where the resident warps spend their time
        samp = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    samp[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(samp.values())
        if tot > 0:
            print("  warp-state samples: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot)
                                                       for k, v in sorted(samp.items(), key=lambda x: -x[1]) if v / tot >= 0.01))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        (launches if p.endswith(".csv") else full)(p)
