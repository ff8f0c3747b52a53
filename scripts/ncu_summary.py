"""Print the metrics we track from an .ncu-rep (raw page):  python scripts/ncu_summary.py file.ncu-rep [more...]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64", "pipe_fp64",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct", "lts__throughput.avg.pct", "smsp__issue_active.avg.pct",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
        "launch__shared_mem_per_block", "sm__sass_thread_inst_executed_op_d", "local_load", "local_store",
        "smsp__pcsamp_warps_issue_stalled", "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared"]
for f in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", f, r[hdr.index("Kernel Name")][:90])
        for h, u, v in zip(hdr, units, r):
            if any(w in h for w in WANT) and v not in ("", "0", "n/a"):
                print("  %-95s %s %s" % (h, v, u))
