"""Prints the handful of ncu metrics DESIGN.md / profiles/ cite from an .ncu-rep (first kernel in it)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_fmaheavy.avg.pct", "sm__inst_executed_pipe_xu.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "sm__pipe_tensor", "sm__cycles_elapsed.avg",
        "lts__t_sector_hit_rate", "issue_stalled")
extra = sys.argv[2:]
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w) or (w == "issue_stalled" and "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h) for w in want) or any(x in h for x in extra):
        print("%s [%s] = %s" % (h, u, v))
