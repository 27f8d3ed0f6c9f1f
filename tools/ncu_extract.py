#!/usr/bin/env python
"""Extract the handful of counters the design argues from out of an ncu report (run here, no GPU needed):
    python tools/ncu_extract.py gpurun_out/r02_gemv_10M.ncu-rep "title" "command line" > profiles/<name>_ncu_summary.md"""
import csv, io, subprocess, sys
rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
WANT = ["launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.min", "sm__cycles_active.max",
        "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
print(f"# {title}\n\nCommand: `{cmd}`\n")
print("| metric | " + " | ".join(f"launch {i + 1}" for i in range(len(data))) + " | unit |")
print("|---|" + "---|" * (len(data) + 1))
print("| Kernel Name | " + " | ".join(r[col["Kernel Name"]] for r in data) + " |  |")
for m in WANT:
    hits = [h for h in hdr if h == m] or [h for h in hdr if h.startswith(m)]
    for h in hits[:1]:
        print(f"| {h} | " + " | ".join(r[col[h]] for r in data) + f" | {units[col[h]]} |")
