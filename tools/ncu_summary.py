"""Text summary of an .ncu-rep (key raw metrics per kernel + hottest source lines) for profiles/.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/name.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
print("# ncu summary of", rep.split("/")[-1], "(ncu --set full --clock-control none; replayed, cold-cache: compare shares, not absolutes)")
for r in rows[2:]:
    d = dict(zip(h, r))
    print("\n## kernel:", d.get("Kernel Name", "?")[:100], " id", d.get("ID"))
    for k in KEYS:
        if k in d:
            print("  %-82s %s %s" % (k, d[k], u[h.index(k)]))
    try:
        rd = float(d["dram__bytes_read.sum"]); wr = float(d["dram__bytes_write.sum"])
        ur, uw = u[h.index("dram__bytes_read.sum")], u[h.index("dram__bytes_write.sum")]
        sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        print("  %-82s %.0f byte" % ("dram traffic per launch (read+write)", rd * sc[ur] + wr * sc[uw]))
    except Exception:
        pass
if len(sys.argv) > 2 and sys.argv[2] == "lines":
    out = subprocess.run([sys.executable, __file__.replace("ncu_summary", "ncu_lines"), rep, "30"], capture_output=True, text=True).stdout
    print("\n## hottest source lines of the first kernel (share of executed warp instructions, avg active lanes, share of stall samples)")
    print(out)
