"""Summarise an .ncu-rep (raw page) into the handful of metrics the design notes quote."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
        ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe%"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit%"), ("lts__t_sector_hit_rate.pct", "l2_hit%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("lts__t_sectors.sum", "l2_sectors(32B)"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("smsp__warps_eligible.avg.per_cycle_active", "eligible"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_not_sel"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_no_inst"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_dispatch"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar"),
        ("local_load... ", "")]
for d in data:
    name = d[idx["Kernel Name"]]
    short = name.split("::")[-1][:40]
    out = [short]
    for m, lab in want:
        if m in idx and lab:
            v = d[idx[m]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            out.append(f"{lab}={v}{units[idx[m]] if lab in ('time','dram_rd','dram_wr','l2_bytes') else ''}")
    print("  ".join(out))
    print()
