"""Dev: compact per-kernel summary of an .ncu-rep (raw page) -> CSV on stdout."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "hmma%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short")]
w = csv.writer(sys.stdout)
w.writerow([c for _, c in cols])
w.writerow([units[idx[h]] if h in idx else "" for h, _ in cols])
for r in rows[2:]:
    out = []
    for h, c in cols:
        v = r[idx[h]] if h in idx else ""
        if c == "kernel":
            v = v.split("(")[0].replace("void ", "")[:48]
        out.append(v)
    w.writerow(out)
