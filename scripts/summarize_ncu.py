"""Turn an .ncu-rep into a small tracked summary:  python scripts/summarize_ncu.py in.ncu-rep out.csv"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
        "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_atom_dot_alu_lookup_hit.sum",
        "lts__t_sectors_srcunit_tex_op_atom_dot_alu_lookup_miss.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines()) if r]
hdr, units = rows[0], rows[1]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "metric", "unit", "value"])
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        for i, h in enumerate(hdr):
            if h in KEYS or ("issue_stalled" in h and h.endswith("_per_warp_active.pct")):
                w.writerow([name[:80], h, units[i], vals[i]])
print("wrote", sys.argv[2])
