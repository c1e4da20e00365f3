"""Summarise an .ncu-rep of the relaxation kernel: headline metrics + per-region instruction counts.
usage: python tools/ncu_summary.py report.ncu-rep probes flips_per_probe warps_per_cta"""
import csv, subprocess, sys, io

rep, probes, fpp, warps = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_issued.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
keys += sorted(k for k in d if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"))
for k in keys:
    if k in d and d[k] not in ("", "0"):
        print(f"{k:90s} {d[k]:>18s} {u.get(k, '')}")
flips = probes * fpp
print(f"\nflips {flips:.0f}; instructions per warp per flip = {float(d['smsp__inst_issued.sum']) / flips / warps:.1f}")
