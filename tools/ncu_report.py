"""Summarise every kernel launch of an .ncu-rep (ncu --set full): headline metrics, tensor / issue / memory utilisation
and the top stall reasons, one block per launch.   usage: python tools/ncu_report.py report.ncu-rep [name-filter]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_issued.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
name_col = hdr.index("Kernel Name")
for r in rows[2:]:
    if flt and flt not in r[name_col]:
        continue
    d = dict(zip(hdr, r))
    print(f"== {r[name_col][:110]}")
    for k in keys:
        if k in d and d[k] not in ("", "0"):
            print(f"  {k:88s} {d[k]:>20s} {units[hdr.index(k)]}")
    stalls = sorted(((float(d[k]), k) for k in d if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and d[k]),
                    reverse=True)[:5]
    print("  top stalls per issue: " + ", ".join(f"{k.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for v, k in stalls))
    print()
