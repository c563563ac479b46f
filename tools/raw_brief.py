"""Key figures of `ncu --page raw --csv` exports.  Usage: raw_brief.py file_raw.csv [...]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "smsp__cycles_active.avg",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_inst0.avg.pct_of_peak_sustained_active"]
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("==", f, vals[hdr.index("Kernel Name")][:60] if "Kernel Name" in hdr else "")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:90s} {vals[i]:>14s} {units[i]}")
    st = []
    for i, k in enumerate(hdr):
        if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                st.append((float(vals[i].replace(",", "")), k.split("stalled_")[1]))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    print("  stalls:", ", ".join(f"{k} {100 * v / tot:.0f}" for v, k in sorted(st, reverse=True)[:7]))
