"""Summarise an .ncu-rep (raw page) into a compact per-kernel table.  Usage: python tools/ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
M = [("time_us", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"), ("occ_regs", "launch__occupancy_limit_registers"),
     ("occ_smem", "launch__occupancy_limit_shared_mem"), ("warps_act%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
     ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("fma_pipe%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
     ("fma_inst%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), ("lsu_inst%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
     ("smem_wf%", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
     ("bank_confl", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), ("smem_wf", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
     ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
     ("l2_hit%", "lts__t_sector_hit_rate.pct"), ("inst", "smsp__inst_executed.sum"),
     ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
     ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
     ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
     ("stall_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
     ("stall_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
     ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
     ("stall_notsel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
     ("stall_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
     ("stall_dispatch", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"),
     ("stall_nosch", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio")]
ki = hdr.index("Kernel Name")
out = []
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "")
    line = [f"{name[:28]:28s}"]
    for short, m in M:
        if m in hdr:
            v = r[hdr.index(m)].replace(",", "")
            try:
                fv = float(v); u = units[hdr.index(m)]
                if u == "byte": fv /= 1e6
                if u == "Kbyte": fv /= 1e3
                if u == "Gbyte": fv *= 1e3
                if u == "ns": fv /= 1e3
                if u == "ms": fv *= 1e3
                v = f"{fv:.4g}"
            except ValueError:
                pass
            line.append(f"{short}={v}")
    out.append("  ".join(line))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
