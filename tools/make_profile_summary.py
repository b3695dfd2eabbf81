"""profiles/ summary from an ncu report: key raw metrics + hottest source lines.  usage: <rep> <lib.so> <kernel-substr> <out.txt>"""
import csv, subprocess, sys
rep, so, kern, out = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
lines = [f"ncu --set full --clock-control none --import-source on  ({rep})", ""]
d = dict(zip(hdr, zip(units, vals)))
for k in want:
    if k in d:
        lines.append(f"{k:70s} {d[k][1]} {d[k][0]}")
lines.append("")
lines.append("warp stall reasons (average warps stalled per issue-active cycle):")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        lines.append(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {float(d[h][1]):8.3f}")
lines.append("")
lines.append("hottest source lines (ncu source page joined with nvdisasm -g line info; instr = share of executed warp instructions, samples = share of stall samples):")
src = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, so, kern, "30"], capture_output=True, text=True).stdout
lines += src.splitlines()
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
