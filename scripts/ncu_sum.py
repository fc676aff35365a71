"""Key metrics of one ncu report (raw page).  usage: ncu_sum.py file.ncu-rep [nodes]"""
import csv, subprocess, sys
rep = sys.argv[1]; nodes = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines())); h = r[0]; v = r[2]
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in h:
        x = v[h.index(k)]
        try:
            f = float(x.replace(",", ""))
            extra = f"   ({f / nodes:.2f} per node)" if nodes and f > 1e5 else ""
        except ValueError:
            extra = ""
        print(f"{k:75s} {x} {r[1][h.index(k)]}{extra}")
