"""Prints the metrics of an .ncu-rep that the profiles/ summaries quote:  python profiles/ncu_keys.py file.ncu-rep [kernel-substring]"""
import csv, subprocess, sys
rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    if sub and sub not in r[ki]:
        continue
    print("==", r[ki][:90])
    for i, h in enumerate(hdr):
        if h in want:
            print("  %-75s %-12s %s" % (h, units[i], r[i]))
    st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h and r[i]]
    st.sort(reverse=True)
    for v, h in st[:7]:
        print("  stall %-68s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
