"""ncu report -> short text summary (run here, no GPU needed):
    python profiles/summarize.py gpurun_out/encode_r01.ncu-rep > profiles/encode_r01.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__cycles_elapsed.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed_op_global_atom.sum",
    "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("kernel:", name)
        for k in KEYS:
            if k in hdr:
                print("  %-62s %s %s" % (k, row[hdr.index(k)], units[hdr.index(k)]))
        print("  stall reasons (warps per issue-active cycle, > 0.05):")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    if float(row[i]) > 0.05:
                        print("    %-40s %s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), row[i]))
                except ValueError:
                    pass


if __name__ == "__main__":
    main(sys.argv[1])
