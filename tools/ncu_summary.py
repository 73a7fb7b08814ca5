"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md / profiles/ cite.
usage: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
]


def main():
    for rep in sys.argv[1:]:
        raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        for r in data:
            print("== %s :: %s" % (rep, r[hdr.index("Kernel Name")]))
            for k in KEYS:
                if k in hdr:
                    print("  %-75s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
            for i, h in enumerate(hdr):
                if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct") or ("average_warps_issue_stalled" in h and "not_issued" not in h):
                    try:
                        v = float(r[i])
                    except ValueError:
                        continue
                    if v >= 0.05:
                        print("  %-75s %s %s" % (h, r[i], units[i]))


if __name__ == "__main__":
    main()
