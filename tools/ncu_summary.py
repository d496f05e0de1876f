"""Summarise an .ncu-rep (read here, no GPU needed):  python tools/ncu_summary.py file.ncu-rep [more metrics...]"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum"]


def main():
    path = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for w in WANT + extra:
        idx = [i for i, h in enumerate(hdr) if h == w] or [i for i, h in enumerate(hdr) if w in h][:3]
        for i in idx:
            print("%-92s %-8s %s" % (hdr[i][:92], units[i][:8], [d[i] for d in data]))


if __name__ == "__main__":
    main()
