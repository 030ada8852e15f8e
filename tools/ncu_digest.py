"""Digest an .ncu-rep: key throughput / stall metrics per kernel.  Usage: python tools/ncu_digest.py rep [filter]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        print("==", row[hdr.index("Kernel Name")][:70])
        for i, h in enumerate(hdr):
            if h in KEYS or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                v = row[i]
                try:
                    if float(v) == 0:
                        continue
                except ValueError:
                    pass
                print(f"  {h.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', ''):75s} {v:>16s} {units[i]}")


if __name__ == "__main__":
    main()
