"""Profiling helper (not a test): compact CSV of the metrics that matter from ncu --set full reports.

    python tools/ncu_summary.py OUT.csv label1=report1.ncu-rep [label2=report2.ncu-rep ...]
"""
import csv
import subprocess
import sys

METRICS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "math_pipe_throttle", "mio_throttle", "lg_throttle", "wait",
          "not_selected", "no_instruction", "branch_resolving", "dispatch_stall"]


def main():
    out = sys.argv[1]
    rows_out = []
    head = ["capture", "kernel"] + METRICS + ["stall_" + s for s in STALLS]
    for arg in sys.argv[2:]:
        label, rep = arg.split("=", 1)
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            def g(k):
                return r[hdr.index(k)] if k in hdr else ""
            row = [label, g("Kernel Name").split("(")[0].replace("void ", "")]
            for m in METRICS:
                v = g(m)
                u = units[hdr.index(m)] if m in hdr else ""
                row.append((v + " " + u).strip())
            for s in STALLS:
                row.append(g("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s)[:6])
            rows_out.append(row)
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(head)
        w.writerows(rows_out)
    print("wrote %d kernels to %s" % (len(rows_out), out))


if __name__ == "__main__":
    main()
