"""Print selected metrics of every launch in an .ncu-rep (read here, on the CPU box): python tools/ncu_metrics.py rep [substr ...]"""
import csv
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct", "sm__pipe_tc_cycles_active",
           "sm__inst_executed_pipe_tc", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
           "lts__throughput.avg.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct",
           "smsp__issue_active.avg.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp", "smsp__warp_issue_stalled",
           "l1tex__throughput.avg.pct", "sm__throughput.avg.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
           "sm__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed"]


def main():
    rep = sys.argv[1]
    keys = sys.argv[2:] or DEFAULT
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:100])
        for h, u, v in zip(hdr, units, r):
            if any(k in h for k in keys):
                print("   %-95s %-10s %s" % (h, u, v))


if __name__ == "__main__":
    main()
