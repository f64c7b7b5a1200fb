"""Key metrics of an ncu --set full report (raw page), one block per kernel launch."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp32.sum" ,"smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_red.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("====", d["Kernel Name"][:80])
        for k in hdr:
            if k in WANT or ("issue_stalled" in k and k.endswith("per_issue_active.ratio")):
                try:
                    val = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                if "issue_stalled" in k and val < 0.15:
                    continue
                print(f"  {k:82s} {d[k]:>16s} {u[k]}")


if __name__ == "__main__":
    main(sys.argv[1])
