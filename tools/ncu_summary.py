#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full) into the compact per-launch table kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_ntt_r01b.ncu-rep profiles/ncu_full_ntt_pass_r01b.csv
"""
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
KEEP_PREFIX = ("smsp__average_warps_issue_stalled_",)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h in KEEP or h.startswith(KEEP_PREFIX) or h == "Kernel Name"]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % k for k in range(len(data))])
        for i in cols:
            w.writerow([hdr[i], units[i]] + [r[i] for r in data])
    print("wrote", out, len(cols), "metrics x", len(data), "launches")


if __name__ == "__main__":
    main()
