"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small CSV of the metrics DESIGN.md quotes."""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
        "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in KEYS if k in idx]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keys)
        w.writerow([units[idx[k]] for k in keys])
        for r in rows[2:]:
            w.writerow([r[idx[k]] for k in keys])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
