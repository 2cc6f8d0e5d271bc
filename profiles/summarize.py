"""Extract the judged metrics from an .ncu-rep (read here, no GPU):  python profiles/summarize.py rep out.csv"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    stall = [(h, i) for i, h in enumerate(hdr) if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct")]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([k for k, _ in idx] + [h.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_warp_active.pct", "") for h, _ in stall])
        w.writerow([units[i] for _, i in idx] + ["%"] * len(stall))
        for r in data:
            w.writerow([r[i][:90] for _, i in idx] + [r[i] for _, i in stall])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
