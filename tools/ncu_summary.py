"""Condense `ncu -i report.ncu-rep --page raw --csv` output into the small JSON summaries kept under profiles/.
Usage: python tools/ncu_summary.py raw.csv out.json "free-text source line" """
import csv
import json
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def main():
    raw, out, source = sys.argv[1], sys.argv[2], sys.argv[3]
    with open(raw, newline="") as f:
        rows = [r for r in csv.reader(f) if r]
    # the raw page has a header row, a units row, then one row per kernel instance
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, body = rows[hdr_i], rows[hdr_i + 1], rows[hdr_i + 2:]
    kernels = []
    for r in body:
        d = {}
        for k in KEEP:
            if k in hdr:
                j = hdr.index(k)
                u = units[j] if j < len(units) else ""
                d[k] = (r[j] + (" " + u if u and k != "Kernel Name" else "")).strip()
        kernels.append(d)
    with open(out, "w") as f:
        json.dump({"source": source, "kernels": kernels}, f, indent=1)
    print(out, len(kernels), "kernels")


if __name__ == "__main__":
    main()
