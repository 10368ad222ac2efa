"""Condense an .ncu-rep into the few numbers the roofline discussion needs.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {path}: {len(data)} profiled launch(es)")
    name_i = hdr.index("Kernel Name")
    for d in data:
        print("kernel:", d[name_i][:120])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} [{units[i]}]: " + ", ".join(d[i] for d in data))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-id", ":::1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        hdr = rows[1]
        data = [r for r in rows[2:] if len(r) == len(hdr)]
        cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = {h: sum(int(r[i]) for r in data if r[i].isdigit()) for i, h in cols}
        s = sum(tot.values()) or 1
        print("warp stall samples (first launch): " +
              ", ".join(f"{k}={100 * v / s:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
        i_s, i_src = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
        top = sorted(data, key=lambda r: -int(r[i_s]) if r[i_s].isdigit() else 0)[:8]
        print("hottest SASS instructions (samples):")
        for r in top:
            print(f"  {r[i_s]:>6}  {r[i_src].strip()[:80]}")


if __name__ == "__main__":
    main(sys.argv[1])
