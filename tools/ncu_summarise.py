"""Summarise an `ncu --set full` report into the handful of rows DESIGN.md / profiles/README.md quote.

Usage:  python tools/ncu_summarise.py gpurun_out/x.ncu-rep > profiles/rNN_x_summary.md
Reads the report with `ncu -i … --page raw --csv` (works without a GPU)."""
import csv
import io
import subprocess
import sys

ROWS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (hmma) active, % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput, % of peak"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of peak"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy, % of peak warps"),
    ("smsp__cycles_active.avg", "SMSP active cycles (avg)"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of `{path.split('/')[-1]}`\n")
    for r in rows[2:]:
        print(f"## {r[col['Kernel Name']]}\n")
        print("| metric | value |")
        print("|---|---|")
        for key, label in ROWS:
            if key in col:
                print(f"| {label} (`{key}`) | {r[col[key]]} {units[col[key]]} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
