"""Summarise .ncu-rep captures (ncu --set full) into a small markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_top.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_summary.md
"""
import csv
import io
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
]


def main():
    print("| kernel | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|" + "---|" * len(WANT))
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print(f"| **{rep.split('/')[-1]}** |" + " |" * len(WANT))
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
            name = re.sub(r"^b200::", "", name)
            cells = []
            for key, _ in WANT:
                if key in idx:
                    v, u = r[idx[key]], units[idx[key]]
                    try:
                        f = float(v.replace(",", ""))
                        v = f"{f:.3g}" if abs(f) < 1000 else f"{f:.0f}"
                    except ValueError:
                        pass
                    cells.append(f"{v} {u}".strip())
                else:
                    cells.append("-")
            print(f"| `{name[:60]}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
