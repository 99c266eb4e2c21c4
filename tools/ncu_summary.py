"""Key metrics of every kernel in an .ncu-rep (ncu --set full capture) as a markdown table.
usage: python tools/ncu_summary.py gpurun_out/<name>.ncu-rep [more.ncu-rep ...] > profiles/<name>.md"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % (of ncu peak)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def main(paths):
    for path in paths:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            print(f"## {path}\n\n(no kernels)\n")
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"## {path}\n")
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
            print(f"### `{name}`  (launch id {r[col['ID']]})\n")
            print("| metric | value | unit |")
            print("|---|---:|---|")
            for key, label in METRICS:
                if key in col:
                    print(f"| {label} (`{key}`) | {r[col[key]]} | {units[col[key]]} |")
            print()


if __name__ == "__main__":
    main(sys.argv[1:])
