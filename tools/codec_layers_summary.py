"""Summarise the ncu per-launch csv made with tools/codec_layers.py (see its docstring) into a table."""
import collections
import csv
import re
import sys

U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    by = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = by.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("vnb::", "")})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * U.get(r["Metric Unit"], 1)
    tot = 0.0
    print(f"# {path}: one encode + one decode of 32 ten-second clips, ncu --clock-control none (cold-cache, serialised)")
    print(f"# {'id':>3} {'kernel':24s} {'us':>9} {'tensor%':>7} {'rd GB':>6} {'wr GB':>6} {'TB/s':>5}")
    for i, d in by.items():
        t = d["gpu__time_duration.sum"]
        rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        tot += t
        print(f"{i:>5} {d['name'][:24]:24s} {t:9.1f} {d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']:7.1f} "
              f"{rd / 1e9:6.2f} {wr / 1e9:6.2f} {(rd + wr) / t / 1e6:5.2f}")
    print(f"# total {tot / 1e3:.1f} ms")


if __name__ == "__main__":
    main(sys.argv[1])
