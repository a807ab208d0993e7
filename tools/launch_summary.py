"""Summarise an ncu --metrics gpu__time_duration.sum launch list (csv): per-kernel totals of the LAST
proof in the log (from the last k_to_mont before the last k_matvec to the end).
Usage: python tools/launch_summary.py gpurun_out/x_launches.csv [--all]"""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    seq = []
    for row in csv.DictReader(lines):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"us": 1e-3, "ns": 1e-6, "ms": 1, "s": 1e3}[row["Metric Unit"]]
        seq.append((row["Kernel Name"], v, row["Grid Size"], row["Block Size"]))
    return seq


def main():
    seq = load(sys.argv[1])
    idx = [i for i, s in enumerate(seq) if s[0].startswith("k_matvec")]
    start = 0 if "--all" in sys.argv or not idx else idx[-1] - 1
    agg = collections.OrderedDict()
    for n, v, g, b in seq[start:]:
        a = agg.setdefault(n[:72], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{len(seq)} launches in log; summarising {len(seq) - start}; total {tot:.3f} ms")
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v:9.3f} ms {100 * v / tot:5.1f}% x{c:3d}  {k}")


if __name__ == "__main__":
    main()
