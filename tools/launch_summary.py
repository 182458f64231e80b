"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = row["Kernel Name"].replace("void ", "").replace("(anonymous namespace)::", "")
        m = re.search(r"(\w+)\s*(<|\()", name)
        rows.append((m.group(1) if m else name[:40], v, row["Grid Size"], row["Block Size"]))
    return rows


def main():
    path = sys.argv[1]
    nsteps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    rows = load(path)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v, _, _ in rows:
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v, _, _ in rows)
    print("launches %d, total %.1f us (%.2f ms per step over %g steps)" % (len(rows), tot, tot / nsteps / 1e3, nsteps))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
        print("%-36s %6d %11.1f us %5.1f%%  avg %7.1f" % (k[:36], n, t, 100 * t / tot, t / n))


if __name__ == "__main__":
    main()
