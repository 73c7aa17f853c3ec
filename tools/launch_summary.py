"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        agg[(row["Kernel Name"].split("(")[0][:70], row["Grid Size"], row["Block Size"])].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v):10.1f}us {100 * sum(v) / tot:5.1f}% n={len(v):4d} avg={sum(v) / len(v):8.2f} min={min(v):7.2f} max={max(v):8.2f} {k}")


if __name__ == "__main__":
    main(sys.argv[1])
