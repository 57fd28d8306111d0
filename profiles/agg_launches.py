"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path, top=25):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        name = row["Kernel Name"][:72]
        t = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        t = t / 1000 if unit == "ns" else t * 1000 if unit == "ms" else t
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += t
        tot += t
    print("%10s %6s %12s %6s  %s" % ("total_us", "n", "us/launch", "share", "kernel"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%10.1f %6d %12.1f %5.1f%%  %s" % (v[1], v[0], v[1] / v[0], 100 * v[1] / tot, k))
    print("%10.1f total" % tot)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
