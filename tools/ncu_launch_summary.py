"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, average, share."""
import collections
import csv
import sys


def main():
    lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("total %.1f us over %d launches" % (tot, sum(len(v) for v in agg.values())))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-44s n=%4d  sum=%10.1f us  avg=%8.2f us  share=%5.1f%%" % (k[:44], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    main()
