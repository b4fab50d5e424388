"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.

    python tools/launch_summary.py gpurun_out/launches.csv [top]
"""
import collections
import csv
import sys


def main():
    fn = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = [r for r in csv.reader(open(fn)) if len(r) > 5]
    hdr = next(r for r in rows if r[0] == "ID")
    agg = collections.OrderedDict()
    for r in rows:
        if r[0] == "ID":
            continue
        name = r[hdr.index("Kernel Name")][:70]
        v = float(r[hdr.index("Metric Value")].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    unit = next(r for r in rows if r[0] != "ID")[hdr.index("Metric Unit")]
    tot = sum(a[1] for a in agg.values())
    print("%s: total %.1f %s" % (fn, tot, unit))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print("  %-72s n=%5d %14.1f %5.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))


if __name__ == "__main__":
    main()
