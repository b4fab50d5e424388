"""Per-source-line instruction / stall-sample shares of one kernel of an ncu report, in SOURCE order
with the source text next to it (companion of ncu_by_line.py, which sorts by samples).

    python tools/ncu_lines.py prof.ncu-rep contact_map_b200/libcmb200.so MANGLED_SUBSTR [min_pct]
"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_by_line  # noqa: E402


def main():
    rep, so, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    table = ncu_by_line.line_table(so, kernel)
    if len(body) != len(table):
        print("warning: %d SASS rows in ncu vs %d in nvdisasm (pick a more specific kernel substring)"
              % (len(body), len(table)))
    col = {n: hdr.index(n) for n in hdr}
    agg = defaultdict(lambda: [0.0, 0.0, 0])
    for r, (_, where, _) in zip(body, table):
        a = agg[where]
        a[0] += float(r[col["Instructions Executed"]] or 0)
        a[1] += float(r[col["# Samples"]] or 0)
        a[2] += 1
    tot = sum(a[0] for a in agg.values())
    samp = sum(a[1] for a in agg.values())
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "contact_map_b200", "csrc")
    src = {f: open(os.path.join(root, f), errors="replace").read().split("\n") for f in os.listdir(root)}
    print("total warp instructions %.4e, samples %d" % (tot, samp))
    print("%-26s %4s %7s %7s  source" % ("file:line", "sass", "inst%", "samp%"))
    for where, a in sorted(agg.items()):
        if 100 * a[0] / tot < min_pct and 100 * a[1] / max(samp, 1) < min_pct:
            continue
        line = src[where[0]][where[1] - 1].strip()[:100] if where[0] in src and where[1] <= len(src[where[0]]) else ""
        print("%-26s %4d %6.2f%% %6.2f%%  %s" % ("%s:%d" % where, a[2], 100 * a[0] / tot, 100 * a[1] / max(samp, 1), line))


if __name__ == "__main__":
    main()
