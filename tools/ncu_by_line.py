"""Aggregate an ncu SASS-level source page by CUDA source line.

    ncu -i prof.ncu-rep --page source --csv > sass.csv
    python tools/ncu_by_line.py sass.csv contact_map_b200/libcmb200.so k_frames_small [top]

Joins the per-instruction rows of ncu (same order as the SASS of the kernel) with the line
table of `nvdisasm -g`, and prints instructions executed / stall samples per source line.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_table(so_path, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so_path)], cwd=tmp, check=True,
                   capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True,
                         text=True).stdout.splitlines()
    out, inside, cur = [], False, ("?", 0)
    for ln in dis:
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def main():
    sass_csv, so_path, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    table = line_table(so_path, kernel)
    if len(table) != len(body):
        print("warning: %d SASS rows in ncu vs %d in nvdisasm" % (len(body), len(table)))
    col = {name: hdr.index(name) for name in hdr}
    agg = defaultdict(lambda: defaultdict(float))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r, (_, where, _) in zip(body, table):
        a = agg[where]
        a["inst"] += float(r[col["Instructions Executed"]] or 0)
        a["samples"] += float(r[col["# Samples"]] or 0)
        a["sass"] += 1
        for s in stall_cols:
            a[s] += float(r[col[s]] or 0)
    tot_inst = sum(a["inst"] for a in agg.values())
    tot_samp = sum(a["samples"] for a in agg.values())
    print("total warp instructions %.3e, samples %d" % (tot_inst, tot_samp))
    print("%-28s %6s %8s %8s  top stalls" % ("file:line", "sass", "inst%", "samp%"))
    for where, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        stalls = sorted(((a[s], s[6:]) for s in stall_cols), reverse=True)[:3]
        print("%-28s %6d %7.2f%% %7.2f%%  %s" % ("%s:%d" % where, a["sass"], 100 * a["inst"] / tot_inst,
                                                  100 * a["samples"] / max(tot_samp, 1),
                                                  ", ".join("%s %.0f" % (n, v) for v, n in stalls if v)))


    # the same numbers by code region of the pair phase (line ranges of the CURRENT sources: the
    # function a line belongs to is found from the nearest preceding definition)
    regions = source_regions()
    by_region = defaultdict(lambda: [0.0, 0.0])
    for (fname, line), a in agg.items():
        name = "%s: %s" % (fname, regions.get(fname, lambda _l: "-")(line))
        by_region[name][0] += a["inst"]
        by_region[name][1] += a["samples"]
    print("\n%-58s %8s %8s" % ("function (file: name)", "inst%", "samp%"))
    for name, (inst, samp) in sorted(by_region.items(), key=lambda kv: -kv[1][0])[:24]:
        print("%-58s %7.2f%% %7.2f%%" % (name, 100 * inst / tot_inst, 100 * samp / max(tot_samp, 1)))


def source_regions():
    """{file name: line -> name of the enclosing __device__/__global__ function} for csrc/."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "contact_map_b200", "csrc")
    out = {}
    for fname in os.listdir(root):
        starts = []
        for no, text in enumerate(open(os.path.join(root, fname), errors="replace"), 1):
            text = re.sub(r"__launch_bounds__\([^)]*\)", "", text)
            m = re.match(r"\s*(?:static\s+)?(?:__device__|__global__|__host__).*?\b(\w+)\s*\(", text) or \
                re.match(r"\s*(?:__launch_bounds__\S*\s+)?(\w+)\(const .*Params", text)
            if m and m.group(1) not in ("if", "for", "while", "__launch_bounds__"):
                starts.append((no, m.group(1)))

        def find(line, _s=starts):
            name = "-"
            for no, n in _s:
                if no > line:
                    break
                name = n
            return name
        out[fname] = find
    return out


if __name__ == "__main__":
    main()
