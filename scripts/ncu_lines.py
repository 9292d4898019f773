#!/usr/bin/env python
"""Aggregate an ncu SASS source page per CUDA source line.

    ncu -i X.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all lib.so ; nvdisasm -g -c lib.cubin > dis.txt
    python scripts/ncu_lines.py sass.csv dis.txt <mangled-kernel-substring> [source.cu]
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass_csv, dis, kern = sys.argv[1:4]
    src = sys.argv[4] if len(sys.argv) > 4 else None
    rows = list(csv.reader(open(sass_csv)))
    for i, r in enumerate(rows[:6]):
        if "Source" in r and "Address" in r:
            hdr, start = r, i + 1
            break
    ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    sass = [(r[1].strip(), int(r[ci] or 0), int(r[cs] or 0), int(r[ct] or 0)) for r in rows[start:] if len(r) > ci]
    # line info from nvdisasm
    lines = open(dis).read().split("\n")
    in_k, cur, per_instr = False, (None, 0), []
    for ln in lines:
        if ln.startswith(".text.") and ln.endswith(":"):
            in_k = kern in ln
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
        if m:
            per_instr.append((cur, m.group(2).strip()))
    n = min(len(sass), len(per_instr))
    agg = defaultdict(lambda: [0, 0, 0])
    tot = 0
    for k in range(n):
        (f, l), _ = per_instr[k]
        agg[(f, l)][0] += sass[k][1]
        agg[(f, l)][1] += sass[k][2]
        agg[(f, l)][2] += sass[k][3]
        tot += sass[k][1]
    text = {}
    if src:
        for i, t in enumerate(open(src).read().split("\n")):
            text[i + 1] = t.strip()[:100]
    print("sass instrs", len(sass), "disasm instrs", len(per_instr), "total warp-instr executed", tot)
    tsamp = sum(v[1] for v in agg.values()) or 1
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:50]:
        print("%-14s:%4d  inst %9d %5.1f%%  samples %6d %5.1f%%  | %s" % (f, l, v[0], 100.0 * v[0] / tot, v[1], 100.0 * v[1] / tsamp,
                                                                      text.get(l, "") if f.endswith(".cu") else ""))


if __name__ == "__main__":
    main()


def regions(sass_csv, dis, kern, bounds):
    """Attribute instructions/samples to program regions given as {name: (first_line, last_line)} of the
    kernel's own source file; helper-header lines inherit the region of the last kernel-file line seen."""
    rows = list(csv.reader(open(sass_csv)))
    for i, r in enumerate(rows[:6]):
        if "Source" in r and "Address" in r:
            hdr, start = r, i + 1
            break
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    sass = [(int(r[ci] or 0), int(r[cs] or 0)) for r in rows[start:] if len(r) > ci]
    lines = open(dis).read().split("\n")
    in_k, cur, per = False, None, []
    main_file = None
    for ln in lines:
        if ln.startswith(".text.") and ln.endswith(":"):
            in_k = kern in ln
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            f = m.group(1).split("/")[-1]
            if main_file is None:
                main_file = f
            if f == bounds["_file"]:
                cur = int(m.group(2))
            continue
        if re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln):
            per.append(cur)
    agg = defaultdict(lambda: [0, 0])
    for k in range(min(len(sass), len(per))):
        name = "other"
        for nm, rng in bounds.items():
            if nm != "_file" and per[k] is not None and rng[0] <= per[k] <= rng[1]:
                name = nm
        agg[name][0] += sass[k][0]
        agg[name][1] += sass[k][1]
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    for nm, v in agg.items():
        print("%-10s inst %10d %5.1f%%   samples %7d %5.1f%%" % (nm, v[0], 100.0 * v[0] / ti, v[1], 100.0 * v[1] / ts))
