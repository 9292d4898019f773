#!/usr/bin/env python
"""Split an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) into loop regions and print, per region,
executed warp-instructions, samples and the stall mix.  Regions = maximal address ranges between backward branches
whose executed count is large (the sweeps) -- a quick way to see which sweep the samples sit in.

    python scripts/ncu_regions.py source.csv [min_exec]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) > 10]
ci, cs = ix["Instructions Executed"], ix["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_i = sum(int(r[ci]) for r in body); tot_s = sum(int(r[cs]) for r in body)
print("total warp-inst %d, samples %d" % (tot_i, tot_s))
# group consecutive instructions with similar executed counts (within 2x) into regions
regions = []; cur = None
for k, r in enumerate(body):
    e = int(r[ci])
    if cur is None or not (0.5 * cur["e0"] <= e <= 2.0 * cur["e0"]) :
        cur = {"start": k, "e0": max(e, 1), "rows": []}; regions.append(cur)
    cur["rows"].append(r)
minexec = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
for g in regions:
    ei = sum(int(r[ci]) for r in g["rows"]); es = sum(int(r[cs]) for r in g["rows"])
    if ei < minexec * tot_i and es < minexec * tot_s: continue
    mix = {s: sum(int(r[ix[s]] or 0) for r in g["rows"]) for s in stalls}
    top = sorted(mix.items(), key=lambda kv: -kv[1])[:5]
    print("instr %4d..%4d n=%3d exec/inst %8d  inst %5.1f%%  samples %5.1f%%  | %s" % (
        g["start"], g["start"] + len(g["rows"]) - 1, len(g["rows"]), g["e0"], 100.0 * ei / tot_i, 100.0 * es / tot_s,
        " ".join("%s %.0f%%" % (s[6:], 100.0 * v / max(es, 1)) for s, v in top)))
    print("      first: %s" % g["rows"][0][ix["Source"]].strip()[:80])
