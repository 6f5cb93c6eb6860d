#!/usr/bin/env python3
"""ncu_regions.py <rep> [split-regex ...] -> per-region totals (instructions executed, stall samples by reason) of the
first kernel in the report.  Regions are cut at SASS lines matching BAR.SYNC / BAR.ARV (the stage boundaries)."""
import csv, subprocess, sys, re
from collections import Counter
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(r for r in rows if "Address" in r)
isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
rr, seen = [], set()
for r in rows:
    if len(r) > iex and r[isamp].isdigit():
        if r[0] in seen:
            break
        seen.add(r[0])
        rr.append(r)
regions, cur = [], {"start": 0, "ex": 0, "samp": 0, "st": Counter(), "ops": Counter(), "name": "entry"}
for i, r in enumerate(rr):
    cur["ex"] += int(r[iex]); cur["samp"] += int(r[isamp])
    op = r[isrc].split()[1] if r[isrc].startswith("@") else r[isrc].split()[0]
    cur["ops"][op.split(".")[0]] += int(r[iex])
    for j, c in stall:
        if r[j].isdigit():
            cur["st"][c[6:]] += int(r[j])
    if re.search(r"\bBAR\.|\bRET\b|\bEXIT\b", r[isrc]):
        cur["end"] = i; cur["endsrc"] = r[isrc][:50]
        regions.append(cur)
        cur = {"start": i + 1, "ex": 0, "samp": 0, "st": Counter(), "ops": Counter()}
tot = sum(x["samp"] for x in regions) or 1
for x in regions:
    if x["samp"] < 0.002 * tot: continue
    print("[%4d..%4d] ends %-40s ex %11d samples %6d (%.1f%%)" % (x["start"], x["end"], x["endsrc"], x["ex"], x["samp"], 100 * x["samp"] / tot))
    print("      stalls: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in x["st"].most_common(6)))
    print("      ops: " + ", ".join("%s %.1fM" % (k, v / 1e6) for k, v in x["ops"].most_common(10)))
