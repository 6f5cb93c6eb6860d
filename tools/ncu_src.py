#!/usr/bin/env python3
"""ncu_src.py <rep> [kernel-regex] -> <rep>.src.txt : per-SASS-instruction stall samples of the first matching kernel."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(r for r in rows if "Address" in r)
isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
rr, seen = [], set()
for r in rows:
    if len(r) > iex and r[isamp].isdigit():
        if r[0] in seen:
            break
        seen.add(r[0])
        rr.append(r)
tot = sum(int(r[isamp]) for r in rr)
with open(rep + ".src.txt", "w") as f:
    f.write("%d instrs; samples %d executed %d\n" % (len(rr), tot, sum(int(r[iex]) for r in rr)))
    for i, r in enumerate(rr):
        f.write("%4d %7d %5.2f%% %9d  %s\n" % (i, int(r[isamp]), 100 * int(r[isamp]) / max(tot, 1), int(r[iex]), r[isrc][:120]))
print(rep + ".src.txt", len(rr), tot)
