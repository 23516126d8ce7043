#!/usr/bin/env python
"""Top stall-sample SASS lines of an ncu report: python tools/sass_hot.py rep.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(h)]
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[isamp] or 0) for r in body)
print("total samples", tot, "instructions", sum(int(r[iex] or 0) for r in body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][isamp] or 0))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[j] or 0), h[j][6:]) for j in stalls), reverse=True)[:2]
    print("%5d %5.1f%% ex=%9s  %-70s %s" % (i, 100.0 * int(r[isamp] or 0) / tot, r[iex], r[isrc][:70], top))
