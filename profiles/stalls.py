#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: top SASS instructions by stall samples and
stall-reason totals.  usage: stalls.py file.csv [topN]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = rows[hdr_i + 1:]
S = col["# Samples"]
reasons = [h for h in hdr if h.startswith("stall_") and not h.endswith("(Not Issued)")]
tot = sum(int(r[S] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {h: sum(int(r[col[h]] or 0) for r in data) for h in reasons}
for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-28s %8d %5.1f%%" % (h, v, 100.0 * v / max(1, tot)))
print("top instructions:")
for idx, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][S] or 0))[:top]:
    rs = sorted(((int(r[col[h]] or 0), h) for h in reasons), reverse=True)[:2]
    print("  #%4d %6d %5.1f%%  %-60s %s" % (idx, int(r[S] or 0), 100.0 * int(r[S] or 0) / max(1, tot),
                                           r[col["Source"]].strip()[:60], " ".join("%s=%d" % (h[6:], v) for v, h in rs)))
