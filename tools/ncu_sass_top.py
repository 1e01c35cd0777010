#!/usr/bin/env python3
"""Hottest SASS instructions of one kernel in an ncu report (by executed count or by stall samples),
optionally filtered by a mnemonic substring.
    python tools/ncu_sass_top.py rep.ncu-rep kernel_regex [filter] [top_n] [inst|samp]"""
import csv, io, subprocess, sys
rep, k = sys.argv[1], sys.argv[2]
flt = sys.argv[3] if len(sys.argv) > 3 else ""
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
by = sys.argv[5] if len(sys.argv) > 5 else "inst"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + k], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) <= ix["Instructions Executed"]: continue
    if r[ix["Instructions Executed"]] == "Instructions Executed": break   # next launch of the report
    data.append(r)
tot = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
tots = sum(int(r[ix["# Samples"]] or 0) for r in data)
sel = [(int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0), i, r) for i, r in enumerate(data) if flt in r[ix["Source"]]]
sel.sort(key=lambda t: -(t[0] if by == "inst" else t[1]))
print("total inst %d samples %d; filter '%s': inst %d (%.1f%%) samples %d (%.1f%%)" % (tot, tots, flt, sum(s[0] for s in sel), 100.0 * sum(s[0] for s in sel) / max(tot, 1), sum(s[1] for s in sel), 100.0 * sum(s[1] for s in sel) / max(tots, 1)))
for n, sm, i, r in sel[:topn]:
    print("%5d %11d %5.2f%% samp %6d %5.2f%% thr %5s  %s" % (i, n, 100.0 * n / max(tot, 1), sm, 100.0 * sm / max(tots, 1), r[ix["Avg. Threads Executed"]], r[ix["Source"]]))
