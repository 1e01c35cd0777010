#!/usr/bin/env python3
"""Opcode histogram of one kernel in an ncu report, weighted by executed warp instructions.
    python tools/ncu_opmix.py rep.ncu-rep kernel_regex [launch_index]"""
import csv, io, subprocess, re, collections, sys
rep, k = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + k], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
h = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) <= ix["Instructions Executed"]: continue
    try: n = int(r[ix["Instructions Executed"]] or 0)
    except ValueError: break   # second launch of the report starts with its own header
    s = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
    op = s.split()[0] if s else "?"
    p = op.split(".")
    op = p[0] + ("." + p[1] if len(p) > 1 and p[0] in ("LDS", "STS", "LDG", "STG", "LD", "ST", "IMAD", "LDL", "STL") else "")
    h[op] += n; tot += n
print("total", tot)
for kk, v in h.most_common(45): print("%-14s %6.2f%%" % (kk, 100 * v / tot))
