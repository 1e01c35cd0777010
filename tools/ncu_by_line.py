#!/usr/bin/env python3
"""Per-source-line view of an ncu report for one kernel: joins `ncu --page source --csv` (SASS rows
with instruction counts and stall samples) with `nvdisasm -g` line info of the in-tree libvpz.so
(row i of the ncu page == instruction i of the function).

    python tools/ncu_by_line.py gpurun_out/prof.ncu-rep vpz_k1b_spectrumILb0 [top_n]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(func_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "vorbispizza_b200", "libvpz.so")], cwd=tmp,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = []
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_func, cur_line, inlined = None, None, None
        for ln in txt.splitlines():
            m = re.match(r"^\.text\.(\S+):", ln)
            if m:
                cur_func = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if cur_func and func_substr in cur_func and re.match(r"^\s+/\*[0-9a-f]{4}\*/", ln):
                out.append((cur_line, ln.strip()))
    return out


def main():
    rep, func = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    kname = re.sub(r"ILb\d.*", "", func)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kname],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    # ncu prints the SASS listing once per matching launch / view, each behind its own "Kernel Name" row:
    # the first block is the one joined with the disassembly
    body = rows[2:]
    for k, r in enumerate(body):
        if r and r[0] == "Kernel Name":
            body = body[:k]
            break
    data = [r for r in body if len(r) > ix["Instructions Executed"]]
    sass = sass_lines(func)
    if len(sass) != len(data):
        print("warning: %d SASS instructions vs %d ncu rows (stale .so?)" % (len(sass), len(data)))
    n = min(len(sass), len(data))
    agg = {}
    tot_i = tot_s = 0
    stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    for i in range(n):
        line = sass[i][0]
        inst = int(data[i][ix["Instructions Executed"]] or 0)
        samp = int(data[i][ix["# Samples"]] or 0)
        a = agg.setdefault(line, [0, 0, {}])
        a[0] += inst
        a[1] += samp
        for k in stall_cols:
            v = int(data[i][ix[k]] or 0)
            if v:
                a[2][k] = a[2].get(k, 0) + v
        tot_i += inst
        tot_s += samp
    src_cache = {}

    def src(line):
        if not line:
            return ""
        f, l = line
        if f not in src_cache:
            p = os.path.join(ROOT, "vorbispizza_b200", "csrc", f)
            src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        s = src_cache[f]
        return s[l - 1].strip() if 0 < l <= len(s) else ""

    print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
    print("%7s %7s  %-22s %s" % ("inst%", "samp%", "where", "top stalls | source"))
    for line, (inst, samp, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
        tops = ",".join("%s:%d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        where = "%s:%d" % line if line else "?"
        print("%6.2f%% %6.2f%%  %-22s %s | %s" % (100.0 * inst / max(tot_i, 1), 100.0 * samp / max(tot_s, 1), where, tops,
                                                  src(line)[:90]))


if __name__ == "__main__":
    main()
