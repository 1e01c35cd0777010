#!/usr/bin/env python3
"""Times the device page scan (vpz_scan_pages: stage + H2D + K0 + records) on many files and on one large file."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vorbispizza_b200 import Context, scan_pages

files = [open(os.path.join(ROOT, "tests", "data", n + ".ogg"), "rb").read() for n in ("1test", "2test", "3test", "issue6test")]
with Context(0) as ctx:
    for label, datas in (("4096 files (187 MB)", [files[i % 4] for i in range(4096)]),
                         ("256 files", [files[i % 4] for i in range(256)]),
                         ("1 file of 118 KB", [files[2]]),
                         ("1 file of 47 MB (3test x 400, chained)", [files[2] * 400])):
        scan_pages(ctx, datas)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            r = scan_pages(ctx, datas)
        dt = (time.perf_counter() - t0) / reps
        nbytes = sum(len(d) for d in datas)
        print("%-42s %8.2f ms  %7.1f MB/s  pages %d" % (label, dt * 1e3, nbytes / dt / 1e6, sum(len(p[0]) for p in r)))
