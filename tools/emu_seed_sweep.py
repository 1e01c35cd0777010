#!/usr/bin/env python3
"""One-off wider sweep of the generated-stream parity (tests/synthvorbis.py) on the EMULATED kernels: every shape x many
seeds, integer stages / residue / spectrum bit-exact against the oracle, plus whole-stream PCM of the gather-path shapes.
Test infrastructure (uses oracle/ through tests/cases.py); the committed tests run fixed seeds, this runs more of them.
    python tools/emu_seed_sweep.py        # ~5 min; last run: 2,736 packets, 0 failures"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases, synthvorbis
from vorbispizza_b200 import Context
ctx = Context(0, lib_path=os.path.join(ROOT, 'tests', 'emu', 'libvpz_emu.so'))
t0 = time.time(); n = 0; bad = []
for seed in range(100, 112):
    for shape in synthvorbis.SHAPES:
        try:
            n += cases.synth_stage_parity(ctx, shape, seed=seed, n_packets=6)
        except Exception as ex:   # collect, keep going
            bad.append((shape, seed, repr(ex)[:200]))
for seed in range(200, 204):
    for shape in ["stereo_res2", "big_classbook", "long_codes", "sparse_ordered", "posts_beyond_block", "equal_blocks", "multi_submap_stereo"]:
        try:
            cases.synth_stream_parity(ctx, shape, seed=seed, n_packets=12, clip=True, eos_trim=50)
        except Exception as ex:
            bad.append(("stream:" + shape, seed, repr(ex)[:200]))
print("packets checked", n, "failures", len(bad), "in %.0f s" % (time.time() - t0))
for b in bad[:10]: print(b)
ctx.close()
