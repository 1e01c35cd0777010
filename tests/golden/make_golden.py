#!/usr/bin/env python3
"""Generates tests/golden/oracle_golden.json and 1test_pcm.npy from the CPU oracle.

The reference ships no golden vectors for this path and cannot be executed here (C#, no .NET), so
these are REGRESSION pins of the oracle itself (every integer stage of all 1,306 audio packets of
the four TestFiles, PCM digests, sample totals), generated once after the oracle was pinned against
an independent decoder (make_ffmpeg_pin.py).  Run from the repo root: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402

FILES = ["1test", "2test", "3test", "issue6test"]


def stage_digest(s, packets):
    h = hashlib.sha256()
    n_scalars = n_classes = 0
    for p in packets:
        d = s.dump_packet(p["data"], want_floats=False)
        h.update(np.int32([d["status"], d["mode"], d["block_size"], d["bits_read"]] + d["info"] + d["post_count"]).tobytes())
        h.update(d["scalars"].tobytes())
        h.update(d["classes"].tobytes())
        h.update(d["raw_posts"].tobytes())
        for c, k in enumerate(d["post_count"]):
            h.update(d["final_y"][c][:k].tobytes())
            h.update(d["step_flags"][c][:k].tobytes())
        n_scalars += d["scalars_n"]
        n_classes += d["classes_n"]
    return h.hexdigest(), n_scalars, n_classes


def main():
    out = {}
    for name in FILES:
        data = open(os.path.join(ROOT, "tests", "data", name + ".ogg"), "rb").read()
        s = ob.OracleStream(data)
        pk = s.audio_packets()
        digest, nsc, ncl = stage_digest(s, pk)
        entry = dict(channels=s.channels, sample_rate=s.sample_rate, block_sizes=list(s.block_sizes),
                     audio_packets=len(pk), packet_bytes=sum(len(p["data"]) for p in pk),
                     total_samples=s.total_samples, stage_sha256=digest, decode_scalar_calls=nsc,
                     partition_classes=ncl)
        for clip in (True, False):
            t = ob.OracleStream(data)
            t.set_clip(clip)
            pcm, counts, fault = t.decode_all()
            key = "clip" if clip else "noclip"
            entry["pcm_sha256_" + key] = hashlib.sha256(pcm.tobytes()).hexdigest()
            entry["samples_" + key] = int(pcm.shape[0])
            entry["read_calls_" + key] = len(counts)
            entry["fault_" + key] = fault
            entry["has_clipped_" + key] = t.has_clipped
            if name == "1test" and clip:
                np.save(os.path.join(HERE, "1test_pcm.npy"), pcm)
        out[name] = entry
        print(name, entry)
    with open(os.path.join(HERE, "oracle_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
