#!/usr/bin/env python3
"""Early GPU probe: stage parity on a few packets, whole-file batch decode vs the oracle, and a
first timing of a replicated batch.  Developer tool, not part of the test-suite."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_binding as ob  # noqa: E402
from vorbispizza_b200 import _native as N  # noqa: E402


def main():
    libpath = sys.argv[1] if len(sys.argv) > 1 else N.DEFAULT_LIB
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    lib = C.CDLL(libpath)
    for name, (res, args) in N._SIGS.items():
        if hasattr(lib, name):
            f = getattr(lib, name)
            f.restype = res
            f.argtypes = args
    print(lib.vpz_version(), "devices", lib.vpz_device_count(), "nproc", os.cpu_count())
    ctx = C.c_void_p()
    rc = lib.vpz_ctx_create(0, C.byref(ctx))
    assert rc == 0, rc
    files = {}
    for fname in ["1test", "2test", "3test", "issue6test"]:
        d = open(os.path.join(ROOT, "tests/data/%s.ogg" % fname), "rb").read()
        s = ob.OracleStream(d)
        idp, sp = s.header_packet(0), s.header_packet(2)
        st = C.c_void_p()
        rc = lib.vpz_setup_create(ctx, idp, len(idp), sp, len(sp), C.byref(st))
        assert rc == 0, (rc, lib.vpz_last_error(ctx))
        pk = s.audio_packets()
        ch = s.channels
        # ---- stage parity on a sample of packets
        bad = 0
        idxs = list(range(0, len(pk), max(1, len(pk) // 24)))
        for i in idxs:
            p = pk[i]
            o = s.dump_packet(p["data"])
            dump = N.PacketDump()
            scal = np.zeros(8192, np.int32)
            cls = np.zeros(8192, np.int32)
            n1 = s.block_sizes[1]
            res = np.zeros(ch * n1 // 2, np.float32)
            spec = np.zeros(ch * n1 // 2, np.float32)
            imd = np.zeros(ch * n1, np.float32)
            buf = np.frombuffer(p["data"], np.uint8)
            rc = lib.vpz_debug_decode_packet(ctx, st, buf.ctypes.data if buf.size else None, buf.size, C.byref(dump),
                                             scal.ctypes.data, 8192, cls.ctypes.data, 8192, res.ctypes.data,
                                             spec.ctypes.data, imd.ctypes.data)
            assert rc == 0, (rc, lib.vpz_last_error(ctx))
            ok = dump.status == o["status"]
            if o["status"] == 0:
                n = o["block_size"]
                ok &= dump.scalars_n == o["scalars_n"] and np.array_equal(scal[:dump.scalars_n], o["scalars"])
                ok &= dump.classes_n == o["classes_n"] and np.array_equal(cls[:dump.classes_n], o["classes"])
                ok &= dump.bits_read == o["bits_read"]
                for c in range(ch):
                    k = o["post_count"][c]
                    ok &= dump.post_count[c] == k and list(dump.raw_posts[c]) == list(o["raw_posts"][c])
                    if k > 0:
                        ok &= list(dump.final_y[c])[:k] == list(o["final_y"][c])[:k]
                        ok &= list(dump.step_flags[c])[:k] == list(o["step_flags"][c])[:k]
                r = res[:ch * n // 2].reshape(ch, n // 2)
                sp_ = spec[:ch * n // 2].reshape(ch, n // 2)
                im = imd[:ch * n].reshape(ch, n)
                ok &= np.array_equal(r.view(np.uint32), o["residue"].view(np.uint32))
                ok &= np.array_equal(sp_.view(np.uint32), o["spectrum"].view(np.uint32))
                e = np.abs(im - o["imdct"]).max()
                ok &= bool(e <= 2e-6 * max(np.abs(o["imdct"]).max(), 1) + 1e-7)
            if not ok:
                bad += 1
                print("  packet", i, "MISMATCH")
        print(fname, "stage parity: %d packets, %d bad" % (len(idxs), bad))
        # ---- whole-file batch decode vs oracle PCM (no clip)
        s2 = ob.OracleStream(d)
        s2.set_clip(False)
        ref, _, fault = s2.decode_all()
        blob = b"".join(p["data"] for p in pk)
        offs = np.zeros(len(pk) + 1, np.uint32)
        offs[1:] = np.cumsum([len(p["data"]) for p in pk])
        bbuf = np.frombuffer(blob, np.uint8)
        bt = C.c_void_p()
        assert lib.vpz_batch_create(ctx, C.byref(bt)) == 0
        run = lib.vpz_batch_add_run(bt, st, bbuf.ctypes.data, offs.ctypes.data, len(pk), None)
        assert run == 0, (run, lib.vpz_last_error(ctx))
        ns = lib.vpz_batch_run_samples(bt, run)
        assert lib.vpz_batch_decode(bt, 0) == 0, lib.vpz_last_error(ctx)
        assert lib.vpz_batch_sync(bt) == 0, lib.vpz_last_error(ctx)
        out = np.zeros((ns, ch), np.float32)
        assert lib.vpz_batch_read_run(bt, run, out.ctypes.data) == 0
        m = min(ns, ref.shape[0])
        err = np.abs(out[:m] - ref[:m]).max()
        q = lambda x: np.clip((x * 32768.0).astype(np.int64), -32768, 32767)
        lsb = np.abs(q(out[:m]) - q(ref[:m])).max()
        print(fname, "batch decode: samples gpu %d oracle %d (fault %d), max abs err %.3g, max 16-bit diff %d, status %d"
              % (ns, ref.shape[0], fault, err, lsb, lib.vpz_batch_run_status(bt, run, None)))
        lib.vpz_batch_destroy(bt)
        files[fname] = (st, bbuf, offs, len(pk), ch, ns)
    # ---- throughput: `reps` replicas of every file in one batch
    bt = C.c_void_p()
    assert lib.vpz_batch_create(ctx, C.byref(bt)) == 0
    t0 = time.time()
    total = 0
    for r in range(reps):
        for fname, (st, bbuf, offs, n, ch, ns) in files.items():
            run = lib.vpz_batch_add_run(bt, st, bbuf.ctypes.data, offs.ctypes.data, n, None)
            assert run >= 0
            total += ns * ch
    t1 = time.time()
    assert lib.vpz_batch_upload(bt) == 0, lib.vpz_last_error(ctx)
    t2 = time.time()
    for it in range(4):
        assert lib.vpz_batch_decode(bt, 1) == 0, lib.vpz_last_error(ctx)
        assert lib.vpz_batch_sync(bt) == 0, lib.vpz_last_error(ctx)
        ln = C.c_int(0)
        ms = lib.vpz_batch_last_ms(bt, 0, C.byref(ln))
        print("iter %d: total %.3f ms (K1 %.3f, K3 %.3f), launches %d -> %.3f G channel-samples/s"
              % (it, ms, lib.vpz_batch_last_ms(bt, 1, None), lib.vpz_batch_last_ms(bt, 3, None), ln.value,
                 total / ms / 1e6))
    print("streams %d packets %d bytes %d floats %d; host add_run %.3f s, upload %.3f s"
          % (reps * 4, lib.vpz_batch_total_packets(bt), lib.vpz_batch_total_bytes(bt), total, t1 - t0, t2 - t1))
    host = np.zeros(total, np.float32)
    t3 = time.time()
    assert lib.vpz_batch_read_all(bt, host.ctypes.data) == 0
    print("D2H (pageable) %.3f s" % (time.time() - t3))
    lib.vpz_batch_destroy(bt)
    lib.vpz_ctx_destroy(ctx)


if __name__ == "__main__":
    main()
