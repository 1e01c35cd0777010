"""Backend-agnostic parity cases: each takes a vorbispizza_b200.Context (CUDA library on a B200 for
`-m gpu`, the emulated build for the CPU suite) and compares it with the CPU oracle on the same
inputs.  Bars (BASELINE.json north_star): integer stages bit-exact (decoded codeword indices, floor
Y values, partition classes -- plus residue and spectrum floats, which are bit-exact by
construction); float PCM within 1e-5 max-abs and <= 1 LSB after the tests' 16-bit quantisation
(NVorbis.Tests/AssetTest.cs:131-132)."""
import numpy as np

import oracle_binding as ob
from conftest import load_file
from vorbispizza_b200 import Batch, SynthBatch, VorbisReader, VpzError, decode_excerpts, decode_files
from vorbispizza_b200 import _native as N

PCM_TOL = 1e-5
IMDCT_RTOL = 2e-6


def q16(x):
    """(int)(x * 32768f) with clamp, NVorbis.Tests/AssetTest.cs:131-132"""
    return np.clip((x.astype(np.float32) * np.float32(32768.0)).astype(np.int64), -32768, 32767)


def assert_pcm_close(got, ref, what=""):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if got.size == 0:
        return
    err = float(np.abs(got - ref).max())
    lsb = int(np.abs(q16(got) - q16(ref)).max())
    assert err <= PCM_TOL, "%s: max abs err %.3g > %.1g" % (what, err, PCM_TOL)
    assert lsb <= 1, "%s: %d LSB at 16 bit" % (what, lsb)


def setup_for(ctx, ostream):
    return ctx.create_setup(ostream.header_packet(0), ostream.header_packet(2))


def compare_stage_dump(g, o, ch, what):
    assert g["status"] == o["status"], what
    if o["status"] != 0:
        return
    assert g["mode"] == o["mode"] and g["block_size"] == o["block_size"], what
    assert g["info"] == o["info"], what
    assert g["bits_read"] == o["bits_read"], (what, g["bits_read"], o["bits_read"])
    assert g["scalars_n"] == o["scalars_n"], what
    assert np.array_equal(g["scalars"], o["scalars"]), what + ": DecodeScalar sequence"
    assert g["classes_n"] == o["classes_n"], what
    assert np.array_equal(g["classes"], o["classes"]), what + ": partition classes"
    noexec = sum((1 << c) for c in range(ch) if o["no_execute"][c])
    assert g["no_execute_mask"] == noexec, what
    for c in range(ch):
        k = o["post_count"][c]
        assert g["post_count"][c] == k, what
        assert np.array_equal(g["raw_posts"][c], o["raw_posts"][c]), what + ": raw posts"
        if k > 0:
            assert np.array_equal(g["final_y"][c][:k], o["final_y"][c][:k]), what + ": final Y"
            assert np.array_equal(g["step_flags"][c][:k], o["step_flags"][c][:k]), what + ": step flags"
    assert np.array_equal(g["residue"].view(np.uint32), o["residue"].view(np.uint32)), what + ": residue bits"
    assert np.array_equal(g["spectrum"].view(np.uint32), o["spectrum"].view(np.uint32)), what + ": spectrum bits"
    scale = max(float(np.abs(o["imdct"]).max()), 1.0)
    assert float(np.abs(g["imdct"] - o["imdct"]).max()) <= IMDCT_RTOL * scale + 1e-7, what + ": imdct"


def stage_parity(ctx, name, stride=1, truncate=False):
    """Integer stages + residue/spectrum bit-exact, raw IMDCT within tolerance, packet by packet."""
    data = load_file(name)
    s = ob.OracleStream(data)
    st = setup_for(ctx, s)
    pk = s.audio_packets()
    n = 0
    try:
        for i in range(0, len(pk), stride):
            p = pk[i]["data"]
            variants = [p]
            if truncate and len(p) > 4:
                # quirk Q3: zero-padded reads / one extra symbol at the tail of a truncated packet
                variants = [p[:len(p) * 3 // 4], p[:len(p) // 2], p[:3], p[:-1]]
            for v in variants:
                o = s.dump_packet(v)
                g = ctx.debug_decode_packet(st, v, s.channels, s.block_sizes[1])
                compare_stage_dump(g, o, s.channels, "%s packet %d len %d" % (name, i, len(v)))
                n += 1
    finally:
        ctx.release_setup(st)
    return n


def batch_pcm_parity(ctx, name, clip):
    """Whole file through the batch layer (packets in, interleaved PCM out) vs the oracle's reader."""
    data = load_file(name)
    s = ob.OracleStream(data)
    s.set_clip(clip)
    ref, _, fault = s.decode_all()
    st = setup_for(ctx, s)
    pk = s.audio_packets()
    try:
        with Batch(ctx) as b:
            # end-of-stream granule trim (StreamDecoder.cs:658-666) is the packet provider side's job
            trim = np.zeros(len(pk), np.int32)
            run = b.add_run(st, [p["data"] for p in pk], trim)
            counts = b.run_packet_samples(run, len(pk))
            last = len(pk) - 1
            if pk[last]["is_eos"] and pk[last]["granule"] != -1:
                # position before the last packet = everything emitted so far
                pos_before = int(counts[:last].sum())
                info = ctx.packet_info(st, pk[last - 1]["data"])[1]
                diff = pos_before + (info[5] - info[4]) - pk[last]["granule"]
                if diff > 0:
                    trim[last] = diff
                    b.reset()
                    run = b.add_run(st, [p["data"] for p in pk], trim)
            b.decode(clip=clip)
            got = b.read_run(run)
            status, stop = b.run_status(run)
            if fault:
                assert status == N.VPZ_E_REF_FAULT, (status, fault)
            else:
                assert status == 0
            assert_pcm_close(got, ref, "%s batch clip=%s" % (name, clip))
            assert b.has_clipped == s.has_clipped
    finally:
        ctx.release_setup(st)
    return got.shape


def reader_parity(ctx, name, clip=True, lookahead=None, chunk=48000):
    """ReadSamples loop: same per-call counts, positions, flags and PCM as the oracle's restated reader."""
    data = load_file(name)
    return reader_parity_bytes(ctx, data, name, clip, lookahead, chunk)


def reader_parity_bytes(ctx, data, name, clip=True, lookahead=None, chunk=48000):
    s = ob.OracleStream(data)
    s.set_clip(clip)
    with VorbisReader(ctx, data, lookahead=lookahead) as r:
        r.clip_samples = clip
        assert r.channels == s.channels and r.sample_rate == s.sample_rate
        assert (r.upper_bitrate, r.nominal_bitrate, r.lower_bitrate) == s.bitrates()
        assert r.vendor == s.vendor() and r.comments == s.comments()
        ch = r.channels
        chunk -= chunk % ch
        a = np.zeros(chunk, np.float32)
        b = np.zeros(chunk, np.float32)
        total = 0
        calls = 0
        while True:
            no = s.read(a)
            ng = r.lib.vpz_reader_read(r._h, b.ctypes.data, b.size)
            if no < 0:
                assert ng == N.VPZ_E_REF_FAULT and no == -6, (name, ng, no)
                break
            assert ng == no, "%s call %d: product %d oracle %d" % (name, calls, ng, no)
            assert r.sample_position == s.sample_position, name
            assert r.is_end_of_stream == s.is_end_of_stream, name
            if no == 0:
                break
            assert_pcm_close(b[:ng * ch], a[:no * ch], "%s call %d" % (name, calls))
            assert r.has_clipped == s.has_clipped, name
            total += no
            calls += 1
        ot = s.total_samples   # a negative value is the error the reference would throw (damaged granules)
        try:
            gt = r.total_samples
        except VpzError as e:
            gt = {N.VPZ_E_INVALID_DATA: -1, N.VPZ_E_ARGUMENT: -2, N.VPZ_E_SEEK_RANGE: -3}.get(e.code, e.code)
        assert gt == ot, (name, gt, ot)
    return total, calls


def reader_planar_and_partial(ctx, name):
    """Planar reads with a stride and small odd-sized requests (partial packets)."""
    data = load_file(name)
    s = ob.OracleStream(data)
    with VorbisReader(ctx, data) as r:
        ch = r.channels
        stride = 300
        a = np.zeros(ch * stride, np.float32)
        b = np.zeros(ch * stride, np.float32)
        sizes = [1, 7, 100, 257, 300, 64]
        for k in range(40):
            want = sizes[k % len(sizes)]
            no = s.read_planar(a, want, stride)
            ng = r.read_samples_planar(b, want, stride)
            assert ng == no, (name, k, ng, no)
            if no == 0:
                break
            for c in range(ch):
                assert_pcm_close(b[c * stride:c * stride + ng], a[c * stride:c * stride + no], "%s planar %d" % (name, k))
            assert r.sample_position == s.sample_position
        # argument errors (StreamDecoder.cs:423-430)
        try:
            r.ctx.check(r.lib.vpz_reader_read(r._h, b.ctypes.data, ch * 4 + (1 if ch > 1 else 0)))
            bad = ch == 1
        except VpzError as e:
            bad = e.code == N.VPZ_E_ARGUMENT
        assert bad
        try:
            r.read_samples_planar(b[:ch * 4], 5, 4)
            assert False, "buffer too small must be rejected"
        except VpzError as e:
            assert e.code == N.VPZ_E_ARGUMENT


def seek_parity(ctx, name, positions, nread=4096, lookahead=64):
    """SeekTo + read (pre-roll, roll-forward, position bookkeeping) vs the oracle."""
    data = load_file(name)
    s = ob.OracleStream(data)
    with VorbisReader(ctx, data, lookahead=lookahead) as r:
        ch = r.channels
        a = np.zeros(8192 * ch, np.float32)
        b = np.zeros(8192 * ch, np.float32)
        for pos in positions:
            try:
                s.seek(pos)
                oerr = 0
            except ob.OracleError as e:
                oerr = e.code
            try:
                r.seek_to(pos)
                gerr = 0
            except VpzError as e:
                gerr = e.code
            omap = {0: 0, -1: N.VPZ_E_INVALID_DATA, -2: N.VPZ_E_ARGUMENT, -3: N.VPZ_E_SEEK_RANGE, -4: N.VPZ_E_PREROLL,
                    -6: N.VPZ_E_REF_FAULT}
            assert gerr == omap[oerr], "%s seek %d: product %d oracle %d" % (name, pos, gerr, oerr)
            if oerr:
                continue
            assert r.sample_position == s.sample_position == pos
            got = 0
            while got < nread:
                no = s.read(a)
                ng = r.lib.vpz_reader_read(r._h, b.ctypes.data, b.size)
                if no < 0:
                    assert ng == N.VPZ_E_REF_FAULT
                    break
                assert ng == no, "%s seek %d: product %d oracle %d" % (name, pos, ng, no)
                if no == 0:
                    break
                assert_pcm_close(b[:ng * ch], a[:no * ch], "%s after seek %d" % (name, pos))
                assert r.sample_position == s.sample_position
                got += no


def decode_files_s16_parity(ctx, names, clip=True):
    """16-bit bulk output (vpz_decode_files_s16) vs the oracle's float PCM converted by the reference tests'
    rule v = (int)(x * 32768f) clamped (AssetTest.cs:131-132): at most 1 LSB apart (the float PCM itself agrees
    to ~1e-6, which can move a sample across a truncation boundary), and identical to the product's own float
    output converted by the same rule."""
    datas = [load_file(n) for n in names]
    pcm16, counts = decode_files(ctx, datas, clip=clip, s16=True)
    pcmf, countsf = decode_files(ctx, datas, clip=clip)
    assert pcm16.dtype == np.int16 and (counts == countsf).all()

    def to_s16(x):
        v = np.trunc(x.astype(np.float32) * np.float32(32768.0)).astype(np.int64)
        return np.clip(v, -32768, 32767).astype(np.int16)

    assert np.array_equal(pcm16, to_s16(pcmf)), "GPU conversion differs from the rule applied to the GPU float output"
    off = 0
    for n, d, cnt in zip(names, datas, counts):
        s = ob.OracleStream(d)
        s.set_clip(clip)
        ref, _, _ = s.decode_all()
        assert cnt == ref.shape[0]
        got = pcm16[off:off + cnt * s.channels].reshape(-1, s.channels)
        off += cnt * s.channels
        diff = np.abs(got.astype(np.int32) - to_s16(ref).astype(np.int32))
        assert diff.max() <= 1, "%s: max 16-bit difference %d" % (n, diff.max())
    assert off == pcm16.size


def excerpts_parity(ctx, names, n_excerpts, nread=4096, seed=0x5EED0005, extra_positions=(), clip=True, datas=None):
    """BASELINE config 5: a batch of short excerpts (SeekTo + read nread samples, each like a fresh reader)
    through vpz_decode_excerpts vs the oracle's reader, excerpt by excerpt.  Start samples: seeded uniform
    in [0, total) of a uniformly chosen file, plus `extra_positions` on every file (edge cases).
    datas: container images to use instead of the TestFiles `names` (damaged streams: names label them)."""
    if datas is None:
        datas = [load_file(n) for n in names]
    totals, chans = [], []
    for d in datas:
        s = ob.OracleStream(d)
        try:
            ref, _, _ = s.decode_all()
            totals.append(max(ref.shape[0], 2))
        except ob.OracleError:
            totals.append(100000)
        chans.append(s.channels)
    rng = np.random.default_rng(seed)
    file_of = list(rng.integers(0, len(names), n_excerpts))
    start = [int(rng.integers(0, max(1, totals[f] - 1))) for f in file_of]
    for f in range(len(names)):
        for p in extra_positions:
            file_of.append(f)
            start.append(int(p) if p >= 0 else totals[f] + int(p))
    file_of = np.array(file_of, np.uint32)
    start = np.array(start, np.int64)
    count = np.full(file_of.size, nread, np.int32)
    pcm, offsets, got = decode_excerpts(ctx, datas, file_of, start, count, clip=clip)
    omap = {0: 0, -1: N.VPZ_E_INVALID_DATA, -2: N.VPZ_E_ARGUMENT, -3: N.VPZ_E_SEEK_RANGE, -4: N.VPZ_E_PREROLL,
            -6: N.VPZ_E_REF_FAULT}
    buf = np.zeros(8192 * 2, np.float32)
    checked = 0
    for i in range(file_of.size):
        f = int(file_of[i])
        ch = chans[f]
        s = ob.OracleStream(datas[f])   # a fresh reader per excerpt
        s.set_clip(clip)
        try:
            s.seek(int(start[i]))
        except ob.OracleError as e:
            assert got[i] == omap[e.code], "excerpt %d (%s @ %d): product %d oracle error %d" % (i, names[f], start[i], got[i], e.code)
            assert not pcm[offsets[i]:offsets[i] + nread * ch].any(), "a failed excerpt reads as zeros"
            continue
        ref = []
        n = 0
        while n < nread:
            want = min(nread - n, 8192)
            no = s.read(buf[:want * ch])
            if no <= 0:
                break
            ref.append(buf[:no * ch].copy())
            n += no
        assert got[i] == n, "excerpt %d (%s @ %d): product %d samples, oracle %d" % (i, names[f], start[i], got[i], n)
        assert not pcm[offsets[i] + n * ch:offsets[i] + nread * ch].any(), "samples past the end of the stream read as zeros"
        if n:
            a = pcm[offsets[i]:offsets[i] + n * ch]
            assert_pcm_close(a, np.concatenate(ref), "excerpt %d (%s @ %d)" % (i, names[f], start[i]))
        checked += 1
    return checked


def reference_imdct_ola(flags, spectra, channels, size0, size1):
    """Oracle-side restatement of a synthetic stream: Mdct.Reverse per block (oracle C), then the
    window / overlap-add / valid-range rules of StreamDecoder.ReadNextPacket in numpy fp32."""
    w = {size0: ob.window_slope(size0 // 2), size1: ob.window_slope(size1 // 2)}
    out = [[] for _ in range(channels)]
    prev = None
    off = 0
    n_blocks = len(flags)
    for i in range(n_blocks):
        lb = bool(flags[i] & 1)
        pf = True if i == 0 else bool(flags[i - 1] & 1)
        nf = True if i + 1 == n_blocks else bool(flags[i + 1] & 1)
        n = size1 if lb else size0
        info = ob.packet_info(size0, size1, lb, pf, nf)
        ls, rs, re = info[2], info[4], info[5]
        cur = []
        for c in range(channels):
            y = ob.imdct(spectra[off:off + n // 2])
            off += n // 2
            cur.append(y)
        if prev is not None:
            pcur, prs, pre = prev
            L = pre - prs
            slope = w[size1 if info[1] else size0]
            for c in range(channels):
                y = cur[c]
                a = (y[ls:ls + L] * slope[:L]).astype(np.float32)
                b = (pcur[c][prs:prs + L] * slope[:L][::-1]).astype(np.float32)
                y[ls:ls + L] = (a + b).astype(np.float32)
                out[c].append(y[ls:rs].copy())
        prev = (cur, rs, re)
    return np.stack([np.concatenate(o) for o in out], axis=1)


def synth_flags(rng, n_streams, n_blocks):
    """Markov block-size sequence of SURVEY 8(d) config 3: P(long->short)=1/16, P(short->long)=1/4."""
    f = np.zeros((n_streams, n_blocks), np.uint8)
    for s in range(n_streams):
        cur = 1
        for i in range(n_blocks):
            f[s, i] = cur
            u = rng.random()
            cur = (0 if u < 1 / 16 else 1) if cur else (1 if u < 1 / 4 else 0)
    return f


def synth_spectra(rng, flags, channels, size0, size1):
    parts = []
    for s in range(flags.shape[0]):
        for i in range(flags.shape[1]):
            m = (size1 if flags[s, i] & 1 else size0) // 2
            k = np.arange(m, dtype=np.float32)
            for c in range(channels):
                parts.append((rng.standard_normal(m).astype(np.float32) * np.exp2(-k / 128.0) * 0.02).astype(np.float32))
    return np.concatenate(parts)


def synth_parity(ctx, channels=2, n_streams=3, n_blocks=24, lg0=8, lg1=11, seed=0x5EED0001, clip=False):
    rng = np.random.default_rng(seed)
    flags = synth_flags(rng, n_streams, n_blocks)
    spectra = synth_spectra(rng, flags, channels, 1 << lg0, 1 << lg1)
    with SynthBatch(ctx, channels, lg0, lg1, flags, spectra) as b:
        b.decode(clip=clip)
        off = 0
        for s in range(n_streams):
            got = b.read_run(s)
            nfl = sum(((1 << lg1) if f & 1 else (1 << lg0)) // 2 * channels for f in flags[s])
            ref = reference_imdct_ola(flags[s], spectra[off:off + nfl], channels, 1 << lg0, 1 << lg1)
            off += nfl
            if clip:
                ref = np.clip(ref, -0.99999994, 0.99999994).astype(np.float32)
            assert_pcm_close(got, ref, "synth stream %d" % s)
    return True


def decode_files_parity(ctx, names, clip=True):
    datas = [load_file(n) for n in names]
    pcm, counts = decode_files(ctx, datas, clip=clip)
    off = 0
    for n, d, cnt in zip(names, datas, counts):
        s = ob.OracleStream(d)
        s.set_clip(clip)
        ref, _, _ = s.decode_all()
        assert cnt == ref.shape[0], (n, cnt, ref.shape)
        got = pcm[off:off + cnt * s.channels].reshape(-1, s.channels)
        off += cnt * s.channels
        assert_pcm_close(got, ref, "decode_files " + n)
    assert off == pcm.size


# ---- generated streams (tests/synthvorbis.py): every setup shape the TestFiles do not have ---------------
FLOOR0_RTOL = 2e-5   # floor 0 evaluates cos / sqrt / exp in fp32: library functions, not bit-reproducible


def compare_stage_dump_tol(g, o, ch, what, rtol):
    """compare_stage_dump with a relative tolerance on the spectrum (floor 0): integer stages and the
    residue stay bit-exact."""
    keep_g, keep_o = g["spectrum"], o["spectrum"]
    fin = np.isfinite(keep_o)
    assert np.array_equal(fin, np.isfinite(keep_g)), what + ": non-finite spectrum bins differ"
    scale = np.maximum(np.abs(keep_o[fin]), 1e-30)
    assert float((np.abs(keep_g[fin] - keep_o[fin]) / scale).max(initial=0.0)) <= rtol, what + ": spectrum (floor 0)"
    g = dict(g)
    o = dict(o)
    g["spectrum"] = o["spectrum"] = np.zeros(1, np.float32)
    ok = np.isfinite(o["imdct"]).all() and np.isfinite(g["imdct"]).all()
    if ok:
        scale = max(float(np.abs(o["imdct"]).max()), 1.0)
        assert float(np.abs(g["imdct"] - o["imdct"]).max()) <= (IMDCT_RTOL + 40 * rtol) * scale + 1e-7, what + ": imdct"
    g["imdct"] = o["imdct"] = np.zeros(1, np.float32)
    compare_stage_dump(g, o, ch, what)


def synth_stage_parity(ctx, shape, seed, n_packets=40, mean_len=120):
    """Integer stages + residue (+ spectrum) of every generated packet, plus truncated variants."""
    import synthvorbis as sv
    st = sv.make_stream(seed, shape, n_packets=n_packets, mean_len=mean_len)
    s = ob.OracleStream(st["ogg"])
    setup = setup_for(ctx, s)
    floor0 = any(f["type"] == 0 for f in st["info"]["floors"])
    n = 0
    try:
        for i, p in enumerate(st["packets"]):
            for v in (p, p[:max(len(p) * 2 // 3, 1)]):
                o = s.dump_packet(v)
                g = ctx.debug_decode_packet(setup, v, s.channels, s.block_sizes[1])
                what = "%s seed %d packet %d len %d" % (shape, seed, i, len(v))
                if floor0:
                    compare_stage_dump_tol(g, o, s.channels, what, FLOOR0_RTOL)
                else:
                    compare_stage_dump(g, o, s.channels, what)
                n += 1
    finally:
        ctx.release_setup(setup)
    return n


def assert_pcm_close_scaled(got, ref, what="", rtol=PCM_TOL, peak=None):
    """PCM of generated streams is not normalised to +-1: the 1e-5 bar applies relative to the peak (of the
    UNCLIPPED signal: clipping hides the peak but not the rounding error of the samples below it)."""
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if got.size == 0:
        return
    scale = max(float(np.abs(ref).max()), 1.0) if peak is None else max(peak, 1.0)
    err = float(np.abs(got - ref).max())
    assert err <= rtol * scale, "%s: max abs err %.3g > %.1g x peak %.3g" % (what, err, rtol, scale)


def synth_stream_parity(ctx, shape, seed, n_packets=40, clip=False, lookahead=7, eos_trim=0):
    """A generated stream end to end: bulk decode (vpz_decode_files) and the reader's ReadSamples loop vs the
    oracle's reader -- same counts, positions, PCM."""
    import synthvorbis as sv
    st = sv.make_stream(seed, shape, n_packets=n_packets, eos_trim=eos_trim)
    data = st["ogg"]
    s = ob.OracleStream(data)
    s.set_clip(clip)
    ref, _, fault = s.decode_all()
    assert not fault
    assert np.isfinite(ref).all(), "generated stream overflows in the oracle: not a usable fixture"
    su = ob.OracleStream(data)
    su.set_clip(False)
    peak = float(np.abs(su.decode_all()[0]).max(initial=0.0))
    floor0 = any(f["type"] == 0 for f in st["info"]["floors"])
    rtol = PCM_TOL + (40 * FLOOR0_RTOL if floor0 else 0.0)
    pcm, counts = decode_files(ctx, [data], clip=clip)
    assert counts[0] == ref.shape[0], (shape, seed, counts[0], ref.shape)
    assert_pcm_close_scaled(pcm.reshape(-1, s.channels), ref, "%s seed %d bulk" % (shape, seed), rtol, peak)
    s2 = ob.OracleStream(data)
    s2.set_clip(clip)
    with VorbisReader(ctx, data, lookahead=lookahead) as r:
        r.clip_samples = clip
        ch = r.channels
        a = np.zeros(4096 * ch, np.float32)
        b = np.zeros(4096 * ch, np.float32)
        total = 0
        while True:
            no = s2.read(a)
            ng = r.lib.vpz_reader_read(r._h, b.ctypes.data, b.size)
            assert ng == no, "%s seed %d: product %d oracle %d" % (shape, seed, ng, no)
            assert r.sample_position == s2.sample_position and r.is_end_of_stream == s2.is_end_of_stream
            if no <= 0:
                break
            assert_pcm_close_scaled(b[:ng * ch], a[:no * ch], "%s seed %d read" % (shape, seed), rtol, peak)
            total += no
        assert total == ref.shape[0]
    return total


def synth_mixed_batch_parity(ctx, shape_seeds, n_packets=20, with_files=(), clip=False):
    """One vpz_decode_files call over generated streams of different kernel classes plus TestFiles: the
    per-setup path selection must give every stream the result it has when decoded alone."""
    import synthvorbis as sv
    datas, names = [], []
    for shape, seed in shape_seeds:
        datas.append(sv.make_stream(seed, shape, n_packets=n_packets)["ogg"])
        names.append("%s/%d" % (shape, seed))
    for n in with_files:
        datas.append(load_file(n))
        names.append(n)
    pcm, counts = decode_files(ctx, datas, clip=clip)
    off = 0
    for name, d, cnt in zip(names, datas, counts):
        s = ob.OracleStream(d)
        s.set_clip(clip)
        ref, _, _ = s.decode_all()
        su = ob.OracleStream(d)
        su.set_clip(False)
        peak = float(np.abs(su.decode_all()[0]).max(initial=0.0))
        assert cnt == ref.shape[0], (name, cnt, ref.shape)
        got = pcm[off:off + cnt * s.channels].reshape(-1, s.channels)
        off += cnt * s.channels
        rtol = PCM_TOL + (40 * FLOOR0_RTOL if name.startswith("floor0") else 0.0)
        assert_pcm_close_scaled(got, ref, "mixed batch " + name, rtol, peak)
    assert off == pcm.size
