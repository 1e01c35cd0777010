"""Container-level behaviour of the reader surface on DAMAGED and re-wrapped streams (tests/oggmux.py):
corrupt / truncated / empty audio packets, dropped pages (resync), CRC failures, garbage between pages,
streams without an end-of-stream flag, shortened end granules and chained logical streams.  The reference
never throws on bad AUDIO packets (StreamDecoder.cs:750-761: they yield nothing and decoding goes on);
every case checks per-call counts, positions, end-of-stream flags and PCM against the oracle's restated
reader on the same bytes.  CPU: emulated build; GPU: the product library (same cases, larger)."""
import numpy as np
import pytest

import cases
import oggmux
import oracle_binding as ob
from vorbispizza_b200 import VorbisReader


def _parts(name):
    data = cases.load_file(name)
    s = ob.OracleStream(data)
    hdr = [s.header_packet(i) for i in range(3)]
    pk = s.audio_packets()
    return hdr, pk


def _pages(hdr, pk):
    """(packets, granule) per page like oggmux.remux builds them, as a list that can be edited."""
    pages = [([hdr[0]], 0), ([hdr[1]], 0), ([hdr[2]], 0)]
    cur, cur_page, cur_gran, last = [], None, -1, 0
    for p in pk:
        if cur_page is not None and p["page_index"] != cur_page:
            g = cur_gran if cur_gran != -1 else last
            last = g
            pages.append((cur, g))
            cur, cur_gran = [], -1
        cur_page = p["page_index"]
        if sum(len(q) // 255 + 1 for q in cur) + len(p["data"]) // 255 + 1 > 255:
            g = cur_gran if cur_gran != -1 else last
            last = g
            pages.append((cur, g))
            cur, cur_gran = [], -1
        cur.append(p["data"])
        if p["granule"] != -1:
            cur_gran = p["granule"]
    if cur:
        pages.append((cur, cur_gran if cur_gran != -1 else last))
    return pages


def damaged_streams(name, limit=None):
    hdr, pk = _parts(name)
    if limit:
        pk = pk[:limit]
    n = len(pk)
    out = {}
    out["remux"] = oggmux.remux(hdr, pk)
    # audio packets the decoder has to skip or cut short
    bad = [dict(p) for p in pk]
    bad[n // 3]["data"] = bytes([bad[n // 3]["data"][0] | 1]) + bad[n // 3]["data"][1:]   # header-type bit set: not audio
    bad[n // 2]["data"] = bad[n // 2]["data"][:max(1, len(bad[n // 2]["data"]) // 3)]      # truncated
    bad[2 * n // 3]["data"] = b""                                                          # empty packet
    out["bad_packets"] = oggmux.remux(hdr, bad)
    # a mode index that does not exist (StreamDecoder.cs:732-735) when the stream has spare mode numbers
    pages = _pages(hdr, pk)
    # dropped page: the sequence numbers jump, the packet provider resynchronises
    if len(pages) > 8:
        dropped = [pg for i, pg in enumerate(pages) if i != len(pages) // 2]
        seqs = [i for i in range(len(pages)) if i != len(pages) // 2]
        blob = bytearray()
        for k, ((packets, gran), seq) in enumerate(zip(dropped, seqs)):
            segs = []
            for p in packets:
                segs += [255] * (len(p) // 255) + [len(p) % 255]
            flags = (2 if k == 0 else 0) | (4 if k == len(dropped) - 1 else 0)
            blob += oggmux.page(0x1234, seq, gran, flags, segs, b"".join(packets))
        out["dropped_page"] = bytes(blob)
    # CRC failure: one damaged byte in the body of a page in the middle
    good = bytearray(oggmux.mux(pages))
    pos = [i for i in range(len(good) - 4) if good[i:i + 4] == b"OggS"]
    if len(pos) > 8:
        victim = pos[len(pos) // 2]
        good[victim + 27 + good[victim + 26] + 3] ^= 0x5A
        out["crc_failure"] = bytes(good)
    # garbage between two pages (container waste bits)
    clean = oggmux.mux(pages)
    pos = [i for i in range(len(clean) - 4) if clean[i:i + 4] == b"OggS"]
    cut = pos[len(pos) // 2]
    out["garbage"] = clean[:cut] + bytes(range(37)) + clean[cut:]
    # no end-of-stream flag on the last page
    out["no_eos"] = oggmux.mux(pages, eos=False)
    # end granule pulled back into the last packet (end-of-stream trim, StreamDecoder.cs:658-666)
    if pages[-1][1] > 300:
        short = list(pages)
        short[-1] = (short[-1][0], short[-1][1] - 173)
        out["short_end"] = oggmux.mux(short)
    return out


CPU_CASES = [("1test", None), ("2test", 40), ("3test", 110)]


@pytest.mark.parametrize("name,limit", CPU_CASES)
def test_damaged_streams_emulated(emu_ctx, name, limit):
    for kind, data in damaged_streams(name, limit).items():
        total, _ = cases.reader_parity_bytes(emu_ctx, data, "%s/%s" % (name, kind), lookahead=16)
        assert total >= 0


def _chained(names, limit=None):
    blobs = []
    for i, n in enumerate(names):
        hdr, pk = _parts(n)
        if limit:
            pk = pk[:limit]
        blobs.append(oggmux.remux(hdr, pk, serial=0x4000 + i))
    return b"".join(blobs)


def _chained_parity(ctx, names, limit=None):
    """Chained logical streams: VorbisReader.FindNextStream / SwitchStreams (VorbisReader.cs:127-190) walk them."""
    data = _chained(names, limit)
    with VorbisReader(ctx, data, lookahead=32) as r:
        found = 1
        while r.find_next_stream():
            found += 1
        assert found == len(names) and r.stream_count == len(names)
        for i, n in enumerate(names):
            r.switch_streams(i)
            hdr, pk = _parts(n)
            if limit:
                pk = pk[:limit]
            s = ob.OracleStream(oggmux.remux(hdr, pk, serial=0x4000 + i))
            ch = r.channels
            assert ch == s.channels
            a = np.zeros(4096 * ch, np.float32)
            b = np.zeros(4096 * ch, np.float32)
            while True:
                no = s.read(a)
                ng = r.lib.vpz_reader_read(r._h, b.ctypes.data, b.size)
                assert ng == no, (n, ng, no)
                if no <= 0:
                    break
                cases.assert_pcm_close(b[:ng * ch], a[:no * ch], "chained %s" % n)


def test_chained_streams_emulated(emu_ctx):
    _chained_parity(emu_ctx, ["1test", "2test"], limit=30)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["1test", "2test", "3test", "issue6test"])
def test_damaged_streams(gpu_ctx, name):
    for kind, data in damaged_streams(name).items():
        cases.reader_parity_bytes(gpu_ctx, data, "%s/%s" % (name, kind), lookahead=64)


@pytest.mark.gpu
def test_chained_streams(gpu_ctx):
    _chained_parity(gpu_ctx, ["1test", "3test", "2test"])


# ---- K0: the physical Ogg layer on the GPU ------------------------------------------
def _reference_page_scan(data):
    """PageReaderBase.ReadNextPage / VerifyHeader / VerifyPage (Ogg/PageReaderBase.cs:41-84,176-212,286-361) in
    plain Python: byte-by-byte sync search, segment table and body inside the data, CRC over the page with a
    zeroed CRC field.  Returns (pages, waste_bits, crc_failures)."""
    import struct
    pages, pos, waste, crcf, resync, n = [], 0, 0, 0, False, len(data)
    while pos + 4 <= n:
        ok = pos + 27 <= n and data[pos:pos + 4] == b"OggS"
        if ok:
            nseg = data[pos + 26]
            ok = pos + 27 + nseg <= n
        if ok:
            lacing = data[pos + 27:pos + 27 + nseg]
            total = 27 + nseg + sum(lacing)
            ok = pos + total <= n
        if ok:
            want = struct.unpack_from("<I", data, pos + 22)[0]
            page = bytearray(data[pos:pos + total])
            page[22:26] = b"\0\0\0\0"
            if oggmux.crc32_ogg(bytes(page)) != want:
                crcf += 1
                ok = False
        if not ok:
            pos += 1
            waste += 8
            resync = True
            continue
        gran, serial, seq = struct.unpack_from("<qII", data, pos + 6)
        cont = nseg > 0 and lacing[-1] == 255
        pages.append(dict(offset=pos, body_len=total - 27 - nseg, granule=gran, serial=serial, sequence=seq,
                          flags=data[pos + 5], segments=nseg, is_resync=int(resync), is_continued=int(cont),
                          packet_count=sum(1 for v in lacing if v < 255) + (1 if cont else 0)))
        resync = False
        pos += total
    if pos < n:
        waste += 8 * (n - pos)
    return pages, waste, crcf


def _scan_parity(ctx, datas, what):
    from vorbispizza_b200 import scan_pages
    got = scan_pages(ctx, datas)
    assert len(got) == len(datas)
    for k, (data, (pages, waste, crcf)) in enumerate(zip(datas, got)):
        ref, rwaste, rcrc = _reference_page_scan(data)
        assert len(pages) == len(ref), (what, k, len(pages), len(ref))
        for i, r in enumerate(ref):
            for key, v in r.items():
                assert int(pages[i][key]) == v, (what, k, "page", i, key, int(pages[i][key]), v)
        assert (waste, crcf) == (rwaste, rcrc), (what, k, waste, crcf, rwaste, rcrc)
        # the oracle's own Ogg layer agrees on the verdicts that it exposes
        try:
            s = ob.OracleStream(data)
            assert s.crc_failures() == crcf, (what, k)
        except ob.OracleError:
            pass


def _scan_cases(names, limit):
    out = []
    for name in names:
        out.append((name, cases.load_file(name)))
        for kind, data in damaged_streams(name, limit).items():
            out.append(("%s/%s" % (name, kind), data))
    out.append(("garbage only", bytes(range(256)) * 3))
    out.append(("empty", b""))
    out.append(("three bytes", b"Ogg"))
    out.append(("truncated page", cases.load_file("1test")[:5000]))
    out.append(("capture pattern inside garbage", b"xxOggSOggS" + b"\0" * 40 + cases.load_file("1test")[:4500] + b"OggS"))
    return out


def test_gpu_page_scan_emulated(emu_ctx):
    cs = _scan_cases(["1test"], None) + _scan_cases(["3test"], 60)[1:4]
    _scan_parity(emu_ctx, [d for _, d in cs], "page scan")


def test_reader_on_device_scanned_pages_emulated(emu_lib_path):
    """The whole reader surface on top of the DEVICE page scan ("gpu_scan" 2): damaged streams included."""
    from vorbispizza_b200 import Context
    ctx = Context(0, lib_path=emu_lib_path)
    try:
        ctx.set("gpu_scan", 2)
        for kind, data in damaged_streams("1test").items():
            cases.reader_parity_bytes(ctx, data, "1test/%s (device scan)" % kind, lookahead=16)
        _chained_parity(ctx, ["1test", "2test"], limit=20)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_page_scan(gpu_ctx):
    cs = _scan_cases(["1test", "2test", "3test", "issue6test"], None)
    _scan_parity(gpu_ctx, [d for _, d in cs], "page scan")
    # many images in one call (one warp each)
    _scan_parity(gpu_ctx, [cases.load_file(n) for n in ["1test", "2test", "3test", "issue6test"]] * 40, "page scan x160")


@pytest.mark.gpu
def test_reader_on_device_scanned_pages():
    from vorbispizza_b200 import Context
    ctx = Context(0)
    try:
        ctx.set("gpu_scan", 2)
        for name in ["1test", "3test", "issue6test"]:
            for kind, data in damaged_streams(name).items():
                cases.reader_parity_bytes(ctx, data, "%s/%s (device scan)" % (name, kind), lookahead=64)
        _chained_parity(ctx, ["1test", "3test", "2test"])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_bulk_decode_same_with_host_and_device_scan(gpu_ctx):
    from vorbispizza_b200 import Context, decode_files
    datas = [cases.load_file(n) for n in ["1test", "2test", "3test", "issue6test"]]
    for kind, d in damaged_streams("3test").items():
        datas.append(d)
    a, ca = decode_files(gpu_ctx, datas, clip=True)
    host = Context(0)
    try:
        host.set("gpu_scan", 0)
        b, cb = decode_files(host, datas, clip=True)
    finally:
        host.close()
    assert (ca == cb).all() and np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ---- K0g: the seek index on the GPU -------------------------------------------------
def _granule_index_parity(ctx, named, what):
    """Device-built page-end granule index == the oracle's PacketProvider cache == the product's host walk, for
    every file that qualifies; files that do not (damaged / chained streams) must say so and keep the host walk,
    which is compared with the oracle as well."""
    from vorbispizza_b200 import VpzError
    from vorbispizza_b200 import _native as N
    from vorbispizza_b200.api import page_end_granules
    on_device = 0
    for name, data in named:
        try:
            ref = ob.OracleStream(data).page_end_granules()
        except ob.OracleError:
            continue   # no stream / a page list the reference refuses
        host = page_end_granules(ctx, data, False)
        assert np.array_equal(host, ref), (what, name, "host walk")
        try:
            dev = page_end_granules(ctx, data, True)
        except VpzError as e:
            assert e.code == N.VPZ_E_UNSUPPORTED, (what, name, e.code)
            continue
        assert np.array_equal(dev, ref), (what, name, "device index", dev[:8], ref[:8])
        on_device += 1
    return on_device


def _index_cases(names, limit):
    import synthvorbis
    out = []
    for name in names:
        out.append((name, cases.load_file(name)))
        for kind, data in damaged_streams(name, limit).items():
            out.append(("%s/%s" % (name, kind), data))
    return out


def test_granule_index_emulated(emu_ctx):
    n = _granule_index_parity(emu_ctx, _index_cases(["1test"], None) + _index_cases(["3test"], 60), "seek index")
    assert n >= 6   # the intact files and the damaged ones whose PAGES are in order


@pytest.mark.gpu
def test_granule_index(gpu_ctx):
    n = _granule_index_parity(gpu_ctx, _index_cases(["1test", "2test", "3test", "issue6test"], None), "seek index")
    assert n >= 16


# ---- random access on damaged containers: files that do not qualify for the device seek index keep the host walk,
# ---- and every excerpt still has to be what a fresh reference reader delivers (SeekTo errors included)
def _excerpts_on_damaged(ctx, name, limit, n_excerpts, nread):
    named = [(k, d) for k, d in damaged_streams(name, limit).items()]
    named.append(("intact", cases.load_file(name)))
    return cases.excerpts_parity(ctx, ["%s/%s" % (name, k) for k, _ in named], n_excerpts=n_excerpts, nread=nread,
                                 extra_positions=(0, 1, -1, -200, 10 ** 7), datas=[d for _, d in named], seed=1234)


def test_excerpts_on_damaged_streams_emulated(emu_ctx):
    _excerpts_on_damaged(emu_ctx, "1test", None, 10, 900)


@pytest.mark.gpu
def test_excerpts_on_damaged_streams(gpu_ctx):
    _excerpts_on_damaged(gpu_ctx, "3test", None, 200, 3000)
    _excerpts_on_damaged(gpu_ctx, "2test", None, 120, 4096)


# ---- a file with more pages than the device scan sizes its record area for (len / 64 + 16) falls back to the host scan
def _many_empty_pages(name, limit, every):
    hdr, pk = _parts(name)
    pages = _pages(hdr, pk[:limit] if limit else pk)
    out = []
    for i, (packets, gran) in enumerate(pages):
        out.append((packets, gran))
        if i >= 2:
            out += [([], gran)] * every      # 27-byte pages that complete no packet and repeat the granule
    return oggmux.mux(out)


def _overflow_parity(lib_path, name, limit, every):
    from vorbispizza_b200 import Context, decode_files, scan_pages, VpzError
    from vorbispizza_b200 import _native as N
    data = _many_empty_pages(name, limit, every)
    n_pages = data.count(b"OggS")
    assert n_pages > len(data) // 64 + 16, "the case must overflow the record area"
    outs = []
    for gpu_scan in (1, 0):
        ctx = Context(0, lib_path=lib_path) if lib_path else Context(0)
        try:
            ctx.set("gpu_scan", gpu_scan)
            if gpu_scan:
                with pytest.raises(VpzError) as e:      # the explicit scan call reports it ...
                    scan_pages(ctx, [data])
                assert e.value.code == N.VPZ_E_UNSUPPORTED
            outs.append(decode_files(ctx, [data, cases.load_file(name)], clip=True))   # ... the bulk path scans that file on the host
        finally:
            ctx.close()
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][0].view(np.uint32), outs[1][0].view(np.uint32))
    cases.reader_parity_bytes(Context(0, lib_path=lib_path) if lib_path else Context(0), data, name + " with empty pages", lookahead=32)


def test_record_area_overflow_emulated(emu_lib_path):
    _overflow_parity(emu_lib_path, "1test", None, 120)


@pytest.mark.gpu
def test_record_area_overflow():
    _overflow_parity(None, "2test", None, 150)
