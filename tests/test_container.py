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
