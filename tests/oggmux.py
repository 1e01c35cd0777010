"""Tiny Ogg page writer for tests: re-wraps (possibly modified) Vorbis packets into a valid
container so that the oracle and the product decode the SAME synthetic stream (truncated packets,
bad packets, dropped pages, chained streams ...).  Test infrastructure only."""
import struct

_CRC_TABLE = []
for _i in range(256):
    _r = _i << 24
    for _ in range(8):
        _r = ((_r << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if _r & 0x80000000 else (_r << 1) & 0xFFFFFFFF
    _CRC_TABLE.append(_r)


def crc32_ogg(data, crc=0):
    for b in data:
        crc = ((crc << 8) & 0xFFFFFFFF) ^ _CRC_TABLE[((crc >> 24) ^ b) & 0xFF]
    return crc


def page(serial, seq, granule, flags, segments, body):
    hdr = b"OggS" + struct.pack("<BBqIII", 0, flags, granule, serial, seq, 0) + bytes([len(segments)]) + bytes(segments)
    crc = crc32_ogg(hdr + body)
    return hdr[:22] + struct.pack("<I", crc) + hdr[26:] + body


def mux(pages, serial=0x1234, eos=True, first_seq=0):
    """pages: list of (packets, granule) where packets is a list of bytes that all END on that page
    (no continuation across pages; every packet must need <= 255 lacing values in total per page)."""
    out = bytearray()
    for i, (packets, granule) in enumerate(pages):
        segs = []
        for p in packets:
            n = len(p)
            segs += [255] * (n // 255) + [n % 255]
        assert len(segs) <= 255, "page too large for this simple muxer"
        flags = (2 if i == 0 else 0) | (4 if eos and i == len(pages) - 1 else 0)
        out += page(serial, first_seq + i, granule, flags, segs, b"".join(packets))
    return bytes(out)


def remux(header_packets, audio_packets, serial=0x1234, eos=True):
    """header_packets: [id, comment, setup]; audio_packets: list of dicts with keys data, granule,
    page_index (as oracle_binding.OracleStream.audio_packets returns).  Packets that shared a page
    share a page again and the page keeps its granule."""
    pages = [([header_packets[0]], 0)]
    # comment + setup: one packet per page unless small enough to share
    for h in header_packets[1:]:
        if len(h) // 255 + 1 <= 255:
            pages.append(([h], 0))
        else:
            raise ValueError("header packet too large for the simple muxer")
    cur, cur_page, cur_gran = [], None, -1
    last = [0]

    def flush():
        # a page on which packets complete must carry a granule (StreamPageReader.AddPage rejects -1)
        g = cur_gran if cur_gran != -1 else last[0]
        last[0] = g
        pages.append((cur, g))

    for pk in audio_packets:
        if cur_page is not None and pk["page_index"] != cur_page:
            flush()
            cur, cur_gran = [], -1
        cur_page = pk["page_index"]
        # split when the lacing table would overflow
        segs = sum(len(p) // 255 + 1 for p in cur) + len(pk["data"]) // 255 + 1
        if segs > 255:
            flush()
            cur, cur_gran = [], -1
        cur.append(pk["data"])
        if pk["granule"] != -1:
            cur_gran = pk["granule"]
    if cur:
        flush()
    return mux(pages, serial=serial, eos=eos)
