"""Generator of VALID Vorbis I streams with arbitrary setups (test infrastructure only).

The four TestFiles are mono / stereo, floor 1, residue 1 / 2, one submap, power-of-two VQ dimensions,
block sizes 256 / 2048.  Everything else the reference's decode path accepts -- residue type 0, up to 8
channels with several coupling steps, several submaps, floor 0, VQ dimensions that do not divide the
partition, codewords up to 32 bits, ordered / sparse codebooks, classbooks with more entries than
classwords, floor posts at and beyond the end of the block, 64-post floors, other block sizes -- is
produced here: a random setup header of the requested SHAPE plus audio packets whose payload is random
bits.  Huffman-complete codebooks decode any bit string, so random payloads walk every decode path and
run into the end of the packet at random places (the reference's zero-padded end-of-packet rules).

The same bytes go to the CPU oracle (checker) and to the product; nothing here decodes anything.
Header layouts follow the reference's readers: Codebook.cs:21-144,220-288, Floor0.cs:39-76,
Floor1.cs:39-155, Residue0.cs:25-115, Mapping.cs:19-95, Mode.cs:14-28, StreamDecoder.cs:213-321.
"""
import numpy as np

import oggmux


class BitWriter:
    """LSB-first bit packer (the order VorbisPacket.ReadBits consumes, VorbisPacket.cs:157-186)."""

    def __init__(self):
        self.acc = 0
        self.n = 0

    def write(self, value, bits):
        assert 0 <= value < (1 << bits) or bits == 0, (value, bits)
        self.acc |= int(value) << self.n
        self.n += bits

    def bytes(self):
        nbytes = (self.n + 7) // 8
        return self.acc.to_bytes(nbytes, "little")


def ilog(x):
    return int(x).bit_length() if x > 0 else 0


def lookup1_values(entries, dims):
    """Codebook.lookup1_values (Codebook.cs:290-298)."""
    r = int(np.floor(np.exp(np.log(entries) / dims)))
    if np.floor(float(r + 1) ** dims) <= entries:
        r += 1
    return r


def pack_float32(mantissa, exponent, negative=False):
    """Utils.ConvertFromVorbisFloat32's inverse: value = mantissa * 2^exponent (mantissa < 2^21)."""
    assert 0 <= mantissa < (1 << 21) and 0 <= exponent + 788 < 1024
    return (0x80000000 if negative else 0) | ((exponent + 788) << 21) | mantissa


def random_lengths(rng, n_used, max_len, deep=False):
    """Codeword lengths of a COMPLETE prefix code with n_used leaves, none longer than max_len: split random
    leaves of a binary tree.  deep: keep splitting the deepest leaf first (one very long code)."""
    if n_used == 1:
        return [1]   # Huffman.cs:52-58: a single used entry must have length 1
    leaves = [1, 1]
    if deep:
        while len(leaves) < n_used and max(leaves) < max_len:
            d = max(leaves)
            leaves.remove(d)
            leaves += [d + 1, d + 1]
    while len(leaves) < n_used:
        cand = [i for i, d in enumerate(leaves) if d < max_len]
        assert cand, "max_len too small for the entry count"
        i = cand[int(rng.integers(0, len(cand)))]
        d = leaves.pop(i)
        leaves += [d + 1, d + 1]
    return leaves


def make_book(rng, entries, dims, max_len=12, mode="dense", lookup=0, deep=False, complete=True,
              value_bits=None, scale_exp=-6):
    """One codebook.  mode: dense | sparse | ordered.  lookup: 0 none, 1 lattice, 2 explicit."""
    if mode == "sparse":
        n_used = max(1, int(entries * float(rng.uniform(0.1, 0.6))))
    else:
        n_used = entries
    need = max(1, int(np.ceil(np.log2(max(n_used, 2)))))
    max_len = max(max_len, need)
    lens = random_lengths(rng, n_used, max_len, deep=deep)
    if not complete and n_used > 2:
        # under-subscribed: one code is made longer, so some bit patterns match nothing (DecodeScalar -> -1)
        k = int(np.argmin(lens))
        if lens[k] < 32:
            lens[k] += 1
    lengths = [-1] * entries
    if mode == "ordered":
        lens.sort()
        for i in range(entries):
            lengths[i] = lens[i]
    else:
        rng.shuffle(lens)
        used = sorted(rng.choice(entries, n_used, replace=False).tolist()) if mode == "sparse" else range(entries)
        for i, e in enumerate(used):
            lengths[e] = lens[i]
    bk = dict(dims=dims, entries=entries, lengths=lengths, mode=mode, lookup=lookup)
    if lookup:
        vb = value_bits if value_bits else int(rng.integers(1, 9))
        bk["value_bits"] = vb
        bk["sequence_p"] = int(rng.integers(0, 4) == 0)
        nvals = lookup1_values(entries, dims) if lookup == 1 else entries * dims
        bk["mults"] = [int(x) for x in rng.integers(0, 1 << vb, nvals)]
        # delta = m * 2^scale_exp, min ~ -half the range: values of order 2^(scale_exp + value_bits)
        m = int(rng.integers(1, 8))
        bk["delta"] = pack_float32(m, scale_exp)
        bk["min"] = pack_float32(m * (1 << vb) // 2, scale_exp, negative=True)
    return bk


def write_book(bw, bk):
    bw.write(0x564342, 24)
    bw.write(bk["dims"], 16)
    bw.write(bk["entries"], 24)
    lengths = bk["lengths"]
    if bk["mode"] == "ordered":
        bw.write(1, 1)
        cur = lengths[0]
        bw.write(cur - 1, 5)
        i = 0
        n = bk["entries"]
        while i < n:
            cnt = 0
            while i + cnt < n and lengths[i + cnt] == cur:
                cnt += 1
            bw.write(cnt, ilog(n - i))
            i += cnt
            cur += 1
    else:
        bw.write(0, 1)
        sparse = bk["mode"] == "sparse"
        bw.write(1 if sparse else 0, 1)
        for l in lengths:
            if sparse:
                bw.write(1 if l > 0 else 0, 1)
                if l <= 0:
                    continue
            bw.write(l - 1, 5)
    bw.write(bk["lookup"], 4)
    if bk["lookup"]:
        bw.write(bk["min"], 32)
        bw.write(bk["delta"], 32)
        bw.write(bk["value_bits"] - 1, 4)
        bw.write(bk["sequence_p"], 1)
        for v in bk["mults"]:
            bw.write(v, bk["value_bits"])


def write_floor1(bw, fl):
    bw.write(1, 16)
    bw.write(len(fl["part_class"]), 5)
    for c in fl["part_class"]:
        bw.write(c, 4)
    for cl in fl["classes"]:
        bw.write(cl["dim"] - 1, 3)
        bw.write(cl["sub_bits"], 2)
        if cl["sub_bits"]:
            bw.write(cl["master"], 8)
        for b in cl["books"]:
            bw.write(b + 1, 8)
    bw.write(fl["multiplier"] - 1, 2)
    bw.write(fl["range_bits"], 4)
    for x in fl["xs"]:
        bw.write(x, fl["range_bits"])


def write_floor0(bw, fl):
    bw.write(0, 16)
    bw.write(fl["order"], 8)
    bw.write(fl["rate"], 16)
    bw.write(fl["bark_map_size"], 16)
    bw.write(fl["amp_bits"], 6)
    bw.write(fl["amp_ofs"], 8)
    bw.write(len(fl["books"]) - 1, 4)
    for b in fl["books"]:
        bw.write(b, 8)


def write_residue(bw, rs):
    bw.write(rs["type"], 16)
    bw.write(rs["begin"], 24)
    bw.write(rs["end"], 24)
    bw.write(rs["part_size"] - 1, 24)
    bw.write(len(rs["cascade"]) - 1, 6)
    bw.write(rs["class_book"], 8)
    for c in rs["cascade"]:
        bw.write(c & 7, 3)
        if c >> 3:
            bw.write(1, 1)
            bw.write(c >> 3, 5)
        else:
            bw.write(0, 1)
    for b in rs["books"]:
        bw.write(b, 8)


def write_mapping(bw, mp, channels):
    bw.write(0, 16)
    if mp["submaps"] > 1:
        bw.write(1, 1)
        bw.write(mp["submaps"] - 1, 4)
    else:
        bw.write(0, 1)
    if mp["coupling"]:
        bw.write(1, 1)
        bw.write(len(mp["coupling"]) - 1, 8)
        for mag, ang in mp["coupling"]:
            bw.write(mag, ilog(channels - 1))
            bw.write(ang, ilog(channels - 1))
    else:
        bw.write(0, 1)
    bw.write(0, 2)
    if mp["submaps"] > 1:
        for m in mp["mux"]:
            bw.write(m, 4)
    for fl, rs in mp["sub"]:
        bw.write(0, 8)
        bw.write(fl, 8)
        bw.write(rs, 8)


SHAPES = {
    # name: keyword arguments of make_setup
    "stereo_res2": dict(channels=2, res_types=(2,), coupling=1),
    "mono_res1_dims3": dict(channels=1, res_types=(1,), vq_dims=(3, 5, 6), psize=30),
    "res0_4ch": dict(channels=4, res_types=(0,), coupling=2, vq_dims=(1, 2, 4)),
    "res012_3ch": dict(channels=3, res_types=(0, 1, 2), coupling=2, lg=(8, 10)),
    "ch6_coupled": dict(channels=6, res_types=(1, 2), coupling=5, lg=(9, 11)),
    "ch8": dict(channels=8, res_types=(2,), coupling=7, lg=(8, 9), psize=16),
    "long_codes": dict(channels=2, res_types=(1,), deep_books=True, coupling=1),
    "sparse_ordered": dict(channels=2, res_types=(2, 1), book_modes=("sparse", "ordered"), coupling=1),
    "incomplete_books": dict(channels=2, res_types=(1,), incomplete=True, coupling=1),
    "big_classbook": dict(channels=2, res_types=(1, 2), classbook_extra=7, coupling=1),
    "posts_beyond_block": dict(channels=2, res_types=(2,), range_bits=11, floor_posts=40, coupling=1),
    "posts64": dict(channels=1, res_types=(1,), floor_posts=64),
    "multi_submap": dict(channels=4, res_types=(1, 2, 0), submaps=3, coupling=2),
    "multi_submap_stereo": dict(channels=2, res_types=(1, 1), submaps=2, coupling=1),
    "floor0": dict(channels=2, res_types=(1,), floor0=True, coupling=1),
    "floor0_mixed": dict(channels=3, res_types=(2, 1), floor0=True, floor1_too=True, submaps=2, coupling=1),
    "equal_blocks": dict(channels=2, res_types=(2,), lg=(10, 10), coupling=1),
    "blocks_512_4096": dict(channels=1, res_types=(1,), lg=(9, 12)),
    # partitions of 2,048 positions: above the 2,047 entries per unit that K1b's gather path packs into 16-bit prefix
    # sums (engine.cpp), so the setup must take the general path
    "big_partitions": dict(channels=2, res_types=(1,), lg=(9, 13), psize=2048, vq_dims=(1, 2, 4, 8), coupling=1),
}


def make_setup(rng, channels=2, res_types=(1,), coupling=0, lg=(8, 11), vq_dims=(1, 2, 4, 8), psize=None,
               deep_books=False, book_modes=("dense",), incomplete=False, classbook_extra=0, range_bits=None,
               floor_posts=None, submaps=1, floor0=False, floor1_too=False, rate=44100):
    """Returns (id_packet, setup_packet, info dict)."""
    size0, size1 = 1 << lg[0], 1 << lg[1]
    books = []

    def add_book(**kw):
        mode = book_modes[int(rng.integers(0, len(book_modes)))]
        kw.setdefault("mode", mode)
        if incomplete and kw.get("lookup", 0) and int(rng.integers(0, 3)) == 0:
            kw["complete"] = False
        books.append(make_book(rng, **kw))
        return len(books) - 1

    # ---- floor books / floors ------------------------------------------------------------------
    floors = []
    n_floors = 2 if (floor0 and floor1_too) or submaps > 1 else 1
    for fi in range(n_floors):
        use0 = floor0 and not (floor1_too and fi == 1)
        if use0:
            # LSP coefficients as an encoder makes them: increasing angles in (0, pi).  Every VQ component is a
            # positive gap (min > 0, delta > 0, sequence_p accumulates inside a vector, Floor0.Unpack's
            # "averaging" across vectors), so the roots of the two polynomials interlace and p + q
            # (Floor0.cs:196-214) stays away from zero: the curve is finite for random payloads.
            order = int(rng.integers(1, 9)) if fi == 0 else 5
            gap = np.pi / (1.7 * order)
            fb = []
            for _ in range(int(rng.integers(1, 4))):
                bi = add_book(entries=int(rng.integers(8, 64)), dims=int(rng.integers(1, 5)), lookup=int(rng.integers(1, 3)),
                              max_len=10, value_bits=4)
                bk = books[bi]
                bk["sequence_p"] = 1
                bk["delta"] = pack_float32(int(round(gap / 15.0 * (1 << 24))), -24)
                bk["min"] = pack_float32(int(round(gap * 0.5 * (1 << 20))), -20)
                fb.append(bi)
            floors.append(dict(type=0, order=order, rate=rate, bark_map_size=int(rng.integers(8, size0 // 2 + 1)),
                               # a small amplitude offset keeps Amp / sqrt(p + q) - ampOfs (Floor0.cs:217) inside exp's range
                               amp_bits=int(rng.integers(3, 9)), amp_ofs=int(rng.integers(1, 4)), books=fb))
            continue
        n_posts = floor_posts if floor_posts else int(rng.integers(4, 30))
        rb = range_bits if range_bits else int(rng.integers(ilog(size0 // 2 - 1), ilog(size1 // 2 - 1) + 1))
        rb = max(rb, ilog(n_posts))   # enough distinct X values
        classes = []
        for _ in range(int(rng.integers(1, 5))):
            sub_bits = int(rng.integers(0, 3))
            cl = dict(dim=int(rng.integers(1, 5)), sub_bits=sub_bits, books=[])
            if sub_bits:
                cl["master"] = add_book(entries=int(rng.integers(2, 32)), dims=1, max_len=10 if not deep_books else 24,
                                        deep=deep_books)
            for _ in range(1 << sub_bits):
                if int(rng.integers(0, 5)) == 0:
                    cl["books"].append(-1)
                else:
                    cl["books"].append(add_book(entries=int(rng.integers(2, 130)), dims=1, max_len=12 if not deep_books else 30,
                                                deep=deep_books))
            classes.append(cl)
        if floor_posts:
            if len(classes) == 1:
                classes.append(dict(classes[0]))      # the same books, another dimension
            classes[0]["dim"] = 1                    # so that the requested post count can be hit exactly
            classes[-1]["dim"] = max(classes[-1]["dim"], 3 if len(classes) > 1 else 1)
        part_class, total = [], 2
        while total < n_posts and len(part_class) < 31:
            c = int(rng.integers(0, len(classes)))
            left = 31 - len(part_class)          # partitions still available: do not run out before the count
            if total + classes[c]["dim"] > n_posts or (floor_posts and (n_posts - total) > (left - 1) * 1 + classes[c]["dim"]
                                                        and classes[c]["dim"] < max(cl["dim"] for cl in classes)):
                fit = [i for i, cl in enumerate(classes) if total + cl["dim"] <= n_posts]
                if not fit:
                    break
                c = max(fit, key=lambda i: classes[i]["dim"]) if floor_posts and (n_posts - total) > left else fit[0]
            part_class.append(c)
            total += classes[c]["dim"]
        used = sorted(set(part_class))   # classes are numbered 0..max(part_class): drop unused tail classes
        classes = classes[:max(used) + 1] if used else classes[:0]
        xs = rng.choice(np.arange(1, 1 << rb), total - 2, replace=False).tolist() if total > 2 else []
        floors.append(dict(type=1, part_class=part_class, classes=classes, multiplier=int(rng.integers(1, 5)), range_bits=rb,
                           xs=[int(x) for x in xs], posts=total))

    # ---- residues --------------------------------------------------------------------------------
    residues = []
    for rt in res_types:
        ps = psize if psize else int(rng.choice([8, 16, 32]))
        nclass = int(rng.integers(2, 7))
        cdim = int(rng.integers(1, 4))
        while nclass ** cdim > 4096:
            cdim -= 1
        partvals = nclass ** cdim
        cb = add_book(entries=partvals + classbook_extra, dims=cdim, max_len=12 if not deep_books else 26, deep=deep_books,
                      lookup=0)
        cascade, rbooks = [], []
        for _ in range(nclass):
            c = int(rng.integers(0, 8)) if int(rng.integers(0, 4)) else int(rng.integers(0, 256)) & 0x1f
            cascade.append(c)
            for bit in range(8):
                if c >> bit & 1:
                    d = int(rng.choice(vq_dims))
                    ent = int(rng.integers(2, 82))
                    lk = int(rng.integers(1, 3))
                    if lk == 1 and lookup1_values(ent, d) < 1:
                        lk = 2
                    rbooks.append(add_book(entries=ent, dims=d, lookup=lk, max_len=11 if not deep_books else 32, deep=deep_books,
                                           value_bits=int(rng.integers(2, 7)), scale_exp=-8))
        half1 = size1 // 2
        mult = channels if rt == 2 else 1
        begin = int(rng.choice([0, 0, ps, 3 * ps]))
        end = int(rng.choice([half1 * mult, half1 * mult // 2, half1 * mult - ps, half1 * mult + 5 * ps]))
        # a VQ dimension that does not divide the partition writes past the partition's end; past the end of
        # the type-2 vector the reference slices out of range (Residue1.cs:24), so leave room there
        maxd = max(vq_dims)
        if rt == 2:
            end = min(end, (size0 // 2) * mult - maxd, half1 * mult - maxd)
        if (end - begin) // ps * (1 if rt == 2 else channels) > 500:   # the GPU path takes 512 units per residue
            end = begin + 500 // (1 if rt == 2 else channels) * ps
        residues.append(dict(type=rt, begin=begin, end=max(end, 0), part_size=ps, class_book=cb, cascade=cascade, books=rbooks))

    # ---- mappings: one per block size so that long and short blocks use different ones ------------
    mappings = []
    for mi in range(2):
        pairs = []
        while len(pairs) < coupling and channels > 1:
            a, b = (int(x) for x in rng.choice(channels, 2, replace=False))
            pairs.append((a, b))
        if submaps > 1:
            mux = [int(rng.integers(0, submaps)) for _ in range(channels)]
            if mi == 0:
                mux[0] = 0   # at least one populated submap with a known index
        else:
            mux = [0] * channels
        sub = []
        for j in range(submaps):
            sub.append((int(rng.integers(0, len(floors))) if j or mi else 0, (j + mi) % len(residues)))
        mappings.append(dict(submaps=submaps, coupling=pairs, mux=mux, sub=sub))
    modes = [(0, 0), (1, 1), (1, 0), (0, 1)][:int(rng.integers(2, 5))]

    # ---- packets -----------------------------------------------------------------------------------
    idw = BitWriter()
    for b in b"\x01vorbis":
        idw.write(b, 8)
    idw.write(0, 32)
    idw.write(channels, 8)
    idw.write(rate, 32)
    idw.write(0, 32)
    idw.write(128000, 32)
    idw.write(0, 32)
    idw.write(lg[0], 4)
    idw.write(lg[1], 4)
    idw.write(1, 1)
    sw = BitWriter()
    for b in b"\x05vorbis":
        sw.write(b, 8)
    sw.write(len(books) - 1, 8)
    for bk in books:
        write_book(sw, bk)
    sw.write(0, 6)
    sw.write(0, 16)
    sw.write(len(floors) - 1, 6)
    for fl in floors:
        (write_floor0 if fl["type"] == 0 else write_floor1)(sw, fl)
    sw.write(len(residues) - 1, 6)
    for rs in residues:
        write_residue(sw, rs)
    sw.write(len(mappings) - 1, 6)
    for mp in mappings:
        write_mapping(sw, mp, channels)
    sw.write(len(modes) - 1, 6)
    for flag, mp in modes:
        sw.write(flag, 1)
        sw.write(0, 16)
        sw.write(0, 16)
        sw.write(mp, 8)
    sw.write(1, 1)
    assert len(books) <= 256
    info = dict(channels=channels, size0=size0, size1=size1, modes=modes, n_books=len(books), floors=floors,
                residues=residues, mappings=mappings)
    return idw.bytes(), sw.bytes(), info


def packet_counts(info, kinds):
    """Samples each packet of a stream makes available (Mode.GetPacketInfo, Mode.cs:30-66 + ReadNextPacket):
    kinds = list of long-block flags; window flags follow the neighbours."""
    s0, s1 = info["size0"], info["size1"]
    out = []
    for i, lb in enumerate(kinds):
        prev = kinds[i - 1] if i else 1
        nxt = kinds[i + 1] if i + 1 < len(kinds) else 1
        n = s1 if lb else s0
        if not lb:
            prev = nxt = 1
        ls = 0 if prev else (n - s0) // 4
        rs = n // 2 if nxt else (n * 3 - s0) // 4
        out.append(0 if i == 0 else rs - ls)
    return out


def make_packets(rng, info, n_packets, mean_len=120, kinds=None):
    """Audio packets: header bit 0, a mode of the wanted block size, window flags that agree with the
    neighbours, then random bits."""
    modes = info["modes"]
    mode_bits = ilog(len(modes) - 1)
    if kinds is None:
        kinds, cur = [], 1
        for _ in range(n_packets):
            kinds.append(cur)
            u = rng.random()
            cur = (0 if u < 0.2 else 1) if cur else (1 if u < 0.4 else 0)
    packets = []
    for i, lb in enumerate(kinds):
        cand = [k for k, (flag, _) in enumerate(modes) if flag == lb]
        bw = BitWriter()
        bw.write(0, 1)
        bw.write(cand[int(rng.integers(0, len(cand)))], mode_bits)
        if lb:
            bw.write(kinds[i - 1] if i else 1, 1)
            bw.write(kinds[i + 1] if i + 1 < len(kinds) else 1, 1)
        n = int(rng.integers(0, 2 * mean_len))
        head = bw.bytes()
        body = bytearray(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
        if body and bw.n % 8:
            # keep the random bits that share the last header byte
            head = head[:-1] + bytes([head[-1] | (body[0] & (0xff << (bw.n % 8)) & 0xff)])
            body = body[1:]
        packets.append(head + bytes(body))
    return packets, kinds


def make_stream(seed, shape, n_packets=40, mean_len=120, eos_trim=0, comment=b"synthvorbis"):
    """A whole Ogg Vorbis stream of the given shape: dict(id, setup, packets, ogg, info, kinds)."""
    rng = np.random.default_rng(seed)
    kw = SHAPES[shape] if isinstance(shape, str) else shape
    idp, setup, info = make_setup(rng, **kw)
    packets, kinds = make_packets(rng, info, n_packets, mean_len)
    cw = BitWriter()
    for b in b"\x03vorbis":
        cw.write(b, 8)
    cw.write(len(comment), 32)
    for b in comment:
        cw.write(b, 8)
    cw.write(0, 32)
    cw.write(1, 1)
    counts = packet_counts(info, kinds)
    pages = [([idp], 0)]
    for h in (cw.bytes(), setup):
        # a header packet may need more than one page's worth of lacing values: the simple muxer cannot
        # continue packets, so the caller keeps setups below 255 * 255 bytes
        assert len(h) < 255 * 255, "setup header too large for the simple muxer"
        pages.append(([h], 0))
    pos, cur, segs = 0, [], 0
    for i, p in enumerate(packets):
        need = len(p) // 255 + 1
        if cur and (segs + need > 255 or len(cur) >= 7):
            pages.append((cur, pos))
            cur, segs = [], 0
        cur.append(p)
        segs += need
        pos += counts[i]
    if cur:
        pages.append((cur, max(pos - eos_trim, 0)))
    ogg = oggmux.mux(pages, serial=0x5EED)
    return dict(id=idp, setup=setup, packets=packets, ogg=ogg, info=info, kinds=kinds, total=pos)
