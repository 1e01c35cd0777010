#!/usr/bin/env python3
"""Independent symbol-level walker -> tests/golden/stage_pin.json   (test infrastructure only).

A second reading of the reference's integer stages, written from the algorithm notes in SURVEY.md
Appendix A (A1 bitstream, A2 codebooks, A3 floor 1, A4 residues, A5 mapping) and the Vorbis I
specification -- NOT from oracle/ and sharing no code with it: pure Python, its own Ogg page splitter,
its own codeword assignment (canonical "lowest free leaf" tree walk instead of stb's marker table), a
dictionary decoder instead of lookup tables, big-integer bit access.  It walks every audio packet of the
four TestFiles and records, per packet, digests of
    * the Codebook.DecodeScalar result sequence (floor books, classwords, VQ entries, in call order),
    * the raw floor-1 posts and the unwrapped final Y + step flags of every channel,
    * the residue partition classes in the order the reference visits them.
tests/test_oracle.py asserts the C oracle against these digests, which removes the single-author
common-mode risk on the integer stages (it does not pin them to the reference binary: nothing here can).

    python tests/golden/make_stage_pin.py        # rewrites tests/golden/stage_pin.json
"""
import hashlib
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(os.path.dirname(HERE), "data")
FILES = ["1test", "2test", "3test", "issue6test"]


# ---- Ogg: pages -> packets (A8) -------------------------------------------------------------------
def ogg_packets(data):
    """All packets of the first logical stream, in order (continued packets stitched)."""
    pos, serial, cur, out = 0, None, b"", []
    while pos + 27 <= len(data):
        assert data[pos:pos + 4] == b"OggS", "lost sync at %d" % pos
        flags = data[pos + 5]
        ser = struct.unpack_from("<I", data, pos + 14)[0]
        nseg = data[pos + 26]
        lacing = data[pos + 27:pos + 27 + nseg]
        body = pos + 27 + nseg
        if serial is None:
            serial = ser
        if ser == serial:
            if not (flags & 1):
                cur = b""   # a fresh packet starts this page (an unfinished one would be dropped)
            for l in lacing:
                cur += data[body:body + l]
                body += l
                if l < 255:
                    out.append(cur)
                    cur = b""
        else:
            body += sum(lacing)
        pos = body if ser == serial else pos + 27 + nseg + sum(lacing)
    return out


class Bits:
    """LSB-first reader over one packet: reads past the end return the bits that exist, zero-extended (A1)."""

    def __init__(self, data):
        self.v = int.from_bytes(data, "little")
        self.n = len(data) * 8
        self.pos = 0

    def left(self):
        return self.n - self.pos

    def peek(self, k):
        return (self.v >> self.pos) & ((1 << k) - 1) if self.pos < self.n else 0

    def read(self, k):
        x = self.peek(k)
        self.pos = min(self.pos + k, self.n)
        return x


def ilog(x):
    return x.bit_length() if x > 0 else 0


# ---- codebooks (A2): lengths -> prefix code, decode by walking a dict of (length, code) ----------------
class Book:
    def __init__(self, br):
        assert br.read(24) == 0x564342
        self.dims = br.read(16)
        self.entries = br.read(24)
        lens = [0] * self.entries
        if br.read(1):   # ordered
            cur = br.read(5) + 1
            i = 0
            while i < self.entries:
                cnt = br.read(ilog(self.entries - i))
                for _ in range(cnt):
                    lens[i] = cur
                    i += 1
                cur += 1
        else:
            sparse = br.read(1)
            for i in range(self.entries):
                if not sparse or br.read(1):
                    lens[i] = br.read(5) + 1
        self.lens = lens
        # Vorbis codeword assignment: every entry, in order, takes the lowest-valued free leaf at its depth,
        # the tree being filled MSB-first.  free[d] = list of free node values at depth d.
        self.code = {}   # (length, value read LSB-first from the stream) -> entry
        used = [(i, l) for i, l in enumerate(lens) if l > 0]
        self.single = len(used) == 1
        taken = set()    # node prefixes (depth, msb value) that are occupied or have an occupied descendant

        def is_free(depth, val):
            # free when no ancestor-or-self is a leaf and no descendant is taken
            for d in range(1, depth + 1):
                if (d, val >> (depth - d)) in leaves:
                    return False
            return (depth, val) not in taken

        leaves = set()
        for idx, l in used:
            val = 0
            # smallest MSB-first value at depth l that is free
            while not is_free(l, val):
                val += 1
                assert val < (1 << l), "over-subscribed codebook"
            leaves.add((l, val))
            for d in range(1, l + 1):
                taken.add((d, val >> (l - d)))
            # the stream delivers the MSB of the codeword first, bits are read LSB-first: reverse
            rev = int(format(val, "0%db" % l)[::-1], 2)
            self.code[(l, rev)] = idx
        self.maxlen = max([l for _, l in used], default=0)
        self.map_type = br.read(4)
        if self.map_type:
            br.read(32)
            br.read(32)
            vbits = br.read(4) + 1
            br.read(1)
            if self.map_type == 1:
                n = lookup1(self.entries, self.dims)
            else:
                n = self.entries * self.dims
            for _ in range(n):
                br.read(vbits)

    def decode(self, br):
        """One symbol or -1 (no bits left / no code matches); consumes the code's length, at most what is left."""
        if br.left() <= 0:
            return -1
        for l in range(1, self.maxlen + 1):
            e = self.code.get((l, br.peek(l)))
            if e is not None:
                br.read(l)
                return e
        return -1


def lookup1(entries, dims):
    r = 0
    while (r + 1) ** dims <= entries:
        r += 1
    return r


# ---- setup header (A3-A5) ---------------------------------------------------------------------------
class Floor1:
    def __init__(self, br):
        self.parts = [br.read(4) for _ in range(br.read(5))]
        self.classes = []
        for _ in range(max(self.parts, default=-1) + 1):
            dim = br.read(3) + 1
            sub = br.read(2)
            master = br.read(8) if sub else None
            books = [br.read(8) - 1 for _ in range(1 << sub)]
            self.classes.append((dim, sub, master, books))
        self.mult = br.read(2) + 1
        self.range = [256, 128, 86, 64][self.mult - 1]
        self.ybits = [8, 7, 7, 6][self.mult - 1]
        rb = br.read(4)
        self.x = [0, 1 << rb]
        for c in self.parts:
            for _ in range(self.classes[c][0]):
                self.x.append(br.read(rb))

    def neighbours(self, i):
        lo = max((j for j in range(i) if self.x[j] < self.x[i]), key=lambda j: self.x[j])
        hi = min((j for j in range(i) if self.x[j] > self.x[i]), key=lambda j: self.x[j])
        return lo, hi


class Residue:
    def __init__(self, br, rtype, books):
        self.type = rtype
        self.begin, self.end = br.read(24), br.read(24)
        self.psize = br.read(24) + 1
        self.nclass = br.read(6) + 1
        self.classbook = br.read(8)
        self.cascade = []
        for _ in range(self.nclass):
            low = br.read(3)
            high = br.read(5) if br.read(1) else 0
            self.cascade.append(low | (high << 3))
        self.books = [[br.read(8) if (c >> s) & 1 else None for s in range(8)] for c in self.cascade]
        self.stages = max((ilog(c) for c in self.cascade), default=0)


def parse_setup(idp, setup):
    chans = idp[11]
    bs = idp[28]
    sizes = (1 << (bs & 15), 1 << (bs >> 4))
    br = Bits(setup)
    assert br.read(8) == 5 and bytes(br.read(8) for _ in range(6)) == b"vorbis"
    books = [Book(br) for _ in range(br.read(8) + 1)]
    for _ in range(br.read(6) + 1):
        br.read(16)
    floors = []
    for _ in range(br.read(6) + 1):
        assert br.read(16) == 1, "floor 1 only in the TestFiles"
        floors.append(Floor1(br))
    residues = []
    for _ in range(br.read(6) + 1):
        t = br.read(16)
        residues.append(Residue(br, t, books))
    mappings = []
    for _ in range(br.read(6) + 1):
        assert br.read(16) == 0
        submaps = br.read(4) + 1 if br.read(1) else 1
        steps = []
        if br.read(1):
            for _ in range(br.read(8) + 1):
                steps.append((br.read(ilog(chans - 1)), br.read(ilog(chans - 1))))
        assert br.read(2) == 0
        mux = [br.read(4) for _ in range(chans)] if submaps > 1 else [0] * chans
        sub = []
        for _ in range(submaps):
            br.read(8)
            sub.append((br.read(8), br.read(8)))
        mappings.append((steps, mux, sub))
    modes = []
    for _ in range(br.read(6) + 1):
        flag = br.read(1)
        br.read(32)
        modes.append((flag, br.read(8)))
    assert br.read(1) == 1
    return chans, sizes, books, floors, residues, mappings, modes


# ---- one audio packet: the symbol walk -----------------------------------------------------------------
def walk_packet(pkt, st):
    chans, sizes, books, floors, residues, mappings, modes = st
    br = Bits(pkt)
    scal, out = [], {}

    def dec(b):
        v = books[b].decode(br)
        scal.append(v)
        return v

    if br.read(1) != 0:
        return None
    mode = br.read(ilog(len(modes) - 1))
    flag, mapping = modes[mode]
    if flag:
        br.read(2)
    half = sizes[flag] // 2
    steps, mux, sub = mappings[mapping]
    raw, final, flags_out, energy = [], [], [], []
    for ch in range(chans):
        fl = floors[sub[mux[ch]][0]]
        posts = []
        ok = br.read(1) == 1
        if ok:
            posts = [br.read(fl.ybits), br.read(fl.ybits)]
            for c in fl.parts:
                dim, subb, master, cb = fl.classes[c]
                cval = 0
                if subb:
                    cval = dec(master)
                    if cval < 0:
                        ok = False
                        break
                for _ in range(dim):
                    b = cb[cval & ((1 << subb) - 1)]
                    cval >>= subb
                    v = 0
                    if b >= 0:
                        v = dec(b)
                        if v < 0:
                            ok = False
                            break
                    posts.append(v)
                if not ok:
                    break
        if not ok:
            posts = []
        raw.append(list(posts))
        energy.append(bool(posts))
        # unwrap (A3): predicted value from the neighbours among earlier posts, "room" logic
        y = list(posts)
        step = [1, 1] + [0] * max(len(posts) - 2, 0)
        for i in range(2, len(posts)):
            lo, hi = fl.neighbours(i)
            dy, adx = y[hi] - y[lo], fl.x[hi] - fl.x[lo]
            off = abs(dy) * (fl.x[i] - fl.x[lo]) // adx
            pred = y[lo] - off if dy < 0 else y[lo] + off
            val = posts[i]
            hiroom, loroom = fl.range - pred, pred
            room = 2 * min(hiroom, loroom)
            if val:
                step[lo] = step[hi] = step[i] = 1
                if val >= room:
                    y[i] = val - loroom + pred if hiroom > loroom else pred - val + hiroom - 1
                elif val & 1:
                    y[i] = pred - (val + 1) // 2
                else:
                    y[i] = pred + val // 2
            else:
                y[i] = pred
        final.append(y)
        flags_out.append(step[:len(posts)])
    # no-energy propagation through the coupling steps (A5)
    noexec = [not e for e in energy]
    for mag, ang in steps:
        if not (noexec[mag] and noexec[ang]):
            noexec[mag] = noexec[ang] = False
    # residue: stage -> partition group -> [stage 0: classwords] -> partition -> channel (A4)
    classes = []
    for si, (_, rnum) in enumerate(sub):
        rs = residues[rnum]
        chs = [c for c in range(chans) if mux[c] == si]
        skip = [noexec[c] for c in chs]
        if rs.type == 2:
            if all(skip):
                continue
            n_vec, vlen, skip = 1, half * len(chs), [False]
        else:
            n_vec, vlen = len(chs), half
        begin, end = min(rs.begin, vlen), min(rs.end, vlen)
        nparts = (end - begin) // rs.psize if end > begin else 0
        cb = books[rs.classbook]
        cdim = cb.dims
        partvals = rs.nclass ** cdim
        cls = [[0] * nparts for _ in range(n_vec)]
        stop = False
        for stage in range(rs.stages):
            p = 0
            while p < nparts and not stop:
                if stage == 0:
                    for v in range(n_vec):
                        if skip[v]:
                            continue
                        w = dec(rs.classbook)
                        if w < 0 or w >= partvals:
                            stop = True
                            break
                        for k in range(cdim - 1, -1, -1):
                            if p + k < nparts:
                                cls[v][p + k] = w % rs.nclass
                            w //= rs.nclass
                    if stop:
                        break
                for k in range(cdim):
                    if p >= nparts or stop:
                        break
                    for v in range(n_vec):
                        if skip[v]:
                            continue
                        c = cls[v][p]
                        if stage == 0:
                            classes.append(c)
                        b = rs.books[c][stage]
                        if b is None:
                            continue
                        d = books[b].dims
                        n = rs.psize // d if rs.type == 0 else -(-rs.psize // d)
                        for _ in range(n):
                            if dec(b) < 0:
                                stop = True
                                break
                        if stop:
                            break
                    p += 1
            if stop:
                break
    out = dict(scalars=scal, raw=raw, final=final, flags=flags_out, classes=classes, bits=br.pos)
    return out


def digest(seq):
    return hashlib.sha256(",".join(str(int(x)) for x in seq).encode()).hexdigest()[:16]


def pin_file(name):
    with open(os.path.join(DATA, name + ".ogg"), "rb") as f:
        data = f.read()
    pk = ogg_packets(data)
    st = parse_setup(pk[0], pk[2])
    rows = []
    for p in pk[3:]:
        w = walk_packet(p, st)
        if w is None:
            rows.append(None)
            continue
        rows.append(dict(
            n_scalars=len(w["scalars"]), scalars=digest(w["scalars"]),
            n_classes=len(w["classes"]), classes=digest(w["classes"]),
            post_counts=[len(r) for r in w["raw"]],
            raw_posts=digest([v for r in w["raw"] for v in r]),
            final_y=digest([v for r in w["final"] for v in r]),
            step_flags=digest([v for r in w["flags"] for v in r]),
            bits=w["bits"]))
    return rows


def main():
    out = {}
    for name in FILES:
        rows = pin_file(name)
        out[name] = rows
        n = sum(r["n_scalars"] for r in rows if r)
        print("%s: %d audio packets, %d DecodeScalar calls" % (name, len(rows), n), file=sys.stderr)
    with open(os.path.join(HERE, "stage_pin.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
