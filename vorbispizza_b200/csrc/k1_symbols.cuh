// k1_symbols.cuh -- the entropy stage, split by the shape of its parallelism.
//
//   K1a  vpz_k1a_symbols : ONE LANE PER PACKET.  The bit cursor of a packet is strictly serial, so a
//        warp decodes 32 packets at once, each lane running the reference's symbol sequence for its
//        own packet: floor1 unpack + unwrap, partition classwords, VQ entry indices.  It produces a
//        compact symbol record (floor line segments, partition classes, entry indices) and never
//        touches a float.  The residue walk is flattened into one loop with one DecodeScalar per
//        iteration so that lanes at different (stage, partition) positions still share every
//        instruction.
//   K1b  vpz_k1b_spectrum : ONE WARP PER PACKET, all data-parallel: VQ lookups accumulated in shared
//        memory in the reference's order, inverse coupling, floor line render, dB multiply, spectrum
//        store.
//
// Replaces, per audio packet (reference file:line):
//   Floor1.Unpack / UnwrapPosts         Floor1.cs:162-219, 270-370          (K1a)
//   Mapping.DecodePacket flag logic     Mapping.cs:121-130                  (K1a)
//   Residue0/1/2.Decode control flow    Residue0.cs:117-206, Residue2.cs:12-52   (K1a)
//   Codebook.DecodeScalar               Codebook.cs:301-335  (two-level table, same symbol + bits) (K1a)
//   Residue0/1.WriteVectors             Residue0.cs:208-231, Residue1.cs:12-34   (K1b)
//   Mapping.ApplyCoupling               Mapping.cs:198-269                  (K1b)
//   Floor1.Apply / RenderLineMulti      Floor1.cs:222-268, 372-397          (K1b, closed-form line)
#pragma once
#include "k1_params.h"

#ifndef VPZ_EMU
#define VPZ_DEV __device__ __forceinline__
#define VPZ_DEVN __device__ __noinline__
#define VPZ_LDG(p) __ldg(p)
// packet bytes are read once, one 128-byte line per lane at a time: keep them out of L1 (ld.global.cg) so
// the lines stay available for the Huffman tables that every lane re-reads
#ifdef VPZ_STREAM_CA
#define VPZ_LDSTREAM(p) __ldca(p)
#else
#define VPZ_LDSTREAM(p) __ldcg(p)
#endif
// setup images are reached through a pointer table in global memory: tell the compiler that what they point to is
// global memory too (LDG instead of generic loads with an address-space check)
#ifndef VPZ_NO_ASSUME
#define VPZ_ASSUME_GLOBAL(p) __builtin_assume(__isGlobal(p))
#else
#define VPZ_ASSUME_GLOBAL(p)
#endif
#else
#define VPZ_ASSUME_GLOBAL(p)
#define VPZ_DEV inline
#define VPZ_DEVN inline
#define VPZ_LDG(p) (*(p))
#define VPZ_LDSTREAM(p) (*(p))
#endif

// word offsets inside vpz_packet_dump (include/vpz.h) -- keep in sync
#define DUMP_STATUS 0
#define DUMP_MODE 1
#define DUMP_BLOCK 2
#define DUMP_INFO 3
#define DUMP_BITS 9
#define DUMP_EXEC 10
#define DUMP_NOEXEC 11
#define DUMP_SCALARS_N 12
#define DUMP_CLASSES_N 13
#define DUMP_POSTCOUNT 14
#define DUMP_RAWPOSTS (14 + 8)
#define DUMP_FINALY (14 + 8 + 8 * 64)
#define DUMP_STEPFLAGS (14 + 8 + 16 * 64)

// ---- symbol record (K1a -> K1b), 32-bit words at P.rec + pkt.rec_off --------------------------
//   [0] own_mask | noexec_mask << 8 | status << 16 | long_block << 24
//   [1] number of VQ entries written at P.ent + pkt.ent_off (uint16 each)
//   [2] final bit cursor
//   [3] mapping index | residue instance of submap 0 << 8 | floor-0 entry indices that precede the residue's << 16
//       (found through the mode; K1b does not walk the packet again)
//   then per channel K1_SEG_WORDS words: floor 1: [0] = number of line segments n, [1..n+1] = points x | y << 16;
//       floor 0 with energy: [0] = 0x80000000 | book number, [1] = raw amplitude, [2] = first entry index
//       (relative to the packet's entry area), [3] = number of entry indices; 0 in [0] = silent channel
//   then one byte per unit (partition * nvec + vector, the decode order inside a stage): the partition class;
//       the submaps of a mapping follow each other, each padded to whole words
#define K1_REC_HDR 4
#define K1_SEG_WORDS 68
#define K1_MAX_UNITS 512      // vectors * partitions per packet (setup.cpp refuses more)
#define K1B_THREADS 128       // K1b: threads (one CTA) per packet

// Asynchronous 16-byte copies global -> shared (LDGSTS, L2 only): how K1a streams its packet bytes.
#ifndef VPZ_EMU
#define K1_CP16(dst, src) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory")
#define K1_CP_WAIT() asm volatile("cp.async.wait_all;" ::: "memory")
#else
#define K1_CP16(dst, src) memcpy((dst), (src), 16)
#define K1_CP_WAIT() ((void)0)
#endif

#define K1A_RING_BYTES(threads) ((size_t)(threads) * 32)   // two 16-byte slots per thread
// K1a keeps the partition classes of a packet's first K1A_CLS_CACHE units in shared memory as well: the walk reads a
// unit's class back to find its book, and the read-back of its own global store missed L1 (stores do not allocate)
#define K1A_CLS_CACHE 64
#define K1A_CLS_BYTES(threads) ((size_t)(threads) * K1A_CLS_CACHE)

struct K1Bits {
  uint32_t woff;   // word offset of the packet in the batch byte buffer (a multiple of 4: packets start on 16 bytes)
  int pos, nbits;
  // register window over the packet: lo = w[wi], hi = w[wi + 1], nx = w[wi + 2] with wi = pos >> 5; a peek is
  // one funnel shift.  The words come from a per-thread RING of two 16-byte chunks in shared memory that
  // cp.async fills one chunk AHEAD: when the window enters chunk c, chunk c + 1 is requested and lands while
  // the ~28 codewords of chunk c are decoded.  (A plain load of w[wi + 2] at every word crossing looked like a
  // prefetch in the source, but the compiler parked the result in a temporary and moved it into `nx` at the
  // end of the same trip: 19 % of K1a's stall samples waited for that L2 round trip.)  The packet is followed
  // by at least 12 zero bytes (VpzPktIn), so w[wi + 2] is packet data or zero padding for every pos <= nbits;
  // chunk reads may run up to 32 bytes further (into the next packet or the buffer's slack), never used.
  uint32_t lo, hi, nx;
  int wi;
  uint4* ring;     // this thread's two slots (32 consecutive bytes): word j of the packet sits at word j & 7
  uint32_t* cls;   // class cache: this thread's word of row 0, rows are blockDim.x words apart (NULL: no cache)
};

VPZ_DEV uint32_t k1_ring_word(const K1Bits& b, int j) {   // word j of the packet; its chunk is resident
  return reinterpret_cast<const uint32_t*>(b.ring)[j & 7];
}

VPZ_DEV uint8_t* k1_cls_slot(const K1Bits& b, int u) {   // u < K1A_CLS_CACHE
  return reinterpret_cast<uint8_t*>(b.cls + (u >> 2) * blockDim.x) + (u & 3);
}

VPZ_DEV void k1_bits_init(K1Bits& b, const uint32_t* bytes, uint32_t woff, int byte_len, uint4* ring) {
  b.woff = woff;
  b.pos = 0;
  b.nbits = byte_len * 8;
  b.wi = 0;
  b.ring = ring;
  b.cls = nullptr;
  K1_CP16(ring, bytes + woff);
  K1_CP16(ring + 1, bytes + woff + 4);
  K1_CP_WAIT();
  b.lo = k1_ring_word(b, 0);
  b.hi = k1_ring_word(b, 1);
  b.nx = k1_ring_word(b, 2);
}

VPZ_DEV uint32_t k1_peek32(const K1Bits& b) { return __funnelshift_r(b.lo, b.hi, b.pos & 31); }

// move the cursor forward by at most 32 bits
VPZ_DEV void k1_bits_seek(K1Bits& b, const uint32_t* bytes, int np) {
  b.pos = np;
  const int ni = np >> 5;
  if (ni != b.wi) {
    b.lo = b.hi;
    b.hi = b.nx;
    b.wi = ni;
    const int j = ni + 2;
    if ((j & 3) == 0) {
      // the window enters a new chunk: it was requested four words ago; its predecessor's slot is free now
      // (those words are in lo / hi or used up), so the chunk after it goes there
      K1_CP_WAIT();
      b.nx = k1_ring_word(b, j);
      K1_CP16(b.ring + (((j >> 2) + 1) & 1), bytes + b.woff + j + 4);
    } else {
      b.nx = k1_ring_word(b, j);
    }
  }
}

// VorbisPacket.ReadBits (VorbisPacket.cs:157-164): zero-extended, truncated at the end, n <= 32
VPZ_DEV uint32_t k1_read(K1Bits& b, const uint32_t* bytes, int n) {
  if (n <= 0) return 0;
  uint32_t v = k1_peek32(b);
  if (n < 32) v &= (1u << n) - 1u;
  int np = b.pos + n;
  k1_bits_seek(b, bytes, np < b.nbits ? np : b.nbits);
  return v;
}

// a codebook as the decoder needs it: three registers, no pointers
struct K1Book {
  uint32_t l1_off;   // word offset of the first-level table in the blob
  uint32_t l1_mask;
  uint32_t meta;     // l1_bits | book index << 8 | shared-memory offset of the first-level table << 16 (K1A_SM_NONE: global)
};

// K1a shared-memory tables (vpz_k1a_symbols_sm): [K1A_SM_WORDS words of first-level tables][256 x uint16 offsets]
#ifndef VPZ_EMU
extern __shared__ uint32_t k1a_sm[];
#define K1A_SM k1a_sm
#else
#define K1A_SM (reinterpret_cast<uint32_t*>(emu::t_block->smem))
#endif
// shared-memory word offset of a book's first-level table, K1A_SM_NONE when the lane is not on the staged setup
// or the book is not staged
template <bool SM>
VPZ_DEV uint32_t k1a_soff(uint32_t sm_on, int book) {
  if (!SM || !sm_on) return K1A_SM_NONE;
  return (K1A_SM[K1A_SM_WORDS + (book >> 1)] >> ((book & 1) * 16)) & 0xffffu;
}

// dims (u16) and l1_bits (u8) share one 32-bit word at byte offset 24 of VpzBook
VPZ_DEV int k1_book_dims(const VpzBook* bk) {
  return (int)(VPZ_LDG(reinterpret_cast<const uint32_t*>(bk) + 6) & 0xffffu);
}

VPZ_DEV K1Book k1_book(const uint32_t* blob, const VpzBook* books, int idx, uint32_t soff = K1A_SM_NONE) {
  const VpzBook* bk = books + idx;
  K1Book r;
  r.l1_off = VPZ_LDG(&bk->l1_off);
  const uint32_t l1_bits = (VPZ_LDG(reinterpret_cast<const uint32_t*>(bk) + 6) >> 16) & 0xffu;
  r.l1_mask = (1u << l1_bits) - 1u;
  r.meta = l1_bits | ((uint32_t)idx << 8) | (soff << 16);
  return r;
}

// Codebook.DecodeScalar (Codebook.cs:301-335).  -1: no bits left or no code matches.
template <bool DEBUG, bool SM = false>
VPZ_DEV int k1_decode(K1Bits& b, const K1Book& bk, const uint32_t* blob, const K1Params& P, int& nscal) {
  int sym = -1;
  if (b.pos < b.nbits) {
    uint32_t x = k1_peek32(b);
    const uint32_t* l1 = blob + bk.l1_off;
    // first level from shared memory when the book's table is staged there: a warp's 32 lookups go to 32
    // unrelated entries, and from L1 the warp waits for its slowest lane (any miss = an L2 round trip)
    uint32_t e;
    if (SM && (bk.meta >> 16) != K1A_SM_NONE) e = K1A_SM[(bk.meta >> 16) + (x & bk.l1_mask)];
    else e = VPZ_LDG(l1 + (x & bk.l1_mask));
    if (e & 0x80000000u) {  // longer than the first-level table: second-level table over the next bits
      const uint32_t l2b = e & 31u;
      e = VPZ_LDG(l1 + ((e >> 5) & 0x3ffffffu) + ((x >> (bk.meta & 0xffu)) & ((1u << l2b) - 1u)));
      if (e & 0x80000000u) {  // longer still: binary search in the sorted long codes of this prefix
        const VpzBook* vb = reinterpret_cast<const VpzBook*>(blob + reinterpret_cast<const VpzSetupHdr*>(blob)->books_off) +
                            ((bk.meta >> 8) & 0xffu);
        const uint32_t* ranges = blob + VPZ_LDG(&vb->range_off);
        const uint32_t* lcode = blob + VPZ_LDG(&vb->lcode_off);
        const uint32_t* linfo = blob + VPZ_LDG(&vb->linfo_off);
        uint32_t id = e & 0x7fffffffu;
        uint32_t lo = VPZ_LDG(ranges + 2 * id), hi = VPZ_LDG(ranges + 2 * id + 1);
        uint32_t m = __brev(x);
        while (hi - lo > 1) {
          uint32_t mid = (lo + hi) >> 1;
          if (VPZ_LDG(lcode + mid) <= m) lo = mid; else hi = mid;
        }
        uint32_t info = VPZ_LDG(linfo + lo);
        uint32_t len = info & 0xffu;
        e = (((m ^ VPZ_LDG(lcode + lo)) >> (32u - len)) == 0u) ? info : 0u;
      }
    }
    if (e != 0u) {
      sym = (int)(e >> 8);
      int np = b.pos + (int)(e & 0xffu);
      // SkipBits past the end (VorbisPacket.cs:248-292) stops at the end of the packet
      k1_bits_seek(b, P.bytes, np > b.nbits ? b.nbits : np);
    }
  }
  if (DEBUG) {
    if (P.dbg.scalars && nscal < P.dbg.scalars_cap) P.dbg.scalars[nscal] = sym;
    nscal++;
  }
  return sym;
}

// Floor1.RenderPoint (Floor1.cs:355-370)
// (the reciprocal table of the setup -- VpzSetupHdr.rcp_off, used by K1b -- was tried here as well: K1a 2.164 vs 2.151 ms,
// the table load is one more dependent memory access on a path that waits for memory already)
VPZ_DEV int k1_render_point(int x0, int y0, int x1, int y1, int X) {
  int dy = y1 - y0, adx = x1 - x0;
  int ady = dy < 0 ? -dy : dy;
  int off = ady * (X - x0) / adx;
  return dy < 0 ? y0 - off : y0 + off;
}

// Geometry of the residue vectors of one packet (Residue0.Decode, Residue0.cs:117-143; type 2 via
// Residue2.cs:12-52).  Shared by K1a and K1b so both walk the same units.
struct K1ResGeom {
  int rtype, nvec, vlen, begin, psize, part_count;
  uint32_t skip;       // bit v: vector v is not decoded
  bool any;
};
// C: channels of the submap, noexec: their do-not-decode flags (bit c = c-th channel of the submap)
VPZ_DEV K1ResGeom k1_res_geom(const VpzResidue* rs, int C, int half, uint32_t noexec) {
  K1ResGeom g;
  g.rtype = rs->type;
  g.nvec = g.rtype == 2 ? 1 : C;
  g.vlen = g.rtype == 2 ? half * C : half;
  g.skip = g.rtype == 2 ? ((noexec == ((1u << C) - 1u)) ? 1u : 0u) : noexec;
  g.begin = (int)rs->begin < g.vlen ? (int)rs->begin : g.vlen;
  int end = (int)rs->end < g.vlen ? (int)rs->end : g.vlen;
  int n = end - g.begin;
  g.psize = (int)rs->part_size;
  // partition sizes are powers of two in every stream an encoder has produced: a shift instead of the ~20
  // instructions of an integer division (same quotient)
  if (n <= 0) g.part_count = 0;
  else if ((g.psize & (g.psize - 1)) == 0) g.part_count = n >> (31 - __clz(g.psize));
  else g.part_count = n / g.psize;
  g.any = false;
  for (int v = 0; v < g.nvec; v++) g.any |= !((g.skip >> v) & 1u);
  return g;
}
// entries one (class, stage) unit holds: Residue0.WriteVectors decodes psize / dims entries,
// Residue1.WriteVectors steps i += dims until i >= psize
VPZ_DEV int k1_unit_entries(int rtype, int psize, int dims) {
  return rtype == 0 ? psize / dims : (psize + dims - 1) / dims;
}

// The residue walk of one submap (Residue0.Decode, Residue0.cs:117-206; Residue2.Decode, Residue2.cs:12-52):
// classwords and VQ entry indices.  C = channels of the submap, noexec = their do-not-decode flags,
// rec_cls = where the class bytes of this submap's units go.  Entry indices continue at ent_pos.
struct K1aOut {
  uint32_t ent_pos, ent_lo, ent_hi;   // four entry indices per 8-byte store (k1a_emit)
  int status, nscal, ncls;
};
VPZ_DEV void k1a_emit(const K1Params& P, K1aOut& o, int sym) {
  // four entry indices per 8-byte store: a lane's store is its own L1 tag lookup, and the entry
  // stream is the bulk of K1a's memory requests.  The indices are shifted in from the top of a 64-bit window
  // (two instructions, no branch); after four of them the window is exactly the four in order.
  o.ent_lo = __funnelshift_r(o.ent_lo, o.ent_hi, 16);
  o.ent_hi = (o.ent_hi >> 16) | ((uint32_t)sym << 16);
  o.ent_pos++;
  if ((o.ent_pos & 3u) == 0u) *reinterpret_cast<uint2*>(P.ent + (o.ent_pos - 4)) = uint2{o.ent_lo, o.ent_hi};
}
// the 1..3 indices behind the last full store: bring them down to the bottom of the window (zeros above)
VPZ_DEV void k1a_emit_flush(const K1Params& P, uint32_t ent_pos, uint32_t ent_lo, uint32_t ent_hi) {
  const uint32_t k = ent_pos & 3u;
  if (k) {
    const unsigned long long w = (((unsigned long long)ent_hi << 32) | ent_lo) >> (16u * (4u - k));
    *reinterpret_cast<uint2*>(P.ent + (ent_pos & ~3u)) = uint2{(uint32_t)w, (uint32_t)(w >> 32)};
  }
}
template <bool DEBUG, bool SM>
VPZ_DEV void k1a_residue(const K1Params& P, const uint32_t* blob, const VpzBook* books, K1Bits& b, const VpzResidue* rs,
                         int C, uint32_t noexec, int half, uint8_t* rec_cls, K1aOut& o, uint32_t sm_on) {
  const K1ResGeom g = k1_res_geom(rs, C, half, noexec);
  int& status = o.status;
  int& nscal = o.nscal;
  int& ncls = o.ncls;
  if (g.part_count > 0 && g.any && rs->max_stages > 0) {
    // Residue0.Decode (Residue0.cs:117-206) walks  stage -> partition group -> [stage 0: classwords] ->
    // partition -> vector.  Here the walk is flattened so that every trip round the loop decodes
    // exactly ONE codeword (lanes at different places of their packets still share every instruction)
    // and finding the next codeword never loops over classes or cascades:
    //  * a classword value indexes cw_tab (setup.cpp), which holds for every stage the units of its
    //    partition group that carry codewords; stage 0 consumes that mask right away, the masks of the
    //    later stages are OR-ed into a per-stage bit array over all units (unit = partition * nvec + vector,
    //    which is also the decode order inside a stage);
    //  * stages >= 1 walk their bit array with find-first-set;
    //  * the book and the codeword count of a unit come from unit_tab[class][stage] in one 8-byte load.
    uint32_t smask[8 * (K1_MAX_UNITS / 32)];         // [stage][chunk of 32 units]; row 0 unused
    // class per unit in decode order: written to the record for K1b and read back from there (plain
    // loads: the lane reads its own stores)
    const K1Book cb = k1_book(blob, books, rs->class_book, k1a_soff<SM>(sm_on, rs->class_book));
    const int cdim = rs->cdim, nvec = g.nvec, part_count = g.part_count;
    const int partvals = (int)rs->partvals;
    const int max_stages = rs->max_stages;
    const uint32_t dmap_off = rs->decode_map_off, unit_tab_off = rs->unit_tab_off, cw_tab_off = rs->cw_tab_off;
    const int nunits = part_count * nvec;
    const int nchunks = (nunits + 31) >> 5;
    // classes of groups a truncated packet never reaches must still be valid indices for K1b
    for (int u = 0; u < (nunits + 3) >> 2; u++) reinterpret_cast<uint32_t*>(rec_cls)[u] = 0;
    for (int i = nchunks; i < max_stages * nchunks; i++) smask[i] = 0;
    uint32_t vecmask = 0;                  // vectors that are decoded
    for (int v = 0; v < nvec; v++)
      if (!((g.skip >> v) & 1u)) vecmask |= 1u << v;
    int stage = 0, gpart = 0, chunk = 0;
    uint32_t cwmask = vecmask;             // classwords still to read for the current group (stage 0)
    uint32_t grp_acc = 0;                  // stage-0 units of the current group, collected from its classwords
    uint32_t cur_mask = 0;                 // units still to decode; bit i = unit ubase + i
    int ubase = 0;
    int dbg_slot = 0;                      // DEBUG: next unit of the group whose class goes to the dump
    int cw_v = 0;                          // vector of the classword being decoded
    bool in_class = false;
    int rem = 0;                           // entries still to decode in the current unit
    K1Book cur = cb;
    bool done = false;
    // GATING.  A trip of this loop has a light part (decode one VQ entry index, emit it) and a heavy part
    // (find the next unit, or decode a classword and spread its classes and masks).  The 32 lanes of a warp
    // sit at unrelated places of their packets, so when every lane may enter the heavy part on any trip,
    // nearly every trip pays for it while most lanes only need the light part.  The heavy part is therefore
    // only OPEN on every K1A_GATE-th trip; a lane that runs out of entries in between idles until the gate
    // opens.  Units hold 2 / 4 / 8 / 16 entries in every stream seen so far, so a lane mostly finishes a unit
    // when the gate opens again.  Scheduling only: the symbol sequence of a packet is unchanged.  Measured at
    // 4,096 streams: period 1 / 2 / 4 / 8 = 3.02 / 2.73 / 2.61 / 2.53 ms (the ungated loop it replaces: 2.69).
#ifndef K1A_GATE
#define K1A_GATE 8
#endif
    int trip = 0;
    while (!done) {
      if (rem == 0 && (K1A_GATE <= 1 || (trip & (K1A_GATE - 1)) == 0)) {
        // at most two rounds per open gate: a classword and the unit (or classword) behind it
        for (int round = 0; round < 2 && rem == 0 && !done; round++) {
          // ---- find the next codeword to decode ----
          for (;;) {
            if (cwmask) {                    // classword of the lowest pending vector
              cw_v = __ffs((int)cwmask) - 1;
              in_class = true;
              break;
            }
            in_class = false;
            if (DEBUG && stage == 0) {
              // the dump lists the class of every unit the reference VISITS, in order, idle ones too
              const int upto = cur_mask ? __ffs((int)cur_mask) - 1 : cdim * nvec - 1;
              for (; dbg_slot <= upto; dbg_slot++) {
                const int k = dbg_slot / nvec, v = dbg_slot - k * nvec;
                if (gpart + k >= part_count || !((vecmask >> v) & 1u)) continue;
                if (P.dbg.classes && ncls < P.dbg.classes_cap) P.dbg.classes[ncls] = rec_cls[(gpart + k) * nvec + v];
                ncls++;
              }
            }
            if (cur_mask) {
              const int u = ubase + __ffs((int)cur_mask) - 1;
              cur_mask &= cur_mask - 1;
              const int cu = (b.cls && u < K1A_CLS_CACHE) ? *k1_cls_slot(b, u) : rec_cls[u];
              const uint2 t = VPZ_LDG(reinterpret_cast<const uint2*>(blob + unit_tab_off) + (cu * 8 + stage));
              cur.l1_off = t.x;
              cur.l1_mask = (1u << (t.y & 0xffu)) - 1u;
              cur.meta = (t.y & 0xffffu) | (k1a_soff<SM>(sm_on, (int)((t.y >> 8) & 0xffu)) << 16);
              rem = (int)(t.y >> 16);
              break;
            }
            if (stage == 0) {                // next partition group: its classwords come first
              gpart += cdim;
              if (gpart < part_count) {
                cwmask = vecmask;
                dbg_slot = 0;
                continue;
              }
              chunk = -1;                    // stage 0 is finished: fall through to the first chunk of stage 1
              stage = 1;
            } else if (chunk + 1 >= nchunks) {
              stage++;
              chunk = -1;
            }
            if (stage >= max_stages) {
              done = true;
              break;
            }
            chunk++;
            cur_mask = smask[stage * nchunks + chunk];
            ubase = chunk * 32;
          }
          if (done || !in_class) break;
          const int sym = k1_decode<DEBUG, SM>(b, cb, blob, P, nscal);
          // quirk Q8 accepts idx < partvals*dim; beyond partvals the reference indexes past
          // _decodeMap and throws, so both ends are treated as "stop decoding this packet"
          if (sym < 0 || sym >= partvals) {
            status = 1;
            done = true;
            break;
          }
          const int left = part_count - gpart;   // partitions of this (possibly partial, last) group
          const uint8_t* dmap = reinterpret_cast<const uint8_t*>(blob + dmap_off) + sym * cdim;
          for (int kk = 0; kk < cdim; kk++)
            if (kk < left) {
              const int u2 = (gpart + kk) * nvec + cw_v;
              const uint8_t c2 = dmap[kk];
              rec_cls[u2] = c2;
              if (b.cls && u2 < K1A_CLS_CACHE) *k1_cls_slot(b, u2) = c2;
            }
          const uint32_t lim = left * nvec >= 32 ? 0xffffffffu : (1u << (left * nvec)) - 1u;
          const uint32_t* ct = blob + cw_tab_off + ((size_t)cw_v * partvals + sym) * max_stages;
          grp_acc |= VPZ_LDG(ct) & lim;
          const int u0 = gpart * nvec, c0 = u0 >> 5, sh = u0 & 31;
          for (int s2 = 1; s2 < max_stages; s2++) {
            const uint32_t m = VPZ_LDG(ct + s2) & lim;
            if (m) {
              smask[s2 * nchunks + c0] |= m << sh;
              if (sh && (m >> (32 - sh))) smask[s2 * nchunks + c0 + 1] |= m >> (32 - sh);
            }
          }
          cwmask &= cwmask - 1;
          if (!cwmask) {                     // the group's classwords are complete: its stage-0 units follow
            cur_mask = grp_acc;
            grp_acc = 0;
            ubase = u0;
          }
        }
        if (done) break;
      }
      trip++;
      if (rem > 0) {
        const int sym = k1_decode<DEBUG, SM>(b, cur, blob, P, nscal);
        if (sym < 0) {  // Residue0.cs:195-201: keep what was decoded
          status = 1;
          break;
        }
        k1a_emit(P, o, sym);
        rem--;
      }
    }
  }
}

// =============================================================================================
// K1a: one lane decodes one packet
// =============================================================================================
// staged_key: (setup slot * 2 + long block flag) whose tables the CTA holds in shared memory (SM variant only)
// ring: this thread's two 16-byte slots for the packet bytes (K1Bits)
template <bool DEBUG, bool FULL, bool SM = false>
VPZ_DEV void k1a_decode_packet(const K1Params& P, uint32_t pkt_idx, uint4* ring, uint32_t* cls, uint32_t staged_key = 0xffffffffu) {
  const VpzPktIn pk = P.pkts[pkt_idx];
  const uint32_t* blob = P.setups[pk.setup_slot];
  VPZ_ASSUME_GLOBAL(blob);
  const VpzSetupHdr* H = reinterpret_cast<const VpzSetupHdr*>(blob);
  const int C = H->channels;
  const VpzBook* books = reinterpret_cast<const VpzBook*>(blob + H->books_off);
  const int half_max = 1 << (H->log2_size1 - 1);
  uint32_t* rec = P.rec + pk.rec_off;

  K1Bits b;
  k1_bits_init(b, P.bytes, pk.byte_off >> 2, (int)pk.byte_len, ring);
  b.cls = cls;
  int nscal = 0, ncls = 0;

  // StreamDecoder.DecodeNextPacket (StreamDecoder.cs:728-741): the host only queues packets whose
  // type bit is 0 and whose mode exists, so these reads just advance the cursor.
  k1_read(b, P.bytes, 1);
  const int mode_idx = (int)k1_read(b, P.bytes, H->mode_bits);
  const VpzMode* modes = reinterpret_cast<const VpzMode*>(blob + H->modes_off);
  const int long_block = modes[mode_idx].block_flag;
  const VpzMapping* mp = reinterpret_cast<const VpzMapping*>(blob + H->mappings_off) + modes[mode_idx].mapping;
  if (long_block) k1_read(b, P.bytes, 2);  // prev/next window flags (Mode.cs:38), geometry is the host's job
  const int half = long_block ? half_max : (1 << (H->log2_size0 - 1));
  const uint32_t sm_on = SM && (pk.setup_slot * 2u + (uint32_t)long_block) == staged_key ? 1u : 0u;

  // ---- floor unpack + unwrap, channel by channel (Mapping.cs:106-116, Floor1.cs:162-219, 270-353) ----
  uint32_t own_mask = 0;  // bit ch: floor has energy (FloorData.ExecuteChannel)
  uint32_t ent_pos = pk.ent_off;   // multiple of 4 (engine.cpp)
  uint32_t ent_lo = 0, ent_hi = 0;
  for (int ch = 0; ch < C; ch++) {
    const VpzFloor1* fl = reinterpret_cast<const VpzFloor1*>(blob + H->floors_off) + mp->submap_floor[mp->mux[ch]];
    if (FULL && fl->floor_type == 0) {
      // Floor0.Unpack (Floor0.cs:115-167): amplitude, book number, then the coefficient vectors -- which the
      // reference reads WHATEVER the amplitude is.  The entry indices go to the packet's entry area (they
      // precede the residue's); K1b turns them into the LSP curve.
      uint32_t* seg = rec + K1_REC_HDR + ch * K1_SEG_WORDS;
      const uint32_t amp = k1_read(b, P.bytes, fl->f0.amp_bits);
      const uint32_t book_num = k1_read(b, P.bytes, fl->f0.book_bits);
      bool ok = book_num < fl->f0.nbooks;
      const uint32_t first = ent_pos - pk.ent_off;
      if (ok) {
        const int bi = fl->f0.books[book_num];
        const K1Book fb = k1_book(blob, books, bi);
        const int dims = k1_book_dims(books + bi);
        K1aOut o;
        o.ent_pos = ent_pos;
        o.ent_lo = ent_lo;
        o.ent_hi = ent_hi;
        for (int i = 0; i < (int)fl->f0.order; i += dims) {
          const int sym = k1_decode<DEBUG>(b, fb, blob, P, nscal);
          if (sym < 0) {
            ok = false;
            break;
          }
          k1a_emit(P, o, sym);
        }
        ent_pos = o.ent_pos;
        ent_lo = o.ent_lo;
        ent_hi = o.ent_hi;
      }
      // ExecuteChannel = Amp != 0 with Amp = amp * ampOfs / (2^ampBits - 1) (Floor0.cs:21,121-123)
      const bool exec = ok && amp != 0 && fl->f0.amp_ofs != 0;
      seg[0] = exec ? (0x80000000u | book_num) : 0u;
      seg[1] = amp;
      seg[2] = first;
      seg[3] = ent_pos - pk.ent_off - first;
      if (exec) own_mask |= 1u << ch;
      if (DEBUG && P.dbg.hdr) {
        P.dbg.hdr[DUMP_POSTCOUNT + ch] = exec ? 2 : 0;
        for (int i = 0; i < 64; i++) {
          const int v = !exec ? 0 : (i == 0 ? (int)(amp & 0x7fffffffu) : (i == 1 ? (int)book_num : 0));
          P.dbg.hdr[DUMP_RAWPOSTS + ch * 64 + i] = v;
          P.dbg.hdr[DUMP_FINALY + ch * 64 + i] = v;
          P.dbg.hdr[DUMP_STEPFLAGS + ch * 64 + i] = 0;
        }
      }
      continue;
    }
    // raw posts, unwrapped in place into the final Y (post i is read once, at step i, and only earlier
    // posts are looked at afterwards): one per-lane array in local memory instead of two
    short po[VPZ_MAX_POSTS + 1];
    short* const fy = po;
    int count = 0, written = 0;  // written: posts stored before a failed decode reset the count
    if (k1_read(b, P.bytes, 1) == 1) {
      const int ybits = fl->ybits;
      po[0] = (short)k1_read(b, P.bytes, ybits);
      po[1] = (short)k1_read(b, P.bytes, ybits);
      count = written = 2;
      const int nparts = fl->partitions;
      // book of a codeword: one 8-byte load from the floor's book table (VpzFloor1.fbook_tab_off)
      const uint2* ftab = reinterpret_cast<const uint2*>(blob + fl->fbook_tab_off);
      for (int i = 0; i < nparts && count > 0; i++) {
        const int c = fl->part_class[i];
        const int cdim = fl->class_dim[c], cbits = fl->class_sub[c];
        const uint32_t csub = (1u << cbits) - 1u;
        uint32_t cval = 0;
        if (cbits > 0) {
          const uint2 t = VPZ_LDG(ftab + c * 9);
          K1Book mb;
          mb.l1_off = t.x;
          mb.l1_mask = (1u << (t.y & 0xffu)) - 1u;
          mb.meta = (t.y & 0xffffu) | (k1a_soff<SM>(sm_on, (int)((t.y >> 8) & 0xffu)) << 16);
          int v = k1_decode<DEBUG, SM>(b, mb, blob, P, nscal);
          if (v < 0) {
            count = 0;
            break;
          }
          cval = (uint32_t)v;
        }
        for (int j = 0; j < cdim; j++) {
          const uint2 t = VPZ_LDG(ftab + c * 9 + 1 + (int)(cval & csub));
          cval >>= cbits;
          int post = 0;
          if (t.y != 0u) {
            K1Book sb;
            sb.l1_off = t.x;
            sb.l1_mask = (1u << (t.y & 0xffu)) - 1u;
            sb.meta = (t.y & 0xffffu) | (k1a_soff<SM>(sm_on, (int)((t.y >> 8) & 0xffu)) << 16);
            post = k1_decode<DEBUG, SM>(b, sb, blob, P, nscal);
            if (post < 0) {
              count = 0;
              break;
            }
          }
          po[count++] = (short)post;
          written = count;
        }
      }
    }
    if (DEBUG && P.dbg.hdr) {
      P.dbg.hdr[DUMP_POSTCOUNT + ch] = count;
      for (int i = 0; i < 64; i++) P.dbg.hdr[DUMP_RAWPOSTS + ch * 64 + i] = i < written ? po[i] : 0;
    }
    uint32_t* seg = rec + K1_REC_HDR + ch * K1_SEG_WORDS;
    if (count == 0) {
      seg[0] = 0;
      continue;
    }
    own_mask |= 1u << ch;
    // UnwrapPosts (Floor1.cs:270-353): serial dependency through earlier posts
    const int range = fl->range;
    unsigned long long flags = 3ull;
    for (int i = 2; i < count; i++) {
      const int lo = fl->lneigh[i], hi = fl->hneigh[i];
      const int predicted = k1_render_point(fl->xlist[lo], fy[lo], fl->xlist[hi], fy[hi], fl->xlist[i]);
      const int val = po[i];
      const int highroom = range - predicted, lowroom = predicted;
      const int room = (highroom < lowroom ? highroom : lowroom) * 2;
      int result = predicted;
      if (val != 0) {
        flags |= (1ull << lo) | (1ull << hi) | (1ull << i);
        if (val >= room)
          result = highroom > lowroom ? val - lowroom + predicted : predicted - val + highroom - 1;
        else
          result = (val & 1) ? predicted - ((val + 1) >> 1) : predicted + (val >> 1);
      }
      fy[i] = (short)result;
    }
    if (DEBUG && P.dbg.hdr) {
      for (int i = 0; i < 64; i++) {
        P.dbg.hdr[DUMP_FINALY + ch * 64 + i] = i < count ? fy[i] : 0;
        P.dbg.hdr[DUMP_STEPFLAGS + ch * 64 + i] = i < count ? (int)((flags >> i) & 1ull) : 0;
      }
    }
    // flagged posts in X order -> line segments (Floor1.Apply, Floor1.cs:222-268).  A segment that
    // is clamped at `half` (quirk Q1: clamp before the slope) is the last one: the loop breaks.
    const int mult = fl->multiplier;
    int nseg = 0, lx = 0, ly = fy[0] * mult;
    seg[1] = (uint32_t)0 | ((uint32_t)(ly & 0xffff) << 16);
    for (int i = 1; i < count; i++) {
      const int idx = fl->sortidx[i];
      if ((flags >> idx) & 1ull) {
        const int hx = fl->xlist[idx], hy = fy[idx] * mult;
        if (lx < half) {
          nseg++;
          seg[1 + nseg] = (uint32_t)(hx < half ? hx : half) | ((uint32_t)(hy & 0xffff) << 16);
        }
        lx = hx;
        ly = hy;
      }
      if (lx >= half) break;
    }
    if (lx < half) {  // flat tail
      nseg++;
      seg[1 + nseg] = (uint32_t)half | ((uint32_t)(ly & 0xffff) << 16);
    }
    seg[0] = (uint32_t)nseg;
  }
  // no-energy propagation through the coupling steps (Mapping.cs:121-130)
  uint32_t noexec = ~own_mask & ((1u << C) - 1u);
  for (int i = 0; i < mp->coupling_steps; i++) {
    uint32_t mb = 1u << mp->mag[i], ab = 1u << mp->ang[i];
    if (!((noexec & mb) && (noexec & ab))) noexec &= ~(mb | ab);
  }

  // ---- residue: classwords + VQ entry indices, submap after submap (Mapping.cs:136-163) ------------
  const uint32_t floor_ents = ent_pos - pk.ent_off;   // floor-0 coefficient entries (0 without floor 0)
  K1aOut o;
  o.ent_pos = ent_pos;
  o.ent_lo = ent_lo;
  o.ent_hi = ent_hi;
  o.status = 0;
  o.nscal = nscal;
  o.ncls = ncls;
  uint8_t* rec_cls = reinterpret_cast<uint8_t*>(P.rec + pk.rec_off + K1_REC_HDR + C * K1_SEG_WORDS);
  const VpzResidue* residues = reinterpret_cast<const VpzResidue*>(blob + H->residues_off);
  if (!FULL || mp->submaps == 1) {
    k1a_residue<DEBUG, SM>(P, blob, books, b, residues + mp->submap_residue[0], C, noexec, half, rec_cls, o, sm_on);
  } else {
    // where a submap's entry indices end is only known here: a residue that stops early (no matching
    // code, quirk Q8) does not stop the submaps after it, which read on from the same bit position.  The
    // end positions go behind the class bytes of the last submap.
    int cls_total = 0;
    for (int sm = 0; sm < mp->submaps; sm++) {
      int nch = 0;
      for (int ch = 0; ch < C; ch++) nch += mp->mux[ch] == sm;
      if (nch == 0) continue;
      const K1ResGeom g = k1_res_geom(residues + mp->submap_residue[sm], nch, half, 0u);
      cls_total += (g.part_count * g.nvec + 3) >> 2;
    }
    uint32_t* sub_end = reinterpret_cast<uint32_t*>(rec_cls) + cls_total;
    for (int sm = 0; sm < mp->submaps; sm++) {
      // the channels of this submap, in channel order, and their flags (Mapping.cs:138-146)
      int nch = 0;
      uint32_t flags = 0;
      for (int ch = 0; ch < C; ch++)
        if (mp->mux[ch] == sm) {
          flags |= ((noexec >> ch) & 1u) << nch;
          nch++;
        }
      if (nch > 0) {   // Residue0.Decode with no channels reads nothing
        const VpzResidue* rs = residues + mp->submap_residue[sm];
        const K1ResGeom g = k1_res_geom(rs, nch, half, flags);
        k1a_residue<DEBUG, SM>(P, blob, books, b, rs, nch, flags, half, rec_cls, o, sm_on);
        rec_cls += ((g.part_count * g.nvec + 3) >> 2) << 2;
      }
      sub_end[sm] = o.ent_pos - pk.ent_off;
    }
  }
  const int status = o.status;
  ent_pos = o.ent_pos;
  ent_lo = o.ent_lo;
  ent_hi = o.ent_hi;
  nscal = o.nscal;
  ncls = o.ncls;
  rec[0] = own_mask | (noexec << 8) | ((uint32_t)status << 16) | ((uint32_t)long_block << 24);
  k1a_emit_flush(P, ent_pos, ent_lo, ent_hi);
  rec[1] = ent_pos - pk.ent_off;
  rec[2] = (uint32_t)b.pos;
  rec[3] = (uint32_t)modes[mode_idx].mapping | ((uint32_t)mp->submap_residue[0] << 8) | (floor_ents << 16);
  if (DEBUG && P.dbg.hdr) {
    int32_t* h = P.dbg.hdr;
    h[DUMP_STATUS] = 0;
    h[DUMP_MODE] = mode_idx;
    h[DUMP_BLOCK] = half * 2;
    h[DUMP_BITS] = b.pos;
    h[DUMP_EXEC] = (int)own_mask;
    h[DUMP_NOEXEC] = (int)noexec;
    h[DUMP_SCALARS_N] = nscal;
    h[DUMP_CLASSES_N] = ncls;
  }
}

// ---------------------------------------------------------------------------------------------
// K1a with the first-level Huffman tables in SHARED memory (vpz_k1a_symbols_sm; the north-star's "shared-memory
// lookahead table").  A CTA of K1A_SM_THREADS threads takes K1A_SM_THREADS consecutive packets of the grouped
// order at a time; the tables of the (setup, block size) the range starts with are staged once (VpzSetupHdr.
// k1a_stage_off, <= 64 KB) and stay until a range starts with another group.  Lanes whose packet belongs to
// another group (the few at a group boundary) keep the global-memory path.
// ---------------------------------------------------------------------------------------------
#define K1A_SM_THREADS 512
VPZ_DEV void k1a_sm_loop(const K1Params& P, uint32_t* s_ctl) {
  const uint32_t tid = threadIdx.x;
  uint32_t staged = 0xffffffffu;   // the same in every thread of the CTA
  for (;;) {
    __syncthreads();               // every warp has left the previous range: its tables may go
    if (tid == 0) {
      const uint32_t base = atomicAdd(P.counter, (uint32_t)K1A_SM_THREADS);
      uint32_t key = 0xffffffffu;
      if (base < P.n_pkts) {
        const VpzPktIn pk = P.pkts[P.order ? P.order[base] : base];
        const uint32_t* blob = P.setups[pk.setup_slot];
        const VpzSetupHdr* H = reinterpret_cast<const VpzSetupHdr*>(blob);
        // bit 0 of an audio packet is its type, the mode number follows (StreamDecoder.cs:728-741)
        const uint32_t mode = (P.bytes[pk.byte_off >> 2] >> 1) & ((1u << H->mode_bits) - 1u);
        const VpzMode* modes = reinterpret_cast<const VpzMode*>(blob + H->modes_off);
        key = pk.setup_slot * 2u + ((mode < H->nmodes && modes[mode].block_flag) ? 1u : 0u);
      }
      s_ctl[0] = base;
      s_ctl[1] = key;
    }
    __syncthreads();
    const uint32_t base = s_ctl[0];
    if (base >= P.n_pkts) break;
    const uint32_t key = s_ctl[1];
    if (key != staged) {
      const uint32_t* blob = P.setups[key >> 1];
      const VpzSetupHdr* H = reinterpret_cast<const VpzSetupHdr*>(blob);
      const uint32_t* plan = blob + H->k1a_stage_off[key & 1u];
      const uint32_t n = VPZ_LDG(plan);
      for (uint32_t i = 0; i < n; i++) {
        const uint32_t* src = blob + VPZ_LDG(plan + 2 + 3 * i);
        const uint32_t words = VPZ_LDG(plan + 3 + 3 * i);
        uint32_t* dst = K1A_SM + VPZ_LDG(plan + 4 + 3 * i);
        for (uint32_t w = tid; w < words; w += K1A_SM_THREADS) dst[w] = VPZ_LDG(src + w);
      }
      if (tid < 128) K1A_SM[K1A_SM_WORDS + tid] = VPZ_LDG(blob + H->k1a_soff_off[key & 1u] + tid);
      staged = key;
      __syncthreads();
    }
    const uint32_t i = base + tid;
    if (i < P.n_pkts)
      k1a_decode_packet<false, false, true>(P, P.order ? P.order[i] : i, reinterpret_cast<uint4*>(K1A_SM + K1A_SM_WORDS + 128) + 2 * tid, nullptr, staged);
  }
}

// =============================================================================================
// K1b: one warp turns one symbol record into a spectrum
// =============================================================================================
// swizzled shared index: one pad word per 32 keeps lanes that own consecutive partitions (a
// partition is 16 or 32 floats) on different banks
VPZ_DEV int k1b_sw(int i) { return i + (i >> 5); }

// GENERAL path (any channel count, several submaps, residue type 0, dimensions that do not divide the
// partition, floor 0): one CTA of K1B_THREADS threads per packet.  Shared memory (32-bit words):
// res[sw(C * half_max)] then ustart[K1_MAX_UNITS + 32] (reused for the floor segments / LSP coefficients),
// uinfo[K1_MAX_UNITS], uvq[K1_MAX_UNITS], then 8 words of scan scratch
template <bool DEBUG>
VPZ_DEV void k1b_build_packet_general(const K1Params& P, uint32_t pkt_idx, uint32_t* smem, int tid) {
  const int lane = tid & 31, wid = tid >> 5;
  const VpzPktIn pk = P.pkts[pkt_idx];
  const uint32_t* blob = P.setups[pk.setup_slot];
  VPZ_ASSUME_GLOBAL(blob);
  const VpzSetupHdr* H = reinterpret_cast<const VpzSetupHdr*>(blob);
  const int C = H->channels;
  const VpzBook* books = reinterpret_cast<const VpzBook*>(blob + H->books_off);
  const int half_max = 1 << (H->log2_size1 - 1);
  const uint32_t* rec = P.rec + pk.rec_off;
  const uint16_t* ent = P.ent + pk.ent_off;
  const uint32_t hdr = rec[0];
  const uint32_t own_mask = hdr & 0xffu, noexec = (hdr >> 8) & 0xffu;
  const int long_block = (int)(hdr >> 24) & 1;
  const uint32_t n_ent_all = rec[1];
  const int half = long_block ? half_max : (1 << (H->log2_size0 - 1));

  float* res = reinterpret_cast<float*>(smem);
  int* ustart = reinterpret_cast<int*>(smem + k1b_sw(C * half_max) + 1);
  uint32_t* uinfo = reinterpret_cast<uint32_t*>(ustart + K1_MAX_UNITS + 32);
  uint32_t* uvq = uinfo + K1_MAX_UNITS;
  int* scan = reinterpret_cast<int*>(uvq + K1_MAX_UNITS);

  if (tid == 0) {
    VpzPktRes r;
    r.exec_mask = (uint8_t)own_mask;
    r.status = (uint8_t)((hdr >> 16) & 0xffu);
    r.end16[0] = r.end16[1] = 255;   // this path writes every bin
    P.res[pkt_idx] = r;
  }
  if (own_mask == 0 && !(DEBUG && P.dbg.residue)) return;  // every channel silent: K3 outputs zeros

  // mapping index and the number of floor-0 entry indices come with the record (K1a found them)
  const VpzMapping* mp = reinterpret_cast<const VpzMapping*>(blob + H->mappings_off) + (rec[3] & 0xffu);
  const VpzResidue* residues = reinterpret_cast<const VpzResidue*>(blob + H->residues_off);
  const int nsub = mp->submaps;
  const bool multi = nsub > 1;
  float* out = P.spec + pk.spec_off;

  const int total = C * half;
  for (int i = tid; i < total; i += K1B_THREADS) res[k1b_sw(i)] = 0.f;
  __syncthreads();

  // ---- residue, submap after submap (Mapping.cs:136-163).  The reference decodes every submap into ONE
  // scratch buffer that is zeroed once per packet (Mapping.cs:133): channel slot c of a later submap
  // starts from what slot c of the earlier submaps left behind (types 0 / 1 accumulate; type 2 overwrites).
  // res[] is that scratch (slot stride = half); with several submaps the slots are copied out to the
  // channels' spectrum buffers after every submap and fetched back, channel-major, at the end.
  const uint8_t* rec_cls = reinterpret_cast<const uint8_t*>(rec + K1_REC_HDR + C * K1_SEG_WORDS);
  uint32_t stage_base = rec[3] >> 16;   // the residue's entry indices follow the floor-0 ones
  // several submaps: K1a left the end of every submap's entry indices behind the class bytes (a residue
  // that stops early does not stop the submaps after it)
  const uint32_t* sub_end = nullptr;
  if (multi) {
    int cls_total = 0;
    for (int sm = 0; sm < nsub; sm++) {
      int nch = 0;
      for (int ch = 0; ch < C; ch++) nch += mp->mux[ch] == sm;
      if (nch == 0) continue;
      const K1ResGeom g = k1_res_geom(residues + mp->submap_residue[sm], nch, half, 0u);
      cls_total += (g.part_count * g.nvec + 3) >> 2;
    }
    sub_end = rec + K1_REC_HDR + C * K1_SEG_WORDS + cls_total;
  }
  int rtype_single = 0;
  for (int sm = 0; sm < nsub; sm++) {
    const uint32_t n_ent = multi ? sub_end[sm] : n_ent_all;
    int nch = 0;
    uint32_t flags = 0, chans = 0;   // chans: 4 bits per slot = the channel it belongs to
    for (int ch = 0; ch < C; ch++)
      if (!multi || mp->mux[ch] == sm) {
        flags |= ((noexec >> ch) & 1u) << nch;
        chans |= (uint32_t)ch << (4 * nch);
        nch++;
      }
    if (nch == 0) continue;
    const VpzResidue* rs = residues + mp->submap_residue[sm];
    const K1ResGeom g = k1_res_geom(rs, nch, half, flags);
    rtype_single = g.rtype;
    if (multi && g.rtype == 2) {
      // Residue2.Decode (Residue2.cs:12-52): all flagged -> the slots are cleared; else a FRESH zeroed vector
      // is decoded and de-interleaved over the slots (both overwrite slots 0 .. nch-1 completely)
      for (int i = tid; i < nch * half; i += K1B_THREADS) res[k1b_sw(i)] = 0.f;
      __syncthreads();
    }
    const int nunits = g.part_count * g.nvec;       // decode order inside a stage: partition major
    // ---- per stage: unit scan, then entry-parallel accumulate --------------------------------------
    if (g.part_count > 0 && g.any && stage_base < n_ent) {
      for (int stage = 0; stage < rs->max_stages && stage_base < n_ent; stage++) {
        // entries per unit -> exclusive prefix sum = where each unit's entries start
        uint32_t carry = 0;
        for (int u0 = 0; u0 < nunits; u0 += K1B_THREADS) {
          const int u = u0 + tid;
          int cnt = 0;
          uint32_t info = 0, vqoff = 0;   // info: dims | first bin << 8
          if (u < nunits) {
            const int part = u / g.nvec, v = u - part * g.nvec;
            if (!((g.skip >> v) & 1u)) {
              const int c = rec_cls[u];
              if (((rs->cascade[c] >> stage) & 1u) && rs->has_books[c]) {
                const VpzBook* bk = books + rs->books[c][stage];
                const int dims = k1_book_dims(bk);
                cnt = k1_unit_entries(g.rtype, g.psize, dims);
                vqoff = VPZ_LDG(&bk->vq_off);
                info = (uint32_t)(dims & 0xff) | ((uint32_t)(g.begin + part * g.psize + v * half) << 8);   // rtype 2: v == 0
              }
            }
          }
          int incl = cnt;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
          }
          if (lane == 31) scan[wid] = incl;
          __syncthreads();
          int before = 0, round_total = 0;
#pragma unroll
          for (int w = 0; w < K1B_THREADS / 32; w++) {
            const int t = scan[w];
            if (w < wid) before += t;
            round_total += t;
          }
          if (u < nunits) {
            ustart[u] = (int)carry + before + incl - cnt;
            uinfo[u] = info;
            uvq[u] = vqoff;
          }
          carry += (uint32_t)round_total;
          __syncthreads();
        }
        // ---- every thread takes single ENTRIES (codewords) of the stage, not whole units: the work is
        // spread evenly however the active units are distributed.  The unit of entry e is the last one
        // whose start is <= e (idle units have zero length and sort before the active unit that shares
        // their start).  A bin is touched once per stage and the stages run in order, so the sums
        // round exactly like the reference's stage-major accumulation.
        uint32_t stage_n = carry;
        if (stage_base + stage_n > n_ent) stage_n = n_ent - stage_base;   // truncated packet: keep what was decoded
        const uint16_t* ep = ent + stage_base;
        for (uint32_t e = (uint32_t)tid; e < stage_n; e += K1B_THREADS) {
          int lo = 0, hi = nunits;   // ustart[lo] <= e < ustart[hi] (virtual ustart[nunits] = +inf)
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if ((uint32_t)ustart[mid] <= e) lo = mid; else hi = mid;
          }
          const uint32_t info = uinfo[lo];
          const int dims = (int)(info & 0xffu);
          const int si = (int)(e - (uint32_t)ustart[lo]);        // index of the entry inside its unit
          const int dest = (int)(info >> 8);                     // first bin of the unit
          const float* lk = reinterpret_cast<const float*>(blob + uvq[lo]) + (size_t)ep[e] * dims;
          if (g.rtype == 0) {
            // Residue0.WriteVectors (Residue0.cs:208-231), quirk Q6: dims summed into one bin
            float r = 0.f;
            for (int d = 0; d < dims; d++) r = __fadd_rn(r, VPZ_LDG(lk + d));
            const int at = k1b_sw(dest + si);
            res[at] = __fadd_rn(res[at], r);
          } else if (dims == 2) {
            // Residue1.WriteVectors (Residue1.cs:12-34)
            const float2 v2 = VPZ_LDG(reinterpret_cast<const float2*>(lk));
            const int o = dest + si * 2;
            const int a0 = k1b_sw(o), a1 = k1b_sw(o + 1);
            res[a0] = __fadd_rn(res[a0], v2.x);
            res[a1] = __fadd_rn(res[a1], v2.y);
          } else if (dims == 4 || dims == 8) {
            const int o = dest + si * dims;
            for (int d = 0; d < dims; d += 4) {
              const float4 v4 = VPZ_LDG(reinterpret_cast<const float4*>(lk + d));
              const int a0 = k1b_sw(o + d), a1 = k1b_sw(o + d + 1), a2 = k1b_sw(o + d + 2), a3 = k1b_sw(o + d + 3);
              res[a0] = __fadd_rn(res[a0], v4.x);
              res[a1] = __fadd_rn(res[a1], v4.y);
              res[a2] = __fadd_rn(res[a2], v4.z);
              res[a3] = __fadd_rn(res[a3], v4.w);
            }
          } else {
            // any other dimension; the last entry of a unit may run past the partition when dims does
            // not divide it, and the reference then keeps writing (bounded by the vector length)
            const int vend = g.rtype == 2 ? g.vlen : (dest / half) * half + half;
            for (int d = 0; d < dims; d++) {
              const int o = dest + si * dims + d;
              if (o < vend) {
                const int at = k1b_sw(o);
                res[at] = __fadd_rn(res[at], VPZ_LDG(lk + d));
              }
            }
          }
        }
        stage_base += carry;
        __syncthreads();
      }
    }
    rec_cls += ((nunits + 3) >> 2) << 2;
    if (multi) stage_base = n_ent;   // the next submap's entry indices start where this one's really end
    if (multi) {
      // Mapping.cs:150-160: slot c -> channel buffer of the c-th channel of the submap, always (also when
      // nothing was decoded: the slot then still holds what earlier submaps left)
      __syncthreads();
      const bool inter = g.rtype == 2 && g.any && nch > 1;   // the interleaved type 2 vector sits in [0, nch * half)
      for (int c = 0; c < nch; c++) {
        const int ch = (int)((chans >> (4 * c)) & 15u);
        for (int i = tid; i < half; i += K1B_THREADS) out[ch * half + i] = res[k1b_sw(inter ? i * nch + c : c * half + i)];
      }
      __syncthreads();
      if (inter) {   // the slots now hold the de-interleaved channels (Residue2.cs:42-50) for the submaps to come
        for (int c = 0; c < nch; c++) {
          const int ch = (int)((chans >> (4 * c)) & 15u);
          for (int i = tid; i < half; i += K1B_THREADS) res[k1b_sw(c * half + i)] = out[ch * half + i];
        }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  if (multi) {   // every channel belongs to exactly one submap: fetch them back, channel-major
    for (int i = tid; i < total; i += K1B_THREADS) res[k1b_sw(i)] = out[i];
    __syncthreads();
  }
  const bool interleaved = !multi && rtype_single == 2;

  // accessor of channel c, bin i after the Residue2 de-interleave (Residue2.cs:42-50)
#define RES_AT(c, i) res[k1b_sw(interleaved ? (i) * C + (c) : (c) * half + (i))]

  if (DEBUG && P.dbg.residue) {
    for (int c = 0; c < C; c++)
      for (int i = tid; i < half; i += K1B_THREADS) P.dbg.residue[c * half + i] = RES_AT(c, i);
  }

  // ---- inverse coupling, last step first (Mapping.cs:166-172, 235-267) ----------------------
  for (int s = mp->coupling_steps - 1; s >= 0; s--) {
    const int cm = mp->mag[s], ca = mp->ang[s];
    for (int i = tid; i < half; i += K1B_THREADS) {
      float m = RES_AT(cm, i), a = RES_AT(ca, i);
      float nm = m, na = m;
      if (m > 0.f) {
        if (a > 0.f) na = __fsub_rn(m, a); else nm = __fadd_rn(m, a);
      } else {
        if (a > 0.f) na = __fadd_rn(m, a); else nm = __fsub_rn(m, a);
      }
      RES_AT(cm, i) = nm;
      RES_AT(ca, i) = na;
    }
    __syncthreads();
  }

  // ---- floor curve + multiply + store ------------------------------------------------------------
  // Floor 1 (Floor1.cs:222-268, 372-397): RenderLineMulti in closed form: after k steps of the DDA
  // y = y0 + sy * floor(k * |dy| / adx) (base*k + sy*floor(k*rem/adx) with |dy| = |base|*adx + rem).  This
  // path is the rare one: a plain integer division per bin.  Lanes walk the bins K1B_THREADS at a time;
  // every thread keeps its own (monotone) segment cursor.
  // Floor 0 (Floor0.cs:169-224): the LSP curve, one value per bark band, evaluated per bin.
  const float* db = reinterpret_cast<const float*>(blob + H->db_off);
  uint32_t* sg = reinterpret_cast<uint32_t*>(ustart);   // per segment: x0 | x1 << 16, y0 | (|dy| << 16), sign
  for (int ch = 0; ch < C; ch++) {
    if (!((own_mask >> ch) & 1u)) continue;  // Mapping.cs:185-194: silent channel, K3 sees zeros
    const uint32_t* seg = rec + K1_REC_HDR + ch * K1_SEG_WORDS;
    if (seg[0] & 0x80000000u) {
      const VpzFloor1* fl = reinterpret_cast<const VpzFloor1*>(blob + H->floors_off) + mp->submap_floor[mp->mux[ch]];
      const int order = fl->f0.order;
      float* coeff = reinterpret_cast<float*>(ustart);
      __syncthreads();
      if (tid == 0) {
        // Floor0.Unpack, second half (Floor0.cs:138-166): lookup vectors, the "averaging", then 2 cos (:186-189)
        const VpzBook* bk = books + fl->f0.books[seg[0] & 0xffu];
        const int dims = k1_book_dims(bk);
        const float* vq = reinterpret_cast<const float*>(blob + VPZ_LDG(&bk->vq_off));
        const uint16_t* fe = ent + seg[2];
        for (int i = 0, e = 0; i < order; e++) {
          const float* lk = vq + (size_t)fe[e] * dims;
          for (int j = 0; i < order && j < dims; j++, i++) coeff[i] = VPZ_LDG(lk + j);
        }
        float last = 0.f;
        for (int j = 0; j < order;) {
          for (int k = 0; j < order && k < dims; j++, k++) coeff[j] = __fadd_rn(coeff[j], last);
          last = coeff[j - 1];
        }
        for (int j = 0; j < order; j++) coeff[j] = __fmul_rn(2.f, cosf(coeff[j]));
      }
      __syncthreads();
      // Amp = (float)(amp * ampOfs / (double)((1 << ampBits) - 1)) (Floor0.cs:121-123; int shift, wrapping)
      const double amp_div = (double)(int32_t)((1u << (fl->f0.amp_bits & 31)) - 1u);
      const float amp = (float)((double)((uint64_t)seg[1] * (uint64_t)fl->f0.amp_ofs) / amp_div);
      const float amp_ofs = (float)fl->f0.amp_ofs;
      const int w = long_block ? 1 : 0;   // BlockSizes.IndexOf: equal sizes resolve to index 0, whose tables are identical
      const uint16_t* bark = reinterpret_cast<const uint16_t*>(blob + fl->f0.bark_off[w]);
      const float* wmap = reinterpret_cast<const float*>(blob + fl->f0.wmap_off[w]);
      for (int x = tid; x < half; x += K1B_THREADS) {
        const float wv = VPZ_LDG(wmap + VPZ_LDG(bark + x));
        float p = .5f, q = .5f;
        int j;
        for (j = 1; j < order; j += 2) {
          q = __fmul_rn(q, __fsub_rn(wv, coeff[j - 1]));
          p = __fmul_rn(p, __fsub_rn(wv, coeff[j]));
        }
        if (j == order) {   // odd order filter; slightly asymmetric
          q = __fmul_rn(q, __fsub_rn(wv, coeff[j - 1]));
          p = __fmul_rn(p, __fmul_rn(p, __fsub_rn(4.f, __fmul_rn(wv, wv))));
          q = __fmul_rn(q, q);
        } else {            // even order filter; still symmetric
          p = __fmul_rn(p, __fmul_rn(p, __fsub_rn(2.f, wv)));
          q = __fmul_rn(q, __fmul_rn(q, __fadd_rn(2.f, wv)));
        }
        q = __fsub_rn(amp / sqrtf(__fadd_rn(p, q)), amp_ofs);
        q = expf(__fmul_rn(q, 0.11512925f));
        out[ch * half + x] = __fmul_rn(RES_AT(ch, x), q);
      }
      continue;
    }
    const int nseg = (int)seg[0];
    __syncthreads();
    for (int s = tid; s < nseg; s += K1B_THREADS) {
      const uint32_t p0 = seg[1 + s], p1 = seg[2 + s];
      const int x0 = (int)(p0 & 0xffffu), y0 = (int)(short)(p0 >> 16);
      const int x1 = (int)(p1 & 0xffffu), y1 = (int)(short)(p1 >> 16);
      const int dy = y1 - y0;
      sg[4 * s] = (uint32_t)x0 | ((uint32_t)x1 << 16);
      sg[4 * s + 1] = (uint32_t)(y0 & 0xffff) | ((uint32_t)(dy < 0 ? -dy : dy) << 16);
      sg[4 * s + 2] = dy < 0 ? 1u : 0u;
    }
    __syncthreads();
    if (nseg == 0) continue;
    int si = 0;
    uint32_t w0 = sg[0], w1 = sg[1], w2 = sg[2];
    const int xend = (int)(sg[4 * (nseg - 1)] >> 16);   // the last segment ends at `half` (or where the floor ends)
    for (int x = tid; x < xend; x += K1B_THREADS) {
      while (x >= (int)(w0 >> 16) && si + 1 < nseg) {
        si++;
        w0 = sg[4 * si];
        w1 = sg[4 * si + 1];
        w2 = sg[4 * si + 2];
      }
      const int x0 = (int)(w0 & 0xffffu), adx = (int)(w0 >> 16) - x0;
      const int y0 = (int)(short)(w1 & 0xffffu), ady = (int)(w1 >> 16);
      const int q = adx > 0 ? ((x - x0) * ady) / adx : 0;
      int y = w2 ? y0 - q : y0 + q;
      y = y < 0 ? 0 : (y > 255 ? 255 : y);  // the reference reads the table unchecked (quirk Q2)
      out[ch * half + x] = __fmul_rn(RES_AT(ch, x), VPZ_LDG(db + y));
    }
  }
#undef RES_AT
  __syncthreads();
}

// =============================================================================================
// K1b, gather path (mono / stereo, residue 1 / 2, partitions aligned to 8 / 16 positions, power-of-two
// VQ dimensions): no residue buffer at all.  Every thread owns 8 CONSECUTIVE FREQUENCY BINS (for the
// interleaved type 2 vector of a stereo stream: the 16 positions 2x .. 2x+15, both channels).  Such
// a chunk lies inside one partition, so per stage ONE lookup finds the unit, its first entry and its
// VQ table; the 16/dims entries that cover the chunk are fetched and summed into registers in stage
// order -- the same rounded sums as the reference's stage-major accumulation into a zeroed buffer.
// Inverse coupling, floor multiply and two 128-bit stores per channel follow.  The floor curve is
// rendered beforehand by an exact integer DDA (RenderLineMulti, Floor1.cs:372-397), 16 bins per
// thread, into one byte per bin of shared memory.
//   per-warp words: urec[stages][U][2] | per channel ybuf[half_max/4] | sg[C][P.seg_stride]
// =============================================================================================

struct K1Gather {
  const uint32_t* urec;       // [stage][U][2]: (log2 dims + 1) << 28 | first entry (absolute; 0: idle unit), vq word offset
  const uint32_t* blob;
  const uint16_t* ent;
  uint32_t n_ent;
  int max_stages, U, nvec, begin, psize, pshift, span;   // span = part_count * psize
};

// acc[0..CH) += the VQ components of positions p .. p+CH-1 of vector v, stage after stage.
// CH = 16 (stereo type 2: 8 bins of both channels) or 8; p - begin is a multiple of CH.
// PAIR (CH = 16, stereo type 2): position i of the interleaved vector is channel i & 1, bin i >> 1; the
// accumulators are kept de-interleaved (channel 0 in the first eight, channel 1 in the last eight): Residue2.cs:42-50
#define K1G_AT(i) (PAIR ? (((i) & 1) * 8 + ((i) >> 1)) : (i))
template <int CH, bool PAIR>
VPZ_DEV void k1g_fetch_chunk(const K1Gather& G, int v, int p, float* acc) {
  const int rel = p - G.begin;
  if (rel < 0 || rel >= G.span) return;
  const int part = G.pshift >= 0 ? (rel >> G.pshift) : rel / G.psize;
  const int off = rel - part * G.psize;
  const int u = part * G.nvec + v;
  for (int s = 0; s < G.max_stages; s++) {
    const uint2 R = *reinterpret_cast<const uint2*>(G.urec + (size_t)(s * G.U + u) * 2);
    if (!R.x) continue;
    const int dsh = (int)(R.x >> 28) - 1;
    const uint32_t e0 = (R.x & 0x0fffffffu) + (uint32_t)(off >> dsh);
    const float* vq = reinterpret_cast<const float*>(G.blob + R.y);
    const uint16_t* ep = G.ent + e0;
    const uint32_t left = e0 < G.n_ent ? G.n_ent - e0 : 0u;   // truncated packet: keep what was decoded
    if (dsh == 1) {
#pragma unroll
      for (int k = 0; k < CH / 2; k++)
        if ((uint32_t)k < left) {
          const float2 t = VPZ_LDG(reinterpret_cast<const float2*>(vq + (size_t)ep[k] * 2));
          acc[K1G_AT(2 * k)] = __fadd_rn(acc[K1G_AT(2 * k)], t.x);
          acc[K1G_AT(2 * k + 1)] = __fadd_rn(acc[K1G_AT(2 * k + 1)], t.y);
        }
    } else if (dsh == 2) {
#pragma unroll
      for (int k = 0; k < CH / 4; k++)
        if ((uint32_t)k < left) {
          const float4 t = VPZ_LDG(reinterpret_cast<const float4*>(vq + (size_t)ep[k] * 4));
          acc[K1G_AT(4 * k)] = __fadd_rn(acc[K1G_AT(4 * k)], t.x);
          acc[K1G_AT(4 * k + 1)] = __fadd_rn(acc[K1G_AT(4 * k + 1)], t.y);
          acc[K1G_AT(4 * k + 2)] = __fadd_rn(acc[K1G_AT(4 * k + 2)], t.z);
          acc[K1G_AT(4 * k + 3)] = __fadd_rn(acc[K1G_AT(4 * k + 3)], t.w);
        }
    } else if (dsh == 3) {
#pragma unroll
      for (int k = 0; k < CH / 8; k++)
        if ((uint32_t)k < left) {
          const float* lk = vq + (size_t)ep[k] * 8;
#pragma unroll
          for (int d = 0; d < 8; d += 4) {
            const float4 t = VPZ_LDG(reinterpret_cast<const float4*>(lk + d));
            acc[K1G_AT(8 * k + d)] = __fadd_rn(acc[K1G_AT(8 * k + d)], t.x);
            acc[K1G_AT(8 * k + d + 1)] = __fadd_rn(acc[K1G_AT(8 * k + d + 1)], t.y);
            acc[K1G_AT(8 * k + d + 2)] = __fadd_rn(acc[K1G_AT(8 * k + d + 2)], t.z);
            acc[K1G_AT(8 * k + d + 3)] = __fadd_rn(acc[K1G_AT(8 * k + d + 3)], t.w);
          }
        }
    } else if (dsh == 0) {
#pragma unroll
      for (int k = 0; k < CH; k++)
        if ((uint32_t)k < left) acc[K1G_AT(k)] = __fadd_rn(acc[K1G_AT(k)], VPZ_LDG(vq + ep[k]));
    } else if (CH == 16 && dsh == 4) {
      if (left > 0) {
        const float* lk = vq + (size_t)ep[0] * 16;
#pragma unroll
        for (int d = 0; d < 16; d += 4) {
          const float4 t = VPZ_LDG(reinterpret_cast<const float4*>(lk + d));
          acc[K1G_AT(d)] = __fadd_rn(acc[K1G_AT(d)], t.x);
          acc[K1G_AT(d + 1)] = __fadd_rn(acc[K1G_AT(d + 1)], t.y);
          acc[K1G_AT(d + 2)] = __fadd_rn(acc[K1G_AT(d + 2)], t.z);
          acc[K1G_AT(d + 3)] = __fadd_rn(acc[K1G_AT(d + 3)], t.w);
        }
      }
    }
  }
}

// Inverse square-polar coupling of one (magnitude, angle) pair (Mapping.cs:235-267)
VPZ_DEV void k1_uncouple(float& m, float& a) {
  // the four cases of the reference as selects: with t = (m > 0 ? a : -a) the new angle is m - t when
  // a > 0 and the new magnitude m + t otherwise (x - y and x + (-y) round identically)
  const float t = m > 0.f ? a : -a;
  const float d = __fsub_rn(m, t), sm = __fadd_rn(m, t);
  const bool p = a > 0.f;
  a = p ? d : m;
  m = p ? m : sm;
}

// Floor1.RenderLineMulti (Floor1.cs:372-397), one byte per bin, exactly the reference's integer DDA:
//   y(x0 + k) = y0 + sy * (k * base + floor(k * rem / adx)).
// The curve is cut into PIECES of at most 16 bins that never cross a post: piece p of a channel belongs
// to the segment whose piece prefix (phase A) is the largest one <= p.  Every lane renders whole pieces,
// so no lane ever switches segments: the lanes of a warp stay convergent, and four bins at a time come
// from the remainder at the start of the group without a carried dependency (remainders e + j * rem,
// quotients by multiply-high with M = ceil(2^32 / adx): exact while dividend * adx < 2^32, i.e. for
// e + 4 rem < 5 adx <= 5 * 2^12; the state at the start of a piece, dividend up to adx^2, may come out
// one too high and is fixed up).
// sg: per segment {x0 | x1 << 16, y0 | |base| << 16, rem | piece prefix << 16 | sign << 31, M}.
VPZ_DEV uint32_t k1b_ybyte(int y) { return (uint32_t)(y < 0 ? 0 : (y > 255 ? 255 : y)); }   // the reference reads the table unchecked (quirk Q2)
#ifndef VPZ_EMU
#define K1B_MULHI(a, b) __umulhi(a, b)
#else
#define K1B_MULHI(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
#endif
VPZ_DEV void k1b_render_floor(const uint32_t* sg, int nseg, int npieces, uint8_t* yb, int x_end, int tid) {
  for (int p = tid; p < npieces; p += 32) {
    int lo = 0, hi = nseg;                                   // prefix[lo] <= p < prefix[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((int)((sg[4 * mid + 2] >> 16) & 0x7fffu) <= p) lo = mid; else hi = mid;
    }
    const uint4 w = *reinterpret_cast<const uint4*>(sg + 4 * lo);
    const int x0 = (int)(w.x & 0xffffu), x1 = (int)(w.x >> 16);
    const int adx = x1 - x0;
    const int rem = (int)(w.z & 0xffffu), sy = (w.z >> 31) ? -1 : 1;
    const int ystep = sy * (int)(w.y >> 16);
    const uint32_t magic = w.w;
    const int k = 16 * (p - (int)((w.z >> 16) & 0x7fffu));   // steps into the segment
    const int xs = x0 + k;
    if (xs >= x_end) break;                                  // pieces are ordered by x: nothing is coded from here on
    int len = (x1 < x_end ? x1 : x_end) - xs;                // bins of this piece
    len = len > 16 ? 16 : len;
    const int t = k * rem;
    int q = (int)K1B_MULHI((uint32_t)t, magic);
    int err = t - q * adx;
    if (err < 0) {
      q--;
      err += adx;
    }
    int y = (int)(short)(w.y & 0xffffu) + k * ystep + sy * q;
    uint8_t* out = yb + xs;
    int left = len;
#pragma unroll 1
    for (; left >= 4; left -= 4, out += 4) {
      // four bins and the state after them, each from the remainder at the start of the group
      const int n1 = err + rem, n2 = n1 + rem, n3 = n2 + rem, n4 = n3 + rem;
      const int q1 = (int)K1B_MULHI((uint32_t)n1, magic), q2 = (int)K1B_MULHI((uint32_t)n2, magic);
      const int q3 = (int)K1B_MULHI((uint32_t)n3, magic), q4 = (int)K1B_MULHI((uint32_t)n4, magic);
      out[0] = (uint8_t)k1b_ybyte(y);
      out[1] = (uint8_t)k1b_ybyte(y + ystep + sy * q1);
      out[2] = (uint8_t)k1b_ybyte(y + 2 * ystep + sy * q2);
      out[3] = (uint8_t)k1b_ybyte(y + 3 * ystep + sy * q3);
      y += 4 * ystep + sy * q4;
      err = n4 - q4 * adx;
    }
    if (left > 0) {   // the last 1..3 bins of the piece
      const int n1 = err + rem, n2 = n1 + rem;
      out[0] = (uint8_t)k1b_ybyte(y);
      if (left > 1) out[1] = (uint8_t)k1b_ybyte(y + ystep + sy * (int)K1B_MULHI((uint32_t)n1, magic));
      if (left > 2) out[2] = (uint8_t)k1b_ybyte(y + 2 * ystep + sy * (int)K1B_MULHI((uint32_t)n2, magic));
    }
  }
}

// What the gather path needs of a packet's descriptor, plus the first four words of its symbol record.  A warp
// fetches these for K1B_GRAB packets at once (one lane each) and hands them round by shuffle: the chain
// work-stealing counter -> order -> descriptor -> record header is four dependent memory round trips, which
// cost 0.9 of K1b's 4.3 ms when every packet walked it on its own.
struct K1bPkt {
  uint32_t idx, spec_off, setup_slot, rec_off, ent_off, byte_len;
  uint4 rh;
};
#ifndef K1B_GRAB
#define K1B_GRAB 8
#endif
// 1: phase B scans the entry counts of two stages in one 32-bit prefix sum
#ifndef K1B_PIN_REGS
#define K1B_PIN_REGS 1
#endif
#ifndef K1B_SCAN2
#define K1B_SCAN2 1
#endif

VPZ_DEV void k1b_prefetch_packet(const K1Params& P, uint32_t rec_off, uint32_t ent_off, uint32_t byte_len, int lane) {
#ifndef VPZ_EMU
  // the record (header, floor segments, classes) and the entry indices are read in dependent steps; ask
  // L2 for all of their lines now (one 128-byte line per lane: the record first, then ~4 bytes of entry
  // indices per packet byte, the typical rate) so those steps find them there
  const char* r0 = reinterpret_cast<const char*>(P.rec + rec_off);
  const char* e0 = reinterpret_cast<const char*>(P.ent + ent_off);
  const uint32_t rec_lines = 6, ent_bytes = 4u * byte_len;
  const char* line = lane < (int)rec_lines ? r0 + 128 * lane : e0 + 128 * (lane - (int)rec_lines);
  if (lane < (int)rec_lines || 128u * (uint32_t)(lane - (int)rec_lines) < ent_bytes)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
#endif
}

template <bool DEBUG>
VPZ_DEV void k1b_build_packet_gather(const K1Params& P, const K1bPkt& pk, uint32_t* smem, const float* dbtab, int tid) {
  const int lane = tid;   // ONE WARP per packet: many packets in flight per SM hide the per-packet load latency
  const uint32_t pkt_idx = pk.idx;
  const uint32_t* blob = P.setups[pk.setup_slot];
  VPZ_ASSUME_GLOBAL(blob);
  const VpzSetupHdr* H = reinterpret_cast<const VpzSetupHdr*>(blob);
  const int C = H->channels;
  const int half_max = 1 << (H->log2_size1 - 1);
  const uint32_t* rec = P.rec + pk.rec_off;
  const uint4 rh = pk.rh;   // records start on 16-byte boundaries (engine.cpp)
  const uint32_t hdr = rh.x;
  const uint32_t own_mask = hdr & 0xffu, noexec = (hdr >> 8) & 0xffu;
  const int long_block = (int)(hdr >> 24) & 1;
  const int half = long_block ? half_max : (1 << (H->log2_size0 - 1));

  if (own_mask == 0 && !(DEBUG && P.dbg.residue)) {  // every channel silent: K3 outputs zeros
    if (tid == 0) {
      VpzPktRes r;
      r.exec_mask = 0;
      r.status = (uint8_t)((hdr >> 16) & 0xffu);
      r.end16[0] = r.end16[1] = 0;
      P.res[pkt_idx] = r;
    }
    return;
  }

#if defined(K1B_ABLATE) && (K1B_ABLATE & 16)
  if (tid == 0) P.res[pkt_idx] = VpzPktRes{(uint8_t)own_mask, 0, {0, 0}};
  return;
#endif
  // mapping and residue index come with the record (K1a found them through the mode): no second walk
  const VpzMapping* mp = reinterpret_cast<const VpzMapping*>(blob + H->mappings_off) + (rh.w & 0xffu);
  const VpzResidue* rs = reinterpret_cast<const VpzResidue*>(blob + H->residues_off) + ((rh.w >> 8) & 0xffu);
  const K1ResGeom g = k1_res_geom(rs, C, half, noexec);
  const int nunits = g.part_count * g.nvec;
  const int U = (nunits + 31) & ~31;
  const int max_stages = rs->max_stages;

  uint32_t* urec = smem;
  uint8_t* ybuf = reinterpret_cast<uint8_t*>(urec + (size_t)max_stages * U * 2);   // [C][half_max] bytes
  uint32_t* sgbase = reinterpret_cast<uint32_t*>(ybuf) + (C * half_max) / 4;        // [C][4*66]

  K1Gather G;
  G.urec = urec;
  G.blob = blob;
  G.ent = P.ent + pk.ent_off;
  G.n_ent = rh.y;
  G.max_stages = max_stages;
  G.U = U;
  G.nvec = g.nvec;
  G.begin = g.begin;
  G.psize = g.psize;
  G.pshift = (g.psize & (g.psize - 1)) == 0 ? 31 - __clz(g.psize) : -1;
  G.span = g.part_count * g.psize;
  const bool have_res = g.part_count > 0 && g.any && G.n_ent > 0;
  const bool pair = g.rtype == 2 && C == 2;

  // Bins at and above the end of the coded residue range hold an exact +0 residue; phase A cuts the floor
  // into pieces below that bound, phase B then finds the packet's last ACTIVE partition, which is where the
  // render and the gather really stop.
  int res_cfg = 0;
  if (have_res) {
    const int span_end = g.begin + G.span;                          // in vector positions
    const int bins = pair ? (span_end + 1) >> 1 : span_end;
    res_cfg = (bins + 15) & ~15;
    if (res_cfg > half) res_cfg = half;
  }
  if (DEBUG && P.dbg.residue) res_cfg = half;   // the debug dump wants every bin

  // ---- phase A: floor segments of every channel with energy, and their pieces of <= 16 bins -------
  int npieces0 = 0, npieces1 = 0;   // per channel (the gather path has at most two)
  const uint32_t* rcp = blob + H->rcp_off;
  for (int ch = 0; ch < C; ch++) {
    if (!((own_mask >> ch) & 1u)) continue;
    const uint32_t* seg = rec + K1_REC_HDR + ch * K1_SEG_WORDS;
    uint32_t* sg = sgbase + ch * P.seg_stride;
    // the points of the first 32 segments are requested together with the count (the record has room for 67 points
    // per channel whatever the count says): one memory round trip instead of two dependent ones
    const uint32_t sp0 = seg[1 + tid], sp1 = seg[2 + tid];
    const int nseg = (int)seg[0];
    int carry = 0;
    for (int s0 = 0; s0 < nseg; s0 += 32) {   // uniform trip count: every lane takes part in the scan
      const int s = s0 + tid;
      int np = 0;
      uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
      if (s < nseg) {
        const uint32_t p0 = s0 == 0 ? sp0 : seg[1 + s], p1 = s0 == 0 ? sp1 : seg[2 + s];
        const int x0 = (int)(p0 & 0xffffu), y0 = (int)(short)(p0 >> 16);
        const int x1 = (int)(p1 & 0xffffu), y1 = (int)(short)(p1 >> 16);
        const int dy = y1 - y0, adx = x1 - x0;
        const int ady = dy < 0 ? -dy : dy;
        // ceil(2^32 / adx) from the setup's table (post distances on this path are <= half <= 4096); adx = 1 has no
        // remainder steps.  base = floor(ady / adx) by the same multiply-high: ady < 2^16, so ady * adx < 2^32
        w3 = adx > 1 ? (adx <= VPZ_RCP_MAX ? VPZ_LDG(rcp + adx) : 0xffffffffu / (uint32_t)adx + 1u) : 0u;
        const int base = adx > 1 ? (adx <= VPZ_RCP_MAX ? (int)K1B_MULHI((uint32_t)ady, w3) : ady / adx) : (adx == 1 ? ady : 0);
        const int xe = x1 < res_cfg ? x1 : res_cfg;
        np = xe > x0 ? (xe - x0 + 15) >> 4 : 0;
        w0 = (uint32_t)x0 | ((uint32_t)x1 << 16);
        w1 = (uint32_t)(y0 & 0xffff) | ((uint32_t)base << 16);          // |base| of the DDA
        w2 = (uint32_t)(ady - base * adx) | (dy < 0 ? 0x80000000u : 0u); // remainder step (< adx <= 4096), sign
      }
      int incl = np;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += n;
      }
      if (s < nseg) {
        sg[4 * s] = w0;
        sg[4 * s + 1] = w1;
        sg[4 * s + 2] = w2 | ((uint32_t)(carry + incl - np) << 16);     // pieces before this segment
        sg[4 * s + 3] = w3;
      }
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    // pieces | segments << 16: phase C needs both
    if (ch == 0) npieces0 = carry | (nseg << 16); else npieces1 = carry | (nseg << 16);
  }
  __syncwarp();

  // ---- phase B: per unit and stage, how many entries it holds -> first entry (all stages at once) ----
  int act_units = 0;   // last unit that carries codewords in any stage, + 1
  if (have_res) {
    const uint8_t* rec_cls = reinterpret_cast<const uint8_t*>(rec + K1_REC_HDR + C * K1_SEG_WORDS);
    int carry[8];
#pragma unroll
    for (int s = 0; s < 8; s++) carry[s] = 0;
    for (int u0 = 0; u0 < nunits; u0 += 32) {
      const int u = u0 + tid;
      int cnt[8];
      uint32_t info[8], vqo[8];
      int c = -1;
      if (u < nunits) {
        const int v = g.nvec == 2 ? (u & 1) : 0;   // the gather path has one or two vectors
        if (!((g.skip >> v) & 1u)) c = rec_cls[u];
      }
      const uint2* ut = reinterpret_cast<const uint2*>(blob + rs->unit_tabb_off) + (c < 0 ? 0 : c) * 8;
#if K1B_SCAN2
#pragma unroll
      for (int s = 0; s < 8; s++) {
        cnt[s] = 0;
        info[s] = vqo[s] = 0;
        if (s < max_stages && c >= 0) {   // {dims | log2 dims << 8 | entries << 16, vq word offset}; 0: no codewords for (class, stage)
          const uint2 t = VPZ_LDG(ut + s);
          info[s] = t.x & 0xffffu;
          cnt[s] = (int)(t.x >> 16);
          vqo[s] = t.y;
        }
      }
      // two stages per prefix sum, 16 bits each: a unit holds <= part_size <= 2047 entries (engine.cpp), so the
      // 32 lanes of a round stay below 2^16 and no field carries into its neighbour
#pragma unroll
      for (int s = 0; s < 8; s += 2) {
        if (s < max_stages) {
          const uint32_t both = (uint32_t)cnt[s] | ((uint32_t)cnt[s + 1] << 16);
          uint32_t incl = both;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
          }
          const uint32_t round_total = __shfl_sync(0xffffffffu, incl, 31);
          const uint32_t excl = incl - both;
          cnt[s] = carry[s] + (int)(excl & 0xffffu);   // first entry of the unit inside the stage
          cnt[s + 1] = carry[s + 1] + (int)(excl >> 16);
          carry[s] += (int)(round_total & 0xffffu);
          carry[s + 1] += (int)(round_total >> 16);
        }
      }
#else
#pragma unroll
      for (int s = 0; s < 8; s++) {
        cnt[s] = 0;
        info[s] = vqo[s] = 0;
        if (s < max_stages) {
          if (c >= 0) {   // {dims | log2 dims << 8 | entries << 16, vq word offset}; 0: no codewords for (class, stage)
            const uint2 t = VPZ_LDG(ut + s);
            info[s] = t.x & 0xffffu;
            cnt[s] = (int)(t.x >> 16);
            vqo[s] = t.y;
          }
          int incl = cnt[s];
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
          }
          const int round_total = __shfl_sync(0xffffffffu, incl, 31);
          cnt[s] = carry[s] + incl - cnt[s];   // first entry of the unit inside the stage
          carry[s] += round_total;
        }
      }
#endif
      // a unit is active when any stage has codewords for its class (info holds the dimension: never 0 then)
      const uint32_t act = __ballot_sync(0xffffffffu, (info[0] | info[1] | info[2] | info[3] | info[4] | info[5] | info[6] | info[7]) != 0u);
      if (act) act_units = u0 + 32 - __clz((int)act);
      // absolute first entry = entries of all earlier stages + offset inside the stage; the stage
      // totals are only complete after the last round, so the stage bases are added below
      if (u < nunits) {
#pragma unroll
        for (int s = 0; s < 8; s++)
          if (s < max_stages) {
            // idle units keep 0; the stage base added below never reaches bit 28 (engine.cpp bounds the packet size)
            uint32_t* R = urec + (size_t)(s * U + u) * 2;
            R[0] = info[s] ? ((((info[s] >> 8) & 0xfu) + 1u) << 28) | (uint32_t)cnt[s] : 0u;
            R[1] = vqo[s];
          }
      }
      __syncwarp();
    }
    // add the stage bases
    uint32_t sbase[8];
    uint32_t acc = 0;
#pragma unroll
    for (int s = 0; s < 8; s++) {
      sbase[s] = acc;
      if (s < max_stages) acc += (uint32_t)carry[s];
    }
    for (int u = tid; u < nunits; u += 32) {
#pragma unroll
      for (int s = 1; s < 8; s++)
        if (s < max_stages && urec[(size_t)(s * U + u) * 2]) urec[(size_t)(s * U + u) * 2] += sbase[s];
    }
  }

  // Bins above the packet's last ACTIVE residue partition hold an exact +0 residue (zeroed buffer, and the
  // inverse coupling of (+0, +0) is (+0, +0)), so their spectrum is +0 whatever the floor says: the floor
  // is rendered and the gather runs only below res_end (a multiple of 16 bins).  The TestFiles code nothing
  // in 20 % (160 kb/s) to 87 % (48 kb/s) of the bins of a block.  For the 256 / 2048 IMDCT kernel the zero
  // tail is not even written: end16 tells it where to stop reading.
  int res_end = 0;
  if (act_units > 0) {
    const int part = g.nvec == 2 ? (act_units - 1) >> 1 : (act_units - 1);   // the gather path has one or two vectors
    const int span_end = g.begin + (part + 1) * g.psize;              // in vector positions
    const int bins = pair ? (span_end + 1) >> 1 : span_end;
    res_end = (bins + 15) & ~15;
    if (res_end > half) res_end = half;
  }
  if (DEBUG && P.dbg.residue) res_end = half;   // the debug dump wants every bin
#if defined(VPZ_K3_END) && !VPZ_K3_END
  const bool k3_reads_end = false;
#else
  const bool k3_reads_end = !DEBUG && H->log2_size0 == 8 && H->log2_size1 == 11;   // engine.cpp k3_fast (C <= 2 here)
#endif
  if (tid == 0) {
    VpzPktRes r;
    r.exec_mask = (uint8_t)own_mask;
    r.status = (uint8_t)((hdr >> 16) & 0xffu);
    // K3 tests the end per 128-bin unit (uniform over its threads): the zero fill below goes up to the next one
    r.end16[0] = r.end16[1] = (uint8_t)(k3_reads_end ? ((res_end + 127) & ~127) >> 4 : 255);
    P.res[pkt_idx] = r;
  }
  __syncwarp();

  // ---- phase C: floor curve as one byte per bin: exact integer DDA, a lane renders pieces of <= 16 bins
  for (int ch = 0; ch < C; ch++) {
    if (!((own_mask >> ch) & 1u)) continue;
    const uint32_t* sg = sgbase + ch * P.seg_stride;
    uint8_t* yb = ybuf + ch * half_max;
    const int np_ns = ch == 0 ? npieces0 : npieces1;
#if !(defined(K1B_ABLATE) && (K1B_ABLATE & 1))
    k1b_render_floor(sg, np_ns >> 16, np_ns & 0xffff, yb, res_end, tid);
#endif
  }
  __syncwarp();

  // ---- phase D: 8 bins per thread: gather the residue, inverse coupling, floor, store ------------
  float* out = P.spec + pk.spec_off;
  const bool coupled = C == 2 && mp->coupling_steps > 0;   // stereo: the only possible pair is (0,1) / (1,0)
#if defined(K1B_ABLATE) && (K1B_ABLATE & 4)
  for (int x0 = tid * 8; x0 < 0; x0 += 32 * 8) {
#else
  for (int x0 = tid * 8; x0 < res_end; x0 += 32 * 8) {
#endif
    float r[16];   // r[i] = channel 0, r[8+i] = channel 1 of bin x0+i
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = 0.f;
#if defined(K1B_ABLATE) && (K1B_ABLATE & 2)
    if (false) {
#else
    if (have_res) {
#endif
      if (pair) {
        k1g_fetch_chunk<16, true>(G, 0, 2 * x0, r);   // Residue2 de-interleave: positions (2x, 2x+1) = (ch0, ch1)
      } else {   // type 2 mono: the one vector is channel 0
        if (g.rtype == 2 || !(g.skip & 1u)) k1g_fetch_chunk<8, false>(G, 0, x0, r);
        if (g.rtype != 2 && C == 2 && !(g.skip & 2u)) k1g_fetch_chunk<8, false>(G, 1, x0, r + 8);
      }
    }
    float c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      c0[i] = r[i];
      c1[i] = r[8 + i];
    }
    if (DEBUG && P.dbg.residue) {
      for (int i = 0; i < 8; i++) {
        P.dbg.residue[x0 + i] = c0[i];
        if (C == 2) P.dbg.residue[half + x0 + i] = c1[i];
      }
    }
    // inverse coupling, last step first (Mapping.cs:166-172)
    if (coupled) {
      for (int sidx = mp->coupling_steps - 1; sidx >= 0; sidx--) {
        if (mp->mag[sidx] == 0) {
#pragma unroll
          for (int i = 0; i < 8; i++) k1_uncouple(c0[i], c1[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; i++) k1_uncouple(c1[i], c0[i]);
        }
      }
    }
    // floor multiply (one rounded product per bin, Floor1.cs:383,395) and store
    if (own_mask & 1u) {
      const uint2 yy = *reinterpret_cast<const uint2*>(ybuf + x0);
      float4 o0, o1;
      o0.x = __fmul_rn(c0[0], dbtab[yy.x & 0xffu]);
      o0.y = __fmul_rn(c0[1], dbtab[(yy.x >> 8) & 0xffu]);
      o0.z = __fmul_rn(c0[2], dbtab[(yy.x >> 16) & 0xffu]);
      o0.w = __fmul_rn(c0[3], dbtab[yy.x >> 24]);
      o1.x = __fmul_rn(c0[4], dbtab[yy.y & 0xffu]);
      o1.y = __fmul_rn(c0[5], dbtab[(yy.y >> 8) & 0xffu]);
      o1.z = __fmul_rn(c0[6], dbtab[(yy.y >> 16) & 0xffu]);
      o1.w = __fmul_rn(c0[7], dbtab[yy.y >> 24]);
      reinterpret_cast<float4*>(out + x0)[0] = o0;
      reinterpret_cast<float4*>(out + x0)[1] = o1;
    }
    if (C == 2 && (own_mask & 2u)) {
      const uint2 yy = *reinterpret_cast<const uint2*>(ybuf + half_max + x0);
      float4 o0, o1;
      o0.x = __fmul_rn(c1[0], dbtab[yy.x & 0xffu]);
      o0.y = __fmul_rn(c1[1], dbtab[(yy.x >> 8) & 0xffu]);
      o0.z = __fmul_rn(c1[2], dbtab[(yy.x >> 16) & 0xffu]);
      o0.w = __fmul_rn(c1[3], dbtab[yy.x >> 24]);
      o1.x = __fmul_rn(c1[4], dbtab[yy.y & 0xffu]);
      o1.y = __fmul_rn(c1[5], dbtab[(yy.y >> 8) & 0xffu]);
      o1.z = __fmul_rn(c1[6], dbtab[(yy.y >> 16) & 0xffu]);
      o1.w = __fmul_rn(c1[7], dbtab[yy.y >> 24]);
      reinterpret_cast<float4*>(out + half + x0)[0] = o0;
      reinterpret_cast<float4*>(out + half + x0)[1] = o1;
    }
  }
  const int fill_end = !k3_reads_end ? half : (((res_end + 127) & ~127) < half ? ((res_end + 127) & ~127) : half);
  for (int x0 = res_end + tid * 8; x0 < fill_end; x0 += 32 * 8) {   // no coded residue up here: +0
    const float4 z = float4{0.f, 0.f, 0.f, 0.f};
    if (own_mask & 1u) {
      reinterpret_cast<float4*>(out + x0)[0] = z;
      reinterpret_cast<float4*>(out + x0)[1] = z;
    }
    if (C == 2 && (own_mask & 2u)) {
      reinterpret_cast<float4*>(out + half + x0)[0] = z;
      reinterpret_cast<float4*>(out + half + x0)[1] = z;
    }
  }
  __syncwarp();
}

// The kernel bodies: gather path = every warp takes its own packets; general path = the CTA takes them.
template <bool DEBUG>
VPZ_DEV void k1b_gather_loop(const K1Params& P, uint32_t* smem) {
  const int tid = (int)threadIdx.x;
  int lane = tid & 31;
  const int warp = tid >> 5;
  // inverse_dB_table (Floor1.cs:407-473): identical in every setup image, staged once per CTA
  float* dbtab = reinterpret_cast<float*>(smem);
  {
    const uint32_t* blob0 = P.setups[0];
    const float* db = reinterpret_cast<const float*>(blob0 + reinterpret_cast<const VpzSetupHdr*>(blob0)->db_off);
    for (int i = tid; i < 256; i += K1B_THREADS) dbtab[i] = VPZ_LDG(db + i);
  }
  __syncthreads();
#if K1B_PIN_REGS && !defined(VPZ_EMU)
  // At 56 registers the compiler re-derived the lane and the warp's shared-memory base from threadIdx wherever
  // they were used (56 S2R per packet, a tenth of the kernel's instructions on these three source lines): opaque
  // values cannot be rematerialised
  uint32_t woff = 256u + (uint32_t)warp * P.smem_words_per_warp;
  asm volatile("" : "+r"(woff));
  asm volatile("" : "+r"(lane));
  uint32_t* my = smem + woff;
#else
  uint32_t* my = smem + 256 + (size_t)warp * P.smem_words_per_warp;
#endif
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(P.counter, (uint32_t)K1B_GRAB);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= P.n_pkts) break;
    const int n = (int)(P.n_pkts - base < (uint32_t)K1B_GRAB ? P.n_pkts - base : (uint32_t)K1B_GRAB);
    // lane j < n fetches packet base + j: same order as K1a (grouped by (setup, block size)), so the warps
    // resident on an SM run the same code paths on the same VQ tables
    K1bPkt mine;
    mine.idx = mine.spec_off = mine.setup_slot = mine.rec_off = mine.ent_off = mine.byte_len = 0;
    mine.rh = uint4{0, 0, 0, 0};
    if (lane < n) {
      mine.idx = P.order ? P.order[base + lane] : base + lane;
      const uint4* d = reinterpret_cast<const uint4*>(P.pkts + mine.idx);   // VpzPktIn: 32 bytes, 16-byte aligned
      const uint4 d0 = d[0], d1 = d[1];
      mine.byte_len = d0.y;
      mine.spec_off = d0.z;
      mine.setup_slot = d0.w;
      mine.rec_off = d1.x;
      mine.ent_off = d1.y;
      mine.rh = *reinterpret_cast<const uint4*>(P.rec + mine.rec_off);
    }
    k1b_prefetch_packet(P, __shfl_sync(0xffffffffu, mine.rec_off, 0), __shfl_sync(0xffffffffu, mine.ent_off, 0),
                        __shfl_sync(0xffffffffu, mine.byte_len, 0), lane);
    for (int j = 0; j < n; j++) {
      // the lines of the NEXT packet are requested while this one is built
      if (j + 1 < n)
        k1b_prefetch_packet(P, __shfl_sync(0xffffffffu, mine.rec_off, j + 1), __shfl_sync(0xffffffffu, mine.ent_off, j + 1),
                            __shfl_sync(0xffffffffu, mine.byte_len, j + 1), lane);
      K1bPkt pk;
      pk.idx = __shfl_sync(0xffffffffu, mine.idx, j);
      pk.spec_off = __shfl_sync(0xffffffffu, mine.spec_off, j);
      pk.setup_slot = __shfl_sync(0xffffffffu, mine.setup_slot, j);
      pk.rec_off = __shfl_sync(0xffffffffu, mine.rec_off, j);
      pk.ent_off = __shfl_sync(0xffffffffu, mine.ent_off, j);
      pk.byte_len = __shfl_sync(0xffffffffu, mine.byte_len, j);
      pk.rh.x = __shfl_sync(0xffffffffu, mine.rh.x, j);
      pk.rh.y = __shfl_sync(0xffffffffu, mine.rh.y, j);
      pk.rh.z = __shfl_sync(0xffffffffu, mine.rh.z, j);
      pk.rh.w = __shfl_sync(0xffffffffu, mine.rh.w, j);
      k1b_build_packet_gather<DEBUG>(P, pk, my, dbtab, lane);
    }
  }
}
template <bool DEBUG>
VPZ_DEV void k1b_general_loop(const K1Params& P, uint32_t* smem, uint32_t* s_idx) {
  const int tid = (int)threadIdx.x;
  for (;;) {
    __syncthreads();
    if (tid == 0) *s_idx = atomicAdd(P.counter, 1u);
    __syncthreads();
    const uint32_t idx = *s_idx;
    if (idx >= P.n_pkts) break;
    k1b_build_packet_general<DEBUG>(P, P.order ? P.order[idx] : idx, smem, tid);
  }
}
