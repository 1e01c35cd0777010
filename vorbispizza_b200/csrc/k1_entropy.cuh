// k1_entropy.cuh -- K1: packet bytes -> spectrum (IMDCT input), one warp per packet.
//
// Replaces, per audio packet (reference file:line):
//   Floor1.Unpack                      Floor1.cs:162-219      (serial bit decode)
//   Mapping.DecodePacket flag logic    Mapping.cs:121-130
//   Residue0/1/2.Decode + WriteVectors Residue0.cs:117-231, Residue1.cs:12-34, Residue2.cs:12-52
//   Codebook.DecodeScalar              Codebook.cs:301-335    (two-level table, same symbol + bits)
//   Mapping.ApplyCoupling              Mapping.cs:198-269
//   Floor1.UnwrapPosts / Apply         Floor1.cs:222-397      (warp-parallel closed-form line render)
//
// Execution model: the bit cursor of a packet is strictly serial, so all 32 lanes of the warp run
// the SAME decode redundantly (uniform addresses -> broadcast loads, no divergence, no shuffles) and
// split only the data-parallel parts by lane: VQ vector adds, zeroing, coupling, floor render and
// the spectrum store.  The residue of the whole packet lives in shared memory; the only HBM
// traffic is the packet bytes in and the fp32 spectrum out.  Tables sit in global memory behind
// L1 (ld.global.nc); all lanes hit the same line.
#pragma once
#include "k1_params.h"

#ifndef VPZ_EMU
#define VPZ_DEV __device__ __forceinline__
#define VPZ_LDG(p) __ldg(p)
#else
#define VPZ_DEV inline
#define VPZ_LDG(p) (*(p))
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

struct K1Bits {
  const uint32_t* w;
  int pos, nbits;
  int is_short;
};

VPZ_DEV uint32_t k1_peek32(const K1Bits& b) {
  int i = b.pos >> 5;
  uint32_t lo = VPZ_LDG(b.w + i), hi = VPZ_LDG(b.w + i + 1);
  return __funnelshift_r(lo, hi, b.pos & 31);
}

// VorbisPacket.ReadBits (VorbisPacket.cs:157-164): zero-extended, truncated at the end, n <= 32
VPZ_DEV uint32_t k1_read(K1Bits& b, int n) {
  if (n <= 0) return 0;
  uint32_t v = k1_peek32(b);
  if (n < 32) v &= (1u << n) - 1u;
  int np = b.pos + n;
  b.pos = np < b.nbits ? np : b.nbits;
  return v;
}

struct K1Book {
  const uint32_t* l1;
  const uint32_t* ranges;
  const uint32_t* lcode;
  const uint32_t* linfo;
  const float* vq;
  uint32_t l1_mask;
  int dims;
};

VPZ_DEV K1Book k1_book(const uint32_t* blob, const VpzBook* books, int idx) {
  const VpzBook* bk = books + idx;
  K1Book r;
  r.l1 = blob + VPZ_LDG(&bk->l1_off);
  r.ranges = blob + VPZ_LDG(&bk->range_off);
  r.lcode = blob + VPZ_LDG(&bk->lcode_off);
  r.linfo = blob + VPZ_LDG(&bk->linfo_off);
  r.vq = reinterpret_cast<const float*>(blob + VPZ_LDG(&bk->vq_off));
  // dims (u16) and l1_bits (u8) share one 32-bit word at byte offset 24
  uint32_t packed = VPZ_LDG(reinterpret_cast<const uint32_t*>(bk) + 6);
  r.dims = (int)(packed & 0xffffu);
  r.l1_mask = (1u << ((packed >> 16) & 0xffu)) - 1u;
  return r;
}

// Codebook.DecodeScalar (Codebook.cs:301-335).  -1: no bits left or no code matches.
template <bool DEBUG>
VPZ_DEV int k1_decode(K1Bits& b, const K1Book& bk, const K1Params& P, int& nscal, int lane) {
  int sym = -1;
  if (b.pos < b.nbits) {
    uint32_t x = k1_peek32(b);
    uint32_t e = VPZ_LDG(bk.l1 + (x & bk.l1_mask));
    if (e & 0x80000000u) {
      uint32_t id = e & 0x7fffffffu;
      uint32_t lo = VPZ_LDG(bk.ranges + 2 * id), hi = VPZ_LDG(bk.ranges + 2 * id + 1);
      uint32_t m = __brev(x);
      while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (VPZ_LDG(bk.lcode + mid) <= m) lo = mid; else hi = mid;
      }
      uint32_t info = VPZ_LDG(bk.linfo + lo);
      uint32_t len = info & 0xffu;
      e = (((m ^ VPZ_LDG(bk.lcode + lo)) >> (32u - len)) == 0u) ? info : 0u;
    }
    if (e != 0u) {
      int len = (int)(e & 0xffu);
      sym = (int)(e >> 8);
      int np = b.pos + len;
      if (np > b.nbits) {  // SkipBits past the end: VorbisPacket.cs:248-292
        np = b.nbits;
        b.is_short = 1;
      }
      b.pos = np;
    }
  }
  if (DEBUG) {
    if (lane == 0 && P.dbg.scalars && nscal < P.dbg.scalars_cap) P.dbg.scalars[nscal] = sym;
    nscal++;
  }
  return sym;
}

// Floor1.RenderPoint (Floor1.cs:355-370)
VPZ_DEV int k1_render_point(int x0, int y0, int x1, int y1, int X) {
  int dy = y1 - y0, adx = x1 - x0;
  int ady = dy < 0 ? -dy : dy;
  int off = ady * (X - x0) / adx;
  return dy < 0 ? y0 - off : y0 + off;
}

// Per-warp shared memory layout (32-bit words):
//   res   [C * half_max]          residue / spectrum accumulators (float)
//   posts [C * 64]                raw floor posts (int)
//   fy    [64]                    unwrapped Y of the channel being rendered (int)
//   segx  [66], segy [66]         flagged posts in X order (int)
//   cls   [cls_words]             partition classes, one byte each
template <bool DEBUG>
VPZ_DEV void k1_decode_packet(const K1Params& P, uint32_t pkt_idx, uint32_t* smem, int lane) {
  const VpzPktIn pk = P.pkts[pkt_idx];
  const uint32_t* blob = P.setups[pk.setup_slot];
  const VpzSetupHdr* H = reinterpret_cast<const VpzSetupHdr*>(blob);
  const int C = H->channels;
  const VpzBook* books = reinterpret_cast<const VpzBook*>(blob + H->books_off);
  const int half_max = 1 << (H->log2_size1 - 1);

  float* res = reinterpret_cast<float*>(smem);
  int* posts = reinterpret_cast<int*>(smem + C * half_max);
  int* fy = posts + C * 64;
  int* segx = fy + 64;
  int* segy = segx + 66;
  uint8_t* cls = reinterpret_cast<uint8_t*>(segy + 66);

  K1Bits b;
  b.w = P.bytes + (pk.byte_off >> 2);
  b.pos = 0;
  b.nbits = (int)pk.byte_len * 8;
  b.is_short = 0;
  int nscal = 0, ncls = 0;

  // StreamDecoder.DecodeNextPacket (StreamDecoder.cs:728-741): the host only queues packets whose
  // type bit is 0 and whose mode exists, so these reads just advance the cursor.
  k1_read(b, 1);
  int mode_idx = (int)k1_read(b, H->mode_bits);
  const VpzMode* modes = reinterpret_cast<const VpzMode*>(blob + H->modes_off);
  int long_block = modes[mode_idx].block_flag;
  const VpzMapping* mp = reinterpret_cast<const VpzMapping*>(blob + H->mappings_off) + modes[mode_idx].mapping;
  if (long_block) k1_read(b, 2);  // prev/next window flags (Mode.cs:38), geometry is the host's job
  const int half = long_block ? half_max : (1 << (H->log2_size0 - 1));

  // ---- floor unpack, channel by channel (Mapping.cs:106-116, Floor1.cs:162-219) ----------
  uint32_t own_mask = 0;  // bit ch: floor has energy (FloorData.ExecuteChannel)
  int post_count[VPZ_MAX_CH];
#pragma unroll
  for (int ch = 0; ch < VPZ_MAX_CH; ch++) post_count[ch] = 0;
  for (int ch = 0; ch < C; ch++) {
    const VpzFloor1* fl = reinterpret_cast<const VpzFloor1*>(blob + H->floors_off) + mp->submap_floor[mp->mux[ch]];
    int* po = posts + ch * 64;
    for (int i = lane; i < 64; i += 32) po[i] = 0;  // FloorData.Reset
    __syncwarp();
    int count = 0;
    if (k1_read(b, 1) == 1) {
      int ybits = fl->ybits;
      int p0 = (int)k1_read(b, ybits), p1 = (int)k1_read(b, ybits);
      if (lane == 0) {
        po[0] = p0;
        po[1] = p1;
      }
      count = 2;
      int nparts = fl->partitions;
      for (int i = 0; i < nparts && count > 0; i++) {
        int c = fl->part_class[i];
        int cdim = fl->class_dim[c], cbits = fl->class_sub[c];
        uint32_t csub = (1u << cbits) - 1u, cval = 0;
        if (cbits > 0) {
          K1Book mb = k1_book(blob, books, fl->class_master[c]);
          int v = k1_decode<DEBUG>(b, mb, P, nscal, lane);
          if (v < 0) {
            count = 0;
            break;
          }
          cval = (uint32_t)v;
        }
        for (int j = 0; j < cdim; j++) {
          int book_idx = fl->sub_books[c][cval & csub];
          cval >>= cbits;
          int post = 0;
          if (book_idx >= 0) {
            K1Book sb = k1_book(blob, books, book_idx);
            post = k1_decode<DEBUG>(b, sb, P, nscal, lane);
            if (post < 0) {
              count = 0;
              break;
            }
          }
          if (lane == 0) po[count] = post;
          count++;
        }
      }
    }
    post_count[ch] = count;
    if (count > 0) own_mask |= 1u << ch;
  }
  // no-energy propagation through the coupling steps (Mapping.cs:121-130)
  uint32_t noexec = ~own_mask & ((1u << C) - 1u);
  for (int i = 0; i < mp->coupling_steps; i++) {
    uint32_t mb = 1u << mp->mag[i], ab = 1u << mp->ang[i];
    if (!((noexec & mb) && (noexec & ab))) noexec &= ~(mb | ab);
  }
  __syncwarp();
  if (DEBUG && P.dbg.hdr) {
    for (int ch = 0; ch < C; ch++)
      for (int i = lane; i < 64; i += 32) P.dbg.hdr[DUMP_RAWPOSTS + ch * 64 + i] = posts[ch * 64 + i];
  }

  // ---- residue (single submap: Setup::parse refuses more) -----------------------------------
  for (int i = lane; i < C * half; i += 32) res[i] = 0.f;
  __syncwarp();
  const VpzResidue* rs = reinterpret_cast<const VpzResidue*>(blob + H->residues_off) + mp->submap_residue[0];
  const int rtype = rs->type;
  int status = 0;
  {
    // Residue2 (Residue2.cs:12-52): one interleaved vector of length half*C, flags ignored unless
    // every channel is silent; Residue0/1: per channel vectors, silent channels skipped.
    int nvec = rtype == 2 ? 1 : C;
    int vlen = rtype == 2 ? half * C : half;
    uint32_t skip = rtype == 2 ? ((noexec == ((1u << C) - 1u)) ? 1u : 0u) : noexec;
    int begin = (int)rs->begin < vlen ? (int)rs->begin : vlen;
    int end = (int)rs->end < vlen ? (int)rs->end : vlen;
    int n = end - begin;
    int psize = (int)rs->part_size;
    int part_count = n > 0 ? n / psize : 0;
    bool any = false;
    for (int v = 0; v < nvec; v++) any |= !((skip >> v) & 1u);
    if (part_count > 0 && any) {
      K1Book cb = k1_book(blob, books, rs->class_book);
      const int cdim = cb.dims;
      const uint8_t* dmap = reinterpret_cast<const uint8_t*>(blob + rs->decode_map_off);
      const int partvals = (int)rs->decode_map_len / cdim;
      const int max_stages = rs->max_stages;
      bool abort = false;
      for (int stage = 0; stage < max_stages && !abort; stage++) {
        for (int part = 0; part < part_count && !abort;) {
          if (stage == 0) {
            for (int v = 0; v < nvec; v++) {
              if ((skip >> v) & 1u) continue;
              int idx = k1_decode<DEBUG>(b, cb, P, nscal, lane);
              // quirk Q8 accepts idx < partvals*dim; beyond partvals the reference indexes past
              // _decodeMap and throws, so both ends are treated as "stop decoding this packet"
              if (idx < 0 || idx >= partvals) {
                abort = true;
                break;
              }
              for (int k = lane; k < cdim; k += 32)
                if (part + k < part_count) cls[v * part_count + part + k] = dmap[idx * cdim + k];
            }
            __syncwarp();
            if (abort) break;
          }
          for (int k = 0; k < cdim && part < part_count && !abort; k++, part++) {
            int offset = begin + part * psize;
            for (int v = 0; v < nvec; v++) {
              if ((skip >> v) & 1u) continue;
              int c = cls[v * part_count + part];
              if (DEBUG && stage == 0) {
                if (lane == 0 && P.dbg.classes && ncls < P.dbg.classes_cap) P.dbg.classes[ncls] = c;
                ncls++;
              }
              if (!((rs->cascade[c] >> stage) & 1u) || !rs->has_books[c]) continue;
              K1Book vb = k1_book(blob, books, rs->books[c][stage]);
              float* dst = res + v * half;  // rtype 2: v == 0
              if (rtype == 0) {
                // Residue0.WriteVectors (Residue0.cs:208-231), quirk Q6: dims summed into one bin
                int steps = psize / vb.dims;
                for (int s = 0; s < steps; s++) {
                  int entry = k1_decode<DEBUG>(b, vb, P, nscal, lane);
                  if (entry < 0) {
                    abort = true;
                    break;
                  }
                  if (lane == 0) {
                    float r = 0.f;
                    const float* lk = vb.vq + (size_t)entry * vb.dims;
                    for (int d = 0; d < vb.dims; d++) r = __fadd_rn(r, VPZ_LDG(lk + d));
                    if (offset + s < vlen) dst[offset + s] = __fadd_rn(dst[offset + s], r);
                  }
                }
              } else {
                // Residue1.WriteVectors (Residue1.cs:12-34)
                for (int i = 0; i < psize;) {
                  int entry = k1_decode<DEBUG>(b, vb, P, nscal, lane);
                  if (entry < 0) {
                    abort = true;
                    break;
                  }
                  const float* lk = vb.vq + (size_t)entry * vb.dims;
                  for (int j = lane; j < vb.dims; j += 32) {
                    int at = offset + i + j;
                    if (at < vlen) dst[at] = __fadd_rn(dst[at], VPZ_LDG(lk + j));
                  }
                  i += vb.dims;
                }
              }
              if (abort) break;
            }
          }
          __syncwarp();
        }
        __syncwarp();
      }
      if (abort) status = 1;
    }
  }
  __syncwarp();

  // accessor of channel c, bin i after the Residue2 de-interleave (Residue2.cs:42-50)
#define RES_AT(c, i) res[rtype == 2 ? (i) * C + (c) : (c) * half + (i)]

  if (DEBUG && P.dbg.residue) {
    for (int c = 0; c < C; c++)
      for (int i = lane; i < half; i += 32) P.dbg.residue[c * half + i] = RES_AT(c, i);
  }

  // ---- inverse coupling, last step first (Mapping.cs:166-172, 235-267) ----------------------
  for (int s = mp->coupling_steps - 1; s >= 0; s--) {
    int cm = mp->mag[s], ca = mp->ang[s];
    for (int i = lane; i < half; i += 32) {
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
    __syncwarp();
  }

  // ---- floor synthesis + store (Floor1.cs:222-397) ------------------------------------------
  const float* db = reinterpret_cast<const float*>(blob + H->db_off);
  float* out = P.spec + pk.spec_off;
  for (int ch = 0; ch < C; ch++) {
    if (!((own_mask >> ch) & 1u)) continue;  // Mapping.cs:185-194: silent channel, K3 sees zeros
    const VpzFloor1* fl = reinterpret_cast<const VpzFloor1*>(blob + H->floors_off) + mp->submap_floor[mp->mux[ch]];
    const int* po = posts + ch * 64;
    const int count = post_count[ch];
    const int range = fl->range;
    // UnwrapPosts (Floor1.cs:270-353): serial dependency through earlier posts
    unsigned long long flags = 3ull;
    if (lane == 0) {
      fy[0] = po[0];
      fy[1] = po[1];
    }
    __syncwarp();
    for (int i = 2; i < count; i++) {
      int lo = fl->lneigh[i], hi = fl->hneigh[i];
      int predicted = k1_render_point(fl->xlist[lo], fy[lo], fl->xlist[hi], fy[hi], fl->xlist[i]);
      int val = po[i];
      int highroom = range - predicted, lowroom = predicted;
      int room = (highroom < lowroom ? highroom : lowroom) * 2;
      int result = predicted;
      if (val != 0) {
        flags |= (1ull << lo) | (1ull << hi) | (1ull << i);
        if (val >= room)
          result = highroom > lowroom ? val - lowroom + predicted : predicted - val + highroom - 1;
        else
          result = (val & 1) ? predicted - ((val + 1) >> 1) : predicted + (val >> 1);
      }
      __syncwarp();
      if (lane == 0) fy[i] = result;
      __syncwarp();
    }
    if (DEBUG && P.dbg.hdr) {
      for (int i = lane; i < 64; i += 32) {
        P.dbg.hdr[DUMP_FINALY + ch * 64 + i] = i < count ? fy[i] : 0;
        P.dbg.hdr[DUMP_STEPFLAGS + ch * 64 + i] = i < count ? (int)((flags >> i) & 1ull) : 0;
      }
    }
    // flagged posts in X order -> segment list (Floor1.Apply, Floor1.cs:222-268)
    const int mult = fl->multiplier;
    int nseg = 0;
    {
      int lx = 0, ly = fy[0] * mult;
      if (lane == 0) {
        segx[0] = 0;
        segy[0] = ly;
      }
      for (int i = 1; i < count; i++) {
        int idx = fl->sortidx[i];
        if ((flags >> idx) & 1ull) {
          int hx = fl->xlist[idx], hy = fy[idx] * mult;
          if (lx < half) {
            nseg++;
            if (lane == 0) {
              segx[nseg] = hx < half ? hx : half;  // quirk Q1: clamp before the slope
              segy[nseg] = hy;
            }
          }
          lx = hx;
          ly = hy;
        }
        if (lx >= half) break;
      }
      if (lx < half) {  // flat tail
        nseg++;
        if (lane == 0) {
          segx[nseg] = half;
          segy[nseg] = ly;
        }
      }
    }
    __syncwarp();
    // RenderLineMulti (Floor1.cs:372-397) in closed form: after k steps of the DDA
    // y = y0 + k*base + sy*floor(k*rem/adx), rem = |dy| - |base|*adx
    // a segment that was clamped at `half` keeps lx/ly of the unclamped post for the NEXT segment
    // in the reference, but no segment follows a clamped one (the loop breaks), so segx/segy chain.
    for (int s = 0; s < nseg; s++) {
      int x0 = segx[s], y0 = segy[s], x1 = segx[s + 1], y1 = segy[s + 1];
      int dy = y1 - y0, adx = x1 - x0;
      int ady = dy < 0 ? -dy : dy;
      int sy = dy < 0 ? -1 : 1;
      int base = dy / adx;
      int rem = ady - (base < 0 ? -base : base) * adx;
      for (int x = x0 + lane; x < x1; x += 32) {
        int k = x - x0;
        int y = y0 + k * base + sy * ((k * rem) / adx);
        y = y < 0 ? 0 : (y > 255 ? 255 : y);  // the reference reads the table unchecked (quirk Q2)
        out[ch * half + x] = __fmul_rn(RES_AT(ch, x), VPZ_LDG(db + y));
      }
    }
    __syncwarp();
  }
#undef RES_AT

  if (lane == 0) {
    VpzPktRes r;
    r.exec_mask = (uint8_t)own_mask;
    r.status = (uint8_t)status;
    r.bits_used_lo = (uint16_t)b.pos;
    P.res[pkt_idx] = r;
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
      for (int ch = 0; ch < C; ch++) h[DUMP_POSTCOUNT + ch] = post_count[ch];
    }
  }
}
