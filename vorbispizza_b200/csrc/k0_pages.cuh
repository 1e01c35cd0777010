// k0_pages.cuh -- K0: the physical Ogg layer on the GPU.  ONE WARP PER CONTAINER IMAGE walks the page chain of
// its file: capture-pattern search, header parse, lacing sums, packet counts, and the page CRC-32 computed by the
// 32 lanes together.  It emits one VpzPageRec per valid page plus the file's waste / CRC-failure counters; the
// host turns the records into its per-serial page lists (ogg.cpp, OggContainer::scan_from_records) without
// touching the page bytes again.
//
// Replaces (reference file:line):
//   PageReaderBase.ReadNextPage / VerifyHeader / VerifyPage   Ogg/PageReaderBase.cs:41-84,176-212,286-361
//     (sync search byte by byte, "OggS", segment table inside the data, CRC over the page with a zeroed
//      CRC field; a candidate that fails is skipped ONE byte and counted as waste)
//   Crc.Update / Crc.Test                                     Ogg/Crc.cs:20-63, Ogg/Crc.Table.cs:14 (poly 0x04c11db7)
//   PageHeader.GetPacketCount                                 Ogg/PageHeader.cs:35-59
#pragma once
#include "k1_params.h"

#ifndef VPZ_EMU
#define K0_DEV __device__ __forceinline__
#else
#define K0_DEV inline
#endif

#define K0_POLY 0x04c11db7u
#define K0_THREADS 128
// shared memory words: byte table [256], lane constants [32] (x^(128 (31 - l)) mod P), x^4096 mod P
#define K0_SMEM_WORDS (256 + 32 + 1)

// carry-less a * b mod P (degree < 32 operands)
K0_DEV uint32_t k0_mulmod(uint32_t a, uint32_t b) {
  uint32_t r = 0;
#pragma unroll 4
  for (int i = 31; i >= 0; i--) {
    r = (r << 1) ^ ((r & 0x80000000u) ? K0_POLY : 0u);
    if ((b >> i) & 1u) r ^= a;
  }
  return r;
}

// x^n mod P
K0_DEV uint32_t k0_xpow(uint32_t n) {
  uint32_t r = 1;
  for (uint32_t i = 0; i < n; i++) r = (r << 1) ^ ((r & 0x80000000u) ? K0_POLY : 0u);
  return r;
}

// Filled by the whole CTA once: tab[256] = the MSB-first byte table, then the combine constants.
K0_DEV void k0_init_tables(uint32_t* sm, int tid) {
  for (int i = tid; i < 256; i += K0_THREADS) {
    uint32_t r = (uint32_t)i << 24;
    for (int j = 0; j < 8; j++) r = (r << 1) ^ ((r & 0x80000000u) ? K0_POLY : 0u);
    sm[i] = r;
  }
  if (tid < 32) sm[256 + tid] = k0_xpow(128u * (uint32_t)(31 - tid));
  if (tid == 32) sm[256 + 32] = k0_xpow(4096u);
  __syncthreads();
}

// Ogg page CRC by one warp: the page is `n` bytes at img + pos; bytes 22..25 (the CRC field) count as zero.
// The CRC starts from 0 and has no final xor, so zero bytes in FRONT of the message do not change it: the
// message is padded at the front to a multiple of 512 bytes, every round the 32 lanes take 16 bytes each from
// state 0, their values are brought to the end of the round by the constants x^(128 (31 - lane)) and summed;
// the running value moves on by x^4096 per round.
K0_DEV uint32_t k0_page_crc(const uint8_t* img, uint32_t pos, uint32_t n, const uint32_t* sm, int lane) {
  const int pad = (int)((512u - (n & 511u)) & 511u);
  const int rounds = (int)((n + (uint32_t)pad) >> 9);
  const uint32_t klane = sm[256 + lane], kround = sm[256 + 32];
  uint32_t crc = 0;
  for (int r = 0; r < rounds; r++) {
    const int i0 = r * 512 + lane * 16 - pad;   // first message byte of this lane's 16
    uint32_t c = 0;
    if (i0 + 16 > 0) {
#pragma unroll 4
      for (int k = 0; k < 16; k++) {
        const int i = i0 + k;
        const uint32_t b = (i < 0 || (i >= 22 && i < 26)) ? 0u : (uint32_t)img[pos + (uint32_t)i];
        c = (c << 8) ^ sm[((c >> 24) ^ b) & 0xffu];
      }
    }
    c = k0_mulmod(c, klane);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, d);
    crc = k0_mulmod(crc, kround) ^ c;
  }
  return crc;
}

K0_DEV uint32_t k0_load_le32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// One file by one warp.  All control flow is warp-uniform (values come from ballots / shuffles).
K0_DEV void k0_scan_file(const K0Params& P, uint32_t fi, const uint32_t* sm, int lane) {
  const VpzScanFile f = P.files[fi];
  const uint8_t* img = P.images + f.data_off;
  const uint32_t len = f.len;
  VpzPageRec* out = P.pages + f.page_base;
  uint32_t n_pages = 0, crc_fail = 0, overflow = 0;
  unsigned long long waste = 0;   // bytes
  uint32_t pos = 0;
  bool resync = false;
  while (pos + 4 <= len) {
    // ---- capture pattern at pos?  else the next candidate among the following positions --------------
    {
      const uint32_t q = pos + (uint32_t)lane;
      const bool hit = q + 4 <= len && img[q] == 'O' && img[q + 1] == 'g' && img[q + 2] == 'g' && img[q + 3] == 'S';
      const uint32_t m = __ballot_sync(0xffffffffu, hit);
      if (!m) {
        // every position tried and refused is one wasted byte (PageReaderBase.cs:56-70); positions with fewer
        // than four bytes behind them are not tried (they are counted after the loop)
        const uint32_t valid = len - 3u - pos;   // >= 1 here
        const uint32_t tried = valid < 32u ? valid : 32u;
        waste += tried;
        pos += tried;
        resync = true;
        continue;
      }
      const uint32_t skip = (uint32_t)__ffs((int)m) - 1u;
      if (skip) {
        waste += skip;
        pos += skip;
        resync = true;
      }
    }
    // ---- try_page(pos): header inside the data, segment table inside, body inside, CRC ----------------
    bool ok = pos + 27 <= len;
    uint32_t nseg = 0, body = 0, npk = 0, last_seg = 0;
    if (ok) {
      nseg = img[pos + 26];
      ok = pos + 27 + nseg <= len;
    }
    if (ok) {
      for (uint32_t s0 = 0; s0 < nseg; s0 += 32) {
        const uint32_t s = s0 + (uint32_t)lane;
        const uint32_t v = s < nseg ? (uint32_t)img[pos + 27 + s] : 0u;
        uint32_t sum = v;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        body += sum;
        npk += (uint32_t)__popc(__ballot_sync(0xffffffffu, s < nseg && v < 255u));
        if (s0 + 32 >= nseg) last_seg = __shfl_sync(0xffffffffu, v, (int)((nseg - 1u) & 31u));
      }
      ok = pos + 27 + nseg + body <= len;
    }
    const uint32_t total = 27u + nseg + body;
    if (ok) {
      const uint32_t want = k0_load_le32(img + pos + 22);
      const uint32_t crc = k0_page_crc(img, pos, total, sm, lane);
      if (crc != want) {
        crc_fail++;
        ok = false;
      }
    }
    if (!ok) {   // not a page: one byte of waste, search on (PageReaderBase.cs:56-70)
      pos++;
      waste++;
      resync = true;
      continue;
    }
    if (n_pages >= f.page_cap) {   // more pages than the caller sized for: the host scans this file itself
      overflow = 1;
      break;
    }
    if (lane == 0) {
      // the record leaves as two 16-byte stores (the array may be pinned host memory written over the link)
      const bool cont = nseg > 0 && last_seg == 255u;
      uint4 a, b;
      a.x = pos;                                  // offset
      a.y = body;                                 // body_len
      a.z = k0_load_le32(img + pos + 6);          // granule_lo
      a.w = k0_load_le32(img + pos + 10);         // granule_hi
      b.x = k0_load_le32(img + pos + 14);         // serial
      b.y = k0_load_le32(img + pos + 18);         // seq
      b.z = (uint32_t)img[pos + 5] | (nseg << 8) | ((resync ? 1u : 0u) << 16) | ((cont ? 1u : 0u) << 24);
      b.w = (npk + (cont ? 1u : 0u)) & 0xffffu;   // packet_count, pad
      uint4* dst = reinterpret_cast<uint4*>(out + n_pages);
      dst[0] = a;
      dst[1] = b;
    }
    n_pages++;
    resync = false;
    pos += total;
  }
  if (!overflow && pos < len) waste += len - pos;
  if (lane == 0) {
    uint4 a, b;
    a.x = n_pages;
    a.y = crc_fail;
    a.z = (uint32_t)waste;
    a.w = (uint32_t)(waste >> 32);
    b.x = overflow;
    b.y = b.z = b.w = 0;
    uint4* dst = reinterpret_cast<uint4*>(P.out + fi);
    dst[0] = a;
    dst[1] = b;
  }
}

// kernel body of the SERIAL pass: warps take files from the counter (word 3); with only_irregular set, the files the
// fast path finished are skipped
K0_DEV void k0_cta(const K0Params& P, uint32_t* sm) {
  const int tid = (int)threadIdx.x, lane = tid & 31;
  k0_init_tables(sm, tid);
  for (;;) {
    uint32_t fi = 0;
    if (lane == 0) fi = atomicAdd(P.counter + 3, 1u);
    fi = __shfl_sync(0xffffffffu, fi, 0);
    if (fi >= P.n_files) break;
    if (P.only_irregular && P.irregular[fi] == 0u) continue;
    k0_scan_file(P, fi, sm, lane);
    __syncwarp();
  }
}

// =============================================================================================================
// Fast path.  One warp per file is enough for thousands of small files, but it is a serial walk: a single 47 MB
// file took 200 ms (234 MB/s).  Almost every file is ONE GAPLESS CHAIN of valid pages, and for those the work
// splits: k0_walk follows the chain from header to header (capture pattern exactly where the previous page ended,
// segment table and body inside the file; no CRC) and writes the page records; k0_crc then checks the CRC of every
// such page with ALL warps of the GPU, one page per warp.  A file that is not such a chain -- garbage, a truncated
// page, trailing bytes, or a CRC that fails -- is flagged `irregular` and scanned by the serial pass above, which
// alone knows the reference's resynchronisation and waste accounting (PageReaderBase.cs:56-70).
// =============================================================================================================
K0_DEV void k0_walk_file(const K0Params& P, uint32_t fi, int lane) {
  const VpzScanFile f = P.files[fi];
  const uint8_t* img = P.images + f.data_off;
  const uint32_t len = f.len;
  VpzPageRec* out = P.pages + f.page_base;
  uint32_t n_pages = 0, overflow = 0, pos = 0;
  bool regular = len >= 27u;
  while (regular && pos < len) {
    if (pos + 27u > len) {
      regular = false;
      break;
    }
    // the 27 header bytes, one per lane
    const uint32_t hb = lane < 27 ? (uint32_t)img[pos + (uint32_t)lane] : 0u;
    const uint32_t want = lane == 0 ? 'O' : (lane == 1 || lane == 2 ? 'g' : 'S');
    if ((__ballot_sync(0xffffffffu, lane < 4 && hb == want) & 0xfu) != 0xfu) {
      regular = false;
      break;
    }
    const uint32_t nseg = __shfl_sync(0xffffffffu, hb, 26);
    if (pos + 27u + nseg > len) {
      regular = false;
      break;
    }
    uint32_t body = 0, npk = 0, last_seg = 0;
    for (uint32_t s0 = 0; s0 < nseg; s0 += 32) {
      const uint32_t s = s0 + (uint32_t)lane;
      const uint32_t v = s < nseg ? (uint32_t)img[pos + 27 + s] : 0u;
      uint32_t sum = v;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
      body += sum;
      npk += (uint32_t)__popc(__ballot_sync(0xffffffffu, s < nseg && v < 255u));
      if (s0 + 32 >= nseg) last_seg = __shfl_sync(0xffffffffu, v, (int)((nseg - 1u) & 31u));
    }
    const uint32_t total = 27u + nseg + body;
    if (pos + total > len) {
      regular = false;
      break;
    }
    if (n_pages >= f.page_cap) {   // more pages than the caller sized for: the host scans this file itself
      overflow = 1;
      break;
    }
    if (lane == 0) {
      const bool cont = nseg > 0 && last_seg == 255u;
      uint4 a, b;
      a.x = pos;
      a.y = body;
      a.z = k0_load_le32(img + pos + 6);
      a.w = k0_load_le32(img + pos + 10);
      b.x = k0_load_le32(img + pos + 14);
      b.y = k0_load_le32(img + pos + 18);
      b.z = (uint32_t)img[pos + 5] | (nseg << 8) | ((cont ? 1u : 0u) << 24);   // no resync in a gapless chain
      b.w = (npk + (cont ? 1u : 0u)) & 0xffffu;
      uint4* dst = reinterpret_cast<uint4*>(out + n_pages);
      dst[0] = a;
      dst[1] = b;
      P.jobs[atomicAdd(P.counter + 1, 1u)] = VpzCrcJob{fi, pos, total, 0u};   // its CRC is checked by k0_crc
    }
    n_pages++;
    pos += total;
  }
  if (!regular) {
    if (lane == 0) P.irregular[fi] = 1u;   // the serial pass redoes this file and writes its records and counters
    return;
  }
  if (lane == 0) {
    uint4 a, b;
    a.x = n_pages;
    a.y = a.z = a.w = 0u;     // no CRC failures, no waste: otherwise the file is not regular
    b.x = overflow;
    b.y = b.z = b.w = 0u;
    uint4* dst = reinterpret_cast<uint4*>(P.out + fi);
    dst[0] = a;
    dst[1] = b;
  }
}

K0_DEV void k0_walk_cta(const K0Params& P) {
  const int lane = (int)threadIdx.x & 31;
  for (;;) {
    uint32_t fi = 0;
    if (lane == 0) fi = atomicAdd(P.counter, 1u);
    fi = __shfl_sync(0xffffffffu, fi, 0);
    if (fi >= P.n_files) break;
    k0_walk_file(P, fi, lane);
    __syncwarp();
  }
}

// one page per warp: a CRC that does not match flags the file
K0_DEV void k0_crc_cta(const K0Params& P, uint32_t* sm) {
  const int tid = (int)threadIdx.x, lane = tid & 31;
  k0_init_tables(sm, tid);
  const uint32_t njobs = P.counter[1];
  for (;;) {
    uint32_t j = 0;
    if (lane == 0) j = atomicAdd(P.counter + 2, 1u);
    j = __shfl_sync(0xffffffffu, j, 0);
    if (j >= njobs) break;
    const VpzCrcJob job = P.jobs[j];
    const uint8_t* img = P.images + P.files[job.file].data_off;
    const uint32_t want = k0_load_le32(img + job.offset + 22);
    const uint32_t crc = k0_page_crc(img, job.offset, job.length, sm, lane);
    if (crc != want && lane == 0) P.irregular[job.file] = 1u;
    __syncwarp();
  }
}

// =============================================================================================================
// K0g: the seek index.  PacketProvider.SeekTo (Ogg/PacketProvider.cs:90-169) searches by the GRANULES EACH PAGE
// COMPLETES, which the reference derives lazily by walking every packet of every page up to the target and asking
// the decoder for its sample count (FillPageEndGranuleCache, Ogg/PacketProvider.cs:203-307, with
// StreamDecoder.GetPacketGranuleCount, StreamDecoder.cs:882-913: mode number and window flags in the packet's
// first bits).  Here the whole index of a file is built at once: ONE WARP PER FILE, a lane per page (lacing walk,
// two bytes per packet), a warp prefix sum over the pages.  Only clean single-stream files come here (no resync,
// consecutive sequence numbers, continuation flags that match; scan.cpp checks), so every packet is valid.
// =============================================================================================================
K0_DEV int k0g_packet_granules(const uint8_t* p, uint32_t len, const VpzGranFile& f) {
  // the first 1 + mode_bits + 2 bits (<= 9); bits past the end of the packet read as 0 (VorbisPacket.cs:157-164)
  const uint32_t w = (len > 0 ? (uint32_t)p[0] : 0u) | (len > 1 ? (uint32_t)p[1] << 8 : 0u);
  if (w & 1u) return 0;                                   // not an audio packet (StreamDecoder.cs:889-893)
  const uint32_t mode = (w >> 1) & ((1u << f.mode_bits) - 1u);
  if (mode >= f.nmodes) return 0;
  const bool lb = mode < 32 ? (f.mode_flags_lo >> mode) & 1u : (f.mode_flags_hi >> (mode - 32)) & 1u;
  const int s0 = 1 << f.log2_size0, s1 = 1 << f.log2_size1;
  if (!lb) return s0 / 2;
  const bool prev = (w >> (1 + f.mode_bits)) & 1u, next = (w >> (2 + f.mode_bits)) & 1u;
  const int left_start = prev ? 0 : (s1 - s0) / 4;        // Mode.GetPacketInfo, Mode.cs:41-66
  const int right_start = next ? s1 / 2 : (s1 * 3 - s0) / 4;
  return right_start - left_start;
}

K0_DEV void k0g_file(const K0gParams& P, uint32_t fi, int lane) {
  const VpzGranFile f = P.files[fi];
  const uint8_t* img = P.images + f.data_off;
  const VpzPageRec* recs = P.pages + f.page_base;
  long long* out = P.page_end + f.page_base;
  long long carry = 0;
  bool found = false;          // the first data page (first header granule > 0) has been seen
  uint32_t first_data = 0;
  for (uint32_t p0 = 0; p0 < f.n_pages; p0 += 32) {
    const uint32_t p = p0 + (uint32_t)lane;
    const bool in = p < f.n_pages;
    uint4 a = uint4{0, 0, 0, 0}, b = uint4{0, 0, 0, 0};
    if (in) {
      a = reinterpret_cast<const uint4*>(recs + p)[0];     // offset, body_len, granule lo / hi
      b = reinterpret_cast<const uint4*>(recs + p)[1];     // serial, seq, flags | nseg << 8 | resync << 16 | continued << 24, packet_count
    }
    const bool data_page = in && (int)a.w >= 0 && (a.z | a.w) != 0u;   // granule > 0 as a signed 64-bit number
    const uint32_t m = __ballot_sync(0xffffffffu, data_page);
    if (!found && m) {
      found = true;
      first_data = p0 + (uint32_t)__ffs((int)m) - 1u;
    }
    int length = 0;
    if (in && found && p >= first_data) {
      int first_real = 0;
      if (p > 0) {
        const uint4 pa = reinterpret_cast<const uint4*>(recs + p - 1)[0];
        const uint4 pb = reinterpret_cast<const uint4*>(recs + p - 1)[1];
        if ((pb.z >> 24) & 1u) {   // the previous page ends inside a packet, which this page (or a later one) completes
          const uint32_t nsp = (pb.z >> 8) & 0xffu;
          const uint8_t* seg = img + pa.x + 27;
          uint32_t off = 0, start = 0;
          for (uint32_t s = 0; s < nsp; s++) {
            const uint32_t v = seg[s];
            off += v;
            if (v < 255u) start = off;
          }
          length += k0g_packet_granules(seg + nsp + start, pa.y - start, f);   // >= 255 bytes of it lie in that page
          first_real = 1;
        }
      }
      uint32_t k = (uint32_t)first_real;
      if (p == first_data) k = 1;                             // PacketProvider.cs:266-270
      const uint32_t cont = (b.z >> 24) & 1u;
      const uint32_t p_count = (b.w & 0xffffu) - cont;
      const uint32_t ns = (b.z >> 8) & 0xffu;
      const uint8_t* seg = img + a.x + 27;
      const uint8_t* body = seg + ns;
      uint32_t acc = 0, pkt_start = 0, idx = 0;
      for (uint32_t s = 0; s < ns; s++) {
        const uint32_t v = seg[s];
        acc += v;
        if (v < 255u) {
          if (idx >= k && idx < p_count) length += k0g_packet_granules(body + pkt_start, acc - pkt_start, f);
          idx++;
          pkt_start = acc;
        }
      }
    }
    int incl = length;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += n;
    }
    if (in) out[p] = carry + (long long)incl;
    carry += (long long)__shfl_sync(0xffffffffu, incl, 31);
  }
}

K0_DEV void k0g_cta(const K0gParams& P) {
  const int lane = (int)threadIdx.x & 31;
  const uint32_t warps_per_cta = blockDim.x >> 5;
  for (uint32_t fi = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); fi < P.n_files; fi += gridDim.x * warps_per_cta)
    k0g_file(P, fi, lane);
}
