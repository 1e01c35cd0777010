/* vo_internal.h -- TEST INFRASTRUCTURE ONLY (see vorbis_oracle.h). */
#ifndef VO_INTERNAL_H
#define VO_INTERNAL_H

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "vorbis_oracle.h"

/* ------------------------------------------------------------------ bits --
 * Externally visible behaviour of VorbisPacket.TryPeekBits / SkipBits /
 * ReadBits (VorbisPacket.cs:157-292): LSB-first; peek(n) yields
 * min(n, remaining) bits zero-extended; skipping past the end consumes what is
 * left and raises IsShort.  The 64-bit bucket + 8 overflow bits of the
 * reference have no other observable effect, so the restatement keeps only a
 * bit cursor.  `data` must be followed by >= 16 readable zero bytes. */
typedef struct {
  const uint8_t* data;
  int64_t total_bits;
  int64_t pos;
  int is_short;
} vo_bits;

static inline void vo_bits_init(vo_bits* b, const uint8_t* data, int len_bytes) {
  b->data = data;
  b->total_bits = (int64_t)len_bytes * 8;
  b->pos = 0;
  b->is_short = 0;
}

static inline uint64_t vo_load64(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v; /* little-endian host */
}

/* TryPeekBits (VorbisPacket.cs:195-206) */
static inline uint64_t vo_peek(const vo_bits* b, int n, int* got) {
  int64_t rem = b->total_bits - b->pos;
  int m = (int64_t)n < rem ? n : (int)rem;
  *got = m;
  if (m <= 0) return 0;
  const uint8_t* p = b->data + (b->pos >> 3);
  int sh = (int)(b->pos & 7);
  uint64_t v = vo_load64(p) >> sh;
  if (sh && m > 64 - sh) v |= (uint64_t)p[8] << (64 - sh);
  if (m < 64) v &= (~(uint64_t)0) >> (64 - m);
  return v;
}

/* SkipBits / SkipExtraBits (VorbisPacket.cs:213-292) */
static inline int vo_skip(vo_bits* b, int n) {
  if (n <= 0) return 0;
  int64_t rem = b->total_bits - b->pos;
  if ((int64_t)n <= rem) {
    b->pos += n;
    return n;
  }
  b->pos = b->total_bits;
  b->is_short = 1;
  return (int)rem;
}

/* ReadBits (VorbisPacket.cs:157-164): truncated value, never sets IsShort */
static inline uint64_t vo_read_bits(vo_bits* b, int n) {
  int got;
  uint64_t v = vo_peek(b, n, &got);
  b->pos += got;
  return v;
}

static inline int vo_read_bit(vo_bits* b) { return vo_read_bits(b, 1) == 1; }

/* Utils.ilog (Utils.cs:19-28) */
static inline int vo_ilog(int x) {
  int c = 0;
  while (x > 0) {
    ++c;
    x >>= 1;
  }
  return c;
}

/* Utils.BitReverse (Utils.cs:30-42) */
static inline uint32_t vo_bitrev32(uint32_t n) {
  n = ((n & 0xAAAAAAAAu) >> 1) | ((n & 0x55555555u) << 1);
  n = ((n & 0xCCCCCCCCu) >> 2) | ((n & 0x33333333u) << 2);
  n = ((n & 0xF0F0F0F0u) >> 4) | ((n & 0x0F0F0F0Fu) << 4);
  n = ((n & 0xFF00FF00u) >> 8) | ((n & 0x00FF00FFu) << 8);
  return (n >> 16) | (n << 16);
}
static inline uint32_t vo_bitrev(uint32_t n, int bits) {
  return bits <= 0 ? 0 : vo_bitrev32(n) >> (32 - bits);
}

/* ----------------------------------------------------------------- setup -- */
typedef struct {
  int32_t value, length, bits, mask; /* Contracts/HuffmanListNode.cs:7-11 */
} vo_hnode;

typedef struct {
  int dims, entries, map_type;
  int max_bits;     /* Codebook._maxBits */
  int prefix_bits;  /* Huffman.TableBits */
  int* lengths;     /* -1 unused */
  vo_hnode* prefix; /* 1 << prefix_bits (NULL when empty) */
  vo_hnode* overflow;
  int overflow_n;
  float* lookup;    /* entries*dims, NULL for map type 0 */
} vo_book;

typedef struct {
  int type; /* 1 = floor 1 (Floor1.cs), 0 = floor 0 (Floor0.cs) */
  /* floor 0 (Floor0.cs:29-76) */
  int order, rate, bark_map_size, amp_bits, amp_ofs, nbooks0;
  uint8_t books0[16];
  int* bark_map[2];  /* n + 1 entries per block size (Floor0.cs:83-96) */
  float* wmap[2];    /* n entries per block size (Floor0.cs:103-113) */
  /* floor 1 */
  int partitions;
  uint8_t part_class[32];
  int class_count;
  uint8_t class_dim[16], class_sub[16], class_master[16];
  int16_t sub_books[16][8];
  int multiplier, range, ybits;
  int xcount;
  int xlist[256], lneigh[256], hneigh[256], sortidx[256];
} vo_floor1;

typedef struct {
  int type; /* 0, 1, 2 */
  int begin, end, part_size, classifications, class_book, max_stages;
  uint8_t cascade[64];
  int16_t books[64][8]; /* -1 = none */
  int has_books[64];
  int* decode_map;
  int decode_map_len;
  int* part_word_cache;
  int part_word_cache_len;
} vo_residue;

typedef struct {
  int submaps, coupling_steps;
  uint8_t mag[256], ang[256];
  uint8_t mux[VO_MAX_CH * 32];
  uint8_t submap_floor[16], submap_residue[16];
} vo_mapping;

typedef struct {
  int block_flag, mapping;
} vo_mode;

typedef struct {
  int channels, sample_rate, br_upper, br_nominal, br_lower;
  int size0, size1;
  int nbooks, nfloors, nresidues, nmappings, nmodes, mode_bits;
  vo_book* books;
  vo_floor1* floors;
  vo_residue* residues;
  vo_mapping* mappings;
  vo_mode* modes;
  float* slope[2];
} vo_setup;

int vo_setup_parse_id(vo_setup* st, const uint8_t* pkt, int len);
int vo_setup_parse_books(vo_setup* st, const uint8_t* pkt, int len);
void vo_setup_free(vo_setup* st);
int vo_book_decode_scalar(const vo_book* bk, vo_bits* br);

/* ---------------------------------------------------------------- decode -- */
typedef struct {
  int length, left_use_size1, left_start, left_end, right_start, right_end;
} vo_pinfo;

int vo_mode_packet_info(const vo_setup* st, const vo_mode* m, vo_bits* br, vo_pinfo* info);
/* Mapping.DecodePacket; buf = channels * size1 floats, stride size1 */
void vo_mapping_decode(vo_setup* st, const vo_mapping* mp, vo_bits* br, int block_size, float* buf,
                       vo_packet_dump* dump);

#endif
