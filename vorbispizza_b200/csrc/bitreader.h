// bitreader.h -- host-side LSB-first bit cursor used for header parsing and packet geometry.
// Same externally visible behaviour as VorbisPacket.ReadBits/TryPeekBits/SkipBits
// (VorbisPacket.cs:157-292): reads past the end return the zero-extended remainder and do not
// flag the packet; only skipping past the end raises `is_short`.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace vpz {

struct BitReader {
  const uint8_t* p;
  int64_t nbits;
  int64_t pos = 0;
  bool is_short = false;

  BitReader(const uint8_t* data, size_t len) : p(data), nbits((int64_t)len * 8) {}

  int64_t remaining() const { return nbits - pos; }

  // up to 32 bits
  uint32_t peek(int n, int* got) const {
    int64_t rem = nbits - pos;
    int m = rem < n ? (int)rem : n;
    if (m < 0) m = 0;
    *got = m;
    uint64_t v = 0;
    int64_t byte = pos >> 3;
    int sh = (int)(pos & 7);
    int need = (sh + m + 7) >> 3;
    for (int i = 0; i < need; i++) v |= (uint64_t)p[byte + i] << (8 * i);
    v >>= sh;
    if (m < 64) v &= ((uint64_t)1 << m) - 1;
    return (uint32_t)v;
  }
  uint32_t read(int n) {
    uint32_t v = 0;
    int done = 0;
    while (n > 0) {  // header fields are <= 32 bits; loop keeps the helper general
      int take = n > 24 ? 24 : n, got;
      uint32_t part = peek(take, &got);
      v |= part << done;
      pos += got;
      if (got < take) break;
      done += take;
      n -= take;
    }
    return v;
  }
  bool read_bit() { return read(1) == 1; }
  void skip(int64_t n) {
    if (n <= remaining()) {
      pos += n;
    } else {
      pos = nbits;
      is_short = true;
    }
  }
};

inline int ilog(int x) {  // Utils.ilog (Utils.cs:19-28)
  int c = 0;
  while (x > 0) {
    ++c;
    x >>= 1;
  }
  return c;
}

inline uint32_t bitrev32(uint32_t n) {
  n = ((n & 0xAAAAAAAAu) >> 1) | ((n & 0x55555555u) << 1);
  n = ((n & 0xCCCCCCCCu) >> 2) | ((n & 0x33333333u) << 2);
  n = ((n & 0xF0F0F0F0u) >> 4) | ((n & 0x0F0F0F0Fu) << 4);
  n = ((n & 0xFF00FF00u) >> 8) | ((n & 0x00FF00FFu) << 8);
  return (n >> 16) | (n << 16);
}

}  // namespace vpz
