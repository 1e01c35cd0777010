// setup.cpp -- Vorbis id/setup header parsing and GPU table construction (host side).
// Behaviour follows the reference (file:line cited per function); layout follows vpz_dev.h.
#include "setup.h"

#include <math.h>
#include <string.h>

#include <algorithm>

#include "../../include/vpz.h"
#include "bitreader.h"

namespace vpz {

static const uint32_t k_db_bits[256] = {
#include "db_table.inc"
};

uint64_t fnv1a64(const uint8_t* p, size_t n, uint64_t h) {
  for (size_t i = 0; i < n; i++) {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

// BlocksizeDerivedCache.CalcWindowSlope (BlocksizeDerivedCache.cs:24-35): every step in fp32.
void window_slope(float* slope, int n) {
  const float half_pi = 0.5f * 3.14159274f;
  for (int i = 0; i < n; i++) {
    volatile float a = half_pi * ((float)i + 0.5f);
    volatile float q = a / (float)n;
    volatile float v = sinf(q);
    volatile float b = half_pi * v;
    volatile float c = b * v;
    slope[i] = sinf(c);
  }
}

void compute_geometry(int size0, int size1, bool long_block, bool prev, bool next, PacketGeom* g) {
  int size = long_block ? size1 : size0;
  if (!long_block) prev = next = true;
  g->block_size = size;
  g->long_block = long_block;
  g->prev_flag = prev;
  g->next_flag = next;
  if (prev) {
    g->left_start = 0;
    g->left_end = size / 2;
    g->length = size / 2;
    g->left_use_size1 = long_block;
  } else {
    g->left_start = (size - size0) / 4;
    g->left_end = (size + size0) / 4;
    g->length = size0 / 2;
    g->left_use_size1 = false;
  }
  if (next) {
    g->right_start = size / 2;
    g->right_end = size;
  } else {
    g->right_start = (size * 3 - size0) / 4;
    g->right_end = (size * 3 + size0) / 4;
  }
}

// StreamDecoder.LoadStreamHeader (StreamDecoder.cs:213-240)
int parse_id_header(const uint8_t* pkt, size_t len, IdHeader* out) {
  static const uint8_t sig[11] = {0x01, 'v', 'o', 'r', 'b', 'i', 's', 0, 0, 0, 0};
  if (len < 11 || memcmp(pkt, sig, 11) != 0) return VPZ_E_INVALID_DATA;
  BitReader br(pkt, len);
  br.skip(88);
  out->channels = (int)br.read(8);
  out->sample_rate = (int)br.read(32);
  out->br_upper = (int)br.read(32);
  out->br_nominal = (int)br.read(32);
  out->br_lower = (int)br.read(32);
  out->size0 = 1 << br.read(4);
  out->size1 = 1 << br.read(4);
  if (out->br_nominal == 0 && out->br_upper > 0 && out->br_lower > 0)
    out->br_nominal = (out->br_upper + out->br_lower) / 2;
  return VPZ_OK;
}

namespace {

struct HostBook {
  int dims = 0, entries = 0, map_type = 0, max_bits = 0;
  std::vector<int> lengths;          // -1 unused
  std::vector<uint32_t> codes;       // LSB-first codeword per entry (valid where lengths > 0)
  std::vector<float> lookup;
};

// Codebook.ComputeCodewords (Codebook.cs:147-218): lowest free tree node at depth <= length,
// stored bit-reversed so it matches LSB-first stream bits.  false = over-subscribed.
bool assign_codewords(const std::vector<int>& len, std::vector<uint32_t>& codes) {
  uint32_t avail[33];
  memset(avail, 0, sizeof(avail));
  int n = (int)len.size();
  codes.assign(n, 0);
  int k = 0;
  while (k < n && len[k] <= 0) ++k;
  if (k == n) return true;
  codes[k] = 0;
  for (int i = 1; i <= len[k]; ++i) avail[i] = 1u << (32 - i);
  for (int i = k + 1; i < n; ++i) {
    int z = len[i];
    if (z <= 0) continue;
    while (z > 0 && avail[z] == 0) --z;
    if (z == 0) return false;
    uint32_t res = avail[z];
    avail[z] = 0;
    codes[i] = bitrev32(res);
    if (z != len[i])
      for (int y = len[i]; y > z; --y) avail[y] = res + (1u << (32 - y));
  }
  return true;
}

// Codebook.lookup1_values (Codebook.cs:290-298)
int lookup1_values(int entries, int dims) {
  int r = (int)floor(exp(log((double)entries) / dims));
  if (floor(pow((double)r + 1, dims)) <= entries) ++r;
  return r;
}

// Utils.ConvertFromVorbisFloat32 (Utils.cs:92-105)
float vorbis_float32(uint32_t bits) {
  int32_t sign = (int32_t)bits >> 31;
  int exponent = (int)((bits & 0x7fe00000u) >> 21) - 788;
  float mantissa = (float)((((int32_t)(bits & 0x1fffff)) ^ sign) + (sign & 1));
  return scalbnf(mantissa, exponent);
}

// Codebook ctor: InitTree (Codebook.cs:44-144) + InitLookupTable (:220-288)
int parse_book(BitReader& br, HostBook& bk, std::string& err) {
  if (br.read(24) != 0x564342u) {
    err = "codebook sync";
    return VPZ_E_INVALID_DATA;
  }
  bk.dims = (int)br.read(16);
  bk.entries = (int)br.read(24);
  bk.lengths.assign(bk.entries, -1);
  int max_len = -1;
  if (br.read_bit()) {  // ordered
    int len = (int)br.read(5) + 1;
    for (int i = 0; i < bk.entries;) {
      int cnt = (int)br.read(ilog(bk.entries - i));
      if (i + cnt > bk.entries || (cnt == 0 && br.remaining() <= 0)) {
        err = "ordered codebook overruns its entry count";
        return VPZ_E_INVALID_DATA;
      }
      while (--cnt >= 0) bk.lengths[i++] = len;
      ++len;
    }
    max_len = len;  // quirk Q7: lastLen + 1
  } else {
    bool sparse = br.read_bit();
    for (int i = 0; i < bk.entries; i++) {
      if (!sparse || br.read_bit()) {
        bk.lengths[i] = (int)br.read(5) + 1;
        max_len = std::max(max_len, bk.lengths[i]);
      }
    }
  }
  bk.max_bits = max_len < 0 ? 0 : max_len;
  if (max_len >= 0) {
    int used = 0, last = -1;
    for (int i = 0; i < bk.entries; i++)
      if (bk.lengths[i] > 0) {
        used++;
        last = i;
      }
    if (used == 1 && bk.lengths[last] != 1) {  // Huffman.cs:52-58 "Invalid single entry"
      err = "single-entry codebook with length != 1";
      return VPZ_E_INVALID_DATA;
    }
    if (!assign_codewords(bk.lengths, bk.codes)) {
      err = "over-subscribed codebook";
      return VPZ_E_INVALID_DATA;
    }
  }
  bk.map_type = (int)br.read(4);
  if (bk.map_type == 0) return VPZ_OK;
  if (bk.dims < 1) {
    // a lookup of zero dimensions: lookup1_values divides by it (Codebook.cs:290-298) and every residue /
    // floor that names the book divides the partition size by it
    err = "VQ codebook with zero dimensions";
    return VPZ_E_INVALID_DATA;
  }
  if (bk.map_type > 2) {
    // the reference builds no table for other values and faults on first use; we refuse early
    err = "codebook lookup type > 2";
    return VPZ_E_INVALID_DATA;
  }
  float min_value = vorbis_float32(br.read(32));
  float delta_value = vorbis_float32(br.read(32));
  int value_bits = (int)br.read(4) + 1;
  bool sequence_p = br.read_bit();
  int64_t total = (int64_t)bk.entries * bk.dims;
  if (total > (1 << 26)) {
    err = "codebook lookup too large";
    return VPZ_E_UNSUPPORTED;
  }
  int mult_n = bk.map_type == 1 ? lookup1_values(bk.entries, bk.dims) : (int)total;
  std::vector<uint16_t> mult((size_t)std::max(mult_n, 1));
  for (int i = 0; i < mult_n; i++) mult[i] = (uint16_t)br.read(value_bits);
  bk.lookup.assign((size_t)total, 0.f);
  for (int e = 0; e < bk.entries; e++) {
    volatile float last = 0.f;
    uint32_t div = 1;
    for (int i = 0; i < bk.dims; i++) {
      uint32_t moff = bk.map_type == 1 ? ((uint32_t)e / div) % (uint32_t)mult_n : (uint32_t)(e * bk.dims + i);
      volatile float v = (float)mult[moff] * delta_value;  // fp32 mul, add, add: Codebook.cs:262-285
      v = v + min_value;
      v = v + last;
      bk.lookup[(size_t)e * bk.dims + i] = v;
      if (sequence_p) last = v;
      if (bk.map_type == 1) div *= (uint32_t)mult_n;
    }
  }
  return VPZ_OK;
}

struct BlobWriter {
  std::vector<uint32_t>& w;
  explicit BlobWriter(std::vector<uint32_t>& v) : w(v) {}
  uint32_t here() const { return (uint32_t)w.size(); }
  uint32_t reserve(size_t words) {
    uint32_t off = here();
    w.resize(w.size() + words, 0u);
    return off;
  }
  template <typename T>
  uint32_t put_struct_array(const std::vector<T>& v) {
    size_t bytes = v.size() * sizeof(T);
    uint32_t off = reserve((bytes + 3) / 4);
    if (bytes) memcpy(&w[off], v.data(), bytes);
    return off;
  }
  uint32_t put_floats(const float* f, size_t n) {
    uint32_t off = reserve(n);
    if (n) memcpy(&w[off], f, n * 4);
    return off;
  }
  void align(size_t words) {
    while (w.size() % words) w.push_back(0);
  }
};

// Huffman.GenerateTable (Huffman.cs:24-105) re-expressed for the GPU decoder: same symbol and same
// bit consumption for every bit pattern, different table shape (see VpzBook in vpz_dev.h).
void build_decode_tables(const HostBook& hb, int l1_bits_cfg, BlobWriter& bw, VpzBook& out) {
  int max_len = 0;
  for (int i = 0; i < hb.entries; i++) max_len = std::max(max_len, hb.lengths[i]);
  int l1 = std::min(max_len, l1_bits_cfg);
  out.l1_bits = (uint8_t)l1;
  std::vector<uint32_t> l1tab((size_t)1 << l1, 0u);
  struct LongCode {
    uint32_t msb;  // left-aligned MSB-first code
    uint32_t info;
    uint32_t prefix;  // first l1 stream bits (LSB-first value)
  };
  std::vector<LongCode> longs;
  for (int e = 0; e < hb.entries; e++) {
    int len = hb.lengths[e];
    if (len <= 0) continue;
    uint32_t code = hb.codes[e];
    if (len <= l1) {
      int reps = 1 << (l1 - len);
      for (int j = 0; j < reps; j++) l1tab[((uint32_t)j << len) | code] = ((uint32_t)e << 8) | (uint32_t)len;
    } else {
      LongCode lc;
      lc.msb = bitrev32(code);  // code occupies the low `len` bits LSB-first -> top `len` bits MSB-first
      lc.info = ((uint32_t)e << 8) | (uint32_t)len;
      lc.prefix = l1 ? (code & ((1u << l1) - 1u)) : 0u;
      longs.push_back(lc);
    }
  }
  std::sort(longs.begin(), longs.end(), [](const LongCode& a, const LongCode& b) { return a.msb < b.msb; });
  // Second level: per long prefix a table over the next l2 bits (l2 = longest code under the prefix
  // minus l1, at most VPZ_L2_BITS_MAX), entries as in the first level.  Codes that are longer still
  // (probability <= 2^-(l1 + VPZ_L2_BITS_MAX) per symbol) keep the sorted-array bisection: their
  // second-level slot holds 0x80000000 | range id.
  std::vector<uint32_t> ranges;  // {lo, hi} pairs
  std::vector<uint32_t> l2tab;
  struct Fix {
    uint32_t prefix, rel, bits;
  };
  std::vector<Fix> fix;
  for (size_t i = 0; i < longs.size();) {
    size_t j = i;
    int longest = 0;
    while (j < longs.size() && longs[j].prefix == longs[i].prefix) {
      longest = std::max(longest, (int)(longs[j].info & 0xffu));
      ++j;
    }
    uint32_t id = (uint32_t)(ranges.size() / 2);
    ranges.push_back((uint32_t)i);
    ranges.push_back((uint32_t)j);
    const int l2 = std::min(longest - l1, VPZ_L2_BITS_MAX);
    const uint32_t rel = (uint32_t)l2tab.size();
    l2tab.resize(l2tab.size() + ((size_t)1 << l2), 0u);
    for (size_t k = i; k < j; k++) {
      const int len = (int)(longs[k].info & 0xffu);
      const uint32_t code = bitrev32(longs[k].msb);          // back to the LSB-first stream value
      const uint32_t sub = code >> l1;                       // the len - l1 bits after the prefix
      if (len - l1 <= l2) {
        const int reps = 1 << (l2 - (len - l1));
        for (int q = 0; q < reps; q++) l2tab[rel + (((uint32_t)q << (len - l1)) | sub)] = longs[k].info;
      } else {
        l2tab[rel + (sub & ((1u << l2) - 1u))] = 0x80000000u | id;
      }
    }
    fix.push_back(Fix{longs[i].prefix, rel, (uint32_t)l2});
    i = j;
  }
  bw.align(4);
  out.l1_off = bw.reserve(l1tab.size());
  const uint32_t l2_off = bw.reserve(l2tab.size());
  if (!l2tab.empty()) memcpy(&bw.w[l2_off], l2tab.data(), l2tab.size() * 4);
  // first-level entry of a long prefix: 0x80000000 | (second-level word offset from l1_off) << 5 | l2 bits
  for (const Fix& f : fix) l1tab[f.prefix] = 0x80000000u | ((l2_off - out.l1_off + f.rel) << 5) | f.bits;
  memcpy(&bw.w[out.l1_off], l1tab.data(), l1tab.size() * 4);
  out.range_off = bw.reserve(ranges.size());
  if (!ranges.empty()) memcpy(&bw.w[out.range_off], ranges.data(), ranges.size() * 4);
  out.long_n = (uint32_t)longs.size();
  out.lcode_off = bw.reserve(longs.size());
  out.linfo_off = bw.reserve(longs.size());
  for (size_t i = 0; i < longs.size(); i++) {
    bw.w[out.lcode_off + i] = longs[i].msb;
    bw.w[out.linfo_off + i] = longs[i].info;
  }
}

// Floor1 ctor (Floor1.cs:39-155)
int parse_floor1(BitReader& br, int nbooks, VpzFloor1& f, std::string& err) {
  static const uint16_t range_lookup[4] = {256, 128, 86, 64};
  static const uint8_t ybits_lookup[4] = {8, 7, 7, 6};
  memset(&f, 0, sizeof(f));
  f.floor_type = 1;
  int max_class = -1;
  f.partitions = (uint8_t)br.read(5);
  for (int i = 0; i < f.partitions; i++) {
    f.part_class[i] = (uint8_t)br.read(4);
    max_class = std::max(max_class, (int)f.part_class[i]);
  }
  for (int c = 0; c <= max_class; c++) {
    f.class_dim[c] = (uint8_t)(br.read(3) + 1);
    f.class_sub[c] = (uint8_t)br.read(2);
    if (f.class_sub[c] > 0) {
      f.class_master[c] = (uint8_t)br.read(8);
      if (f.class_master[c] >= nbooks) {
        err = "floor1 master book out of range";
        return VPZ_E_INVALID_DATA;
      }
    }
    for (int j = 0; j < (1 << f.class_sub[c]); j++) {
      int book = (int)br.read(8) - 1;
      if (book >= nbooks) {
        err = "floor1 subclass book out of range";
        return VPZ_E_INVALID_DATA;
      }
      f.sub_books[c][j] = (int16_t)book;
    }
  }
  int m = (int)br.read(2);
  f.multiplier = (uint8_t)(m + 1);
  f.range = range_lookup[m];
  f.ybits = ybits_lookup[m];
  int range_bits = (int)br.read(4);
  int n = 2;
  for (int i = 0; i < f.partitions; i++) n += f.class_dim[f.part_class[i]];
  if (n > VPZ_MAX_POSTS) {  // reference Posts[] holds 64 (Floor1.cs:17); a 65-post floor faults there
    err = "floor1 with more than 64 posts";
    return VPZ_E_UNSUPPORTED;
  }
  f.xcount = (uint8_t)n;
  int k = 0;
  f.xlist[k++] = 0;
  f.xlist[k++] = (uint16_t)(1 << range_bits);
  for (int i = 0; i < f.partitions; i++)
    for (int j = 0; j < f.class_dim[f.part_class[i]]; j++) f.xlist[k++] = (uint16_t)br.read(range_bits);
  // neighbours among earlier posts (Floor1.cs:109-133)
  for (int i = 2; i < n; i++) {
    int lo = 0, hi = 1;
    for (int j = 2; j < i; j++) {
      int t = f.xlist[j];
      if (t < f.xlist[i]) {
        if (t > f.xlist[lo]) lo = j;
      } else if (t < f.xlist[hi]) {
        hi = j;
      }
    }
    f.lneigh[i] = (uint8_t)lo;
    f.hneigh[i] = (uint8_t)hi;
  }
  // index sorted by X; duplicates are an error (Floor1.cs:136-149)
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++)
      if (f.xlist[i] == f.xlist[j]) {
        err = "floor1 duplicate X";
        return VPZ_E_INVALID_DATA;
      }
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return f.xlist[a] < f.xlist[b]; });
  for (int i = 0; i < n; i++) f.sortidx[i] = (uint8_t)idx[i];
  return VPZ_OK;
}

// Floor0 ctor (Floor0.cs:39-76) with SynthesizeBarkCurve (:83-96) and SynthesizeWDelMap (:103-113).
// The tables go into the blob later (bark[w], wmap[w]).
struct HostFloor0 {
  std::vector<uint16_t> bark[2];
  std::vector<float> wmap[2];
};
float floor0_to_bark(double lsp) {   // Floor0.ToBARK (Floor0.cs:98-101): double arithmetic, rounded to float
  return (float)(13.1 * atan(0.00074 * lsp) + 2.24 * atan(0.0000000185 * lsp * lsp) + .0001 * lsp);
}
int parse_floor0(BitReader& br, const std::vector<HostBook>& books, const IdHeader& id, VpzFloor1& f, HostFloor0& t,
                 std::string& err) {
  memset(&f, 0, sizeof(f));
  f.floor_type = 0;
  const int order = (int)br.read(8), rate = (int)br.read(16), bark_map_size = (int)br.read(16);
  const int amp_bits = (int)br.read(6), amp_ofs = (int)br.read(8), nb = (int)br.read(4) + 1;
  if (order < 1 || rate < 1 || bark_map_size < 1) {
    err = "floor0 with zero order / rate / bark map size";
    return VPZ_E_INVALID_DATA;
  }
  if (amp_bits < 1 || amp_bits > 32) {
    // amp_bits 0: the reference divides 0 by 0 and renders NaN; > 32 does not fit the GPU bit reader
    err = "floor0 amplitude width outside 1..32 bits is not on the GPU path";
    return VPZ_E_UNSUPPORTED;
  }
  f.f0.order = (uint8_t)order;
  f.f0.rate = (uint16_t)rate;
  f.f0.bark_map_size = (uint16_t)bark_map_size;
  f.f0.amp_bits = (uint8_t)amp_bits;
  f.f0.amp_ofs = (uint8_t)amp_ofs;
  f.f0.nbooks = (uint8_t)nb;
  f.f0.book_bits = (uint8_t)ilog(nb);
  for (int i = 0; i < nb; i++) {
    const int num = (int)br.read(8);
    if (num >= (int)books.size() || books[(size_t)num].map_type == 0 || books[(size_t)num].dims < 1) {
      err = "floor0 book without lookup";
      return VPZ_E_INVALID_DATA;
    }
    f.f0.books[i] = (uint8_t)num;
  }
  for (int w = 0; w < 2; w++) {
    const int n = (w ? id.size1 : id.size0) / 2;
    volatile float scale = (float)bark_map_size / floor0_to_bark(rate / 2.0);   // ushort / float in fp32
    t.bark[w].assign((size_t)n, 0);
    for (int i = 0; i < n - 1; i++) {   // i < map.Length - 2: bin n-1 keeps the default 0 (Floor0.cs:88-94)
      volatile float prod = floor0_to_bark((rate / 2.0) / n * i) * scale;
      const int v = std::min(bark_map_size - 1, (int)floor((double)prod));
      if (v >= n || v < 0) {
        // Apply reads wMap[barkMap[i]] with wMap of n entries (Floor0.cs:192): the reference faults here
        err = "floor0 bark map index beyond the block (the reference reads out of range)";
        return VPZ_E_UNSUPPORTED;
      }
      t.bark[w][(size_t)i] = (uint16_t)v;
    }
    volatile float wdel = (float)(M_PI / bark_map_size);
    t.wmap[w].assign((size_t)n, 0.f);
    for (int i = 0; i < n; i++) {
      volatile float a = wdel * (float)i;
      t.wmap[w][(size_t)i] = 2.0f * cosf(a);
    }
  }
  return VPZ_OK;
}

// Residue0 ctor (Residue0.cs:25-115)
int parse_residue(BitReader& br, int type, const std::vector<HostBook>& books, VpzResidue& r,
                  std::vector<uint8_t>& decode_map, std::string& err) {
  memset(&r, 0, sizeof(r));
  int nbooks = (int)books.size();
  r.type = (uint8_t)type;
  r.begin = br.read(24);
  r.end = br.read(24);
  r.part_size = br.read(24) + 1;
  int classifications = (int)br.read(6) + 1;
  r.classifications = (uint8_t)classifications;
  r.class_book = (uint8_t)br.read(8);
  int acc = 0;
  for (int i = 0; i < classifications; i++) {
    uint32_t low = br.read(4);
    uint32_t bits = low & 7u;
    if (low & 8u) bits |= br.read(5) << 3;
    r.cascade[i] = (uint8_t)bits;
    acc += __builtin_popcount(bits);
  }
  std::vector<uint8_t> book_nums((size_t)acc);
  for (int i = 0; i < acc; i++) {
    book_nums[i] = (uint8_t)br.read(8);
    if (book_nums[i] >= nbooks || books[book_nums[i]].map_type == 0) {
      err = "residue book without lookup";
      return VPZ_E_INVALID_DATA;
    }
  }
  if (r.class_book >= nbooks) {
    err = "residue classbook out of range";
    return VPZ_E_INVALID_DATA;
  }
  const HostBook& cb = books[r.class_book];
  int partvals = 1;
  for (int i = 0; i < cb.dims; i++) {
    partvals *= classifications;
    if (partvals > cb.entries) {
      err = "residue classbook too small";
      return VPZ_E_INVALID_DATA;
    }
  }
  acc = 0;
  int maxstage = 0;
  for (int j = 0; j < classifications; j++) {
    int stages = ilog(r.cascade[j]);
    if (stages <= 0) continue;
    r.has_books[j] = 1;
    maxstage = std::max(maxstage, stages);
    for (int k = 0; k < stages; k++) r.books[j][k] = (r.cascade[j] & (1 << k)) ? book_nums[acc++] : 0;
  }
  r.max_stages = (uint8_t)maxstage;
  r.decode_map_len = (uint32_t)(partvals * cb.dims);
  decode_map.assign((size_t)partvals * cb.dims, 0);
  for (int j = 0; j < partvals; j++) {
    int val = j, mult = partvals / classifications;
    for (int k = 0; k < cb.dims; k++) {
      int deco = val / mult;
      val -= deco * mult;
      mult /= classifications;
      decode_map[(size_t)j * cb.dims + k] = (uint8_t)deco;
    }
  }
  return VPZ_OK;
}

// Mapping ctor (Mapping.cs:19-95)
int parse_mapping(BitReader& br, int channels, int nfloors, int nresidues, VpzMapping& m, std::string& err) {
  memset(&m, 0, sizeof(m));
  m.submaps = 1;
  if (br.read_bit()) m.submaps = (uint8_t)(1 + br.read(4));
  int steps = 0;
  if (br.read_bit()) steps = (int)br.read(8) + 1;
  if (steps > 32) {
    err = "more than 32 coupling steps";
    return VPZ_E_UNSUPPORTED;
  }
  m.coupling_steps = (uint8_t)steps;
  int cbits = ilog(channels - 1);
  for (int j = 0; j < steps; j++) {
    int mag = (int)br.read(cbits), ang = (int)br.read(cbits);
    if (mag == ang || mag > channels - 1 || ang > channels - 1) {
      err = "bad coupling pair";
      return VPZ_E_INVALID_DATA;
    }
    m.mag[j] = (uint8_t)mag;
    m.ang[j] = (uint8_t)ang;
  }
  if (br.read(2) != 0) {
    err = "mapping reserved bits";
    return VPZ_E_INVALID_DATA;
  }
  if (m.submaps > 1)
    for (int c = 0; c < channels; c++) {
      m.mux[c] = (uint8_t)br.read(4);
      if (m.mux[c] >= m.submaps) {
        err = "mapping mux out of range";
        return VPZ_E_INVALID_DATA;
      }
    }
  for (int j = 0; j < m.submaps; j++) {
    br.read(8);
    int fl = (int)br.read(8), rs = (int)br.read(8);
    if (fl >= nfloors || rs >= nresidues) {
      err = "mapping floor/residue out of range";
      return VPZ_E_INVALID_DATA;
    }
    m.submap_floor[j] = (uint8_t)fl;
    m.submap_residue[j] = (uint8_t)rs;
  }
  return VPZ_OK;
}

}  // namespace

int Setup::parse(const uint8_t* id_pkt, size_t id_len, const uint8_t* setup_pkt, size_t setup_len, int l1_bits) {
  int rc = parse_id_header(id_pkt, id_len, &id);
  if (rc) {
    error = "bad identification header";
    return rc;
  }
  if (id.channels < 1 || id.channels > VPZ_MAX_CH) {
    error = "channel count outside the GPU path (1..8)";
    return id.channels < 1 ? VPZ_E_INVALID_DATA : VPZ_E_UNSUPPORTED;
  }
  if (id.size0 < 256 && id.size0 >= 64 && id.size1 >= id.size0 && id.size1 <= 8192) {
    // the reference's Mdct.CalcReverse (Mdct.cs:200-249) is not a valid IMDCT for
    // n = 64 / 128, so "identical to the reference" would mean reproducing noise; refuse instead.
    error = "block size 64/128: the reference MDCT is not a valid transform below 256 (quirk Q10)";
    return VPZ_E_UNSUPPORTED;
  }
  if (id.size0 < 64 || id.size1 < id.size0 || id.size1 > 8192) {
    error = "block sizes outside 64..8192";
    return VPZ_E_INVALID_DATA;
  }
  hash = fnv1a64(setup_pkt, setup_len, fnv1a64(id_pkt, id_len));
  static const uint8_t sig[7] = {0x05, 'v', 'o', 'r', 'b', 'i', 's'};
  if (setup_len < 7 || memcmp(setup_pkt, sig, 7) != 0) {
    error = "bad setup header signature";
    return VPZ_E_INVALID_DATA;
  }
  BitReader br(setup_pkt, setup_len);
  br.skip(56);

  int nbooks = (int)br.read(8) + 1;
  std::vector<HostBook> books((size_t)nbooks);
  for (int i = 0; i < nbooks; i++)
    if ((rc = parse_book(br, books[i], error)) != VPZ_OK) return rc;

  int times = (int)br.read(6) + 1;
  br.skip(16 * times);  // StreamDecoder.cs:278

  int nfloors = (int)br.read(6) + 1;
  std::vector<VpzFloor1> floors((size_t)nfloors);
  std::vector<HostFloor0> floors0((size_t)nfloors);
  for (int i = 0; i < nfloors; i++) {
    int type = (int)br.read(16);
    if (type == 0) {
      if ((rc = parse_floor0(br, books, id, floors[i], floors0[i], error)) != VPZ_OK) return rc;
      continue;
    }
    if (type != 1) {
      error = "invalid floor type";
      return VPZ_E_INVALID_DATA;
    }
    if ((rc = parse_floor1(br, nbooks, floors[i], error)) != VPZ_OK) return rc;
  }

  int nres = (int)br.read(6) + 1;
  std::vector<VpzResidue> residues((size_t)nres);
  std::vector<std::vector<uint8_t>> dmaps((size_t)nres);
  for (int i = 0; i < nres; i++) {
    int type = (int)br.read(16);
    if (type > 2) {
      error = "invalid residue type";
      return VPZ_E_INVALID_DATA;
    }
    if ((rc = parse_residue(br, type, books, residues[i], dmaps[i], error)) != VPZ_OK) return rc;
  }

  int nmaps = (int)br.read(6) + 1;
  std::vector<VpzMapping> mappings((size_t)nmaps);
  for (int i = 0; i < nmaps; i++) {
    if (br.read(16) != 0) {
      error = "invalid mapping type";
      return VPZ_E_INVALID_DATA;
    }
    if ((rc = parse_mapping(br, id.channels, nfloors, nres, mappings[i], error)) != VPZ_OK) return rc;
  }
  // Residue INSTANCES (vpz_dev.h): the first nres are the header's residues serving all channels (what a
  // single-submap mapping uses); a submap with fewer channels gets its own instance, because the walk
  // tables depend on the number of vectors.  submap_residue is rewritten to instance indices.
  n_residues_hdr = nres;
  {
    std::vector<int> inst_nch((size_t)nres, id.channels), inst_src((size_t)nres);
    for (int i = 0; i < nres; i++) inst_src[(size_t)i] = i;
    for (int i = 0; i < nmaps; i++) {
      VpzMapping& m = mappings[(size_t)i];
      for (int j = 0; j < m.submaps; j++) {
        int nch = 0;
        for (int c = 0; c < id.channels; c++) nch += (m.submaps > 1 ? m.mux[c] : 0) == j;
        const int src = m.submap_residue[j];
        int found = -1;
        for (size_t k = 0; k < inst_src.size(); k++)
          if (inst_src[k] == src && inst_nch[k] == std::max(nch, 1)) found = (int)k;
        if (found < 0) {
          found = (int)inst_src.size();
          inst_src.push_back(src);
          inst_nch.push_back(std::max(nch, 1));
          dmaps.push_back(dmaps[(size_t)src]);
          residues.push_back(residues[(size_t)src]);
        }
        if (found > 255) {
          error = "more than 255 (residue, submap width) combinations are not on the GPU path";
          return VPZ_E_UNSUPPORTED;
        }
        m.submap_residue[j] = (uint8_t)found;
      }
    }
    for (size_t k = 0; k < residues.size(); k++) residues[k].nch = (uint16_t)inst_nch[k];
    nres = (int)residues.size();
  }

  int nmodes = (int)br.read(6) + 1;
  modes.assign((size_t)nmodes, VpzMode{0, 0});
  for (int i = 0; i < nmodes; i++) {
    modes[i].block_flag = br.read_bit() ? 1 : 0;
    if (br.read(32) != 0) {  // Mode.cs:17-20
      error = "mode header had invalid window or transform type";
      return VPZ_E_INVALID_DATA;
    }
    int mp = (int)br.read(8);
    if (mp >= nmaps) {
      error = "mode header had invalid mapping index";
      return VPZ_E_INVALID_DATA;
    }
    modes[i].mapping = (uint8_t)mp;
  }
  if (!br.read_bit()) {  // StreamDecoder.cs:312
    error = "setup header framing bit";
    return VPZ_E_INVALID_DATA;
  }
  mode_bits = ilog(nmodes - 1);

  // ---- emit the device image -------------------------------------------------------------
  blob.clear();
  BlobWriter bw(blob);
  bw.reserve((sizeof(VpzSetupHdr) + 3) / 4);
  bw.align(4);
  std::vector<VpzBook> dbooks((size_t)nbooks);
  max_codeword_bits = 0;
  for (int i = 0; i < nbooks; i++) {
    VpzBook& d = dbooks[i];
    memset(&d, 0, sizeof(d));
    d.entries = (uint32_t)books[i].entries;
    if (books[i].map_type != 0 && (books[i].entries > 65535 || books[i].dims > 255)) {
      // K1a stores VQ entry indices as uint16 and K1b packs the dimension in 8 bits
      error = "VQ codebook with more than 65535 entries or 255 dimensions is not on the GPU path";
      return VPZ_E_UNSUPPORTED;
    }
    d.dims = (uint16_t)books[i].dims;
    d.max_bits = (uint8_t)books[i].max_bits;
    d.map_type = (uint8_t)books[i].map_type;
    for (int e = 0; e < books[i].entries; e++) max_codeword_bits = std::max(max_codeword_bits, books[i].lengths[e]);
    build_decode_tables(books[i], l1_bits, bw, d);
    if (books[i].map_type != 0) {
      bw.align(4);
      d.vq_off = bw.put_floats(books[i].lookup.data(), books[i].lookup.size());
    }
  }
  for (int i = 0; i < nres; i++) {
    bw.align(4);
    residues[i].decode_map_off = bw.reserve((dmaps[i].size() + 3) / 4);
    if (!dmaps[i].empty()) memcpy(&blob[residues[i].decode_map_off], dmaps[i].data(), dmaps[i].size());
    // ---- K1a walk tables (vpz_dev.h): which (class, stage) units carry codewords and with which book,
    // and per classword value the units of its partition group that are active in every stage.  They
    // restate the tests of Residue0.Decode (Residue0.cs:160-190: cascade bit + book present) as lookups.
    VpzResidue& r = residues[i];
    const int cdim = books[r.class_book].dims;
    const int nvec = r.type == 2 ? 1 : r.nch;
    const int partvals = (int)(r.decode_map_len / (uint32_t)std::max(cdim, 1));
    r.cdim = (uint16_t)cdim;
    r.nvec = (uint16_t)nvec;
    r.partvals = (uint32_t)partvals;
    if (cdim < 1 || cdim * nvec > 32) {
      error = "residue classbook dimension x vectors above 32 is not on the GPU path";
      return VPZ_E_UNSUPPORTED;
    }
    if ((uint64_t)partvals * (uint64_t)nvec * std::max<int>(r.max_stages, 1) > (1u << 20)) {
      error = "residue classword table too large for the GPU path";
      return VPZ_E_UNSUPPORTED;
    }
    uint32_t active[64] = {0};   // per class: bit s set when stage s decodes codewords for a partition of that class
    bw.align(2);
    r.unit_tab_off = bw.reserve((size_t)r.classifications * 8 * 2);
    r.unit_tabb_off = bw.reserve((size_t)r.classifications * 8 * 2);
    for (int c = 0; c < r.classifications; c++)
      for (int s = 0; s < 8; s++) {
        if (!(((r.cascade[c] >> s) & 1) && r.has_books[c])) continue;
        const int bk = r.books[c][s];
        const int dims = books[bk].dims;
        // Residue0.WriteVectors decodes psize / dims entries, Residue1.WriteVectors steps by dims until psize
        const uint32_t n = r.type == 0 ? r.part_size / (uint32_t)dims : (r.part_size + (uint32_t)dims - 1) / (uint32_t)dims;
        if (n == 0) continue;
        if (n > 65535u) {
          error = "residue partition with more than 65535 codewords is not on the GPU path";
          return VPZ_E_UNSUPPORTED;
        }
        blob[r.unit_tab_off + (size_t)(c * 8 + s) * 2] = dbooks[bk].l1_off;
        blob[r.unit_tab_off + (size_t)(c * 8 + s) * 2 + 1] = (uint32_t)dbooks[bk].l1_bits | ((uint32_t)bk << 8) | (n << 16);
        int lg = 0;
        while ((1 << (lg + 1)) <= dims) lg++;   // exact for the power-of-two dimensions the gather path accepts
        blob[r.unit_tabb_off + (size_t)(c * 8 + s) * 2] = (uint32_t)(dims & 0xff) | ((uint32_t)lg << 8) | (n << 16);
        blob[r.unit_tabb_off + (size_t)(c * 8 + s) * 2 + 1] = dbooks[bk].vq_off;
        active[c] |= 1u << s;
      }
    const int ns = std::max<int>(r.max_stages, 1);
    r.cw_tab_off = bw.reserve((size_t)nvec * partvals * ns);
    for (int v = 0; v < nvec; v++)
      for (int sym = 0; sym < partvals; sym++)
        for (int s = 0; s < ns; s++) {
          uint32_t m = 0;
          for (int k = 0; k < cdim; k++)
            if ((active[dmaps[i][(size_t)sym * cdim + k]] >> s) & 1u) m |= 1u << (k * nvec + v);
          blob[r.cw_tab_off + ((size_t)v * partvals + sym) * ns + s] = m;
        }
  }
  VpzSetupHdr h;
  memset(&h, 0, sizeof(h));
  h.magic = 0x315A5056u;
  h.channels = (uint8_t)id.channels;
  h.log2_size0 = (uint8_t)(ilog(id.size0) - 1);
  h.log2_size1 = (uint8_t)(ilog(id.size1) - 1);
  h.mode_bits = (uint8_t)mode_bits;
  h.nmodes = (uint8_t)nmodes;
  h.nmappings = (uint8_t)nmaps;
  h.nfloors = (uint8_t)nfloors;
  h.nresidues = (uint8_t)nres;
  h.nbooks = (uint32_t)nbooks;
  bw.align(4);
  h.books_off = bw.put_struct_array(dbooks);
  for (int i = 0; i < nfloors; i++) {
    if (floors[i].floor_type != 0) continue;
    for (int w = 0; w < 2; w++) {
      bw.align(4);
      floors[i].f0.bark_off[w] = bw.reserve((floors0[i].bark[w].size() + 1) / 2);
      memcpy(&blob[floors[i].f0.bark_off[w]], floors0[i].bark[w].data(), floors0[i].bark[w].size() * 2);
      bw.align(4);
      floors[i].f0.wmap_off[w] = bw.put_floats(floors0[i].wmap[w].data(), floors0[i].wmap[w].size());
    }
  }
  for (int i = 0; i < nfloors; i++) {
    if (floors[i].floor_type != 1) continue;
    VpzFloor1& f = floors[i];
    bw.align(2);
    f.fbook_tab_off = bw.reserve(16 * 9 * 2);
    auto put = [&](int c, int k, int bk) {
      if (bk < 0 || bk >= nbooks) return;
      blob[f.fbook_tab_off + (size_t)(c * 9 + k) * 2] = dbooks[(size_t)bk].l1_off;
      blob[f.fbook_tab_off + (size_t)(c * 9 + k) * 2 + 1] = (uint32_t)dbooks[(size_t)bk].l1_bits | ((uint32_t)bk << 8) | 0x10000u;   // bit 16: a book is there (l1_bits may be 0)
    };
    for (int c = 0; c < 16; c++) {
      if (f.class_sub[c] > 0) put(c, 0, f.class_master[c]);
      for (int k = 0; k < 8; k++) put(c, 1 + k, f.sub_books[c][k]);
    }
  }
  bw.align(4);
  h.floors_off = bw.put_struct_array(floors);
  bw.align(4);
  h.residues_off = bw.put_struct_array(residues);
  bw.align(4);
  h.mappings_off = bw.put_struct_array(mappings);
  bw.align(4);
  h.modes_off = bw.put_struct_array(modes);
  for (int w = 0; w < 2; w++) {
    int size = w ? id.size1 : id.size0;
    std::vector<float> slope((size_t)size / 2);
    window_slope(slope.data(), size / 2);
    bw.align(4);
    h.slope_off[w] = bw.put_floats(slope.data(), slope.size());
    // IMDCT as a DCT-IV through an N/4-point complex FFT: pre/post twiddle exp(-i*pi*(n+1/8)/M),
    // M = N/2; roots exp(-2*pi*i*k/H), H = N/4.  Computed in double, rounded once to fp32.
    int M = size / 2, H = size / 4;
    std::vector<float> tw((size_t)2 * H), roots((size_t)2 * H);
    for (int n = 0; n < H; n++) {
      double a = -M_PI * ((double)n + 0.125) / (double)M;
      tw[2 * n] = (float)cos(a);
      tw[2 * n + 1] = (float)sin(a);
      double b = -2.0 * M_PI * (double)n / (double)H;
      roots[2 * n] = (float)cos(b);
      roots[2 * n + 1] = (float)sin(b);
    }
    bw.align(4);
    h.tw_off[w] = bw.put_floats(tw.data(), tw.size());
    bw.align(4);
    h.fft_off[w] = bw.put_floats(roots.data(), roots.size());
  }
  bw.align(4);
  h.db_off = bw.put_floats(reinterpret_cast<const float*>(k_db_bits), 256);
  {
    std::vector<uint32_t> rcp((size_t)VPZ_RCP_MAX + 1, 0u);
    for (uint32_t d = 2; d <= VPZ_RCP_MAX; d++) rcp[d] = 0xffffffffu / d + 1u;   // ceil(2^32 / d): d never divides 2^32 - 1 + 1 unevenly here
    h.rcp_off = bw.put_struct_array(rcp);
  }
  // ---- K1a shared-memory table plan (vpz_dev.h, VpzSetupHdr.k1a_stage_off) ------------------------------
  for (int flag = 0; flag < 2; flag++) {
    std::vector<int> list;
    std::vector<char> seen((size_t)nbooks, 0);
    auto add = [&](int bk) {
      if (bk < 0 || bk >= nbooks || seen[(size_t)bk]) return;
      seen[(size_t)bk] = 1;
      list.push_back(bk);
    };
    for (int pass = 0; pass < 3; pass++)
      for (const VpzMode& md : modes) {
        if ((md.block_flag != 0) != (flag != 0)) continue;
        const VpzMapping& mp = mappings[md.mapping];
        for (int j = 0; j < mp.submaps; j++) {
          const VpzResidue& r = residues[mp.submap_residue[j]];
          if (pass == 0) {
            for (int c = 0; c < r.classifications; c++)
              for (int st = 0; st < 8; st++)
                if (((r.cascade[c] >> st) & 1) && r.has_books[c]) add(r.books[c][st]);
          } else if (pass == 1) {
            add(r.class_book);
          } else {
            const VpzFloor1& fl = floors[mp.submap_floor[j]];
            if (fl.floor_type == 1) {
              for (int i = 0; i < fl.partitions; i++) {
                const int c = fl.part_class[i];
                if (fl.class_sub[c] > 0) add(fl.class_master[c]);
                for (int k = 0; k < (1 << fl.class_sub[c]); k++) add(fl.sub_books[c][k]);
              }
            } else {
              for (int b2 = 0; b2 < fl.f0.nbooks; b2++) add(fl.f0.books[b2]);
            }
          }
        }
      }
    std::vector<uint32_t> plan;
    std::vector<uint16_t> soff(256, (uint16_t)K1A_SM_NONE);
    uint32_t cur = 0;
    for (int bk : list) {
      const uint32_t words = 1u << dbooks[(size_t)bk].l1_bits;
      if (bk > 255 || cur + words > K1A_SM_WORDS) continue;
      soff[(size_t)bk] = (uint16_t)cur;
      plan.push_back(dbooks[(size_t)bk].l1_off);
      plan.push_back(words);
      plan.push_back(cur);
      cur += words;
    }
    bw.align(4);
    h.k1a_stage_off[flag] = bw.reserve(2 + plan.size());
    blob[h.k1a_stage_off[flag]] = (uint32_t)(plan.size() / 3);
    blob[h.k1a_stage_off[flag] + 1] = cur;
    if (!plan.empty()) memcpy(&blob[h.k1a_stage_off[flag] + 2], plan.data(), plan.size() * 4);
    bw.align(4);
    h.k1a_soff_off[flag] = bw.reserve(128);
    memcpy(&blob[h.k1a_soff_off[flag]], soff.data(), 512);
  }
  bw.align(4);
  h.total_words = bw.here();
  memcpy(blob.data(), &h, sizeof(h));
  return VPZ_OK;
}

PacketGeom Setup::packet_geometry(const uint8_t* pkt, size_t len) const {
  PacketGeom g;
  BitReader br(pkt, len);
  if (br.read(1) != 0) return g;  // StreamDecoder.cs:728: not an audio packet
  int mode = (int)br.read(mode_bits);
  if ((size_t)mode >= modes.size()) {  // StreamDecoder.cs:732-735
    g.bad_mode = true;
    return g;
  }
  g.mode = mode;
  bool lb = modes[mode].block_flag != 0;
  bool prev = true, next = true;
  if (lb) {  // Mode.cs:38: ReadBit() is false past the end of the packet
    prev = br.read_bit();
    next = br.read_bit();
  }
  compute_geometry(id.size0, id.size1, lb, prev, next, &g);
  g.valid = true;
  return g;
}

}  // namespace vpz
