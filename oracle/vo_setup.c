/* vo_setup.c -- TEST INFRASTRUCTURE ONLY (see vorbis_oracle.h).
 * Restates the header parsing + table construction of the reference:
 *   StreamDecoder.LoadStreamHeader / LoadBooks  (StreamDecoder.cs:213-355)
 *   Codebook ctor / InitTree / ComputeCodewords / InitLookupTable (Codebook.cs:21-298)
 *   Huffman.GenerateTable (Huffman.cs:24-105)
 *   Floor1 ctor (Floor1.cs:39-155), Residue0 ctor (Residue0.cs:25-115),
 *   Mapping ctor (Mapping.cs:19-95), Mode ctor (Mode.cs:14-28),
 *   BlocksizeDerivedCache (BlocksizeDerivedCache.cs:9-35).
 */
#include "vo_internal.h"

/* ---------------------------------------------------------------- huffman -- */

/* HuffmanListNode.CompareTo (Contracts/HuffmanListNode.cs:12-20).  The
 * reference subtracts ints; a three-way compare gives the same order for every
 * code shorter than 32 bits. */
static int hnode_cmp(const void* a, const void* b) {
  const vo_hnode* x = (const vo_hnode*)a;
  const vo_hnode* y = (const vo_hnode*)b;
  if (x->length != y->length) return x->length < y->length ? -1 : 1;
  if (x->bits != y->bits) return x->bits < y->bits ? -1 : 1;
  return 0;
}

/* Huffman.GenerateTable (Huffman.cs:24-105) */
static int huff_generate(vo_book* bk, const int* values, const int* length_list, const int* code_list,
                         int count_in) {
  vo_hnode* list = (vo_hnode*)calloc((size_t)(count_in > 0 ? count_in : 1), sizeof(vo_hnode));
  if (!list) return VO_E_NOMEM;
  int nonzero = 0, last_valid = -1, max_len = 0;
  for (int i = 0; i < count_in; i++) {
    int l = length_list[i];
    if (l != 0) {
      nonzero++;
      last_valid = i;
    }
    list[i].value = values ? values[i] : i;
    list[i].length = l <= 0 ? 99999 : l;
    list[i].bits = code_list[i];
    list[i].mask = (int32_t)((1u << (l & 31)) - 1u);
    if (l > 0 && l > max_len) max_len = l;
  }
  if (nonzero == 1 && length_list[last_valid] != 1) {
    free(list);
    return VO_E_INVALID_DATA; /* "Invalid single entry." */
  }
  qsort(list, (size_t)count_in, sizeof(vo_hnode), hnode_cmp);

  int table_bits = max_len > 10 ? 10 : max_len; /* MAX_TABLE_BITS, Huffman.cs:12 */
  bk->prefix_bits = table_bits;
  bk->prefix = (vo_hnode*)calloc((size_t)1 << table_bits, sizeof(vo_hnode));
  bk->overflow = NULL;
  bk->overflow_n = 0;
  if (!bk->prefix) {
    free(list);
    return VO_E_NOMEM;
  }
  for (int i = 0; i < count_in && list[i].length < 99999; i++) {
    int item_bits = list[i].length;
    if (item_bits > table_bits) {
      int n = 0;
      for (int j = i; j < count_in && list[j].length < 99999; j++) n++;
      bk->overflow = (vo_hnode*)malloc(sizeof(vo_hnode) * (size_t)n);
      if (!bk->overflow) {
        free(list);
        return VO_E_NOMEM;
      }
      memcpy(bk->overflow, list + i, sizeof(vo_hnode) * (size_t)n);
      bk->overflow_n = n;
      break;
    }
    int reps = 1 << (table_bits - item_bits);
    for (int j = 0; j < reps; j++) bk->prefix[(j << item_bits) | list[i].bits] = list[i];
  }
  free(list);
  return VO_OK;
}

/* Codebook.ComputeCodewords (Codebook.cs:147-218): stb-style assignment, codes
 * stored bit-reversed so they match LSB-first peeked bits. */
static int compute_codewords(int sparse, int* codewords, int* codeword_lengths, const int* len, int n,
                             int* values) {
  uint32_t available[33];
  memset(available, 0, sizeof(available));
  int k, m = 0;
  for (k = 0; k < n; ++k)
    if (len[k] > 0) break;
  if (k == n) return 1;

#define ADD_ENTRY(code, sym, cnt, l)          \
  do {                                        \
    if (sparse) {                             \
      codewords[cnt] = (int)(code);           \
      codeword_lengths[cnt] = (l);            \
      values[cnt] = (sym);                    \
    } else {                                  \
      codewords[sym] = (int)(code);           \
    }                                         \
  } while (0)

  ADD_ENTRY(0u, k, m, len[k]);
  m++;
  for (int i = 1; i <= len[k]; ++i) available[i] = 1u << (32 - i);
  for (int i = k + 1; i < n; ++i) {
    int z = len[i];
    if (z <= 0) continue;
    while (z > 0 && available[z] == 0) --z;
    if (z == 0) return 0;
    uint32_t res = available[z];
    available[z] = 0;
    ADD_ENTRY(vo_bitrev32(res), i, m, len[i]);
    m++;
    if (z != len[i])
      for (int y = len[i]; y > z; --y) available[y] = res + (1u << (32 - y));
  }
#undef ADD_ENTRY
  return 1;
}

/* Codebook.lookup1_values (Codebook.cs:290-298) */
static int lookup1_values(int entries, int dims) {
  int r = (int)floor(exp(log((double)entries) / dims));
  if (floor(pow((double)r + 1, dims)) <= entries) ++r;
  return r;
}

/* Utils.ConvertFromVorbisFloat32 (Utils.cs:92-105) */
static float vorbis_float32(uint32_t bits) {
  int32_t sign = (int32_t)bits >> 31;
  int exponent = (int)((bits & 0x7fe00000u) >> 21) - 788;
  float mantissa = (float)((((int32_t)(bits & 0x1fffff)) ^ sign) + (sign & 1));
  return scalbnf(mantissa, exponent);
}

/* Codebook ctor (Codebook.cs:21-42) + InitTree (:44-144) + InitLookupTable (:220-288) */
static int book_parse(vo_book* bk, vo_bits* br) {
  memset(bk, 0, sizeof(*bk));
  if (vo_read_bits(br, 24) != 0x564342u) return VO_E_INVALID_DATA;
  bk->dims = (int)vo_read_bits(br, 16);
  int entries = bk->entries = (int)vo_read_bits(br, 24);
  bk->lengths = (int*)calloc((size_t)(entries > 0 ? entries : 1), sizeof(int));
  if (!bk->lengths) return VO_E_NOMEM;

  int sparse, total = 0, max_len;
  if (vo_read_bit(br)) { /* ordered */
    int len = (int)vo_read_bits(br, 5) + 1;
    for (int i = 0; i < entries;) {
      int cnt = (int)vo_read_bits(br, vo_ilog(entries - i));
      while (--cnt >= 0) {
        if (i >= entries) return VO_E_INVALID_DATA; /* reference: IndexOutOfRange */
        bk->lengths[i++] = len;
      }
      ++len;
      if (br->pos >= br->total_bits && i < entries) return VO_E_INVALID_DATA; /* runaway guard */
    }
    total = 0;
    sparse = 0;
    max_len = len; /* quirk Q7: one more than the last length used */
  } else {
    max_len = -1;
    sparse = vo_read_bit(br);
    for (int i = 0; i < entries; i++) {
      if (!sparse || vo_read_bit(br)) {
        bk->lengths[i] = (int)vo_read_bits(br, 5) + 1;
        ++total;
      } else {
        bk->lengths[i] = -1;
      }
      if (bk->lengths[i] > max_len) max_len = bk->lengths[i];
    }
  }

  if (max_len <= -1) {
    bk->max_bits = 0; /* Huffman.Empty */
    bk->prefix_bits = 0;
  } else {
    bk->max_bits = max_len;
    int* codeword_lengths = NULL;
    if (sparse && total >= (entries >> 2)) {
      codeword_lengths = (int*)malloc(sizeof(int) * (size_t)entries);
      memcpy(codeword_lengths, bk->lengths, sizeof(int) * (size_t)entries);
      sparse = 0;
    }
    int sorted_count = sparse ? total : 0;
    int *values = NULL, *codewords = NULL;
    int list_n = entries;
    if (!sparse) {
      codewords = (int*)calloc((size_t)(entries > 0 ? entries : 1), sizeof(int));
    } else if (sorted_count != 0) {
      codeword_lengths = (int*)calloc((size_t)sorted_count, sizeof(int));
      codewords = (int*)calloc((size_t)sorted_count, sizeof(int));
      values = (int*)calloc((size_t)sorted_count, sizeof(int));
      list_n = sorted_count;
    }
    int ok = compute_codewords(sparse, codewords, codeword_lengths, bk->lengths, entries, values);
    int rc = VO_OK;
    if (!ok || !codewords) {
      rc = VO_E_INVALID_DATA;
    } else {
      const int* length_list = codeword_lengths ? codeword_lengths : bk->lengths;
      rc = huff_generate(bk, values, length_list, codewords, list_n);
    }
    free(codeword_lengths);
    free(values);
    free(codewords);
    if (rc != VO_OK) return rc;
  }

  bk->map_type = (int)vo_read_bits(br, 4);
  if (bk->map_type == 0) return VO_OK;

  float min_value = vorbis_float32((uint32_t)vo_read_bits(br, 32));
  float delta_value = vorbis_float32((uint32_t)vo_read_bits(br, 32));
  int value_bits = (int)vo_read_bits(br, 4) + 1;
  int sequence_p = vo_read_bit(br);
  int dims = bk->dims;
  int64_t lookup_n = (int64_t)entries * dims;
  int mult_n = (int)lookup_n;
  if (bk->map_type == 1) mult_n = lookup1_values(entries, dims);
  if (lookup_n > (1 << 26) || mult_n < 0) return VO_E_INVALID_DATA;
  bk->lookup = (float*)calloc((size_t)(lookup_n > 0 ? lookup_n : 1), sizeof(float));
  uint16_t* mult = (uint16_t*)calloc((size_t)(mult_n > 0 ? mult_n : 1), sizeof(uint16_t));
  if (!bk->lookup || !mult) return VO_E_NOMEM;
  for (int i = 0; i < mult_n; i++) mult[i] = (uint16_t)vo_read_bits(br, value_bits);

  if (bk->map_type == 1) {
    for (int idx = 0; idx < entries; idx++) {
      float last = 0.f;
      uint32_t idx_div = 1;
      for (int i = 0; i < dims; i++) {
        uint32_t moff = (uint32_t)idx / idx_div % (uint32_t)mult_n;
        float value = (float)mult[moff] * delta_value;
        value = value + min_value;
        value = value + last;
        bk->lookup[(size_t)idx * dims + i] = value;
        if (sequence_p) last = value;
        idx_div *= (uint32_t)mult_n;
      }
    }
  } else {
    for (int idx = 0; idx < entries; idx++) {
      float last = 0.f;
      for (int i = 0; i < dims; i++) {
        float value = (float)mult[(size_t)idx * dims + i] * delta_value;
        value = value + min_value;
        value = value + last;
        bk->lookup[(size_t)idx * dims + i] = value;
        if (sequence_p) last = value;
      }
    }
  }
  free(mult);
  return VO_OK;
}

/* Codebook.DecodeScalar + DecodeOverflowScalar (Codebook.cs:301-335) */
int vo_book_decode_scalar(const vo_book* bk, vo_bits* br) {
  int got;
  uint64_t data = vo_peek(br, bk->prefix_bits, &got);
  if (got != 0) {
    const vo_hnode* node = &bk->prefix[data];
    if (node->length != 0) {
      vo_skip(br, node->length);
      return node->value;
    }
  }
  int32_t wide = (int32_t)vo_peek(br, bk->max_bits, &got);
  if (got != 0) {
    for (int i = 0; i < bk->overflow_n; i++) {
      const vo_hnode* node = &bk->overflow[i];
      if (node->bits == (wide & node->mask)) {
        vo_skip(br, node->length);
        return node->value;
      }
    }
  }
  return -1;
}

/* ----------------------------------------------------------------- floor1 -- */
/* Floor0.ToBARK (Floor0.cs:98-101): double arithmetic, rounded to float */
static float floor0_to_bark(double lsp) {
  return (float)(13.1 * atan(0.00074 * lsp) + 2.24 * atan(0.0000000185 * lsp * lsp) + .0001 * lsp);
}

/* Floor0 ctor (Floor0.cs:39-76), SynthesizeBarkCurve (:83-96), SynthesizeWDelMap (:103-113) */
static int floor0_parse(vo_floor1* f, vo_bits* br, const vo_setup* st) {
  memset(f, 0, sizeof(*f));
  f->type = 0;
  f->order = (int)vo_read_bits(br, 8);
  f->rate = (int)vo_read_bits(br, 16);
  f->bark_map_size = (int)vo_read_bits(br, 16);
  f->amp_bits = (int)vo_read_bits(br, 6);
  f->amp_ofs = (int)vo_read_bits(br, 8);
  f->nbooks0 = (int)vo_read_bits(br, 4) + 1;
  if (f->order < 1 || f->rate < 1 || f->bark_map_size < 1) return VO_E_INVALID_DATA;
  for (int i = 0; i < f->nbooks0; i++) {
    int num = (int)vo_read_bits(br, 8);
    if (num >= st->nbooks) return VO_E_INVALID_DATA;
    if (st->books[num].map_type == 0 || st->books[num].dims < 1) return VO_E_INVALID_DATA;
    f->books0[i] = (uint8_t)num;
  }
  for (int w = 0; w < 2; w++) {
    const int n = (w ? st->size1 : st->size0) / 2;
    /* float scale = _bark_map_size / ToBARK(_rate / 2.0) : ushort / float in fp32 */
    const float scale = (float)f->bark_map_size / floor0_to_bark(f->rate / 2.0);
    int* map = (int*)calloc((size_t)n + 1, sizeof(int));
    for (int i = 0; i < n + 1 - 2; i++) { /* i < map.Length - 2: entry n-1 keeps its default 0 */
      const float prod = floor0_to_bark((f->rate / 2.0) / n * i) * scale; /* float * float */
      int v = (int)floor((double)prod);
      map[i] = v < f->bark_map_size - 1 ? v : f->bark_map_size - 1;
    }
    map[n] = -1;
    f->bark_map[w] = map;
    const float wdel = (float)(3.14159265358979323846 / f->bark_map_size);
    float* wm = (float*)calloc((size_t)n, sizeof(float));
    for (int i = 0; i < n; i++) {
      const float a = wdel * (float)i;
      wm[i] = 2.0f * cosf(a);
    }
    f->wmap[w] = wm;
    /* Apply reads wMap[barkMap[i]] (Floor0.cs:192): a bark index >= n is an IndexOutOfRangeException there */
    for (int i = 0; i < n; i++)
      if (map[i] >= n) return VO_E_REF_FAULT;
  }
  return VO_OK;
}

static int floor1_parse(vo_floor1* f, vo_bits* br, int nbooks) {
  static const uint8_t range_lookup[4] = {128, 64, 43, 32};
  static const uint8_t ybits_lookup[4] = {8, 7, 7, 6};
  memset(f, 0, sizeof(*f));
  f->type = 1;
  int maximum_class = -1;
  f->partitions = (int)vo_read_bits(br, 5);
  for (int i = 0; i < f->partitions; i++) {
    f->part_class[i] = (uint8_t)vo_read_bits(br, 4);
    if (f->part_class[i] > maximum_class) maximum_class = f->part_class[i];
  }
  f->class_count = maximum_class + 1;
  for (int i = 0; i < f->class_count; i++) {
    f->class_dim[i] = (uint8_t)(vo_read_bits(br, 3) + 1);
    f->class_sub[i] = (uint8_t)vo_read_bits(br, 2);
    if (f->class_sub[i] > 0) f->class_master[i] = (uint8_t)vo_read_bits(br, 8);
    int nsub = 1 << f->class_sub[i];
    for (int j = 0; j < nsub; j++) {
      int book_num = (int)vo_read_bits(br, 8) - 1;
      if (book_num >= nbooks) return VO_E_INVALID_DATA;
      f->sub_books[i][j] = (int16_t)book_num;
    }
  }
  int multiplier = (int)vo_read_bits(br, 2);
  f->range = range_lookup[multiplier] * 2;
  f->ybits = ybits_lookup[multiplier];
  f->multiplier = multiplier + 1;
  int range_bits = (int)vo_read_bits(br, 4);
  int n = 2;
  for (int i = 0; i < f->partitions; i++) n += f->class_dim[f->part_class[i]];
  if (n > 256) return VO_E_INVALID_DATA;
  f->xcount = n;
  int k = 0;
  f->xlist[k++] = 0;
  f->xlist[k++] = 1 << range_bits;
  for (int i = 0; i < f->partitions; i++)
    for (int j = 0; j < f->class_dim[f->part_class[i]]; j++) f->xlist[k++] = (int)vo_read_bits(br, range_bits);

  /* low / high neighbours among earlier posts (Floor1.cs:109-133) */
  f->sortidx[0] = 0;
  f->sortidx[1] = 1;
  for (int i = 2; i < n; i++) {
    f->lneigh[i] = 0;
    f->hneigh[i] = 1;
    f->sortidx[i] = i;
    for (int j = 2; j < i; j++) {
      int t = f->xlist[j];
      if (t < f->xlist[i]) {
        if (t > f->xlist[f->lneigh[i]]) f->lneigh[i] = j;
      } else {
        if (t < f->xlist[f->hneigh[i]]) f->hneigh[i] = j;
      }
    }
  }
  /* exchange sort of the index by X; duplicate X is an error (Floor1.cs:136-149) */
  for (int i = 0; i < n - 1; i++) {
    for (int j = i + 1; j < n; j++) {
      if (f->xlist[i] == f->xlist[j]) return VO_E_INVALID_DATA;
      if (f->xlist[f->sortidx[i]] > f->xlist[f->sortidx[j]]) {
        int t = f->sortidx[i];
        f->sortidx[i] = f->sortidx[j];
        f->sortidx[j] = t;
      }
    }
  }
  return VO_OK;
}

/* ---------------------------------------------------------------- residue -- */
static int residue_parse(vo_residue* r, int type, vo_bits* br, const vo_book* books, int nbooks) {
  memset(r, 0, sizeof(*r));
  r->type = type;
  r->begin = (int)vo_read_bits(br, 24);
  r->end = (int)vo_read_bits(br, 24);
  r->part_size = (int)vo_read_bits(br, 24) + 1;
  r->classifications = (int)vo_read_bits(br, 6) + 1;
  r->class_book = (int)vo_read_bits(br, 8);
  int acc = 0;
  for (int i = 0; i < r->classifications; i++) {
    uint32_t low = (uint32_t)vo_read_bits(br, 4);
    uint32_t bits = low & 7u;
    if (low & 8u) bits |= (uint32_t)vo_read_bits(br, 5) << 3;
    r->cascade[i] = (uint8_t)bits;
    acc += __builtin_popcount(bits);
  }
  uint8_t* book_nums = (uint8_t*)malloc((size_t)(acc > 0 ? acc : 1));
  for (int i = 0; i < acc; i++) {
    book_nums[i] = (uint8_t)vo_read_bits(br, 8);
    if (book_nums[i] >= nbooks || books[book_nums[i]].map_type == 0) {
      free(book_nums);
      return VO_E_INVALID_DATA;
    }
  }
  if (r->class_book >= nbooks) {
    free(book_nums);
    return VO_E_INVALID_DATA;
  }
  const vo_book* cb = &books[r->class_book];
  int partvals = 1;
  for (int i = 0; i < cb->dims; i++) {
    partvals *= r->classifications;
    if (partvals > cb->entries) {
      free(book_nums);
      return VO_E_INVALID_DATA;
    }
  }
  acc = 0;
  int maxstage = 0;
  for (int j = 0; j < r->classifications; j++) {
    for (int k = 0; k < 8; k++) r->books[j][k] = -1;
    int stages = vo_ilog(r->cascade[j]);
    if (stages <= 0) continue;
    r->has_books[j] = 1;
    if (stages > maxstage) maxstage = stages;
    for (int k = 0; k < stages; k++) {
      /* unset stages keep book 0 in the reference's byte[] (Residue0.cs:88-95) */
      r->books[j][k] = (r->cascade[j] & (1 << k)) ? (int16_t)book_nums[acc++] : 0;
    }
  }
  free(book_nums);
  r->max_stages = maxstage;
  r->decode_map_len = partvals * cb->dims;
  r->decode_map = (int*)calloc((size_t)(r->decode_map_len > 0 ? r->decode_map_len : 1), sizeof(int));
  for (int j = 0; j < partvals; j++) {
    int val = j;
    int mult = partvals / r->classifications;
    for (int k = 0; k < cb->dims; k++) {
      int deco = val / mult;
      val -= deco * mult;
      mult /= r->classifications;
      r->decode_map[j * cb->dims + k] = deco;
    }
  }
  return VO_OK;
}

/* ---------------------------------------------------------------- mapping -- */
static int mapping_parse(vo_mapping* m, vo_bits* br, int channels, int nfloors, int nresidues) {
  memset(m, 0, sizeof(*m));
  m->submaps = 1;
  if (vo_read_bit(br)) m->submaps += (int)vo_read_bits(br, 4);
  m->coupling_steps = 0;
  if (vo_read_bit(br)) m->coupling_steps = (int)vo_read_bits(br, 8) + 1;
  int coupling_bits = vo_ilog(channels - 1);
  for (int j = 0; j < m->coupling_steps; j++) {
    int mag = (int)vo_read_bits(br, coupling_bits);
    int ang = (int)vo_read_bits(br, coupling_bits);
    if (mag == ang || mag > channels - 1 || ang > channels - 1) return VO_E_INVALID_DATA;
    m->ang[j] = (uint8_t)ang;
    m->mag[j] = (uint8_t)mag;
  }
  if (vo_read_bits(br, 2) != 0) return VO_E_INVALID_DATA;
  if (m->submaps > 1) {
    for (int c = 0; c < channels; c++) {
      m->mux[c] = (uint8_t)vo_read_bits(br, 4);
      if (m->mux[c] > m->submaps) return VO_E_INVALID_DATA;
      if (m->mux[c] >= m->submaps) return VO_E_INVALID_DATA; /* reference: index out of range later */
    }
  }
  for (int j = 0; j < m->submaps; j++) {
    vo_read_bits(br, 8); /* SkipBits(8): unused time configuration placeholder */
    int fl = (int)vo_read_bits(br, 8);
    if (fl >= nfloors) return VO_E_INVALID_DATA;
    int rs = (int)vo_read_bits(br, 8);
    if (rs >= nresidues) return VO_E_INVALID_DATA;
    m->submap_floor[j] = (uint8_t)fl;
    m->submap_residue[j] = (uint8_t)rs;
  }
  return VO_OK;
}

/* BlocksizeDerivedCache.CalcWindowSlope (BlocksizeDerivedCache.cs:24-35) */
void vo_window_slope(float* slope, int n) {
  const float half_pi = 0.5f * 3.14159274f; /* 0.5f * MathF.PI, folded in fp32 */
  for (int i = 0; i < n; i++) {
    float a = half_pi * ((float)i + 0.5f);
    a = a / (float)n;
    float v = sinf(a);
    float b = half_pi * v;
    b = b * v;
    slope[i] = sinf(b);
  }
}

/* --------------------------------------------------------------- id + setup -- */
static int check_sig(vo_bits* br, const uint8_t* sig, int n) {
  for (int i = 0; i < n; i++)
    if (vo_read_bits(br, 8) != sig[i]) return 0;
  return 1;
}

/* StreamDecoder.LoadStreamHeader (StreamDecoder.cs:213-240) */
int vo_setup_parse_id(vo_setup* st, const uint8_t* pkt, int len) {
  static const uint8_t sig[11] = {0x01, 0x76, 0x6f, 0x72, 0x62, 0x69, 0x73, 0, 0, 0, 0};
  vo_bits br;
  vo_bits_init(&br, pkt, len);
  if (!check_sig(&br, sig, 11)) return VO_E_INVALID_DATA;
  st->channels = (int)vo_read_bits(&br, 8);
  st->sample_rate = (int)vo_read_bits(&br, 32);
  st->br_upper = (int)vo_read_bits(&br, 32);
  st->br_nominal = (int)vo_read_bits(&br, 32);
  st->br_lower = (int)vo_read_bits(&br, 32);
  int b0 = (int)vo_read_bits(&br, 4), b1 = (int)vo_read_bits(&br, 4);
  st->size0 = 1 << b0;
  st->size1 = 1 << b1;
  if (st->br_nominal == 0 && st->br_upper > 0 && st->br_lower > 0)
    st->br_nominal = (st->br_upper + st->br_lower) / 2;
  if (st->channels < 1 || st->channels > VO_MAX_CH) return VO_E_UNSUPPORTED;
  st->slope[0] = (float*)malloc(sizeof(float) * (size_t)(st->size0 / 2 + 1));
  st->slope[1] = (float*)malloc(sizeof(float) * (size_t)(st->size1 / 2 + 1));
  vo_window_slope(st->slope[0], st->size0 / 2);
  vo_window_slope(st->slope[1], st->size1 / 2);
  return VO_OK;
}

/* StreamDecoder.LoadBooks (StreamDecoder.cs:262-321) */
int vo_setup_parse_books(vo_setup* st, const uint8_t* pkt, int len) {
  static const uint8_t sig[7] = {0x05, 0x76, 0x6f, 0x72, 0x62, 0x69, 0x73};
  vo_bits br;
  vo_bits_init(&br, pkt, len);
  if (!check_sig(&br, sig, 7)) return VO_E_INVALID_DATA;
  int rc;

  st->nbooks = (int)vo_read_bits(&br, 8) + 1;
  st->books = (vo_book*)calloc((size_t)st->nbooks, sizeof(vo_book));
  for (int i = 0; i < st->nbooks; i++)
    if ((rc = book_parse(&st->books[i], &br)) != VO_OK) return rc;

  int times = (int)vo_read_bits(&br, 6) + 1;
  vo_skip(&br, 16 * times);

  st->nfloors = (int)vo_read_bits(&br, 6) + 1;
  st->floors = (vo_floor1*)calloc((size_t)st->nfloors, sizeof(vo_floor1));
  for (int i = 0; i < st->nfloors; i++) {
    int type = (int)vo_read_bits(&br, 16);
    if (type == 0) {
      if ((rc = floor0_parse(&st->floors[i], &br, st)) != VO_OK) return rc;
      continue;
    }
    if (type != 1) return VO_E_INVALID_DATA;
    if ((rc = floor1_parse(&st->floors[i], &br, st->nbooks)) != VO_OK) return rc;
    if (st->floors[i].xcount > 64) return VO_E_UNSUPPORTED; /* Posts[64], quirk Q2 */
  }

  st->nresidues = (int)vo_read_bits(&br, 6) + 1;
  st->residues = (vo_residue*)calloc((size_t)st->nresidues, sizeof(vo_residue));
  for (int i = 0; i < st->nresidues; i++) {
    int type = (int)vo_read_bits(&br, 16);
    if (type < 0 || type > 2) return VO_E_INVALID_DATA;
    if ((rc = residue_parse(&st->residues[i], type, &br, st->books, st->nbooks)) != VO_OK) return rc;
  }

  st->nmappings = (int)vo_read_bits(&br, 6) + 1;
  st->mappings = (vo_mapping*)calloc((size_t)st->nmappings, sizeof(vo_mapping));
  for (int i = 0; i < st->nmappings; i++) {
    if (vo_read_bits(&br, 16) != 0) return VO_E_INVALID_DATA;
    if ((rc = mapping_parse(&st->mappings[i], &br, st->channels, st->nfloors, st->nresidues)) != VO_OK)
      return rc;
  }

  st->nmodes = (int)vo_read_bits(&br, 6) + 1;
  st->modes = (vo_mode*)calloc((size_t)st->nmodes, sizeof(vo_mode));
  for (int i = 0; i < st->nmodes; i++) {
    st->modes[i].block_flag = vo_read_bit(&br);
    if (vo_read_bits(&br, 32) != 0) return VO_E_INVALID_DATA;
    st->modes[i].mapping = (int)vo_read_bits(&br, 8);
    if (st->modes[i].mapping >= st->nmappings) return VO_E_INVALID_DATA;
  }
  if (!vo_read_bit(&br)) return VO_E_INVALID_DATA; /* framing bit */
  st->mode_bits = vo_ilog(st->nmodes - 1);
  return VO_OK;
}

void vo_setup_free(vo_setup* st) {
  for (int i = 0; i < st->nbooks && st->books; i++) {
    free(st->books[i].lengths);
    free(st->books[i].prefix);
    free(st->books[i].overflow);
    free(st->books[i].lookup);
  }
  for (int i = 0; i < st->nresidues && st->residues; i++) {
    free(st->residues[i].decode_map);
    free(st->residues[i].part_word_cache);
  }
  for (int i = 0; i < st->nfloors && st->floors; i++)
    for (int w = 0; w < 2; w++) {
      free(st->floors[i].bark_map[w]);
      free(st->floors[i].wmap[w]);
    }
  free(st->books);
  free(st->floors);
  free(st->residues);
  free(st->mappings);
  free(st->modes);
  free(st->slope[0]);
  free(st->slope[1]);
  memset(st, 0, sizeof(*st));
}
