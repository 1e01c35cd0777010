/* vo_stream.c -- TEST INFRASTRUCTURE ONLY (see vorbis_oracle.h).
 * Restates, over an in-memory file:
 *   Ogg physical layer: PageReaderBase.ReadNextPage / VerifyPage / VerifyHeader
 *       (Ogg/PageReaderBase.cs:41-84,176-212,286-361), PageReader.AddPage (Ogg/PageReader.cs:58-102),
 *       Crc (Ogg/Crc.cs:20-63, polynomial Ogg/Crc.Table.cs:14), PageHeader.GetPacketCount
 *       (Ogg/PageHeader.cs:35-59), PageData.GetPacket (Ogg/PageData.cs:53-83)
 *   Ogg logical layer: StreamPageReader.AddPage / FindPage (Ogg/StreamPageReader.cs:44-110,152-305),
 *       PacketProvider.CreatePacket / SeekTo / GetTargetPageInfo / FillPageEndGranuleCache /
 *       NormalizePacketIndex / GetGranuleCount (Ogg/PacketProvider.cs:35-560)
 *   Stream decoder: StreamDecoder.Read / ReadNextPacket / DecodeNextPacket / OverlapBuffers /
 *       StoreInterleaved / StoreContiguous / SeekTo / GetPacketGranuleCount
 *       (StreamDecoder.cs:418-498,515-638,640-791,817-913), Utils.ClipValue (Utils.cs:44-58)
 * Pages are "read" lazily (pages_loaded) because HasAllPages / IsEndOfStream in the
 * reference depend on how far the reader has got.
 */
#include "vo_internal.h"

typedef struct {
  int64_t offset;        /* byte offset in the file */
  int64_t granule;
  uint32_t seq;
  uint8_t flags;         /* 1 continuation, 2 BOS, 4 EOS */
  int nseg;
  const uint8_t* seg;    /* lacing table */
  const uint8_t* body;
  int body_len;
  int is_resync;         /* stored as negative page offset in the reference */
  int packet_count;
  int is_continued;
} vo_page;

typedef struct {
  uint8_t* data; /* assembled, zero padded */
  int len;
  int valid;
  int is_resync, is_eos;
  int64_t granule;
  int page_index, packet_index;
} vo_packet;

enum { EOS_NONE = 0, EOS_INVALID_PACKET = 1, EOS_PACKET_FLAG = 2, EOS_INVALID_PREROLL = 4 };

struct vo_stream {
  const uint8_t* file;
  size_t file_len;
  uint32_t serial;
  /* physical accounting */
  int64_t container_bits, waste_bits;
  int crc_failures;
  /* logical stream pages (everything the physical reader would hand to this serial) */
  vo_page* pages;
  int npages, pages_cap;
  int pages_loaded;     /* lazily "read" prefix */
  int has_all_pages;
  int first_data_page;  /* -1 unknown */
  int64_t max_granule;
  /* packet provider cursor */
  int64_t page_index;
  int packet_index;
  int64_t* page_end_granules;
  int page_end_n, page_end_cap;
  /* headers */
  vo_packet hdr[3];
  char* vendor;
  int vendor_len;
  char** comments;
  int* comment_lens;
  int ncomments;
  vo_setup setup;
  /* decoder state (StreamDecoder.cs:40-49) */
  int clip;
  int64_t current_position;
  int has_clipped, has_position, eos_found;
  float *next_buf, *prev_buf, *pool[2];
  int pool_n;
  int prev_start, prev_end, prev_stop;
  int fault; /* sticky VO_E_REF_FAULT */
  /* cached forward packet table for vo_audio_packet */
  vo_packet* table;
  int table_n;
  int table_built;
};

/* -------------------------------------------------------------------- CRC -- */
static uint32_t g_crc_table[256];
static int g_crc_ready;
static void crc_init(void) {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t r = i << 24;
    for (int j = 0; j < 8; j++) r = (r << 1) ^ ((r & 0x80000000u) ? 0x04c11db7u : 0);
    g_crc_table[i] = r;
  }
  g_crc_ready = 1;
}
uint32_t vo_crc_ogg(const uint8_t* data, size_t len, uint32_t crc) {
  if (!g_crc_ready) crc_init();
  for (size_t i = 0; i < len; i++) crc = (crc << 8) ^ g_crc_table[((crc >> 24) ^ data[i]) & 0xff];
  return crc;
}

/* ------------------------------------------------------- physical page scan -- */
static void page_counts(vo_page* p) {
  int cnt = 0;
  for (int i = 0; i < p->nseg; i++)
    if (p->seg[i] < 255) cnt++;
  p->is_continued = p->nseg > 0 && p->seg[p->nseg - 1] == 255;
  if (p->is_continued) cnt++;
  p->packet_count = cnt;
}

/* Tries to verify a page at `pos` (VerifyHeader + VerifyPage). Returns page length or 0. */
static size_t try_page(const uint8_t* f, size_t len, size_t pos, int* crc_fail) {
  if (pos + 27 > len) return 0;
  if (!(f[pos] == 0x4f && f[pos + 1] == 0x67 && f[pos + 2] == 0x67 && f[pos + 3] == 0x53)) return 0;
  int nseg = f[pos + 26];
  if (pos + 27 + (size_t)nseg > len) return 0;
  size_t body = 0;
  for (int i = 0; i < nseg; i++) body += f[pos + 27 + i];
  size_t total = 27 + (size_t)nseg + body;
  if (pos + total > len) return 0;
  uint32_t want = ((uint32_t)f[pos + 22]) | ((uint32_t)f[pos + 23] << 8) | ((uint32_t)f[pos + 24] << 16) |
                  ((uint32_t)f[pos + 25] << 24);
  static const uint8_t zero4[4] = {0, 0, 0, 0};
  uint32_t crc = vo_crc_ogg(f + pos, 22, 0);
  crc = vo_crc_ogg(zero4, 4, crc);
  crc = vo_crc_ogg(f + pos + 26, total - 26, crc);
  if (crc != want) {
    if (crc_fail) (*crc_fail)++;
    return 0;
  }
  return total;
}

static int push_page(vo_stream* s, const vo_page* p) {
  if (s->npages == s->pages_cap) {
    int nc = s->pages_cap ? s->pages_cap * 2 : 64;
    vo_page* np = (vo_page*)realloc(s->pages, sizeof(vo_page) * (size_t)nc);
    if (!np) return VO_E_NOMEM;
    s->pages = np;
    s->pages_cap = nc;
  }
  s->pages[s->npages++] = *p;
  return VO_OK;
}

/* Scans the whole file once.  The first serial whose first page passes becomes
 * the stream (VorbisReader.Initialize picks _decoders[0]); pages of other serials are
 * skipped.  Stream-level filtering that depends on laziness (HasAllPages) is applied in
 * load_page(). */
static int scan_file(vo_stream* s) {
  size_t pos = 0;
  int have_serial = 0;
  int resync = 0;
  int serial_ignored = 0;
  while (pos + 4 <= s->file_len) {
    size_t plen = try_page(s->file, s->file_len, pos, &s->crc_failures);
    if (!plen) {
      pos++;
      s->waste_bits += 8;
      resync = 1;
      continue;
    }
    const uint8_t* h = s->file + pos;
    vo_page p;
    memset(&p, 0, sizeof(p));
    p.offset = (int64_t)pos;
    p.flags = h[5];
    memcpy(&p.granule, h + 6, 8);
    uint32_t serial;
    memcpy(&serial, h + 14, 4);
    memcpy(&p.seq, h + 18, 4);
    p.nseg = h[26];
    p.seg = h + 27;
    p.body = h + 27 + p.nseg;
    p.body_len = (int)(plen - 27 - (size_t)p.nseg);
    p.is_resync = resync;
    page_counts(&p);
    resync = 0;
    pos += plen;
    if (!have_serial) {
      s->serial = serial;
      have_serial = 1;
    }
    if (serial != s->serial || serial_ignored) {
      s->waste_bits += (int64_t)plen * 8;
      continue;
    }
    if (p.packet_count == 0) {
      /* PageReader.AddPage refuses it and the serial lands in _ignoredSerials
       * (PageReaderBase.cs:86-102, PageReader.cs:63-67) */
      serial_ignored = 1;
      s->waste_bits += (int64_t)plen * 8;
      continue;
    }
    int rc = push_page(s, &p);
    if (rc != VO_OK) return rc;
  }
  if (pos < s->file_len) s->waste_bits += 8 * (int64_t)(s->file_len - pos);
  return have_serial ? VO_OK : VO_E_INVALID_DATA;
}

/* StreamPageReader.AddPage (Ogg/StreamPageReader.cs:44-110) applied when page `idx`
 * of the scan is first touched.  Returns <0 on InvalidDataException. */
static int g_last_seq_dummy;
static int load_pages_upto(vo_stream* s, int64_t idx) {
  (void)g_last_seq_dummy;
  while (s->pages_loaded <= idx && !s->has_all_pages) {
    if (s->pages_loaded >= s->npages) {
      s->has_all_pages = 1; /* SetEndOfStreams on physical end */
      break;
    }
    vo_page* p = &s->pages[s->pages_loaded];
    if (p->granule != -1) {
      if (s->first_data_page < 0 && p->granule > 0) {
        s->first_data_page = s->pages_loaded;
      } else if (s->max_granule > p->granule) {
        return VO_E_INVALID_DATA; /* "Granule Position regressed?!" */
      }
      s->max_granule = p->granule;
    } else if (s->first_data_page >= 0 && (!p->is_continued || p->packet_count != 1)) {
      return VO_E_INVALID_DATA;
    }
    if (p->flags & 4) s->has_all_pages = 1;
    if (s->pages_loaded > 0) {
      uint32_t last = s->pages[s->pages_loaded - 1].seq;
      if (last != 0 && last + 1 != p->seq) p->is_resync = 1;
    }
    s->container_bits += 8 * (27 + p->nseg);
    s->pages_loaded++;
  }
  return VO_OK;
}

/* IStreamPageReader.GetPage(index, out ...) (Ogg/StreamPageReader.cs:335-424) */
static const vo_page* get_page(vo_stream* s, int64_t idx) {
  if (idx < 0) return NULL;
  if (load_pages_upto(s, idx) != VO_OK) return NULL;
  if (idx < s->pages_loaded) return &s->pages[idx];
  return NULL;
}

static int page_count(vo_stream* s) { return s->pages_loaded; }

/* PageData.GetPacket (Ogg/PageData.cs:53-83) */
static void page_packet_slice(const vo_page* p, int packet_index, const uint8_t** data, int* len) {
  int pk = 0, ofs = 0, size = 0;
  for (int i = 0; i < p->nseg; i++) {
    size += p->seg[i];
    if (p->seg[i] < 255) {
      if (pk == packet_index) {
        *data = p->body + ofs;
        *len = size;
        return;
      }
      pk++;
      ofs += size;
      size = 0;
    }
  }
  if (pk == packet_index) {
    *data = p->body + ofs;
    *len = size;
    return;
  }
  *data = p->body;
  *len = 0;
}

static void packet_free(vo_packet* pk) {
  free(pk->data);
  memset(pk, 0, sizeof(*pk));
}

static void packet_append(vo_packet* pk, const uint8_t* d, int n) {
  pk->data = (uint8_t*)realloc(pk->data, (size_t)pk->len + (size_t)n + 16);
  if (n) memcpy(pk->data + pk->len, d, (size_t)n);
  pk->len += n;
  memset(pk->data + pk->len, 0, 16);
}

/* PacketProvider.CreatePacket (Ogg/PacketProvider.cs:427-560) */
static void create_packet(vo_stream* s, int64_t* page_index, int* packet_index, int advance,
                          int64_t granule_pos, int is_resync, int is_continued, int packet_count,
                          vo_packet* out) {
  memset(out, 0, sizeof(*out));
  const vo_page* first = get_page(s, *page_index);
  const uint8_t* d;
  int n;
  if (!first) return;
  out->page_index = (int)*page_index;
  out->packet_index = *packet_index;
  page_packet_slice(first, *packet_index, &d, &n);
  packet_append(out, d, n);

  int is_last;
  int64_t final_page = *page_index;
  if (is_continued && *packet_index == packet_count - 1) {
    int64_t cont = *page_index;
    while (is_continued) {
      const vo_page* np = get_page(s, ++cont);
      if (!np) {
        packet_free(out);
        return; /* default(VorbisPacket) */
      }
      granule_pos = np->granule;
      is_resync = np->is_resync;
      is_continued = np->is_continued;
      packet_count = np->packet_count;
      if (!(np->flags & 1) || is_resync) break;
      if (is_continued && packet_count > 1) is_continued = 0;
      page_packet_slice(np, 0, &d, &n);
      packet_append(out, d, n);
    }
    is_last = packet_count == 1;
    final_page = cont;
  } else {
    is_last = *packet_index == packet_count - 1;
  }
  out->valid = 1;
  out->is_resync = is_resync;
  if (is_last) {
    out->granule = granule_pos;
    if (s->has_all_pages && final_page == page_count(s) - 1) out->is_eos = 1;
  } else {
    out->granule = -1;
  }
  if (advance) {
    if (final_page != *page_index) {
      *page_index = final_page;
      *packet_index = 0;
    }
    if (*packet_index == packet_count - 1) {
      ++*page_index;
      *packet_index = 0;
    } else {
      ++*packet_index;
    }
  }
}

/* PacketProvider.GetNextPacket (Ogg/PacketProvider.cs:51-54,350-366) */
static void next_packet(vo_stream* s, vo_packet* out) {
  const vo_page* p = get_page(s, s->page_index);
  if (!p) {
    memset(out, 0, sizeof(*out));
    return;
  }
  create_packet(s, &s->page_index, &s->packet_index, 1, p->granule, p->is_resync, p->is_continued,
                p->packet_count, out);
}

/* StreamDecoder.GetPacketGranuleCount (StreamDecoder.cs:882-913) */
static int packet_granule_count(vo_stream* s, const vo_packet* pk) {
  if (pk->is_resync) return 0;
  vo_bits br;
  vo_bits_init(&br, pk->data, pk->len);
  if (vo_read_bit(&br)) return 0;
  uint32_t mode = (uint32_t)vo_read_bits(&br, s->setup.mode_bits);
  if (mode >= (uint32_t)s->setup.nmodes) return 0;
  vo_pinfo info;
  if (vo_mode_packet_info(&s->setup, &s->setup.modes[mode], &br, &info)) return info.right_start - info.left_start;
  return 0;
}

/* PacketProvider.CreateValidPacket (Ogg/PacketProvider.cs:413-425) */
static int create_valid_packet(vo_stream* s, int64_t* page_index, int* packet_index, int is_resync,
                               int is_continued, int packet_count, vo_packet* out) {
  create_packet(s, page_index, packet_index, 0, 0, *packet_index == 0 && is_resync, is_continued, packet_count,
                out);
  return out->valid ? VO_OK : VO_E_INVALID_DATA;
}

static int first_data_page_index(vo_stream* s) {
  /* StreamPageReader.FindFirstDataPage (Ogg/StreamPageReader.cs:191-208) */
  int64_t idx = s->pages_loaded - 1;
  if (idx < 0) idx = 0;
  while (s->first_data_page < 0) {
    if (!get_page(s, idx)) return -1;
    idx++;
  }
  return s->first_data_page;
}

/* PacketProvider.FillPageEndGranuleCache (Ogg/PacketProvider.cs:203-307) */
static int fill_page_end_cache(vo_stream* s, int64_t target) {
  int64_t p_index = s->page_end_n;
  int64_t first_data = first_data_page_index(s);
  if (first_data < 0) first_data = 0;
#define PUSH_END(v)                                                                        \
  do {                                                                                     \
    if (s->page_end_n == s->page_end_cap) {                                                \
      s->page_end_cap = s->page_end_cap ? s->page_end_cap * 2 : 64;                        \
      s->page_end_granules = (int64_t*)realloc(s->page_end_granules, 8 * (size_t)s->page_end_cap); \
    }                                                                                      \
    s->page_end_granules[s->page_end_n++] = (v);                                           \
  } while (0)
  while (p_index < first_data) {
    PUSH_END(0);
    p_index++;
  }
  while (p_index <= target) {
    if (s->has_all_pages && p_index >= page_count(s)) break;
    int64_t page_length = 0;
    int first_real = 0;
    int64_t prev = p_index - 1;
    if (prev >= 0) {
      const vo_page* pp = get_page(s, prev);
      if (!pp) return VO_E_INVALID_DATA;
      if (pp->is_continued) {
        int last_idx = pp->packet_count - 1;
        vo_packet pk;
        int64_t pi = prev;
        create_packet(s, &pi, &last_idx, 0, 0, 0, pp->is_continued, pp->packet_count, &pk);
        if (!pk.valid) {
          if (!s->has_all_pages) return VO_E_INVALID_DATA;
          break;
        }
        page_length += packet_granule_count(s, &pk);
        packet_free(&pk);
        first_real = 1;
      }
    }
    const vo_page* p = get_page(s, p_index);
    if (!p) {
      if (!s->has_all_pages) return VO_E_INVALID_DATA;
      break;
    }
    int packet_index = first_real;
    if (p_index == first_data) packet_index = 1;
    int p_count = p->packet_count;
    if (p->is_continued) p_count--;
    for (; packet_index < p_count; packet_index++) {
      vo_packet pk;
      int64_t pi = p_index;
      int rc = create_valid_packet(s, &pi, &packet_index, p->is_resync, p->is_continued, p->packet_count, &pk);
      if (rc != VO_OK) return rc;
      page_length += packet_granule_count(s, &pk);
      packet_free(&pk);
    }
    int64_t g = page_length;
    if (p_index > 0) g += s->page_end_granules[p_index - 1];
    PUSH_END(g);
    p_index++;
  }
#undef PUSH_END
  return VO_OK;
}

/* PacketProvider.GetPageRange (Ogg/PacketProvider.cs:171-201) */
static int get_page_range(vo_stream* s, int64_t page_index, int64_t* start, int64_t* end, int* err) {
  if ((uint64_t)page_index >= (uint64_t)s->page_end_n) {
    int rc = fill_page_end_cache(s, page_index);
    if (rc != VO_OK) {
      *err = rc;
      *start = *end = 0;
      return 0;
    }
    if ((uint64_t)page_index > (uint64_t)s->page_end_n) page_index = s->page_end_n;
  }
  if ((uint64_t)(page_index - 1) < (uint64_t)s->page_end_n)
    *start = s->page_end_granules[page_index - 1];
  else
    *start = 0;
  if ((uint64_t)page_index < (uint64_t)s->page_end_n) {
    *end = s->page_end_granules[page_index];
    return 1;
  }
  *end = *start;
  return 0;
}

/* PacketProvider.GetGranuleCount (Ogg/PacketProvider.cs:35-49) */
int64_t vo_total_samples(vo_stream* s) {
  int64_t start, end;
  int err = 0;
  get_page_range(s, INT64_MAX, &start, &end, &err);
  if (err) return err;
  if (s->has_all_pages && start > s->max_granule) start = s->max_granule;
  return start;
}

/* Test access to the cache FillPageEndGranuleCache (Ogg/PacketProvider.cs:203-307) builds: fills it to the end
 * of the stream (what GetGranuleCount does) and copies min(n, cap) entries; returns n or a negative error. */
int vo_page_end_granules(vo_stream* s, int64_t* out, int cap) {
  int64_t start, end;
  int err = 0;
  get_page_range(s, INT64_MAX, &start, &end, &err);
  if (err) return err;
  for (int i = 0; i < s->page_end_n && i < cap; i++) out[i] = s->page_end_granules[i];
  return s->page_end_n;
}

/* StreamPageReader.FindPage (Ogg/StreamPageReader.cs:152-305): the three search
 * strategies all land on the first page whose header granule exceeds the target
 * (index+1 on a direct hit); restated as a forward scan over header granules. */
static int64_t find_page(vo_stream* s, int64_t granule_pos, int* err) {
  if (granule_pos == 0) {
    int fd = first_data_page_index(s);
    if (fd < 0) *err = VO_E_SEEK_RANGE;
    return fd;
  }
  int last = s->pages_loaded - 1;
  if (last < 0) {
    if (!get_page(s, 0)) {
      *err = VO_E_SEEK_RANGE;
      return -1;
    }
    last = s->pages_loaded - 1;
  }
  int64_t last_gp = s->pages[last].granule;
  if (granule_pos < last_gp) {
    /* FindPageBisection between the first data page and `last` */
    int64_t low = first_data_page_index(s), high = last, high_gp = last_gp, low_gp = 0, dist;
    if (low < 0) {
      *err = VO_E_SEEK_RANGE;
      return -1;
    }
    while ((dist = high - low) > 0) {
      int64_t index = low + (int64_t)((double)dist * ((double)(granule_pos - low_gp) / (double)(high_gp - low_gp)));
      int64_t gp = s->pages[index].granule;
      if (gp > granule_pos) {
        high = index;
        high_gp = gp;
      } else if (gp < granule_pos) {
        low = index + 1;
        low_gp = gp + 1;
      } else {
        return index + 1;
      }
    }
    return low;
  } else if (granule_pos > last_gp) {
    /* FindPageForward */
    int64_t idx = last, gp = last_gp;
    while (gp <= granule_pos) {
      ++idx;
      const vo_page* p = get_page(s, idx);
      if (!p) {
        if (s->max_granule < granule_pos) {
          *err = VO_E_SEEK_RANGE;
          return -1;
        }
        break;
      }
      gp = p->granule;
    }
    return idx;
  }
  return last + 1;
}

/* PacketProvider.NormalizePacketIndex (Ogg/PacketProvider.cs:312-348) */
static int normalize_packet_index(vo_stream* s, int64_t* page_index, int* packet_index) {
  const vo_page* p = get_page(s, *page_index);
  if (!p) return 0;
  int is_resync = p->is_resync, is_continuation = (p->flags & 1) != 0;
  int64_t pg = *page_index;
  int pk = *packet_index;
  while (pk < (is_continuation ? 1 : 0)) {
    if (is_continuation && is_resync) return 0;
    int was_continuation = is_continuation;
    const vo_page* q = get_page(s, --pg);
    if (!q) return 0;
    is_resync = q->is_resync;
    is_continuation = (q->flags & 1) != 0;
    if (was_continuation && !q->is_continued) return 0;
    pk += q->packet_count - (was_continuation ? 1 : 0);
  }
  *page_index = pg;
  *packet_index = pk;
  return 1;
}

/* PacketProvider.SeekTo + GetTargetPageInfo (Ogg/PacketProvider.cs:56-169) */
static int64_t provider_seek(vo_stream* s, int64_t granule_pos, int pre_roll, int* err) {
  if (granule_pos < 0) {
    *err = VO_E_ARGUMENT;
    return 0;
  }
  int64_t page_index = find_page(s, granule_pos, err);
  if (*err) return 0;
  int64_t page_start = 0, page_end = 0;
  for (;;) {
    if (!get_page_range(s, page_index, &page_start, &page_end, err)) {
      return page_start; /* "We're at the last page": cursor untouched */
    }
    if (granule_pos >= page_start && granule_pos <= page_end) break;
    if (granule_pos - page_end > 0)
      page_index++;
    else
      page_index--;
  }
  const vo_page* p = get_page(s, page_index);
  if (!p) {
    *err = VO_E_INVALID_DATA;
    return 0;
  }
  int is_continuation = (p->flags & 1) != 0;
  int first_real = is_continuation ? 1 : 0;
  int64_t cur = page_end;
  int packet_index = p->packet_count - 1;
  if (p->is_continued) packet_index--;
  for (; packet_index >= first_real; packet_index--) {
    vo_packet pk;
    int64_t pi = page_index;
    int rc = create_valid_packet(s, &pi, &packet_index, packet_index == 0 && p->is_resync, p->is_continued,
                                 p->packet_count, &pk);
    if (rc != VO_OK) {
      *err = rc;
      return 0;
    }
    cur -= packet_granule_count(s, &pk);
    packet_free(&pk);
    if (granule_pos >= cur) break;
  }
  if (packet_index == 0 && first_real == 1) {
    int64_t prev = page_index - 1;
    const vo_page* pp = get_page(s, prev);
    if (!pp) {
      *err = VO_E_INVALID_DATA;
      return 0;
    }
    int last_idx = pp->packet_count - 1;
    vo_packet pk;
    int64_t pi = prev;
    int rc = create_valid_packet(s, &pi, &last_idx, packet_index == 0 && p->is_resync, p->is_continued,
                                 p->packet_count, &pk);
    if (rc != VO_OK) {
      *err = rc;
      return 0;
    }
    cur -= packet_granule_count(s, &pk);
    packet_free(&pk);
    page_index = prev;
    packet_index = last_idx;
  }
  if (page_index > first_data_page_index(s) || packet_index > 0) packet_index -= pre_roll;
  if (!normalize_packet_index(s, &page_index, &packet_index)) {
    *err = VO_E_SEEK_RANGE;
    return 0;
  }
  s->page_index = page_index;
  s->packet_index = (uint8_t)packet_index;
  return cur;
}

/* ---------------------------------------------------------- stream decoder -- */
static float* get_buffer(vo_stream* s) {
  if (s->pool_n > 0) return s->pool[--s->pool_n];
  return (float*)calloc((size_t)s->setup.size1 * s->setup.channels, sizeof(float));
}
static void return_buffer(vo_stream* s, float* b) {
  if (!b) return;
  if (s->pool_n < 2)
    s->pool[s->pool_n++] = b;
  else
    free(b);
}

/* StreamDecoder.ResetDecoder (StreamDecoder.cs:357-369) */
static void reset_decoder(vo_stream* s) {
  return_buffer(s, s->prev_buf);
  s->prev_buf = NULL;
  s->prev_start = s->prev_end = s->prev_stop = 0;
  return_buffer(s, s->next_buf);
  s->next_buf = NULL;
  s->eos_found = EOS_NONE;
  s->has_clipped = 0;
  s->has_position = 0;
}

/* StreamDecoder.DecodeNextPacket (StreamDecoder.cs:696-762).  Returns the buffer or NULL. */
static float* decode_next_packet(vo_stream* s, vo_pinfo* info, int* is_eos, int64_t* sample_position, int* err) {
  vo_packet pk;
  next_packet(s, &pk);
  memset(info, 0, sizeof(*info));
  *sample_position = -1;
  if (!pk.valid) {
    *is_eos = EOS_INVALID_PACKET;
    return NULL;
  }
  float* result = NULL;
  *is_eos = pk.is_eos ? EOS_PACKET_FLAG : EOS_NONE;
  if (pk.is_resync) s->has_position = 0;
  vo_bits br;
  vo_bits_init(&br, pk.data, pk.len);
  if (vo_read_bits(&br, 1) == 0) {
    int mode_idx = (int)vo_read_bits(&br, s->setup.mode_bits);
    if ((uint32_t)mode_idx >= (uint32_t)s->setup.nmodes) {
      *err = VO_E_INVALID_DATA; /* "Unused mode index." */
      packet_free(&pk);
      return NULL;
    }
    const vo_mode* mode = &s->setup.modes[mode_idx];
    if (!s->next_buf) s->next_buf = get_buffer(s);
    /* Mode.Decode (Mode.cs:68-85) */
    if (vo_mode_packet_info(&s->setup, mode, &br, info)) {
      int block_size = mode->block_flag ? s->setup.size1 : s->setup.size0;
      vo_mapping_decode(&s->setup, &s->setup.mappings[mode->mapping], &br, block_size, s->next_buf, NULL);
      *sample_position = pk.granule;
      result = s->next_buf;
    }
  }
  packet_free(&pk);
  return result;
}

/* StreamDecoder.OverlapBuffers (StreamDecoder.cs:764-791) */
static int overlap_buffers(vo_stream* s, const vo_pinfo* info, float* prev, float* next, int packet_len) {
  const float* slope = s->setup.slope[info->left_use_size1 ? 1 : 0];
  int slope_len = (info->left_use_size1 ? s->setup.size1 : s->setup.size0) / 2;
  int size1 = s->setup.size1;
  if (packet_len > slope_len || packet_len < 0 || info->left_start + packet_len > size1)
    return VO_E_REF_FAULT; /* AsSpan(0, packetLen) throws in the reference */
  for (int ch = 0; ch < s->setup.channels; ch++) {
    const float* pv = prev + (size_t)size1 * ch + s->prev_end;
    float* chan = next + info->left_start + (size_t)size1 * ch;
    for (int i = 0; i < packet_len; i++) {
      float a = chan[i] * slope[i];
      float b = pv[i] * slope[packet_len - 1 - i]; /* slope.AsSpan(0, packetLen)[^(i + 1)] */
      chan[i] = a + b;
    }
  }
  return VO_OK;
}

/* StreamDecoder.ReadNextPacket (StreamDecoder.cs:640-694) */
static int read_next_packet(vo_stream* s, int64_t* sample_position, int* err) {
  vo_pinfo info;
  int is_eos = 0;
  float* cur = decode_next_packet(s, &info, &is_eos, sample_position, err);
  if (*err) return 0;
  s->eos_found |= is_eos;
  if (!cur) return 0;
  int packet_len = s->prev_stop - s->prev_end;
  int right_start = info.right_start;
  if (*sample_position != -1 && is_eos != EOS_NONE) {
    int64_t actual_end = s->current_position + packet_len;
    int diff = (int)(actual_end - *sample_position);
    if (diff > 0) {
      right_start = right_start - diff;
      if (right_start < 0) right_start = 0;
    }
  }
  if (s->prev_buf) {
    int rc = overlap_buffers(s, &info, s->prev_buf, cur, packet_len);
    if (rc != VO_OK) {
      *err = rc;
      return 0;
    }
    s->prev_start = info.left_start;
  } else {
    s->prev_start = right_start;
  }
  s->prev_end = right_start;
  s->prev_stop = info.right_end;
  s->next_buf = s->prev_buf;
  s->prev_buf = cur;
  return 1;
}

static inline float clip_value(float v, int* clipped) {
  /* Utils.ClipValue (Utils.cs:44-58) */
  if (v > 0.99999994f) {
    *clipped = 1;
    return 0.99999994f;
  }
  if (v < -0.99999994f) {
    *clipped = 1;
    return -0.99999994f;
  }
  return v;
}

/* StreamDecoder.Read (StreamDecoder.cs:418-498) */
static int stream_read(vo_stream* s, float* buffer, int nfloats, int samples_to_read, int channel_stride,
                       int interleave) {
  int channels = s->setup.channels;
  if (s->fault) return s->fault;
  if (nfloats % channels != 0) return VO_E_ARGUMENT;
  if (nfloats < samples_to_read * channels) return VO_E_ARGUMENT;
  int idx = 0;
  int size1 = s->setup.size1;
  while (idx == 0) {
    if (s->prev_start == s->prev_end) {
      if (s->eos_found != EOS_NONE) {
        return_buffer(s, s->prev_buf);
        s->prev_buf = NULL;
        break;
      }
      int64_t sample_position = -1;
      int err = 0;
      if (!read_next_packet(s, &sample_position, &err)) {
        if (err) {
          if (err == VO_E_REF_FAULT) {
            /* the reference throws here; the restatement ends the stream instead */
            s->fault = err;
            s->eos_found |= EOS_INVALID_PACKET;
            return_buffer(s, s->prev_buf);
            s->prev_buf = NULL;
            s->prev_start = s->prev_end = s->prev_stop = 0;
          }
          return err;
        }
        if (s->eos_found & EOS_PACKET_FLAG) s->prev_end = s->prev_stop;
      }
      if (sample_position != -1 && !s->has_position) {
        s->has_position = 1;
        s->current_position = sample_position - (s->prev_end - s->prev_start) - idx;
      }
    }
    int copy_len = samples_to_read - idx;
    if (s->prev_end - s->prev_start < copy_len) copy_len = s->prev_end - s->prev_start;
    if (copy_len <= 0) {
      if (samples_to_read - idx <= 0) break; /* reference would spin forever on a zero-length request */
      if (s->prev_end < s->prev_start) return VO_E_REF_FAULT; /* reference: Debug.Assert / endless loop */
      continue;
    }
    int clipped = 0;
    for (int ch = 0; ch < channels; ch++) {
      const float* src = s->prev_buf + s->prev_start + (size_t)size1 * ch;
      for (int i = 0; i < copy_len; i++) {
        float p = s->clip ? clip_value(src[i], &clipped) : src[i];
        if (interleave)
          buffer[(size_t)(idx + i) * channels + ch] = p;
        else
          buffer[(size_t)ch * channel_stride + idx + i] = p;
      }
    }
    s->has_clipped |= clipped;
    idx += copy_len;
    s->prev_start += copy_len;
    s->current_position += copy_len;
  }
  return idx;
}

int vo_read(vo_stream* s, float* buf, int nfloats) {
  return stream_read(s, buf, nfloats, nfloats / s->setup.channels, 0, 1);
}
int vo_read_planar(vo_stream* s, float* buf, int nfloats, int samples_to_read, int channel_stride) {
  return stream_read(s, buf, nfloats, samples_to_read, channel_stride, 0);
}

/* StreamDecoder.SeekTo (StreamDecoder.cs:817-880), SeekOrigin.Begin */
int vo_seek(vo_stream* s, int64_t sample_position) {
  if (sample_position < 0) return VO_E_ARGUMENT;
  int err = 0;
  int64_t pos = provider_seek(s, sample_position, 1, &err);
  if (err) return err;
  int roll_forward = (int)(sample_position - pos);
  reset_decoder(s);
  s->fault = 0;
  s->has_position = 1;
  int64_t sp;
  if (!read_next_packet(s, &sp, &err)) {
    if (err) return err;
    s->eos_found |= EOS_INVALID_PREROLL;
    int64_t max_granule = vo_total_samples(s);
    if (sample_position > max_granule) return VO_E_SEEK_RANGE;
    s->prev_start = s->prev_stop;
    s->current_position = sample_position;
    return VO_OK;
  }
  if (!read_next_packet(s, &sp, &err)) {
    if (err == VO_E_REF_FAULT) s->fault = err;
    if (err) return err;
    reset_decoder(s);
    s->eos_found |= EOS_INVALID_PACKET;
    return VO_E_PREROLL;
  }
  s->prev_start += roll_forward;
  s->current_position = sample_position;
  return VO_OK;
}

/* ------------------------------------------------------------ open / close -- */
static int parse_comments(vo_stream* s, const vo_packet* pk) {
  static const uint8_t sig[7] = {0x03, 0x76, 0x6f, 0x72, 0x62, 0x69, 0x73};
  vo_bits br;
  vo_bits_init(&br, pk->data, pk->len);
  for (int i = 0; i < 7; i++)
    if (vo_read_bits(&br, 8) != sig[i]) return VO_E_INVALID_DATA;
  uint32_t vlen = (uint32_t)vo_read_bits(&br, 32);
  if ((int64_t)vlen * 8 > br.total_bits - br.pos) return VO_E_INVALID_DATA;
  s->vendor = (char*)malloc((size_t)vlen + 1);
  for (uint32_t i = 0; i < vlen; i++) s->vendor[i] = (char)vo_read_bits(&br, 8);
  s->vendor[vlen] = 0;
  s->vendor_len = (int)vlen;
  uint32_t n = (uint32_t)vo_read_bits(&br, 32);
  if ((int64_t)n * 32 > br.total_bits - br.pos) return VO_E_INVALID_DATA;
  s->comments = (char**)calloc(n ? n : 1, sizeof(char*));
  s->comment_lens = (int*)calloc(n ? n : 1, sizeof(int));
  s->ncomments = (int)n;
  for (uint32_t c = 0; c < n; c++) {
    uint32_t l = (uint32_t)vo_read_bits(&br, 32);
    if ((int64_t)l * 8 > br.total_bits - br.pos) return VO_E_INVALID_DATA;
    s->comments[c] = (char*)malloc((size_t)l + 1);
    for (uint32_t i = 0; i < l; i++) s->comments[c][i] = (char)vo_read_bits(&br, 8);
    s->comments[c][l] = 0;
    s->comment_lens[c] = (int)l;
  }
  return VO_OK;
}

vo_stream* vo_open(const uint8_t* data, size_t len, int* err) {
  int e = VO_OK;
  vo_stream* s = (vo_stream*)calloc(1, sizeof(vo_stream));
  if (!s) {
    if (err) *err = VO_E_NOMEM;
    return NULL;
  }
  s->file = data;
  s->file_len = len;
  s->first_data_page = -1;
  s->clip = 1;
  e = scan_file(s);
  if (e == VO_OK) {
    /* StreamDecoder.Initialize -> ProcessHeaderPackets (StreamDecoder.cs:71-165) */
    for (int i = 0; i < 3 && e == VO_OK; i++) {
      next_packet(s, &s->hdr[i]);
      if (!s->hdr[i].valid) e = VO_E_INVALID_DATA;
    }
  }
  if (e == VO_OK) e = vo_setup_parse_id(&s->setup, s->hdr[0].data, s->hdr[0].len);
  if (e == VO_OK) e = parse_comments(s, &s->hdr[1]);
  if (e == VO_OK) e = vo_setup_parse_books(&s->setup, s->hdr[2].data, s->hdr[2].len);
  if (e == VO_OK) {
    s->current_position = 0;
    reset_decoder(s);
    s->has_position = 1;
  }
  if (err) *err = e;
  if (e != VO_OK) {
    vo_close(s);
    return NULL;
  }
  return s;
}

void vo_close(vo_stream* s) {
  if (!s) return;
  for (int i = 0; i < 3; i++) free(s->hdr[i].data);
  for (int i = 0; i < s->table_n; i++) free(s->table[i].data);
  free(s->table);
  free(s->vendor);
  for (int i = 0; i < s->ncomments; i++) free(s->comments[i]);
  free(s->comments);
  free(s->comment_lens);
  free(s->pages);
  free(s->page_end_granules);
  free(s->prev_buf);
  free(s->next_buf);
  for (int i = 0; i < s->pool_n; i++) free(s->pool[i]);
  vo_setup_free(&s->setup);
  free(s);
}

/* --------------------------------------------------------------- accessors -- */
int vo_channels(const vo_stream* s) { return s->setup.channels; }
int vo_sample_rate(const vo_stream* s) { return s->setup.sample_rate; }
int vo_block_size(const vo_stream* s, int which) { return which ? s->setup.size1 : s->setup.size0; }
int vo_bitrate(const vo_stream* s, int which) {
  return which == 0 ? s->setup.br_upper : which == 1 ? s->setup.br_nominal : s->setup.br_lower;
}
int64_t vo_container_bits(const vo_stream* s) { return s->container_bits; }
int64_t vo_waste_bits(const vo_stream* s) { return s->waste_bits; }
int vo_page_count(const vo_stream* s) { return s->npages; }
int vo_crc_failures(const vo_stream* s) { return s->crc_failures; }
const char* vo_vendor(const vo_stream* s, int* len) {
  if (len) *len = s->vendor_len;
  return s->vendor;
}
int vo_comment_count(const vo_stream* s) { return s->ncomments; }
const char* vo_comment(const vo_stream* s, int i, int* len) {
  if (i < 0 || i >= s->ncomments) return NULL;
  if (len) *len = s->comment_lens[i];
  return s->comments[i];
}
const uint8_t* vo_header_packet(const vo_stream* s, int which, int* len) {
  if (which < 0 || which > 2) return NULL;
  if (len) *len = s->hdr[which].len;
  return s->hdr[which].data;
}
void vo_set_clip(vo_stream* s, int clip) { s->clip = clip != 0; }
int vo_has_clipped(const vo_stream* s) { return s->has_clipped; }
int vo_is_end_of_stream(const vo_stream* s) { return s->eos_found != EOS_NONE && s->prev_buf == NULL; }
int64_t vo_sample_position(const vo_stream* s) { return s->current_position; }

/* Forward walk from the first audio packet on a private second reader (so the lazy
 * page state of `s` is untouched); all pages are loaded first, so IsEndOfStream matches
 * a reader that has seen the whole file. */
static int build_table(vo_stream* s) {
  if (s->table_built) return VO_OK;
  s->table_built = 1;
  int e = 0;
  vo_stream* t = vo_open(s->file, s->file_len, &e);
  if (!t) return e;
  load_pages_upto(t, INT32_MAX);
  t->page_index = 0;
  t->packet_index = 0;
  int cap = 0;
  vo_packet pk;
  for (int i = 0;; i++) {
    next_packet(t, &pk);
    if (!pk.valid) break;
    if (i < 3) {
      packet_free(&pk);
      continue;
    }
    if (s->table_n == cap) {
      cap = cap ? cap * 2 : 256;
      s->table = (vo_packet*)realloc(s->table, sizeof(vo_packet) * (size_t)cap);
    }
    s->table[s->table_n++] = pk;
  }
  vo_close(t);
  return VO_OK;
}

int vo_audio_packet_count(vo_stream* s) {
  build_table(s);
  return s->table_n;
}
int vo_audio_packet(vo_stream* s, int i, vo_packet_view* out) {
  build_table(s);
  if (i < 0 || i >= s->table_n) return VO_E_ARGUMENT;
  const vo_packet* p = &s->table[i];
  out->data = p->data;
  out->len = p->len;
  out->is_resync = p->is_resync;
  out->is_eos = p->is_eos;
  out->granule = p->granule;
  out->page_index = p->page_index;
  out->packet_index = p->packet_index;
  return VO_OK;
}

/* ----------------------------------------------------------- book accessors -- */
int vo_book_count(const vo_stream* s) { return s->setup.nbooks; }
int vo_book_info(const vo_stream* s, int b, int* dims, int* entries, int* max_bits, int* map_type,
                 int* prefix_bits, int* overflow_count) {
  if (b < 0 || b >= s->setup.nbooks) return VO_E_ARGUMENT;
  const vo_book* bk = &s->setup.books[b];
  if (dims) *dims = bk->dims;
  if (entries) *entries = bk->entries;
  if (max_bits) *max_bits = bk->max_bits;
  if (map_type) *map_type = bk->map_type;
  if (prefix_bits) *prefix_bits = bk->prefix_bits;
  if (overflow_count) *overflow_count = bk->overflow_n;
  return VO_OK;
}
int vo_book_lengths(const vo_stream* s, int b, int* lengths) {
  if (b < 0 || b >= s->setup.nbooks) return VO_E_ARGUMENT;
  memcpy(lengths, s->setup.books[b].lengths, sizeof(int) * (size_t)s->setup.books[b].entries);
  return VO_OK;
}
const float* vo_book_lookup(const vo_stream* s, int b, int* count) {
  if (b < 0 || b >= s->setup.nbooks) return NULL;
  const vo_book* bk = &s->setup.books[b];
  if (count) *count = bk->lookup ? bk->entries * bk->dims : 0;
  return bk->lookup;
}
int vo_book_decode(const vo_stream* s, int b, const uint8_t* data, int len_bytes, int64_t* bitpos,
                   int* is_short) {
  if (b < 0 || b >= s->setup.nbooks) return VO_E_ARGUMENT;
  /* copy into a padded buffer so peeks never run off the caller's array */
  uint8_t* tmp = (uint8_t*)calloc((size_t)len_bytes + 16, 1);
  memcpy(tmp, data, (size_t)len_bytes);
  vo_bits br;
  vo_bits_init(&br, tmp, len_bytes);
  br.pos = *bitpos;
  int v = vo_book_decode_scalar(&s->setup.books[b], &br);
  *bitpos = br.pos;
  if (is_short) *is_short = br.is_short;
  free(tmp);
  return v;
}
double vo_book_kraft(const vo_stream* s, int b) {
  const vo_book* bk = &s->setup.books[b];
  double k = 0;
  for (int i = 0; i < bk->entries; i++)
    if (bk->lengths[i] > 0) k += ldexp(1.0, -bk->lengths[i]);
  return k;
}

/* ------------------------------------------------------ single packet dump -- */
int vo_decode_packet_dump(vo_stream* s, const uint8_t* data, int len, vo_packet_dump* d) {
  uint8_t* tmp = (uint8_t*)calloc((size_t)len + 16, 1);
  memcpy(tmp, data, (size_t)len);
  vo_bits br;
  vo_bits_init(&br, tmp, len);
  d->status = 1;
  d->scalars_n = 0;
  d->classes_n = 0;
  memset(d->post_count, 0, sizeof(d->post_count));
  memset(d->raw_posts, 0, sizeof(d->raw_posts));
  memset(d->final_y, 0, sizeof(d->final_y));
  memset(d->step_flags, 0, sizeof(d->step_flags));
  memset(d->no_execute, 0, sizeof(d->no_execute));
  int rc = VO_OK;
  if (vo_read_bits(&br, 1) == 0) {
    int mode_idx = (int)vo_read_bits(&br, s->setup.mode_bits);
    if ((uint32_t)mode_idx >= (uint32_t)s->setup.nmodes) {
      rc = VO_E_INVALID_DATA;
    } else {
      const vo_mode* mode = &s->setup.modes[mode_idx];
      vo_pinfo info;
      d->mode = mode_idx;
      d->block_flag = mode->block_flag;
      d->block_size = mode->block_flag ? s->setup.size1 : s->setup.size0;
      if (vo_mode_packet_info(&s->setup, mode, &br, &info)) {
        d->info[0] = info.length;
        d->info[1] = info.left_use_size1;
        d->info[2] = info.left_start;
        d->info[3] = info.left_end;
        d->info[4] = info.right_start;
        d->info[5] = info.right_end;
        float* buf = (float*)calloc((size_t)s->setup.size1 * s->setup.channels, sizeof(float));
        vo_mapping_decode(&s->setup, &s->setup.mappings[mode->mapping], &br, d->block_size, buf, d);
        free(buf);
        d->status = 0;
      }
    }
  }
  d->bits_read = (int)br.pos;
  d->is_short = br.is_short;
  free(tmp);
  return rc;
}
