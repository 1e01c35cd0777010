/* vo_decode.c -- TEST INFRASTRUCTURE ONLY (see vorbis_oracle.h).
 * Restates the per-packet decode of the reference:
 *   Floor1.Unpack / Apply / UnwrapPosts / RenderPoint / RenderLineMulti (Floor1.cs:162-397)
 *   Residue0.Decode / WriteVectors, Residue1.WriteVectors, Residue2.Decode
 *       (Residue0.cs:117-231, Residue1.cs:12-34, Residue2.cs:12-52)
 *   Mapping.DecodePacket / ApplyCoupling (Mapping.cs:98-269)
 *   Mode.GetPacketInfo (Mode.cs:30-66)
 *   Mdct (Mdct.cs:29-726; the stb_vorbis inverse MDCT, same operation order)
 * Compile with -ffp-contract=off: the CLR JIT never fuses a*b+c.
 */
#include "vo_internal.h"

/* ------------------------------------------------------------- dump hooks -- */
static inline void dump_scalar(vo_packet_dump* d, int v) {
  if (!d || !d->scalars) return;
  if (d->scalars_n < d->scalars_cap) d->scalars[d->scalars_n] = v;
  d->scalars_n++;
}
static inline void dump_class(vo_packet_dump* d, int v) {
  if (!d || !d->classes) return;
  if (d->classes_n < d->classes_cap) d->classes[d->classes_n] = v;
  d->classes_n++;
}
static inline int decode_scalar(const vo_book* bk, vo_bits* br, vo_packet_dump* d) {
  int v = vo_book_decode_scalar(bk, br);
  dump_scalar(d, v);
  return v;
}

/* ------------------------------------------------------------------ floor1 -- */
typedef struct {
  int posts[64];
  int post_count;
  /* floor 0 (Floor0.Data, Floor0.cs:11-28) */
  float amp;
  float coeff[256];
  uint32_t amp_raw;
  int book_num;
} floor1_data;

/* ------------------------------------------------------------------ floor0 -- */
/* Floor0.Unpack (Floor0.cs:115-167).  The amplitude is read, then the book number and the
 * coefficients are read WHATEVER the amplitude is (libvorbis stops at amplitude 0; the reference
 * does not), so a silent channel still consumes its bits. */
static void floor0_unpack(const vo_floor1* f, const vo_book* books, vo_bits* br, floor1_data* fd,
                          vo_packet_dump* d) {
  memset(fd->coeff, 0, sizeof(fd->coeff));
  memset(fd->posts, 0, sizeof(fd->posts));
  fd->post_count = 0;
  fd->amp_raw = 0;
  fd->book_num = 0;
  const uint64_t amp = vo_read_bits(br, f->amp_bits);
  /* double ampDiv = (1 << _ampBits) - 1 : int arithmetic, shift count taken mod 32, wrapping */
  const int32_t one_shift = (int32_t)(1u << (f->amp_bits & 31));
  const double amp_div = (double)(int32_t)((uint32_t)one_shift - 1u);
  /* (float)(amp * _ampOfs / ampDiv): ulong product (wraps), then double division */
  fd->amp = (float)((double)(amp * (uint64_t)f->amp_ofs) / amp_div);
  fd->amp_raw = (uint32_t)amp;
  const uint32_t book_num = (uint32_t)vo_read_bits(br, vo_ilog(f->nbooks0));
  fd->book_num = (int)book_num;
  if (book_num >= (uint32_t)f->nbooks0) {
    fd->amp = 0;
    return;
  }
  const vo_book* bk = &books[f->books0[book_num]];
  for (int i = 0; i < f->order;) {
    int entry = decode_scalar(bk, br, d);
    if (entry == -1) {
      fd->amp = 0;
      return;
    }
    const float* lookup = bk->lookup + (size_t)entry * bk->dims;
    for (int j = 0; i < f->order && j < bk->dims; j++, i++) fd->coeff[i] = lookup[j];
  }
  /* the "averaging" */
  const int dim = bk->dims;
  float last = 0.f;
  for (int j = 0; j < f->order;) {
    for (int k = 0; j < f->order && k < dim; j++, k++) fd->coeff[j] += last;
    last = fd->coeff[j - 1];
  }
  /* the dump keeps the floor-1 fields: post_count = 2 when the channel has energy (Amp != 0,
   * Floor0.cs:21), posts = {raw amplitude, book number} */
  if (fd->amp != 0) {
    fd->post_count = 2;
    fd->posts[0] = (int)(fd->amp_raw & 0x7fffffffu);
    fd->posts[1] = fd->book_num;
  }
}

/* Floor0.Apply (Floor0.cs:169-224) */
static void floor0_apply(const vo_floor1* f, floor1_data* fd, int w, int block_size, float* residue) {
  const int n = block_size / 2;
  if (fd->amp <= 0.f) {
    memset(residue, 0, sizeof(float) * (size_t)n);
    return;
  }
  const int* bark = f->bark_map[w];
  const float* wmap = f->wmap[w];
  const int order = f->order;
  for (int j = 0; j < order; j++) fd->coeff[j] = 2.0f * cosf(fd->coeff[j]);
  const float amp_ofs = (float)f->amp_ofs;
  int i = 0;
  while (i < n) {
    int j;
    const int k = bark[i];
    float p = .5f, q = .5f;
    const float wv = wmap[k];
    for (j = 1; j < order; j += 2) {
      q *= wv - fd->coeff[j - 1];
      p *= wv - fd->coeff[j];
    }
    if (j == order) {
      /* odd order filter; slightly asymmetric */
      q *= wv - fd->coeff[j - 1];
      p *= p * (4.f - wv * wv);
      q *= q;
    } else {
      /* even order filter; still symmetric */
      p *= p * (2.f - wv);
      q *= q * (2.f + wv);
    }
    q = fd->amp / sqrtf(p + q) - amp_ofs;
    q = expf(q * 0.11512925f);
    residue[i] *= q;
    while (bark[++i] == k) residue[i] *= q;
  }
}

/* Floor1.Unpack (Floor1.cs:162-219) */
static void floor1_unpack(const vo_floor1* f, const vo_book* books, vo_bits* br, floor1_data* fd,
                          vo_packet_dump* d) {
  memset(fd->posts, 0, sizeof(fd->posts)); /* Data.Reset, Mapping.cs:111 */
  fd->post_count = 0;
  if (!vo_read_bit(br)) return;
  int post_count = 2;
  fd->posts[0] = (int)vo_read_bits(br, f->ybits);
  fd->posts[1] = (int)vo_read_bits(br, f->ybits);
  for (int i = 0; i < f->partitions; i++) {
    int cls = f->part_class[i];
    int cdim = f->class_dim[cls];
    int cbits = f->class_sub[cls];
    int csub = (1 << cbits) - 1;
    uint32_t cval = 0;
    if (cbits > 0) {
      cval = (uint32_t)decode_scalar(&books[f->class_master[cls]], br, d);
      if (cval == 0xFFFFFFFFu) {
        post_count = 0;
        break;
      }
    }
    int bail = 0;
    for (int j = 0; j < cdim; j++) {
      int book_idx = f->sub_books[cls][cval & (uint32_t)csub];
      cval >>= cbits;
      int post = 0;
      if (book_idx >= 0) {
        post = decode_scalar(&books[book_idx], br, d);
        if (post == -1) {
          post_count = 0;
          bail = 1;
          break;
        }
      }
      fd->posts[post_count++] = post;
    }
    if (bail) break;
  }
  fd->post_count = post_count;
}

/* Floor1.RenderPoint (Floor1.cs:355-370) */
static int render_point(int x0, int y0, int x1, int y1, int X) {
  int dy = y1 - y0;
  int adx = x1 - x0;
  int ady = dy < 0 ? -dy : dy;
  int off = ady * (X - x0) / adx;
  return dy < 0 ? y0 - off : y0 + off;
}

/* Floor1.UnwrapPosts (Floor1.cs:270-353): posts -> final Y in place, step flags out */
static void floor1_unwrap(const vo_floor1* f, floor1_data* fd, uint8_t* step) {
  int final_y[64];
  memset(final_y, 0, sizeof(final_y));
  step[0] = 1;
  step[1] = 1;
  final_y[0] = fd->posts[0];
  final_y[1] = fd->posts[1];
  for (int i = 2; i < fd->post_count; i++) {
    int lo = f->lneigh[i], hi = f->hneigh[i];
    int predicted = render_point(f->xlist[lo], final_y[lo], f->xlist[hi], final_y[hi], f->xlist[i]);
    int val = fd->posts[i];
    int highroom = f->range - predicted;
    int lowroom = predicted;
    int room = (highroom < lowroom ? highroom : lowroom) * 2;
    int result;
    if (val != 0) {
      step[lo] = 1;
      step[hi] = 1;
      step[i] = 1;
      if (val >= room) {
        result = highroom > lowroom ? val - lowroom + predicted : predicted - val + highroom - 1;
      } else {
        result = (val % 2) == 1 ? predicted - ((val + 1) / 2) : predicted + (val / 2);
      }
    } else {
      step[i] = 0;
      result = predicted;
    }
    final_y[i] = result;
  }
  memcpy(fd->posts, final_y, sizeof(final_y));
}

/* inverse_dB_table (Floor1.cs:407-473) as fp32 bit patterns */
static const uint32_t k_inverse_db_bits[256] = {
#include "vo_db_table.inc"
};

/* Floor1.RenderLineMulti (Floor1.cs:372-397).  The table index is unchecked in
 * the reference (quirk Q2); the oracle clamps only to stay memory-safe. */
static inline float db_at(int y) {
  float f;
  memcpy(&f, &k_inverse_db_bits[y < 0 ? 0 : (y > 255 ? 255 : y)], 4);
  return f;
}

static void render_line_multi(int x0, int y0, int x1, int y1, float* v) {
  int dy = y1 - y0;
  int adx = x1 - x0;
  int ady = dy < 0 ? -dy : dy;
  int sy = dy < 0 ? -1 : 1;
  int b = dy / adx;
  int x = x0, y = y0;
  int err = -adx;
  v[x] *= db_at(y);
  ady -= (b < 0 ? -b : b) * adx;
  while (++x < x1) {
    y += b;
    err += ady;
    if (err >= 0) {
      err -= adx;
      y += sy;
    }
    v[x] *= db_at(y);
  }
}

/* Floor1.Apply (Floor1.cs:222-268); quirk Q1: x1 = min(hx, n) before the slope */
static void floor1_apply(const vo_floor1* f, floor1_data* fd, int block_size, float* residue,
                         uint8_t* step_out) {
  int n = block_size / 2;
  if (fd->post_count <= 0) return;
  uint8_t step[64];
  memset(step, 0, sizeof(step));
  floor1_unwrap(f, fd, step);
  if (step_out) memcpy(step_out, step, 64);
  int lx = 0;
  int ly = fd->posts[0] * f->multiplier;
  for (int i = 1; i < fd->post_count; i++) {
    int idx = f->sortidx[i];
    if (step[idx]) {
      int hx = f->xlist[idx];
      int hy = fd->posts[idx] * f->multiplier;
      if (lx < n) render_line_multi(lx, ly, hx < n ? hx : n, hy, residue);
      lx = hx;
      ly = hy;
    }
    if (lx >= n) break;
  }
  if (lx < n) render_line_multi(lx, ly, n, ly, residue);
}

/* ----------------------------------------------------------------- residue -- */

/* Residue0.WriteVectors (Residue0.cs:208-231), quirk Q6: all dims summed into one bin */
static int write_vectors0(const vo_book* bk, vo_bits* br, float* ch, int offset, int part_size,
                          vo_packet_dump* d) {
  int steps = part_size / bk->dims;
  for (int step = 0; step < steps; step++) {
    int entry = decode_scalar(bk, br, d);
    if (entry == -1) return 1;
    float r = 0;
    const float* lk = bk->lookup + (size_t)entry * bk->dims;
    for (int k = 0; k < bk->dims; k++) r += lk[k];
    ch[offset + step] += r;
  }
  return 0;
}

/* Residue1.WriteVectors (Residue1.cs:12-34) */
static int write_vectors1(const vo_book* bk, vo_bits* br, float* ch, int offset, int part_size,
                          vo_packet_dump* d) {
  for (int i = 0; i < part_size;) {
    int entry = decode_scalar(bk, br, d);
    if (entry == -1) return 1;
    const float* lk = bk->lookup + (size_t)entry * bk->dims;
    float* res = ch + offset + i;
    for (int j = 0; j < bk->dims; j++) res[j] += lk[j];
    i += bk->dims;
  }
  return 0;
}

/* Residue0.Decode (Residue0.cs:117-206).  `chbuf` holds nch channels at `stride`. */
static void residue_decode(vo_residue* r, const vo_book* books, vo_bits* br, const uint8_t* no_decode,
                           int nch, int block_size, float* chbuf, int stride, vo_packet_dump* d) {
  int half = block_size / 2;
  int begin = r->begin < half ? r->begin : half;
  int end = r->end < half ? r->end : half;
  int n = end - begin;
  if (n <= 0) return;
  int part_count = n / r->part_size;
  const vo_book* cb = &books[r->class_book];
  int dim = cb->dims;
  int part_words = (part_count + dim - 1) / dim;
  int cache_len = nch * part_words;
  if (r->part_word_cache_len < cache_len) {
    /* Array.Resize keeps old content; new tail is zero (Residue0.cs:140-141) */
    int* nc = (int*)calloc((size_t)cache_len, sizeof(int));
    if (r->part_word_cache) memcpy(nc, r->part_word_cache, sizeof(int) * (size_t)r->part_word_cache_len);
    free(r->part_word_cache);
    r->part_word_cache = nc;
    r->part_word_cache_len = cache_len;
  }
  int* cache = r->part_word_cache;
  int max_stages = r->max_stages;

  for (int stage = 0; stage < max_stages; stage++) {
    for (int part = 0, entry = 0; part < part_count; entry++) {
      if (stage == 0) {
        for (int ch = 0; ch < nch; ch++) {
          if (no_decode[ch]) continue;
          int idx = decode_scalar(cb, br, d);
          /* quirk Q8: the reference accepts idx < decodeMap.Length (= partvals * dim, Residue0.cs:158) and
           * then indexes decodeMap[idx * dim + k] (:175-176), which for idx >= partvals is an
           * IndexOutOfRangeException out of Read.  There is no output to agree with; the oracle and the
           * product both end the packet's residue decode at such a classword, like a failed one. */
          if (idx >= 0 && idx < r->decode_map_len && idx < r->decode_map_len / (dim > 0 ? dim : 1)) {
            cache[ch * part_words + entry] = idx;
          } else {
            part = part_count;
            stage = max_stages;
            break;
          }
        }
      }
      for (int k = 0; part < part_count && k < dim; k++, part++) {
        int offset = begin + part * r->part_size;
        for (int ch = 0; ch < nch; ch++) {
          if (no_decode[ch]) continue;
          int map_index = cache[ch * part_words + entry] * dim;
          int cls = r->decode_map[map_index + k];
          if (stage == 0) dump_class(d, cls);
          if ((r->cascade[cls] & (1 << stage)) == 0) continue;
          if (!r->has_books[cls]) continue;
          const vo_book* bk = &books[r->books[cls][stage]];
          int bad = r->type == 0 ? write_vectors0(bk, br, chbuf + (size_t)ch * stride, offset, r->part_size, d)
                                 : write_vectors1(bk, br, chbuf + (size_t)ch * stride, offset, r->part_size, d);
          if (bad) {
            part = part_count;
            stage = max_stages;
            break;
          }
        }
      }
    }
  }
}

/* Residue2.Decode (Residue2.cs:12-52) */
static void residue2_decode(vo_residue* r, const vo_book* books, vo_bits* br, const uint8_t* no_decode,
                            int nch, int block_size, float* chbuf, int stride, vo_packet_dump* d) {
  int half = block_size / 2;
  int any = 0;
  for (int c = 0; c < nch; c++)
    if (!no_decode[c]) any = 1;
  if (!any) {
    for (int c = 0; c < nch; c++) memset(chbuf + (size_t)c * stride, 0, sizeof(float) * (size_t)half);
    return;
  }
  float* tmp = (float*)calloc((size_t)half * nch, sizeof(float));
  uint8_t one_false = 0;
  residue_decode(r, books, br, &one_false, 1, block_size * nch, tmp, half * nch, d);
  if (nch == 1) {
    memcpy(chbuf, tmp, sizeof(float) * (size_t)half);
  } else {
    for (int c = 0; c < nch; c++) {
      float* dst = chbuf + (size_t)c * stride;
      for (int i = 0; i < half; i++) dst[i] = tmp[i * nch + c];
    }
  }
  free(tmp);
}

/* ---------------------------------------------------------------- coupling -- */
/* Mapping.ApplyCoupling scalar form (Mapping.cs:235-267) */
static void apply_coupling(float* mag, float* ang, int n) {
  for (int j = 0; j < n; j++) {
    float m = mag[j], a = ang[j];
    float nm = m, na = m;
    if (m > 0) {
      if (a > 0) na = m - a; else nm = m + a;
    } else {
      if (a > 0) na = m + a; else nm = m - a;
    }
    mag[j] = nm;
    ang[j] = na;
  }
}

/* ------------------------------------------------------------------- IMDCT -- */
typedef struct {
  int n, ld;
  float *A, *B, *C;
  uint16_t* bitrev;
} mdct_setup;

static mdct_setup g_mdct[16];

/* MdctImpl ctor (Mdct.cs:29-66); twiddles in fp32 like MathF.SinCos */
static const mdct_setup* mdct_get(int n) {
  int ld = vo_ilog(n) - 1;
  if (ld < 5 || ld > 15 || (1 << ld) != n) return NULL;
  mdct_setup* s = &g_mdct[ld];
  if (s->n == n) return s;
  const float pi = 3.14159274f;
  int n2 = n >> 1, n4 = n >> 2, n8 = n >> 3;
  float* A = (float*)malloc(sizeof(float) * (size_t)n2);
  float* B = (float*)malloc(sizeof(float) * (size_t)n2);
  float* C = (float*)malloc(sizeof(float) * (size_t)n4);
  uint16_t* br = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)n8);
  for (int k = 0, k2 = 0; k < n4; ++k, k2 += 2) {
    float a = (float)(4 * k) * pi;
    a = a / (float)n;
    A[k2] = cosf(a);
    A[k2 + 1] = -sinf(a);
    float b = (float)(k2 + 1) * pi;
    b = b / (float)n;
    b = b / 2.0f;
    B[k2] = cosf(b) * .5f;
    B[k2 + 1] = sinf(b) * .5f;
  }
  for (int k = 0, k2 = 0; k < n8; ++k, k2 += 2) {
    float c = (float)(2 * (k2 + 1)) * pi;
    c = c / (float)n;
    C[k2] = cosf(c);
    C[k2 + 1] = -sinf(c);
  }
  for (int i = 0; i < n8; ++i) br[i] = (uint16_t)(vo_bitrev((uint32_t)i, ld - 3) << 2);
  s->A = A;
  s->B = B;
  s->C = C;
  s->bitrev = br;
  s->ld = ld;
  __sync_synchronize();
  s->n = n;
  return s;
}

/* One radix-2 butterfly pair of the step-3 family: (hi, lo) are two complex
 * values stored as (re at [0], im at [-1]); hi += lo, lo = (hi-lo) * (a0 + i*a1).
 * Shared by step3 iter0 / r-loop / s-loop (Mdct.cs:424-474, 556-597, 600-649). */
static inline void bfly(float* e0, float* e2, float a0, float a1) {
  float d0 = e0[0] - e2[0];
  float d1 = e0[-1] - e2[-1];
  e0[0] = e0[0] + e2[0];
  e0[-1] = e0[-1] + e2[-1];
  e2[0] = d0 * a0 - d1 * a1;
  e2[-1] = d1 * a0 + d0 * a1;
}

/* step3_iter0_loop (Mdct.cs:424-470): A advances 8 per butterfly */
static void s3_iter0(int n, float* e, int i_off, int k_off, const float* A) {
  float* e0 = e + i_off;
  float* e2 = e0 + k_off;
  for (int i = n >> 2; i > 0; --i) {
    for (int j = 0; j < 4; j++) {
      bfly(e0 - 2 * j, e2 - 2 * j, A[0], A[1]);
      A += 8;
    }
    e0 -= 8;
    e2 -= 8;
  }
}

/* step3_inner_r_loop (Mdct.cs:472-598, scalar branch): A advances k1 per butterfly */
static void s3_r_loop(int lim, float* e, int d0, int k_off, const float* A, int k1) {
  float* e0 = e + d0;
  float* e2 = e0 + k_off;
  for (int i = lim >> 2; i > 0; --i) {
    for (int j = 0; j < 4; j++) {
      bfly(e0 - 2 * j, e2 - 2 * j, A[0], A[1]);
      A += k1;
    }
    e0 -= 8;
    e2 -= 8;
  }
}

/* step3_inner_s_loop (Mdct.cs:600-649): 4 fixed twiddles, stride k0 */
static void s3_s_loop(int n, float* e, int i_off, int k_off, const float* A, int a_off, int k0) {
  float tw[8];
  for (int j = 0; j < 4; j++) {
    tw[2 * j] = A[a_off * j];
    tw[2 * j + 1] = A[a_off * j + 1];
  }
  float* e0 = e + i_off;
  float* e2 = e0 + k_off;
  for (int i = n; i > 0; --i) {
    for (int j = 0; j < 4; j++) bfly(e0 - 2 * j, e2 - 2 * j, tw[2 * j], tw[2 * j + 1]);
    e0 -= k0;
    e2 -= k0;
  }
}

/* iter_54 (Mdct.cs:696-726) */
static inline void iter_54(float* z) {
  float k00 = z[0] - z[-4];
  float y0 = z[0] + z[-4];
  float y2 = z[-2] + z[-6];
  float k22 = z[-2] - z[-6];
  z[0] = y0 + y2;
  z[-2] = y0 - y2;
  float k33 = z[-3] - z[-7];
  z[-4] = k00 + k33;
  z[-6] = k00 - k33;
  float k11 = z[-1] - z[-5];
  float y1 = z[-1] + z[-5];
  float y3 = z[-3] + z[-7];
  z[-1] = y1 + y3;
  z[-3] = y1 - y3;
  z[-5] = k11 - k22;
  z[-7] = k11 + k22;
}

/* step3_inner_s_loop_ld654 (Mdct.cs:651-694) */
static void s3_ld654(int n, float* e, int i_off, const float* A, int base_n) {
  int a_off = base_n >> 3;
  float A2 = A[a_off];
  float* z = e + i_off;
  float* base = z - 16 * n;
  while (z > base) {
    float k00 = z[0] - z[-8];
    float k11 = z[-1] - z[-9];
    float l00 = z[-2] - z[-10];
    float l11 = z[-3] - z[-11];
    z[0] = z[0] + z[-8];
    z[-1] = z[-1] + z[-9];
    z[-2] = z[-2] + z[-10];
    z[-3] = z[-3] + z[-11];
    z[-8] = k00;
    z[-9] = k11;
    z[-10] = (l00 + l11) * A2;
    z[-11] = (l11 - l00) * A2;

    k00 = z[-4] - z[-12];
    k11 = z[-5] - z[-13];
    l00 = z[-6] - z[-14];
    l11 = z[-7] - z[-15];
    z[-4] = z[-4] + z[-12];
    z[-5] = z[-5] + z[-13];
    z[-6] = z[-6] + z[-14];
    z[-7] = z[-7] + z[-15];
    z[-12] = k11;
    z[-13] = -k00;
    z[-14] = (l11 - l00) * A2;
    z[-15] = (l00 + l11) * -A2;

    iter_54(z);
    iter_54(z - 8);
    z -= 16;
  }
}

/* MdctImpl.CalcReverse (Mdct.cs:77-419) */
static void mdct_reverse(const mdct_setup* s, float* buffer, float* buf2) {
  const int n = s->n, n2 = n >> 1, n4 = n >> 2, n8 = n >> 3, ld = s->ld;
  const float* A = s->A;
  float* u = buffer;
  float* v = buf2;

  /* step 0 (Mdct.cs:98-126): fold + first twiddle into buf2, written back to front */
  {
    int d = n2 - 2, a = 0;
    for (int e = 0; e < n2; e += 4, d -= 2, a += 2) {
      v[d + 1] = buffer[e] * A[a] - buffer[e + 2] * A[a + 1];
      v[d] = buffer[e] * A[a + 1] + buffer[e + 2] * A[a];
    }
    for (int e = n2 - 3; d >= 0; e -= 4, d -= 2, a += 2) {
      v[d + 1] = -buffer[e + 2] * A[a] - -buffer[e] * A[a + 1];
      v[d] = -buffer[e + 2] * A[a + 1] + -buffer[e] * A[a];
    }
  }

  /* step 2 (Mdct.cs:140-178): v -> u */
  {
    int aa = n2 - 8;
    for (int k = 0; aa >= 0; k += 4, aa -= 8) {
      const float* e0 = v + n4 + k;
      const float* e1 = v + k;
      float* d0 = u + n4 + k;
      float* d1 = u + k;
      float v41 = e0[1] - e1[1];
      float v40 = e0[0] - e1[0];
      d0[1] = e0[1] + e1[1];
      d0[0] = e0[0] + e1[0];
      d1[1] = v41 * A[aa + 4] - v40 * A[aa + 5];
      d1[0] = v40 * A[aa + 4] + v41 * A[aa + 5];
      v41 = e0[3] - e1[3];
      v40 = e0[2] - e1[2];
      d0[3] = e0[3] + e1[3];
      d0[2] = e0[2] + e1[2];
      d1[3] = v41 * A[aa] - v40 * A[aa + 1];
      d1[2] = v40 * A[aa] + v41 * A[aa + 1];
    }
  }

  /* step 3 (Mdct.cs:184-247) */
  s3_iter0(n >> 4, u, n2 - 1 - n4 * 0, -(n >> 3), A);
  s3_iter0(n >> 4, u, n2 - 1 - n4 * 1, -(n >> 3), A);
  for (int q = 0; q < 4; q++) s3_r_loop(n >> 5, u, n2 - 1 - n8 * q, -(n >> 4), A, 16);
  int l = 2;
  for (; l < (ld - 3) >> 1; ++l) {
    int k0 = n >> (l + 2), k0_2 = k0 >> 1;
    int lim = 1 << (l + 1);
    for (int i = 0; i < lim; ++i) s3_r_loop(n >> (l + 4), u, n2 - 1 - k0 * i, -k0_2, A, 1 << (l + 3));
  }
  for (; l < ld - 6; ++l) {
    int k0 = n >> (l + 2), k1 = 1 << (l + 3), k0_2 = k0 >> 1;
    int rlim = n >> (l + 6);
    int lim = 1 << (l + 1);
    const float* A0 = A;
    int i_off = n2 - 1;
    for (int r = rlim; r > 0; --r) {
      s3_s_loop(lim, u, i_off, -k0_2, A0, k1, k0);
      A0 += k1 * 4;
      i_off -= 8;
    }
  }
  s3_ld654(n >> 5, u, n2 - 1, A, n);

  /* steps 4,5,6 (Mdct.cs:256-292): bit-reversed gather u -> v */
  {
    const uint16_t* br = s->bitrev;
    for (int d0 = n4 - 4, d1 = n2 - 4; d0 >= 0; d0 -= 4, d1 -= 4, br += 2) {
      int k4 = br[0];
      v[d1 + 3] = u[k4 + 0];
      v[d1 + 2] = u[k4 + 1];
      v[d0 + 3] = u[k4 + 2];
      v[d0 + 2] = u[k4 + 3];
      k4 = br[1];
      v[d1 + 1] = u[k4 + 0];
      v[d1 + 0] = u[k4 + 1];
      v[d0 + 1] = u[k4 + 2];
      v[d0 + 0] = u[k4 + 3];
    }
  }

  /* step 7 (Mdct.cs:302-349): in place on v */
  {
    const float* C = s->C;
    float* d = v;
    float* e = v + n2 - 4;
    while (d < e) {
      float a02 = d[0] - e[2];
      float a11 = d[1] + e[3];
      float b0 = C[1] * a02 + C[0] * a11;
      float b1 = C[1] * a11 - C[0] * a02;
      float b2 = d[0] + e[2];
      float b3 = d[1] - e[3];
      d[0] = b2 + b0;
      d[1] = b3 + b1;
      e[2] = b2 - b0;
      e[3] = b1 - b3;
      a02 = d[2] - e[0];
      a11 = d[3] + e[1];
      b0 = C[3] * a02 + C[2] * a11;
      b1 = C[3] * a11 - C[2] * a02;
      b2 = d[2] + e[0];
      b3 = d[3] - e[1];
      d[2] = b2 + b0;
      d[3] = b3 + b1;
      e[0] = b2 - b0;
      e[1] = b1 - b3;
      C += 4;
      d += 4;
      e -= 4;
    }
  }

  /* step 8 + decode (Mdct.cs:360-414): v -> buffer with the TDAC symmetries */
  {
    const float* B = s->B + n2 - 8;
    const float* e = buf2 + n2 - 8;
    float* d0 = buffer;
    float* d1 = buffer + n2 - 4;
    float* d2 = buffer + n2;
    float* d3 = buffer + n - 4;
    while (e >= v) {
      for (int j = 0; j < 4; j++) {
        int q = 6 - 2 * j;
        float pa = e[q] * B[q + 1] - e[q + 1] * B[q];
        float pb = -e[q] * B[q] - e[q + 1] * B[q + 1];
        d0[j] = pa;
        d1[3 - j] = -pa;
        d2[j] = pb;
        d3[3 - j] = pb;
      }
      B -= 8;
      e -= 8;
      d0 += 4;
      d2 += 4;
      d1 -= 4;
      d3 -= 4;
    }
  }
}

/* the same with caller-provided scratch of n / 2 floats (no allocation per call) */
int vo_imdct2(float* buf, int n, float* buf2) {
  const mdct_setup* s = mdct_get(n);
  if (!s) return VO_E_ARGUMENT;
  mdct_reverse(s, buf, buf2);
  return VO_OK;
}

int vo_imdct(float* buf, int n) {
  const mdct_setup* s = mdct_get(n);
  if (!s) return VO_E_ARGUMENT;
  float* buf2 = (float*)malloc(sizeof(float) * (size_t)(n / 2));
  mdct_reverse(s, buf, buf2);
  free(buf2);
  return VO_OK;
}

/* -------------------------------------------------------------------- mode -- */
void vo_packet_info(int size0, int size1, int block_flag, int prev_flag, int next_flag, int32_t info[6]) {
  int size = block_flag ? size1 : size0;
  int center = size / 2;
  int prev = block_flag ? prev_flag : 1;
  int next = block_flag ? next_flag : 1;
  if (prev) {
    info[2] = 0;
    info[3] = center;
    info[0] = size / 2;
    info[1] = block_flag;
  } else {
    info[2] = (size - size0) / 4;
    info[3] = (size + size0) / 4;
    info[0] = size0 / 2;
    info[1] = 0;
  }
  if (next) {
    info[4] = center;
    info[5] = size;
  } else {
    info[4] = (size * 3 - size0) / 4;
    info[5] = (size * 3 + size0) / 4;
  }
}

/* Mode.GetPacketInfo (Mode.cs:30-66) */
int vo_mode_packet_info(const vo_setup* st, const vo_mode* m, vo_bits* br, vo_pinfo* info) {
  if (br->is_short) {
    memset(info, 0, sizeof(*info));
    return 0;
  }
  int prev = 1, next = 1;
  if (m->block_flag) {
    prev = vo_read_bit(br);
    next = vo_read_bit(br);
  }
  int32_t a[6];
  vo_packet_info(st->size0, st->size1, m->block_flag, prev, next, a);
  info->length = a[0];
  info->left_use_size1 = a[1];
  info->left_start = a[2];
  info->left_end = a[3];
  info->right_start = a[4];
  info->right_end = a[5];
  return 1;
}

/* ----------------------------------------------------------------- mapping -- */
/* Mapping.DecodePacket (Mapping.cs:98-196) */
void vo_mapping_decode(vo_setup* st, const vo_mapping* mp, vo_bits* br, int block_size, float* buf,
                       vo_packet_dump* d) {
  const int half = block_size / 2;
  const int channels = st->channels;
  const int stride = st->size1;
  floor1_data fd[VO_MAX_CH];
  uint8_t no_exec[VO_MAX_CH];

  for (int ch = 0; ch < channels; ch++) {
    const vo_floor1* f = &st->floors[mp->submap_floor[mp->mux[ch]]];
    if (f->type == 0)
      floor0_unpack(f, st->books, br, &fd[ch], d);
    else
      floor1_unpack(f, st->books, br, &fd[ch], d);
    no_exec[ch] = fd[ch].post_count <= 0;
    memset(buf + (size_t)ch * stride, 0, sizeof(float) * (size_t)stride);
    if (d) {
      d->post_count[ch] = fd[ch].post_count;
      memcpy(d->raw_posts[ch], fd[ch].posts, sizeof(int) * 64);
    }
  }
  for (int i = 0; i < mp->coupling_steps; i++) {
    int mag = mp->mag[i], ang = mp->ang[i];
    if (!(no_exec[mag] && no_exec[ang])) {
      no_exec[mag] = 0;
      no_exec[ang] = 0;
    }
  }
  if (d)
    for (int ch = 0; ch < channels; ch++) d->no_execute[ch] = no_exec[ch];

  /* one zeroed scratch for all submaps, never re-zeroed in between (Mapping.cs:133) */
  float* scratch = (float*)calloc((size_t)channels * stride, sizeof(float));
  for (int i = 0; i < mp->submaps; i++) {
    uint8_t flags[VO_MAX_CH];
    int nch = 0;
    for (int j = 0; j < channels; j++)
      if (mp->mux[j] == i) flags[nch++] = no_exec[j];
    vo_residue* r = &st->residues[mp->submap_residue[i]];
    /* ChannelBuffer(decodeBuffer, count, blockSize): stride is the block size */
    if (r->type == 2)
      residue2_decode(r, st->books, br, flags, nch, block_size, scratch, block_size, d);
    else
      residue_decode(r, st->books, br, flags, nch, block_size, scratch, block_size, d);
    int c = 0;
    for (int j = 0; j < channels; j++) {
      if (mp->mux[j] == i) {
        memcpy(buf + (size_t)j * stride, scratch + (size_t)c * block_size, sizeof(float) * (size_t)half);
        c++;
      }
    }
  }
  free(scratch);
  if (d && d->residue)
    for (int ch = 0; ch < channels; ch++)
      memcpy(d->residue + (size_t)ch * half, buf + (size_t)ch * stride, sizeof(float) * (size_t)half);

  for (int i = mp->coupling_steps - 1; i >= 0; i--)
    apply_coupling(buf + (size_t)mp->mag[i] * stride, buf + (size_t)mp->ang[i] * stride, half);

  const mdct_setup* ms = mdct_get(block_size);
  float* buf2 = (float*)malloc(sizeof(float) * (size_t)half);
  for (int ch = 0; ch < channels; ch++) {
    float* span = buf + (size_t)ch * stride;
    if (fd[ch].post_count > 0) {
      const vo_floor1* f = &st->floors[mp->submap_floor[mp->mux[ch]]];
      uint8_t step[64];
      memset(step, 0, sizeof(step));
      if (f->type == 0)
        floor0_apply(f, &fd[ch], block_size == st->size0 ? 0 : 1, block_size, span);
      else
        floor1_apply(f, &fd[ch], block_size, span, step);
      if (d) {
        memcpy(d->final_y[ch], fd[ch].posts, sizeof(int) * 64);
        for (int k = 0; k < 64; k++) d->step_flags[ch][k] = step[k];
        if (d->spectrum) memcpy(d->spectrum + (size_t)ch * half, span, sizeof(float) * (size_t)half);
      }
      mdct_reverse(ms, span, buf2);
    } else {
      memset(span, 0, sizeof(float) * (size_t)half);
      if (d && d->spectrum) memset(d->spectrum + (size_t)ch * half, 0, sizeof(float) * (size_t)half);
    }
    if (d && d->imdct) memcpy(d->imdct + (size_t)ch * block_size, span, sizeof(float) * (size_t)block_size);
  }
  free(buf2);
}

const float* vo_inverse_db_table(void) { return (const float*)(const void*)k_inverse_db_bits; }
