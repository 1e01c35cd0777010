/*
 * vorbis_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the VorbisPizza (NVorbis fork) decode path, used as
 * the parity checker for the B200 kernels and as bench.py's cpu_baseline leg.
 * Nothing under vorbispizza_b200/ may include, link or call this.
 *
 * PARITY PINNING: the reference is C# (net8/net9); no .NET runtime exists in
 * this image, so the reference itself cannot be executed and it ships no golden
 * vectors for this path ("parity unpinned" against the reference binary).  The
 * restatement is pinned instead by (a) self-checks (page CRCs, Kraft sums,
 * float64 direct-form IMDCT, closed-form window/dB table, sample totals) and
 * (b) the reference's own test convention -- <= 2 LSB @ 16 bit against an
 * independent native Vorbis decoder (tests/golden/, generated with FFmpeg's
 * native decoder by tests/golden/make_ffmpeg_pin.py).
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/NVorbis/).
 */
#ifndef VORBIS_ORACLE_H
#define VORBIS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vo_stream vo_stream;

/* error codes (negative) */
enum {
  VO_OK = 0,
  VO_E_INVALID_DATA = -1,      /* InvalidDataException */
  VO_E_ARGUMENT = -2,          /* ArgumentException / ArgumentOutOfRange */
  VO_E_SEEK_RANGE = -3,        /* SeekOutOfRangeException */
  VO_E_PREROLL = -4,           /* PreRollPacketException */
  VO_E_UNSUPPORTED = -5,       /* floor0 etc: not restated */
  VO_E_REF_FAULT = -6,         /* the reference would throw an unintended exception here */
  VO_E_NOMEM = -7
};

/* ---- container + stream ------------------------------------------------- */

/* Opens the first logical Vorbis stream found in an in-memory Ogg file
 * (VorbisReader.Initialize, VorbisReader.cs:56).  The data must stay alive. */
vo_stream* vo_open(const uint8_t* data, size_t len, int* err);
void vo_close(vo_stream* s);

int vo_channels(const vo_stream* s);
int vo_sample_rate(const vo_stream* s);
int vo_block_size(const vo_stream* s, int which);
int vo_bitrate(const vo_stream* s, int which); /* 0 upper, 1 nominal, 2 lower */
int64_t vo_container_bits(const vo_stream* s);
int64_t vo_waste_bits(const vo_stream* s);
int vo_page_count(const vo_stream* s);
int vo_crc_failures(const vo_stream* s);

/* comment header (StreamDecoder.LoadComments, StreamDecoder.cs:242) */
const char* vo_vendor(const vo_stream* s, int* len);
int vo_comment_count(const vo_stream* s);
const char* vo_comment(const vo_stream* s, int i, int* len);

/* raw header packets (id=0, comment=1, setup=2) as assembled by the Ogg layer */
const uint8_t* vo_header_packet(const vo_stream* s, int which, int* len);

/* IStreamDecoder surface (StreamDecoder.cs:407-498, 817-880, 933-1007) */
void vo_set_clip(vo_stream* s, int clip);
int vo_has_clipped(const vo_stream* s);
int vo_is_end_of_stream(const vo_stream* s);
int64_t vo_sample_position(const vo_stream* s);
int64_t vo_total_samples(vo_stream* s);
/* the page-end granule cache of PacketProvider (Ogg/PacketProvider.cs:203-307), filled to the end of the stream */
int vo_page_end_granules(vo_stream* s, int64_t* out, int cap);
/* interleaved read; nfloats must be a multiple of channels; returns samples per
 * channel (>=0) or a negative error */
int vo_read(vo_stream* s, float* buf, int nfloats);
/* planar read: buf[ch*channel_stride + i] */
int vo_read_planar(vo_stream* s, float* buf, int nfloats, int samples_to_read, int channel_stride);
int vo_seek(vo_stream* s, int64_t sample_position);

/* ---- audio packet table (what IPacketProvider.GetNextPacket would hand out
 *      when walking forward from the first audio packet) -------------------- */
typedef struct {
  const uint8_t* data;   /* assembled, contiguous; >= 16 readable zero bytes follow */
  int32_t len;
  int32_t is_resync;
  int32_t is_eos;
  int64_t granule;       /* -1 when not the last packet completed on its page */
  int32_t page_index;    /* page the packet starts on */
  int32_t packet_index;  /* index inside that page */
} vo_packet_view;

int vo_audio_packet_count(vo_stream* s);
int vo_audio_packet(vo_stream* s, int i, vo_packet_view* out);

/* ---- setup introspection (for table cross-checks) ------------------------ */
int vo_book_count(const vo_stream* s);
int vo_book_info(const vo_stream* s, int b, int* dims, int* entries, int* max_bits, int* map_type,
                 int* prefix_bits, int* overflow_count);
/* fills lengths[entries] (-1 = unused) */
int vo_book_lengths(const vo_stream* s, int b, int* lengths);
const float* vo_book_lookup(const vo_stream* s, int b, int* count);
/* decode one scalar from a standalone bit buffer with book b; returns symbol or -1,
 * *bitpos is advanced exactly like VorbisPacket would (Codebook.cs:301-335) */
int vo_book_decode(const vo_stream* s, int b, const uint8_t* data, int len_bytes, int64_t* bitpos,
                   int* is_short);
double vo_book_kraft(const vo_stream* s, int b);

/* ---- single-packet decode with stage dumps ------------------------------- */
#define VO_MAX_CH 8
typedef struct {
  /* outputs */
  int32_t status;        /* 0 decoded; 1 not an audio packet / rejected */
  int32_t mode;
  int32_t block_flag;
  int32_t block_size;
  int32_t info[6];       /* Length, LeftUseSize1, LeftStart, LeftEnd, RightStart, RightEnd (PacketInfo.cs) */
  int32_t bits_read;
  int32_t is_short;      /* VorbisPacket.IsShort at the end of decode */
  /* every Codebook.DecodeScalar result in call order */
  int32_t* scalars; int32_t scalars_cap; int32_t scalars_n;
  /* floor1 per channel */
  int32_t post_count[VO_MAX_CH];
  int32_t raw_posts[VO_MAX_CH][64];
  int32_t final_y[VO_MAX_CH][64];
  int32_t step_flags[VO_MAX_CH][64];
  int32_t no_execute[VO_MAX_CH];      /* after coupling propagation (Mapping.cs:121-130) */
  /* residue partition classes in decode order: (submap, channel-in-submap, partition) */
  int32_t* classes; int32_t classes_cap; int32_t classes_n;
  /* per channel float arrays, each block_size/2 long (caller supplies storage or NULL) */
  float* residue;        /* [ch][n/2]  after residue decode, before coupling */
  float* spectrum;       /* [ch][n/2]  after coupling + floor multiply (IMDCT input) */
  float* imdct;          /* [ch][n]    IMDCT output (zeros for non-executed channels) */
} vo_packet_dump;

/* Decodes audio packet i standalone (Mode.Decode -> Mapping.DecodePacket,
 * Mode.cs:68-85, Mapping.cs:98-196) filling the dump. */
int vo_decode_packet_dump(vo_stream* s, const uint8_t* data, int len, vo_packet_dump* d);

/* ---- building blocks exposed for unit tests ------------------------------ */
/* Mdct.Reverse (Mdct.cs:15-19,77-419): in-place, buf has n floats, first n/2 are input */
int vo_imdct(float* buf, int n);
/* BlocksizeDerivedCache.CalcWindowSlope (BlocksizeDerivedCache.cs:24-35): slope[n] */
void vo_window_slope(float* slope, int n);
const float* vo_inverse_db_table(void);
uint32_t vo_crc_ogg(const uint8_t* data, size_t len, uint32_t crc);
/* Mode.GetPacketInfo given the sizes and flags (Mode.cs:30-66) */
void vo_packet_info(int size0, int size1, int block_flag, int prev_flag, int next_flag, int32_t info[6]);

/* ---- host-core baseline (vo_bench.c) ---------------------------------------------------- */
/* Decodes njobs whole streams (job j = file j % nfiles) TestApp-style on nthreads threads; returns
 * channel-samples decoded, *seconds = wall clock. */
int64_t vo_bench_decode(const uint8_t* const* datas, const size_t* lens, int nfiles, int njobs, int nthreads,
                        double* seconds);
/* n excerpts: SeekTo(start[i]) + read count[i] samples per channel of file file_of[i]; every thread keeps one
 * open reader per file.  Returns the channel-samples delivered. */
/* BASELINE config 3 on host cores: synthetic streams of `n_blocks` blocks each (flags bit 0: long block; window
 * flags of a long block follow its neighbours like Mode.GetPacketInfo would read them), `channels` spectra of n / 2
 * floats per block back to back in `spectra`.  Per block Mdct.Reverse (Mdct.cs:77-419), OverlapBuffers
 * (StreamDecoder.cs:764-791) and the interleaved clipped store of StoreInterleaved (StreamDecoder.cs:515-592) into a
 * per-thread buffer; streams are handed to `nthreads` threads by a shared cursor.  Returns the channel-samples
 * produced; *seconds = wall time; *checksum (may be NULL) = sum of all samples, so the work cannot be elided. */
int64_t vo_bench_imdct_ola(const float* spectra, const uint8_t* flags, int n_streams, int n_blocks, int channels,
                           int size0, int size1, int nthreads, double* seconds, double* checksum);
int vo_imdct2(float* buf, int n, float* scratch);

int64_t vo_bench_excerpts(const uint8_t* const* datas, const size_t* lens, int nfiles, int n, const uint32_t* file_of,
                          const int64_t* start, const int32_t* count, int nthreads, double* seconds);

#ifdef __cplusplus
}
#endif
#endif
