/*
 * vpz.h -- C ABI of the B200-native Vorbis decode path (libvpz.so).
 *
 * This is the boundary a VorbisPizza maintainer binds with P/Invoke to replace the managed hot
 * path below StreamDecoder.Read (reference: NVorbis/StreamDecoder.cs:418) and above
 * IPacketProvider.GetNextPacket (NVorbis/Contracts/IPacketProvider.cs:9-48).  Plain pointers and
 * sizes only; no exceptions, callbacks or caller-freed allocations cross it.  Every entry point
 * returns 0 (VPZ_OK) / a non-negative count, or a negative VPZ_E_* code that the managed shim maps
 * to the reference's exception types.  There is NO CPU fallback: without a
 * usable sm_100 device every compute entry point fails with VPZ_E_NO_DEVICE / VPZ_E_CUDA.
 *
 * Three layers, lowest first:
 *   1. setup  -- vpz_setup_*  : replaces StreamDecoder.LoadStreamHeader/LoadBooks table building
 *                               (StreamDecoder.cs:213-355); tables are uploaded once per distinct
 *                               (id, setup) header pair.
 *   2. batch  -- vpz_batch_*  : replaces Mode.Decode -> Mapping.DecodePacket -> Mdct.Reverse ->
 *                               OverlapBuffers -> StoreInterleaved (Mode.cs:68, Mapping.cs:98,
 *                               Mdct.cs:15, StreamDecoder.cs:515-592,764-791) for MANY packets of
 *                               MANY streams per call.  This is the IPacketProvider-side seam.
 *   3. reader -- vpz_reader_* : the IStreamDecoder / IVorbisReader-side seam
 *                               (Contracts/IStreamDecoder.cs:9-152, Contracts/IVorbisReader.cs:10-150)
 *                               including the host-side Ogg layer, for callers that hand over the
 *                               container bytes instead of packets.
 * All calls that share one vpz_ctx must come from one host thread at a time.
 */
#ifndef VPZ_H
#define VPZ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vpz_ctx vpz_ctx;
typedef struct vpz_setup vpz_setup;
typedef struct vpz_batch vpz_batch;
typedef struct vpz_reader vpz_reader;

enum {
  VPZ_OK = 0,
  VPZ_E_INVALID_DATA = -1,  /* System.IO.InvalidDataException (StreamDecoder.cs:77,84,313,330,338,351,734) */
  VPZ_E_ARGUMENT = -2,      /* ArgumentException / ArgumentOutOfRangeException (StreamDecoder.cs:423-430,825,842) */
  VPZ_E_SEEK_RANGE = -3,    /* SeekOutOfRangeException (StreamDecoder.cs:861) */
  VPZ_E_PREROLL = -4,       /* PreRollPacketException (StreamDecoder.cs:874) */
  VPZ_E_UNSUPPORTED = -5,   /* valid Vorbis the GPU path does not cover (> 8 channels, block size < 256, > 64 floor posts) */
  VPZ_E_CUDA = -6,          /* CUDA runtime error; vpz_last_error() has the text */
  VPZ_E_NOMEM = -7,
  VPZ_E_DISPOSED = -8,      /* ObjectDisposedException (StreamDecoder.cs:401) */
  VPZ_E_INVALID_OP = -9,    /* InvalidOperationException (StreamDecoder.cs:822) */
  VPZ_E_NO_DEVICE = -10,    /* no sm_100 GPU visible: the product has no CPU path */
  VPZ_E_REF_FAULT = -11     /* the reference faults here; we stop the stream instead */
};

const char* vpz_strerror(int code);
/* Text of the last failure on this context (valid until the next call on it). */
const char* vpz_last_error(const vpz_ctx* ctx);
/* Library version / build string, e.g. "vpz 0.1 sm_100a". */
const char* vpz_version(void);

/* ---- context: one per GPU ------------------------------------------------------------- */
/* device < 0: the current CUDA device.  Creates the stream, pinned staging and table caches. */
int vpz_ctx_create(int device, vpz_ctx** out);
void vpz_ctx_destroy(vpz_ctx* ctx);
int vpz_device_count(void);
/* Tunables (call before the first batch): key one of "l1_bits" (Huffman first-level table width,
 * default 9), "ola_chunk" (packets per IMDCT work item, default: 16..63 chosen per batch), "k1_warps" (warps per entropy
 * CTA, default 4), "bulk_group" (streams per pipeline group of vpz_decode_files, default 256), "bulk_group_mib" (a group also closes
 * at this many MiB of container images, default 128, at most 384),
 * "host_threads" (size of the host worker pool of the bulk calls, default 0 = all cores up to 32), "bulk_threads"
 * (how many of them vpz_decode_files uses, default 4: the call is bound by the PCM copy, which more staging
 * threads slow down; 0 = all), "gpu_scan" (bulk
 * path: Ogg page scan + CRC on the GPU, default 1), "force_general" (tests: route every packet through the
 * general kernels). */
int vpz_ctx_set(vpz_ctx* ctx, const char* key, int value);

/* Device-side stopwatch on the context's stream: vpz_ctx_mark records CUDA event `slot` (0..7) on
 * the stream all of this context's kernels and copies are issued on; vpz_ctx_elapsed_ms waits for
 * slot b and returns the milliseconds between two marks (< 0 on error). */
int vpz_ctx_mark(vpz_ctx* ctx, int slot);
float vpz_ctx_elapsed_ms(vpz_ctx* ctx, int slot_a, int slot_b);
/* Number of kernels this library has launched on the context so far. */
int64_t vpz_ctx_kernel_launches(const vpz_ctx* ctx);

/* ---- setup: StreamDecoder.LoadStreamHeader + LoadBooks -------------------------------- */
typedef struct {
  int32_t channels, sample_rate;
  int32_t bitrate_upper, bitrate_nominal, bitrate_lower;
  int32_t block_size0, block_size1;
  int32_t n_books, n_floors, n_residues, n_mappings, n_modes;
  int32_t max_codeword_bits;
  uint64_t table_bytes;      /* size of the device image */
} vpz_setup_info;

/* Parses the identification and setup header packets, builds the GPU tables and uploads them.
 * Identical header pairs share one vpz_setup (content hash); each create needs one release. */
int vpz_setup_create(vpz_ctx* ctx, const uint8_t* id_pkt, size_t id_len, const uint8_t* setup_pkt,
                     size_t setup_len, vpz_setup** out);
void vpz_setup_release(vpz_setup* s);
int vpz_setup_get_info(const vpz_setup* s, vpz_setup_info* info);

/* Mode.GetPacketInfo (Mode.cs:30-66) for one audio packet, from its first bytes: the callback the
 * reference's seek code needs (IPacketGranuleCountProvider, StreamDecoder.cs:882-913).
 * info = {Length, LeftUseSize1, LeftStart, LeftEnd, RightStart, RightEnd}.  Returns 1 when the
 * packet is a decodable audio packet, 0 when the reference would skip it, VPZ_E_INVALID_DATA for
 * an unused mode index (StreamDecoder.cs:732-735). */
int vpz_packet_info(const vpz_setup* s, const uint8_t* pkt, size_t len, int32_t info[6]);

/* ---- batch: many packets of many streams in one pass ---------------------------------- */
int vpz_batch_create(vpz_ctx* ctx, vpz_batch** out);
void vpz_batch_destroy(vpz_batch* b);
/* Forget all runs but keep the buffers (steady-state reuse). */
int vpz_batch_reset(vpz_batch* b);

/* Adds one RUN: a fresh decoder state (as after StreamDecoder.ResetDecoder, StreamDecoder.cs:357)
 * fed `n_pkts` consecutive audio packets.  As in the reference the first decodable packet of a run
 * only seeds the overlap and yields no samples, so a caller continuing a stream passes its previous
 * packet again as the first packet of the next run (this is also the seek pre-roll).
 *   bytes/offsets : packet i is bytes[offsets[i] .. offsets[i+1])
 *   trim          : NULL, or per packet the number of samples to pull RightStart back by
 *                   (end-of-stream granule trim, StreamDecoder.cs:658-666); a negative value
 *                   instead extends the packet's output by that many samples of its raw right
 *                   half (the drain of StreamDecoder.cs:451-455)
 * Returns the run index (>= 0).  Packets the reference would skip (header bit set, empty decode)
 * are skipped; an unused mode index fails the call with VPZ_E_INVALID_DATA. */
int vpz_batch_add_run(vpz_batch* b, vpz_setup* s, const uint8_t* bytes, const uint32_t* offsets,
                      uint32_t n_pkts, const int32_t* trim);
/* Samples per channel run `run` will produce (known before decoding: geometry is in the packet
 * headers), and its channel count. */
int64_t vpz_batch_run_samples(const vpz_batch* b, int run);
int vpz_batch_run_channels(const vpz_batch* b, int run);
/* 0, or VPZ_E_REF_FAULT when the run was cut at submitted packet *stop_packet because the reference
 * itself faults there (overlap longer than the window slope). */
int vpz_batch_run_status(const vpz_batch* b, int run, int32_t* stop_packet);
/* Per-packet sample counts of a run (PacketInfo.SampleCount after trim; 0 for skipped packets and
 * for the seeding packet).  counts has n_pkts entries. */
int vpz_batch_run_packet_samples(const vpz_batch* b, int run, int32_t* counts);
int64_t vpz_batch_total_floats(const vpz_batch* b);
int64_t vpz_batch_total_packets(const vpz_batch* b);
int64_t vpz_batch_total_bytes(const vpz_batch* b);

/* Uploads the queued packets (host -> device) and builds the device work lists. */
int vpz_batch_upload(vpz_batch* b);
/* Launches symbol decode (K1a), spectrum build (K1b: VQ accumulate, coupling, floor) and IMDCT +
 * window + overlap-add + store (K3) on the context's stream.  clip != 0 clamps to +-0.99999994f like Utils.ClipValue (Utils.cs:44-58).
 * Asynchronous; vpz_batch_sync waits.  May be called repeatedly on the same uploaded batch. */
int vpz_batch_decode(vpz_batch* b, int clip);
int vpz_batch_sync(vpz_batch* b);
/* 1 when any sample was clamped in the last decode (HasClipped, StreamDecoder.cs:569-570). */
int vpz_batch_has_clipped(vpz_batch* b);

/* Interleaved float PCM of one run, device -> host.  dst has run_samples * channels floats. */
int vpz_batch_read_run(vpz_batch* b, int run, float* dst);
/* All runs back to back (run order), device -> host; dst has vpz_batch_total_floats floats.
 * dst may be pinned memory from vpz_host_alloc for full PCIe rate. */
int vpz_batch_read_all(vpz_batch* b, float* dst);
/* Float offset of a run inside the batch PCM buffer / the device pointer of that buffer, for
 * consumers that keep PCM on the GPU. */
int64_t vpz_batch_run_offset(const vpz_batch* b, int run);
const float* vpz_batch_device_pcm(const vpz_batch* b);

/* Device time of the last vpz_batch_decode in milliseconds: which = 0 total, 1 entropy stage (K1a +
 * K1b), 11 symbol decode (K1a), 12 spectrum build (K1b), 3 IMDCT/OLA (K3); number of kernel launches
 * in `launches` (may be NULL). */
float vpz_batch_last_ms(vpz_batch* b, int which, int* launches);

/* Bytes this process has copied so far: which = 0 host->device, 1 device->host. */
uint64_t vpz_transfer_bytes(int which);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost). */
void* vpz_host_alloc(size_t bytes);
void vpz_host_free(void* p);

/* ---- stage dumps for parity tests (integer stages must be bit-exact) -------- */
#define VPZ_DUMP_MAX_CH 8
typedef struct {
  int32_t status;
  int32_t mode, block_size;
  int32_t info[6];
  int32_t bits_read;
  int32_t exec_mask;
  int32_t no_execute_mask;
  int32_t scalars_n;
  int32_t classes_n;
  int32_t post_count[VPZ_DUMP_MAX_CH];
  int32_t raw_posts[VPZ_DUMP_MAX_CH][64];
  int32_t final_y[VPZ_DUMP_MAX_CH][64];
  int32_t step_flags[VPZ_DUMP_MAX_CH][64];
} vpz_packet_dump;

/* Decodes ONE audio packet on the GPU with every stage written out.  scalars/classes receive up to
 * *_cap entries (the _n fields of the dump hold the true counts); residue (before coupling) and
 * spectrum (IMDCT input) receive channels*block_size/2 floats, imdct channels*block_size floats
 * (raw, unwindowed transform output); any of the five may be NULL. */
int vpz_debug_decode_packet(vpz_ctx* ctx, vpz_setup* s, const uint8_t* pkt, size_t len,
                            vpz_packet_dump* dump, int32_t* scalars, int32_t scalars_cap,
                            int32_t* classes, int32_t classes_cap, float* residue, float* spectrum,
                            float* imdct);

/* Kernel-only IMDCT + window + overlap-add + interleaved store on caller-provided spectra
 * (BASELINE config 3).  n_streams runs of n_blocks blocks; flags[s*n_blocks+i] bit0 = long block;
 * window flags follow the Vorbis rule (prev/next flag = neighbour is long).  spectra holds, run
 * after run, block after block, channel after channel, block_size/2 floats.  The spectra stay on
 * the device; returns a handle used like a batch whose entropy stage is skipped. */
int vpz_synth_create(vpz_ctx* ctx, int channels, int log2_size0, int log2_size1, uint32_t n_streams,
                     uint32_t n_blocks, const uint8_t* flags, const float* spectra, vpz_batch** out);

/* ---- reader: IVorbisReader / IStreamDecoder surface ------------------------------------- */
/* VorbisReader(Stream) + Initialize() (VorbisReader.cs:37-66) over container bytes in memory.
 * copy != 0: the library keeps its own copy.  Fails with VPZ_E_INVALID_DATA when no Vorbis stream
 * is found (VorbisReader.cs:63). */
int vpz_reader_open_memory(vpz_ctx* ctx, const uint8_t* data, size_t len, int copy, vpz_reader** out);
void vpz_reader_close(vpz_reader* r);

int vpz_reader_stream_count(const vpz_reader* r);        /* Streams.Count */
int vpz_reader_stream_index(const vpz_reader* r);        /* StreamIndex */
int vpz_reader_switch_stream(vpz_reader* r, int index);  /* SwitchStreams (VorbisReader.cs:191-210) */
int vpz_reader_find_next_stream(vpz_reader* r);          /* FindNextStream: 1 found, 0 none */
int vpz_reader_can_seek(const vpz_reader* r);

int vpz_reader_channels(const vpz_reader* r);
int vpz_reader_sample_rate(const vpz_reader* r);
int vpz_reader_bitrate(const vpz_reader* r, int which);  /* 0 upper, 1 nominal, 2 lower */
int vpz_reader_stream_serial(const vpz_reader* r);
int64_t vpz_reader_total_samples(vpz_reader* r);
int64_t vpz_reader_sample_position(const vpz_reader* r);
int vpz_reader_is_end_of_stream(const vpz_reader* r);
int vpz_reader_has_clipped(const vpz_reader* r);
int vpz_reader_get_clip(const vpz_reader* r);
void vpz_reader_set_clip(vpz_reader* r, int clip);       /* ClipSamples, default 1 */
int64_t vpz_reader_container_overhead_bits(const vpz_reader* r);
int64_t vpz_reader_container_waste_bits(const vpz_reader* r);

/* Tags (TagData.cs): vendor string and raw "KEY=value" comments, UTF-8, not NUL terminated. */
const char* vpz_reader_vendor(const vpz_reader* r, int* len);
int vpz_reader_comment_count(const vpz_reader* r);
const char* vpz_reader_comment(const vpz_reader* r, int i, int* len);

/* IStreamDecoder.Read(Span<float>) (StreamDecoder.cs:407): interleaved; nfloats must be a multiple
 * of Channels; returns samples per channel, at most one packet's worth per call, 0 at the end. */
int vpz_reader_read(vpz_reader* r, float* buf, int nfloats);
/* IStreamDecoder.Read(Span<float>, samplesToRead, channelStride) (StreamDecoder.cs:413): planar. */
int vpz_reader_read_planar(vpz_reader* r, float* buf, int nfloats, int samples_to_read, int channel_stride);
/* SeekTo(long, SeekOrigin) (StreamDecoder.cs:817-880); origin 0 Begin, 1 Current, 2 End (Current and
 * End SUBTRACT the offset like the reference, StreamDecoder.cs:833-839). */
int vpz_reader_seek(vpz_reader* r, int64_t sample_position, int origin);
/* How many packets the reader decodes ahead per GPU batch (default 256; 0 = whole stream). */
int vpz_reader_set_lookahead(vpz_reader* r, int packets);

/* Packet access for callers that keep their own Ogg layer but want ours for tests/benchmarks:
 * audio packet i of the current stream as the packet provider hands it out. */
int vpz_reader_audio_packet_count(vpz_reader* r);
int vpz_reader_audio_packet(vpz_reader* r, int i, const uint8_t** data, uint32_t* len, int64_t* granule,
                            int32_t* flags /* bit0 resync, bit1 end of stream */);
const uint8_t* vpz_reader_header_packet(vpz_reader* r, int which, uint32_t* len); /* 0 id, 1 comment, 2 setup */
vpz_setup* vpz_reader_setup(vpz_reader* r);

/* ---- bulk: many whole files, one call (BASELINE config 4 / bench e2e) --------------------- */
/* Parses n container images on the host, queues every stream's packets (replica k of a file may
 * start at a different packet with start_packet[k], wrapping is the caller's business), decodes
 * them in one batch and leaves interleaved PCM in dst (host) when dst != NULL.  sample_counts[k]
 * receives samples per channel of file k.  Returns total floats written or a negative error. */
int64_t vpz_decode_files(vpz_ctx* ctx, uint32_t n, const uint8_t* const* datas, const size_t* lens,
                         int clip, float* dst, size_t dst_floats, int64_t* sample_counts);

/* The same with 16-bit output: every sample is converted on the GPU by the rule the reference's own
 * tests apply to the float output, v = (int)(x * 32768f) clamped to [-32768, 32767] (AssetTest.cs:131-132),
 * after ClipSamples when clip != 0.  Halves the device-to-host bytes.  Block sizes 256 / 2048 and mono /
 * stereo only (VPZ_E_UNSUPPORTED otherwise).  dst_samples / return value: int16 elements. */
int64_t vpz_decode_files_s16(vpz_ctx* ctx, uint32_t n, const uint8_t* const* datas, const size_t* lens,
                             int clip, int16_t* dst, size_t dst_samples, int64_t* sample_counts);

/* ---- physical Ogg layer on the GPU ----------------------------------------------- */
/* One valid Ogg page as PageReaderBase.ReadNextPage / VerifyPage accept it (Ogg/PageReaderBase.cs:41-84,
 * 286-361): found by capture-pattern search, segment table and body inside the data, CRC-32 (Ogg/Crc.cs:20-63)
 * verified, packets counted (Ogg/PageHeader.cs:35-59). */
typedef struct {
  uint32_t offset;          /* of the page in its container image */
  uint32_t body_len;
  uint32_t granule_lo, granule_hi;
  uint32_t serial, sequence;
  uint8_t flags;            /* 1 continuation, 2 BOS, 4 EOS (Contracts/Ogg/PageFlags.cs) */
  uint8_t segments;
  uint8_t is_resync;        /* bytes were skipped in front of this page */
  uint8_t is_continued;     /* the last lacing value is 255 */
  uint16_t packet_count;
  uint16_t reserved;
} vpz_page_info;

/* Scans n container images on the GPU, one warp per image: sync search, header parse, lacing sums, page CRC.
 * The pages of image i are pages[first[i] .. first[i] + count[i]); waste_bits[i] = bits that belong to no
 * valid page (IVorbisReader.ContainerWasteBits before per-stream filtering), crc_failures[i] = candidates whose
 * CRC did not match.  Any output pointer may be NULL.  Returns the total number of pages or a negative error.
 * vpz_decode_files uses the same scan ("gpu_scan" tunable, default 1). */
int64_t vpz_scan_pages(vpz_ctx* ctx, uint32_t n, const uint8_t* const* datas, const size_t* lens,
                       vpz_page_info* pages, size_t pages_cap, uint32_t* first, uint32_t* count,
                       uint64_t* waste_bits, uint32_t* crc_failures);

/* ---- bulk random access: many short excerpts, one call (BASELINE config 5) -------------------- */
/* Excerpt i is what a fresh VorbisReader over container image file_of[i] delivers for
 *     reader.SeekTo(start[i]);  then ReadSamples until count[i] samples per channel have been read
 * (StreamDecoder.SeekTo, StreamDecoder.cs:817-880: one packet of pre-roll, roll forward inside the
 * target packet; ReadSamples: StreamDecoder.cs:418-498).  The host runs the provider side of every
 * SeekTo, the windows of a group of excerpts are decoded in one GPU batch, interleaved PCM goes to
 * dst (host): excerpt i starts at float offset dst_offsets[i] = sum over j < i of count[j] * channels
 * (fixed layout; returned in dst_offsets when non-NULL).  got[i] receives the samples per channel that
 * were delivered (fewer than count[i] at the end of the stream) or the negative vpz_status SeekTo
 * raised (VPZ_E_SEEK_RANGE, VPZ_E_PREROLL, ...); the floats of an excerpt that were not delivered are set
 * to zero.  clip: ClipSamples.  dst == NULL: only the layout is computed.  The samples are gathered on the
 * device (K4) and arrive in dst by one copy per ~2,048 excerpts: pinned dst (vpz_host_alloc) is fastest.
 * Returns the total floats of the layout or a negative error. */
int64_t vpz_decode_excerpts(vpz_ctx* ctx, uint32_t n_files, const uint8_t* const* datas, const size_t* lens,
                            uint32_t n, const uint32_t* file_of, const int64_t* start, const int32_t* count,
                            int clip, float* dst, size_t dst_floats, int64_t* dst_offsets, int32_t* got);

/* Debug / tests: the page-end granule index SeekTo searches in (PacketProvider.FillPageEndGranuleCache,
 * Ogg/PacketProvider.cs:203-307) of one container image's first logical stream: entry p = granules up to and
 * including page p.  on_device 1: built by the GPU (page scan + one warp per file over its pages), VPZ_E_UNSUPPORTED
 * when the file is not a clean single stream (those keep the host's packet walk); 0: the host's walk.  Returns the
 * number of pages and writes min(pages, cap) entries. */
int64_t vpz_debug_page_end_granules(vpz_ctx* ctx, const uint8_t* data, size_t len, int on_device, int64_t* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* VPZ_H */
